"""Stage-by-stage numeric check of the CUDA engine against numpy / the oracle (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from madaiemulator_b200 import engine, datasets as ds
from oracle.pyoracle import PortOracle

ctx = engine.Context(0)
for (n, d, order) in [(100, 3, 1), (300, 6, 2), (700, 10, 0)]:
    X = ds.synthetic_design(n, d); y = ds.synthetic_response(X)
    o = PortOracle(X, y, 1, order)
    m = engine.Model(ctx, X, y, 1, order, max_slots=2)
    th = ds.default_theta_less_amp(d)
    full = np.concatenate([[0.3], th])
    C = m.cov_matrix(full); Cr = o.cov_matrix(full)
    print(n, "cov relerr", np.max(np.abs(C - Cr) / np.abs(Cr).clip(1e-300)))
    rc, L, ld = m.debug_cholesky(th)
    C1 = o.cov_matrix(np.concatenate([[0.0], th]))
    Lr = np.linalg.cholesky(C1)
    print(n, "chol rc", rc, "L err", np.max(np.abs(L - Lr)), "logdet", ld, 2 * np.log(np.diag(Lr)).sum())
    r = m.loglik_grad(th)
    ref = o.loglik_grad(th)
    W = np.tril(m.debug_fetch(0, 1)); Wr = np.linalg.inv(Lr)
    print(n, "W err", np.max(np.abs(W - Wr)) / np.max(np.abs(Wr)))
    Ci = np.tril(m.debug_fetch(0, 0)); Cir = np.tril(np.linalg.inv(C1))
    print(n, "Cinv err", np.max(np.abs(Ci - Cir)) / np.max(np.abs(Cir)))
    print(n, "negL", r["negL"], ref["negL"], "sigma2", r["sigma2"], ref["sigma2"], "status", r["status"])
    print(n, "beta", r["beta"][:4], ref["beta"][:4])
    print(n, "grad", r["grad"][:4], ref["grad"][:4])
    pts = ds.synthetic_queries(10, d); pts[0] = X[3]
    e = m.emulator(full)
    mu, var = e.emulate(pts)
    mr, vr = o.emulator(full).emulate(pts)
    print(n, "mean err", np.max(np.abs(mu - mr)), "var err", np.max(np.abs(var - vr)))
    e.close(); m.close()

# timing at n=4096
n, d = 4096, 10
X = ds.synthetic_design(n, d); y = ds.synthetic_response(X)
m = engine.Model(ctx, X, y, 1, 0, max_slots=16)
th = np.tile(ds.default_theta_less_amp(d), (16, 1))
for g in (1, 2, 4):
    ctx.set_groups(g)
    for B in (1, 8, 16):
        m.loglik_grad_batch(th[:B])
        t0 = time.time(); r = m.loglik_grad_batch(th[:B]); t1 = time.time()
        print("n=4096 groups", g, "B", B, "ms", (t1 - t0) * 1e3, "evals/s", B / (t1 - t0), "negL", r["negL"][0], "status", r["status"][0])
ctx.set_groups(1)
ctx.profile(True)
m.loglik_grad_batch(th[:4])
pr = ctx.profile_read()
tot = sum(v["ms"] for v in pr.values())
for k, v in pr.items():
    if v["launches"]:
        rate = v["work"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0
        print("%-12s %6d launches %9.3f ms (%5.1f%%)  work/time = %8.3f T/s" % (k, v["launches"], v["ms"], 100 * v["ms"] / tot, rate))
ctx.profile(False)
full = np.concatenate([[0.0], th[0]])
e = m.emulator(full)
pts = ds.synthetic_queries(65536, d)
e.emulate(pts[:1000])
t0 = time.time(); mu, var = e.emulate(pts); t1 = time.time()
print("predict 65536 pts: ms", (t1 - t0) * 1e3, "pts/s", 65536 / (t1 - t0))
ctx.profile(True)
e.emulate(pts[:16384])
pr = ctx.profile_read()
for k, v in pr.items():
    if v["launches"]:
        rate = v["work"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0
        print("%-12s %6d launches %9.3f ms  work/time = %8.3f T/s" % (k, v["launches"], v["ms"], rate))
import time as _t
t0 = _t.time(); e2 = m.emulator(full); t1 = _t.time()
print("emulator set-up (covariance + factor + inverse + regression) at n=4096: %.1f ms" % ((t1 - t0) * 1e3))
e2.close()
