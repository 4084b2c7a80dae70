"""Latency of small batches at n=4096, d=10 (B = 1, 2, 4, 8): wall time per call and the per-family split."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from madaiemulator_b200 import engine, datasets as ds

n, d = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 10)
X = ds.synthetic_design(n, d); y = ds.synthetic_response(X)
ctx = engine.Context(0)
for B in (1, 2, 4, 8):
    m = engine.Model(ctx, X, y, 1, 0, max_slots=B)
    th = np.tile(ds.default_theta_less_amp(d), (B, 1))
    for _ in range(3):
        m.loglik_grad_batch(th)
    t0 = time.time()
    for _ in range(10):
        m.loglik_grad_batch(th)
    ms = (time.time() - t0) * 100.0
    for _ in range(3):
        m.loglik_grad_batch(th, want_grad=False)
    t0 = time.time()
    for _ in range(10):
        m.loglik_grad_batch(th, want_grad=False)
    ms_val = (time.time() - t0) * 100.0
    ctx.profile(True)
    m.loglik_grad_batch(th)
    prof = ctx.profile_read()
    ctx.profile(False)
    print("B=%d: %.3f ms per call (%.1f evals/s), value only %.3f ms (%.1f evals/s); profile-mode split:" % (B, ms, B / ms * 1e3, ms_val, B / ms_val * 1e3),
          ", ".join("%s %.3f ms/%d" % (k, v["ms"], v["launches"]) for k, v in prof.items() if v["launches"]))
    m.close()
