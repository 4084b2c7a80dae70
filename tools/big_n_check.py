"""One-off: n = 16384 (128 blocks) sanity -- 64-bit offsets, task counts, memory -- against numpy."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from madaiemulator_b200 import engine, datasets as ds
n, d = 16384, 5
X = ds.synthetic_design(n, d); y = ds.synthetic_response(X)
ctx = engine.Context(0)
m = engine.Model(ctx, X, y, 1, 0, max_slots=2)
th = ds.default_theta_less_amp(d)
t0 = time.time(); r = m.loglik_grad_batch(np.stack([th, th])); t1 = time.time()
r = m.loglik_grad_batch(np.stack([th, th])); t2 = time.time()
print("n=16384 B=2: first %.1f ms, second %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), r["status"], r["negL"], r["sigma2"])
C = m.cov_matrix(np.concatenate([[0.0], th]))
t0 = time.time(); sign, ld = np.linalg.slogdet(C); print("numpy slogdet %.1f s" % (time.time() - t0))
a = np.linalg.solve(C, np.stack([y, np.ones(n)], axis=1))
beta = (np.ones(n) @ a[:, 0]) / (np.ones(n) @ a[:, 1])
res = y - beta
negL = 0.5 * ld + (n / 2.0) * 1.83788 + 0.5 * res @ (a[:, 0] - beta * a[:, 1])
print("negL gpu %.12g numpy %.12g rel %.2e" % (r["negL"][0], negL, abs(r["negL"][0] - negL) / abs(negL)))
print("flops n^3 = %.3g -> %.1f TFLOP/s effective" % (n ** 3, 2 * n ** 3 / ((t2 - t1)) / 1e12))
