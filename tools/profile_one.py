"""One likelihood+gradient batch (B=8, one stream group) and one prediction chunk at n=4096, d=10: the
command that the ncu captures under profiles/ are taken from.  GEMM launch order: 124 factorisation
launches, then W^T W (launch 125), then the prediction product (launch 126)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from madaiemulator_b200 import engine, datasets as ds

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n, d = 4096, 10
X = ds.synthetic_design(n, d); y = ds.synthetic_response(X)
ctx = engine.Context(0, groups=1)
m = engine.Model(ctx, X, y, 1, 0, max_slots=B)
th = np.tile(ds.default_theta_less_amp(d), (B, 1))
t0 = time.time(); r = m.loglik_grad_batch(th); t1 = time.time()
print("batch", B, "ms", (t1 - t0) * 1e3, "negL", r["negL"][0])
e = m.emulator(np.concatenate([[0.0], th[0]]))
mu, var = e.emulate(ds.synthetic_queries(16384, d))
print("pred ok", mu[:2], var[:2])
