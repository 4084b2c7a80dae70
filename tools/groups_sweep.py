"""evals/s at small batches (n=4096, d=10) against the number of stream groups and the side-stream switch.
Usage: EMUB_AUX_MAX=.. python tools/groups_sweep.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from madaiemulator_b200 import engine, datasets as ds
n, d = 4096, 10
X = ds.synthetic_design(n, d); y = ds.synthetic_response(X)
ctx = engine.Context(0)
for B in (4, 8, 16):
    m = engine.Model(ctx, X, y, 1, 0, max_slots=B)
    th = np.tile(ds.default_theta_less_amp(d), (B, 1))
    for g in (1, 2, 4):
        if g > B: continue
        ctx.set_groups(g)
        out = []
        for want in (True, False):
            for _ in range(3): m.loglik_grad_batch(th, want_grad=want)
            t0 = time.perf_counter()
            for _ in range(10): m.loglik_grad_batch(th, want_grad=want)
            out.append(B * 10 / (time.perf_counter() - t0))
        print("AUX_MAX=%s B=%d groups=%d: %.1f evals/s with gradient, %.1f value-only" % (os.environ.get("EMUB_AUX_MAX", "default"), B, g, out[0], out[1]), flush=True)
    m.close()
