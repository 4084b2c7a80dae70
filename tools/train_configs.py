"""Wall time of the whole restart driver (host C BFGS chains + batched GPU evaluations) on BASELINE configs 3 and 4
(one component of config 4 on one GPU).  Usage: python tools/train_configs.py [cfg3|cfg4|all]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from madaiemulator_b200 import engine, datasets as ds

which = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = engine.Context(0)
cases = []
if which in ("cfg3", "all"):
    cases.append(("cfg3: n=2048, d=6, Matern52, order 2, 64 restarts in one front", 2048, 6, engine.MATERN52, 2, 64, 64, 64))
if which in ("cfg4", "all"):
    cases.append(("cfg4 (one of the 8 components): n=8192, d=15, power-exp, order 0, 32 restarts", 8192, 15, engine.POWEREXP, 0, 32, 32, 32))
for name, n, d, kernel, order, tries, chains, slots in cases:
    X = ds.synthetic_design(n, d)
    y = ds.synthetic_response(X)
    m = engine.Model(ctx, X, y, kernel, order, max_slots=slots)
    for polish in (0, 100):
        t0 = time.time()
        th, best, st = engine.estimate_thetas(m, max_tries=tries, nchains=chains, seed=1, polish_steps=polish)
        dt = time.time() - t0
        print("%s | polish %d: %.2f s, %d evaluations in %d batches (%.0f evals/s), best log-likelihood %.6f, %d/%d chains converged" %
              (name, polish, dt, st["evaluations"], st["batches"], st["evaluations"] / dt, best, st["success_count"], tries))
    m.close()
