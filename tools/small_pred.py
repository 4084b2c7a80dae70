"""Latency of small prediction calls (m = 1, 8, 128, 1024 points) at n=4096, d=10 and at n=100 (multi-simple size)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from madaiemulator_b200 import engine, datasets as ds

ctx = engine.Context(0)
for n, d in ((4096, 10), (1024, 10), (100, 3)):
    X = ds.synthetic_design(n, d); y = ds.synthetic_response(X)
    m = engine.Model(ctx, X, y, 1, 0, max_slots=1)
    e = m.emulator(np.concatenate([[0.0], ds.default_theta_less_amp(d)]))
    for mq in (1, 8):
        pts = ds.synthetic_queries(mq, d)
        for _ in range(5):
            e.emulate_few(pts)
        t0 = time.time()
        for _ in range(200):
            e.emulate_few(pts)
        us = (time.time() - t0) / 200 * 1e6
        a, b = e.emulate_few(pts), e.emulate(pts)
        ctx.profile(True)
        e.emulate_few(pts)
        prof = ctx.profile_read()
        ctx.profile(False)
        print("n=%d m=%d FEW: %.1f us per call (%.0f points/s), max diff vs batch path %.1e %.1e | " % (n, mq, us, mq / us * 1e6, np.max(np.abs(a[0] - b[0])), np.max(np.abs(a[1] - b[1]))) +
              ", ".join("%s %.1f us" % (k, v["ms"] * 1e3) for k, v in prof.items() if v["launches"]))
    for mq in (1, 8, 128, 1024):
        pts = ds.synthetic_queries(mq, d)
        for _ in range(5):
            e.emulate(pts)
        t0 = time.time()
        reps = 200
        for _ in range(reps):
            e.emulate(pts)
        us = (time.time() - t0) / reps * 1e6
        ctx.profile(True)
        e.emulate(pts)
        prof = ctx.profile_read()
        ctx.profile(False)
        print("n=%d m=%d: %.1f us per call (%.0f points/s) | " % (n, mq, us, mq / us * 1e6) +
              ", ".join("%s %.1f us" % (k, v["ms"] * 1e3) for k, v in prof.items() if v["launches"]))
    e.close(); m.close()

# all PCA components of one model + back-projection for one point (emulate_point_multi in the glue)
for n, d, nr, nt in ((100, 3, 5, 6), (1024, 10, 8, 12)):
    X = ds.synthetic_design(n, d)
    rng = np.random.default_rng(2)
    Z = rng.normal(size=(n, nr))
    m = engine.Model(ctx, X, Z[:, 0], 1, 0, max_slots=1)
    m.set_training_multi(Z)
    emus = [m.emulator(np.concatenate([[0.0], ds.default_theta_less_amp(d)]), comp=c) for c in range(nr)]
    ybar, U, lam = rng.normal(size=nt), rng.normal(size=(nt, nr)), rng.uniform(0.5, 2, nr)
    one = ds.synthetic_queries(1, d)
    for few in (False, True):
        for _ in range(5):
            engine.predict_multi(emus, one, ybar, U, lam, few=few)
        t0 = time.time()
        for _ in range(200):
            engine.predict_multi(emus, one, ybar, U, lam, few=few)
        print("multi n=%d nr=%d nt=%d one point, %s path: %.1f us per call" % (n, nr, nt, "few" if few else "batched", (time.time() - t0) / 200 * 1e6))
    for e in emus:
        e.close()
    m.close()
