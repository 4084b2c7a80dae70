"""Small end-to-end pass over every kernel (for compute-sanitizer --tool memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from madaiemulator_b200 import engine, datasets as ds
ctx = engine.Context(0)
for (n, d, order, kernel) in [(300, 3, 1, 1), (130, 2, 2, 3)]:
    X = ds.synthetic_design(n, d); y = ds.synthetic_response(X)
    m = engine.Model(ctx, X, y, kernel, order, max_slots=3)
    th = np.tile(ds.default_theta_less_amp(d, kernel), (5, 1))
    r = m.loglik_grad_batch(th)
    assert np.all(r["status"] == 0)
    full = np.concatenate([[0.1], th[0]]) if kernel == 1 else np.array([1.2, 0.05, 0.3])
    C = m.cov_matrix(full)
    e = m.emulator(full)
    mu, var = e.emulate(ds.synthetic_queries(200, d))
    K = m.k_vectors(full, ds.synthetic_queries(3, d))
    Z = np.stack([y, y[::-1]], axis=1)
    m.set_training_multi(Z)
    r2 = m.loglik_grad_batch(th[:4], comp=[0, 1, 1, 0])
    e0, e1 = m.emulator(full, 0), m.emulator(full, 1)
    pm, pv = engine.predict_multi([e0, e1], ds.synthetic_queries(150, d), np.zeros(3), np.ones((3, 2)), np.ones(2))
    for x in (e, e0, e1): x.close()
    m.close()
print("sanitize pass ok", ctx.debug_exp(np.array([-1.0]))[0])
ctx.close()
