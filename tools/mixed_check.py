"""Bit-consistency of the batched evaluator under every call shape a narrow restart front produces: counts 1..slots,
arbitrary gradient patterns, value-only calls in between (stale workspace), repeated many times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from madaiemulator_b200 import engine, datasets as ds

n, d, slots = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (2048, 15, 8)
X = ds.synthetic_design(n, d); y = ds.synthetic_response(X)
ctx = engine.Context(0)
m = engine.Model(ctx, X, y, 1, 0, max_slots=slots)
ranges = engine.optimization_ranges(1, X)
rng = np.random.default_rng(1)
P = 24
ths = ranges[1:, 0] + rng.uniform(size=(P, d + 1)) * (ranges[1:, 1] - ranges[1:, 0])
truth = m.loglik_grad_batch(ths, want_grad=True)
truth2 = m.loglik_grad_batch(ths, want_grad=True)
assert np.array_equal(truth["negL"], truth2["negL"]) and np.array_equal(truth["grad"], truth2["grad"]), "uniform gradient calls differ between runs"
bad = 0
for it in range(300):
    cnt = int(rng.integers(1, slots + 1))
    idx = rng.choice(P, cnt, replace=False)
    pat = rng.integers(0, 2, cnt) if it % 3 else np.zeros(cnt, int)
    r = m.loglik_grad_batch(ths[idx], want_grad=pat)
    ok = np.array_equal(r["negL"], truth["negL"][idx]) and np.array_equal(r["sigma2"], truth["sigma2"][idx])
    for k in range(cnt):
        if pat[k]:
            ok = ok and np.array_equal(r["grad"][k], truth["grad"][idx[k]])
    if not ok:
        bad += 1
        if bad <= 5:
            print("MISMATCH it=%d cnt=%d pat=%s idx=%s" % (it, cnt, pat, idx))
            print("  negL got ", r["negL"]); print("  negL true", truth["negL"][idx])
            for k in range(cnt):
                if pat[k] and not np.array_equal(r["grad"][k], truth["grad"][idx[k]]):
                    print("  grad[%d] got %s\n          true %s" % (k, r["grad"][k][:4], truth["grad"][idx[k]][:4]))
print("mismatching calls:", bad, "of 300")
sys.exit(1 if bad else 0)
