"""Component sharding must not change a bit: all PCA components in one evaluation front (one device) against one
component at a time (what every rank does at N = ncomp).  Usage: python tools/shard_check.py [n] [d] [ncomp] [restarts]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from madaiemulator_b200 import engine, datasets as ds

n, d, ncomp, restarts = [int(a) for a in sys.argv[1:5]] if len(sys.argv) > 4 else (2048, 15, 8, 8)
step_max = int(os.environ.get("STEP_MAX", "4"))
X, Y = ds.synthetic_model(n, d, nt=ncomp + 1)
Z = np.ascontiguousarray(ds.pca_decompose(Y, vfrac=2.0)["Z"][:, :ncomp])
ranges = engine.optimization_ranges(engine.POWEREXP, X)
ctx = engine.Context(0)


def train(components, first, stride):
    m = engine.Model(ctx, X, Z[:, components[0]], engine.POWEREXP, 0, max_slots=restarts * len(components))
    m.set_training_multi(Z[:, components])
    th, best, st = engine.estimate_thetas_multi(m, len(components), ranges, max_tries=restarts, nchains=restarts, seed=ds.SEED,
                                                step_max=step_max, first_component=first, component_stride=stride)
    m.close()
    return th, best, st


th_all, best_all, st_all = train(list(range(ncomp)), 0, 1)
print("one front:", st_all["evaluations"], "evaluations,", st_all["value_evaluations"], "value-only,", st_all["batches"], "calls")
bad = 0
for c in range(ncomp):
    th, best, st = train([c], c, ncomp)
    same = np.array_equal(th[0], th_all[c]) and best[0] == best_all[c]
    bad += not same
    print("component %d alone: %s  best %.9f vs %.9f  (%d evaluations)" % (c, "identical" if same else "DIFFERENT", best[0], best_all[c], st["evaluations"]))
print("sharded_identical:", bad == 0)
sys.exit(1 if bad else 0)
