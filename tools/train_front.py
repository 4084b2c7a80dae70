"""Wall time of one restart front: ncomp components x restarts chains at (n, d), step_max BFGS iterations (the cfg4
workload of bench.py's strong-scaling section for one rank).  Usage: python tools/train_front.py n d ncomp restarts [groups]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from madaiemulator_b200 import engine, datasets as ds
n, d, ncomp, restarts = [int(a) for a in sys.argv[1:5]]
groups = int(sys.argv[5]) if len(sys.argv) > 5 else 4
X, Y = ds.synthetic_model(n, d, nt=9)
Z = np.ascontiguousarray(ds.pca_decompose(Y, vfrac=2.0)["Z"][:, :ncomp])
ranges = engine.optimization_ranges(engine.POWEREXP, X)
ctx = engine.Context(0)
ctx.set_groups(groups)
m = engine.Model(ctx, X, Z[:, 0], engine.POWEREXP, 0, max_slots=restarts * ncomp)
m.set_training_multi(Z)
for rep in range(2):
    t0 = time.time()
    th, best, st = engine.estimate_thetas_multi(m, ncomp, ranges, max_tries=restarts, nchains=restarts, seed=ds.SEED, step_max=4,
                                                first_component=0, component_stride=2)
    dt = time.time() - t0
    print("n=%d d=%d front %d x %d, groups %d, EMUB_AUX_MAX=%s: %.2f s, %d evaluations (%d value-only) in %d calls, %.1f evals/s" %
          (n, d, ncomp, restarts, groups, os.environ.get("EMUB_AUX_MAX", "default"), dt, st["evaluations"], st["value_evaluations"], st["batches"],
           st["evaluations"] / dt))
