"""The reference's own evalFnGradMulti MEASURED at the headline size n=4096, d=10 (one evaluation per host thread, all
cores; about 20 minutes) -- the anchor that bench.py's default CPU baseline extrapolates to from n=2048.
Usage: python tools/cpu_ref_n4096.py [n]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ncpu = os.cpu_count() or 1
t0 = time.time()
evals, secs, kind, cores = bench.cpu_eval_sample(n, ncpu)
print(json.dumps({"n": n, "d": bench.D_MODEL, "evals": evals, "seconds": secs, "evals_per_s": evals / secs, "threads": cores, "kind": kind,
                  "wall": time.time() - t0}))
