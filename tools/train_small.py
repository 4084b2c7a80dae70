"""configs[0] head to head (BASELINE.json: the reference's own CPU-runnable cases): the reference's maxWithMultiMin
(oracle/_ref, compiled from the reference sources, ONE host thread -- what each of its pool threads does,
estimate_threaded.c:97-113) against the engine's restart front on one B200, same fixture, same number of restarts, both
scored with the port oracle.  No extrapolation.  Usage: python tools/train_small.py [tries]   (measurement tool: it
may use oracle/, the product does not)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import load_golden
from madaiemulator_b200 import engine
from oracle.pyoracle import PortOracle, RefOracle, ref_available

tries_arg = int(sys.argv[1]) if len(sys.argv) > 1 else 50
ctx = engine.Context(0)
rows = []
for name in ("uni-simple-o1", "uni-2d-o0", "multi-simple-pc0-o0", "multi-simple-pc0-o1", "synthetic-n256-d10-o1"):
    c = load_golden(name)
    tries = tries_arg if c["n"] <= 100 else max(1, tries_arg * 2 // 5)  # the reference needs ~3 s per restart at n=256
    po = PortOracle(c["X"], c["y"], c["kernel"], c["order"])
    score = lambda th: -po.loglik_grad(th[1:], want_grad=False)["negL"]
    m = engine.Model(ctx, c["X"], c["y"], c["kernel"], c["order"], max_slots=64)
    engine.estimate_thetas(m, max_tries=4, nchains=4, seed=3)  # warm-up: graphs, workspaces
    row = {"case": name, "n": int(c["n"]), "d": int(c["d"]), "restarts": tries}
    for chains in (min(tries, 64), 1):
        t0 = time.perf_counter()
        th, best, st = engine.estimate_thetas(m, max_tries=tries, nchains=chains, seed=11)
        dt = time.perf_counter() - t0
        row["engine_chains%d" % chains] = {"seconds": dt, "loglik": score(th), "evaluations": st["evaluations"], "calls": st["batches"]}
    if ref_available():
        ref = RefOracle(c["X"], c["y"], c["kernel"], c["order"])
        t0 = time.perf_counter()
        ref_best, ref_th = ref.max_with_multimin(tries, 11)
        dt = time.perf_counter() - t0
        row["reference_1thread"] = {"seconds": dt, "loglik": score(ref_th)}
        row["speedup_front"] = dt / row["engine_chains%d" % min(tries, 64)]["seconds"]
        row["speedup_one_chain"] = dt / row["engine_chains1"]["seconds"]
    rows.append(row)
    print(json.dumps(row), flush=True)
    m.close()
ctx.close()
