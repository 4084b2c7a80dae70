#!/usr/bin/env python
"""bench.py -- the reference's headline hot path on B200: GP log-likelihood + theta-gradient evaluations per
second at n=4096, d=10, FP64 (BASELINE.json), plus emulated points per second on the same model.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

A "step" is one batched call of the hot path: B theta points (one optimizer-restart front) -> (-L, gradient,
sigma^2) each, i.e. B x evalFnGradMulti (reference src/libEmu/maxmultimin.c:615).  Multi-GPU: one process per
GPU (torchrun), every rank evaluates its own B restarts of the same model (independent units, no data-path
collective), `value` = all ranks' evaluations / max-over-ranks device time ("scaling": "weak").

Fixed-work companions (`extra.strong`, SURVEY 8e / BASELINE configs 4 and 5), through the product's own sharding
API, timed the same way at every N:
  cfg4  8 PCA components at n=8192, d=15, a fixed restart budget per component: component c -> rank c mod N
        (emub_estimate_thetas_multi with first_component / component_stride; the serial loop it replaces is
        src/multivar_support.c:20-27), final all_gather of the thetas;
  cfg5  a fixed 10^7 query points on the n=4096, d=10 emulator, split in contiguous blocks (the per-point loop it
        replaces is src/interactive_emulator.c:416-441), final all_gather of (mean, variance).
At N > 1 rank 0 then repeats the whole job on its single device: that gives the speed-up in the same run and
`sharded_identical` (the gathered answers equal the single-device ones bit for bit).  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_MODEL, D_MODEL, ORDER = 4096, 10, 0
CPU_SAMPLE_SIZES = (1024, 2048)
CPU_POINTS = 1000
# dram__bytes_read.sum + dram__bytes_write.sum over the 125 k_gemm launches of ONE likelihood+gradient batch of 8
# matrices at n=4096 (ncu, profiles/r01_gemm_family_traffic.txt); the matrices of a batch are independent, so the
# traffic of a batch of B is B/8 of this
GEMM_FAMILY_DRAM_BYTES_B8, GEMM_FAMILY_LAUNCHES = 5.348e9, 125
FP64_DATASHEET_TFLOPS = 40.0
# fixed-work (strong scaling) workloads
CFG4_N, CFG4_D, CFG4_COMPONENTS, CFG4_RESTARTS, CFG4_STEP_MAX = 8192, 15, 8, 8, 4
CFG5_POINTS = 10_000_000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="theta points per step per GPU (BASELINE cfg3: 64 restarts batched)")
    ap.add_argument("--pred-points", type=int, default=1 << 18, help="query points per prediction step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the fixed-work cfg4 / cfg5 section")
    ap.add_argument("--cpu-sizes", default=",".join(str(s) for s in CPU_SAMPLE_SIZES),
                    help="model sizes the CPU reference is timed at (the largest one is extrapolated to n=4096)")
    ap.add_argument("--groups", type=int, default=0, help="stream groups (0 = min(4, batch/4))")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def theta_batch(ranges, B, seed):
    """B starting points drawn uniformly in the optimizer's search ranges, as set_random_init_value does
    (reference maxmultimin.c:789-804); theta without the amplitude."""
    from madaiemulator_b200 import datasets as ds
    nth1 = ranges.shape[0] - 1
    u = ds.uniform01(seed, np.arange(B * nth1, dtype=np.uint64)).reshape(B, nth1)
    lo, hi = ranges[1:, 0], ranges[1:, 1]
    return lo[None, :] + u * (hi - lo)[None, :]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_eval_sample(n, nthreads):
    """nthreads concurrent evalFnGradMulti calls of the reference's own C (oracle/_ref; the plain-C port if that build
    is absent) at model size n, one independent model per thread -- the reference's own parallel model
    (estimate_threaded.c:97,172).  Returns (evals, seconds, kind, threads)."""
    from madaiemulator_b200 import datasets as ds
    from oracle import pyoracle as po
    X = ds.synthetic_design(n, D_MODEL)
    y = ds.synthetic_response(X)
    th = ds.default_theta_less_amp(D_MODEL)
    if po.ref_available():
        t = po.time_ref_eval_grad(X, y, th, 1, ORDER, nthreads=nthreads, reps=1)
        return nthreads, t, "reference", nthreads
    o = po.PortOracle(X, y, 1, ORDER)
    t0 = time.perf_counter()
    o.loglik_grad(th)
    return 1, time.perf_counter() - t0, "port", 1


def cpu_reference_rate(nthreads, sizes):
    """CPU baseline of the headline metric: evalFnGradMulti MEASURED at every size in `sizes` (default 1024 and 2048; one
    n=2048 evaluation is ~100 s on one core, n=4096 would be ~15 min per evaluation), the largest one extrapolated to
    n=4096 by the cubic cost of the path ((2(T-1)+1) n^3 flops).  The exponent fitted between the two
    largest measured sizes is reported too: above 3 (cache misses of the naive BLAS grow with n) it makes the cubic
    extrapolation an upper bound of the CPU rate, i.e. the GPU/CPU ratio a lower bound."""
    samples = []
    kind, cores = "reference", nthreads
    for n in sorted(sizes):
        evals, secs, kind, cores = cpu_eval_sample(n, nthreads)
        samples.append({"n": n, "evals": evals, "seconds": round(secs, 3), "evals_per_s": evals / secs})
    big = samples[-1]
    value = big["evals_per_s"] * (big["n"] / float(N_MODEL)) ** 3
    exponent = None
    if len(samples) >= 2:
        a, b = samples[-2], samples[-1]
        exponent = float(np.log(a["evals_per_s"] / b["evals_per_s"]) / np.log(b["n"] / float(a["n"])))
    txt = "; ".join("n=%d: %d evals in %.1f s (%.4f evals/s)" % (s["n"], s["evals"], s["seconds"], s["evals_per_s"]) for s in samples)
    out = dict(value=value, unit="evals/s", cores=cores, kind=kind,
               sample="evalFnGradMulti, d=%d, %d threads, measured at %s; fitted cost exponent %s; value = the n=%d rate "
                      "x (%d/%d)^3 (extrapolated, not measured at n=%d)"
                      % (D_MODEL, cores, txt, ("%.2f" % exponent) if exponent is not None else "n/a", big["n"], big["n"], N_MODEL, N_MODEL),
               samples=samples, fitted_exponent=exponent,
               value_with_fitted_exponent=(big["evals_per_s"] * (big["n"] / float(N_MODEL)) ** exponent) if exponent else None,
               seconds=sum(s["seconds"] for s in samples))
    return out


def cpu_points_baseline(nthreads, npoints=CPU_POINTS):
    """CPU baseline of the second metric: the reference's emulate_point (emulator_struct.c:124) on a trained-model shape
    n=4096, d=10 -- MEASURED at the full size: alloc_emulator_struct once (set-up, reported separately), then npoints
    points shared out over nthreads threads (emulate_point is re-entrant)."""
    from madaiemulator_b200 import datasets as ds
    from oracle import pyoracle as po
    X = ds.synthetic_design(N_MODEL, D_MODEL)
    y = ds.synthetic_response(X)
    full = np.concatenate([[0.0], ds.default_theta_less_amp(D_MODEL)])
    pts = ds.synthetic_queries(npoints, D_MODEL)
    if po.ref_available():
        o = po.RefOracle(X, y, 1, ORDER)
        t0 = time.perf_counter()
        e = o.emulator(full)
        t_setup = time.perf_counter() - t0
        mean, var = np.empty(npoints), np.empty(npoints)
        t = po.RefOracle.lib().ref_time_emulate(e.h, po._P(po._c(pts)), npoints, nthreads, po._P(mean), po._P(var))
        kind, cores = "reference", nthreads
    else:
        o = po.PortOracle(X, y, 1, ORDER)
        t0 = time.perf_counter()
        e = o.emulator(full)
        t_setup = time.perf_counter() - t0
        t0 = time.perf_counter()
        mean, var = e.emulate(pts)
        t = time.perf_counter() - t0
        kind, cores = "port", 1
    return dict(value=npoints / t, unit="points/s", cores=cores, kind=kind,
                sample="emulate_point at n=%d, d=%d: %d points in %.2f s on %d threads (measured at the full size); "
                       "set-up (alloc_emulator_struct: covariance, Cholesky, inverse) %.1f s on one thread, not included"
                       % (N_MODEL, D_MODEL, npoints, t, cores, t_setup),
                setup_seconds=round(t_setup, 2), seconds=round(t, 3)), mean, var, pts, full


def run_reference(args, rank, world):
    if rank != 0:
        return
    ncpu = os.cpu_count() or 1
    sizes = [int(s) for s in args.cpu_sizes.split(",") if s]
    # a step = one bounded sample of the workload: nthreads concurrent evaluations at the smallest sample size; the
    # larger sizes are measured once (they anchor the extrapolation to n = 4096)
    small = min(sizes)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_eval_sample(small, ncpu)
    anchor = cpu_reference_rate(ncpu, sizes)
    rates, secs = [], 0.0
    for _ in range(args.steps):
        evals, t, kind, cores = cpu_eval_sample(small, ncpu)
        rates.append(evals / t)
        secs += t
    # per-step rate at the small size, carried to n = 4096 through the measured ratio to the largest size and the cubic law
    small_anchor = [s for s in anchor["samples"] if s["n"] == small][0]["evals_per_s"]
    v = float(np.mean(rates)) / small_anchor * anchor["value"]
    cb = dict(anchor)
    cb["value"] = v
    cb.pop("seconds", None)
    out = {"impl": "reference", "metric": "loglik_grad_evals_per_s", "value": v, "unit": "evals/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, world), "cpu_baseline": cb,
           "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def workload_config(args, world):
    return {"workload": "cfg5/headline: synthetic design n=4096, d=10, power-exponential, regression order 0; "
                        "B theta points per step per GPU drawn in the optimizer ranges -> (-L, gradient, sigma2) each",
            "n": N_MODEL, "d": D_MODEL, "kernel": "power-exponential", "regression_order": ORDER,
            "batch_per_gpu": args.batch, "parallelism": "restarts x%d (independent, no collective)" % world,
            "l2": "inputs larger than L2: %.1f GB of matrix workspace touched per step vs 126 MB L2"
                  % (args.batch * 3 * N_MODEL * N_MODEL * 8 / 1e9)}


def strong_section(engine, ds, ctx, rank, world, local, barrier, max_over_ranks):
    """Fixed-work cfg5 (prediction, query blocks over ranks) and cfg4 (training, components over ranks)."""
    import torch
    import torch.distributed as dist
    from madaiemulator_b200 import sharding
    dev = torch.device("cuda", local)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    last_per_rank = []  # device seconds of every rank in the last collective timed() region

    def timed(fn, collective=True):
        """seconds between two events on the library's stream around fn(); collective: barrier + synchronize on both
        sides and the max over ranks (the sharded runs); otherwise this rank alone (the single-device repeat)"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if collective:
            barrier()
        else:
            torch.cuda.synchronize()
        e0.record(stream)
        t0 = time.perf_counter()
        r = fn()
        e1.record(stream)
        e1.synchronize()
        wall = time.perf_counter() - t0
        dt = e0.elapsed_time(e1) * 1e-3
        if collective:
            barrier()
            per_rank = [dt]
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                outs = [torch.empty_like(t) for _ in range(world)]
                dist.all_gather(outs, t)
                per_rank = [float(o.item()) for o in outs]
            last_per_rank[:] = per_rank
            return r, max(per_rank), max_over_ranks(wall)
        return r, dt, wall

    def gather_rows(local_rows, counts):
        """all_gather of per-rank row blocks: the final gather of thetas / predictions, the only cross-rank traffic"""
        return sharding.gather_row_blocks(local_rows, counts, device=dev if world > 1 else None)

    out = {}
    identical = {}
    # ---- cfg5: 10^7 query points on the n=4096, d=10 emulator, contiguous blocks -------------------------------
    X = ds.synthetic_design(N_MODEL, D_MODEL)
    y = ds.synthetic_response(X)
    model = engine.Model(ctx, X, y, engine.POWEREXP, ORDER, max_slots=1)
    emu = model.emulator(np.concatenate([[0.0], ds.default_theta_less_amp(D_MODEL)]))
    qseed = ds.SEED + 5
    lo, hi = sharding.block_range(CFG5_POINTS, world, rank)
    pts = ds.synthetic_queries(hi - lo, D_MODEL, seed=qseed, row0=lo)
    emu.emulate(pts[:1 << 15])  # warm-up: workspace allocation, first launches
    (mean, var), t_sh, wall_sh = timed(lambda: emu.emulate(pts))
    counts = [b - a for a, b in (sharding.block_range(CFG5_POINTS, world, r) for r in range(world))]
    cfg5 = {"workload": "cfg5: %d query points on a trained-model shape n=%d, d=%d, host buffers in and out "
                        "(emub_predict_batch); contiguous block per rank, final all_gather of (mean, variance)" % (CFG5_POINTS, N_MODEL, D_MODEL),
            "seconds": t_sh, "wall_seconds": wall_sh, "points_per_s": CFG5_POINTS / t_sh, "points_per_rank": counts,
            "seconds_per_rank": [round(x, 4) for x in last_per_rank]}
    if world > 1:
        blocks = gather_rows(np.column_stack([mean, var]), counts)
        if rank == 0:
            allpts = ds.synthetic_queries(CFG5_POINTS, D_MODEL, seed=qseed)
            (m1, v1), t1, _ = timed(lambda: emu.emulate(allpts), collective=False)
            got = np.concatenate(blocks, axis=0)
            identical["cfg5"] = bool(np.array_equal(got[:, 0], m1) and np.array_equal(got[:, 1], v1))
            cfg5.update(single_device_seconds=t1, speedup=t1 / t_sh, efficiency=t1 / t_sh / world)
            del allpts, got
        barrier()
    else:
        identical["cfg5"] = True
    out["cfg5_predict"] = cfg5
    emu.close()
    model.close()
    del pts

    # ---- cfg4: 8 PCA components of a 9-output model on a shared n=8192, d=15 design -----------------------------
    X, Y = ds.synthetic_model(CFG4_N, CFG4_D, nt=CFG4_COMPONENTS + 1)
    pca = ds.pca_decompose(Y, vfrac=2.0)  # keeps all nt-1 components (gen_pca_decomp's loop, multi_modelstruct.c:267-272)
    Z = np.ascontiguousarray(pca["Z"][:, :CFG4_COMPONENTS])
    ranges = engine.optimization_ranges(engine.POWEREXP, X)
    nth = CFG4_D + 2

    def train(components, first, stride, collective):
        m = engine.Model(ctx, X, Z[:, components[0]], engine.POWEREXP, 0, max_slots=CFG4_RESTARTS * len(components))
        m.set_training_multi(Z[:, components])
        slots = m.slots
        res, dt, wall = timed(lambda: engine.estimate_thetas_multi(m, len(components), ranges, max_tries=CFG4_RESTARTS,
                                                                   nchains=CFG4_RESTARTS, seed=ds.SEED, step_max=CFG4_STEP_MAX,
                                                                   first_component=first, component_stride=stride), collective)
        m.close()
        return res, dt, wall, slots

    mine = sharding.round_robin(CFG4_COMPONENTS, world, rank)
    (th_loc, best_loc, st), t_sh, wall_sh, slots = train(mine, rank, world, True)
    cfg4_per_rank = [round(x, 4) for x in last_per_rank]
    counts = [len(sharding.round_robin(CFG4_COMPONENTS, world, r)) for r in range(world)]
    blocks = gather_rows(np.column_stack([th_loc, best_loc]), counts)
    gathered = sharding.scatter_round_robin(blocks, CFG4_COMPONENTS)
    evals = np.array([float(st["evaluations"]), float(st["batches"]), float(st["value_evaluations"])])
    if world > 1:
        t = torch.from_numpy(evals).to(dev)
        dist.all_reduce(t)
        evals = t.cpu().numpy()
    cfg4 = {"workload": "cfg4: %d PCA components, n=%d, d=%d, power-exponential, %d restarts per component, <= %d BFGS "
                        "iterations each (emub_estimate_thetas_multi); component c -> rank c mod N, final all_gather of thetas"
                        % (CFG4_COMPONENTS, CFG4_N, CFG4_D, CFG4_RESTARTS, CFG4_STEP_MAX),
            "seconds": t_sh, "wall_seconds": wall_sh, "seconds_per_rank": cfg4_per_rank, "evaluations": int(evals[0]),
            "value_only_evaluations": int(evals[2]),
            "batched_calls": int(evals[1]), "evals_per_s": evals[0] / t_sh, "components_per_rank": counts,
            "front_width_per_rank": CFG4_RESTARTS * len(mine), "slots_rank0": slots,
            "finite_components": int(np.sum(gathered[:, nth] > -1e6)),
            "best_loglik": [float(v) for v in gathered[:, nth]]}
    if world > 1:
        if rank == 0:
            (th1, best1, st1), t1, _, _ = train(list(range(CFG4_COMPONENTS)), 0, 1, False)
            identical["cfg4"] = bool(np.array_equal(gathered[:, :nth], th1) and np.array_equal(gathered[:, nth], best1))
            cfg4.update(single_device_seconds=t1, single_device_evals_per_s=st1["evaluations"] / t1, speedup=t1 / t_sh,
                        efficiency=t1 / t_sh / world,
                        limit="per-rank front shrinks from %d to %d concurrent chains: fewer matrices per launch (see the "
                              "batch-size sweep in DESIGN.md) and the host BFGS turn-around between batched calls is paid "
                              "per front; no device-to-device traffic besides the final gather"
                              % (CFG4_RESTARTS * CFG4_COMPONENTS, CFG4_RESTARTS * len(mine)))
        barrier()
    else:
        identical["cfg4"] = True
    out["cfg4_train"] = cfg4
    return out, identical


def main():
    args = parse_args()
    rank, world, local = dist_env()
    if args.gpus > 1 and "RANK" not in os.environ:
        # not launched by torchrun: re-launch ourselves the way the driver would
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from madaiemulator_b200 import datasets as ds
    from madaiemulator_b200 import engine

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # the CPU baselines run on rank 0's host cores while the GPUs are idle (before the timed GPU regions), at every N
    cpu = cpu_pts = None
    cpu_thread = None
    X = ds.synthetic_design(N_MODEL, D_MODEL)
    y = ds.synthetic_response(X)
    ranges = engine.optimization_ranges(engine.POWEREXP, X)  # host C: optstruct.c:142-226
    B = args.batch
    thetas = theta_batch(ranges, B, ds.SEED + 17 + rank)
    nth1 = D_MODEL + 1

    ctx = engine.Context(local)
    ctx.set_groups(args.groups if args.groups > 0 else min(4, max(1, B // 4)))
    model = engine.Model(ctx, X, y, engine.POWEREXP, ORDER, max_slots=B)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    d_thetas = torch.tensor(thetas, dtype=torch.float64, device="cuda").contiguous()
    d_out = torch.zeros(B, nth1 + 4, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()

    def step_dev(want_grad=True):
        model.loglik_grad_batch_dev(d_thetas.data_ptr(), B, want_grad, d_out.data_ptr())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # samples through warm-up and the timed region (nvidia-smi takes ~1 s to start reporting)
    for _ in range(max(3, args.warmup)):
        step_dev()
    barrier()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    res = d_out.cpu().numpy()
    nfail = int(np.sum(res[:, 2] != 0))
    value = world * B * args.steps / (ms * 1e-3)

    # the value-only evaluation (evalFnMulti alone, maxmultimin.c:288: factor without the inverse), same batch
    for _ in range(3):
        step_dev(False)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_dev(False)
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms_val = max_over_ranks(e0.elapsed_time(e1))
    res_val = d_out.cpu().numpy()
    value_only = world * B * args.steps / (ms_val * 1e-3)
    value_only_same_bits = bool(np.array_equal(res_val[:, 0], res[:, 0], equal_nan=True))

    # end to end through the host-pointer C-ABI call: host thetas in, host results out, every step
    for _ in range(2):
        model.loglik_grad_batch(thetas)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r_host = model.loglik_grad_batch(thetas)
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * B * args.steps / t_e2e
    assert np.allclose(r_host["negL"], res[:, 0], rtol=0, atol=0, equal_nan=True)

    # prediction: emulated points per second on the trained-model shape (cfg5), same run
    full = np.concatenate([[0.0], ds.default_theta_less_amp(D_MODEL)])
    emu = model.emulator(full)
    mq = args.pred_points
    pts = ds.synthetic_queries(mq, D_MODEL, seed=ds.SEED + 1 + rank)
    d_pts = torch.tensor(pts, dtype=torch.float64, device="cuda").contiguous()
    d_mean = torch.empty(mq, dtype=torch.float64, device="cuda")
    d_var = torch.empty(mq, dtype=torch.float64, device="cuda")
    psteps = max(1, min(args.steps, 3))
    emu.emulate_dev(d_pts.data_ptr(), mq, d_mean.data_ptr(), d_var.data_ptr())
    barrier()
    e0.record(stream)
    for _ in range(psteps):
        emu.emulate_dev(d_pts.data_ptr(), mq, d_mean.data_ptr(), d_var.data_ptr())
    e1.record(stream)
    e1.synchronize()
    barrier()
    pms = max_over_ranks(e0.elapsed_time(e1))
    pred_value = world * mq * psteps / (pms * 1e-3)
    barrier()
    t0 = time.perf_counter()
    for _ in range(psteps):
        mean_h, var_h = emu.emulate(pts)
    t_pe = max_over_ranks(time.perf_counter() - t0)
    pred_e2e = world * mq * psteps / t_pe
    # latency of the per-point call pattern (emulate_point through the glue -> emub_predict_few), host pointers, rank 0
    single_us = None
    if rank == 0:
        one = pts[:1].copy()
        for _ in range(10):
            emu.emulate_few(one)
        t0 = time.perf_counter()
        for _ in range(200):
            emu.emulate_few(one)
        single_us = (time.perf_counter() - t0) / 200 * 1e6

    out = None
    prof = prof_pred = None
    peak = None
    if rank == 0:
        # per-kernel-family durations: one more step with every launch bracketed by CUDA events (one stream group)
        ctx.profile(True)
        step_dev()
        ctx.synchronize()
        prof = ctx.profile_read()
        ctx.profile(True)  # resets the accumulators
        emu.emulate_dev(d_pts.data_ptr(), min(mq, 16384), d_mean.data_ptr(), d_var.data_ptr())
        ctx.synchronize()
        prof_pred = ctx.profile_read()
        ctx.profile(True)
        step_dev(False)
        ctx.synchronize()
        prof_val = ctx.profile_read()
        ctx.profile(False)
        # FP64 tensor peak: MEASURED_PEAKS.json has no FP64 entry, so measure cuBLAS DGEMM 8192^3 here (burst, best of 5)
        a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        bm = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        torch.matmul(a, bm)
        best = 1e30
        t_e0, t_e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(5):
            t_e0.record()
            torch.matmul(a, bm)
            t_e1.record()
            t_e1.synchronize()
            best = min(best, t_e0.elapsed_time(t_e1))
        peak = 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
        del a, bm
    # a query-point sample of the CPU points baseline doubles as an end-to-end parity check of the prediction path
    emu_check = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu_pts, m_ref, v_ref, pts_ref, full_ref = cpu_points_baseline(os.cpu_count() or 1)
        m_gpu, v_gpu = emu.emulate(pts_ref)
        emu_check = {"points": int(len(pts_ref)), "max_rel_err_mean": float(np.max(np.abs(m_gpu - m_ref) / np.maximum(1e-3, np.abs(m_ref)))),
                     "max_abs_err_var_over_kappa": float(np.max(np.abs(v_gpu - v_ref)) / (np.exp(full_ref[0]) + np.exp(full_ref[1])))}
        cpu = cpu_reference_rate(os.cpu_count() or 1, [int(s) for s in args.cpu_sizes.split(",") if s])
        cpu.pop("seconds", None)
    del d_pts, d_mean, d_var, d_thetas
    emu.close()
    model.close()
    barrier()

    strong = identical = None
    if not args.no_strong:
        strong, identical = strong_section(engine, ds, ctx, rank, world, local, barrier, max_over_ranks)

    if rank == 0:
        gem = [prof[k] for k in ("gemm_chol", "gemm_trtri", "gemm_lauum")]
        g_ms = sum(g["ms"] for g in gem)
        g_fl = sum(g["work"] for g in gem)
        g_n = sum(g["launches"] for g in gem)
        tot_ms = sum(v["ms"] for v in prof.values())
        # algorithmic work of the GEMM-shaped part of one evaluation: n^3/3 (Cholesky) + 2n^3/3 (inverse) = n^3 flops
        # (SURVEY 8d); the tile engine executes ~2% more (diagonal tiles are computed at 64 x 64 granularity)
        achieved = B * float(N_MODEL) ** 3 / (g_ms * 1e-3) / 1e12
        executed = g_fl / (g_ms * 1e-3) / 1e12
        kernels = {k: {"ms": round(v["ms"], 4), "launches": v["launches"],
                       "share": round(v["ms"] / tot_ms, 4)} for k, v in prof.items() if v["launches"]}
        roofline = {"bound": "tensor", "kernel": "emub::k_gemm (FP64 DMMA tile engine: Cholesky TRSM/SYRK, inverse merge, W^T W)",
                    "achieved": round(achieved, 3), "peak": round(peak, 3), "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
                    "frac_of_datasheet_%g" % FP64_DATASHEET_TFLOPS: round(achieved / FP64_DATASHEET_TFLOPS, 4),
                    "traffic": GEMM_FAMILY_DRAM_BYTES_B8 * (B / 8.0) / GEMM_FAMILY_LAUNCHES if N_MODEL == 4096 else None,
                    "traffic_source": "ncu dram bytes summed over the 125 k_gemm launches of one batch of 8 (profiles/"
                                      "r01_gemm_family_traffic.txt), scaled by B/8, per launch on average; 669 MB per evaluation "
                                      "for ~10 passes over 67 MB lower-triangular matrices",
                    "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry; "
                                   "DMMA issue-rate microbenchmark 37.1 TFLOP/s, profiles/r01_dmma_probe.txt; datasheet 40)",
                    "launches": g_n, "avg_launch_ms": round(g_ms / max(1, g_n), 4),
                    "algorithmic_flops_per_step": B * float(N_MODEL) ** 3, "executed_flops_per_step": g_fl,
                    "executed_tflops": round(executed, 3), "step_share": round(g_ms / tot_ms, 4),
                    "eval_flops": float(N_MODEL) ** 3,
                    "eval_tflops_timed_region": round(value / world * float(N_MODEL) ** 3 / 1e12, 3)}
        pg = prof_pred["gemm_pred"]
        pred_tot = sum(v["ms"] for v in prof_pred.values())
        pred_roof = {"kernel": "emub::k_gemm<W K, column sum of squares>", "achieved": round(pg["work"] / (pg["ms"] * 1e-3) / 1e12, 3),
                     "peak": round(peak, 3), "unit": "TFLOP/s",
                     "kernels": {k: {"ms": round(v["ms"], 4), "launches": v["launches"], "share": round(v["ms"] / pred_tot, 4)}
                                 for k, v in prof_pred.items() if v["launches"]}} if pg["ms"] > 0 else None
        flops_pt = float(N_MODEL) ** 2 + 2 * N_MODEL + 3 * N_MODEL * D_MODEL
        gv = [prof_val[k] for k in ("gemm_chol", "gemm_trtri", "gemm_lauum")]
        val_tot = sum(v["ms"] for v in prof_val.values())
        out = {"metric": "loglik_grad_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
               "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(B * nth1 * 8),
                       "d2h_bytes_per_step": int(B * engine.RES_STRIDE * 8)},
               "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
               "failed_points": nfail, "kernels": kernels,
               "extra": {"emulated_points_per_s": {"value": pred_value, "unit": "points/s", "points_per_step_per_gpu": mq,
                                                   "steps": psteps, "ms_per_step": pms / psteps,
                                                   "e2e": {"value": pred_e2e, "unit": "points/s",
                                                           "h2d_bytes_per_step": int(mq * D_MODEL * 8), "d2h_bytes_per_step": int(mq * 16)},
                                                   "roofline": pred_roof,
                                                   "algorithmic_flops_per_point": flops_pt,
                                                   "frac_of_peak_on_algorithmic_flops": round(pred_value / world * flops_pt / 1e12 / peak, 4),
                                                   "cpu_baseline": cpu_pts, "parity_vs_cpu_sample": emu_check},
                         "value_only_evals_per_s": {"value": value_only, "unit": "evals/s", "ms_per_step": ms_val / args.steps,
                                                    "ratio_to_gradient_path": value_only / value, "same_bits_as_gradient_path": value_only_same_bits,
                                                    "executed_gemm_flops_per_eval": sum(g["work"] for g in gv) / B,
                                                    "kernels": {k: {"ms": round(v["ms"], 4), "launches": v["launches"], "share": round(v["ms"] / val_tot, 4)}
                                                                for k, v in prof_val.items() if v["launches"]}},
                         "single_point_call_us": {"value": single_us, "unit": "us per emub_predict_few call (host pointers, 1 point)"},
                         "strong": strong, "sharded_identical": (all(identical.values()) if identical else None),
                         "sharded_identical_detail": identical}}
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
