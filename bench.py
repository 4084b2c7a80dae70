#!/usr/bin/env python
"""bench.py -- the reference's headline hot path on B200: GP log-likelihood + theta-gradient evaluations per
second at n=4096, d=10, FP64 (BASELINE.json), plus emulated points per second on the same model.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

A "step" is one batched call of the hot path: B theta points (one optimizer-restart front) -> (-L, gradient,
sigma^2) each, i.e. B x evalFnGradMulti (reference src/libEmu/maxmultimin.c:615).  Multi-GPU: one process per
GPU (torchrun), every rank evaluates its own B restarts of the same model (independent units, no data-path
collective), `value` = all ranks' evaluations / max-over-ranks device time.  Prints ONE JSON line.
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
CPU_SAMPLE_N = 1024
# dram__bytes_read.sum + dram__bytes_write.sum over the 125 k_gemm launches of ONE likelihood+gradient batch of 8
# matrices at n=4096 (ncu, profiles/r01_gemm_family_traffic.txt); the matrices of a batch are independent, so the
# traffic of a batch of B is B/8 of this
GEMM_FAMILY_DRAM_BYTES_B8, GEMM_FAMILY_LAUNCHES = 5.348e9, 125


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="theta points per step per GPU (BASELINE cfg3: 64 restarts batched)")
    ap.add_argument("--pred-points", type=int, default=1 << 18, help="query points per prediction step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


def cpu_reference_rate(nthreads, reps=1):
    """The reference's own evalFnGradMulti (oracle/_ref, compiled from the reference sources; falls back to the
    plain-C port) on a bounded sample: n=1024 (one evaluation is ~13 s on one core), one independent evaluation per
    thread -- the reference's own parallel model (estimate_threaded.c:97,172) -- extrapolated to n=4096 by the
    cubic cost of the path ((2(T-1)+1) n^3 flops)."""
    from madaiemulator_b200 import datasets as ds
    from oracle import pyoracle as po
    X = ds.synthetic_design(CPU_SAMPLE_N, D_MODEL)
    y = ds.synthetic_response(X)
    th = ds.default_theta_less_amp(D_MODEL)
    scale = (CPU_SAMPLE_N / float(N_MODEL)) ** 3
    if po.ref_available():
        t = po.time_ref_eval_grad(X, y, th, 1, ORDER, nthreads=nthreads, reps=reps)
        kind = "reference"
    else:
        o = po.PortOracle(X, y, 1, ORDER)
        t0 = time.perf_counter()
        for _ in range(reps):
            o.loglik_grad(th)
        t = time.perf_counter() - t0
        nthreads = 1
        kind = "port"
    evals = nthreads * reps
    rate_sample = evals / t
    return dict(value=rate_sample * scale, unit="evals/s", cores=nthreads, kind=kind,
                sample="evalFnGradMulti at n=%d,d=%d: %d evals in %.2f s on %d threads (%.3f evals/s), extrapolated to "
                       "n=%d by (n/%d)^3" % (CPU_SAMPLE_N, D_MODEL, evals, t, nthreads, rate_sample, N_MODEL, CPU_SAMPLE_N),
                seconds=t)


def run_reference(args, rank, world):
    if rank != 0:
        return
    ncpu = os.cpu_count() or 1
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_rate(ncpu)
    vals, secs = [], 0.0
    last = None
    for _ in range(args.steps):
        last = cpu_reference_rate(ncpu)
        vals.append(last["value"])
        secs += last["seconds"]
    v = float(np.mean(vals))
    cb = dict(last)
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

    def step_dev():
        model.loglik_grad_batch_dev(d_thetas.data_ptr(), B, True, d_out.data_ptr())

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
    if rank == 0:
        # per-kernel-family durations: one more step with every launch bracketed by CUDA events (one stream group)
        ctx.profile(True)
        step_dev()
        ctx.synchronize()
        prof = ctx.profile_read()
        emu.emulate_dev(d_pts.data_ptr(), min(mq, 16384), d_mean.data_ptr(), d_var.data_ptr())
        ctx.synchronize()
        prof_pred = ctx.profile_read()
        ctx.profile(False)
        gem = [prof[k] for k in ("gemm_chol", "gemm_trtri", "gemm_lauum")]
        g_ms = sum(g["ms"] for g in gem)
        g_fl = sum(g["work"] for g in gem)
        g_n = sum(g["launches"] for g in gem)
        tot_ms = sum(v["ms"] for v in prof.values())
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
        # algorithmic work of the GEMM-shaped part of one evaluation: n^3/3 (Cholesky) + 2n^3/3 (inverse) = n^3 flops
        # (SURVEY 8d); the tile engine executes ~7% more (diagonal tiles are computed at 64 x 64 granularity)
        achieved = B * float(N_MODEL) ** 3 / (g_ms * 1e-3) / 1e12
        executed = g_fl / (g_ms * 1e-3) / 1e12
        kernels = {k: {"ms": round(v["ms"], 4), "launches": v["launches"],
                       "share": round(v["ms"] / tot_ms, 4)} for k, v in prof.items() if v["launches"]}
        roofline = {"bound": "tensor", "kernel": "emub::k_gemm (FP64 DMMA tile engine: Cholesky TRSM/SYRK, inverse merge, W^T W)",
                    "achieved": round(achieved, 3), "peak": round(peak, 3), "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
                    "traffic": GEMM_FAMILY_DRAM_BYTES_B8 * (B / 8.0) / GEMM_FAMILY_LAUNCHES if N_MODEL == 4096 else None,
                    "traffic_source": "ncu dram bytes summed over the 125 k_gemm launches of one batch of 8 (profiles/"
                                      "r01_gemm_family_traffic.txt), scaled by B/8, per launch on average; 669 MB per evaluation "
                                      "for ~10 passes over 67 MB lower-triangular matrices",
                    "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry; "
                                   "DMMA issue-rate microbenchmark 37.1 TFLOP/s, profiles/r01_dmma_probe.txt)",
                    "launches": g_n, "avg_launch_ms": round(g_ms / max(1, g_n), 4),
                    "algorithmic_flops_per_step": B * float(N_MODEL) ** 3, "executed_flops_per_step": g_fl,
                    "executed_tflops": round(executed, 3), "step_share": round(g_ms / tot_ms, 4),
                    "eval_flops": float(N_MODEL) ** 3,
                    "eval_tflops_timed_region": round(value / world * float(N_MODEL) ** 3 / 1e12, 3)}
        pg = prof_pred["gemm_pred"]
        pred_roof = {"kernel": "emub::k_gemm<W K, column sum of squares>", "achieved": round(pg["work"] / (pg["ms"] * 1e-3) / 1e12, 3),
                     "peak": round(peak, 3), "unit": "TFLOP/s"} if pg["ms"] > 0 else None
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_reference_rate(os.cpu_count() or 1)
            cpu.pop("seconds", None)
        out = {"metric": "loglik_grad_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
               "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(B * nth1 * 8),
                       "d2h_bytes_per_step": int(B * 90 * 8)},
               "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
               "failed_points": nfail, "kernels": kernels,
               "extra": {"emulated_points_per_s": {"value": pred_value, "unit": "points/s", "points_per_step_per_gpu": mq,
                                                   "steps": psteps, "ms_per_step": pms / psteps,
                                                   "e2e": {"value": pred_e2e, "unit": "points/s",
                                                           "h2d_bytes_per_step": int(mq * D_MODEL * 8), "d2h_bytes_per_step": int(mq * 16)},
                                                   "roofline": pred_roof,
                                                   "algorithmic_flops_per_point": float(N_MODEL) ** 2 + 2 * N_MODEL + 3 * N_MODEL * D_MODEL},
                         "single_point_call_us": {"value": single_us, "unit": "us per emub_predict_few call (host pointers, 1 point)"}}}
    emu.close()
    model.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
