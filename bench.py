#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: pair-interactions per second of
the O(N^2) force + Euler + kinetic-energy step (BASELINE.json), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json configs): N GPUs = 1 -> C2, N = 1,048,576 uniform cube (the
single-GPU FP32-roofline headline); N GPUs > 1 -> C3, N = 4,194,304 Plummer sphere,
i-sharded strong scaling.  A "step" is one full time step: N^2 pair evaluations, the Euler
update and the kinetic-energy reduction, fused in one kernel launch per GPU.

--impl reference times the reference's own CPU implementation (oracle/_ref ver8: OpenMP +
SIMD + i-tiling, compiled from the unmodified sources) on this box's host cores, on a
bounded sample of the same workload (smaller N; pairs/s is size-independent for O(N^2)).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

FLOP_PER_PAIR = 20.0          # 3 sub + 6 dist2 + 4 rsqrt.cube + 1 mass + 6 accumulate
FP32_LANES_PER_SM = 128

WORKLOADS = {
    "c1": dict(n=16384, ic="uniform", name="C1: N=16384 uniform cube (j-split small-N case)"),
    "c2": dict(n=1 << 20, ic="uniform", name="C2: N=1,048,576 uniform cube, reference ICs (mt19937(42))"),
    "c3": dict(n=1 << 22, ic="plummer", name="C3: N=4,194,304 Plummer sphere, i-sharded"),
    "c4": dict(n=1 << 24, ic="uniform", name="C4: N=16,777,216 uniform cube, i-sharded"),
}


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------
#  reference arm / cpu baseline: the reference's own OpenMP+SIMD code on the host cores
# ---------------------------------------------------------------------------------
def cpu_reference_rate(n_sample: int, steps: int):
    """Run the compiled reference (ver8) -- or the oracle port when it is absent -- for
    `steps` steps at N = n_sample; returns (G pairs/s, kind, cores, seconds)."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    if O.ref_available("ver8"):
        _, _, secs = O.ref_run("ver8", n_sample, steps, threads=cores)
        kind = "reference"
    else:
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        s = O.ic_uniform(n_sample)
        t0 = time.perf_counter()
        O.run(s, steps, variant="ver7")
        secs = time.perf_counter() - t0
        kind = "port"
    return float(n_sample) ** 2 * steps / secs / 1e9, kind, cores, secs


def cpu_baseline_block(budget_s: float = 12.0):
    rate0, kind, cores, _ = cpu_reference_rate(32768, 1)            # calibrate (~0.1-1 s)
    n_s = 131072 if rate0 > 20 else 65536
    steps = int(max(1, min(40, round(budget_s * rate0 * 1e9 / float(n_s) ** 2))))
    rate, kind, cores, secs = cpu_reference_rate(n_s, steps)
    others = {}
    try:   # the north star's other two reported baselines, same box, same run (a few seconds in total)
        from oracle import oracle as O
        if O.ref_available("ver0"):
            _, _, s0 = O.ref_run("ver0", 2000, 50, threads=1)
            others["ver0_1_thread_n2000"] = round(2000.0 ** 2 * 50 / s0 / 1e9, 4)
        if O.ref_available("ver7"):
            _, _, s7 = O.ref_run("ver7", 16384, 20, threads=cores)
            others[f"ver7_{cores}_threads_n16384"] = round(16384.0 ** 2 * 20 / s7 / 1e9, 3)
    except Exception as ex:
        others["error"] = repr(ex)
    return {"value": round(rate, 3), "unit": "G pair-interactions/s", "cores": cores, "kind": kind,
            "gflops_ref_convention": round(rate * 29.0, 1), "other_reference_versions": others,
            "sample": f"{'oracle/_ref ver8 (OpenMP+SIMD+i-tiling, unmodified reference sources)' if kind == 'reference' else 'oracle port of ver7'}"
                      f", N={n_s}, {steps} steps, {secs:.1f} s step-loop time, OMP_NUM_THREADS={cores}; pairs/s is flat in N for O(N^2)"}


def gpu_reference_block():
    """The reference's own (naive) CUDA backend on this GPU, for context: not the reference arm
    (that is the CPU path) and never part of the product."""
    from oracle import oracle as O
    if not O.ref_cuda_available():
        return None
    n_s = 131072
    rate, ke = O.ref_cuda_rate(n_s, 150)
    return {"value": round(rate, 2), "unit": "G pair-interactions/s", "kind": "reference-cuda",
            "sample": f"oracle/_ref/ver5_all_cuda (cuda/Compute.cu:31-66 unmodified, rebuilt -arch sm_100a, block 1024), "
                      f"N={n_s}, 150 steps, its own timer over windows 2-3 (per-step H2D + kernel + D2H + host update)",
            "kenergy_column": ke}


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    rate0, kind, cores, _ = cpu_reference_rate(32768, 1)
    n_s = 131072 if rate0 > 20 else 65536
    for _ in range(args.warmup):
        cpu_reference_rate(n_s, 1)
    t = 0.0
    for _ in range(args.steps):
        _, kind, cores, secs = cpu_reference_rate(n_s, 1)
        t += secs
    value = float(n_s) ** 2 * args.steps / t / 1e9
    line = {
        "impl": "reference", "metric": "pair_interactions_per_second", "value": round(value, 3),
        "unit": "G pair-interactions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * t / args.steps, 3), "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "sample": f"N={n_s} per step (bounded sample of the workload; CPU step at full N takes minutes)"},
        "cpu_baseline": {"value": round(value, 3), "unit": "G pair-interactions/s", "cores": cores, "kind": kind,
                         "sample": f"oracle/_ref ver8, N={n_s}, one step per timed step, OMP_NUM_THREADS={cores}"},
        "e2e": {"value": round(value, 3), "unit": "G pair-interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gflops_ref_convention": round(value * 29.0, 1),
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------
#  native arm
# ---------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS))
    ap.add_argument("--exchange", default="p2p", choices=["nccl", "nccl_overlap", "p2p"])
    ap.add_argument("--variant", type=int, default=-1)
    ap.add_argument("--j-splits", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    wl_key = args.workload if args.workload != "auto" else ("c2" if args.gpus == 1 else "c3")
    wl = WORKLOADS[wl_key]
    if args.impl == "reference":
        return run_reference_arm(args, wl)

    import torch
    pkg = importlib.import_module("nbody-demo-2023_b200")
    nbx, dist = pkg.nbx, importlib.import_module("nbody-demo-2023_b200.dist")
    if not torch.cuda.is_available() or nbx.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    rank, local_rank, world = dist.init("nccl")
    if world != args.gpus:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)

    n = wl["n"]
    # inputs: pinned host buffers (nbx_host_alloc), filled with the workload's ICs
    host = [nbx.pinned_empty(n) for _ in range(7)]
    for h, a in zip(host, nbx.ic(n, wl["ic"])):
        h[:] = a
    out = [nbx.pinned_empty(n) for _ in range(6)]

    exchange = {"p2p": nbx.EXCHANGE_P2P, "nccl": nbx.EXCHANGE_NCCL, "nccl_overlap": nbx.EXCHANGE_NCCL_OVERLAP}[args.exchange]
    ctx = dist.make_sharded_context(nbx, n, exchange, device=local_rank)
    if args.variant >= 0:
        ctx.set_option("variant", args.variant)
    if args.j_splits > 0:
        ctx.set_option("j_splits", args.j_splits)
    ctx.upload(*host)
    dist.barrier()     # every replica is packed before any peer's epilogue may store into it

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def one_step():
        flush.zero_()                      # L2 flush between timed iterations
        torch.cuda.synchronize()
        ke, secs = ctx.run(1)
        return ke[0], secs

    for _ in range(args.warmup):
        one_step()

    sampler = ClockSampler(local_rank)
    info0 = ctx.info()
    dist.barrier(); torch.cuda.synchronize()
    sampler.start()
    t0 = time.perf_counter()
    kernel_s, ke_last = 0.0, 0.0
    for _ in range(args.steps):
        ke_last, secs = one_step()
        kernel_s += secs
    torch.cuda.synchronize(); dist.barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    info1 = ctx.info()

    wall = dist.reduce_scalar(wall, "max")
    kernel_s = dist.reduce_scalar(kernel_s, "max")       # device time (CUDA events in nbx_run), max over ranks
    pairs_per_step = float(n) * float(n)
    value = pairs_per_step * args.steps / kernel_s / 1e9
    launches = (info1["kernel_launches"] - info0["kernel_launches"])

    # ---- e2e: the same step through the C ABI with HOST buffers: H2D of the step's inputs
    # from pinned memory, one step, D2H of the updated state + kinetic energy, every step
    e2e = None
    if not args.no_e2e:
        def e2e_step():
            ctx.upload(*host)
            dist.barrier()
            ke, _ = ctx.run(1)
            ctx.download(*out)
            return ke[0]
        e2e_step()
        e2e_steps = min(args.steps, 5)         # same per-step work every time; keeps long runs bounded
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize(); dist.barrier()
        te = dist.reduce_scalar(time.perf_counter() - t0, "max")
        i_count = info1["i_count"]
        e2e = {"value": round(pairs_per_step * e2e_steps / te / 1e9, 3), "unit": "G pair-interactions/s", "steps": e2e_steps,
               "h2d_bytes_per_step": 7 * 4 * n, "d2h_bytes_per_step": 3 * 4 * n + 3 * 4 * min(i_count, n) + 8,
               "ms_per_step": round(1e3 * te / e2e_steps, 3),
               "what": "nbx_upload(7 host SoA arrays, pinned) + nbx_run(1) + nbx_download(pos, vel) per step, per rank"}

    if rank != 0:
        ctx.close()
        if world > 1:
            torch.distributed.destroy_process_group()
        return 0

    peaks, peak_kind = measured_peaks()
    sm_count = info1["sm_count"]
    peak_tflops = sm_count * FP32_LANES_PER_SM * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12 * args.gpus
    achieved_tflops = FLOP_PER_PAIR * value * 1e9 / 1e12
    traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(wl_key)
        except Exception:
            traffic = None
    line = {
        "metric": "pair_interactions_per_second", "value": round(value, 3), "unit": "G pair-interactions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * kernel_s / args.steps, 4), "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "n_bodies": n, "pairs_per_step": pairs_per_step,
                   "parallelism": f"i-shard x{args.gpus}" + (f", exchange={ {0: 'nccl', 1: 'p2p', 2: 'nccl_overlap'}[ctx.exchange_used] }" if args.gpus > 1 else ""),
                   "kernel_shape": nbx.variant_names()[info1["variant"]],
                   "i_tiles": info1["i_tiles"], "whole_tiles": info1["whole_tiles"], "j_splits": info1["j_splits"], "ctas_per_sm": info1["ctas_per_sm"],
                   "l2": "flushed between timed steps (256 MiB memset); positions (16 B/body) are L2-resident by design within a step",
                   "timing": "CUDA events around each step on the launching stream (inside nbx_run), max over ranks; wall clock alongside"},
        "gflops": round(FLOP_PER_PAIR * value, 1),
        "gflops_ref_convention": round(value * 29.0 + 19.0 * n * args.steps / kernel_s / 1e9, 1),
        "wall_ms_per_step": round(1e3 * wall / args.steps, 4),
        "kenergy_last": ke_last,
        "roofline": {"bound": "fp32", "achieved": round(achieved_tflops, 3), "peak": round(peak_tflops, 3), "unit": "TFLOP/s",
                     "frac": round(achieved_tflops / peak_tflops, 4), "traffic": traffic,
                     "peak_source": f"{sm_count} SMs x 128 FP32 lanes x 2 x sm_max_mhz {peaks['sm_max_mhz']} ({peak_kind} MEASURED_PEAKS.json) x {args.gpus} GPU",
                     "flop_per_pair": FLOP_PER_PAIR,
                     "hbm": {"algorithmic_bytes_per_launch": 64 * n // args.gpus, "achieved_gbs": round(64.0 * n / args.gpus / (kernel_s / args.steps) / 1e9, 2),
                             "peak_gbs": peaks.get("hbm_gbs"), "frac": round(64.0 * n / args.gpus / (kernel_s / args.steps) / 1e9 / peaks.get("hbm_gbs", 6650.0), 6)},
                     "note": "compute-bound on the FP32 pipe, not HBM or tensor: 12 FP32 lane-ops per pair, 6 of them FMAs, so 20 algorithmic flop/pair caps at 20/24 = 83.3% of the FMA peak; HBM need is 64 B/body/step"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
    }
    if args.gpus == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline_block()
        except Exception as ex:   # the baseline is a report, not the product
            line["cpu_baseline"] = {"value": None, "unit": "G pair-interactions/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(ex)}
        try:
            ctx.close()                      # free the GPU before the comparator runs on it
            line["gpu_reference"] = gpu_reference_block()
        except Exception as ex:
            line["gpu_reference"] = {"value": None, "kind": "unavailable", "sample": repr(ex)}
    print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
