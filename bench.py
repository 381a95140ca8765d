#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: pair-interactions per second of
the O(N^2) force + Euler + kinetic-energy step (BASELINE.json), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs): N GPUs = 1 -> C2, N = 1,048,576 uniform cube (the
single-GPU FP32-roofline headline); N GPUs > 1 -> C3, N = 4,194,304 Plummer sphere, i-sharded
STRONG scaling.  A "step" is one full time step: N^2 pair evaluations, the Euler update and the
kinetic-energy reduction, fused in one kernel launch per GPU -- preceded, from 65 536 bodies on, by the 10 us
launch that rewrites the j-records for the q-scaled pair (`gpu_launches_detail` counts both).

Besides the headline the line carries (all outside the headline's timed region):
  parity        correctness of what was just timed: kinetic energies and sampled positions against the
                committed output of the reference itself (tests/golden/large_*_ver8.npz, where the workload
                has one), kinetic energy against an fp64 sum, sampled forces against the oracle's fp64 and
                float (ver2-order) sums, and -- several GPUs -- every replica bit-equal
  strong_anchor (--gpus 1) C3 timed on this one GPU: T1 of the strong-scaling curve; N > 1 lines report
                strong_efficiency = T1 / (N * T_N) with T1 taken from the same box
  also          (--gpus 1) C1 and C0, the small-N configs; (--gpus 8) C4, the 16 M weak-scaling config
  exchange_ab   (N > 1) the position-exchange modes on the headline workload, 2 steps each (each entry names its kernel shape)
  extras_error  only if one of the blocks above failed: the headline, e2e and parity were measured before them and stand
  gpu_reference (--gpus 1) the reference's own CUDA backend on the same GPU, kernel-only and end-to-end,
                with this build timed at the same N beside it

--impl reference times the reference's own CPU implementation (oracle/_ref ver8: OpenMP +
SIMD + i-tiling, compiled from the unmodified sources) on this box's host cores, on a
bounded sample of the same workload (smaller N; the rate at three sizes is reported with it).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import socket
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

FLOP_PER_PAIR = 20.0          # 3 sub + 6 dist2 + 4 rsqrt.cube + 1 mass + 6 accumulate
FP32_LANES_PER_SM = 128
UNIT = "G pair-interactions/s"

WORKLOADS = {
    "c0": dict(n=2000, ic="uniform", fixture=None, name="C0: N=2000 uniform cube (./nbody.x 2000 500)"),
    "c1": dict(n=16384, ic="uniform", fixture="c1", name="C1: N=16384 uniform cube (j-split small-N case)"),
    "c2": dict(n=1 << 20, ic="uniform", fixture="c2", name="C2: N=1,048,576 uniform cube, reference ICs (mt19937(42))"),
    "c3": dict(n=1 << 22, ic="plummer", fixture="c3", name="C3: N=4,194,304 Plummer sphere, i-sharded"),
    "c4": dict(n=1 << 24, ic="uniform", fixture=None, name="C4: N=16,777,216 uniform cube, i-sharded"),
}
XCH_NAMES = {0: "nccl", 1: "p2p", 2: "nccl_overlap"}
ANCHOR_CACHE = os.path.join(REPO, "gpurun_out", "strong_anchor.json")     # T1(C3) measured by the --gpus 1 run on this box


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


def cpu_rate_by_size(budget_s: float = 6.0):
    """ver8's rate is NOT flat in N on these hosts (cache blocking, thread start-up): report it at
    three sizes; the largest rate is the one the ratio is quoted against (most favourable to the CPU)."""
    rates = {}
    for n_s in (65536, 131072, 262144):
        est = float(n_s) ** 2 / 30e9
        steps = int(max(1, min(8, budget_s / 3 / max(est, 1e-3))))
        r, kind, cores, _ = cpu_reference_rate(n_s, steps)
        rates[str(n_s)] = round(r, 2)
    return rates


def cpu_baseline_block(budget_s: float = 12.0):
    rates = cpu_rate_by_size()
    n_s = int(max(rates, key=lambda k: rates[k]))
    rate0 = rates[str(n_s)]
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
    lo, hi = min(rates.values()), max(rates.values())
    return {"value": round(rate, 3), "unit": UNIT, "cores": cores, "kind": kind,
            "gflops_ref_convention": round(rate * 29.0, 1), "rate_by_n": rates, "other_reference_versions": others,
            "sample": f"{'oracle/_ref ver8 (OpenMP+SIMD+i-tiling, unmodified reference sources)' if kind == 'reference' else 'oracle port of ver7'}"
                      f", N={n_s} (the size with the best rate of 65536/131072/262144: {lo}..{hi} G pairs/s, i.e. the CPU rate depends on N), "
                      f"{steps} steps, {secs:.1f} s step-loop time, OMP_NUM_THREADS={cores}"}


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    rates = cpu_rate_by_size()
    n_s = int(max(rates, key=lambda k: rates[k]))
    kind, cores = "reference", os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_reference_rate(n_s, 1)
    t = 0.0
    for _ in range(args.steps):
        _, kind, cores, secs = cpu_reference_rate(n_s, 1)
        t += secs
    value = float(n_s) ** 2 * args.steps / t / 1e9
    lo, hi = min(rates.values()), max(rates.values())
    sample = (f"oracle/_ref ver8, N={n_s} per timed step (bounded sample of the workload: a CPU step at full N takes minutes to hours); "
              f"rate at N=65536/131072/262144 = {rates['65536']}/{rates['131072']}/{rates['262144']} G pairs/s (spread {lo}..{hi}: "
              f"not flat in N), the best of the three sizes is used; OMP_NUM_THREADS={cores}")
    line = {
        "impl": "reference", "metric": "pair_interactions_per_second", "value": round(value, 3),
        "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * t / args.steps, 3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "sample": f"N={n_s} per step", "rate_by_n": rates},
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "rate_by_n": rates},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gflops_ref_convention": round(value * 29.0, 1),
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------
#  native arm
# ---------------------------------------------------------------------------------
class Bench:
    """One process = one GPU.  Holds the plumbing shared by the headline run and the extra blocks."""

    def __init__(self, args):
        import torch
        self.torch = torch
        pkg = importlib.import_module("nbody-demo-2023_b200")
        self.nbx, self.dist = pkg.nbx, importlib.import_module("nbody-demo-2023_b200.dist")
        if not torch.cuda.is_available() or self.nbx.device_count() == 0:
            raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
        self.rank, self.local_rank, self.world = self.dist.init("nccl")
        if self.world != args.gpus:
            raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={self.world}; launch with torch.distributed.run")
        torch.cuda.set_device(self.local_rank)
        self.args = args
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
        self.peaks, self.peak_kind = measured_peaks()
        self.host_cache = {}

    # ---- inputs: pinned host buffers (nbx_host_alloc) filled with the workload's initial conditions
    def host_arrays(self, key):
        if key not in self.host_cache:
            wl = WORKLOADS[key]
            host = [self.nbx.pinned_empty(wl["n"]) for _ in range(7)]
            for h, a in zip(host, self.nbx.ic(wl["n"], wl["ic"])):
                h[:] = a
            self.host_cache = {key: host}          # keep one workload's buffers at a time
        return self.host_cache[key]

    def make_ctx(self, key, exchange=None, variant=-1, j_splits=0, multicast=-1):
        nbx = self.nbx
        ctx = self.dist.make_sharded_context(nbx, WORKLOADS[key]["n"], exchange, device=self.local_rank, multicast=multicast)
        if variant >= 0:
            ctx.set_option("variant", variant)
        if j_splits > 0:
            ctx.set_option("j_splits", j_splits)
        return ctx

    def upload(self, ctx, host):
        if self.world > 1:
            ctx.upload_sharded(*host)      # collective: own shard over PCIe, packed records over NVLink
        else:
            ctx.upload(*host)

    def timed_steps(self, ctx, steps, warmup, sampler=None):
        """`warmup` untimed + `steps` timed single-step runs, L2 flushed before each; device seconds are the
        CUDA-event time inside nbx_run, max over ranks."""
        torch, dist = self.torch, self.dist

        def one_step():
            self.flush.zero_()                 # L2 flush between timed iterations
            torch.cuda.synchronize()
            ke, secs = ctx.run(1)
            return ke[0], secs

        for _ in range(warmup):
            one_step()
        dist.barrier(); torch.cuda.synchronize()
        if sampler:
            sampler.start()
        i0 = ctx.info()
        t0 = time.perf_counter()
        kernel_s, ke_last = 0.0, 0.0
        for _ in range(steps):
            ke_last, secs = one_step()
            kernel_s += secs
        torch.cuda.synchronize(); dist.barrier()
        wall = dist.reduce_scalar(time.perf_counter() - t0, "max")
        kernel_s = dist.reduce_scalar(kernel_s, "max")
        i1 = ctx.info()      # launches in the timed region: the step kernel, and (q-scaled shapes) the record rewrite before it
        self.timed_launches_detail = {"step_kernel": int(i1["kernel_launches"] - i0["kernel_launches"]),
                                      "qscale_kernel": int(i1["aux_launches"] - i0["aux_launches"])}
        self.timed_launches = sum(self.timed_launches_detail.values())
        return kernel_s, wall, ke_last

    # ---- parity of one workload on this set of GPUs (never inside a timed region)
    def parity(self, key, ctx, host):
        """Fresh run from the initial conditions, checked four ways; returns the JSON block."""
        from oracle import oracle as O      # the checker (tests/bench only)
        nbx, dist, world = self.nbx, self.dist, self.world
        wl = WORKLOADS[key]
        n = wl["n"]
        out = {"ok": True, "gates": {"kenergy_rel": 1e-4, "pos_rel_l2": 1e-4}}
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))
        self.upload(ctx, host)
        info = ctx.info()
        lo, hi = min(info["i_begin"], n), min(info["i_begin"] + info["i_count"], n)
        st0 = O.State(n)
        for f, a in zip(O.State.FIELDS, host):
            setattr(st0, f, np.asarray(a))

        # (1) sampled accelerations of this rank's shard: fp64 truth and the reference's float order
        nsel = max(32, 256 // world)
        sel = np.sort(np.random.default_rng(1234 + self.rank).choice(np.arange(lo, hi), min(nsel, hi - lo), replace=False)).astype(np.int32)
        acc = ctx.accelerations()[sel - info["i_begin"]]
        truth = O.acc_fp64(st0, sel)
        ref32 = O.acc_f32(st0, sel)
        tn = np.linalg.norm(truth, axis=1)
        e_gpu = np.linalg.norm(acc - truth, axis=1) / tn
        e_ref = np.linalg.norm(ref32 - truth, axis=1) / tn
        acc_blk = {"samples": int(dist.reduce_scalar(float(sel.size), "sum")),
                   "gpu_vs_fp64_max": dist.reduce_scalar(float(e_gpu.max()), "max"),
                   "reference_float_vs_fp64_max": dist.reduce_scalar(float(e_ref.max()), "max"),
                   "gpu_vs_reference_float_max": dist.reduce_scalar(float((np.linalg.norm(acc - ref32, axis=1) / tn).max()), "max")}
        # bar: as close to the fp64 force as the reference's own float arithmetic (1e-4 where that is tighter)
        acc_blk["ok"] = bool(acc_blk["gpu_vs_fp64_max"] < max(1e-4, 1.5 * acc_blk["reference_float_vs_fp64_max"]))
        out["sampled_forces"] = {k: (float(f"{v:.3e}") if isinstance(v, float) else v) for k, v in acc_blk.items()}

        # (2) steps from the ICs: against the reference's own output AND the fp64 truth where fixtures exist
        # (tests/golden/large_*_ver8.npz: unmodified reference ver8; truth_*_fp64.npz: oracle_run_fp64).
        # Gates as in tests/test_gpu_headline.py: 1e-4 against the truth; against the reference's output
        # 1e-4 + the reference's own distance from the truth (its float sums are biased low from N ~ 1 M on);
        # a 500-step run is chaotic: no further from the truth than twice the reference's own distance.
        fx = tr = None
        if wl["fixture"]:
            gdir = os.path.join(REPO, "tests", "golden")
            path, tpath = os.path.join(gdir, f"large_{wl['fixture']}_ver8.npz"), os.path.join(gdir, f"truth_{wl['fixture']}_fp64.npz")
            if os.path.exists(path) and os.path.exists(tpath):
                fx, tr = np.load(path), np.load(tpath)
        steps = int(fx["steps"]) if fx is not None else 1
        ke, _ = ctx.run(steps)
        full = [np.zeros(n, dtype=np.float32) for _ in range(6)]
        ctx.download(*full)                 # all positions + this shard's velocities
        if fx is not None:
            s_all = fx["sel"]
            mine = (s_all >= lo) & (s_all < hi)
            pos = np.stack([a[s_all] for a in full[:3]], axis=1).astype(np.float64)
            vel = np.stack([a[s_all[mine]] for a in full[3:6]], axis=1).astype(np.float64)

            def dev(ke_a, pos_a, vel_a, ke_b, pos_b, vel_b):
                dv2 = dist.reduce_scalar(float(np.sum((vel_a - vel_b[mine]) ** 2)), "sum")
                return (float(np.max(np.abs(ke_a - ke_b) / ke_b)), float(np.linalg.norm(pos_a - pos_b) / np.linalg.norm(pos_b)),
                        float(np.sqrt(dv2 / np.sum(vel_b.astype(np.float64) ** 2))))
            rk, rp, rv = fx["ke"].astype(np.float64), fx["pos_sel"].astype(np.float64), fx["vel_sel"].astype(np.float64)
            tk, tp, tv = tr["ke"], tr["pos_sel"], tr["vel_sel"]
            g_t, g_r = dev(ke, pos, vel, tk, tp, tv), dev(ke, pos, vel, rk, rp, rv)
            r_t = (float(np.max(np.abs(rk - tk) / tk)), float(np.linalg.norm(rp - tp) / np.linalg.norm(tp)), float(np.linalg.norm(rv - tv) / np.linalg.norm(tv)))
            chaotic = steps > 100
            ok = all((g < max(1e-4, 2 * r) if chaotic else g < 1e-4) and gr < 1e-4 + 1.05 * (r + (g if chaotic else 0.0))
                     for g, r, gr in zip(g_t, r_t, g_r))
            f3 = lambda t: {k: float(f"{v:.3e}") for k, v in zip(("kenergy_max_rel", "pos_rel_l2_sampled", "vel_rel_l2_sampled"), t)}
            out["vs_reference_output"] = {
                "source": f"tests/golden/large_{wl['fixture']}_ver8.npz (unmodified reference ver8) and truth_{wl['fixture']}_fp64.npz (all-double run), "
                          f"{steps} steps from the same ICs, 4096 sampled bodies",
                "steps": steps, "gpu_vs_fp64_truth": f3(g_t), "reference_vs_fp64_truth": f3(r_t), "gpu_vs_reference": f3(g_r),
                "gate": "gpu_vs_truth < 1e-4; gpu_vs_reference < 1e-4 + reference_vs_truth" + ("; 500-step run is chaotic: gpu_vs_truth < 2 x reference_vs_truth" if chaotic else ""),
                "ok": bool(ok)}
        else:
            out["vs_reference_output"] = None

        # (3) kinetic energy against an fp64 sum over the downloaded velocities (all shards)
        st = O.State(hi - lo)
        st.vx, st.vy, st.vz, st.mass = (np.ascontiguousarray(full[3][lo:hi]), np.ascontiguousarray(full[4][lo:hi]),
                                        np.ascontiguousarray(full[5][lo:hi]), np.ascontiguousarray(np.asarray(host[6])[lo:hi]))
        ke64 = dist.reduce_scalar(O.kenergy_fp64(st) if hi > lo else 0.0, "sum")
        out["kenergy_vs_fp64_sum"] = float(f"{abs(ke[-1] - ke64) / ke64:.3e}")
        ke_ok = out["kenergy_vs_fp64_sum"] < 1e-6

        # (4) every replica holds the same positions, bit for bit
        if world > 1:
            crc = float(zlib.crc32(b"".join(a.tobytes() for a in full[:3])))
            same = dist.reduce_scalar(crc, "max") == dist.reduce_scalar(crc, "min")
            out["replicas_bit_equal"] = bool(same)
        else:
            same = True
        out["ok"] = bool(acc_blk["ok"] and ke_ok and same and (out["vs_reference_output"] is None or out["vs_reference_output"]["ok"]))
        return out

    def rate_block(self, n, kernel_s, steps):
        value = float(n) * float(n) * steps / kernel_s / 1e9
        sm = self.sm_count
        peak = sm * FP32_LANES_PER_SM * 2 * self.peaks["sm_max_mhz"] * 1e6 / 1e12 * self.world
        return value, FLOP_PER_PAIR * value / 1e3, peak

    # ---- a secondary workload: timing + parity, as one JSON block
    def side_workload(self, key, steps, warmup, chunk=1, exchange=None):
        """`chunk` steps per nbx_run call (small N: the CUDA-graph replay path needs several steps per call)."""
        wl = WORKLOADS[key]
        host = self.host_arrays(key)
        ctx = self.make_ctx(key, exchange)
        try:
            self.upload(ctx, host)
            if chunk == 1:
                kernel_s, _, _ = self.timed_steps(ctx, steps, warmup)
                nsteps = steps
            else:
                for _ in range(warmup):
                    ctx.run(chunk)
                self.torch.cuda.synchronize()
                kernel_s = 0.0
                for _ in range(steps):
                    _, secs = ctx.run(chunk)
                    kernel_s += secs
                nsteps = steps * chunk
            kernel_s = self.dist.reduce_scalar(kernel_s, "max")
            info = ctx.info()
            value, tf, peak = self.rate_block(wl["n"], kernel_s, nsteps)
            blk = {"workload": wl["name"], "n_bodies": wl["n"], "steps": nsteps, "ms_per_step": round(1e3 * kernel_s / nsteps, 5),
                   "value": round(value, 2), "unit": UNIT, "tflops20": round(tf, 2), "frac_fp32_peak": round(tf / peak, 4),
                   "kernel_shape": self.nbx.variant_names()[info["variant"]], "i_tiles": info["i_tiles"], "whole_tiles": info["whole_tiles"],
                   "j_splits": info["j_splits"], "graph": info["use_graph"]}
            blk["parity"] = self.parity(key, ctx, host)
        finally:
            ctx.close()
        return blk


def gpu_reference_block(bench):
    """The reference's own (naive) CUDA backend on this GPU, kernel-only and end-to-end, with this build
    timed at the same N beside it: not the reference arm (that is the CPU path), never part of the product."""
    from oracle import oracle as O
    if not O.ref_cuda_available():
        return None
    n_s = 131072
    out = {"n_bodies": n_s, "unit": UNIT, "kind": "reference-cuda"}
    rate, ke = O.ref_cuda_rate(n_s, 150)
    out["end_to_end"] = {"value": round(rate, 2),
                         "sample": "oracle/_ref/ver5_all_cuda (cuda/Compute.cu unmodified, rebuilt -arch sm_100a, block 1024), 150 steps, its own "
                                   "timer over windows 2-3: per-step H2D + kernel + D2H + host Euler/energy (cuda/Compute.cu:150-194)",
                         "kenergy_column": ke}
    out["value"] = out["end_to_end"]["value"]
    ko = os.path.join(O.REF_DIR, "ref_cuda_kernel_only")
    if os.path.exists(ko):
        r = subprocess.run([ko, str(n_s), "20"], capture_output=True, text=True, check=True)
        d = json.loads(r.stdout.strip().splitlines()[-1])
        out["kernel_only"] = {"value": round(d["gpairs_per_s"], 2), "ms_per_launch": round(d["ms_per_launch"], 4),
                              "sample": "oracle/_ref/ref_cuda_kernel_only: CUDA events around nbody<<<n/1024,1024>>> alone "
                                        "(cuda/Compute.cu:31-66,159-162), 20 launches"}
    return out


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
    ap.add_argument("--no-extras", action="store_true", help="headline + parity only (no anchor / also / exchange_ab / comparators)")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    wl_key = args.workload if args.workload != "auto" else ("c2" if args.gpus == 1 else "c3")
    wl = WORKLOADS[wl_key]
    if args.impl == "reference":
        return run_reference_arm(args, wl)

    B = Bench(args)
    nbx, dist, torch = B.nbx, B.dist, B.torch
    rank, world = B.rank, B.world
    n = wl["n"]
    host = B.host_arrays(wl_key)
    out = [nbx.pinned_empty(n) for _ in range(6)]
    exchange = {"p2p": nbx.EXCHANGE_P2P, "nccl": nbx.EXCHANGE_NCCL, "nccl_overlap": nbx.EXCHANGE_NCCL_OVERLAP}[args.exchange]
    ctx = B.make_ctx(wl_key, exchange, args.variant, args.j_splits)
    exchange_used, multicast_used = (ctx.exchange_used if world > 1 else exchange), ctx.multicast
    B.upload(ctx, host)
    B.sm_count = ctx.info()["sm_count"]

    # ---- headline: W warm-up + K timed steps
    sampler = ClockSampler(B.local_rank)
    kernel_s, wall, ke_last = B.timed_steps(ctx, args.steps, args.warmup, sampler)
    clocks = sampler.stop()
    info1 = ctx.info()
    launches, launches_detail = B.timed_launches, dict(B.timed_launches_detail)
    pairs_per_step = float(n) * float(n)
    value, achieved_tflops, peak_tflops = B.rate_block(n, kernel_s, args.steps)

    # ---- e2e: the same step through the C ABI with HOST buffers: H2D of the step's inputs from pinned
    # memory, one step, D2H of the updated state + kinetic energy, every step.  Several GPUs: each rank
    # moves its own shard over PCIe (nbx_upload_sharded / nbx_download_shard).
    e2e = None
    if not args.no_e2e:
        def e2e_step():
            B.upload(ctx, host)
            ke, _ = ctx.run(1)
            if world > 1:
                ctx.download_shard(*out)
            else:
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
        e2e = {"value": round(pairs_per_step * e2e_steps / te / 1e9, 3), "unit": UNIT, "steps": e2e_steps,
               "h2d_bytes_per_step": 7 * 4 * n, "d2h_bytes_per_step": 6 * 4 * n + 8 * world,
               "ms_per_step": round(1e3 * te / e2e_steps, 3),
               "what": ("nbx_upload_sharded (each rank: its shard of the 7 pinned host SoA arrays; packed records all-gathered over NVLink) + "
                        "nbx_run(1) + nbx_download_shard per step; bytes are totals over the ranks" if world > 1 else
                        "nbx_upload(7 host SoA arrays, pinned) + nbx_run(1) + nbx_download(pos, vel) per step")}

    # ---- parity of the headline workload on this set of GPUs
    parity = None if args.no_parity else B.parity(wl_key, ctx, host)

    exchange_ab = None
    # Everything below is reporting around the headline (other exchange modes, the strong-scaling anchor, side workloads):
    # a failure there must not cost the line that was just measured.
    strong, also, extras_error = None, {}, None
    try:
        # ---- strong-scaling anchor: T1 on C3 (measured by the 1-GPU run; re-used by N > 1 runs on the same box)
        if not args.no_extras:
            if ctx is not None:
                ctx.close(); ctx = None
            try:       # the anchor is only valid for the GPU it was measured on (boxes share a hostname, GPUs differ by ~1.5 %)
                box = socket.gethostname() + "/" + str(torch.cuda.get_device_properties(0).uuid)
            except Exception:
                box = socket.gethostname()
            if world == 1:
                blk = B.side_workload("c3", 3, 1)
                strong = {"workload": WORKLOADS["c3"]["name"], "ms_per_step": blk["ms_per_step"], "value": blk["value"], "steps": 3,
                          "frac_fp32_peak": blk["frac_fp32_peak"], "parity": blk["parity"], "box": box}
                try:
                    os.makedirs(os.path.dirname(ANCHOR_CACHE), exist_ok=True)
                    with open(ANCHOR_CACHE, "w") as f:
                        json.dump(strong, f)
                except OSError:
                    pass
                also["c1"] = B.side_workload("c1", 5, 2, chunk=500)
                also["c0"] = B.side_workload("c0", 5, 2, chunk=500)
            elif wl_key == "c3":
                anchor = None
                if os.path.exists(ANCHOR_CACHE):
                    try:
                        a = json.load(open(ANCHOR_CACHE))
                        if a.get("box") == box and time.time() - os.path.getmtime(ANCHOR_CACHE) < 6 * 3600:
                            anchor = dict(a, source="the --gpus 1 run on this box (cached in " + ANCHOR_CACHE + ")")
                    except Exception:
                        anchor = None
                if anchor is None:
                    # no 1-GPU run on this box yet: rank 0 times C3 alone (1 warm-up + 2 steps) while the other GPUs idle
                    t1 = [0.0]
                    if rank == 0:
                        with nbx.Context(n, device=B.local_rank) as c1:
                            c1.upload(*host)
                            c1.run(1)
                            s = 0.0
                            for _ in range(2):
                                B.flush.zero_(); torch.cuda.synchronize()
                                s += c1.run(1)[1]
                            t1[0] = 1e3 * s / 2
                    dist.barrier()
                    ms1 = dist.reduce_scalar(t1[0], "max")
                    anchor = {"workload": WORKLOADS["c3"]["name"], "ms_per_step": round(ms1, 4), "value": round(pairs_per_step / ms1 / 1e6, 2),
                              "steps": 2, "box": box, "source": "timed inside this run on rank 0's GPU alone"}
                strong = anchor
            if world == 8:
                also["c4"] = B.side_workload("c4", 2, 1)

        # ---- the exchange modes, same workload, same box (N > 1)
        if world > 1 and not args.no_extras:
            exchange_ab = {}
            if ctx is not None:
                ctx.close(); ctx = None
            for name, mode in (("p2p", nbx.EXCHANGE_P2P), ("p2p_unicast", nbx.EXCHANGE_P2P), ("nccl", nbx.EXCHANGE_NCCL),
                               ("nccl_overlap", nbx.EXCHANGE_NCCL_OVERLAP)):
                c2 = B.make_ctx(wl_key, mode, multicast=0 if name == "p2p_unicast" else -1)
                try:
                    B.upload(c2, host)
                    ks, _, _ = B.timed_steps(c2, 2, 1)
                    exchange_ab[name] = {"ms_per_step": round(1e3 * ks / 2, 3), "value": round(pairs_per_step * 2 / ks / 1e9, 1),
                                         "used": XCH_NAMES[c2.exchange_used], "multicast": c2.multicast,
                                         "kernel_shape": nbx.variant_names()[c2.info()["variant"]]}
                finally:
                    c2.close()

            # the one-process form of the same run (nbx_run_group: the CLI's path; its multicast team needs no descriptor
            # passing): rank 0 drives all the GPUs while the other ranks wait on a CPU barrier
            gloo = torch.distributed.new_group(backend="gloo")
            if rank == 0:
                try:
                    ctxs = [nbx.Context(n, device=g, rank=g, world=world) for g in range(world)]
                    try:
                        nbx.p2p_attach_group(ctxs)
                        nbx.upload_group(ctxs, *host)
                        nbx.run_group(ctxs, 1)
                        ks = sum(nbx.run_group(ctxs, 1)[1] for _ in range(2))
                        exchange_ab["p2p_one_process"] = {"ms_per_step": round(1e3 * ks / 2, 3), "value": round(pairs_per_step * 2 / ks / 1e9, 1),
                                                          "multicast": bool(ctxs[0].info()["multicast"]),
                                                          "what": "nbx_run_group from rank 0's process over all GPUs; multicast = the epilogue exchange is one "
                                                                  "multimem.st per record through a cuMulticast mapping instead of world-1 NVLink stores"}
                    finally:
                        for c in ctxs:
                            c.close()
                except Exception as ex:
                    exchange_ab["p2p_one_process"] = {"error": repr(ex)}
            torch.distributed.barrier(group=gloo)

    except Exception as ex:
        extras_error = repr(ex)
        print(f"[bench] rank {rank}: extras failed: {extras_error}", file=sys.stderr, flush=True)

    if rank != 0:
        if ctx is not None:
            ctx.close()
        if world > 1:
            torch.distributed.destroy_process_group()
        return 0

    sm_count = B.sm_count
    peaks, peak_kind = B.peaks, B.peak_kind
    traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))      # ncu-measured DRAM bytes per step, keyed by workload and kernel shape
            traffic = tj.get(f"{wl_key}:{nbx.variant_names()[info1['variant']]}")
        except Exception:
            traffic = None
    ms_per_step = 1e3 * kernel_s / args.steps
    line = {
        "metric": "pair_interactions_per_second", "value": round(value, 3), "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "n_bodies": n, "pairs_per_step": pairs_per_step,
                   "parallelism": f"i-shard x{args.gpus}" + (f", exchange={XCH_NAMES[exchange_used]}" + (" through NVSwitch multicast (one multimem.st per record)" if multicast_used else "") if args.gpus > 1 else ""),
                   "kernel_shape": nbx.variant_names()[info1["variant"]],
                   "i_tiles": info1["i_tiles"], "whole_tiles": info1["whole_tiles"], "j_splits": info1["j_splits"], "ctas_per_sm": info1["ctas_per_sm"],
                   "l2": "flushed between timed steps (256 MiB memset); positions (16 B/body) are L2-resident by design within a step",
                   "timing": "CUDA events around each step on the launching stream (inside nbx_run), max over ranks; wall clock alongside",
                   "strong_anchor": strong},
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
                     "note": ("compute-bound on the FP32 pipe, not HBM or tensor: q-scaled pair = 11 FP32 lane-ops per pair, 9 of them FMAs, so 20 algorithmic flop/pair caps at "
                              "20/22 = 90.9% of the FMA peak (register-bank reads: 25 cycles where the pipe needs 22); HBM need is 64 B/body/step + 40 B/body for the record rewrite"
                              if nbx.variant_names()[info1["variant"]].endswith("_qi") else
                              "compute-bound on the FP32 pipe, not HBM or tensor: 12 FP32 lane-ops per pair, 6 of them FMAs, so 20 algorithmic flop/pair caps at 20/24 = 83.3% of the FMA peak; HBM need is 64 B/body/step")},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "gpu_launches_detail": launches_detail, "parity": parity,
    }
    if strong and world > 1 and wl_key == "c3":
        line["strong_efficiency"] = round(strong["ms_per_step"] / (world * ms_per_step), 4)
    if exchange_ab:
        line["exchange_ab"] = exchange_ab
    if also:
        line["also"] = also
    if extras_error:
        line["extras_error"] = extras_error
    if args.gpus == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline_block()
        except Exception as ex:   # the baseline is a report, not the product
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(ex)}
    if args.gpus == 1 and not args.no_extras:
        try:
            if ctx is not None:
                ctx.close(); ctx = None      # free the GPU before the comparator runs on it
            gr = gpu_reference_block(B)
            if gr:
                blk = None
                # this build at the comparator's N, device-timed and end-to-end (host buffers), beside it
                n_s = gr["n_bodies"]
                arrs = nbx.ic(n_s)
                with nbx.Context(n_s, device=B.local_rank) as c:
                    c.upload(*arrs)
                    c.run(20)
                    _, secs = c.run(100)
                    t0 = time.perf_counter()
                    for _ in range(20):
                        c.upload(*arrs); c.run(1); st = c.state()
                    te = (time.perf_counter() - t0) / 20
                gr["this_build_same_n"] = {"kernel_only": round(float(n_s) ** 2 * 100 / secs / 1e9, 1),
                                           "end_to_end": round(float(n_s) ** 2 / te / 1e9, 1),
                                           "sample": "nbx_run(100) device time; and nbx_upload + nbx_run(1) + nbx_download per step (pageable host arrays)"}
            line["gpu_reference"] = gr
        except Exception as ex:
            line["gpu_reference"] = {"value": None, "kind": "unavailable", "sample": repr(ex)}
    print(json.dumps(line), flush=True)
    if ctx is not None:
        ctx.close()
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
