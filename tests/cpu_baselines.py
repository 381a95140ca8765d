#!/usr/bin/env python
"""The reference's own CPU versions on this box's host cores (SURVEY.md 8d "CPU baseline timing"):
ver0 single-thread at C0, ver7 and ver8 on all cores at several N.  Uses oracle/_ref (the unmodified
reference compiled by oracle/Makefile).  Writes gpurun_out/cpu_baselines.json.
    python tests/cpu_baselines.py [budget_seconds_per_case]"""
import json, os, platform, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
cores = os.cpu_count() or 1
cpu = subprocess.run("lscpu | grep 'Model name' | head -1", shell=True, capture_output=True, text=True).stdout.strip()
rows = []


def case(ver, n, steps, threads):
    _, ke, secs = O.ref_run(ver, n, steps, threads=threads)
    rate = float(n) * n * steps / secs / 1e9
    rows.append(dict(version=ver, n=n, steps=steps, threads=threads, seconds=secs, gpairs_per_s=rate,
                     gflops_ref_convention=rate * 29.0, kenergy_last=float(ke)))
    print(f"{ver} N={n:7d} steps={steps:4d} threads={threads:3d}  {secs:8.3f} s  {rate:8.3f} Gpairs/s  {rate*29:9.1f} GF(ref conv.)", flush=True)
    return rate


case("ver0", 2000, 500, 1)                       # the reference's default run, single thread
case("ver2", 2000, 500, 1)
for ver in ("ver7", "ver8"):
    case(ver, 2000, 500, cores)
    r = case(ver, 16384, 50, cores)
    for n in (65536, 262144):
        steps = max(1, min(50, int(budget * r * 1e9 / (float(n) * n))))
        r = case(ver, n, steps, cores)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(dict(cpu=cpu, cores=cores, omp_proc_bind="close", rows=rows), open("gpurun_out/cpu_baselines.json", "w"), indent=1)
