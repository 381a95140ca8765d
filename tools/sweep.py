#!/usr/bin/env python
"""Time every compiled kernel shape (and a few j-split counts) on one GPU.
    python tools/sweep.py [N] [steps]         -> table + gpurun_out/sweep.json"""
import importlib
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
pkg = importlib.import_module("nbody-demo-2023_b200")
nbx = pkg.nbx

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
splits_list = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
only = sys.argv[4].split(",") if len(sys.argv) > 4 else None      # substrings of variant names
graph = int(os.environ.get("SWEEP_GRAPH", "0"))
arrs = nbx.ic(n)
rows = []
for v, name in enumerate(nbx.variant_names()):
    if only and not any(o in name for o in only):
        continue
    for sp in splits_list:
        with nbx.Context(n) as c:
            c.set_option("variant", v)
            c.set_option("j_splits", sp)
            c.set_option("graph", graph)
            c.upload(*arrs)
            c.run(1)
            best = 1e9
            for _ in range(2):
                _, secs = c.run(steps)
                best = min(best, secs / steps)
            info = c.info()
        rate = float(n) * n / best / 1e9
        rows.append(dict(variant=name, n=n, j_splits=info["j_splits"], i_tiles=info["i_tiles"],
                         ctas_per_sm=info["ctas_per_sm"], ms=best * 1e3, gpairs=rate, tflops20=rate * 20e-3))
        print(f"{name:14s} N={n} tiles={info['i_tiles']:5d} whole={info['whole_tiles']:5d} splits={info['j_splits']:3d} occ={info['ctas_per_sm']} "
              f"{best*1e3:9.3f} ms  {rate:8.1f} Gpairs/s  {rate*20e-3:6.2f} TF(20)  {rate*20e-3/74.5*100:5.1f}% peak", flush=True)
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
with open(os.path.join(REPO, "gpurun_out", f"sweep_{n}.json"), "w") as f:
    json.dump(rows, f, indent=1)
