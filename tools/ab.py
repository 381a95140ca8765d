#!/usr/bin/env python
"""Same-box A/B of kernel shapes x j-split counts, interleaved repetitions, median reported.
    [NBX_LIB=libnbx_ablation.so] python tools/ab.py N steps reps variants(csv, exact names) splits(csv) [graph [pdls(csv) [smem_pad_kb(csv)]]]"""
import importlib, json, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
nbx = importlib.import_module("nbody-demo-2023_b200").nbx
n, steps, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
names = sys.argv[4].split(",")
splits = [int(x) for x in sys.argv[5].split(",")]
graph = int(sys.argv[6]) if len(sys.argv) > 6 else 0
pdls = [int(x) for x in sys.argv[7].split(",")] if len(sys.argv) > 7 else [-1]
pads = [int(x) for x in sys.argv[8].split(",")] if len(sys.argv) > 8 else [0]
allv = nbx.variant_names()
arrs = nbx.ic(n)
ctxs = {}
for nm in names:
    for sp in splits:
      for pdl in pdls:
       for pad in pads:
        c = nbx.Context(n)
        c.set_option("variant", allv.index(nm)); c.set_option("j_splits", sp); c.set_option("graph", graph); c.set_option("pdl", pdl)
        c.set_option("smem_pad_kb", pad)
        c.upload(*arrs); c.run(max(2, steps // 4))
        ctxs[(nm + {-1: "", 0: "_nopdl", 1: "_pdl"}[pdl] + (f"_pad{pad}" if pad else ""), sp)] = c
res = {k: [] for k in ctxs}
for r in range(reps):
    for k, c in ctxs.items():
        _, secs = c.run(steps)
        res[k].append(secs / steps)
for k, v in sorted(res.items(), key=lambda kv: np.median(kv[1])):
    med = float(np.median(v)); info = ctxs[k].info()
    print(f"{k[0]:30s} occ={info['ctas_per_sm']} S={info['j_splits']:3d} tiles={info['i_tiles']:5d} whole={info['whole_tiles']:5d} med {med*1e3:9.4f} ms  min {min(v)*1e3:9.4f}  {float(n)*n/med/1e9:8.1f} Gpairs/s  {float(n)*n/med/1e9*20e-3/74.45*100:5.1f}% peak", flush=True)
