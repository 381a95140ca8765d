#!/usr/bin/env python
"""Minimal driver for profilers: N bodies, K steps, optional key=value options (no checker, no extras).
    python tools/run_steps.py N K [graph=0 j_splits=9 variant=0 ...]"""
import importlib, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
nbx = importlib.import_module("nbody-demo-2023_b200").nbx
n, k = int(sys.argv[1]), int(sys.argv[2])
with nbx.Context(n) as c:
    for kv in sys.argv[3:]:
        key, v = kv.split("=")
        c.set_option(key, int(v))
    c.upload(*nbx.ic(n))
    ke, secs = c.run(k)
    i = c.info()
    print(f"N={n} steps={k} shape={nbx.variant_names()[i['variant']]} tiles={i['i_tiles']} whole={i['whole_tiles']} splits={i['j_splits']} "
          f"graph={i['use_graph']} {secs/k*1e3:.5f} ms/step {float(n)*n*k/secs/1e9:.1f} Gpairs/s ke={ke[-1]:.9g}")
