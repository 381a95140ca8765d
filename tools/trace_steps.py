#!/usr/bin/env python
"""Where a step's time goes, from per-CTA %globaltimer stamps (trace build: make -C nbody-demo-2023_b200 trace).

    NBX_LIB=libnbx_trace.so python tools/trace_steps.py N [steps] [key=value ...]     e.g. 16384 24 graph=1

Prints, per step (median over the traced steps, the first two skipped): the step period (start of
step s+1 minus start of step s), the launch gap (first CTA start of step s+1 minus last CTA exit of
step s), the prologue (CTA start -> first j tile landed), the j sweep, and the tail after the LAST
sweep ends (split-combine + Euler update + energy, on the last arriver of the slowest tile)."""
import importlib, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
nbx = importlib.import_module("nbody-demo-2023_b200").nbx

n = int(sys.argv[1])
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
opts = dict(kv.split("=") for kv in sys.argv[3:])
arrs = nbx.ic(n)
with nbx.Context(n) as c:
    for k, v in opts.items():
        c.set_option(k, int(v))
    c.set_option("trace_steps", steps)
    c.upload(*arrs)
    c.run(steps)                 # warm-up (also traced, then overwritten)
    ke, secs = c.run(steps)
    t = c.trace().astype(np.int64)
    info = c.info()
start, first, sweep, exit_, smid, last = (t[:, :, k] for k in range(6))
print(f"N={n} steps={steps} shape={nbx.variant_names()[info['variant']]} tiles={info['i_tiles']} whole={info['whole_tiles']} "
      f"splits={info['j_splits']} ctas={t.shape[1]} graph={info['use_graph']} opts={opts}  device time/step {secs/steps*1e6:.2f} us")
rows = []
for s in range(2, steps - 1):
    s0, e0 = start[s].min(), exit_[s].max()
    rows.append(dict(period=(start[s + 1].min() - s0), gap=(start[s + 1].min() - e0), span=(e0 - s0),
                     start_spread=(start[s].max() - s0), prologue=np.median(first[s] - start[s]),
                     sweep=np.median(sweep[s] - first[s]), sweep_max=(sweep[s] - first[s]).max(),
                     last_sweep_end=(sweep[s].max() - s0), tail=(e0 - sweep[s].max()),
                     sms=len(np.unique(smid[s])), per_sm_max=np.bincount(smid[s]).max()))
med = {k: float(np.median([r[k] for r in rows])) for k in rows[0]}
print("median over steps, ns:  " + "  ".join(f"{k}={v:.0f}" for k, v in med.items()))
print(f"  step period {med['period']/1e3:.2f} us = launch gap {med['gap']/1e3:.2f} + kernel span {med['span']/1e3:.2f} "
      f"[CTA start spread {med['start_spread']/1e3:.2f}, prologue {med['prologue']/1e3:.2f}, sweep median {med['sweep']/1e3:.2f} "
      f"max {med['sweep_max']/1e3:.2f}, tail after the last sweep {med['tail']/1e3:.2f}]  on {med['sms']:.0f} SMs (max {med['per_sm_max']:.0f} CTAs/SM)")
