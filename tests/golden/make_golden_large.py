#!/usr/bin/env python
"""Known-answer fixtures at the HEADLINE sizes, generated from the UNMODIFIED reference (ver8, its
fastest CPU version: ver8/GSimulation.cpp:142-215) compiled by oracle/Makefile into oracle/_ref/.

    python tests/golden/make_golden_large.py [n262144] [c2] [c3] [c1] [c1s10]      (default: all)

  n262144  N =   262,144 uniform cube, 10 steps   (~1 min on 8 cores)
  c2       N = 1,048,576 uniform cube,  2 steps   (~3 min)            BASELINE config 2
  c3       N = 4,194,304 Plummer,       1 step    (~25 min)           BASELINE config 3
  c1       N =    16,384 uniform cube, 500 steps  (~1 min)            BASELINE config 1, whole run

Each fixture (tests/golden/large_<name>_ver8.npz, ~100 KB) holds: per-step kinetic energy (float32,
exact), the fp64 sums of the final px/py/pz, and positions + velocities of 4096 sampled bodies
(seeded choice; indices stored).  ICs: the reference's own uniform cube (oracle.ic_uniform, pinned
bit-exact against the reference) or, for c3, the Plummer positions of nbx_ic_plummer fed to the
reference through ref_state_ver8 (the reference cannot generate them itself).  Needs only
oracle/_ref binaries + libnbx.so's host helpers, so it also runs on a GPU box.
"""
import importlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from oracle import oracle as O  # noqa: E402

CASES = {
    "n262144": dict(n=262144, steps=10, ic="uniform"),
    "c2": dict(n=1 << 20, steps=2, ic="uniform"),
    "c3": dict(n=1 << 22, steps=1, ic="plummer"),
    "c1": dict(n=16384, steps=500, ic="uniform"),
    "c1s10": dict(n=16384, steps=10, ic="uniform"),        # the same run after 10 steps (the north star's position gate)
    "c2s10": dict(n=1 << 20, steps=10, ic="uniform"),      # C2 after 10 steps (~15 min reference, ~65 min fp64 truth on 8 cores)
}
NSEL = 4096


def initial_state(n, ic):
    if ic == "uniform":
        return O.ic_uniform(n)
    nbx = importlib.import_module("nbody-demo-2023_b200").nbx
    s = O.State(n)
    for f, a in zip(O.State.FIELDS, nbx.ic(n, ic)):
        setattr(s, f, a)
    return s


def main():
    names = sys.argv[1:] or list(CASES)
    threads = os.cpu_count() or 1
    for name in names:
        c = CASES[name]
        n, steps = c["n"], c["steps"]
        s0 = initial_state(n, c["ic"])
        t0 = time.time()
        s, ke, secs = O.ref_state_run("ver8", s0, steps, threads=threads)
        sel = np.sort(np.random.default_rng(20231).choice(n, min(NSEL, n), replace=False)).astype(np.int32)
        out = os.path.join(HERE, f"large_{name}_ver8.npz")
        np.savez_compressed(
            out, n=n, steps=steps, ic=c["ic"], threads=threads, ke=ke.astype(np.float32),
            sum_pos=np.array([np.sum(getattr(s, f).astype(np.float64)) for f in ("px", "py", "pz")]),
            norm_pos=np.float64(np.linalg.norm(s.pos().astype(np.float64))),
            sel=sel, pos_sel=s.pos()[sel], vel_sel=s.vel()[sel],
            loop_seconds=secs, generator="tests/golden/make_golden_large.py (oracle/_ref/ref_state_ver8)")
        print(f"{name}: N={n} steps={steps} ic={c['ic']} threads={threads} loop {secs:.1f} s "
              f"({float(n) * n * steps / secs / 1e9:.1f} G pairs/s) wall {time.time() - t0:.1f} s  ke={ke[:3]}... -> {out}", flush=True)


if __name__ == "__main__":
    main()
