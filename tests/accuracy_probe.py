#!/usr/bin/env python
"""Accuracy of kernel shapes against the oracle (test infrastructure: this file may import oracle/, tools/ may not).

    [NBX_LIB=libnbx_ablation.so] python tests/accuracy_probe.py forces  N shape[,shape...]     sampled forces vs oracle.acc_fp64
    [NBX_LIB=libnbx_ablation.so] python tests/accuracy_probe.py truth   case shape[,shape...]  steps from the ICs vs tests/golden/truth_<case>_fp64.npz

Used by tools/runs/r02_call{1,3,4}.sh (profiles/r02_xjacc_accuracy.log, r02_accumulation_accuracy_c2.log,
r02_fold_period_accuracy_c2.log)."""
import importlib
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
nbx = importlib.import_module("nbody-demo-2023_b200").nbx
from oracle import oracle as O  # noqa: E402


def state_of(arrs):
    s = O.State(arrs[0].shape[0])
    for f, a in zip(O.State.FIELDS, arrs):
        setattr(s, f, a)
    return s


def forces(n, shapes):
    names = nbx.variant_names()
    arrs = nbx.ic(n)
    s = state_of(arrs)
    sel = np.random.default_rng(5).choice(n, 512, replace=False).astype(np.int32)
    truth = O.acc_fp64(s, sel)
    ref32 = O.acc_f32(s, sel).astype(np.float64)
    tn2 = np.sum(truth * truth, axis=1)

    def report(label, acc):
        proj = np.sum((acc - truth) * truth, axis=1) / tn2
        err = np.linalg.norm(acc - truth, axis=1) / np.sqrt(tn2)
        print(f"N={n} {label:30s} force vs fp64: signed bias along a {proj.mean():+.2e}  |err| median {np.median(err):.2e} max {err.max():.2e}")
    report("reference float order (ver2)", ref32)
    for nm in shapes:
        with nbx.Context(n) as c:
            c.set_option("variant", names.index(nm))
            c.upload(*arrs)
            report(nm, c.accelerations()[sel].astype(np.float64))


def truth(case, shapes):
    names = nbx.variant_names()
    t = np.load(os.path.join(REPO, "tests", "golden", f"truth_{case}_fp64.npz"))
    r = np.load(os.path.join(REPO, "tests", "golden", f"large_{case}_ver8.npz"))
    n, steps, sel = int(t["n"]), int(t["steps"]), t["sel"]
    arrs = nbx.ic(n, str(t["ic"]))
    print(f"{case}: reference ver8 vs truth: kenergy {(r['ke'] - t['ke']) / t['ke']}  pos {np.linalg.norm(r['pos_sel'] - t['pos_sel']) / np.linalg.norm(t['pos_sel']):.2e}")
    for nm in shapes:
        with nbx.Context(n) as c:
            c.set_option("variant", names.index(nm))
            c.upload(*arrs)
            ke, _ = c.run(steps)
            st = c.state()
        pos = np.stack([a[sel] for a in st[:3]], axis=1)
        vel = np.stack([a[sel] for a in st[3:6]], axis=1)
        print(f"{nm:30s} kenergy vs truth {(ke - t['ke']) / t['ke']}  pos {np.linalg.norm(pos - t['pos_sel']) / np.linalg.norm(t['pos_sel']):.2e} "
              f"vel {np.linalg.norm(vel - t['vel_sel']) / np.linalg.norm(t['vel_sel']):.2e}  kenergy vs ver8 {np.max(np.abs(ke - r['ke']) / r['ke']):.2e}")


if __name__ == "__main__":
    mode, arg, shapes = sys.argv[1], sys.argv[2], sys.argv[3].split(",")
    forces(int(arg), shapes) if mode == "forces" else truth(arg, shapes)
