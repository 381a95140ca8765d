#!/bin/bash
# Round 2, GPU call 3 (1 x B200): two-level float accumulation (f2) -- speed A/B and accuracy against the
# fp64 truth fixture; 8-bodies-per-thread shapes at small N.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1 NBX_LIB=libnbx_ablation.so
python tools/ab.py 1048576 1 3 r4_t256_u4_stage,r4_t256_u4_stage_f2,r4_t256_u2_stage_f2,r4_t256_u4_f2,r4_t256_u4_stage_acc64 0 0 > $O/r02c_ab_1m.log 2>&1; cat $O/r02c_ab_1m.log
python tools/ab.py 262144 4 5 r4_t256_u4_stage,r4_t256_u4_stage_f2,r4_t256_u2_stage_f2,r4_t256_u4_f2,r4_t256_u4_stage_acc64,r8_t256_u1_stage,r8_t256_u2_stage,r8_t256_u1_stage_f2,r6_t256_u2_stage 0 0 > $O/r02c_ab_262144.log 2>&1; cat $O/r02c_ab_262144.log
python tools/ab.py 16384 200 7 r4_t256_u4_stage,r4_t256_u4_stage_f2,r4_t256_u2_stage_f2 9 1 > $O/r02c_ab_c1.log 2>&1
python tools/ab.py 16384 200 7 r8_t256_u1_stage,r8_t256_u2_stage,r8_t256_u1_stage_f2,r8_t128_u1_stage 18 1 >> $O/r02c_ab_c1.log 2>&1
python tools/ab.py 16384 200 7 r6_t256_u2_stage 13 1 >> $O/r02c_ab_c1.log 2>&1; cat $O/r02c_ab_c1.log
echo "== accuracy vs the fp64 truth (C2, 2 steps from the ICs)"
python - > $O/r02c_accuracy_c2.log 2>&1 <<'PY'
import importlib, sys, numpy as np
sys.path.insert(0, ".")
nbx = importlib.import_module("nbody-demo-2023_b200").nbx
from oracle import oracle as O
names = nbx.variant_names()
t = np.load("tests/golden/truth_c2_fp64.npz"); r = np.load("tests/golden/large_c2_ver8.npz")
n = int(t["n"]); arrs = nbx.ic(n); sel = t["sel"]
s = O.State(n)
for f, a in zip(O.State.FIELDS, arrs): setattr(s, f, a)
sub = np.random.default_rng(3).choice(n, 512, replace=False).astype(np.int32)
a64 = O.acc_fp64(s, sub)
print("reference ver8 vs truth: ke", (r["ke"] - t["ke"]) / t["ke"], "pos", np.linalg.norm(r["pos_sel"] - t["pos_sel"]) / np.linalg.norm(t["pos_sel"]))
for nm in ("r4_t256_u4_stage", "r4_t256_u4_stage_f2", "r4_t256_u2_stage_f2", "r4_t256_u4_stage_acc64"):
    with nbx.Context(n) as c:
        c.set_option("variant", names.index(nm)); c.upload(*arrs)
        acc = c.accelerations()[sub].astype(np.float64)
        ke, _ = c.run(2); st = c.state()
    proj = np.sum((acc - a64) * a64, axis=1) / np.sum(a64 * a64, axis=1)
    err = np.linalg.norm(acc - a64, axis=1) / np.linalg.norm(a64, axis=1)
    pos = np.stack([a[sel] for a in st[:3]], axis=1); vel = np.stack([a[sel] for a in st[3:6]], axis=1)
    print(f"{nm:26s} ke vs truth {(ke - t['ke']) / t['ke']}  pos {np.linalg.norm(pos - t['pos_sel']) / np.linalg.norm(t['pos_sel']):.2e} vel {np.linalg.norm(vel - t['vel_sel']) / np.linalg.norm(t['vel_sel']):.2e}"
          f"  force: signed bias {proj.mean():+.2e} |err| median {np.median(err):.2e} max {err.max():.2e};  ke vs ver8 {np.max(np.abs(ke - r['ke']) / r['ke']):.2e}")
PY
cat $O/r02c_accuracy_c2.log
echo done
