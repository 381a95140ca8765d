#!/bin/bash
# Round 2, GPU call 4: what the two-level accumulation costs as a function of the fold period.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1 NBX_LIB=libnbx_ablation.so
python tools/ab.py 1048576 1 3 r4_t256_u4_stage,r4_t256_u4_stage_f2,r4_t256_u4_stage_f2p16,r4_t256_u4_stage_f2p64,r4_t256_u4_stage_f2p256,r4_t256_u4_stage_f2_tj1024 0 0 > $O/r02d_ab_1m.log 2>&1; cat $O/r02d_ab_1m.log
python tools/ab.py 262144 4 5 r4_t256_u4_stage,r4_t256_u4_stage_f2,r4_t256_u4_stage_f2p16,r4_t256_u4_stage_f2p64,r4_t256_u4_stage_f2p256,r4_t256_u4_stage_f2_tj1024 0 0 > $O/r02d_ab_262144.log 2>&1; cat $O/r02d_ab_262144.log
python - > $O/r02d_accuracy_c2.log 2>&1 <<'PY'
import importlib, sys, numpy as np
sys.path.insert(0, ".")
nbx = importlib.import_module("nbody-demo-2023_b200").nbx
names = nbx.variant_names()
t = np.load("tests/golden/truth_c2_fp64.npz")
n = int(t["n"]); arrs = nbx.ic(n); sel = t["sel"]
for nm in ("r4_t256_u4_stage_f2", "r4_t256_u4_stage_f2p16", "r4_t256_u4_stage_f2p64", "r4_t256_u4_stage_f2p256"):
    with nbx.Context(n) as c:
        c.set_option("variant", names.index(nm)); c.upload(*arrs)
        ke, _ = c.run(2); st = c.state()
    pos = np.stack([a[sel] for a in st[:3]], axis=1)
    print(f"{nm:26s} ke vs truth {(ke - t['ke']) / t['ke']}  pos {np.linalg.norm(pos - t['pos_sel']) / np.linalg.norm(t['pos_sel']):.2e}")
PY
cat $O/r02d_accuracy_c2.log
