#!/bin/bash
# Round 2, GPU call 19 (1 x B200): second round of source-order permutations (stage-wise reversed body order).
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1 NBX_LIB=libnbx_ablation.so
V=$(python - <<'PY'
import importlib,sys
sys.path.insert(0,".")
nbx=importlib.import_module("nbody-demo-2023_b200").nbx
print(",".join(n for n in nbx.variant_names() if "perm" in n or n in ("r4_t256_u4_stage_f2","r4_t256_u4_stage","r4_t256_u2_stage_f2")))
PY
)
python tools/ab.py 262144 4 3 $V 0 0 > $O/r02i_ab_perm_262144.log 2>&1; head -16 $O/r02i_ab_perm_262144.log; tail -4 $O/r02i_ab_perm_262144.log
BEST=$(head -8 $O/r02i_ab_perm_262144.log | awk '{print $1}' | paste -sd, -)
python tools/ab.py 1048576 1 3 r4_t256_u4_stage_f2,r4_t256_u4_stage,$BEST 0 0 > $O/r02i_ab_perm_1m.log 2>&1; cat $O/r02i_ab_perm_1m.log
python tools/ab.py 16384 200 5 r4_t256_u4_stage_f2,$BEST 0 1 > $O/r02i_ab_perm_c1.log 2>&1; cat $O/r02i_ab_perm_c1.log
