#!/bin/bash
# Round 2, GPU call 18 (1 x B200): source-order permutations of the default inner loop (ptxas register-assignment lottery).
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1 NBX_LIB=libnbx_ablation.so
V=r4_t256_u4_stage_f2,r4_t256_u4_stage
for k in 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15; do V=$V,r4_t256_u4_stage_f2_perm$k; done
V=$V,r4_t256_u2_stage_f2,r4_t256_u2_stage_f2_perm1,r4_t256_u2_stage_f2_perm2,r4_t256_u2_stage_f2_perm4,r4_t256_u2_stage_f2_perm8
python tools/ab.py 262144 4 5 $V 0 0 > $O/r02h_ab_perm_262144.log 2>&1; cat $O/r02h_ab_perm_262144.log
BEST=$(head -6 $O/r02h_ab_perm_262144.log | awk '{print $1}' | paste -sd, -)
python tools/ab.py 1048576 1 3 r4_t256_u4_stage_f2,$BEST 0 0 > $O/r02h_ab_perm_1m.log 2>&1; cat $O/r02h_ab_perm_1m.log
python tools/ab.py 16384 200 5 r4_t256_u4_stage_f2,$BEST 0 1 > $O/r02h_ab_perm_c1.log 2>&1; cat $O/r02h_ab_perm_c1.log
