#!/bin/bash
# Round 2, GPU call 23 (1 x B200): does the operand SLOT of the accumulate / multiply inputs matter (swapped multiplicands)?
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1 NBX_LIB=libnbx_ablation.so
V=r4_t256_u4_stage_f2,r4_t256_u4_stage,r4_t256_u4_stage_f2_perm504,r4_t256_u4_stage_f2_perm760,r4_t256_u4_stage_f2_perm1016,r4_t256_u4_stage_f2_perm256,r4_t256_u2_stage_f2_perm504,r4_t256_u2_stage_f2_perm264,r4_t256_u2_stage_f2_perm8,r4_t256_u4_stage_f2_perm184
python tools/ab.py 262144 4 5 $V 0 0 > $O/r02k_ab_swap_262144.log 2>&1; cat $O/r02k_ab_swap_262144.log
python tools/ab.py 1048576 1 3 $V 0 0 > $O/r02k_ab_swap_1m.log 2>&1; cat $O/r02k_ab_swap_1m.log
