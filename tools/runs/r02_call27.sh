#!/bin/bash
# Round 2, GPU call 27: more, smaller CTAs per SM (4 x 128 threads, 8 x 64 threads) with the final loop body.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1 NBX_LIB=libnbx_ablation.so
V=r4_t256_u4_stage_f2,r4_t128_u4_stage_f2,r4_t128_u4_stage_f2_perm248,r4_t128_u4_stage_f2_perm0,r4_t64_u4_stage_f2
python tools/ab.py 262144 4 3 $V 0 0 > $O/r02n_ab_small_ctas_262144.log 2>&1; cat $O/r02n_ab_small_ctas_262144.log
python tools/ab.py 1048576 1 3 $V 0 0 > $O/r02n_ab_small_ctas_1m.log 2>&1; cat $O/r02n_ab_small_ctas_1m.log
python tools/ab.py 16384 200 5 $V 0 1 > $O/r02n_ab_small_ctas_c1.log 2>&1; cat $O/r02n_ab_small_ctas_c1.log
