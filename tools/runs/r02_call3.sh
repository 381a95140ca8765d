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
python tests/accuracy_probe.py forces 1048576 r4_t256_u4_stage,r4_t256_u4_stage_f2,r4_t256_u2_stage_f2,r4_t256_u4_stage_acc64 > $O/r02c_accuracy_c2.log 2>&1; python tests/accuracy_probe.py truth c2 r4_t256_u4_stage,r4_t256_u4_stage_f2,r4_t256_u2_stage_f2,r4_t256_u4_stage_acc64 >> $O/r02c_accuracy_c2.log 2>&1
cat $O/r02c_accuracy_c2.log
echo done
