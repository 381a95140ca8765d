#!/bin/bash
# Round 2, GPU call 4: what the two-level accumulation costs as a function of the fold period.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1 NBX_LIB=libnbx_ablation.so
python tools/ab.py 1048576 1 3 r4_t256_u4_stage,r4_t256_u4_stage_f2,r4_t256_u4_stage_f2p16,r4_t256_u4_stage_f2p64,r4_t256_u4_stage_f2p256,r4_t256_u4_stage_f2_tj1024 0 0 > $O/r02d_ab_1m.log 2>&1; cat $O/r02d_ab_1m.log
python tools/ab.py 262144 4 5 r4_t256_u4_stage,r4_t256_u4_stage_f2,r4_t256_u4_stage_f2p16,r4_t256_u4_stage_f2p64,r4_t256_u4_stage_f2p256,r4_t256_u4_stage_f2_tj1024 0 0 > $O/r02d_ab_262144.log 2>&1; cat $O/r02d_ab_262144.log
python tests/accuracy_probe.py truth c2 r4_t256_u4_stage_f2p4,r4_t256_u4_stage_f2p16,r4_t256_u4_stage_f2,r4_t256_u4_stage_f2p256 > $O/r02d_accuracy_c2.log 2>&1
cat $O/r02d_accuracy_c2.log
