#!/bin/bash
# Round 2, GPU call 25 (1 x B200): the 16-warp single-wave shape (two threads per i-body group) -- correctness and A/B at small N.
# (the r4_t512_j2_* shape -- two threads per i-body group, 16 warps in the one CTA an SM gets -- and the tj512/1024/2048 shapes these
# scripts name were measured, found no faster (profiles/r02_ab_c1_16warp_cta.log, r02_ab_c1_tile_size.log) and removed from the source again)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
python -m pytest tests/test_gpu_parity.py -q -k "every_kernel_shape or j_split_counts or graph_replay" --timeout 600 2>&1 | tail -3
python tests/accuracy_probe.py truth c1s10 r4_t256_u4_stage_f2,r4_t512_j2_u4_stage_f2 2>&1 | tail -3
python tests/accuracy_probe.py truth n262144 r4_t256_u4_stage_f2,r4_t512_j2_u4_stage_f2 2>&1 | tail -3
python tools/ab.py 16384 200 7 r4_t256_u4_stage_f2,r4_t512_j2_u4_stage_f2 9 1 > $O/r02l_ab_j2_c1.log 2>&1; cat $O/r02l_ab_j2_c1.log
python tools/ab.py 65536 40 5 r4_t256_u4_stage_f2,r4_t512_j2_u4_stage_f2 0,2 1 > $O/r02l_ab_j2_65536.log 2>&1; cat $O/r02l_ab_j2_65536.log
python tools/ab.py 32768 100 5 r4_t256_u4_stage_f2,r4_t512_j2_u4_stage_f2 0,4 1 > $O/r02l_ab_j2_32768.log 2>&1; cat $O/r02l_ab_j2_32768.log
python tools/ab.py 262144 4 3 r4_t256_u4_stage_f2,r4_t512_j2_u4_stage_f2 0 0 > $O/r02l_ab_j2_262144.log 2>&1; cat $O/r02l_ab_j2_262144.log
for v in 0 4; do NBX_LIB=libnbx_trace.so python tools/trace_steps.py 16384 24 graph=1 variant=$v 2>&1 | tail -3; done > $O/r02l_trace_j2.log 2>&1; cat $O/r02l_trace_j2.log
