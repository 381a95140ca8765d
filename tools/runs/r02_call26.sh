#!/bin/bash
# Round 2, GPU call 26: tile hand-over cost at one CTA per SM -- larger TMA tiles at N = 16384.
# (the r4_t512_j2_* shape -- two threads per i-body group, 16 warps in the one CTA an SM gets -- and the tj512/1024/2048 shapes these
# scripts name were measured, found no faster (profiles/r02_ab_c1_16warp_cta.log, r02_ab_c1_tile_size.log) and removed from the source again)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1 NBX_LIB=libnbx_ablation.so
python tools/ab.py 16384 200 7 r4_t256_u4_stage_f2,r4_t256_u4_stage_f2_tj512,r4_t256_u4_stage_f2_tj1024,r4_t256_u4_stage_f2_tj2048 9 1 > $O/r02m_ab_tj_c1.log 2>&1; cat $O/r02m_ab_tj_c1.log
python tools/ab.py 65536 40 5 r4_t256_u4_stage_f2,r4_t256_u4_stage_f2_tj512,r4_t256_u4_stage_f2_tj1024 2 1 > $O/r02m_ab_tj_65536.log 2>&1; cat $O/r02m_ab_tj_65536.log
python tools/ab.py 262144 4 3 r4_t256_u4_stage_f2,r4_t256_u4_stage_f2_tj512,r4_t256_u4_stage_f2_tj1024 0 0 > $O/r02m_ab_tj_262144.log 2>&1; cat $O/r02m_ab_tj_262144.log
