#!/bin/bash
# Round 2, GPU call 5 (1 x B200): full GPU suite on the new default (two-level accumulation), default bench
# line, shape A/B on the final source.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; python -m pytest tests -m gpu -q --timeout 1500 -rs > $O/r02_pytest5.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" $O/r02_pytest5.log | tail -2; grep -E "^_{5,} " $O/r02_pytest5.log
grep -A4 -E "^(N=262144|C2 N=1M|C3 N=4M|C1 N=16384|C2 \{)" $O/r02_pytest5.log | head -60
echo "== bench"; python bench.py > $O/r02_bench5.json 2> $O/r02_bench5.err; echo "bench rc=$?"; cut -c1-300 $O/r02_bench5.json; tail -3 $O/r02_bench5.err
echo "== A/B"
export NBX_LIB=libnbx_ablation.so
python tools/ab.py 16384 200 7 r4_t256_u4_stage_f2,r4_t256_u4_stage,r4_t256_u2_stage_f2,r4_t256_u4_f2 0 1 > $O/r02e_ab_c1.log 2>&1; cat $O/r02e_ab_c1.log
python tools/ab.py 262144 4 5 r4_t256_u4_stage_f2,r4_t256_u4_stage,r4_t256_u2_stage_f2,r4_t256_u4_f2,r4_t256_u4_stage_f2p16 0 0 > $O/r02e_ab_262144.log 2>&1; cat $O/r02e_ab_262144.log
python tools/ab.py 1048576 1 3 r4_t256_u4_stage_f2,r4_t256_u4_stage,r4_t256_u2_stage_f2,r4_t256_u4_f2,r4_t256_u4_stage_f2p16 0 0 > $O/r02e_ab_1m.log 2>&1; cat $O/r02e_ab_1m.log
echo done
