#!/bin/bash
# Round 2, GPU call 2 (1 x B200): full GPU suite (no -x), C1 trace after the combine/PDL changes, shape A/B.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; python -m pytest tests -m gpu -q --timeout 1500 -rs > $O/r02_pytest2.log 2>&1; echo "pytest rc=$?"; tail -40 $O/r02_pytest2.log
echo "== trace C1"
for o in "graph=1" "graph=1 pdl=0" "graph=0"; do
  NBX_LIB=libnbx_trace.so python tools/trace_steps.py 16384 24 $o 2>&1 | tail -3
done > $O/r02b_trace_c1.log 2>&1; cat $O/r02b_trace_c1.log
NBX_LIB=libnbx_trace.so python tools/trace_steps.py 2000 24 graph=1 > $O/r02b_trace_c0.log 2>&1; tail -2 $O/r02b_trace_c0.log
echo "== A/B"
NBX_LIB=libnbx_ablation.so python tools/ab.py 16384 200 7 r4_t256_u4_stage,r4_t256_u2_stage,r4_t256_u2,r4_t256_u4,r4_t256_u1 0 1 > $O/r02b_ab_c1.log 2>&1; cat $O/r02b_ab_c1.log
NBX_LIB=libnbx_ablation.so python tools/ab.py 262144 4 5 r4_t256_u4_stage,r4_t256_u2_stage,r4_t256_u2,r4_t256_u4,r4_t256_u1 0 0 > $O/r02b_ab_262144.log 2>&1; cat $O/r02b_ab_262144.log
NBX_LIB=libnbx_ablation.so python tools/ab.py 1048576 1 3 r4_t256_u4_stage,r4_t256_u2_stage,r4_t256_u2,r4_t256_u4 0 0 > $O/r02b_ab_1m.log 2>&1; cat $O/r02b_ab_1m.log
echo done
