#!/bin/bash
# Round 2, GPU call 1 (1 x B200): full GPU suite, default bench line, C1 time split (trace + ncu),
# inner-loop / small-N A/B of the ablation shapes.  Everything lands in gpurun_out/.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r02_smi.log 2>&1
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest -m gpu"; python -m pytest tests -m gpu -q -x --timeout 1500 -rs > $O/r02_pytest1.log 2>&1; echo "pytest rc=$?"; tail -15 $O/r02_pytest1.log
echo "== bench"; python bench.py > $O/r02_bench1.json 2> $O/r02_bench1.err; echo "bench rc=$?"; cut -c1-600 $O/r02_bench1.json; tail -3 $O/r02_bench1.err
echo "== trace C1"
for o in "graph=1" "graph=1 pdl=1" "graph=1 pdl=1 smem_pad_kb=100" "graph=1 smem_pad_kb=100" "graph=0" "graph=1 j_splits=18"; do
  NBX_LIB=libnbx_trace.so python tools/trace_steps.py 16384 24 $o 2>&1 | tail -3
done > $O/r02_trace_c1.log 2>&1; cat $O/r02_trace_c1.log
NBX_LIB=libnbx_trace.so python tools/trace_steps.py 2000 24 graph=1 > $O/r02_trace_c0.log 2>&1; tail -2 $O/r02_trace_c0.log
echo "== A/B small N"
NBX_LIB=libnbx_ablation.so python tools/ab.py 16384 200 7 r4_t256_u4_stage,r4_t256_u2_stage,r4_t512_u2,r4_t256_u4_stage_xjacc 9,18 1 0,1 0,100 > $O/r02_ab_c1.log 2>&1; head -40 $O/r02_ab_c1.log
echo "== A/B inner loop 262144"
NBX_LIB=libnbx_ablation.so python tools/ab.py 262144 4 5 r4_t256_u4_stage,r4_t256_u2_stage,r4_t256_u2_stage_occ3,r4_t256_u1_stage_occ3,r4_t256_u4_stage_xjacc,r4_t256_u2_stage_xjacc,r4_t256_u4_stage_s8,r4_t256_u4_stage_tj512 0 0 > $O/r02_ab_262144.log 2>&1; cat $O/r02_ab_262144.log
echo "== A/B inner loop 1M"
NBX_LIB=libnbx_ablation.so python tools/ab.py 1048576 1 3 r4_t256_u4_stage,r4_t256_u2_stage_occ3,r4_t256_u4_stage_xjacc,r4_t256_u4_stage_s8 0 0 > $O/r02_ab_1m.log 2>&1; cat $O/r02_ab_1m.log
echo "== accuracy of the xj-accumulate shape"
python tests/accuracy_probe.py forces 65536 r4_t256_u4_stage,r4_t256_u4_stage_xjacc > $O/r02_xjacc_accuracy.log 2>&1; python tests/accuracy_probe.py forces 1048576 r4_t256_u4_stage,r4_t256_u4_stage_xjacc >> $O/r02_xjacc_accuracy.log 2>&1
cat $O/r02_xjacc_accuracy.log
echo "== ncu C1 (after a plain run of the same command)"
python tools/run_steps.py 16384 30 graph=0 > $O/r02_ncu_c1_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 20 --csv --log-file $O/r02_launches_c1.csv python tools/run_steps.py 16384 30 graph=0 > $O/r02_ncu_c1_a.log 2>&1
python tools/run_steps.py 16384 30 graph=0 > $O/r02_ncu_c1_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 12 -c 2 -o $O/r02_prof_c1 python tools/run_steps.py 16384 30 graph=0 > $O/r02_ncu_c1_b.log 2>&1
tail -2 $O/r02_ncu_c1_a.log $O/r02_ncu_c1_b.log; head -5 $O/r02_launches_c1.csv
echo done
