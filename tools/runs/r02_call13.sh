#!/bin/bash
# Round 2, GPU call 13 (1 x B200): ncu evidence for the final default kernel -- launch list of the bench command,
# full capture of one C2 launch and one C1 launch, C1/C0 time split with the trace build.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity --no-e2e"
$CMD > $O/r02f_bench_plain.json 2> $O/r02f_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02f_launches_bench_c2.csv $CMD > $O/r02f_ncu_a.log 2>&1
echo "launch list rc=$?"; cut -c1-200 $O/r02f_bench_plain.json
$CMD > $O/r02f_bench_plain2.json 2>> $O/r02f_bench_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o $O/r02f_prof_c2 $CMD > $O/r02f_ncu_b.log 2>&1
echo "full c2 rc=$?"
python tools/run_steps.py 16384 30 graph=0 > $O/r02f_c1_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 12 -c 1 -o $O/r02f_prof_c1 python tools/run_steps.py 16384 30 graph=0 > $O/r02f_ncu_c.log 2>&1
echo "full c1 rc=$?"; cat $O/r02f_c1_plain.log
for o in "graph=1" "graph=1 pdl=0"; do NBX_LIB=libnbx_trace.so python tools/trace_steps.py 16384 24 $o 2>&1 | tail -3; done > $O/r02f_trace_c1.log 2>&1; cat $O/r02f_trace_c1.log
NBX_LIB=libnbx_trace.so python tools/trace_steps.py 2000 24 graph=1 > $O/r02f_trace_c0.log 2>&1; tail -2 $O/r02f_trace_c0.log
echo done
