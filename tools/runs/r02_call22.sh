#!/bin/bash
# Round 2, GPU call 22 (1 x B200): final evidence on the final kernel -- full GPU suite, ncu launch list + full capture
# (C2 and C1), the full headline config through the CLI (N = 1 M x 50 steps), a driver-style bench run.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; python -m pytest tests -m gpu -q --timeout 1500 -rs > $O/r02_pytest22.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" $O/r02_pytest22.log | tail -2; grep -E "^_{5,} |SKIPPED" $O/r02_pytest22.log | head
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity --no-e2e"
$CMD > $O/r02j_bench_plain.json 2> $O/r02j_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02j_launches_bench_c2.csv $CMD > $O/r02j_ncu_a.log 2>&1
echo "launch list rc=$?"
$CMD > $O/r02j_bench_plain2.json 2>> $O/r02j_bench_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o $O/r02j_prof_c2 $CMD > $O/r02j_ncu_b.log 2>&1
echo "full c2 rc=$?"
python tools/run_steps.py 16384 30 graph=0 > $O/r02j_c1_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 12 -c 1 -o $O/r02j_prof_c1 python tools/run_steps.py 16384 30 graph=0 > $O/r02j_ncu_c.log 2>&1
echo "full c1 rc=$?"
echo "== CLI C2 full config"; NBODY_SFREQ=10 ./nbody-demo-2023_b200/nbody.x 1048576 50 > $O/r02_cli_c2_50steps.log 2>&1; echo "cli rc=$?"; cat $O/r02_cli_c2_50steps.log
echo "== CLI C1 full config"; ./nbody-demo-2023_b200/nbody.x 16384 500 > $O/r02_cli_c1_500steps.log 2>&1; tail -14 $O/r02_cli_c1_500steps.log
echo "== driver-style bench"; SECONDS=0; python bench.py --gpus 1 --steps 20 --warmup 3 > $O/r02_bench22_steps20.json 2> $O/r02_bench22.err; echo "bench rc=$? wall=${SECONDS}s"; cut -c1-230 $O/r02_bench22_steps20.json
for o in "graph=1" "graph=1 pdl=0"; do NBX_LIB=libnbx_trace.so python tools/trace_steps.py 16384 24 $o 2>&1 | tail -3; done > $O/r02j_trace_c1.log 2>&1; cat $O/r02j_trace_c1.log
echo done
