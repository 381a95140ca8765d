#!/bin/bash
# Round 2, GPU call 34 (1 x B200): ncu evidence for the q-scaled default at C2 -- one full capture of qscale_kernel + step_kernel
# (second step), then the launch list of a short bench run.  Each ncu pass only after the same command ran plain and exited 0.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
R="python tools/run_steps.py 1048576 3"
timeout 60 $R > $O/r02k_run_plain.log 2>&1 && \
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"step_kernel|qscale_kernel" -s 2 -c 2 -o $O/r02k_prof_c2_qi $R > $O/r02k_ncu_full.log 2>&1
echo "full rc=$?"; cat $O/r02k_run_plain.log; tail -3 $O/r02k_ncu_full.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity --no-e2e"
timeout 90 $CMD > $O/r02k_bench_plain.json 2> $O/r02k_bench_plain.err && \
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02k_launches_bench_c2_qi.csv $CMD > $O/r02k_ncu_list.log 2>&1
echo "launch list rc=$?"; cut -c1-200 $O/r02k_bench_plain.json; ls -la $O/r02k_*
