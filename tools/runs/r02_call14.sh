#!/bin/bash
# Round 2, GPU call 14: do two co-resident CTAs fill each other's tile-boundary bubbles when their boundaries are staggered?
# (the "stagger" option this script toggles was an experiment that showed no effect and was removed again; see profiles/r02_trace_c1_two_ctas_per_sm_stagger.log)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
for o in "graph=1" "graph=1 j_splits=18" "graph=1 j_splits=18 stagger=1" "graph=1 j_splits=18 stagger=1 pdl=1" "graph=1 j_splits=37 stagger=1" "graph=1 j_splits=37"; do
  NBX_LIB=libnbx_trace.so python tools/trace_steps.py 16384 24 $o 2>&1 | tail -3
done > $O/r02g_trace_stagger.log 2>&1; cat $O/r02g_trace_stagger.log
python tools/run_steps.py 16384 2000 graph=1
python tools/run_steps.py 16384 2000 graph=1 j_splits=18
python tools/run_steps.py 16384 2000 graph=1 j_splits=18 stagger=1
python tools/run_steps.py 262144 20
python tools/run_steps.py 262144 20 stagger=1
python tools/run_steps.py 1048576 4
python tools/run_steps.py 1048576 4 stagger=1
