#!/bin/bash
# Round 2, GPU call 8 (2 x B200): multicast release path, then the multi-GPU suite and the 2-GPU CLI again.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1 NBX_VERBOSE=1
timeout 120 tools/mc_probe 2 lib > $O/r02_mc_probe2.log 2>&1; echo "probe rc=$?"; tail -4 $O/r02_mc_probe2.log
echo "== pytest multi"; timeout 1500 python -m pytest tests/test_gpu_multi.py -q -s --timeout 600 -rs > $O/r02_pytest8.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" $O/r02_pytest8.log | tail -2; grep -E "^_{5,} |multicast|Fatal|nbx.py\", line" $O/r02_pytest8.log | head -20
grep -A3 -E "GPUs, default plan" $O/r02_pytest8.log | head -30
echo "== CLI 2 GPUs"; NBODY_GPUS=2 NBODY_SFREQ=5 timeout 300 ./nbody-demo-2023_b200/nbody.x 262144 10 > $O/r02_cli2.log 2>&1; echo "cli rc=$?"; tail -8 $O/r02_cli2.log
echo done
