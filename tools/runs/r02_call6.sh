#!/bin/bash
# Round 2, GPU call 6 (2 x B200): multi-GPU suite (default plan vs reference + truth, sharded upload, peer timeout,
# torchrun, CLI, NVSwitch multicast), then the 2-GPU bench line.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1 NBX_VERBOSE=1
nvidia-smi topo -m > $O/r02_topo2.log 2>&1
echo "== pytest multi"; timeout 1500 python -m pytest tests/test_gpu_multi.py -q -s --timeout 600 -rs > $O/r02_pytest6.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" $O/r02_pytest6.log | tail -2; grep -E "^_{5,} |multicast|vs fp64 truth|vs reference  " $O/r02_pytest6.log | head -40
echo "== bench 2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r02_bench6_2gpu.json 2> $O/r02_bench6_2gpu.err; echo "bench rc=$?"; tail -c 3000 $O/r02_bench6_2gpu.json; tail -5 $O/r02_bench6_2gpu.err
echo "== CLI 2 GPUs"; NBODY_GPUS=2 NBODY_SFREQ=5 timeout 300 ./nbody-demo-2023_b200/nbody.x 262144 10 > $O/r02_cli2.log 2>&1; tail -12 $O/r02_cli2.log
echo done
