#!/bin/bash
# Round 2, GPU call 15 (2 x B200): multicast across processes (torchrun) + the multi-GPU suite + the 2-GPU bench line.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1 NBX_VERBOSE=1
echo "== pytest multi"; timeout 1500 python -m pytest tests/test_gpu_multi.py -q -s --timeout 600 -rs > $O/r02_pytest15.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" $O/r02_pytest15.log | tail -2; grep -E "^_{5,} |multicast|Fatal|nbx:" $O/r02_pytest15.log | head -20
echo "== bench 2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r02_bench15_2gpu.json 2> $O/r02_bench15_2gpu.err; echo "bench rc=$?"; grep -E "nbx:|Error|error" $O/r02_bench15_2gpu.err | head; python - <<'PY'
import json
line=[l for l in open("gpurun_out/r02_bench15_2gpu.json") if l.startswith("{")][0]
d=json.loads(line)
print(d["value"], d["ms_per_step"], d.get("strong_efficiency"), d["e2e"]["value"], d["config"]["parallelism"])
print(json.dumps(d["exchange_ab"], indent=0))
print(d["parity"]["ok"], d["parity"].get("replicas_bit_equal"), d["parity"]["vs_reference_output"])
PY
echo "== CLI 2 GPUs nccl + version line"; NCCL_DEBUG=VERSION NBODY_EXCHANGE=nccl NBODY_GPUS=2 NBODY_SFREQ=5 timeout 300 ./nbody-demo-2023_b200/nbody.x 4096 10 2>/dev/null | head -4
echo done
