#!/bin/bash
# Round 2, GPU call 17 (4 x B200): 4-GPU bench line and the multi-GPU suite at 2/4 ranks.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1 NBX_VERBOSE=1
echo "== bench 4"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 4 --steps 4 --warmup 3 > $O/r02_bench17_4gpu.json 2> $O/r02_bench17_4gpu.err; echo "bench rc=$?"; grep -E "nbx:|Error|error|Traceback" $O/r02_bench17_4gpu.err | head; python - <<'PY'
import json
line=[l for l in open("gpurun_out/r02_bench17_4gpu.json") if l.startswith("{")][0]
d=json.loads(line)
print(d["value"], d["ms_per_step"], d.get("strong_efficiency"), d["e2e"]["value"], d["config"]["parallelism"], d["roofline"]["frac"])
print(json.dumps(d["exchange_ab"]))
print("parity", json.dumps(d["parity"]))
print("anchor", d["config"]["strong_anchor"])
PY
echo "== pytest multi"; timeout 1800 python -m pytest tests/test_gpu_multi.py -q -s --timeout 900 -rs > $O/r02_pytest17.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" $O/r02_pytest17.log | tail -2; grep -E "^_{5,} |multicast active|Fatal|nbx:" $O/r02_pytest17.log | head -20
grep -A3 -E "GPUs, default plan" $O/r02_pytest17.log | head -40
echo done
