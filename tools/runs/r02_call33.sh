#!/bin/bash
# Round 2, GPU call 33 (2 x B200): the multi-GPU paths with the q-scaled default (qscale_kernel carries the bounded peer
# wait in P2P mode): the multi-GPU suite, then the 2-GPU C3 bench line (headline + e2e + parity only).
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rs > $O/r02j_pytest_multi_2gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r02j_pytest_multi_2gpu.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 2 --steps 3 --warmup 3 --no-extras > $O/r02j_bench_2gpu.json 2> $O/r02j_bench_2gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
line=[l for l in open("gpurun_out/r02j_bench_2gpu.json") if l.startswith("{")][0]
d=json.loads(line)
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["parallelism"], d["config"].get("kernel_shape"), d["roofline"]["frac"], d["gpu_launches"], d["parity"]["ok"])
print(json.dumps(d["parity"])[:1200])
PY
