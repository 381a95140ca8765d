#!/bin/bash
# Round 2, GPU call 28 (2 x B200): the driver's SCALE sequence in miniature -- bench --gpus 1 (writes the C3 anchor), then --gpus 2
# on the same box must pick the cached anchor up.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O; rm -f $O/strong_anchor.json
export PYTHONUNBUFFERED=1
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $O/r02_seq_1gpu.json 2> $O/r02_seq_1gpu.err; echo "n=1 rc=$?"; ls -la $O/strong_anchor.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r02_seq_2gpu.json 2> $O/r02_seq_2gpu.err; echo "n=2 rc=$?"
python - <<'PY'
import json
d1=json.load(open("gpurun_out/r02_seq_1gpu.json"))
line=[l for l in open("gpurun_out/r02_seq_2gpu.json") if l.startswith("{")][0]
d2=json.loads(line)
print(d1["value"], d1["config"]["strong_anchor"]["ms_per_step"])
print(d2["value"], d2["ms_per_step"], d2["strong_efficiency"], d2["config"]["strong_anchor"]["source"], d2["config"]["strong_anchor"]["ms_per_step"], d2["parity"]["ok"])
PY
