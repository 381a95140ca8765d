#!/bin/bash
# Round 2, final scaling lines on the final kernel: bash tools/runs/r02_call21.sh N   (gpurun --gpus N)
set -u
cd "$(dirname "$0")/../.."
N=${1:-8}; O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus $N --steps 5 --warmup 3 > $O/r02_final_bench_${N}gpu.json 2> $O/r02_final_bench_${N}gpu.err; echo "bench rc=$?"
python - $N <<'PY'
import json, sys
n=sys.argv[1]
line=[l for l in open(f"gpurun_out/r02_final_bench_{n}gpu.json") if l.startswith("{")][0]
d=json.loads(line)
print(d["value"], d["ms_per_step"], d.get("strong_efficiency"), d["e2e"]["value"], d["config"]["parallelism"], d["roofline"]["frac"], d["parity"]["ok"])
print("anchor", d["config"]["strong_anchor"]["ms_per_step"], d["config"]["strong_anchor"]["source"])
print({k: v.get("ms_per_step") for k, v in d["exchange_ab"].items()})
if "also" in d and "c4" in d["also"]:
    c4=d["also"]["c4"]; print("c4", c4["ms_per_step"], c4["value"], c4["frac_fp32_peak"], c4["parity"]["ok"])
PY
