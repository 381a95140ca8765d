#!/bin/bash
# Round 2, GPU call 16 (8 x B200): the 8-GPU bench line (C3 strong scaling + C4 + exchange A/B + parity), the
# multi-GPU suite at 2/4/8 ranks, the 8-GPU CLI.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1 NBX_VERBOSE=1
nvidia-smi topo -m > $O/r02_topo8.log 2>&1; nproc > $O/r02_nproc8.log
echo "== bench 8"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r02_bench16_8gpu.json 2> $O/r02_bench16_8gpu.err; echo "bench rc=$?"; grep -E "nbx:|Error|error|Traceback" $O/r02_bench16_8gpu.err | head; python - <<'PY'
import json
line=[l for l in open("gpurun_out/r02_bench16_8gpu.json") if l.startswith("{")][0]
d=json.loads(line)
print(d["value"], d["ms_per_step"], d.get("strong_efficiency"), d["e2e"]["value"], d["config"]["parallelism"], d["roofline"]["frac"])
print(json.dumps(d["exchange_ab"]))
print("parity", json.dumps(d["parity"]))
print("anchor", d["config"]["strong_anchor"])
c4=d["also"]["c4"]; print("c4", c4["ms_per_step"], c4["value"], c4["frac_fp32_peak"], json.dumps(c4["parity"]))
PY
echo "== pytest multi"; timeout 1800 python -m pytest tests/test_gpu_multi.py -q -s --timeout 900 -rs > $O/r02_pytest16.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" $O/r02_pytest16.log | tail -2; grep -E "^_{5,} |multicast active|Fatal|nbx:" $O/r02_pytest16.log | head -20
grep -A3 -E "GPUs, default plan" $O/r02_pytest16.log | head -40
echo "== CLI 8 GPUs C3"; NBODY_GPUS=8 NBODY_IC=plummer NBODY_SFREQ=5 timeout 600 ./nbody-demo-2023_b200/nbody.x 4194304 20 > $O/r02_cli8_c3.log 2>&1; echo "cli rc=$?"; cat $O/r02_cli8_c3.log
echo done
