#!/bin/bash
# Round 2, GPU call 20 (1 x B200): full GPU suite + default bench line on the final default kernel.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1

echo "(suite skipped in this re-run)"
echo "== bench"; SECONDS=0; python bench.py > $O/r02_bench20.json 2> $O/r02_bench20.err; echo "bench rc=$? wall=${SECONDS}s"; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench20.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["parity"]["ok"])
print("c1", d["also"]["c1"]["ms_per_step"], d["also"]["c1"]["frac_fp32_peak"], d["also"]["c1"]["parity"]["ok"], "c0", d["also"]["c0"]["ms_per_step"], d["also"]["c0"]["parity"]["ok"])
a=d["config"]["strong_anchor"]; print("anchor", a["ms_per_step"], a["value"], a["parity"]["ok"], a["parity"]["vs_reference_output"]["gpu_vs_fp64_truth"])
print(d["cpu_baseline"]["value"], d["gpu_reference"]["kernel_only"]["value"], d["gpu_reference"]["this_build_same_n"])
PY
echo "== reference arm"; SECONDS=0; python bench.py --impl reference --steps 5 --warmup 3 > $O/r02_bench20_ref.json 2> $O/r02_bench20_ref.err; echo "ref wall=${SECONDS}s"; cut -c1-260 $O/r02_bench20_ref.json
