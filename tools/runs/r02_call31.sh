#!/bin/bash
# Round 2, GPU call 31 (1 x B200): the q-scaled 11-instruction pair ("_qs" shapes).
#   A  bounds-checked sanity matrix (libnbx_debug.so) including the _qs shape
#   B  accuracy: sampled forces and the C2 truth fixture, 12-instruction default vs _qs
#   C  same-box A/B at N = 1 M, 262 144, 65 536 (source orders of the _qs loop body too)
#   D  only if A passed and _qs is >= 2 % faster at 1 M: the whole GPU suite and the bench with the _qs default
#      (libnbx_qsdefault.so = -DNBX_LARGE_QS)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
D=r4_t256_u4_stage_f2; Q=${D}_qs
NBX_LIB=libnbx_debug.so timeout 400 python tests/sanity_small.py > $O/r02q_sanity_debug.log 2>&1; A=$?; echo "A sanity rc=$A"; tail -2 $O/r02q_sanity_debug.log
export NBX_LIB=libnbx_ablation.so
(timeout 200 python tests/accuracy_probe.py forces 2000 $D,$Q; timeout 200 python tests/accuracy_probe.py forces 262144 $D,$Q,r4_t256_u4_stage_qs
 timeout 300 python tests/accuracy_probe.py truth c2 $D,$Q) > $O/r02q_accuracy.log 2>&1; cat $O/r02q_accuracy.log
timeout 300 python tools/ab.py 1048576 2 3 $D,$Q,${Q}_perm0,${Q}_perm505,${Q}_perm248,${Q}_perm256,${Q}_perm8,r4_t256_u2_stage_f2_qs,r4_t256_u4_stage_qs 0 0 > $O/r02q_ab_1m.log 2>&1; cat $O/r02q_ab_1m.log
timeout 200 python tools/ab.py 262144 4 3 $D,$Q,${Q}_perm0,${Q}_perm505,${Q}_perm248,${Q}_perm256,${Q}_perm8,r4_t256_u2_stage_f2_qs 0 0 > $O/r02q_ab_262144.log 2>&1; cat $O/r02q_ab_262144.log
timeout 200 python tools/ab.py 65536 40 5 $D,$Q 0 0 > $O/r02q_ab_65536.log 2>&1; cat $O/r02q_ab_65536.log
GO=$(python - <<'PY'
import re
r={}
for l in open("gpurun_out/r02q_ab_1m.log"):
    m=re.match(r"(\S+)\s.*med\s+([0-9.]+) ms", l)
    if m: r[m.group(1)]=float(m.group(2))
d, q = r.get("r4_t256_u4_stage_f2"), r.get("r4_t256_u4_stage_f2_qs")
print(1 if d and q and q < 0.98*d else 0)
PY
)
echo "A=$A GO=$GO"
if [ "$A" = "0" ] && [ "$GO" = "1" ]; then
  export NBX_LIB=libnbx_qsdefault.so
  timeout 600 python -m pytest tests -m gpu -x -q -rs > $O/r02q_pytest_gpu_qsdefault.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02q_pytest_gpu_qsdefault.log
  timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02q_bench_qsdefault.json 2> $O/r02q_bench_qsdefault.err; echo "bench rc=$?"
  python - <<'PY'
import json
d=json.load(open("gpurun_out/r02q_bench_qsdefault.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["config"].get("kernel_shape"), d["gpu_launches"], d["parity"]["ok"])
print(json.dumps(d["parity"])[:1500])
print({k: (v.get("value"), v.get("ms_per_step")) for k, v in d.get("also", {}).items()})
print(d["config"].get("strong_anchor"))
PY
fi
