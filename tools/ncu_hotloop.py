#!/usr/bin/env python
"""The hot loop of a profiled kernel with per-instruction warp-stall samples, from an .ncu-rep captured with
`ncu --set full --import-source on` (compile with -lineinfo).
    python tools/ncu_hotloop.py gpurun_out/prof.ncu-rep [kernel-name substring] > profiles/rNN_ncu_<kernel>_hotloop_sass.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
# one section per (kernel, view): a "Kernel Name" row, a header row, then the lines; take the first SASS section whose
# kernel name contains the optional second argument
want = sys.argv[2] if len(sys.argv) > 2 else ""
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
kernel = hdr = data = None
for a, b in zip(starts[:-1], starts[1:]):
    if want in rows[a][1] and b - a > 3 and not rows[a + 2][1].lstrip().startswith(("//", "#", "template", "namespace")):
        kernel, hdr, data = rows[a][1], rows[a + 1], [r for r in rows[a + 2:b] if len(r) >= len(rows[a + 1]) - 1]
        break
if kernel is None:
    sys.exit(f"no SASS section for a kernel matching '{want}'")
col = {h: i for i, h in enumerate(hdr)}
ex = [int(r[col["Instructions Executed"]] or 0) for r in data]
mx = max(ex)
keep = [r for r, e in zip(data, ex) if e >= 0.9 * mx]
cols = ["Warp Stall Sampling (All Samples)", "Warp Stall Sampling (Not-issued Samples)", "stall_math", "stall_wait", "stall_short_sb",
        "stall_not_selected", "stall_dispatch", "stall_long_sb"]
print(f"# {kernel}")
print(f"# instructions executed >= 90% of the maximum ({mx} warp-level executions): the inner j loop; {len(keep)} instructions")
print("# SASS | " + " | ".join(c.replace("Warp Stall Sampling ", "samples ") for c in cols))
tot = [0] * len(cols)
for r in keep:
    vals = [int(float(r[col[c]] or 0)) for c in cols]
    tot = [a + b for a, b in zip(tot, vals)]
    print(f"{r[col['Source']].strip():70s}" + "".join(f"{v:9d}" for v in vals))
print(f"{'# total':70s}" + "".join(f"{v:9d}" for v in tot))
ops = {}
for r in keep:
    op = r[col["Source"]].split()[0]
    ops[op] = ops.get(op, 0) + 1
print("# instruction mix: " + ", ".join(f"{k} x{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])))
