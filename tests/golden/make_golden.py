#!/usr/bin/env python
"""Generate the known-answer fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference): `make -C oracle ref` compiles the
reference sources in place into oracle/_ref/, this script runs those binaries and
records what they produce.  The GPU box has no /root/reference; it only reads the
committed fixtures.  (The reference's own README table, README.md:43-52, does not
reproduce with its current sources -- SURVEY.md section 4 -- so it is recorded here
only under "readme_table_not_reproducible".)

    python tests/golden/make_golden.py
"""
import json
import os
import re
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from oracle import oracle as O  # noqa: E402


def f9(x):
    return float("%.9g" % float(x))


def checksum(s):
    return {k: float(np.sum(getattr(s, k).astype(np.float64))) for k in O.State.FIELDS}


def head(s, k=4):
    return {f: [f9(v) for v in getattr(s, f)[:k]] for f in O.State.FIELDS}


def cli_table(ver, n, steps, threads=None):
    exe = os.path.join(O.REF_DIR, ver, "nbody.x")
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    out = subprocess.run([exe, str(n), str(steps)], check=True, capture_output=True, text=True, env=env).stdout
    rows = []
    for line in out.splitlines():
        m = re.match(r"^ (\d+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s*$", line)
        if m:
            rows.append({"s": int(m.group(1)), "t": m.group(2), "kenergy": m.group(3)})
    return rows, out


def main():
    O.build(with_ref=True)
    g = {"generator": "tests/golden/make_golden.py", "compiler": subprocess.run(
        ["/usr/bin/g++", "--version"], capture_output=True, text=True).stdout.splitlines()[0],
        "flag_map": "see oracle/Makefile"}

    # ---- initial conditions (step-0 dumps)
    g["ic"] = {}
    for n in (2000, 16384):
        s, _, _ = O.ref_run("ver2", n, 0)
        g["ic"][str(n)] = {"head": head(s, 6), "sum": checksum(s)}

    # ---- C0: N = 2000
    c0 = {"n": 2000}
    for ver in ("ver0", "ver2", "ver5", "ver7", "ver8"):
        rows, out = cli_table(ver, 2000, 500, threads=4)
        c0["cli_table_" + ver] = rows
        if ver == "ver2":
            # everything but the two timing columns and the summary numbers
            c0["cli_stdout_shape"] = [re.sub(r"\S+\s+\S+\s*$", "", l) if re.match(r"^ \d+", l) else l
                                      for l in out.splitlines() if not l.startswith("# Total") and not l.startswith("# Average")]
    for ver in ("ver0", "ver2", "ver7", "ver8"):
        kes = []
        for S in range(1, 11):
            s, ke, _ = O.ref_run(ver, 2000, S, threads=4)
            kes.append(f9(ke))
        c0["kenergy_steps_1_10_" + ver] = kes
        c0["sum_after_10_" + ver] = checksum(s)
        c0["head_after_10_" + ver] = head(s, 4)
        if ver == "ver2":
            np.savez_compressed(os.path.join(HERE, "c0_ver2_n2000_s10.npz"),
                                **{f: getattr(s, f) for f in O.State.FIELDS})
    g["c0"] = c0

    # ---- C1-sized: N = 16384 (threaded versions), 3 steps
    c1 = {"n": 16384}
    for ver in ("ver2", "ver7", "ver8"):
        kes = []
        for S in (1, 2, 3):
            s, ke, _ = O.ref_run(ver, 16384, S, threads=8)
            kes.append(f9(ke))
        c1["kenergy_steps_1_3_" + ver] = kes
        c1["sum_after_3_" + ver] = checksum(s)
        c1["head_after_3_" + ver] = head(s, 4)
    g["c1"] = c1

    # ---- N = 65536, ver8, 2 steps (largest size the oracle side is asked for)
    c2 = {"n": 65536}
    kes = []
    for S in (1, 2):
        s, ke, _ = O.ref_run("ver8", 65536, S, threads=8)
        kes.append(f9(ke))
    c2["kenergy_steps_1_2_ver8"] = kes
    c2["sum_after_2_ver8"] = checksum(s)
    g["n65536"] = c2

    g["readme_table_not_reproducible"] = {
        "source": "README.md:43-52",
        "kenergy": ["103.29", "440.49", "809.72", "1194.9", "1589.8", "1991.3", "2397.2", "2807.2", "3220.1", "2666.5"],
        "note": "format documentation only; current sources give c0.cli_table_ver0"}

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", os.path.join(HERE, "golden.json"))


if __name__ == "__main__":
    main()
