"""CPU suite: what the built library contains, read from its SASS and the ptxas log (no GPU needed; cuobjdump ships with the
CUDA toolkit).  Guards the properties DESIGN.md section 4 states about the step kernel against toolchain or source drift."""
import importlib
import os
import re
import shutil
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "nbody-demo-2023_b200")
LIB = os.path.join(PKG, "libnbx.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and a built libnbx.so")


@pytest.fixture(scope="module")
def sass():
    return subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout


def test_sm100a_only_and_native_instructions(sass):
    archs = set(re.findall(r"arch = (sm_\w+)", sass))
    assert archs == {"sm_100a"}, archs
    for mnemonic, what in (("UBLKCP", "1-D TMA bulk copy"), ("SYNCS.ARRIVE.TRANS64", "mbarrier complete_tx"), ("FFMA2", "packed FP32 FMA"),
                           ("FMUL2", "packed FP32 multiply"), ("MUFU.RSQ", "rsqrt on the XU pipe"), ("ACQBULK", "programmatic dependent launch"),
                           ("CCTL.E.RML2", "discard of consumed partials from L2")):
        assert mnemonic in sass, f"{mnemonic} ({what}) not in the library's SASS"
    assert "multimem" in subprocess.run(["cuobjdump", "-ptx", LIB], capture_output=True, text=True).stdout or "STG.E.128.STRONG.SYS" in sass
    for foreign in ("HMMA", "UTCHMMA", "cublas", "triton"):        # nothing on this path is a tensor-core or library kernel
        assert foreign not in sass


def test_inner_loops_match_the_register_bank_count():
    """tools/bank_model.py on the product library: 12-instruction shapes 27 cycles per (i, j-pair), the q-scaled default 25."""
    out = subprocess.run([sys.executable, os.path.join(REPO, "tools", "bank_model.py"), LIB], capture_output=True, text=True, check=True).stdout
    rows = {}
    for line in out.splitlines()[1:]:
        f = line.split()
        rows[f[-1]] = (float(f[0]), float(f[1]), float(f[2]))
    nbx = importlib.import_module("nbody-demo-2023_b200").nbx
    assert len(rows) == len(nbx.variant_names())
    qi = [v for k, v in rows.items() if int(k.strip("<>").split(",")[-1]) & (1 << 22)]
    assert len(qi) == 1 and qi[0] == (25.0, 11.0, 3.0)
    for k, v in rows.items():
        if not int(k.strip("<>").split(",")[-1]) & (1 << 22):
            assert v == (27.0, 12.0, 3.0), (k, v)


def test_no_spills_and_two_ctas_per_sm():
    log = os.path.join(PKG, "build_ptxas.log")
    if not os.path.exists(log):
        pytest.skip("no ptxas log beside the library")
    txt = open(log).read()
    found = 0
    for m in re.finditer(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", txt):
        name, spill_st, spill_ld, regs = m.group(1), int(m.group(3)), int(m.group(4)), int(m.group(5))
        if "step_kernel" not in name:
            continue
        found += 1
        assert spill_st == 0 and spill_ld == 0, name
        threads, minb = (int(x) for x in re.search(r"step_kernelILi\d+ELi(\d+)ELi\d+ELi\d+ELi\d+ELi(\d+)E", name).groups())
        assert regs * threads * minb <= 65536, (name, regs)      # the occupancy the launch bounds ask for fits the register file
    assert found >= 8
