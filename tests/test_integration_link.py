"""The drop-in boundary, proven the way the reference itself selects a backend: LINK-TIME
substitution of GSimulation::start() (ver5_all/Makefile:1-104; ver5_all/GSimulation.cpp:24-235 defines
every member except start()).  integration/b200/Compute.cpp is compiled together with the reference's
UNMODIFIED ver5_all/main.cpp and ver5_all/GSimulation.cpp and linked against libnbx.so
(integration/Makefile).  CPU side: it builds and fails loudly without a GPU.  GPU side: the binary
the build container produced prints the reference's table."""
import os
import re
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(REPO, "integration", "_build", "nbody_ref_b200.x")
HAVE_REF = os.path.isdir("/root/reference")


def test_backend_tu_is_complete():
    src = open(os.path.join(REPO, "integration", "b200", "Compute.cpp")).read()
    assert "void GSimulation ::start()" in src
    assert "..." not in src.split("void GSimulation ::start()")[1]        # no elisions: it is the whole function
    for call in ("nbx_create", "nbx_upload_group", "nbx_run_group", "nbx_download", "nbx_destroy", "print_header()", "print_stats()", "print_flops()"):
        assert call in src, call


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference only exists in the build container")
def test_links_against_the_unmodified_reference_tree(pkg, nbx):
    r = subprocess.run(["make", "-C", os.path.join(REPO, "integration"), "-B"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # exactly the reference's own sources, from where they lie
    assert "/root/reference/ver5_all/main.cpp" in r.stdout and "/root/reference/ver5_all/GSimulation.cpp" in r.stdout
    assert os.path.exists(EXE)
    syms = subprocess.run(["nm", "-C", "--defined-only", EXE], capture_output=True, text=True).stdout
    assert "GSimulation::start()" in syms and "GSimulation::print_stats()" in syms      # ours + the reference's, one binary
    und = subprocess.run(["nm", "-C", "--undefined-only", EXE], capture_output=True, text=True).stdout
    assert "nbx_run_group" in und and "cuda" not in und.lower()                      # only the C ABI crosses
    if nbx.device_count() == 0:
        r = subprocess.run([EXE, "64", "2"], capture_output=True, text=True)
        assert r.returncode == 1 and "no CUDA device" in r.stderr
        assert r.stdout.splitlines()[:2] == ["===============================", " Initialize Gravity Simulation"]


@pytest.mark.gpu
def test_reference_main_with_b200_backend_prints_the_reference_table(golden):
    if not os.path.exists(EXE):
        pytest.skip("integration/_build/nbody_ref_b200.x not built (needs /root/reference at build time)")
    r = subprocess.run([EXE, "2000", "500"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    shape = golden["c0"]["cli_stdout_shape"]
    assert lines[:6] == shape[:6]                      # banner, header, rules: the reference's own code printed them
    rows = [l for l in lines if re.match(r"^ \d+", l)]
    assert len(rows) == 10
    for l, row in zip(rows, golden["c0"]["cli_table_ver2"]):
        f = l.split()
        assert int(f[0]) == row["s"] and f[1] == row["t"]
        assert abs(float(f[2]) - float(row["kenergy"])) / float(row["kenergy"]) < 2e-4
    assert "# Number Threads     : 1" in lines and any(l.startswith("# Average Perfomance : ") for l in lines)
    # ver5_all argv (ver5_all/main.cpp:35-54): nSteps whenever argc > 2, device string echoed first
    r = subprocess.run([EXE, "2000", "100", "gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.splitlines()[0] == "gpu"
    assert len([l for l in r.stdout.splitlines() if re.match(r"^ \d+", l)]) == 2
