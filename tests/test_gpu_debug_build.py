"""GPU suite: the bounds-checked build.  compute-sanitizer is closed on the GPU pool, so memory
safety of the step kernel is checked by the library itself: `make DEBUG=1` (libnbx_debug.so, built by
__graft_entry__.build()) range-checks every index the kernel stores through (partials, velocities,
positions, peer replicas, energy slots), every TMA source range and every ticket value on the device;
a failed check surfaces as NBX_ERR_DEBUG with the source line."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sanity_matrix_on_debug_build(pkg):
    lib = os.path.join(pkg.PKG_DIR, "libnbx_debug.so")
    assert os.path.exists(lib), "libnbx_debug.so not built (make -C nbody-demo-2023_b200 DEBUG=1)"
    env = dict(os.environ, NBX_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "sanity_small.py")], capture_output=True, text=True,
                       env=env, timeout=1200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "libnbx_debug.so" in r.stdout and "cases OK" in r.stdout


def test_debug_check_fires_and_reports_the_line(pkg, tmp_path):
    """The checks are live: ask the debug library for more energy slots than the kernel was told about
    (option only the debug build accepts) and the device-side check must surface as NBX_ERR_DEBUG."""
    lib = os.path.join(pkg.PKG_DIR, "libnbx_debug.so")
    code = f"""
import importlib, sys
sys.path.insert(0, {REPO!r})
nbx = importlib.import_module("nbody-demo-2023_b200").nbx
with nbx.Context(3000) as c:
    c.upload(*nbx.ic(3000))
    c.run(2)
    c.set_option("debug_fault", 1)
    try:
        c.run(2)
    except nbx.NbxError as e:
        print("CODE", e.code, str(e))
        try:
            c.run(1)
        except nbx.NbxError as e2:
            print("AGAIN", e2.code)
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, NBX_LIB=lib), timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "CODE 7" in r.stdout and "nbx_kernels.cuh:" in r.stdout and "AGAIN 7" in r.stdout


def test_product_library_has_no_fault_injection(nbx):
    with nbx.Context(64) as c:
        with pytest.raises(nbx.NbxError) as e:
            c.set_option("debug_fault", 1)
        assert "unknown option" in str(e.value)
