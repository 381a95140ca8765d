"""nbody-demo-2023_b200 -- B200-native backend for the O(N^2) force / Euler /
kinetic-energy path of NTHU-SC/nbody-demo-2023.

The product is native: `libnbx.so` (CUDA kernels for sm_100a behind the C ABI in
include/nbx.h) and `nbody.x` (the reference's GSimulation / CLI surface on top of
it).  This package only builds them (`build()`) and exposes a thin ctypes mirror of
the C ABI (`nbx`) for tests and bench.py.  The directory name has a hyphen, so
import it with importlib:

    import importlib; pkg = importlib.import_module("nbody-demo-2023_b200")
"""
from __future__ import annotations

import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libnbx.so")
CLI_PATH = os.path.join(PKG_DIR, "nbody.x")
CLI_ALL_PATH = os.path.join(PKG_DIR, "nbody_all.x")   # ver5_all-style argv


def build(verbose: bool = False) -> None:
    """Compile libnbx.so (nvcc, sm_100a) and nbody.x in-tree.  Needs no GPU."""
    cmd = ["make", "-C", PKG_DIR, "all"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.run(cmd, check=True)


from . import nbx  # noqa: E402  (ctypes mirror; loading the .so is deferred to first use)

__all__ = ["build", "nbx", "LIB_PATH", "CLI_PATH", "CLI_ALL_PATH", "PKG_DIR", "REPO_DIR"]
