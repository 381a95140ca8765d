"""CPU suite, part 2: the C-ABI library loads, exports what include/nbx.h declares,
the host-side logic is right, and the product fails LOUDLY without a GPU (no fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu(nbx):
    return nbx.device_count() == 0


def test_library_exports_every_declared_symbol(pkg, nbx):
    hdr = open(os.path.join(REPO, "include", "nbx.h")).read()
    declared = sorted(set(re.findall(r"NBX_API[^;]*?\b(nbx_\w+)\s*\(", hdr)))
    assert len(declared) >= 25
    assert sorted(nbx.SYMBOLS) == declared, "nbx.py SYMBOLS out of sync with include/nbx.h"
    L = ctypes.CDLL(pkg.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert L.nbx_abi_version() == 2


def test_every_entry_point_cites_the_reference(pkg):
    hdr = open(os.path.join(REPO, "include", "nbx.h")).read()
    assert hdr.count("GSimulation.cpp:") + hdr.count("Compute.cu:") + hdr.count("Compute.cpp:") >= 10


def test_no_oracle_in_product_path(pkg):
    # the product may never import / link / call anything under oracle/
    for root, _, files in os.walk(pkg.PKG_DIR):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or fn == "Makefile":
                src = open(os.path.join(root, fn), errors="replace").read()
                for line in src.splitlines():
                    code = line.split("//")[0].split("#")[0] if not fn.endswith(".py") else line.split("#")[0]
                    assert "liboracle" not in code and "oracle_run" not in code and "from oracle" not in code, (fn, line)
    out = subprocess.run(["ldd", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_ic_matches_oracle_and_fixture(nbx, oracle, golden):
    for n in (1, 7, 2000, 16384):
        got = nbx.ic(n)
        want = oracle.ic_uniform(n).arrays()
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
    got = nbx.ic(2000)
    fx = golden["ic"]["2000"]
    for a, f in zip(got, oracle.State.FIELDS):
        assert float(np.sum(a.astype(np.float64))) == fx["sum"][f]


def test_plummer_ic_properties(nbx):
    px, py, pz, vx, vy, vz, m = nbx.ic(20000, "plummer")
    r = np.sqrt(px.astype(np.float64) ** 2 + py ** 2 + pz ** 2)
    assert r.max() <= 10.0 + 1e-5
    # half-mass radius of a Plummer sphere with a = 1 is 1.305 (slightly less when cut at 10a)
    assert 1.2 < np.median(r) < 1.4
    assert abs(px.mean()) < 0.05 and abs(py.mean()) < 0.05 and abs(pz.mean()) < 0.05
    u = nbx.ic(20000, "uniform")
    assert np.array_equal(vx, u[3]) and np.array_equal(m, u[6])   # velocities, masses: reference sequence


def test_flop_convention(nbx):
    assert nbx.gflop_per_step(2000) == pytest.approx(1e-9 * (29 * 2000.0 ** 2 + 19 * 2000.0))
    assert nbx.gflop_per_step(1 << 20) == pytest.approx(1e-9 * (29 * float(1 << 20) ** 2 + 19 * float(1 << 20)))


def test_shard_arithmetic(nbx):
    for n in (1, 8, 2000, 16384, 1 << 20, (1 << 22) + 3):
        for world in (1, 2, 4, 8):
            spans = [nbx.shard_of(n, r, world) for r in range(world)]
            n_pad = spans[0][2]
            assert n_pad >= n and n_pad % (8 * world) == 0 and n_pad - n < 8 * world
            assert spans[0][0] == 0
            for a, b in zip(spans, spans[1:]):
                assert a[0] + a[1] == b[0]
            assert spans[-1][0] + spans[-1][1] == n_pad
            assert all(s[1] % 8 == 0 for s in spans)


def test_variant_table(nbx):
    names = nbx.variant_names()
    assert len(names) >= 4 and len(set(names)) == len(names)


def test_bad_arguments_are_rejected(nbx):
    with pytest.raises(nbx.NbxError) as e:
        nbx.Context(0)
    assert e.value.code == 1
    with pytest.raises(nbx.NbxError):
        nbx.Context(100, rank=2, world=2)
    with pytest.raises(nbx.NbxError):
        nbx.Context(100, eps2=0.0)
    with pytest.raises(nbx.NbxError):
        nbx.Context(100, world=9)


def test_fails_loudly_without_gpu(pkg, nbx):
    if not _no_gpu(nbx):
        pytest.skip("a GPU is visible here")
    with pytest.raises(nbx.NbxError) as e:
        nbx.Context(128)
    assert e.value.code == 5 and "no CPU fallback" in str(e.value)
    a = nbx.ic(64)
    with pytest.raises(nbx.NbxError):
        nbx.simulate(1, *a)
    r = subprocess.run([pkg.CLI_PATH, "64", "2"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr
    # the banner is printed by the constructor before start() fails, as in the reference
    assert r.stdout.splitlines()[:2] == ["===============================", " Initialize Gravity Simulation"]


def test_ver5_all_cli_refuses_cpu_share(pkg):
    r = subprocess.run([pkg.CLI_ALL_PATH, "64", "2", "cpu"], capture_output=True, text=True)
    assert r.returncode == 1 and "GPU-only" in r.stderr
    assert r.stdout.splitlines()[0] == "cpu"          # ver5_all/main.cpp:42 echoes the selector first


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 with no C++ or CUDA types, and a
    C program must link against libnbx.so (symbols unmangled)."""
    import importlib
    pkg = importlib.import_module("nbody-demo-2023_b200")
    hdr = os.path.join(REPO, "include", "nbx.h")
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = tmp_path / "t.c"
    src.write_text('#include "nbx.h"\n#include <stdio.h>\nint main(void){ int n = -1; nbx_ctx *c = 0;'
                   ' if (nbx_abi_version() != NBX_ABI_VERSION) return 2; if (nbx_device_count(&n)) return 3;'
                   ' if (n == 0 && nbx_create(&c, 16, 0, 0, 1, 0.1f, 6.67259e-11f, 1e-3f) != NBX_ERR_NODEVICE) return 4;'
                   ' if (c) nbx_destroy(c); printf("%d %d %s\\n", n, (int)sizeof(nbx_info), nbx_last_error()); return 0; }\n')
    exe = tmp_path / "t"
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe),
                        "-L", pkg.PKG_DIR, "-lnbx", "-Wl,-rpath," + pkg.PKG_DIR], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    # the ctypes mirror of nbx_info has the C layout (guards against ABI drift between include/nbx.h and nbx.py)
    assert int(r.stdout.split()[1]) == ctypes.sizeof(pkg.nbx.Info)


def test_launch_plan_host_logic(nbx):
    """nbx_plan is the host-side decomposition nbx_run uses (no device needed)."""
    names = nbx.variant_names()
    # BASELINE configs on a 148-SM B200
    c0 = nbx.plan(2000)
    assert names[c0["variant"]] == "r2_t128_u4" and c0["i_tiles"] == 8 and c0["whole_tiles"] == 0 and c0["j_splits"] > 1
    assert c0["use_graph"] == 1 and c0["n_pad"] == 2000
    c1 = nbx.plan(16384)
    assert len(names) == 8, "ablation shapes must not ship in the product library"
    assert names[c1["variant"]] == "r4_t256_u4_stage_f2" and (c1["i_tiles"], c1["whole_tiles"], c1["j_splits"]) == (16, 0, 9)
    assert c1["use_graph"] == 1
    c2 = nbx.plan(1 << 20)
    assert names[c2["variant"]] == "r4_t256_u4_stage_f2_qi" and names[nbx.plan(65536)["variant"]] == "r4_t256_u4_stage_f2_qi"
    # ... but not by itself in the NCCL-overlap mode (two launches per step)
    assert names[nbx.plan(1 << 22, rank=3, world=8, exchange=2)["variant"]] == "r4_t256_u4_stage_f2"
    assert names[nbx.plan(1 << 22, rank=3, world=8, exchange=1)["variant"]] == "r4_t256_u4_stage_f2_qi"
    assert (c2["i_tiles"], c2["whole_tiles"]) == (1024, 888) and c2["j_splits"] >= 13 and c2["use_graph"] == 0
    mid = nbx.plan(262144)                       # 256 tiles = 1.7 SM rounds: no unsplit tiles
    assert mid["whole_tiles"] == 0 and mid["j_splits"] > 1
    for r in range(8):                           # C3 strong scaling, 8 ranks
        c3 = nbx.plan(1 << 22, rank=r, world=8, exchange=nbx.EXCHANGE_P2P)
        assert c3["i_begin"] == r * (1 << 19) and c3["i_count"] == 1 << 19
        assert (c3["i_tiles"], c3["whole_tiles"]) == (512, 444)
    ov = nbx.plan(1 << 22, rank=3, world=8, exchange=nbx.EXCHANGE_NCCL_OVERLAP)
    assert ov["whole_tiles"] == 0 and ov["j_splits"] >= 2 and ov["use_graph"] == 0
    # invariants over many sizes / worlds / SM counts
    for sm in (148, 132, 64):
        for world in (1, 2, 4, 8):
            for n in (1, 9, 1000, 8191, 8192, 100000, 151552, 155648, 303104, 454656, 1 << 20, (1 << 22) + 5):
                p = nbx.plan(n, rank=world - 1, world=world, sm_count=sm)
                bi = p["threads"] * p["bodies_per_thread"]
                assert p["n_pad"] % (8 * world) == 0 and 0 <= p["n_pad"] - n < 8 * world
                assert p["i_tiles"] == -(-p["i_count"] // bi)
                assert p["whole_tiles"] % sm == 0 or p["whole_tiles"] == p["i_tiles"]
                assert p["whole_tiles"] in (0, p["i_tiles"]) or p["whole_tiles"] // sm >= 3
                assert 1 <= p["j_splits"] <= max(1, p["n_pad"] // 8)
                if p["j_splits"] == 1:
                    assert p["whole_tiles"] == p["i_tiles"]
    forced = nbx.plan(1 << 20, j_splits=4)
    assert forced["whole_tiles"] == 0 and forced["j_splits"] == 4      # a pinned split applies to every tile
    with pytest.raises(nbx.NbxError):
        nbx.plan(100, exchange=7)
    with pytest.raises(nbx.NbxError):
        nbx.plan(100, j_splits=100000)
    with pytest.raises(nbx.NbxError):
        nbx.plan(0)
    with pytest.raises(nbx.NbxError):
        nbx.plan(100, rank=2, world=2)
