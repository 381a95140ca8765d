"""CPU suite, part 1: the oracle is pinned.

The oracle (oracle/nbody_oracle.c) is checked against (a) the committed fixtures that
tests/golden/make_golden.py generated from the compiled, unmodified reference and
(b) -- where /root/reference exists, i.e. in the build container -- the compiled
reference itself, bit for bit for ver0 and ver2.
"""
import os

import numpy as np
import pytest

from conftest import rel_l2

HAVE_REF = os.path.isdir("/root/reference")


def test_ic_matches_reference_fixture(oracle, golden):
    for n_str, fx in golden["ic"].items():
        n = int(n_str)
        s = oracle.ic_uniform(n)
        for f in oracle.State.FIELDS:
            got = getattr(s, f)
            assert [float("%.9g" % v) for v in got[:6]] == fx["head"][f], (n, f)
            assert float(np.sum(got.astype(np.float64))) == fx["sum"][f], (n, f)


def test_ic_known_first_draws(oracle):
    # first draws of libstdc++'s uniform_real_distribution<float> on mt19937(42)
    s = oracle.ic_uniform(8)
    assert np.float32(s.px[0]) == np.float32(0.37454012)
    assert np.float32(s.py[0]) == np.float32(0.796543002)
    assert np.float32(s.pz[0]) == np.float32(0.95071429)
    assert np.float32(s.vx[0]) == np.float32(-0.00025091978)
    assert np.float32(s.mass[0]) == np.float32(8 * 0.37454012)


@pytest.mark.parametrize("variant", ["ver0", "ver2"])
def test_serial_oracle_bit_exact_vs_fixture(oracle, golden, golden_c0_state, variant):
    s = oracle.ic_uniform(2000)
    ke = oracle.run(s, 10, variant=variant)
    want = golden["c0"]["kenergy_steps_1_10_" + variant]
    assert [float("%.9g" % k) for k in ke] == want
    sums = golden["c0"]["sum_after_10_" + variant]
    for f in oracle.State.FIELDS:
        assert float(np.sum(getattr(s, f).astype(np.float64))) == sums[f], f
    if variant == "ver2":
        for f in oracle.State.FIELDS:
            assert np.array_equal(getattr(s, f), golden_c0_state[f]), f


def test_threaded_oracle_close_to_reference_versions(oracle, golden):
    # ver7/ver8 reorder the sums (simd lanes, threads): the reference's own spread is ~1e-6
    s = oracle.ic_uniform(2000)
    ke = oracle.run(s, 10, variant="ver7")
    for ver in ("ver2", "ver7", "ver8"):
        want = np.array(golden["c0"]["kenergy_steps_1_10_" + ver])
        assert np.max(np.abs(ke - want) / want) < 3e-6, ver
    s = oracle.ic_uniform(16384)
    ke = oracle.run(s, 3, variant="ver7")
    for ver in ("ver2", "ver7", "ver8"):
        want = np.array(golden["c1"]["kenergy_steps_1_3_" + ver])
        assert np.max(np.abs(ke - want) / want) < 5e-6, ver
    for f in ("px", "py", "pz"):
        assert abs(float(np.sum(getattr(s, f).astype(np.float64))) - golden["c1"]["sum_after_3_ver2"][f]) < 1e-4


def test_cli_table_fixture_is_version_independent(golden):
    # the 5-digit kenergy column is identical for every reference version
    base = [r["kenergy"] for r in golden["c0"]["cli_table_ver2"]]
    assert base == ["0.1432", "2.4341", "8.1256", "17.877", "32.966", "55.786", "91.132", "150.12", "264.78", "571.53"]
    for ver in ("ver0", "ver5", "ver7", "ver8"):
        assert [r["kenergy"] for r in golden["c0"]["cli_table_" + ver]] == base, ver


def test_fp64_truth_agrees_with_float_oracle(oracle):
    s = oracle.ic_uniform(4096)
    sel = np.arange(0, 4096, 37, dtype=np.int32)
    a64 = oracle.acc_fp64(s, sel)
    s1 = s.copy()
    oracle.run(s1, 1, variant="ver2")
    a32 = (s1.vel()[sel].astype(np.float64) - s.vel()[sel].astype(np.float64)) / np.float64(np.float32(0.1))
    # v' = v + a*dt in float: recovering a loses digits to v's ulp; compare norm-wise
    assert rel_l2(a32, a64) < 2e-3
    ke = oracle.kenergy_fp64(s1)
    ke32 = oracle.run(s.copy(), 1, variant="ver2")[0]
    assert abs(ke - ke32) / ke < 1e-5


def test_float_sampled_accelerations_match_serial_oracle(oracle):
    # oracle_acc_f32 is the force half of oracle_run_ver2: v' - v == a*dt exactly when v == 0
    s = oracle.ic_uniform(3000)
    s.vx[:] = 0; s.vy[:] = 0; s.vz[:] = 0
    sel = np.arange(0, 3000, 7, dtype=np.int32)
    a = oracle.acc_f32(s, sel)
    s1 = s.copy()
    oracle.run(s1, 1, variant="ver2")
    dt = np.float32(0.1)
    assert np.array_equal(s1.vx[sel], a[:, 0] * dt)                       # vel.x: unfused mul
    assert np.array_equal(s1.vy[sel], np.float32(0) + a[:, 1] * dt)       # fma(a, dt, 0) == a*dt


def test_flop_convention(oracle):
    # ver0/GSimulation.cpp:122
    assert oracle.gflop_per_step(2000) == pytest.approx(1e-9 * (29 * 2000.0 ** 2 + 19 * 2000.0))


def test_edge_sizes(oracle):
    for n in (1, 2, 3, 9):
        s = oracle.ic_uniform(n)
        ke = oracle.run(s, 2, variant="ver2")
        assert np.all(np.isfinite(ke)) and np.all(np.isfinite(s.pos()))
    s = oracle.ic_uniform(5)
    assert oracle.run(s, 0).size == 0


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("variant,n,steps", [("ver0", 2000, 3), ("ver2", 2000, 3), ("ver2", 777, 5), ("ver0", 64, 20)])
def test_oracle_bit_exact_vs_compiled_reference(oracle, variant, n, steps):
    ref, ke_ref, _ = oracle.ref_run(variant, n, steps)
    s = oracle.ic_uniform(n)
    ke = oracle.run(s, steps, variant=variant)
    assert ke[-1] == ke_ref
    for f in oracle.State.FIELDS:
        assert np.array_equal(getattr(s, f), getattr(ref, f)), f


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("ver,n,steps", [("ver8", 4096, 5), ("ver7", 3000, 4), ("ver2", 1000, 3)])
def test_reference_from_arbitrary_state_equals_its_own_run(oracle, ver, n, steps):
    """ref_state_verN (the unmodified reference step loop fed an input state, one start() per step)
    must reproduce ref_dump_verN (the reference run as it is) bit for bit from the same ICs."""
    s0 = oracle.ic_uniform(n)
    a, ke, _ = oracle.ref_state_run(ver, s0, steps, threads=4)
    b, ke_b, _ = oracle.ref_run(ver, n, steps, threads=4)
    for f in oracle.State.FIELDS:
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert ke.size == steps and abs(float(ke[-1]) - float(ke_b)) <= 2e-7 * float(ke_b)   # omp reduction order
    # and restarting from a dumped state continues the run exactly
    mid, _, _ = oracle.ref_state_run(ver, s0, 2, threads=4)
    end, _, _ = oracle.ref_state_run(ver, mid, steps - 2, threads=4)
    for f in oracle.State.FIELDS:
        assert np.array_equal(getattr(end, f), getattr(a, f)), f


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference only exists in the build container")
def test_reference_runs_plummer_state_and_oracle_agrees(oracle, nbx):
    """Inputs the reference cannot generate itself (Plummer ICs) reach it through ref_state_ver8; the C
    restatement (ver7 order) agrees with it to the reference's own version-to-version noise."""
    n = 4096
    arrs = nbx.ic(n, "plummer")
    s0 = oracle.State(n)
    for f, a in zip(oracle.State.FIELDS, arrs):
        setattr(s0, f, a.copy())
    ref, ke_ref, _ = oracle.ref_state_run("ver8", s0, 3, threads=4)
    s = s0.copy()
    ke = oracle.run(s, 3, variant="ver7")
    assert np.max(np.abs(ke - ke_ref) / ke_ref) < 5e-6
    assert rel_l2(s.pos(), ref.pos()) < 1e-6


def test_large_fixtures_are_consistent_with_the_small_ones(golden):
    """tests/golden/large_*_ver8.npz (make_golden_large.py) against golden.json (make_golden.py): the first
    three kinetic energies of the 16384-body run were recorded by both generators."""
    import os
    from conftest import GOLDEN_DIR
    fx = np.load(os.path.join(GOLDEN_DIR, "large_c1_ver8.npz"))
    want = np.array(golden["c1"]["kenergy_steps_1_3_ver8"])
    assert np.max(np.abs(fx["ke"][:3] - want) / want) < 1e-6
    assert int(fx["n"]) == 16384 and int(fx["steps"]) == 500 and fx["ke"].size == 500 and fx["sel"].size == 4096
    for name, n, steps in (("n262144", 262144, 10), ("c2", 1 << 20, 2)):
        f = np.load(os.path.join(GOLDEN_DIR, f"large_{name}_ver8.npz"))
        assert int(f["n"]) == n and f["ke"].size == steps and f["pos_sel"].shape == (4096, 3)
        assert np.all(np.diff(f["sel"]) > 0) and np.all(np.isfinite(f["ke"]))


def _dev(fx, tr):
    ke = float(np.max(np.abs(fx["ke"].astype(np.float64) - tr["ke"]) / tr["ke"]))
    return ke, rel_l2(fx["pos_sel"], tr["pos_sel"]), rel_l2(fx["vel_sel"], tr["vel_sel"])


def test_reference_output_vs_fp64_truth_is_as_documented():
    """The claim the large-N parity gates rest on, checkable without a GPU: the reference's own
    float output (ver8 fixtures) is within 1e-4 of the all-double run up to N = 262 144 and NOT beyond."""
    import os
    from conftest import GOLDEN_DIR

    def pair(name):
        return (np.load(os.path.join(GOLDEN_DIR, f"large_{name}_ver8.npz")), np.load(os.path.join(GOLDEN_DIR, f"truth_{name}_fp64.npz")))
    ke, pos, vel = _dev(*pair("c1s10"))
    assert ke < 2e-6 and pos < 1e-6 and vel < 2e-6
    ke, pos, vel = _dev(*pair("n262144"))
    assert 2e-5 < ke < 1e-4 and pos < 1e-4 and vel < 1e-4
    ke, pos, vel = _dev(*pair("c2"))
    assert 5e-4 < ke < 1.2e-3 and 3e-4 < pos < 8e-4 and 3e-4 < vel < 8e-4          # 8.3e-4 / 5.3e-4 / 5.1e-4
    ke, pos, vel = _dev(*pair("c2s10"))                                            # ... and stays there over 10 steps
    assert 5e-4 < ke < 1.2e-3 and 3e-4 < pos < 8e-4 and 3e-4 < vel < 8e-4          # 8.6e-4 / 5.3e-4 / 5.3e-4
    ke, pos, vel = _dev(*pair("c3"))
    assert 1.5e-3 < ke < 4e-3 and 6e-4 < pos < 2e-3 and 8e-4 < vel < 2e-3          # 2.5e-3 / 1.1e-3 / 1.3e-3
    ke, pos, vel = _dev(*pair("c1"))                                               # 500 steps: chaotic
    assert 2e-4 < ke < 1e-3 and vel > 5e-3
    for name in ("c1s10", "n262144", "c2", "c2s10", "c3", "c1"):
        fx, tr = pair(name)
        assert np.array_equal(fx["sel"], tr["sel"]) and int(fx["n"]) == int(tr["n"]) and int(fx["steps"]) == int(tr["steps"])


def test_truth_fixture_is_reproducible(oracle):
    """tests/golden/truth_c1s10_fp64.npz regenerated here with oracle_run_fp64 (2 s): same numbers."""
    import os
    from conftest import GOLDEN_DIR
    tr = np.load(os.path.join(GOLDEN_DIR, "truth_c1s10_fp64.npz"))
    s0 = oracle.ic_uniform(int(tr["n"]))
    pos, vel, ke = oracle.run_fp64(s0, int(tr["steps"]))
    assert np.max(np.abs(ke - tr["ke"]) / tr["ke"]) < 1e-12
    assert rel_l2(pos[tr["sel"]], tr["pos_sel"]) < 1e-12 and rel_l2(vel[tr["sel"]], tr["vel_sel"]) < 1e-12
    # and the float oracle (the reference's ver2 arithmetic) sits where the reference's fixture sits
    s = s0.copy()
    ke32 = oracle.run(s, int(tr["steps"]), variant="ver2")
    assert np.max(np.abs(ke32 - tr["ke"]) / tr["ke"]) < 5e-6


def test_qscaled_pair_rounding_emulation():
    """The arithmetic of the q-scaled pair (the kernel's default from 65 536 bodies on), emulated in numpy with one rounding
    per FP32 instruction: its extra rounding of q*r_j costs a factor ~10 at N = 2000 and nothing once the sum is long."""
    import qscale_emulation as Q
    cur, new = Q.errors(2000, "uniform", 48)
    assert np.all(np.isfinite(new)) and np.median(cur) < 1e-7 and np.median(new) < 1e-6 and new.max() < 1e-5
    cur, new = Q.errors(65536, "uniform", 12)
    assert np.median(new) < 1e-7 and new.max() < 1e-6
