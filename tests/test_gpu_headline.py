"""GPU suite: the HEADLINE configurations against output of the reference itself AND the fp64 truth.

Fixtures (tests/golden/, generators committed beside them):
  large_<case>_ver8.npz   the unmodified reference's fastest CPU version (ver8/GSimulation.cpp:142-215,
                          compiled by oracle/Makefile) run from the same initial conditions: per-step
                          kinetic energy, positions/velocities of 4096 sampled bodies
  truth_<case>_fp64.npz   the same steps with every quantity in double (oracle_run_fp64)

Gates (BASELINE.json north star: kinetic energy within 1e-4 relative per step, positions within 1e-4
relative -- norm-wise over the sampled bodies -- after 10 steps):

  * against the truth: 1e-4, fixed;
  * against the reference's output: 1e-4 + the reference's OWN distance from the truth for that quantity
    (read from the two fixtures, printed).  Up to N = 262 144 that distance is <= 6e-5 and the gate is the
    plain 1e-4.  From N = 1 M on it is not: the reference's single-accumulator float sums are biased low
    (ver8 at N = 1 M: kinetic energy 8.3e-4, positions 5.3e-4 from the truth after two steps; its own
    versions differ from each other by as much), so no implementation -- the reference's other versions
    included -- can be within 1e-4 of that output AND of the exact result.  The triangle bound is what
    "matches the reference within 1e-4" can mean there.
  * a 500-step run (C1) is chaotic: every float implementation drifts from the truth (ver8: 5e-4 in
    kinetic energy, 2e-2 in sampled velocities at step 500).  First 10 steps: the fixed gates; whole run:
    no further from the truth than twice the reference's own distance.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-4


def load(name, kind="large", suffix="ver8"):
    path = os.path.join(GOLDEN_DIR, f"{kind}_{name}_{suffix}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    return np.load(path)


def deviations(ke, pos_sel, vel_sel, fx):
    """(max relative kinetic-energy error over the fixture's steps, rel-l2 of sampled positions, of velocities)."""
    want = fx["ke"].astype(np.float64)
    return (float(np.max(np.abs(np.asarray(ke, dtype=np.float64)[:want.size] - want) / want)),
            rel_l2(pos_sel, fx["pos_sel"]), rel_l2(vel_sel, fx["vel_sel"]))


def check_against_fixtures(ref, truth, ke, out, label="", chaotic=False):
    """`out` = the six downloaded arrays.  Returns the three deviation triples for reporting."""
    sel = ref["sel"]
    assert np.array_equal(sel, truth["sel"])
    pos = np.stack([a[sel] for a in out[:3]], axis=1)
    vel = np.stack([a[sel] for a in out[3:6]], axis=1)
    ref_truth = deviations(ref["ke"], ref["pos_sel"], ref["vel_sel"], truth)
    gpu_truth = deviations(ke, pos, vel, truth)
    gpu_ref = deviations(ke, pos, vel, ref)
    names = ("kenergy", "pos", "vel")
    print(f"\n{label} ({int(ref['steps'])} steps):\n"
          f"   GPU       vs fp64 truth : " + "  ".join(f"{n} {v:.2e}" for n, v in zip(names, gpu_truth)) + "\n"
          f"   reference vs fp64 truth : " + "  ".join(f"{n} {v:.2e}" for n, v in zip(names, ref_truth)) + "\n"
          f"   GPU       vs reference  : " + "  ".join(f"{n} {v:.2e}" for n, v in zip(names, gpu_ref)))
    for n, g_t, r_t, g_r in zip(names, gpu_truth, ref_truth, gpu_ref):
        if chaotic:
            assert g_t < max(TOL, 2.0 * r_t), f"{n}: further from the truth than twice the reference's own distance"
        else:
            assert g_t < TOL, f"{n}: GPU vs fp64 truth"
        assert g_r < TOL + 1.05 * (r_t + (g_t if chaotic else 0.0)), f"{n}: GPU vs the reference's output"
    return gpu_truth, ref_truth, gpu_ref


def run_case(nbx, name, **opts):
    ref = load(name)
    n, steps = int(ref["n"]), int(ref["steps"])
    arrs = nbx.ic(n, str(ref["ic"]))
    with nbx.Context(n) as c:
        for k, v in opts.items():
            c.set_option(k, v)
        c.upload(*arrs)
        ke, secs = c.run(steps)
        out = c.state()
        info = c.info()
    return ref, ke, out, info, secs


def test_n262144_10_steps(nbx):
    """Largest size where the reference itself is within 1e-4 of the truth: the plain north-star gates."""
    ref, ke, out, info, _ = run_case(nbx, "n262144")
    truth = load("n262144", "truth", "fp64")
    g_t, r_t, g_r = check_against_fixtures(ref, truth, ke, out, "N=262144")
    assert int(ref["steps"]) == 10 and max(r_t) < TOL
    assert max(g_r) < 2 * TOL


def test_c2_1m_2_steps(nbx):
    """BASELINE config 2 (N = 1,048,576, uniform cube), the default plan (888 whole tiles + split tail)."""
    ref, ke, out, info, _ = run_case(nbx, "c2")
    assert info["whole_tiles"] > 0 and info["j_splits"] > 1
    g_t, r_t, g_r = check_against_fixtures(ref, load("c2", "truth", "fp64"), ke, out, "C2 N=1M")
    assert g_t[0] < 1e-5 and g_t[1] < 1e-5          # two-level accumulation: far inside the gate


def test_c2_1m_10_steps(nbx):
    """The north star's statement at the headline size itself: N = 1,048,576, kinetic energy per step and
    positions after 10 steps (fixtures: ver8 and the all-double run, ~15 and ~65 min of 8 CPU cores)."""
    ref, ke, out, info, _ = run_case(nbx, "c2s10")
    assert int(ref["steps"]) == 10
    check_against_fixtures(ref, load("c2s10", "truth", "fp64"), ke, out, "C2 N=1M")


def test_c3_4m_plummer_1_step(nbx):
    """BASELINE config 3 (N = 4,194,304 Plummer sphere) on one GPU, one step."""
    ref, ke, out, info, secs = run_case(nbx, "c3")
    check_against_fixtures(ref, load("c3", "truth", "fp64"), ke, out, "C3 N=4M Plummer")
    assert 4.0 < secs < 12.0           # ~16x the 1 M step: throughput is data-independent


def test_c1_first_10_steps(nbx):
    """BASELINE config 1 (N = 16384, j-split + CUDA-graph replay): the fixed gates after 10 steps."""
    ref, ke, out, info, _ = run_case(nbx, "c1s10")
    assert info["j_splits"] > 1 and info["use_graph"] == 1
    g_t, r_t, g_r = check_against_fixtures(ref, load("c1s10", "truth", "fp64"), ke, out, "C1 N=16384, 10 steps")
    assert max(g_r) < TOL


def test_c1_whole_run_500_steps(nbx):
    """All 500 steps of config 1: every kinetic energy and the final sampled state (chaotic regime)."""
    ref, ke, out, info, _ = run_case(nbx, "c1")
    assert info["kernel_launches"] == 500
    check_against_fixtures(ref, load("c1", "truth", "fp64"), ke, out, "C1 N=16384, 500 steps", chaotic=True)
    # the first 100 steps are not chaotic yet: fixed gate on every kinetic energy against the reference
    want = ref["ke"][:100].astype(np.float64)
    assert np.max(np.abs(ke[:100] - want) / want) < TOL


@pytest.mark.parametrize("opts", [dict(accurate=1), dict(variant=3)], ids=["acc64", "single_float_accumulator"])
def test_other_accumulation_modes_on_c2(nbx, opts):
    """The fp64-fold option sits on the truth; the single-accumulator shape (the reference's own summation
    scheme, round 1's default) is closer to the truth than the reference but outside 1e-4 at step 2."""
    ref, ke, out, info, _ = run_case(nbx, "c2", **opts)
    truth = load("c2", "truth", "fp64")
    sel = ref["sel"]
    pos = np.stack([a[sel] for a in out[:3]], axis=1)
    vel = np.stack([a[sel] for a in out[3:6]], axis=1)
    g_t = deviations(ke, pos, vel, truth)
    r_t = deviations(ref["ke"], ref["pos_sel"], ref["vel_sel"], truth)
    print(f"\nC2 {opts}: vs truth kenergy {g_t[0]:.2e} pos {g_t[1]:.2e} vel {g_t[2]:.2e} (reference: {r_t[0]:.2e} {r_t[1]:.2e} {r_t[2]:.2e})")
    if "accurate" in opts:
        assert max(g_t) < 1e-5
    else:
        assert all(g < 0.5 * r for g, r in zip(g_t, r_t))


def test_pdl_multi_wave_bitwise(nbx):
    """Programmatic dependent launch on the configuration that uses it by default (many waves,
    co-resident CTAs of consecutive steps): N = 262144 -> 256 tiles x 15 splits = 3840 CTAs.
    pdl=1 and pdl=0 must agree bit for bit (ticket reuse and the plain i-body loads are what
    this guards)."""
    n = 262144
    arrs = nbx.ic(n)
    res = []
    for pdl in (0, 1, 1):
        with nbx.Context(n) as c:
            c.set_option("pdl", pdl)
            c.upload(*arrs)
            ke, _ = c.run(6)
            info = c.info()
            res.append((ke, c.state()))
        assert info["i_tiles"] * info["j_splits"] >= 600
    for ke, st in res[1:]:
        assert np.array_equal(ke, res[0][0])
        for a, b in zip(st, res[0][1]):
            assert np.array_equal(a, b)


def test_single_wave_pdl_bitwise(nbx):
    """A grid that fits the SMs once (N = 16384: 144 CTAs) runs with programmatic dependent launch and one CTA
    per SM by default; forcing pdl off must not change a bit."""
    arrs = nbx.ic(16384)
    res = []
    for pdl in (-1, 0):
        with nbx.Context(16384) as c:
            c.set_option("pdl", pdl)
            c.upload(*arrs)
            ke, _ = c.run(40)
            res.append((ke, c.state(), c.info()["ctas_per_sm"]))
    assert res[0][2] == 1 and res[1][2] >= 2
    assert np.array_equal(res[0][0], res[1][0])
    for a, b in zip(res[0][1], res[1][1]):
        assert np.array_equal(a, b)
