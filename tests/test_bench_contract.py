"""bench.py prints ONE JSON line with the keys the driver reads.  The reference arm (the
reference's own CPU implementation, oracle/_ref) runs anywhere; the native arm needs a B200."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
          "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, env=None, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py")] + args, capture_output=True, text=True,
                       timeout=timeout, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert COMMON <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "pair_interactions_per_second" and d["unit"] == "G pair-interactions/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.gpu
def test_native_arm_line():
    d = _run(["--steps", "1", "--warmup", "3", "--no-cpu-baseline", "--no-extras"])
    assert (COMMON - {"cpu_baseline"}) <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 3 and d["dtype"] == "f32"
    assert d["gpu_launches_detail"]["step_kernel"] == 1 and d["gpu_launches"] == 1 + d["gpu_launches_detail"]["qscale_kernel"]
    rf = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf)
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-3 and 0.5 < rf["frac"] < 0.84
    assert abs(rf["achieved"] - 20e-3 * d["value"]) < 0.01 * rf["achieved"]
    assert d["e2e"]["h2d_bytes_per_step"] == 7 * 4 * (1 << 20) and d["e2e"]["value"] > 0.9 * d["value"]
    assert d["clocks"]["sm_max_mhz"] and "reasons" in d["clocks"]
    assert d["config"]["n_bodies"] == 1 << 20
    # correctness of what was timed rides on the line: reference output, fp64 sums, sampled forces
    p = d["parity"]
    v = p["vs_reference_output"]
    assert p["ok"] is True and v["ok"] is True and v["gpu_vs_fp64_truth"]["kenergy_max_rel"] < 1e-4
    assert v["gpu_vs_reference"]["kenergy_max_rel"] < 1e-4 + 1.05 * v["reference_vs_fp64_truth"]["kenergy_max_rel"]
    assert p["sampled_forces"]["ok"] is True and p["kenergy_vs_fp64_sum"] < 1e-6


@pytest.mark.gpu
def test_native_arm_default_line_carries_the_other_configs():
    """The default 1-GPU run also times C3 (the strong-scaling anchor), C1 and C0 and the reference's CUDA
    backend, each outside the headline's timed region."""
    d = _run(["--steps", "2", "--warmup", "3", "--no-cpu-baseline"], timeout=900)
    a = d["config"]["strong_anchor"]
    assert a["workload"].startswith("C3") and 4000 < a["ms_per_step"] < 12000 and a["parity"]["ok"] is True
    c1, c0 = d["also"]["c1"], d["also"]["c0"]
    assert c1["n_bodies"] == 16384 and 0.05 < c1["ms_per_step"] < 0.2 and c1["parity"]["ok"] is True
    assert c1["parity"]["vs_reference_output"]["steps"] == 500
    assert c0["n_bodies"] == 2000 and c0["ms_per_step"] < 0.05 and c0["parity"]["ok"] is True
    g = d["gpu_reference"]
    assert g["kernel_only"]["value"] >= g["end_to_end"]["value"] > 0
    assert g["this_build_same_n"]["kernel_only"] > g["kernel_only"]["value"]
