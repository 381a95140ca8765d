"""GPU suite, multi-GPU part (skipped on a 1-GPU box): i-sharded runs against the single-GPU
run.  With the j-split count pinned, every body's force is the same instruction sequence on
any rank, so positions and velocities must agree BIT FOR BIT; kinetic energy is a sum of
per-shard sums and may differ in the last bits."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu(nbx):
    return nbx.device_count()


def _single(nbx, arrs, steps, splits):
    with nbx.Context(arrs[0].shape[0]) as c:
        c.set_option("j_splits", splits)
        c.set_option("graph", 0)
        c.upload(*arrs)
        ke, _ = c.run(steps)
        return ke, c.state()


XCH = {"nccl": 0, "p2p": 1, "nccl_overlap": 2}


@pytest.mark.parametrize("exchange", ["nccl", "p2p", "nccl_overlap"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_one_process_group_matches_single_gpu(nbx, world, exchange):
    if _ngpu(nbx) < world:
        pytest.skip(f"needs {world} GPUs")
    n, steps, splits = 6144, 6, 3      # multiple of 8*world for every world: same padding as 1 GPU
    arrs = nbx.ic(n)
    ke1, st1 = _single(nbx, arrs, steps, splits)
    ctxs = [nbx.Context(n, device=g, rank=g, world=world) for g in range(world)]
    try:
        for c in ctxs:
            c.set_option("j_splits", splits)
            c.set_option("exchange", XCH[exchange])
            c.upload(*arrs)
        nbx.comm_init_all(ctxs)
        if exchange == "p2p":
            blobs = b"".join(c.p2p_export() for c in ctxs)
            for c in ctxs:
                c.p2p_attach(blobs)
        ke, secs = nbx.run_group(ctxs, steps)
        out = [np.zeros(n, dtype=np.float32) for _ in range(6)]
        for c in ctxs:            # each rank fills positions (all) and its own velocity range
            c.download(*out)
        if exchange == "nccl_overlap":
            # the own j-shard is summed first (it needs no remote data): same terms, other order
            for a, b in zip(out, st1):
                assert np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)) < 1e-6
            assert np.max(np.abs(ke - ke1) / ke1) < 1e-6
            assert all(c.info()["kernel_launches"] == 2 * steps for c in ctxs)
        else:
            for a, b in zip(out, st1):
                assert np.array_equal(a, b)
            assert np.max(np.abs(ke - ke1) / ke1) < 1e-12
        # every replica holds the same positions, bit for bit
        ref = ctxs[0].state()
        for c in ctxs[1:]:
            rep = c.state()
            for a, b in zip(rep[:3], ref[:3]):
                assert np.array_equal(a, b)
        assert secs > 0
    finally:
        for c in ctxs:
            c.close()


def test_cli_multi_gpu(pkg, nbx):
    if _ngpu(nbx) < 2:
        pytest.skip("needs 2 GPUs")
    outs = {}
    for g, x in ((1, "nccl"), (2, "nccl"), (2, "p2p"), (2, "nccl_overlap")):
        env = dict(os.environ, NBODY_GPUS=str(g), NBODY_SFREQ="5", NBODY_EXCHANGE=x, NBODY_JSPLITS="2", NBODY_GRAPH="0")
        r = subprocess.run([pkg.CLI_PATH, "4096", "10"], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr
        outs[(g, x)] = [l.split()[2] for l in r.stdout.splitlines() if re.match(r"^ \d+", l)]
        assert f"# Number GPUs        : {g}" in r.stdout
    assert outs[(1, "nccl")] == outs[(2, "nccl")] == outs[(2, "p2p")]
    for a, b in zip(outs[(1, "nccl")], outs[(2, "nccl_overlap")]):
        assert abs(float(a) - float(b)) / float(a) < 1e-4      # 5-digit column, other summation order


@pytest.mark.parametrize("exchange", ["nccl", "p2p", "nccl_overlap"])
def test_torchrun_two_ranks(nbx, exchange, tmp_path):
    """One process per GPU, the way bench.py is launched: ranks step together and agree with
    the single-GPU result."""
    if _ngpu(nbx) < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "rank.py"
    script.write_text(f"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, {REPO!r})
import torch
pkg = importlib.import_module("nbody-demo-2023_b200"); nbx = pkg.nbx
dist = importlib.import_module("nbody-demo-2023_b200.dist")
rank, local_rank, world = dist.init("nccl")
n, steps = 4992, 5      # multiple of 8*world: same padding, hence same j-split boundaries, as 1 GPU
arrs = nbx.ic(n)
ctx = dist.make_sharded_context(nbx, n, {XCH[exchange]})
ctx.set_option("j_splits", 2)
ctx.upload(*arrs)
dist.barrier()
ke, secs = ctx.run(steps)
st = ctx.state()
i0, cnt = ctx.info()["i_begin"], ctx.info()["i_count"]
np.savez({str(tmp_path)!r} + f"/rank{{rank}}.npz", ke=ke, px=st[0], py=st[1], pz=st[2], vx=st[3], i0=i0, cnt=cnt)
dist.barrier()
ctx.close()
torch.distributed.destroy_process_group()
""")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    arrs = nbx.ic(4992)
    ke1, st1 = _single(nbx, arrs, 5, 2)
    for rank in (0, 1):
        d = np.load(tmp_path / f"rank{rank}.npz")
        i0, cnt = int(d["i0"]), int(d["cnt"])
        hi = min(i0 + cnt, 4992)
        if exchange == "nccl_overlap":
            assert np.max(np.abs(d["ke"] - ke1) / ke1) < 1e-6
            for k, f in enumerate(("px", "py", "pz")):
                assert np.allclose(d[f], st1[k], rtol=1e-6, atol=1e-9)
        else:
            assert np.max(np.abs(d["ke"] - ke1) / ke1) < 1e-12
            for k, f in enumerate(("px", "py", "pz")):
                assert np.array_equal(d[f], st1[k])
            assert np.array_equal(d["vx"][i0:hi], st1[3][i0:hi])
