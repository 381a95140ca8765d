"""CPU suite, part 3: the N>1 host path on world_size-2 gloo.  The data path of a multi-GPU
run lives inside libnbx; what the host adds is shard arithmetic and moving a few opaque bytes
between ranks (NCCL unique id, P2P handle blobs) plus reducing timings -- tested here."""
import importlib
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    d = importlib.import_module("nbody-demo-2023_b200.dist")
    nbx = importlib.import_module("nbody-demo-2023_b200").nbx
    r, lr, w = d.init("gloo")
    assert (r, lr, w) == (rank, rank, world)
    uid = bytes(range(128)) if rank == 0 else None
    uid = d.broadcast_bytes(uid, 128, src=0)
    blob = bytes([rank]) * nbx.P2P_BLOB_BYTES
    allb = d.all_gather_bytes(blob)
    tmax = d.reduce_scalar(1.0 + rank, "max")
    tsum = d.reduce_scalar(1.0 + rank, "sum")
    d.barrier()
    # the shards of all ranks tile [0, n_pad) exactly
    n = 1000003
    i0, cnt, n_pad = nbx.shard_of(n, rank, world)
    cover = d.reduce_scalar(float(cnt), "sum")
    # weak-scaling aggregate the way bench.py forms it: units of all ranks / max time
    q.put((rank, uid == bytes(range(128)), allb == b"".join(bytes([g]) * nbx.P2P_BLOB_BYTES for g in range(world)),
           tmax, tsum, cover == n_pad, i0 == rank * cnt))


def test_two_rank_gloo_plumbing():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29533
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, uid_ok, blobs_ok, tmax, tsum, cover_ok, start_ok in res:
        assert uid_ok and blobs_ok and cover_ok and start_ok
        assert tmax == 2.0 and tsum == 3.0


def test_single_process_helpers_are_passthrough():
    sys.path.insert(0, REPO)
    d = importlib.import_module("nbody-demo-2023_b200.dist")
    assert d.env_world() == (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)))
    assert d.broadcast_bytes(b"abc", 3) == b"abc"
    assert d.all_gather_bytes(b"xyz") == b"xyz"
    assert d.reduce_scalar(2.5) == 2.5
