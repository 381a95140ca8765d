"""Host-side plumbing for one-process-per-GPU runs (torchrun): torch.distributed is used
only to move a few bytes between ranks (NCCL unique id, P2P handle blobs) and to reduce
timings; the data path is inside libnbx (ncclAllGather or NVLink stores from the kernel).
Works on the gloo backend too, which is how the CPU tests exercise it."""
from __future__ import annotations

import os

import numpy as np


def env_world():
    """(rank, local_rank, world) from the torchrun environment; (0, 0, 1) when absent."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init(backend: str | None = None):
    """init_process_group from MASTER_ADDR/MASTER_PORT/RANK/WORLD_SIZE; returns (rank, local_rank, world)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def _dev():
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """Rank `src` supplies `payload` (nbytes long); every rank gets it."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return payload
    t = torch.zeros(nbytes, dtype=torch.uint8, device=_dev())
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def all_gather_bytes(blob: bytes) -> bytes:
    """Concatenation over ranks (rank order) of equal-length blobs."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return blob
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(_dev())
    out = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return b"".join(bytes(t.cpu().numpy().tobytes()) for t in out)


def reduce_scalar(x: float, op: str = "max") -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=_dev())
    dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN}[op])
    return float(t.item())


def barrier():
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def make_sharded_context(nbx, n: int, exchange: int | None = None, device: int | None = None, multicast: int = -1, **ctx_kw):
    """Create this rank's nbx.Context (i-shard rank/world) and wire the exchange (default: the
    library's default, P2P): NCCL unique id broadcast from rank 0, and for P2P the all-gather of handle blobs.
    If ANY rank cannot map its peers' buffers (no peer access / IPC in this container), every
    rank switches to the NCCL all-gather together -- both are GPU paths; `ctx.exchange_used`
    says which one runs."""
    rank, local_rank, world = env_world()
    if exchange is None:
        exchange = nbx.EXCHANGE_P2P
    ctx = nbx.Context(n, device=local_rank if device is None else device, rank=rank, world=world, **ctx_kw)
    ctx.exchange_used = exchange if world > 1 else None
    if world > 1:
        ctx.set_option("exchange", exchange)
        ctx.set_option("multicast", multicast)
        uid = nbx.comm_unique_id() if rank == 0 else None
        uid = broadcast_bytes(uid, nbx.UNIQUE_ID_BYTES, src=0)
        ctx.comm_init(uid)
        if exchange == nbx.EXCHANGE_P2P:
            ok = 1.0
            try:
                blob = ctx.p2p_export()
            except nbx.NbxError as e:
                ok, blob, why = 0.0, bytes(nbx.P2P_BLOB_BYTES), str(e)
            blobs = all_gather_bytes(blob)
            if reduce_scalar(ok, "min") > 0:
                try:
                    ctx.p2p_attach(blobs)
                except nbx.NbxError as e:
                    ok, why = 0.0, str(e)
            if reduce_scalar(ok, "min") == 0:
                if rank == 0:
                    import sys
                    print(f"nbx: P2P exchange unavailable on at least one rank ({why if ok == 0 else 'peer'}); "
                          f"all ranks use the NCCL all-gather", file=sys.stderr)
                ctx.set_option("exchange", nbx.EXCHANGE_NCCL)
                ctx.exchange_used = nbx.EXCHANGE_NCCL
    ctx.multicast = bool(world > 1 and ctx.info()["multicast"])      # P2P stores go through an NVSwitch multicast team
    return ctx
