"""Multi-GPU plumbing: one process per GPU (torchrun), MSMs sharded by (base, scalar) chunk with one
NCCL all-gather of partial sums per MSM batch (include/b2m.h b2m_ctx_attach_comm).  torch.distributed is
used only to hand the NCCL unique id from rank 0 to the other ranks."""
import ctypes

from . import _lib

UNIQUE_ID_BYTES = 128


def shard_range(n, rank, world):
    """[lo, hi) of n (base, scalar) pairs owned by `rank` -- the rule csrc/comm.cuh shard_range applies."""
    return n * rank // world, n * (rank + 1) // world


def broadcast_unique_id(dist, rank, make_id, device=None):
    """rank 0 creates the id (bytes), everyone receives it."""
    import torch
    buf = torch.zeros(UNIQUE_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        raw = make_id()
        assert len(raw) == UNIQUE_ID_BYTES
        buf = torch.tensor(list(raw), dtype=torch.uint8)
    if device is not None:
        buf = buf.to(device)
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().tolist())


def _make_id():
    raw = (ctypes.c_uint8 * UNIQUE_ID_BYTES)()
    _lib.check(_lib.lib().b2m_comm_unique_id(raw, UNIQUE_ID_BYTES))
    return bytes(raw)


def attach(ctx, dist, rank, world):
    """Join `ctx` (marlin_b200.api.Context) to the world's NCCL communicator."""
    import torch
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else None
    raw = broadcast_unique_id(dist, rank, _make_id, device)
    arr = (ctypes.c_uint8 * UNIQUE_ID_BYTES)(*raw)
    _lib.check(_lib.lib().b2m_ctx_attach_comm(ctx.handle, arr, UNIQUE_ID_BYTES, rank, world))
