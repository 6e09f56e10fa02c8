"""Multi-GPU plumbing: one process per GPU (torchrun), MSMs sharded by residue class of the SRS index with one
NCCL all-gather of partial sums per MSM batch (include/b2m.h b2m_ctx_attach_comm).  torch.distributed is
used only to hand the NCCL unique id from rank 0 to the other ranks."""
import ctypes

from . import _lib

UNIQUE_ID_BYTES = 128


def shard_slots(base_off, n, rank, world):
    """The pairs of the MSM slice powers[base_off : base_off + n] that GPU `rank` computes -- the rule of
    csrc/msm_impl.cuh run_batch.  GPU r holds the window tables of the powers i = r (mod world) at slot i // world,
    so it takes the pairs with base_off + i = r (mod world).
    Returns (skip, count, first_slot): pairs i = skip + k * world for k < count; pair i uses table slot first_slot + k."""
    skip = (rank - base_off) % world
    count = (n - skip + world - 1) // world if n > skip else 0
    return skip, count, (base_off + skip) // world


def resident_powers(n_srs, rank, world):
    """number of SRS powers whose window tables live on GPU `rank`"""
    return (n_srs - rank + world - 1) // world if n_srs > rank else 0


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
