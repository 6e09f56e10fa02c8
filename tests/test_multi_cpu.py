"""CPU (gloo, world_size = 2): host-side logic of the multi-GPU MSM path -- the unique-id hand-off and
the shard rule -- with the oracle standing in for the per-rank partial MSM: folding the gathered
partial sums in rank order must give the full MSM (what csrc/msm_impl.cuh msm_combine_kernel does)."""
import os
import random

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from marlin_b200 import multi
from oracle import ec
from oracle.params import BLS12_381


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = multi.broadcast_unique_id(dist, rank, lambda: bytes((7 * i + 3) % 256 for i in range(multi.UNIQUE_ID_BYTES)))
        curve = BLS12_381
        rnd = random.Random(99)  # identical on every rank: the prover state is replicated
        bases = [ec.scalar_mul(curve, rnd.randrange(1, curve.fr.p), curve.g) for _ in range(n)]
        scalars = [rnd.randrange(curve.fr.p) for _ in range(n)]
        blind_bases = [ec.scalar_mul(curve, 11 + i, curve.g) for i in range(3)]
        blind = [rnd.randrange(curve.fr.p) for _ in range(3)]
        lo, hi = multi.shard_range(n, rank, world)
        part = ec.msm_naive(curve, bases[lo:hi], scalars[lo:hi])
        if rank == 0:  # the blinding terms ride with rank 0's shard
            part = ec.affine_add(curve, part, ec.msm_naive(curve, blind_bases, blind))
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi, part))
        total = None
        for _, _, p in gathered:
            total = ec.affine_add(curve, total, p)
        want = ec.affine_add(curve, ec.msm_naive(curve, bases, scalars), ec.msm_naive(curve, blind_bases, blind))
        q.put((rank, uid, [(g[0], g[1]) for g in gathered], total == want))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1, 7, 32])
def test_sharded_msm_world2(n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + random.randrange(2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    uids = {r[1] for r in res}
    assert len(uids) == 1 and len(next(iter(uids))) == multi.UNIQUE_ID_BYTES
    for _, _, ranges, ok in res:
        assert ok
        assert ranges[0][0] == 0 and ranges[-1][1] == n and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))


def test_shard_range_partitions():
    for n in (0, 1, 5, 1 << 20, (1 << 22) - 1):
        for world in (1, 2, 4, 8):
            parts = [multi.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1
