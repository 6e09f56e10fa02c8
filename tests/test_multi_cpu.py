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
        # the MSM is the slice powers[off : off + n] of a longer key; this rank holds the powers = rank (mod world)
        off = 5
        skip, cnt, slot0 = multi.shard_slots(off, n, rank, world)
        key = [None] * off + bases
        resident = key[rank::world]  # slot k <-> power k * world + rank
        mine = [skip + k * world for k in range(cnt)]
        assert all(resident[slot0 + k] is key[off + i] for k, i in enumerate(mine))
        part = ec.msm_naive(curve, [bases[i] for i in mine], [scalars[i] for i in mine])
        if rank == 0:  # the blinding terms ride with rank 0's shard
            part = ec.affine_add(curve, part, ec.msm_naive(curve, blind_bases, blind))
        gathered = [None] * world
        dist.all_gather_object(gathered, (skip, cnt, part))
        total = None
        for _, _, p in gathered:
            total = ec.affine_add(curve, total, p)
        want = ec.affine_add(curve, ec.msm_naive(curve, bases, scalars), ec.msm_naive(curve, blind_bases, blind))
        q.put((rank, uid, [(g[0], g[1]) for g in gathered], total == want))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,world", [(1, 2), (7, 2), (32, 2), (10, 3)])
def test_sharded_msm_world2(n, world):
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
    for _, _, parts, ok in res:
        assert ok
        assert sum(cnt for _, cnt in parts) == n


def test_shard_slots_partition():
    """every pair of every slice is taken by exactly one rank, evenly, and lands inside that rank's tables"""
    for n_srs in (1, 5, 64, 1000):
        for world in (1, 2, 4, 8):
            for off in (0, 1, 3, n_srs // 2):
                for n in (0, 1, 5, n_srs - off):
                    if n < 0 or off + n > n_srs:
                        continue
                    seen, counts = [], []
                    for r in range(world):
                        skip, cnt, slot0 = multi.shard_slots(off, n, r, world)
                        counts.append(cnt)
                        for k in range(cnt):
                            i = skip + k * world
                            assert (off + i) % world == r and (off + i) // world == slot0 + k
                            assert slot0 + k < multi.resident_powers(n_srs, r, world)
                            seen.append(i)
                    assert sorted(seen) == list(range(n))
                    assert max(counts) - min(counts) <= 1
            assert sum(multi.resident_powers(n_srs, r, world) for r in range(world)) == n_srs
