"""torchrun entry (not collected by pytest): N-rank sharded prover must emit the very bytes the oracle /
single-GPU prover emits.  Usage: python -m torch.distributed.run --nproc-per-node N tests/multi_gpu_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from marlin_b200 import api, multi, r1cs as gr1cs
from oracle import kzg, marlin as omarlin, r1cs as or1cs
from oracle import rng as orng
from oracle.params import BLS12_381 as curve


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    f = curve.fr
    n = 256
    rng = orng.test_rng()
    a, b = orng.field_rand(f, rng), orng.field_rand(f, rng)
    ocirc = or1cs.dummy_circuit(f, a, b, 10, n)
    osrs = omarlin.universal_setup(curve, n, n, 3 * n, beta=0x1234567, g_scalar=1, gamma=7)
    eng = kzg.Engine(use_trapdoor=True)
    ok = True
    for scheme, oscheme in (("marlin_kzg10", kzg.MARLIN), ("sonic_kzg10", kzg.SONIC)):
        opk = omarlin.index(osrs, ocirc, oscheme, eng)
        want = omarlin.serialize_proof(curve, oscheme, omarlin.prove(opk, ocirc, orng.test_rng(), eng))
        m = api.Marlin("bls12_381", scheme, device=local)
        multi.attach(m.ctx, dist, rank, world)
        srs = m.srs_from_trapdoor(osrs.max_degree, beta=0x1234567, gamma=7, degree_bounds=(n - 2, 4 * n - 2))
        g = gr1cs.dummy_circuit(0, a, b, 10, n)
        pk = m.index(srs, g)
        got = m.prove(pk, g, api.ZkRng.test_rng())
        good = pk.vk_bytes == opk.vk_bytes and got == want
        print(f"rank {rank}/{world} {scheme}: {'OK' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
        pk.close(); srs.close()
    # level 0: slices of every offset / length parity through the residue-class shard rule (b2m_srs_msm)
    import random
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import b2m_testutil as tu
    m = api.Marlin("bls12_381", "marlin_kzg10", device=local)
    multi.attach(m.ctx, dist, rank, world)
    beta, n_srs = 0xabcdef12345, 1000
    srs = m.srs_from_trapdoor(n_srs - 1, beta=beta, gamma=7)
    rnd = random.Random(5)
    for off, cnt in ((0, n_srs), (1, 1), (3, 2), (7, world), (11, world + 1), (999, 1), (0, 0), (13, 500), (world - 1, 64)):
        sc = [rnd.randrange(f.p) for _ in range(cnt)]
        got = tu.srs_msm(srs.handle, curve, off, sc)
        good = got == tu.trapdoor_msm(curve, curve.g, beta, off, sc)
        if not good:
            print(f"rank {rank}/{world} srs_msm off={off} n={cnt}: MISMATCH", flush=True)
        ok = ok and good
    print(f"rank {rank}/{world} level-0 slices: {'OK' if ok else 'MISMATCH'}", flush=True)
    srs.close()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
