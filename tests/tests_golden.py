"""Definition of the golden cases shared by the generator, the oracle tests and the GPU tests."""
import hashlib

from oracle import kzg, marlin, r1cs
from oracle import rng as R
from oracle.params import CURVES

CASES = [
    {"name": "bls_marlin_test_circuit_26x25", "curve": "bls12_381", "scheme": kzg.MARLIN, "circuit": "test", "nc": 26, "nv": 25},
    {"name": "bls_sonic_test_circuit_25x100", "curve": "bls12_381", "scheme": kzg.SONIC, "circuit": "test", "nc": 25, "nv": 100},
    {"name": "bls_marlin_dummy_2p6", "curve": "bls12_381", "scheme": kzg.MARLIN, "circuit": "dummy", "nc": 64, "nv": 10},
    {"name": "bls_sonic_dummy_2p6", "curve": "bls12_381", "scheme": kzg.SONIC, "circuit": "dummy", "nc": 64, "nv": 10},
    {"name": "bn254_marlin_dummy_2p5", "curve": "bn254", "scheme": kzg.MARLIN, "circuit": "dummy", "nc": 32, "nv": 10},
    {"name": "bls_marlin_dummy_2p10_config1", "curve": "bls12_381", "scheme": kzg.MARLIN, "circuit": "dummy", "nc": 1024, "nv": 10},
]
BETA, G_SCALAR, GAMMA = 0x0123456789abcdef0123456789abcdef, 3, 11
ZK_SEED = bytes(range(32))


def case_inputs(case):
    curve = CURVES[case["curve"]]
    f = curve.fr
    rng = R.test_rng()
    a, b = R.field_rand(f, rng), R.field_rand(f, rng)
    if case["circuit"] == "test":
        circ = r1cs.test_circuit(f, a, b, case["nc"], case["nv"])
        pub = [a * b % f.p, a * b % f.p * b % f.p]
    else:
        circ = r1cs.dummy_circuit(f, a, b, case["nv"], case["nc"])
        pub = [a * b % f.p]
    return curve, a, b, circ, pub


def regenerate_case(case):
    curve, a, b, circ, pub = case_inputs(case)
    cs = r1cs.synthesize(curve.fr, circ)
    am, bm, cm = cs.to_matrices()
    nnz = sum(len({i for _, i in ra} | {i for _, i in rb} | {i for _, i in rc}) for ra, rb, rc in zip(am, bm, cm))
    srs = marlin.universal_setup(curve, cs.num_constraints, len(cs.instance) + len(cs.witness), nnz, beta=BETA, g_scalar=G_SCALAR, gamma=GAMMA)
    eng = kzg.Engine(use_trapdoor=True)
    pk = marlin.index(srs, circ, case["scheme"], eng)
    zk = R.ChaChaRng(ZK_SEED, 12)
    proof = marlin.prove(pk, circ, zk, eng)
    assert marlin.verify(pk, pub, proof)
    out = dict(case)
    out.update({"vk_sha256": hashlib.sha256(pk.vk_bytes).hexdigest(), "proof_hex": marlin.serialize_proof(curve, case["scheme"], proof).hex(),
                "zk_rng_word_pos_after": zk.word_pos, "srs_max_degree": srs.max_degree})
    return out
