"""Checker script (test infrastructure, not collected by pytest): verify a proof saved by `bench.py --save-proof F`
with the oracle's restatement of `Marlin::verify` -- once through the SRS trapdoor and once through the reference's
real pairing check -- from public data only.  Lets a proof made at a size or GPU count the oracle cannot reproduce
(2^24 constraints on 8 GPUs) be checked on any CPU afterwards.

    python tests/verify_saved_proof.py gpurun_out/proof.json [--no-pairing]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import kzg, marlin as omarlin  # noqa: E402
from oracle.params import CURVES  # noqa: E402

SCHEMES = {"marlin_kzg10": kzg.MARLIN, "sonic_kzg10": kzg.SONIC}


def parse_vk(curve, scheme, vk):
    """`IndexVerifierKey::write` image (oracle/marlin.py index_vk_bytes): 3 x u64, then six commitments in ToBytes form."""
    nb = curve.fq.nbytes
    info = [int.from_bytes(vk[8 * i:8 * i + 8], "little") for i in range(3)]
    pos, comms = 24, []

    def point():
        nonlocal pos
        x = int.from_bytes(vk[pos:pos + nb], "little")
        y = int.from_bytes(vk[pos + nb:pos + 2 * nb], "little")
        inf = vk[pos + 2 * nb]
        pos += 2 * nb + 1
        return None if inf else (x, y)

    for _ in range(6):
        comms.append(point())
        if scheme == kzg.MARLIN:
            has_shifted = vk[pos]
            pos += 1
            shifted = point()
            assert not has_shifted and shifted is None, "index commitments carry no degree bound"
    assert pos == len(vk), "trailing bytes in the verifier key"
    return info, comms


def verify_file(path, use_pairing=True):
    return verify_blob(json.load(open(path)), use_pairing)


def verify_blob(d, use_pairing=True):
    """d: {"curve", "pc", "max_degree", "beta", "gamma", "public_input": [str], "proof_hex", "vk_hex"} as written by bench.py"""
    curve, scheme = CURVES[d["curve"]], SCHEMES[d["pc"]]
    vk_bytes, proof_bytes = bytes.fromhex(d["vk_hex"]), bytes.fromhex(d["proof_hex"])
    (num_variables, num_constraints, num_non_zero), comms = parse_vk(curve, scheme, vk_bytes)
    lazy = kzg.UniversalParams(curve, d["max_degree"], d["beta"], curve.g, d["gamma"], powers_of_g="lazy")
    vk = omarlin.verifier_key_from_public(curve, scheme, lazy, num_constraints, num_variables, num_non_zero, comms)
    assert vk.vk_bytes == vk_bytes, "verifier key does not re-serialise to the saved bytes"
    proof = omarlin.deserialize_proof(curve, scheme, proof_bytes)
    pub = [int(v) for v in d["public_input"]]
    wrong = [(pub[0] + 1) % curve.fr.p] + pub[1:]
    res = {"trapdoor": omarlin.verify(vk, pub, proof), "trapdoor_rejects_wrong_input": not omarlin.verify(vk, wrong, proof)}
    if use_pairing:
        g2 = kzg.G2Key(lazy, vk.ck.enforced_degree_bounds)
        res["pairing"] = omarlin.verify(vk, pub, proof, g2)
        res["pairing_rejects_wrong_input"] = not omarlin.verify(vk, wrong, proof, g2)
    res["ok"] = all(res.values())
    res.update({"num_constraints": num_constraints, "num_variables": num_variables, "num_non_zero": num_non_zero,
                "proof_bytes": len(proof_bytes), "n_gpus": d.get("n_gpus")})
    return res


if __name__ == "__main__":
    out = verify_file(sys.argv[1], "--no-pairing" not in sys.argv)
    print(json.dumps(out))
    sys.exit(0 if out["ok"] else 1)
