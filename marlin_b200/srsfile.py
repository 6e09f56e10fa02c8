"""SRS files in ark-serialize layout: `kzg10::UniversalParams<E>` written field by field with
`CanonicalSerialize::serialize_uncompressed` [U ark-poly-commit 0.3 kzg10/data_structures.rs, ark-serialize 0.3]:

    powers_of_g        : Vec<G1Affine>             u64-LE length, then the points
    powers_of_gamma_g  : BTreeMap<usize, G1Affine> u64-LE length, then (u64-LE key, point) in ascending key order
    h, beta_h          : G2Affine
    neg_powers_of_h    : BTreeMap<usize, G2Affine> as above (SonicKZG10: the entries its `trim` reads, max_degree - bound)

A point is its canonical little-endian coordinates (x || y; over Fq2: c0 || c1 each) with the infinity flag in bit 6 of the
last byte.  tools/replay_rs reads this file with `deserialize_unchecked` into the public fields of `UniversalParams`, so an SRS
made on the GPU can be replayed through the real `ark_marlin::Marlin::{index, prove, verify}`.  All group arithmetic and the
Montgomery <-> canonical conversions happen in libb2m (GPU for G1, host C++ for the few G2 points); this module only moves bytes.
"""
import struct

import numpy as np

from . import _lib

MAGIC = b"B2MSRS01"  # 8-byte tag + curve id (u64) ahead of the ark-serialize payload, so that a wrong-curve load fails loudly


def fq_bytes(curve_id):
    return 8 * _lib.LIMBS[curve_id][1]


def write_srs(path, curve_id, powers, gamma, h, beta_h, neg_powers):
    """powers: bytes (n * 2 * fq_bytes); gamma: {index: bytes}; h, beta_h: bytes (4 * fq_bytes); neg_powers: {index: bytes}"""
    g1, g2 = 2 * fq_bytes(curve_id), 4 * fq_bytes(curve_id)
    assert len(powers) % g1 == 0 and len(h) == g2 and len(beta_h) == g2
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<Q", curve_id))
        f.write(struct.pack("<Q", len(powers) // g1))
        f.write(powers)
        f.write(struct.pack("<Q", len(gamma)))
        for k in sorted(gamma):
            assert len(gamma[k]) == g1
            f.write(struct.pack("<Q", k) + gamma[k])
        f.write(h + beta_h)
        f.write(struct.pack("<Q", len(neg_powers)))
        for k in sorted(neg_powers):
            assert len(neg_powers[k]) == g2
            f.write(struct.pack("<Q", k) + neg_powers[k])


def read_srs(path):
    """-> dict(curve_id, powers (bytes), gamma {index: bytes}, h, beta_h, neg_powers {index: bytes})"""
    with open(path, "rb") as f:
        head = f.read(16)
        if head[:8] != MAGIC:
            raise ValueError(f"{path}: not a b2m SRS file")
        curve_id = struct.unpack("<Q", head[8:])[0]
        if curve_id not in _lib.LIMBS:
            raise ValueError(f"{path}: unknown curve id {curve_id}")
        g1, g2 = 2 * fq_bytes(curve_id), 4 * fq_bytes(curve_id)

        def u64():
            b = f.read(8)
            if len(b) != 8:
                raise ValueError(f"{path}: truncated")
            return struct.unpack("<Q", b)[0]

        def blob(n):
            b = f.read(n)
            if len(b) != n:
                raise ValueError(f"{path}: truncated")
            return b

        n = u64()
        powers = blob(n * g1)
        gamma = {}
        for _ in range(u64()):
            k = u64()
            gamma[k] = blob(g1)
        h, beta_h = blob(g2), blob(g2)
        neg = {}
        for _ in range(u64()):
            k = u64()
            neg[k] = blob(g2)
        if f.read(1):
            raise ValueError(f"{path}: trailing bytes")
    return {"curve_id": curve_id, "powers": powers, "gamma": gamma, "h": h, "beta_h": beta_h, "neg_powers": neg}


def g2_setup(curve_id, r, beta, max_degree, degree_bounds):
    """h (the standard G2 generator), beta * h and beta^-(max_degree - d) * h per enforced bound d, as uncompressed bytes."""
    L = _lib.lib()
    exps = [1, beta % r] + [pow(pow(beta % r, max_degree - d, r), -1, r) for d in sorted(set(degree_bounds))]
    sc = _lib.ints_to_limbs(exps, 4)
    g2 = 4 * fq_bytes(curve_id)
    out = np.zeros(len(exps) * g2, dtype=np.uint8)
    _lib.check(L.b2m_g2_scalar_muls(curve_id, None, _lib.ptr(sc), len(exps), _lib.ptr(out)))
    raw = out.tobytes()
    pts = [raw[i * g2:(i + 1) * g2] for i in range(len(exps))]
    return pts[0], pts[1], {max_degree - d: pts[2 + i] for i, d in enumerate(sorted(set(degree_bounds)))}
