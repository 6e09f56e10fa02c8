"""CPU: the host-side parts of SRS generation and SRS files: the G2 scalar multiplications of
libb2m (b2m_g2_scalar_muls, host C++ over the device's limb code) against a definitional Fq2 implementation in Python integers,
the standard G2 generators (on the twist, of order r), and the ark-serialize file layout round trip."""
import os
import random

import numpy as np
import pytest

from marlin_b200 import _lib, fields, srsfile
from oracle.params import BLS12_381, BN254


class Fq2Ref:
    def __init__(self, p):
        self.p = p

    def add(self, a, b):
        return ((a[0] + b[0]) % self.p, (a[1] + b[1]) % self.p)

    def sub(self, a, b):
        return ((a[0] - b[0]) % self.p, (a[1] - b[1]) % self.p)

    def mul(self, a, b):
        return ((a[0] * b[0] - a[1] * b[1]) % self.p, (a[0] * b[1] + a[1] * b[0]) % self.p)

    def inv(self, a):
        n = pow((a[0] * a[0] + a[1] * a[1]) % self.p, -1, self.p)
        return (a[0] * n % self.p, (-a[1]) * n % self.p)

    def padd(self, P, Q):
        if P is None:
            return Q
        if Q is None:
            return P
        if P[0] == Q[0]:
            if self.add(P[1], Q[1]) == (0, 0):
                return None
            lam = self.mul(self.mul((3, 0), self.mul(P[0], P[0])), self.inv(self.mul((2, 0), P[1])))
        else:
            lam = self.mul(self.sub(Q[1], P[1]), self.inv(self.sub(Q[0], P[0])))
        x3 = self.sub(self.sub(self.mul(lam, lam), P[0]), Q[0])
        return (x3, self.sub(self.mul(lam, self.sub(P[0], x3)), P[1]))

    def smul(self, k, P):
        R = None
        while k:
            if k & 1:
                R = self.padd(R, P)
            P = self.padd(P, P)
            k >>= 1
        return R


def parse_g2(raw, nb):
    vals = [int.from_bytes(raw[i * nb:(i + 1) * nb], "little") for i in range(4)]
    if raw[-1] & 0x40:
        return None
    return ((vals[0], vals[1]), (vals[2], vals[3]))


@pytest.mark.parametrize("cid,curve", [(0, BLS12_381), (1, BN254)], ids=["bls12_381", "bn254"])
def test_g2_scalar_muls_and_generator(cid, curve):
    L = _lib.lib()
    f2 = Fq2Ref(curve.fq.p)
    nb = curve.fq.nbytes
    r = curve.fr.p
    rnd = random.Random(3 + cid)
    ks = [1, 2, r - 1, 0, rnd.randrange(r), rnd.randrange(r)]
    out = np.zeros(len(ks) * 4 * nb, dtype=np.uint8)
    _lib.check(L.b2m_g2_scalar_muls(cid, None, _lib.ptr(_lib.ints_to_limbs(ks, 4)), len(ks), _lib.ptr(out)))
    raw = out.tobytes()
    pts = [parse_g2(raw[i * 4 * nb:(i + 1) * 4 * nb], nb) for i in range(len(ks))]
    g = pts[0]
    b2 = (4, 4) if cid == 0 else f2.mul((3, 0), f2.inv((9, 1)))  # twist coefficient: 4 (1 + u) / 3 / (9 + u)
    assert f2.mul(g[1], g[1]) == f2.add(f2.mul(f2.mul(g[0], g[0]), g[0]), b2), "generator is not on the twist"
    assert f2.smul(r, g) is None, "generator is not of order r"
    for k, P in zip(ks, pts):
        assert P == f2.smul(k, g), k
    # a caller-supplied base: k2 * (k1 * g) = (k1 k2) * g
    base = raw[4 * 4 * nb:5 * 4 * nb]
    out2 = np.zeros(4 * nb, dtype=np.uint8)
    _lib.check(L.b2m_g2_scalar_muls(cid, _lib.ptr(np.frombuffer(base, dtype=np.uint8).copy()), _lib.ptr(_lib.ints_to_limbs([ks[5]], 4)), 1, _lib.ptr(out2)))
    assert parse_g2(out2.tobytes(), nb) == f2.smul(ks[4] * ks[5] % r, g)


def test_srs_file_round_trip(tmp_path):
    cid = 0
    nb = srsfile.fq_bytes(cid)
    rnd = random.Random(1)
    blob = lambda n: bytes(rnd.randrange(256) for _ in range(n))
    powers = blob(5 * 2 * nb)
    gamma = {0: blob(2 * nb), 1: blob(2 * nb), 2: blob(2 * nb), 17: blob(2 * nb)}
    h, bh = blob(4 * nb), blob(4 * nb)
    neg = {3: blob(4 * nb), 9: blob(4 * nb)}
    path = os.path.join(tmp_path, "srs.bin")
    srsfile.write_srs(path, cid, powers, gamma, h, bh, neg)
    d = srsfile.read_srs(path)
    assert (d["curve_id"], d["powers"], d["gamma"], d["h"], d["beta_h"], d["neg_powers"]) == (cid, powers, gamma, h, bh, neg)
    assert os.path.getsize(path) == 16 + 8 + len(powers) + 8 + 4 * (8 + 2 * nb) + 2 * 4 * nb + 8 + 2 * (8 + 4 * nb)
    with open(path, "ab") as f:
        f.write(b"x")
    with pytest.raises(ValueError):
        srsfile.read_srs(path)
    h2, bh2, neg2 = srsfile.g2_setup(cid, fields.FR_MODULUS[cid], 12345, 63, [10, 40])
    assert len(h2) == len(bh2) == 4 * nb and sorted(neg2) == [23, 53]
