"""CPU: the device field / curve headers (csrc/field.cuh, curve.cuh) compiled for the host with an
emulated carry flag, checked limb for limb against Python integers and the oracle's group law."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle import ec
from oracle.params import BLS12_381, BN254, BLS12_381_FR, BLS12_381_FQ, BN254_FR, BN254_FQ

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "host", "libfield_host.so")


@pytest.fixture(scope="module")
def hostlib():
    src = os.path.join(HERE, "host", "field_host_shim.cpp")
    deps = [src] + [os.path.join(HERE, "..", "marlin_b200", "csrc", h) for h in ("field.cuh", "curve.cuh", "msm_affine.cuh")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O1", "-DB2M_HOST_LIGHT_INLINE", "-shared", "-fPIC", "-x", "c++", src, "-o", SO])
    return ctypes.CDLL(SO)


@pytest.mark.parametrize("fi,field", list(enumerate([BLS12_381_FR, BLS12_381_FQ, BN254_FR, BN254_FQ])), ids=lambda x: getattr(x, "name", x))
def test_montgomery_ops(hostlib, fi, field):
    p = field.p
    n = 8 if p.bit_length() <= 256 else 12
    R = 1 << (32 * n)
    Rinv = pow(R, -1, p)
    A = ctypes.c_uint32 * n
    rnd = random.Random(fi)

    def call(w, a, b):
        r = A()
        hostlib.field_op(fi, w, A(*[(a >> (32 * i)) & 0xffffffff for i in range(n)]), A(*[(b >> (32 * i)) & 0xffffffff for i in range(n)]), r)
        return sum(int(r[i]) << (32 * i) for i in range(n))

    edge = [(0, 0), (1, p - 1), (p - 1, p - 1), (p - 2, 1), (R % p, R % p), ((p - 1) // 2, p - 1), (2, (p + 1) // 2), (p - 1, 0)]
    for it in range(1500):
        a, b = edge[it] if it < len(edge) else (rnd.randrange(p), rnd.randrange(p))
        assert call(0, a, b) == a * b * Rinv % p
        assert call(1, a, b) == (a + b) % p
        assert call(2, a, b) == (a - b) % p
        assert call(3, a, b) == (-a) % p
        assert call(5, a, b) == a * Rinv % p
        assert call(6, a, b) == a * R % p
    for _ in range(3):
        a = rnd.randrange(1, p)
        assert call(4, a, 0) * a % p == R * R % p
    # binary-Euclid inverse (csrc/field.cuh inverse_fast): a^-1 in Montgomery form is a^-1 * R^2 as an integer
    special = [1, 2, 3, 4, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, R % p, R * R % p, 1 << 32, 1 << 64, (1 << 64) + (1 << 33),
               1 << (p.bit_length() - 1), (1 << (p.bit_length() - 1)) - 1, p - (1 << 40), 0xffffffff, 0xffffffff00000000]
    special += [(1 << k) % p for k in range(1, 32 * n, 17)] + [p - ((1 << k) % p) for k in range(1, 32 * n, 29)]
    for it in range(3000):
        a = special[it] if it < len(special) else rnd.randrange(1, p)
        assert call(7, a, 0) == pow(a, -1, p) * R * R % p, hex(a)
    assert call(7, 0, 0) == 0


@pytest.mark.parametrize("ci,curve", list(enumerate([BLS12_381, BN254])), ids=lambda x: getattr(x, "name", x))
def test_xyzz_group_law(hostlib, ci, curve):
    fq = curve.fq
    n32 = 12 if ci == 0 else 8
    rnd = random.Random(5 + ci)

    def pack(P):
        if P is None:
            return [0] * (2 * n32)
        out = []
        for v in P:
            m = fq.to_mont(v)
            out += [(m >> (32 * i)) & 0xffffffff for i in range(n32)]
        return out

    def unpack(arr):
        x = sum(int(arr[i]) << (32 * i) for i in range(n32))
        y = sum(int(arr[n32 + i]) << (32 * i) for i in range(n32))
        return None if x == 0 and y == 0 else (fq.from_mont(x), fq.from_mont(y))

    pts = [ec.scalar_mul(curve, rnd.randrange(1, curve.fr.p), curve.g) for _ in range(10)]
    pts += [pts[0], pts[1], ec.affine_neg(curve, pts[2]), None, pts[0]]  # doubling, cancellation, infinity
    out = (ctypes.c_uint32 * (2 * n32))()
    for _ in range(10):
        rnd.shuffle(pts)
        neg = [rnd.randrange(2) for _ in pts]
        flat = sum((pack(P) for P in pts), [])
        A = (ctypes.c_uint32 * len(flat))(*flat)
        NG = (ctypes.c_uint8 * len(pts))(*neg)
        exp = None
        for P, s in zip(pts, neg):
            exp = ec.affine_add(curve, exp, ec.affine_neg(curve, P) if s else P)
        for which in (0, 1):
            hostlib.curve_op(ci, which, A, NG, len(pts), None, 0, out)
            assert unpack(out) == exp
    P = next(p for p in pts if p)
    k = rnd.randrange(curve.fr.p)
    K = (ctypes.c_uint32 * 8)(*[(k >> (32 * i)) & 0xffffffff for i in range(8)])
    flat = pack(P)
    hostlib.curve_op(ci, 2, (ctypes.c_uint32 * len(flat))(*flat), (ctypes.c_uint8 * 1)(0), 1, K, 8, out)
    assert unpack(out) == ec.scalar_mul(curve, k, P)


@pytest.mark.parametrize("ci,curve", list(enumerate([BLS12_381, BN254])), ids=lambda x: getattr(x, "name", x))
def test_batched_affine_levels(hostlib, ci, curve):
    """csrc/msm_affine.cuh on the host: L levels of pairwise affine additions with one shared inversion per
    thread, then the XYZZ tail, must give every bucket's sum -- including doubled references (P + P),
    cancelling ones (P - P, whose infinity then meets other points) and odd / empty / single-point buckets."""
    import numpy as np
    fq = curve.fq
    n32 = 12 if ci == 0 else 8
    rnd = random.Random(11 + ci)
    n_tab, B = 24, 13
    tab = [ec.scalar_mul(curve, rnd.randrange(1, curve.fr.p), curve.g) for _ in range(n_tab)]

    def limbs(v):
        m = fq.to_mont(v)
        return [(m >> (32 * i)) & 0xffffffff for i in range(n32)]

    tab_l = np.array(sum((limbs(P[0]) + limbs(P[1]) for P in tab), []), dtype=np.uint32)
    for trial in range(4):
        buckets = [[] for _ in range(B)]
        for b in range(B):
            m = [0, 1, 2, 3, 5, 8, 17, 40][rnd.randrange(8)] if trial else [0, 1, 2, 2, 3, 4, 4, 6, 7, 9, 16, 31, 33][b]
            for _ in range(m):
                buckets[b].append((rnd.randrange(n_tab), rnd.randrange(2)))
            if m >= 4:  # adjacent equal and opposite references: doubling and cancellation at level 0 ...
                buckets[b][0] = buckets[b][1]
                buckets[b][2] = (buckets[b][3][0], 1 - buckets[b][3][1])
            if m >= 8:  # ... and at level 1: (A + B) + (A + B), (A + B) - (A + B)
                buckets[b][4:8] = [buckets[b][4], buckets[b][5], buckets[b][4], buckets[b][5]]
            if m >= 17:
                buckets[b][8:12] = [(1, 0), (2, 1), (1, 1), (2, 0)]
        refs, off = [], [0]
        for b, lst in enumerate(buckets):
            for idx, neg in lst:
                refs += [idx | (neg << 31), b]
            off.append(off[-1] + len(lst))
        want = []
        for lst in buckets:
            acc = None
            for idx, neg in lst:
                acc = ec.affine_add(curve, acc, ec.affine_neg(curve, tab[idx]) if neg else tab[idx])
            want.append(acc)
        refs_a = np.array(refs if refs else [0, 0], dtype=np.uint32)
        off_a = np.array(off, dtype=np.uint32)
        # (levels, T, variant, mapping): variant 0 = the fused kernel's thread function (odd T: with operand prefetch), 3 = the
        # software-pipelined one, 4 / 5 = its split (two-kernel) form with the pipelined / plain addition pass; mapping 0 =
        # blocked, 1 = warp-interleaved
        cases = [(0, 1, 0, 0), (1, 1, 0, 0), (1, 4, 0, 0), (2, 3, 0, 0), (3, 8, 0, 0), (7, 5, 0, 0), (3, 64, 0, 0),
                 (1, 1, 0, 1), (1, 4, 0, 1), (2, 3, 0, 1), (3, 2, 0, 1), (4, 1, 0, 1), (3, 8, 0, 1), (7, 5, 0, 1), (3, 64, 0, 1),
                 (1, 1, 3, 1), (1, 2, 3, 1), (2, 3, 3, 1), (3, 8, 3, 1), (4, 5, 3, 0), (7, 4, 3, 1), (3, 64, 3, 1),
                 (1, 1, 4, 1), (2, 3, 4, 1), (3, 8, 4, 0), (3, 64, 4, 1), (1, 1, 5, 1), (2, 3, 5, 1), (3, 8, 5, 0), (3, 64, 5, 1),
                 (1, 1, 6, 1), (2, 3, 6, 1), (3, 8, 6, 0), (3, 16, 6, 1)]  # 6: split with the chain products inverted outside the thread function
        for levels, T, variant, interleaved in cases:
            out = np.zeros(B * 2 * n32, dtype=np.uint32)
            hostlib.affine_levels_host(ci, tab_l.ctypes.data_as(ctypes.c_void_p), refs_a.ctypes.data_as(ctypes.c_void_p),
                                       off_a.ctypes.data_as(ctypes.c_void_p), B, levels, T, out.ctypes.data_as(ctypes.c_void_p), variant,
                                       interleaved)
            for b in range(B):
                x = sum(int(out[b * 2 * n32 + i]) << (32 * i) for i in range(n32))
                y = sum(int(out[b * 2 * n32 + n32 + i]) << (32 * i) for i in range(n32))
                got = None if x == 0 and y == 0 else (fq.from_mont(x), fq.from_mont(y))
                assert got == want[b], (trial, levels, T, variant, interleaved, b)
