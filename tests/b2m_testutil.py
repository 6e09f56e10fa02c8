"""Marshalling helpers shared by the tests: python ints <-> the C ABI's limb arrays."""
import ctypes

import numpy as np

from marlin_b200 import _lib
from oracle import ec
from oracle.params import CURVES

CURVE_ID = {"bls12_381": _lib.CURVE_BLS12_381, "bn254": _lib.CURVE_BN254}


def fr_to_mont_limbs(curve, vals):
    f = curve.fr
    return _lib.ints_to_limbs([f.to_mont(v % f.p) for v in vals], f.limbs64)


def fr_to_canon_limbs(curve, vals):
    f = curve.fr
    return _lib.ints_to_limbs([v % f.p for v in vals], f.limbs64)


def fr_from_mont_limbs(curve, arr):
    f = curve.fr
    return [f.from_mont(v) for v in _lib.limbs_to_ints(arr)]


def points_to_limbs(curve, pts):
    """affine points (or None) -> (n, 2*LQ) uint64 Montgomery limbs; None -> all zero."""
    fq = curve.fq
    flat = []
    for P in pts:
        if P is None:
            flat += [0, 0]
        else:
            flat += [fq.to_mont(P[0]), fq.to_mont(P[1])]
    return _lib.ints_to_limbs(flat, fq.limbs64).reshape(len(pts), 2 * fq.limbs64)


def points_from_limbs(curve, arr):
    fq = curve.fq
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 2 * fq.limbs64)
    out = []
    for row in arr:
        x, y = _lib.limbs_to_ints(row.reshape(2, fq.limbs64))
        out.append(None if x == 0 and y == 0 else (fq.from_mont(x), fq.from_mont(y)))
    return out


def gpu_powers(ctx, curve, g, beta, n):
    """powers_of_g[i] = beta^i * g computed by the library; returns the limb array."""
    L = _lib.lib()
    fq = curve.fq
    out = np.zeros((n, 2 * fq.limbs64), dtype=np.uint64)
    gl = points_to_limbs(curve, [g])
    bl = fr_to_canon_limbs(curve, [beta])
    _lib.check(L.b2m_g1_powers(ctx, CURVE_ID[curve.name], _lib.ptr(gl), _lib.ptr(bl), n, _lib.ptr(out)))
    return out


def make_srs(ctx, curve, powers_limbs, gamma_limbs=None, gamma_indices=None, window_bits=0):
    L = _lib.lib()
    h = ctypes.c_void_p()
    ng = 0 if gamma_limbs is None else len(gamma_limbs)
    gi = None if gamma_indices is None else np.asarray(gamma_indices, dtype=np.uint64)
    _lib.check(L.b2m_srs_create(ctx, CURVE_ID[curve.name], _lib.ptr(powers_limbs), len(powers_limbs),
                                _lib.ptr(gamma_limbs), _lib.ptr(gi), ng, window_bits, ctypes.byref(h)))
    return h


def srs_msm(srs, curve, base_off, scalars):
    """MSM over powers[base_off:base_off+n] with canonical python-int scalars -> affine point or None."""
    L = _lib.lib()
    sc = fr_to_canon_limbs(curve, scalars) if len(scalars) else np.zeros((1, curve.fr.limbs64), dtype=np.uint64)
    out = np.zeros(2 * curve.fq.limbs64, dtype=np.uint64)
    inf = ctypes.c_int(0)
    _lib.check(L.b2m_srs_msm(srs, base_off, _lib.ptr(sc), len(scalars), _lib.ptr(out), ctypes.byref(inf)))
    P = points_from_limbs(curve, out)[0]
    assert (P is None) == bool(inf.value)
    return P


def trapdoor_msm(curve, g, beta, base_off, scalars):
    """(sum_i s_i beta^(off+i)) * g : what an MSM over the powers beta^i*g must equal."""
    r = curve.fr.p
    acc = 0
    cur = pow(beta, base_off, r)
    for s in scalars:
        acc = (acc + s * cur) % r
        cur = cur * beta % r
    return ec.scalar_mul(curve, acc, g)
