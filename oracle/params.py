"""Field and curve constants ([U ark-bls12-381 / ark-bn254 0.3 parameters]).

Every derived constant is re-derived from its defining formula in `FieldParams.__init__`
and the curve generators are checked to lie on the curve at import time.
"""


class FieldParams:
    def __init__(self, name, modulus, generator, repr_shave_bits):
        self.name = name
        self.p = modulus
        self.bits = modulus.bit_length()
        self.limbs64 = (self.bits + 63) // 64
        self.nbytes = self.limbs64 * 8
        self.R = (1 << (64 * self.limbs64)) % modulus
        self.Rinv = pow(self.R, -1, modulus)
        self.generator = generator
        s = 0
        t = modulus - 1
        while t % 2 == 0:
            t //= 2
            s += 1
        self.two_adicity = s
        # TWO_ADIC_ROOT_OF_UNITY = GENERATOR^((p-1)/2^s)   [U ark-ff FftParameters]
        self.two_adic_root = pow(generator, t, modulus)
        assert pow(self.two_adic_root, 1 << s, modulus) == 1
        assert pow(self.two_adic_root, 1 << (s - 1), modulus) == modulus - 1
        # [U ark-ff FpParameters::REPR_SHAVE_BITS] = 64*limbs - MODULUS_BITS
        assert repr_shave_bits == 64 * self.limbs64 - self.bits
        self.repr_shave_bits = repr_shave_bits

    def to_mont(self, x):
        return x * self.R % self.p

    def from_mont(self, x):
        return x * self.Rinv % self.p


class CurveParams:
    """y^2 = x^3 + b over Fq, prime-order subgroup of order r = |Fr|."""

    def __init__(self, name, fq, fr, b, gx, gy):
        self.name = name
        self.fq = fq
        self.fr = fr
        self.b = b
        self.g = (gx, gy)
        assert (gy * gy - gx * gx * gx - b) % fq.p == 0, "generator not on curve"


BLS12_381_FR = FieldParams(
    "bls12_381_fr", 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001, 7, 1)
BLS12_381_FQ = FieldParams(
    "bls12_381_fq",
    0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab, 2, 3)
BN254_FR = FieldParams(
    "bn254_fr", 21888242871839275222246405745257275088548364400416034343698204186575808495617, 5, 2)
BN254_FQ = FieldParams(
    "bn254_fq", 21888242871839275222246405745257275088696311157297823662689037894645226208583, 3, 2)

BLS12_381 = CurveParams(
    "bls12_381", BLS12_381_FQ, BLS12_381_FR, 4,
    0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
    0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1)
BN254 = CurveParams("bn254", BN254_FQ, BN254_FR, 3, 1, 2)

CURVES = {"bls12_381": BLS12_381, "bn254": BN254}
