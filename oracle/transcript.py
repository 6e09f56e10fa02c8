"""`SimpleHashFiatShamirRng<Blake2s, ChaChaRng>` [R src/rng.rs:18-80] and the `ToBytes` encodings
that feed it ([U ark-ff bytes.rs, ark-ec short_weierstrass_jacobian.rs ToBytes])."""
import hashlib
import struct

from .rng import ChaChaRng


class FiatShamirRng:
    def __init__(self, initial_input: bytes):
        # seed = H(initial_input); r = ChaChaRng::from_seed(seed)      [R rng.rs:54-66]
        self.seed = hashlib.blake2s(initial_input, digest_size=32).digest()
        self.r = ChaChaRng(self.seed, rounds=20)

    def absorb(self, new_input: bytes):
        # seed = H(new_input || seed)                                   [R rng.rs:70-79]
        self.seed = hashlib.blake2s(new_input + self.seed, digest_size=32).digest()
        self.r = ChaChaRng(self.seed, rounds=20)

    def next_u32(self):
        return self.r.next_u32()

    def next_u64(self):
        return self.r.next_u64()


def fe_bytes(field, v):
    """ToBytes of a field element: canonical value, little-endian u64 limbs."""
    return (v % field.p).to_bytes(field.nbytes, "little")


def u64_bytes(v):
    return struct.pack("<Q", v)


def g1_affine_bytes(curve, P):
    """ToBytes of `GroupAffine`: x || y || infinity; the identity is (0, 1, true)."""
    fq = curve.fq
    if P is None:
        return fe_bytes(fq, 0) + fe_bytes(fq, 1) + b"\x01"
    return fe_bytes(fq, P[0]) + fe_bytes(fq, P[1]) + b"\x00"


def g1_compressed(curve, P):
    """CanonicalSerialize of a short-Weierstrass affine point [U ark-ec / ark-serialize 0.3 SWFlags]:
    x little-endian, bit 7 of the last byte = (y > -y), bit 6 = infinity."""
    fq = curve.fq
    if P is None:
        b = bytearray(fq.nbytes)
        b[-1] |= 1 << 6
        return bytes(b)
    b = bytearray(P[0].to_bytes(fq.nbytes, "little"))
    if P[1] > (fq.p - P[1]) % fq.p:
        b[-1] |= 1 << 7
    return bytes(b)
