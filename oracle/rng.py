"""Random-number generators of the reference's tests and benches, restated bit for bit.

* `ChaChaRng(seed, rounds)`: rand_chacha 0.3 `ChaCha{8,12,20}Rng::from_seed` [U rand_chacha src/chacha.rs]:
  key = seed, 64-bit block counter starting at 0 in state words 12-13, stream id 0 in words 14-15.
  `BlockRng` hands out the keystream as 32-bit little-endian words in order; `next_u64` takes two
  consecutive words (low word first), also across a buffer boundary, so the generator is a pure
  function of (key, word position).
* `test_rng()`: ark-std 0.3 `test_rng` = rand 0.8 `StdRng` (= ChaCha12) with the fixed seed of SURVEY.md A.4.
* `fr_rand` / `fq_rand`: ark-ff 0.3 `UniformRand for Fp` -- rejection sampling on N u64 limbs, top
  REPR_SHAVE_BITS cleared; the accepted limbs ARE the Montgomery representation.
"""
import struct

MASK32 = 0xffffffff


def _rotl(x, n):
    return ((x << n) & MASK32) | (x >> (32 - n))


def _qr(s, a, b, c, d):
    s[a] = (s[a] + s[b]) & MASK32
    s[d] = _rotl(s[d] ^ s[a], 16)
    s[c] = (s[c] + s[d]) & MASK32
    s[b] = _rotl(s[b] ^ s[c], 12)
    s[a] = (s[a] + s[b]) & MASK32
    s[d] = _rotl(s[d] ^ s[a], 8)
    s[c] = (s[c] + s[d]) & MASK32
    s[b] = _rotl(s[b] ^ s[c], 7)


def chacha_block(key_words, counter, rounds):
    """16 output words of one ChaCha block (djb layout: 64-bit counter, 64-bit stream id = 0)."""
    init = [0x61707865, 0x3320646e, 0x79622d32, 0x6b206574] + list(key_words) + [
        counter & MASK32, (counter >> 32) & MASK32, 0, 0]
    s = list(init)
    for _ in range(rounds // 2):
        _qr(s, 0, 4, 8, 12)
        _qr(s, 1, 5, 9, 13)
        _qr(s, 2, 6, 10, 14)
        _qr(s, 3, 7, 11, 15)
        _qr(s, 0, 5, 10, 15)
        _qr(s, 1, 6, 11, 12)
        _qr(s, 2, 7, 8, 13)
        _qr(s, 3, 4, 9, 14)
    return [(s[i] + init[i]) & MASK32 for i in range(16)]


class ChaChaRng:
    def __init__(self, seed, rounds=20, word_pos=0):
        assert len(seed) == 32
        self.seed = bytes(seed)
        self.key = list(struct.unpack("<8I", self.seed))
        self.rounds = rounds
        self.word_pos = word_pos
        self._blk = None
        self._blk_idx = -1

    def _word(self, pos):
        b = pos >> 4
        if b != self._blk_idx:
            self._blk = chacha_block(self.key, b, self.rounds)
            self._blk_idx = b
        return self._blk[pos & 15]

    def next_u32(self):
        w = self._word(self.word_pos)
        self.word_pos += 1
        return w

    def next_u64(self):
        lo = self._word(self.word_pos)
        hi = self._word(self.word_pos + 1)
        self.word_pos += 2
        return lo | (hi << 32)


TEST_RNG_SEED = bytes([1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0] + [0] * 16)


def test_rng():
    """ark_std::test_rng()"""
    return ChaChaRng(TEST_RNG_SEED, rounds=12)


def field_rand(field, rng):
    """`F::rand(rng)`; returns the CANONICAL value of the sampled element."""
    n = field.limbs64
    top_mask = (1 << (64 - field.repr_shave_bits)) - 1
    while True:
        limbs = [rng.next_u64() for _ in range(n)]
        limbs[-1] &= top_mask
        v = 0
        for i, l in enumerate(limbs):
            v |= l << (64 * i)
        if v < field.p:
            return field.from_mont(v)  # the limbs are the Montgomery representation


def u128_rand(rng):
    lo = rng.next_u64()
    hi = rng.next_u64()
    return lo | (hi << 64)


def poly_rand(field, degree, rng):
    """`DensePolynomial::rand(d, rng)`: d + 1 coefficients, low to high, then trailing zeros stripped."""
    from .poly import strip
    return strip([field_rand(field, rng) for _ in range(degree + 1)])
