// Host-side transcript primitives of the prover: BLAKE2s-256, the ChaCha block function and the
// reference's `SimpleHashFiatShamirRng<Blake2s, ChaChaRng>` [reference src/rng.rs:18-80].
// These run on the CPU in the reference too and hash < 1 KB per round; they are protocol glue,
// not a compute fallback.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace b2m {

// ---- BLAKE2s (RFC 7693), unkeyed, 32-byte digest -------------------------------------------
struct Blake2s {
  uint32_t h[8];
  uint8_t buf[64];
  size_t buflen = 0;
  uint64_t t = 0;
  static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
  Blake2s() {
    static const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    for (int i = 0; i < 8; i++) h[i] = IV[i];
    h[0] ^= 0x01010000u ^ 32u;  // digest length 32, no key, fanout = depth = 1
  }
  void compress(const uint8_t* block, bool last) {
    static const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    static const uint8_t S[10][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
    uint32_t m[16], v[16];
    for (int i = 0; i < 16; i++) memcpy(&m[i], block + 4 * i, 4);
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = IV[i]; }
    v[12] ^= (uint32_t)t;
    v[13] ^= (uint32_t)(t >> 32);
    if (last) v[14] = ~v[14];
    auto G = [&](int a, int b, int c, int d, uint32_t x, uint32_t y) {
      v[a] = v[a] + v[b] + x; v[d] = rotr(v[d] ^ v[a], 16);
      v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 12);
      v[a] = v[a] + v[b] + y; v[d] = rotr(v[d] ^ v[a], 8);
      v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 7);
    };
    for (int r = 0; r < 10; r++) {
      const uint8_t* s = S[r];
      G(0, 4, 8, 12, m[s[0]], m[s[1]]);   G(1, 5, 9, 13, m[s[2]], m[s[3]]);
      G(2, 6, 10, 14, m[s[4]], m[s[5]]);  G(3, 7, 11, 15, m[s[6]], m[s[7]]);
      G(0, 5, 10, 15, m[s[8]], m[s[9]]);  G(1, 6, 11, 12, m[s[10]], m[s[11]]);
      G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
  }
  void update(const uint8_t* in, size_t len) {
    while (len > 0) {
      if (buflen == 64) {  // keep the final block for finalisation
        t += 64;
        compress(buf, false);
        buflen = 0;
      }
      size_t take = 64 - buflen;
      if (take > len) take = len;
      memcpy(buf + buflen, in, take);
      buflen += take; in += take; len -= take;
    }
  }
  void finish(uint8_t out[32]) {
    t += buflen;
    memset(buf + buflen, 0, 64 - buflen);
    compress(buf, true);
    memcpy(out, h, 32);
  }
  static void digest(const std::vector<uint8_t>& in, uint8_t out[32]) {
    Blake2s b;
    b.update(in.data(), in.size());
    b.finish(out);
  }
};

// ---- ChaCha block function (djb layout: 64-bit counter in words 12-13, stream id 0) -------------
inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
inline void chacha_block_host(const uint32_t key[8], uint64_t counter, int rounds, uint32_t out[16]) {
  uint32_t init[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3],
                       key[4], key[5], key[6], key[7], (uint32_t)counter, (uint32_t)(counter >> 32), 0, 0};
  uint32_t s[16];
  memcpy(s, init, sizeof(s));
  auto QR = [&](int a, int b, int c, int d) {
    s[a] += s[b]; s[d] = rotl32(s[d] ^ s[a], 16);
    s[c] += s[d]; s[b] = rotl32(s[b] ^ s[c], 12);
    s[a] += s[b]; s[d] = rotl32(s[d] ^ s[a], 8);
    s[c] += s[d]; s[b] = rotl32(s[b] ^ s[c], 7);
  };
  for (int r = 0; r < rounds / 2; r++) {
    QR(0, 4, 8, 12); QR(1, 5, 9, 13); QR(2, 6, 10, 14); QR(3, 7, 11, 15);
    QR(0, 5, 10, 15); QR(1, 6, 11, 12); QR(2, 7, 8, 13); QR(3, 4, 9, 14);
  }
  for (int i = 0; i < 16; i++) out[i] = s[i] + init[i];
}

// rand_chacha `BlockRng` semantics: the generator is a pure function of (key, word position);
// next_u64 = word[pos] | word[pos+1] << 32.
struct ChaChaHost {
  uint32_t key[8];
  int rounds = 20;
  uint64_t word_pos = 0;
  uint32_t blk[16];
  uint64_t blk_idx = ~0ull;
  ChaChaHost() { memset(key, 0, sizeof(key)); }
  ChaChaHost(const uint8_t seed[32], int rounds_, uint64_t pos = 0) : rounds(rounds_), word_pos(pos) { memcpy(key, seed, 32); }
  uint32_t word(uint64_t pos) {
    uint64_t b = pos >> 4;
    if (b != blk_idx) {
      chacha_block_host(key, b, rounds, blk);
      blk_idx = b;
    }
    return blk[pos & 15];
  }
  uint32_t next_u32() { return word(word_pos++); }
  uint64_t next_u64() {
    uint64_t lo = word(word_pos), hi = word(word_pos + 1);
    word_pos += 2;
    return lo | (hi << 32);
  }
};

// The caller's rng behind the C ABI (include/b2m.h b2m_rng): a ChaCha stream position (device-samplable) or a host
// callback standing for any `RngCore`.  `Rng` below is any type with next_u64().
template <class RngDesc>
struct ZkSource {
  RngDesc* desc;
  ChaChaHost cc;
  bool callback;
  explicit ZkSource(RngDesc* d) : desc(d), callback(d != nullptr && d->kind == 1 /* B2M_RNG_CALLBACK */) {
    if (d && !callback) cc = ChaChaHost(d->key, d->kind, d->word_pos);
  }
  uint64_t next_u64() { return callback ? desc->next_u64(desc->state) : cc.next_u64(); }
  void commit_position() {  // report the stream position back to the caller (ChaCha form)
    if (desc && !callback) desc->word_pos = cc.word_pos;
  }
};

// `F::rand(rng)` of ark-ff 0.3: rejection sampling on limbs; accepted limbs are the Montgomery form.
template <class F, class Rng>
F field_rand(Rng& rng) {
  constexpr int L64 = F::N / 2;
  constexpr int shave = 64 * L64 - F::Params::BITS;
  for (;;) {
    F v;
    for (int i = 0; i < L64; i++) {
      uint64_t x = rng.next_u64();
      if (i == L64 - 1) x &= (~0ull) >> shave;
      v.l[2 * i] = (uint32_t)x;
      v.l[2 * i + 1] = (uint32_t)(x >> 32);
    }
    // valid iff v < p
    bool lt = false;
    for (int i = F::N - 1; i >= 0; i--) {
      uint32_t m = F::Params::mod(i);
      if (v.l[i] != m) { lt = v.l[i] < m; break; }
    }
    if (lt) return v;
  }
}

// `SimpleHashFiatShamirRng<Blake2s, ChaChaRng>`
struct FiatShamir {
  uint8_t seed[32];
  ChaChaHost r;
  explicit FiatShamir(const std::vector<uint8_t>& initial) {
    Blake2s::digest(initial, seed);
    r = ChaChaHost(seed, 20);
  }
  void absorb(const std::vector<uint8_t>& bytes) {
    std::vector<uint8_t> in(bytes);
    in.insert(in.end(), seed, seed + 32);
    Blake2s::digest(in, seed);
    r = ChaChaHost(seed, 20);
  }
  uint32_t next_u32() { return r.next_u32(); }
  uint64_t next_u64() { return r.next_u64(); }
};

inline void put_u64(std::vector<uint8_t>& out, uint64_t v) {
  for (int i = 0; i < 8; i++) out.push_back((uint8_t)(v >> (8 * i)));
}

}  // namespace b2m
