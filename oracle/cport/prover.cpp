/* oracle/cport/prover.cpp -- C++/OpenMP restatement of the reference's CPU prover:
 * `Marlin::index` + `Marlin::prove` [reference src/lib.rs:100-311] over the AHP of src/ahp/{indexer,
 * constraint_systems,prover,verifier,mod}.rs, KZG10 / MarlinKZG10 / SonicKZG10 [U ark-poly-commit 0.3] and
 * SimpleHashFiatShamirRng<Blake2s, ChaChaRng> [reference src/rng.rs], using the reference's algorithms for the
 * heavy steps (ark-ec Pippenger MSM, radix-2 FFTs, polynomial products through FFTs on the reference's domains).
 * It is a transliteration of oracle/{ahp,kzg,marlin}.py (the Python specification) and must emit the same bytes.
 * TEST INFRASTRUCTURE: a fast byte-exact checker for sizes the Python oracle cannot reach, and bench.py's
 * `--impl reference` / cpu_baseline arm.  Never linked into, or called by, the product. */
#include <algorithm>
#include <cstdio>
#include <string>
#include <vector>

extern "C" {
#include "cport_core.h"
}

namespace {

typedef f4_t Fr;

/* ---- curve traits ------------------------------------------------------------------------------------------- */
struct Bls {
  typedef bls_aff Aff;
  typedef bls_jac Jac;
  typedef f6_t Fq;
  static const int FQ_BYTES = 48;
  static const int SCALAR_BITS = 255;
  static const f4_ctx& fr() { return BLS_FR; }
  static const f6_ctx& fq() { return BLS_FQ; }
  static u64 gen() { return BLS_FR_GEN; }
  static int two_adicity() { return BLS_FR_S; }
  static int shave_bits() { return 1; }
  static void msm(Jac* out, const Aff* b, const u64* s, size_t n, int threads) { bls_msm(out, b, s, n, SCALAR_BITS, &BLS_FQ, threads); }
  static void add(Jac* p, const Jac* q) { bls_add(p, q, &BLS_FQ); }
  static void set_inf(Jac* p) { bls_jac_set_inf(p); }
  static void to_affine(Aff* r, const Jac* p) { bls_to_affine(r, p, &BLS_FQ); }
  static bool aff_is_inf(const Aff* p) { return bls_aff_is_inf(p); }
  static void fq_from_mont(Fq* r, const Fq* a) { f6_from_mont(r, a, &BLS_FQ); }
  static void fq_neg(Fq* r, const Fq* a) { f6_neg(r, a, &BLS_FQ); }
};
struct Bn {
  typedef bn_aff Aff;
  typedef bn_jac Jac;
  typedef f4_t Fq;
  static const int FQ_BYTES = 32;
  static const int SCALAR_BITS = 254;
  static const f4_ctx& fr() { return BN_FR; }
  static const f4_ctx& fq() { return BN_FQ; }
  static u64 gen() { return BN_FR_GEN; }
  static int two_adicity() { return BN_FR_S; }
  static int shave_bits() { return 2; }
  static void msm(Jac* out, const Aff* b, const u64* s, size_t n, int threads) { bn_msm(out, b, s, n, SCALAR_BITS, &BN_FQ, threads); }
  static void add(Jac* p, const Jac* q) { bn_add(p, q, &BN_FQ); }
  static void set_inf(Jac* p) { bn_jac_set_inf(p); }
  static void to_affine(Aff* r, const Jac* p) { bn_to_affine(r, p, &BN_FQ); }
  static bool aff_is_inf(const Aff* p) { return bn_aff_is_inf(p); }
  static void fq_from_mont(Fq* r, const Fq* a) { f4_from_mont(r, a, &BN_FQ); }
  static void fq_neg(Fq* r, const Fq* a) { f4_neg(r, a, &BN_FQ); }
};

/* ---- BLAKE2s / ChaCha (independent of the product's hostutil.hpp) -------------------------------------------- */
static inline uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
static void blake2s(const std::vector<uint8_t>& in, uint8_t out[32]) {
  static const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
  static const uint8_t S[10][16] = {{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
      {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
      {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
      {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
      {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
  uint32_t h[8];
  for (int i = 0; i < 8; i++) h[i] = IV[i];
  h[0] ^= 0x01010020u;
  size_t len = in.size(), off = 0;
  uint64_t t = 0;
  for (;;) {
    uint8_t blk[64] = {0};
    size_t take = len - off > 64 ? 64 : len - off;
    bool last = (off + take == len);
    if (take) memcpy(blk, in.data() + off, take);
    t += take;
    uint32_t m[16], v[16];
    for (int i = 0; i < 16; i++) memcpy(&m[i], blk + 4 * i, 4);
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = IV[i]; }
    v[12] ^= (uint32_t)t; v[13] ^= (uint32_t)(t >> 32);
    if (last) v[14] = ~v[14];
#define B2G(a, b, c, d, x, y) v[a] += v[b] + (x); v[d] = rotr32(v[d] ^ v[a], 16); v[c] += v[d]; v[b] = rotr32(v[b] ^ v[c], 12); \
  v[a] += v[b] + (y); v[d] = rotr32(v[d] ^ v[a], 8); v[c] += v[d]; v[b] = rotr32(v[b] ^ v[c], 7);
    for (int r = 0; r < 10; r++) {
      const uint8_t* s = S[r];
      B2G(0, 4, 8, 12, m[s[0]], m[s[1]]) B2G(1, 5, 9, 13, m[s[2]], m[s[3]]) B2G(2, 6, 10, 14, m[s[4]], m[s[5]]) B2G(3, 7, 11, 15, m[s[6]], m[s[7]])
      B2G(0, 5, 10, 15, m[s[8]], m[s[9]]) B2G(1, 6, 11, 12, m[s[10]], m[s[11]]) B2G(2, 7, 8, 13, m[s[12]], m[s[13]]) B2G(3, 4, 9, 14, m[s[14]], m[s[15]])
    }
#undef B2G
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
    off += take;
    if (last) break;
  }
  memcpy(out, h, 32);
}
struct ChaCha {
  uint32_t key[8];
  int rounds;
  uint64_t pos;
  uint32_t blk[16];
  uint64_t blk_idx;
  ChaCha() : rounds(20), pos(0), blk_idx(~0ull) { memset(key, 0, sizeof(key)); }
  ChaCha(const uint8_t* seed, int r, uint64_t p) : rounds(r), pos(p), blk_idx(~0ull) { memcpy(key, seed, 32); }
  void block(uint64_t ctr) {
    uint32_t init[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                         (uint32_t)ctr, (uint32_t)(ctr >> 32), 0, 0};
    uint32_t s[16];
    memcpy(s, init, sizeof(s));
#define CQR(a, b, c, d) s[a] += s[b]; s[d] = rotl32(s[d] ^ s[a], 16); s[c] += s[d]; s[b] = rotl32(s[b] ^ s[c], 12); \
  s[a] += s[b]; s[d] = rotl32(s[d] ^ s[a], 8); s[c] += s[d]; s[b] = rotl32(s[b] ^ s[c], 7);
    for (int r = 0; r < rounds / 2; r++) { CQR(0, 4, 8, 12) CQR(1, 5, 9, 13) CQR(2, 6, 10, 14) CQR(3, 7, 11, 15) CQR(0, 5, 10, 15) CQR(1, 6, 11, 12) CQR(2, 7, 8, 13) CQR(3, 4, 9, 14) }
#undef CQR
    for (int i = 0; i < 16; i++) blk[i] = s[i] + init[i];
    blk_idx = ctr;
  }
  uint32_t word(uint64_t p) { if ((p >> 4) != blk_idx) block(p >> 4); return blk[p & 15]; }
  uint64_t next_u64() { uint64_t lo = word(pos), hi = word(pos + 1); pos += 2; return lo | (hi << 32); }
};
struct FiatShamir {
  uint8_t seed[32];
  ChaCha r;
  explicit FiatShamir(const std::vector<uint8_t>& init) { blake2s(init, seed); r = ChaCha(seed, 20, 0); }
  void absorb(const std::vector<uint8_t>& b) {
    std::vector<uint8_t> in(b);
    in.insert(in.end(), seed, seed + 32);
    blake2s(in, seed);
    r = ChaCha(seed, 20, 0);
  }
  uint64_t next_u64() { return r.next_u64(); }
};

/* ---- the prover, templated on the curve ------------------------------------------------------------------------ */
static int g_skip_index_commit = 0;

template <class C>
struct Marlin {
  typedef typename C::Aff Aff;
  typedef typename C::Jac Jac;
  typedef std::vector<Fr> Poly;
  int threads;
  bool sonic;

  static const f4_ctx& F() { return C::fr(); }
  static Fr zero() { Fr z; memset(&z, 0, sizeof(z)); return z; }
  static Fr one() { Fr o; memcpy(o.l, F().r, 32); return o; }
  static Fr add(const Fr& a, const Fr& b) { Fr r; f4_add(&r, &a, &b, &F()); return r; }
  static Fr sub(const Fr& a, const Fr& b) { Fr r; f4_sub(&r, &a, &b, &F()); return r; }
  static Fr mul(const Fr& a, const Fr& b) { Fr r; f4_mul(&r, &a, &b, &F()); return r; }
  static Fr neg(const Fr& a) { Fr r; f4_neg(&r, &a, &F()); return r; }
  static Fr inv(const Fr& a) { Fr r; f4_inv(&r, &a, &F()); return r; }
  static Fr from_u64(u64 v) { Fr c = {{v, 0, 0, 0}}; Fr r; f4_to_mont(&r, &c, &F()); return r; }
  static Fr pow_u64(const Fr& a, u64 e) { Fr r; u64 ee[1] = {e}; f4_pow(&r, &a, ee, 1, &F()); return r; }
  static bool is_zero(const Fr& a) { return f4_is_zero(&a); }
  static bool eq(const Fr& a, const Fr& b) { return f4_eq(&a, &b); }
  static Fr canonical(const Fr& a) { Fr r; f4_from_mont(&r, &a, &F()); return r; }

  template <class Rng>
  static Fr field_rand(Rng& rng) {  /* ark-ff UniformRand: rejection sampling, accepted limbs ARE the Montgomery form */
    for (;;) {
      Fr v;
      for (int i = 0; i < 4; i++) v.l[i] = rng.next_u64();
      v.l[3] &= (~0ull) >> C::shave_bits();
      if (!f4_geq_p(v.l, &F())) return v;
    }
  }

  struct Domain {  /* Radix2EvaluationDomain */
    size_t n;
    int log;
    Fr gen, n_inv;
    explicit Domain(size_t num_coeffs) {
      n = 1; log = 0;
      while (n < num_coeffs) { n <<= 1; log++; }
      fr_root(&gen, &F(), C::gen(), C::two_adicity(), log);
      n_inv = inv(from_u64(n));
    }
    Fr element(size_t i) const { return pow_u64(gen, i); }
    Fr vanishing(const Fr& tau) const { return sub(pow_u64(tau, n), one()); }
  };
  void fft(const Domain& d, Poly& v, bool inverse) const {
    v.resize(d.n, zero());
    fr_fft(v.data(), d.log, inverse ? 1 : 0, &F(), C::gen(), C::two_adicity(), threads);
  }
  static void strip(Poly& p) { while (!p.empty() && is_zero(p.back())) p.pop_back(); }
  static size_t degree(const Poly& p) { return p.empty() ? 0 : p.size() - 1; }

  void batch_inversion(Poly& v) const {  /* Montgomery's trick per chunk; zeros untouched */
    size_t n = v.size();
    int nt = threads;
    #pragma omp parallel for num_threads(threads)
    for (int t = 0; t < nt; t++) {
      size_t lo = n * t / nt, hi = n * (t + 1) / nt;
      if (lo >= hi) continue;
      std::vector<Fr> pre(hi - lo);
      Fr acc = one();
      for (size_t i = lo; i < hi; i++) { pre[i - lo] = acc; if (!is_zero(v[i])) acc = mul(acc, v[i]); }
      Fr ia = inv(acc);
      for (size_t i = hi; i-- > lo;) {
        if (is_zero(v[i])) continue;
        Fr nv = mul(ia, pre[i - lo]);
        ia = mul(ia, v[i]);
        v[i] = nv;
      }
    }
  }
  Fr evaluate(const Poly& p, const Fr& x) const {
    size_t n = p.size();
    if (n == 0) return zero();
    int nt = threads;
    std::vector<Fr> part(nt, zero());
    size_t chunk = (n + nt - 1) / nt;
    #pragma omp parallel for num_threads(threads)
    for (int t = 0; t < nt; t++) {
      size_t lo = chunk * t, hi = std::min(n, lo + chunk);
      if (lo >= hi) continue;
      Fr acc = zero();
      for (size_t i = hi; i-- > lo;) acc = add(mul(acc, x), p[i]);
      part[t] = mul(acc, pow_u64(x, lo));
    }
    Fr r = zero();
    for (int t = 0; t < nt; t++) r = add(r, part[t]);
    return r;
  }
  Poly poly_mul(const Poly& a, const Poly& b) const {  /* DensePolynomial Mul: FFT on next_pow2(len_a + len_b - 1) */
    if (a.empty() || b.empty()) return Poly();
    Domain d(a.size() + b.size() - 1);
    Poly ea(a), eb(b);
    fft(d, ea, false); fft(d, eb, false);
    #pragma omp parallel for num_threads(threads)
    for (size_t i = 0; i < d.n; i++) ea[i] = mul(ea[i], eb[i]);
    fft(d, ea, true);
    strip(ea);
    return ea;
  }
  static void divide_by_vanishing(const Poly& c, size_t n, Poly& q, Poly& r) {
    if (c.size() < n) { q.clear(); r = c; strip(r); return; }
    q.assign(c.begin() + n, c.end());
    if (q.size() > n) for (size_t i = q.size() - n; i-- > 0;) q[i] = add(q[i], q[i + n]);
    r.assign(c.begin(), c.begin() + n);
    for (size_t i = 0; i < std::min(n, q.size()); i++) r[i] = add(r[i], q[i]);
    strip(q); strip(r);
  }
  static Poly mul_by_vanishing(const Poly& c, size_t n) {
    Poly out(n, zero());
    out.insert(out.end(), c.begin(), c.end());
    for (size_t i = 0; i < c.size(); i++) out[i] = sub(out[i], c[i]);
    strip(out);
    return out;
  }
  static Poly divide_by_linear(const Poly& c, const Fr& z) {  /* quotient of c / (X - z) */
    if (c.size() <= 1) return Poly();
    Poly q(c.size() - 1);
    Fr acc = zero();
    for (size_t i = c.size() - 1; i >= 1; i--) { acc = add(c[i], mul(acc, z)); q[i - 1] = acc; }
    strip(q);
    return q;
  }
  static void axpy(Poly& acc, const Fr& k, const Poly& p) {  /* acc += k * p */
    if (acc.size() < p.size()) acc.resize(p.size(), zero());
    for (size_t i = 0; i < p.size(); i++) acc[i] = add(acc[i], mul(k, p[i]));
  }

  /* ---- SRS / KZG ------------------------------------------------------------------------------------------------ */
  std::vector<Aff> powers, gamma;
  std::vector<u64> gamma_idx;
  size_t D;
  size_t gamma_slot(u64 i) const { for (size_t k = 0; k < gamma_idx.size(); k++) if (gamma_idx[k] == i) return k; return (size_t)-1; }

  Aff msm_powers(size_t off, const Poly& coeffs) const {  /* skip_leading_zeros + into_repr + VariableBaseMSM */
    size_t lz = 0;
    while (lz < coeffs.size() && is_zero(coeffs[lz])) lz++;
    size_t n = coeffs.size() - lz;
    std::vector<u64> sc(4 * (n ? n : 1));
    #pragma omp parallel for num_threads(threads)
    for (size_t i = 0; i < n; i++) { Fr c = canonical(coeffs[lz + i]); memcpy(&sc[4 * i], c.l, 32); }
    Jac r;
    C::msm(&r, powers.data() + off + lz, sc.data(), n, threads);
    Aff a;
    C::to_affine(&a, &r);
    return a;
  }
  Aff msm_gamma(size_t slot, const Poly& coeffs) const {
    std::vector<u64> sc(4 * (coeffs.size() ? coeffs.size() : 1));
    for (size_t i = 0; i < coeffs.size(); i++) { Fr c = canonical(coeffs[i]); memcpy(&sc[4 * i], c.l, 32); }
    Jac r;
    C::msm(&r, gamma.data() + slot, sc.data(), coeffs.size(), 1);
    Aff a;
    C::to_affine(&a, &r);
    return a;
  }
  static Aff aff_add(const Aff& a, const Aff& b) {
    Jac j; C::set_inf(&j);
    Jac ja, jb; C::set_inf(&ja); C::set_inf(&jb);
    /* lift through add_mixed semantics: reuse msm-free path */
    if (!C::aff_is_inf(&a)) { ja.x = a.x; ja.y = a.y; memcpy(&ja.z, one_fq(), sizeof(ja.z)); }
    if (!C::aff_is_inf(&b)) { jb.x = b.x; jb.y = b.y; memcpy(&jb.z, one_fq(), sizeof(jb.z)); }
    j = ja; C::add(&j, &jb);
    Aff r; C::to_affine(&r, &j);
    return r;
  }
  static const void* one_fq() { return C::fq().r; }

  struct Labeled {
    std::string label;
    Poly c;
    long bound;  /* -1 = None */
    bool hiding;
    Poly rand, shifted_rand;
    bool has_shifted_rand;
    Aff comm, shifted;
    Labeled() : bound(-1), hiding(false), has_shifted_rand(false) { memset(&comm, 0, sizeof(comm)); memset(&shifted, 0, sizeof(shifted)); }
  };
  template <class Rng>
  Aff kzg_commit(size_t off, size_t gslot, const Poly& coeffs, bool hiding, Rng* rng, Poly& blinding) const {
    Aff c = msm_powers(off, coeffs);
    blinding.clear();
    if (hiding) { for (int k = 0; k < 3; k++) blinding.push_back(field_rand(*rng)); strip(blinding); }
    if (!blinding.empty()) c = aff_add(c, msm_gamma(gslot, blinding));
    return c;
  }
  template <class Rng>
  void commit(std::vector<Labeled*>& polys, Rng* rng) const {  /* PC::commit */
    for (Labeled* p : polys) {
      if (!sonic) {
        p->comm = kzg_commit(0, gamma_slot(0), p->c, p->hiding, rng, p->rand);
        if (p->bound >= 0) { p->shifted = kzg_commit(D - (size_t)p->bound, gamma_slot(0), p->c, p->hiding, rng, p->shifted_rand); p->has_shifted_rand = true; }
      } else if (p->bound >= 0) {
        p->comm = kzg_commit(D - (size_t)p->bound, p->hiding ? gamma_slot(D - (size_t)p->bound) : 0, p->c, p->hiding, rng, p->rand);
      } else {
        p->comm = kzg_commit(0, gamma_slot(0), p->c, p->hiding, rng, p->rand);
      }
    }
  }
  struct Opening { Aff w; bool has_rv; Fr rv; };
  Opening open_with_witness(size_t off, const Fr& point, const Poly& rand_poly, const Poly& witness, bool has_hw, const Poly& hw) const {
    Opening o;
    o.w = msm_powers(off, witness);
    o.has_rv = false;
    o.rv = zero();
    if (has_hw) {
      o.has_rv = true;
      o.rv = evaluate(rand_poly, point);
      o.w = aff_add(o.w, msm_gamma(gamma_slot(0), hw));
    }
    return o;
  }
  Opening open_at_point(std::vector<Labeled*>& polys, const Fr& point, const Fr& xi, size_t max_bound) const {
    std::vector<Fr> xp(2 * polys.size() + 2);
    xp[0] = one();
    for (size_t i = 1; i < xp.size(); i++) xp[i] = mul(xp[i - 1], xi);
    Poly comb, r;
    if (sonic) {
      size_t k = 0;
      for (Labeled* p : polys) { axpy(comb, xp[k], p->c); axpy(r, xp[k], p->rand); k++; }
      strip(comb); strip(r);
      Poly w = divide_by_linear(comb, point);
      bool hid = !r.empty();
      Poly hw = hid ? divide_by_linear(r, point) : Poly();
      return open_with_witness(0, point, r, w, hid, hw);
    }
    Poly sw, sr, srw;
    bool enforce = false;
    size_t k = 0;
    for (Labeled* p : polys) {
      axpy(comb, xp[k], p->c); axpy(r, xp[k], p->rand);
      k++;
      if (p->bound >= 0) {
        enforce = true;
        Poly w = divide_by_linear(p->c, point);
        Poly sh;
        if (!w.empty()) { sh.assign(max_bound - (size_t)p->bound, zero()); sh.insert(sh.end(), w.begin(), w.end()); }
        axpy(sw, xp[k], sh); axpy(sr, xp[k], p->shifted_rand);
        Poly srp(p->shifted_rand); strip(srp);
        if (!srp.empty()) axpy(srw, xp[k], divide_by_linear(srp, point));
        k++;
      }
    }
    strip(comb); strip(r); strip(sw); strip(sr); strip(srw);
    Poly w = divide_by_linear(comb, point);
    bool hid = !r.empty();
    Poly hw = hid ? divide_by_linear(r, point) : Poly();
    Opening o = open_with_witness(0, point, r, w, hid, hw);
    if (enforce) {
      Opening so = open_with_witness(D - max_bound, point, sr, sw, true, srw);
      o.w = aff_add(o.w, so.w);
      if (o.has_rv) o.rv = add(o.rv, so.rv);
    }
    return o;
  }

  /* ---- index ---------------------------------------------------------------------------------------------------- */
  size_t nc, nv, ni, nnz, H, K, X;
  struct Entry { u64 col; Fr v; };
  std::vector<std::vector<Entry>> A, B, Cm;
  Poly ev_row, ev_col, ev_rc, ev_a, ev_b, ev_c;
  Labeled ip[6];  /* row, col, a_val, b_val, c_val, row_col */
  std::vector<uint8_t> vk_bytes;

  size_t reindex(size_t i) const { size_t period = H / X; if (i < X) return i * period; size_t j = i - X, x = period - 1; return j + j / x + 1; }
  static void put_u64(std::vector<uint8_t>& o, u64 v) { for (int i = 0; i < 8; i++) o.push_back((uint8_t)(v >> (8 * i))); }
  static void put_fr(std::vector<uint8_t>& o, const Fr& m) { Fr c = canonical(m); const uint8_t* p = (const uint8_t*)c.l; o.insert(o.end(), p, p + 32); }
  static void put_fq(std::vector<uint8_t>& o, const typename C::Fq& m) { typename C::Fq c; C::fq_from_mont(&c, &m); const uint8_t* p = (const uint8_t*)c.l; o.insert(o.end(), p, p + C::FQ_BYTES); }
  static void put_aff_tobytes(std::vector<uint8_t>& o, const Aff& P) {
    if (C::aff_is_inf(&P)) { o.insert(o.end(), C::FQ_BYTES, 0); o.push_back(1); o.insert(o.end(), C::FQ_BYTES - 1, 0); o.push_back(1); return; }
    put_fq(o, P.x); put_fq(o, P.y); o.push_back(0);
  }
  void put_commitment(std::vector<uint8_t>& o, const Labeled& p) const {
    put_aff_tobytes(o, p.comm);
    if (!sonic) { Aff none; memset(&none, 0, sizeof(none)); o.push_back(p.bound >= 0 ? 1 : 0); put_aff_tobytes(o, p.bound >= 0 ? p.shifted : none); }
  }
  static void put_compressed(std::vector<uint8_t>& o, const Aff& P) {
    if (C::aff_is_inf(&P)) { o.insert(o.end(), C::FQ_BYTES - 1, 0); o.push_back(1 << 6); return; }
    typename C::Fq x, y, ny, nyc;
    C::fq_from_mont(&x, &P.x); C::fq_from_mont(&y, &P.y);
    C::fq_neg(&ny, &P.y); C::fq_from_mont(&nyc, &ny);
    size_t at = o.size();
    const uint8_t* p = (const uint8_t*)x.l;
    o.insert(o.end(), p, p + C::FQ_BYTES);
    bool larger = false;
    for (int i = (int)(sizeof(y.l) / 8) - 1; i >= 0; i--) if (y.l[i] != nyc.l[i]) { larger = y.l[i] > nyc.l[i]; break; }
    if (larger) o[at + C::FQ_BYTES - 1] |= 1 << 7;
  }

  int build_index(const u64* rp[3], const u64* cl[3], const u64* cf[3]) {
    if (nc != nv) return 5;
    if (ni == 0 || (ni & (ni - 1))) return 4;
    std::vector<std::vector<Entry>>* M[3] = {&A, &B, &Cm};
    for (int m = 0; m < 3; m++) {
      M[m]->assign(nc, std::vector<Entry>());
      for (size_t r = 0; r < nc; r++)
        for (u64 e = rp[m][r]; e < rp[m][r + 1]; e++) { Entry en; en.col = cl[m][e]; memcpy(en.v.l, cf[m] + 4 * e, 32); (*M[m])[r].push_back(en); }
    }
    Domain dh(nc); H = dh.n; X = ni;
    std::vector<Fr> elems(H);
    elems[0] = one();
    for (size_t i = 1; i < H; i++) elems[i] = mul(elems[i - 1], dh.gen);
    Fr h_inv = inv(from_u64(H));
    Poly row, col, va, vb, vc;
    for (size_t r = 0; r < nc; r++) {
      std::vector<u64> cols;
      for (int m = 0; m < 3; m++) for (auto& e : (*M[m])[r]) cols.push_back(e.col);
      std::sort(cols.begin(), cols.end());
      cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
      for (u64 c : cols) {
        Fr colv = elems[reindex(c)], rowv = elems[r];
        Fr sc = mul(colv, h_inv);  /* 1 / u_H(col_val, col_val) */
        row.push_back(colv); col.push_back(rowv);
        Fr v[3];
        for (int m = 0; m < 3; m++) { v[m] = zero(); for (auto& e : (*M[m])[r]) if (e.col == c) v[m] = e.v; }
        va.push_back(mul(v[0], sc)); vb.push_back(mul(v[1], sc)); vc.push_back(mul(v[2], sc));
      }
    }
    nnz = row.size();
    Domain dk(nnz); K = dk.n;
    size_t md = std::max(std::max(2 * H - 1, 3 * H - 1), K - 1);
    if (D < md) return 2;
    row.resize(K, elems[0]); col.resize(K, elems[0]);
    va.resize(K, zero()); vb.resize(K, zero()); vc.resize(K, zero());
    Poly rc(K);
    for (size_t i = 0; i < K; i++) rc[i] = mul(row[i], col[i]);
    ev_row = row; ev_col = col; ev_rc = rc; ev_a = va; ev_b = vb; ev_c = vc;
    const char* labels[6] = {"row", "col", "a_val", "b_val", "c_val", "row_col"};
    Poly* src[6] = {&row, &col, &va, &vb, &vc, &rc};
    std::vector<Labeled*> lp;
    for (int i = 0; i < 6; i++) { ip[i].label = labels[i]; ip[i].c = *src[i]; fft(dk, ip[i].c, true); strip(ip[i].c); lp.push_back(&ip[i]); }
    /* timing-only set-up (bench.py reference arm): the six index commitments are 6 MSMs of |K| that only feed the
       transcript hash -- skipping them leaves every operation of `prove` in place (proofs then do not verify) */
    if (g_skip_index_commit) { for (auto* q : lp) memset(&q->comm, 0, sizeof(q->comm)); }
    else commit<ChaCha>(lp, nullptr);
    vk_bytes.clear();
    put_u64(vk_bytes, nv); put_u64(vk_bytes, nc); put_u64(vk_bytes, nnz);
    for (int i = 0; i < 6; i++) put_commitment(vk_bytes, ip[i]);
    return 0;
  }

  /* ---- prove ---------------------------------------------------------------------------------------------------- */
  Fr sample_outside(FiatShamir& fs) const { for (;;) { Fr t = field_rand(fs); if (!eq(pow_u64(t, H), one())) return t; } }
  void absorb(FiatShamir& fs, std::vector<Labeled*>& ps) const { std::vector<uint8_t> b; for (Labeled* p : ps) put_commitment(b, *p); fs.absorb(b); }

  int prove(const u64* input, size_t n_input, const u64* witness, size_t n_witness, ChaCha& zk, std::vector<uint8_t>& proof) {
    if (n_input + n_witness != nv) return 3;
    if (n_input != ni) return 4;
    std::vector<Fr> z(nv);
    memcpy(z.data(), input, 32 * n_input);
    if (n_witness) memcpy(z.data() + n_input, witness, 32 * n_witness);
    /* prover_init: z_A, z_B (serial mat-vec like the reference) */
    Poly z_a(nc), z_b(nc);
    for (size_t r = 0; r < nc; r++) {
      Fr a = zero(), b = zero();
      for (auto& e : A[r]) a = add(a, mul(e.v, z[e.col]));
      for (auto& e : B[r]) b = add(b, mul(e.v, z[e.col]));
      z_a[r] = a; z_b[r] = b;
    }
    std::vector<uint8_t> init;
    const char* proto = "MARLIN-2019";
    init.insert(init.end(), proto, proto + 11);
    init.insert(init.end(), vk_bytes.begin(), vk_bytes.end());
    for (size_t i = 1; i < ni; i++) put_fr(init, z[i]);
    FiatShamir fs(init);
    Domain dh(nc), dk(nnz), dx(ni);

    /* ---- first round */
    Poly x_poly(z.begin(), z.begin() + ni);
    fft(dx, x_poly, true); strip(x_poly);
    Poly x_evals(x_poly);
    fft(dh, x_evals, false);
    size_t ratio = H / X;
    Poly w_evals(H);
    for (size_t k = 0; k < H; k++) {
      if (k % ratio == 0) { w_evals[k] = zero(); continue; }
      size_t j = k - k / ratio - 1;
      Fr wv = j < n_witness ? z[ni + j] : zero();
      w_evals[k] = sub(wv, x_evals[k]);
    }
    auto blind = [&](Poly& c, const Fr& rho) { c.resize(H + 1, zero()); c[0] = sub(c[0], rho); c[H] = add(c[H], rho); strip(c); };
    Labeled o_w, o_za, o_zb, o_mask;
    { Poly w(w_evals); fft(dh, w, true); blind(w, field_rand(zk)); Poly q, r; divide_by_vanishing(w, X, q, r); o_w.c = q; }
    { Poly p(z_a); fft(dh, p, true); blind(p, field_rand(zk)); o_za.c = p; }
    { Poly p(z_b); fft(dh, p, true); blind(p, field_rand(zk)); o_zb.c = p; }
    {
      size_t md = 3 * H - 1;
      Poly m(md + 1);
      for (size_t i = 0; i <= md; i++) m[i] = field_rand(zk);
      Fr r0 = zero();
      for (size_t i = 0; i <= md / H; i++) r0 = add(r0, m[H * i]);
      m[0] = sub(m[0], r0);
      strip(m);
      o_mask.c = m;
    }
    o_w.label = "w"; o_w.hiding = true; o_za.label = "z_a"; o_za.hiding = true; o_zb.label = "z_b"; o_zb.hiding = true; o_mask.label = "mask_poly";
    std::vector<Labeled*> first = {&o_w, &o_za, &o_zb, &o_mask};
    commit(first, &zk);
    absorb(fs, first);
    Fr alpha = sample_outside(fs), eta_a = field_rand(fs), eta_b = field_rand(fs), eta_c = field_rand(fs);

    /* ---- second round */
    Poly z_c = poly_mul(o_za.c, o_zb.c);
    Poly summed(z_c);
    for (auto& c : summed) c = mul(c, eta_c);
    for (size_t i = 0; i < std::min(summed.size(), std::min(o_za.c.size(), o_zb.c.size())); i++)
      summed[i] = add(summed[i], add(mul(eta_a, o_za.c[i]), mul(eta_b, o_zb.c[i])));
    strip(summed);
    Fr v_h_alpha = dh.vanishing(alpha);
    Poly r_alpha_ev(H);
    { Fr e = one(); for (size_t i = 0; i < H; i++) { r_alpha_ev[i] = sub(alpha, e); e = mul(e, dh.gen); } }
    batch_inversion(r_alpha_ev);
    for (auto& c : r_alpha_ev) c = mul(c, v_h_alpha);
    Poly r_alpha_poly(r_alpha_ev);
    fft(dh, r_alpha_poly, true); strip(r_alpha_poly);
    Poly t_poly(H, zero());
    {
      std::vector<std::vector<Entry>>* M[3] = {&A, &B, &Cm};
      Fr etas[3] = {eta_a, eta_b, eta_c};
      for (int m = 0; m < 3; m++)
        for (size_t r = 0; r < nc; r++)
          for (auto& e : (*M[m])[r]) { size_t j = reindex(e.col); t_poly[j] = add(t_poly[j], mul(mul(etas[m], e.v), r_alpha_ev[r])); }
      fft(dh, t_poly, true); strip(t_poly);
    }
    Poly z_poly = mul_by_vanishing(o_w.c, X);
    if (z_poly.size() < x_poly.size()) z_poly.resize(x_poly.size(), zero());
    for (size_t i = 0; i < x_poly.size(); i++) z_poly[i] = add(z_poly[i], x_poly[i]);
    strip(z_poly);
    size_t mul_size = std::max(o_mask.c.size(), std::max(r_alpha_poly.size() + summed.size(), t_poly.size() + z_poly.size()));
    Domain dm(mul_size);
    Poly ra(r_alpha_poly), sz(summed), ze(z_poly), te(t_poly);
    fft(dm, ra, false); fft(dm, sz, false); fft(dm, ze, false); fft(dm, te, false);
    #pragma omp parallel for num_threads(threads)
    for (size_t i = 0; i < dm.n; i++) ra[i] = sub(mul(ra[i], sz[i]), mul(ze[i], te[i]));
    fft(dm, ra, true); strip(ra);
    Poly q1(o_mask.c);
    if (q1.size() < ra.size()) q1.resize(ra.size(), zero());
    for (size_t i = 0; i < ra.size(); i++) q1[i] = add(q1[i], ra[i]);
    strip(q1);
    Labeled o_t, o_g1, o_h1;
    { Poly h1, xg1; divide_by_vanishing(q1, H, h1, xg1); o_h1.c = h1; if (!xg1.empty()) o_g1.c.assign(xg1.begin() + 1, xg1.end()); strip(o_g1.c); }
    o_t.c = t_poly; o_t.label = "t"; o_g1.label = "g_1"; o_g1.bound = (long)H - 2; o_g1.hiding = true; o_h1.label = "h_1";
    std::vector<Labeled*> second = {&o_t, &o_g1, &o_h1};
    commit(second, &zk);
    absorb(fs, second);
    Fr beta = sample_outside(fs);

    /* ---- third round */
    Fr v_h_beta = dh.vanishing(beta);
    Fr vv = mul(v_h_alpha, v_h_beta);
    Fr ea = mul(eta_a, vv), eb = mul(eta_b, vv), ec = mul(eta_c, vv);
    size_t na = std::min(ip[2].c.size(), std::min(ip[3].c.size(), ip[4].c.size()));
    Poly a_poly(na);
    #pragma omp parallel for num_threads(threads)
    for (size_t i = 0; i < na; i++) a_poly[i] = add(add(mul(ea, ip[2].c[i]), mul(eb, ip[3].c[i])), mul(ec, ip[4].c[i]));
    strip(a_poly);
    Fr ab = mul(alpha, beta);
    Poly b_poly(K), f_ev(K);
    #pragma omp parallel for num_threads(threads)
    for (size_t i = 0; i < K; i++) {
      b_poly[i] = add(sub(sub(ab, mul(alpha, ev_row[i])), mul(beta, ev_col[i])), ev_rc[i]);
      f_ev[i] = mul(sub(beta, ev_row[i]), sub(alpha, ev_col[i]));
    }
    fft(dk, b_poly, true); strip(b_poly);
    batch_inversion(f_ev);
    #pragma omp parallel for num_threads(threads)
    for (size_t i = 0; i < K; i++) f_ev[i] = mul(f_ev[i], add(add(mul(ea, ev_a[i]), mul(eb, ev_b[i])), mul(ec, ev_c[i])));
    Poly f_poly(f_ev);
    fft(dk, f_poly, true); strip(f_poly);
    Labeled o_g2, o_h2;
    {
      Poly bf = poly_mul(b_poly, f_poly);
      Poly diff(a_poly);
      if (diff.size() < bf.size()) diff.resize(bf.size(), zero());
      for (size_t i = 0; i < bf.size(); i++) diff[i] = sub(diff[i], bf[i]);
      strip(diff);
      Poly h2, rem;
      divide_by_vanishing(diff, K, h2, rem);
      o_h2.c = h2;
      if (!f_poly.empty()) o_g2.c.assign(f_poly.begin() + 1, f_poly.end());
      strip(o_g2.c);
    }
    o_g2.label = "g_2"; o_g2.bound = (long)K - 2; o_h2.label = "h_2";
    std::vector<Labeled*> third = {&o_g2, &o_h2};
    commit(third, &zk);
    absorb(fs, third);
    Fr gamma_ch = field_rand(fs);

    /* ---- evaluations and linear combinations [reference lib.rs:264-289, mod.rs:110-221] */
    Fr g1_b = evaluate(o_g1.c, beta), g2_g = evaluate(o_g2.c, gamma_ch), t_b = evaluate(o_t.c, beta), zb_b = evaluate(o_zb.c, beta);
    Fr evals[4] = {g1_b, g2_g, t_b, zb_b};
    { std::vector<uint8_t> eb2; for (auto& e : evals) put_fr(eb2, e); fs.absorb(eb2); }
    Fr xi;
    { u64 lo = fs.next_u64(), hi = fs.next_u64(); Fr c = {{lo, hi, 0, 0}}; f4_to_mont(&xi, &c, &F()); }
    Fr r_alpha_at_beta = mul(sub(v_h_alpha, v_h_beta), inv(sub(alpha, beta)));
    Fr v_x_beta = sub(pow_u64(beta, X), one());
    Fr c_za = mul(r_alpha_at_beta, add(eta_a, mul(eta_c, zb_b)));
    Fr c_w = neg(mul(t_b, v_x_beta));
    Fr c_h1 = neg(v_h_beta);
    Fr v_k_gamma = sub(pow_u64(gamma_ch, K), one());
    Fr bscale = add(mul(gamma_ch, g2_g), mul(t_b, inv(from_u64(K))));
    /* LC polynomials and randomness (constant terms do not enter the committed LC) */
    Labeled outer, inner;
    outer.label = "outer_sumcheck"; inner.label = "inner_sumcheck";
    axpy(outer.c, one(), o_mask.c); axpy(outer.c, c_za, o_za.c); axpy(outer.c, c_w, o_w.c); axpy(outer.c, c_h1, o_h1.c); strip(outer.c);
    axpy(outer.rand, c_za, o_za.rand); axpy(outer.rand, c_w, o_w.rand); strip(outer.rand);
    axpy(inner.c, ea, ip[2].c); axpy(inner.c, eb, ip[3].c); axpy(inner.c, ec, ip[4].c);
    axpy(inner.c, mul(bscale, alpha), ip[0].c); axpy(inner.c, mul(bscale, beta), ip[1].c); axpy(inner.c, neg(bscale), ip[5].c);
    axpy(inner.c, neg(v_k_gamma), o_h2.c); strip(inner.c);
    /* open: "beta" = {g_1, outer_sumcheck, t, z_b}, "gamma" = {g_2, inner_sumcheck} (BTree orders) */
    std::vector<Labeled*> at_beta = {&o_g1, &outer, &o_t, &o_zb}, at_gamma = {&o_g2, &inner};
    size_t max_bound = std::max(H, K) - 2;
    Opening ob = open_at_point(at_beta, beta, xi, max_bound), og = open_at_point(at_gamma, gamma_ch, xi, max_bound);

    /* ---- Proof: CanonicalSerialize */
    proof.clear();
    put_u64(proof, 3);
    std::vector<Labeled*>* rounds[3] = {&first, &second, &third};
    for (auto* rd : rounds) {
      put_u64(proof, rd->size());
      for (Labeled* p : *rd) {
        put_compressed(proof, p->comm);
        if (!sonic) { if (p->bound >= 0) { proof.push_back(1); put_compressed(proof, p->shifted); } else proof.push_back(0); }
      }
    }
    put_u64(proof, 4);
    for (auto& e : evals) put_fr(proof, e);
    put_u64(proof, 3); proof.push_back(0); proof.push_back(0); proof.push_back(0);
    put_u64(proof, 2);
    Opening* os[2] = {&ob, &og};
    for (Opening* o : os) {
      put_compressed(proof, o->w);
      if (o->has_rv) { proof.push_back(1); put_fr(proof, o->rv); } else proof.push_back(0);
    }
    proof.push_back(0);
    return 0;
  }
};

struct Handle {
  int curve;
  Marlin<Bls>* bls;
  Marlin<Bn>* bn;
};

template <class C>
Marlin<C>* make(int pc, int threads, const u64* powers, size_t n_g, const u64* gamma, const u64* gidx, size_t n_gamma, size_t nc, size_t nv,
                size_t ni, const u64* rp[3], const u64* cl[3], const u64* cf[3], int* rc) {
  Marlin<C>* m = new Marlin<C>();
  m->threads = threads; m->sonic = pc == 1;
  m->powers.resize(n_g); memcpy(m->powers.data(), powers, n_g * sizeof(typename C::Aff));
  m->gamma.resize(n_gamma); if (n_gamma) memcpy(m->gamma.data(), gamma, n_gamma * sizeof(typename C::Aff));
  for (size_t k = 0; k < n_gamma; k++) m->gamma_idx.push_back(gidx ? gidx[k] : k);
  m->D = n_g - 1; m->nc = nc; m->nv = nv; m->ni = ni;
  *rc = m->build_index(rp, cl, cf);
  return m;
}

}  // namespace

extern "C" {

void cport_set_skip_index_commit(int on) { g_skip_index_commit = on; }

/* Same inputs as b2m_srs_create + b2m_index_create (include/b2m.h).  Returns 0 or the b2m error code. */
int cport_index_create(int curve, int pc, int threads, const u64* powers, size_t n_g, const u64* gamma, const u64* gidx, size_t n_gamma,
                       size_t nc, size_t nv, size_t ni, const u64* a_rp, const u64* a_cl, const u64* a_cf, const u64* b_rp,
                       const u64* b_cl, const u64* b_cf, const u64* c_rp, const u64* c_cl, const u64* c_cf, void** out) {
  init_all();
  if (threads <= 0) threads = max_threads();
  const u64* rp[3] = {a_rp, b_rp, c_rp};
  const u64* cl[3] = {a_cl, b_cl, c_cl};
  const u64* cf[3] = {a_cf, b_cf, c_cf};
  Handle* h = new Handle();
  h->curve = curve; h->bls = nullptr; h->bn = nullptr;
  int rc = 0;
  if (curve == 0) h->bls = make<Bls>(pc, threads, powers, n_g, gamma, gidx, n_gamma, nc, nv, ni, rp, cl, cf, &rc);
  else h->bn = make<Bn>(pc, threads, powers, n_g, gamma, gidx, n_gamma, nc, nv, ni, rp, cl, cf, &rc);
  *out = h;
  return rc;
}
void cport_index_free(void* hv) { Handle* h = (Handle*)hv; if (!h) return; delete h->bls; delete h->bn; delete h; }
size_t cport_index_vk_bytes(void* hv, uint8_t* out, size_t cap) {
  Handle* h = (Handle*)hv;
  const std::vector<uint8_t>& v = h->curve == 0 ? h->bls->vk_bytes : h->bn->vk_bytes;
  if (out && cap >= v.size()) memcpy(out, v.data(), v.size());
  return v.size();
}
/* Same inputs as b2m_prove; *word_pos is updated; *seconds = wall time of the prove call. */
int cport_prove(void* hv, const u64* input, size_t n_input, const u64* witness, size_t n_witness, int rng_kind, const uint8_t* key,
                u64* word_pos, uint8_t* proof, size_t cap, size_t* len, double* seconds) {
  Handle* h = (Handle*)hv;
  ChaCha zk(key, rng_kind, *word_pos);
  std::vector<uint8_t> bytes;
  double t0 = now();
  int rc = h->curve == 0 ? h->bls->prove(input, n_input, witness, n_witness, zk, bytes) : h->bn->prove(input, n_input, witness, n_witness, zk, bytes);
  if (seconds) *seconds = now() - t0;
  if (rc) return rc;
  *word_pos = zk.pos;
  *len = bytes.size();
  if (cap < bytes.size()) return 1;
  memcpy(proof, bytes.data(), bytes.size());
  return 0;
}
}
