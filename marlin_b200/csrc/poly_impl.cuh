// Device-side polynomial algebra of the AHP prover rounds (everything in reference
// src/ahp/prover.rs that is not an FFT or an MSM): fused pointwise maps, batch inversion,
// linear-recurrence scans (division by X - z and by X^s - 1, Horner evaluation, segmented sums),
// sparse matrix-vector products and on-device sampling of the mask polynomial.
// All field elements are Montgomery-form Fr in HBM; nothing here touches the host except scalars.
#pragma once
#include "common.cuh"
#include "devmem.cuh"
#include "ntt.cuh"

namespace b2m {

// ---- generic fused element-wise launcher ------------------------------------------------------
template <class F>
__global__ void ew_kernel(size_t n, F f) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) f(i);
}
template <class F>
void ew(Ctx& cx, size_t n, F f) {
  if (n == 0) return;
  ew_kernel<<<div_up(n, 256), 256, 0, cx.stream>>>(n, f);
  B2M_CHECK_LAUNCH();
  cx.launches++;
}

// w_n^i for any i in [0, n) through the NTT twiddle table (which holds w_N^j, j < N/2).
template <class Fr>
__device__ __forceinline__ Fr domain_element(const Fr* tw, int max_log, int log_n, size_t i) {
  if (log_n == 0) return Fr::one();
  size_t half = (size_t)1 << (log_n - 1);
  size_t e = (i & (half - 1)) << (max_log - log_n);
  Fr v = ldg_fr(tw + e);
  return (i & half) ? v.neg() : v;
}

// ---- linear recurrence:  out[j] = in[j] + z * out[j + s]  (out[j >= n] = 0) ---------------------
// Chunks of REC_M "super-elements" (s interleaved sequences).  Pass A computes each chunk's local
// Horner value, the recursion turns those into carries, pass C replays the chunk with its carry.
constexpr int REC_M = 32;

template <class Fr, bool MUL>
__global__ void rec_chunk_kernel(const Fr* in, size_t n, size_t s, Fr z, size_t nchunks, Fr* chunk_val) {
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (tid >= nchunks * s) return;
  size_t t = tid / s, r = tid % s;
  size_t k_lo = t * REC_M, k_hi = k_lo + REC_M;  // super-element range
  Fr acc = Fr::zero();
  for (size_t k = k_hi; k-- > k_lo;) {
    size_t j = k * s + r;
    Fr v = j < n ? ld_fr(in + j) : Fr::zero();
    acc = MUL ? v + z * acc : v + acc;
  }
  st_fr(chunk_val + tid, acc);
}

template <class Fr, bool MUL>
__global__ void rec_apply_kernel(const Fr* in, Fr* out, size_t n, size_t s, Fr z, size_t nchunks, const Fr* chunk_scan) {
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (tid >= nchunks * s) return;
  size_t t = tid / s, r = tid % s;
  size_t k_lo = t * REC_M, k_hi = k_lo + REC_M;
  // carry = recurrence value at the first super-element of the next chunk
  Fr acc = (chunk_scan && t + 1 < nchunks) ? ld_fr(chunk_scan + (t + 1) * s + r) : Fr::zero();
  for (size_t k = k_hi; k-- > k_lo;) {
    size_t j = k * s + r;
    if (j < n) {
      Fr v = ld_fr(in + j);
      acc = MUL ? v + z * acc : v + acc;
      st_fr(out + j, acc);
    }
  }
}

template <class Fr>
void rec_suffix(Ctx& cx, const Fr* in, Fr* out, size_t n, size_t s, const Fr& z, bool mul) {
  if (n == 0) return;
  size_t nsuper = (n + s - 1) / s;
  size_t nchunks = (nsuper + REC_M - 1) / REC_M;
  size_t threads = nchunks * s;
  if (nchunks == 1) {
    if (mul) rec_apply_kernel<Fr, true><<<div_up(threads, 128), 128, 0, cx.stream>>>(in, out, n, s, z, nchunks, nullptr);
    else rec_apply_kernel<Fr, false><<<div_up(threads, 128), 128, 0, cx.stream>>>(in, out, n, s, z, nchunks, nullptr);
    B2M_CHECK_LAUNCH();
    cx.launches++;
    return;
  }
  DBuf<Fr> cv(cx, threads);
  if (mul) rec_chunk_kernel<Fr, true><<<div_up(threads, 128), 128, 0, cx.stream>>>(in, n, s, z, nchunks, cv.p);
  else rec_chunk_kernel<Fr, false><<<div_up(threads, 128), 128, 0, cx.stream>>>(in, n, s, z, nchunks, cv.p);
  B2M_CHECK_LAUNCH();
  cx.launches++;
  Fr zm = mul ? z.pow_u64(REC_M) : z;
  rec_suffix(cx, cv.p, cv.p, threads, s, zm, mul);  // in place: chunk values -> chunk-level recurrence
  if (mul) rec_apply_kernel<Fr, true><<<div_up(threads, 128), 128, 0, cx.stream>>>(in, out, n, s, z, nchunks, cv.p);
  else rec_apply_kernel<Fr, false><<<div_up(threads, 128), 128, 0, cx.stream>>>(in, out, n, s, z, nchunks, cv.p);
  B2M_CHECK_LAUNCH();
  cx.launches++;
}

// ---- batch inversion (Montgomery's trick, zeros left untouched) --------------------------------
constexpr int BINV_M = 32;
template <class Fr>
__global__ void __launch_bounds__(128) batch_inverse_kernel(Fr* data, size_t n, size_t nthreads) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= nthreads) return;
  Fr pre[BINV_M];
  Fr acc = Fr::one();
#pragma unroll 1
  for (int k = 0; k < BINV_M; k++) {
    size_t j = t + (size_t)k * nthreads;
    pre[k] = acc;
    if (j < n) {
      Fr v = ld_fr(data + j);
      if (!v.is_zero()) acc = acc * v;
    }
  }
  Fr inv = acc.inverse();
#pragma unroll 1
  for (int k = BINV_M - 1; k >= 0; k--) {
    size_t j = t + (size_t)k * nthreads;
    if (j < n) {
      Fr v = ld_fr(data + j);
      if (!v.is_zero()) {
        st_fr(data + j, inv * pre[k]);
        inv = inv * v;
      }
    }
  }
}
template <class Fr>
void batch_inverse(Ctx& cx, Fr* data, size_t n) {
  if (n == 0) return;
  size_t nthreads = (n + BINV_M - 1) / BINV_M;
  batch_inverse_kernel<Fr><<<div_up(nthreads, 128), 128, 0, cx.stream>>>(data, n, nthreads);
  B2M_CHECK_LAUNCH();
  cx.launches++;
}

// ---- CSR sparse matrix - vector product:  out[r] = sum_e coeff[e] * z[col[e]] --------------------
template <class Fr>
__global__ void spmv_kernel(const uint32_t* row_ptr, const uint32_t* col, const Fr* coeff, const Fr* z, size_t nrows, Fr* out) {
  size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  Fr acc = Fr::zero();
  for (uint32_t e = row_ptr[r]; e < row_ptr[r + 1]; e++) acc = acc + ldg_fr(coeff + e) * ld_fr(z + col[e]);
  st_fr(out + r, acc);
}

// ---- mask polynomial: ChaCha keystream -> rejection-sampled Fr, on the device --------------------
__device__ __forceinline__ uint32_t d_rotl(uint32_t x, int n) { return __funnelshift_l(x, x, n); }
__device__ inline void chacha_block_dev(const uint32_t* key, uint64_t counter, int rounds, uint32_t* out) {
  uint32_t init[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3],
                       key[4], key[5], key[6], key[7], (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
  uint32_t s[16];
#pragma unroll
  for (int i = 0; i < 16; i++) s[i] = init[i];
#define B2M_QR(a, b, c, d)                                   \
  s[a] += s[b]; s[d] = d_rotl(s[d] ^ s[a], 16);              \
  s[c] += s[d]; s[b] = d_rotl(s[b] ^ s[c], 12);              \
  s[a] += s[b]; s[d] = d_rotl(s[d] ^ s[a], 8);               \
  s[c] += s[d]; s[b] = d_rotl(s[b] ^ s[c], 7);
  for (int r = 0; r < rounds / 2; r++) {
    B2M_QR(0, 4, 8, 12) B2M_QR(1, 5, 9, 13) B2M_QR(2, 6, 10, 14) B2M_QR(3, 7, 11, 15)
    B2M_QR(0, 5, 10, 15) B2M_QR(1, 6, 11, 12) B2M_QR(2, 7, 8, 13) B2M_QR(3, 4, 9, 14)
  }
#undef B2M_QR
#pragma unroll
  for (int i = 0; i < 16; i++) out[i] = s[i] + init[i];
}
struct ChaChaKey {
  uint32_t k[8];
};
// attempt a reads stream words [pos0 + 8a, pos0 + 8a + 8): one `F::rand` draw of 4 u64 limbs.
template <class Fr>
__global__ void sample_attempts_kernel(ChaChaKey key, int rounds, uint64_t pos0, size_t nattempts, Fr* cand, uint32_t* accept) {
  size_t a = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (a >= nattempts) return;
  uint64_t pos = pos0 + 8ull * a;
  uint32_t w[32];
  chacha_block_dev(key.k, pos >> 4, rounds, w);
  int off = (int)(pos & 15);
  if (off + Fr::N > 16) chacha_block_dev(key.k, (pos >> 4) + 1, rounds, w + 16);
  Fr v;
#pragma unroll
  for (int i = 0; i < Fr::N; i++) v.l[i] = w[off + i];
  constexpr int shave = 32 * Fr::N - Fr::Params::BITS;
  v.l[Fr::N - 1] &= 0xffffffffu >> shave;
  bool lt = false;
  for (int i = Fr::N - 1; i >= 0; i--) {
    uint32_t m = Fr::Params::mod(i);
    if (v.l[i] != m) { lt = v.l[i] < m; break; }
  }
  st_fr(cand + a, v);
  accept[a] = lt ? 1u : 0u;
}
// out[rank] = cand[a] for accepted attempts with rank < need; last_attempt = attempt holding rank need-1.
template <class Fr>
__global__ void sample_compact_kernel(const Fr* cand, const uint32_t* accept, const uint32_t* rank, size_t nattempts, size_t have,
                                      size_t need, Fr* out, unsigned long long* last_attempt) {
  size_t a = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (a >= nattempts || !accept[a]) return;
  size_t r = have + rank[a];
  if (r < need) {
    st_fr(out + r, ld_fr(cand + a));
    if (r == need - 1) *last_attempt = a;
  }
}

}  // namespace b2m
