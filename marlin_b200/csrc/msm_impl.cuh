// Kernels and member definitions of Msm<Fr, Fq>; included only by the inst_msm_*.cu units.
#pragma once
#include "msm.cuh"
#include "devmem.cuh"
#include "scan.cuh"

namespace b2m {

template <class Fq>
__device__ __forceinline__ Affine<Fq> ld_affine(const Affine<Fq>* p) {
  Affine<Fq> r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < 2 * Fq::N / 4; i++) {
    uint4 v = __ldg(q + i);
    d[4 * i] = v.x; d[4 * i + 1] = v.y; d[4 * i + 2] = v.z; d[4 * i + 3] = v.w;
  }
  return r;
}
// Out-of-line group operations for everything except the accumulate kernel: keeps the cold kernels
// small (and the build fast); the hot loop in msm_accumulate_kernel stays fully inlined.
template <class Fq> __device__ __noinline__ void g1_add(XYZZ<Fq>& a, const XYZZ<Fq>& b) { a.add(b); }
template <class Fq> __device__ __noinline__ void g1_add_mixed(XYZZ<Fq>& a, const Affine<Fq>& b) { a.add_mixed(b); }
template <class Fq> __device__ __noinline__ void g1_dbl(XYZZ<Fq>& a) { a = a.dbl(); }
template <class Fq> __device__ __noinline__ Fq fq_inverse(const Fq& a) { return a.inverse(); }
template <class Fq> __device__ __noinline__ Affine<Fq> g1_to_affine(const XYZZ<Fq>& p) {
  if (p.is_inf()) return Affine<Fq>::inf();
  Fq izzz = fq_inverse(p.ZZZ);
  Fq izz = (p.ZZ * izzz).sqr();
  return Affine<Fq>{p.X * izz, p.Y * izzz};
}
template <class Fq> __device__ __noinline__ XYZZ<Fq> g1_scalar_mul(const Affine<Fq>& p, const uint32_t* k, int nlimbs) {
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
  bool started = false;
  for (int i = nlimbs - 1; i >= 0; i--) {
    for (int b = 31; b >= 0; b--) {
      if (started) g1_dbl(acc);
      if ((k[i] >> b) & 1u) {
        g1_add_mixed(acc, p);
        started = true;
      }
    }
  }
  return acc;
}

// ---- key-load time: window tables ---------------------------------------------------------
// tables[w * n + i] = 2^(c*w) * P_i  (affine).  One thread per power; window w from window w-1.
template <class Fq>
__global__ void msm_precompute_kernel(Affine<Fq>* tables, size_t n, int c, int W) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<Fq> cur = ld_affine(tables + i);
  for (int w = 1; w < W; w++) {
    XYZZ<Fq> acc = XYZZ<Fq>::from_affine(cur);
    for (int k = 0; k < c; k++) g1_dbl(acc);
    cur = g1_to_affine(acc);
    st_words(tables + (size_t)w * n + i, cur);
  }
}

// ---- 1. digits ----------------------------------------------------------------------------
// digits[w * n + i] = (|d| - 1) | sign << 31, or MSM_NO_DIGIT for d == 0; hist[|d| - 1]++.
template <class Fr>
__global__ void msm_digits_kernel(const Fr* scalars, bool MONT, size_t n, int c, int W, uint32_t* digits, uint32_t* hist) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr s = ld_fr(scalars + i);
  if (MONT) s = s.to_canonical();
  const uint32_t half = 1u << (c - 1);
  uint32_t carry = 0;
  for (int w = 0; w < W; w++) {
    int bit = w * c;
    int limb = bit >> 5, off = bit & 31;
    uint32_t raw = 0;
    if (limb < Fr::N) {
      raw = s.l[limb] >> off;
      if (off + c > 32 && limb + 1 < Fr::N) raw |= s.l[limb + 1] << (32 - off);
      raw &= (1u << c) - 1;
    }
    uint32_t v = raw + carry;
    uint32_t out;
    if (v > half) {
      uint32_t mag = (1u << c) - v;  // d = v - 2^c < 0
      carry = 1;
      out = (mag - 1) | 0x80000000u;
    } else {
      carry = 0;
      out = v ? (v - 1) : MSM_NO_DIGIT;
    }
    digits[(size_t)w * n + i] = out;
    if (out != MSM_NO_DIGIT) atomicAdd(hist + (out & 0x7fffffffu), 1u);
  }
}

// ---- 2. exclusive scan of u32: scan.cuh -------------------------------------------------------

// ---- 3. scatter -------------------------------------------------------------------------------
// sorted[cursor[bucket]++] = i | w << 26 | sign << 31
static __global__ void msm_scatter_kernel(const uint32_t* digits, size_t n, int W, uint32_t* cursor, uint32_t* sorted) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int w = 0; w < W; w++) {
    uint32_t d = digits[(size_t)w * n + i];
    if (d == MSM_NO_DIGIT) continue;
    uint32_t pos = atomicAdd(cursor + (d & 0x7fffffffu), 1u);
    sorted[pos] = (uint32_t)i | ((uint32_t)w << MSM_IDX_BITS) | (d & 0x80000000u);
  }
}

// ---- 4. accumulate ----------------------------------------------------------------------------
// buckets[b] = sum of the referenced table points; one thread per bucket.  `ends` is the scatter
// cursor after step 3 (cursor[b] == end of bucket b).  The next point is prefetched while the
// current one is being added.
template <class Fq>
__global__ void __launch_bounds__(128)
msm_accumulate_kernel(const Affine<Fq>* __restrict__ tables, size_t table_stride, size_t base_off,
                      const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ ends,
                      const uint32_t* __restrict__ sorted, uint32_t num_buckets, XYZZ<Fq>* __restrict__ buckets) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= num_buckets) return;
  uint32_t start = offsets[b];
  uint32_t end = ends[b];
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
  if (start < end) {
    uint32_t ref = sorted[start];
    Affine<Fq> p = ld_affine(tables + (size_t)((ref >> MSM_IDX_BITS) & 31u) * table_stride + base_off + (ref & ((1u << MSM_IDX_BITS) - 1)));
    for (uint32_t e = start; e < end; e++) {
      Affine<Fq> cur = p;
      bool neg = ref >> 31;
      if (e + 1 < end) {
        ref = sorted[e + 1];
        p = ld_affine(tables + (size_t)((ref >> MSM_IDX_BITS) & 31u) * table_stride + base_off + (ref & ((1u << MSM_IDX_BITS) - 1)));
      }
      acc.add_mixed(cur, neg);
    }
  }
  st_words(buckets + b, acc);
}

// ---- 5. reduce ----------------------------------------------------------------------------------
// Level kernel: thread j owns in[j*L .. j*L+L):  S[j] = sum_i in[..+i],  T0 = sum_i i * in[..+i];
// block_t[blockIdx] = sum over the block's threads of T0.   W0(in) = L * W0(S) + sum_j T0_j.
template <class Fq>
__global__ void __launch_bounds__(MSM_RED_THREADS)
msm_seg_reduce_kernel(const XYZZ<Fq>* in, uint32_t m, uint32_t L, XYZZ<Fq>* S, XYZZ<Fq>* block_t) {
  __shared__ uint4 sm_raw[MSM_RED_THREADS * sizeof(XYZZ<Fq>) / 16];
  XYZZ<Fq>* sm = reinterpret_cast<XYZZ<Fq>*>(sm_raw);
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t nseg = m / L;
  XYZZ<Fq> running = XYZZ<Fq>::inf(), acc = XYZZ<Fq>::inf();
  if (j < nseg) {
    const XYZZ<Fq>* seg = in + (size_t)j * L;
    for (uint32_t i = L - 1; i >= 1; i--) {
      g1_add(running, ld_words(seg + i));
      g1_add(acc, running);
    }
    g1_add(running, ld_words(seg));
    st_words(S + j, running);
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = MSM_RED_THREADS / 2; s >= 1; s >>= 1) {
    if ((int)threadIdx.x < s) {
      XYZZ<Fq> a = sm[threadIdx.x];
      g1_add(a, sm[threadIdx.x + s]);
      sm[threadIdx.x] = a;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_words(block_t + blockIdx.x, sm[0]);
}

// result = top_sum + U_0 + L_0 * (U_1 + L_1 * (U_2 + ...)),  U_l = sum of level l's partials.
// extra[0..n_extra) are further XYZZ terms added in (hiding commitments, partial results).
template <class Fq>
__global__ void __launch_bounds__(256)
msm_finish_kernel(MsmLevels levels, const XYZZ<Fq>* top_sum, const XYZZ<Fq>* extra, int n_extra, XYZZ<Fq>* out_xyzz,
                  Affine<Fq>* out_affine) {
  __shared__ uint4 sm_raw[256 * sizeof(XYZZ<Fq>) / 16];
  XYZZ<Fq>* sm = reinterpret_cast<XYZZ<Fq>*>(sm_raw);
  XYZZ<Fq> total = XYZZ<Fq>::inf();
  for (int lv = levels.n - 1; lv >= 0; lv--) {
    const XYZZ<Fq>* part = reinterpret_cast<const XYZZ<Fq>*>(levels.lv[lv].partials);
    XYZZ<Fq> a = XYZZ<Fq>::inf();
    for (uint32_t i = threadIdx.x; i < levels.lv[lv].count; i += 256) g1_add(a, ld_words(part + i));
    sm[threadIdx.x] = a;
    __syncthreads();
    for (int s = 128; s >= 1; s >>= 1) {
      if ((int)threadIdx.x < s) {
        XYZZ<Fq> t = sm[threadIdx.x];
        g1_add(t, sm[threadIdx.x + s]);
        sm[threadIdx.x] = t;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      // total = U_lv + L_lv * total
      for (uint32_t k = 0; k < levels.lv[lv].log_l; k++) g1_dbl(total);
      g1_add(total, sm[0]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (top_sum) g1_add(total, ld_words(top_sum));
    for (int i = 0; i < n_extra; i++) g1_add(total, ld_words(extra + i));
    if (out_xyzz) st_words(out_xyzz, total);
    if (out_affine) st_words(out_affine, g1_to_affine(total));
  }
}

// ---- small MSM: one thread per term, double-and-add, block tree (n <= 256) -------------------
template <class Fr, class Fq>
__global__ void __launch_bounds__(256)
msm_small_kernel(const Affine<Fq>* bases, const Fr* scalars, bool MONT, int n, XYZZ<Fq>* out) {
  __shared__ uint4 sm_raw[256 * sizeof(XYZZ<Fq>) / 16];
  XYZZ<Fq>* sm = reinterpret_cast<XYZZ<Fq>*>(sm_raw);
  XYZZ<Fq> a = XYZZ<Fq>::inf();
  for (int i = threadIdx.x; i < n; i += 256) {
    Fr s = ld_fr(scalars + i);
    if (MONT) s = s.to_canonical();
    XYZZ<Fq> t = g1_scalar_mul<Fq>(ld_affine(bases + i), s.l, Fr::N);
    g1_add(a, t);
  }
  sm[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s >= 1; s >>= 1) {
    if ((int)threadIdx.x < s) {
      XYZZ<Fq> t = sm[threadIdx.x];
      g1_add(t, sm[threadIdx.x + s]);
      sm[threadIdx.x] = t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_words(out, sm[0]);
}

// out[i] = beta^i * g  (test-SRS generation; the G1 half of KZG10::setup)
template <class Fr, class Fq>
__global__ void g1_powers_kernel(Affine<Fq> g, Fr beta, size_t n, Affine<Fq>* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr k = beta.pow_u64(i).to_canonical();
  st_words(out + i, g1_to_affine(g1_scalar_mul<Fq>(g, k.l, Fr::N)));
}

// ---- host driver ------------------------------------------------------------------------------
template <class Fr, class Fq>
int Msm<Fr, Fq>::pick_window(size_t n) {
  int lg = 0;
  while (((size_t)1 << (lg + 1)) <= n) lg++;
  int c = lg - 2;
  if (c < MSM_MIN_WINDOW) c = MSM_MIN_WINDOW;
  if (c > 20) c = 20;
  return c;
}

template <class Fr, class Fq>
Msm<Fr, Fq>::Msm(Ctx& cx, const Affine<Fq>* host_powers, size_t n, int window_bits) : ctx(&cx), n_srs(n) {
  B2M_REQUIRE(n >= 1 && n <= ((size_t)1 << MSM_IDX_BITS), B2M_ERR_INVALID_ARG, "SRS size %zu out of range", n);
  c = window_bits > 0 ? window_bits : pick_window(n);
  B2M_REQUIRE(c >= MSM_MIN_WINDOW && c <= 24, B2M_ERR_INVALID_ARG, "window bits %d out of range [%d, 24]", c, MSM_MIN_WINDOW);
  W = (Fr::Params::BITS + 1 + c - 1) / c;
  B2M_REQUIRE(W <= 32, B2M_ERR_INVALID_ARG, "too many windows (%d)", W);
  tables = DBuf<Affine<Fq>>(cx, (size_t)W * n);
  tables.upload(host_powers, n);
  msm_precompute_kernel<Fq><<<div_up(n, 128), 128, 0, cx.stream>>>(tables.p, n, c, W);
  B2M_CHECK_LAUNCH();
  cx.launches++;
  cx.sync();
}

template <class Fr, class Fq>
void Msm<Fr, Fq>::run(const Fr* scalars, bool mont, size_t n, size_t base_off, const XYZZ<Fq>* extra, int n_extra,
                      XYZZ<Fq>* out_xyzz, Affine<Fq>* out_affine) {
  B2M_REQUIRE(base_off + n <= n_srs, B2M_ERR_DEGREE_TOO_LARGE, "MSM slice [%zu, %zu) exceeds the SRS (%zu powers)", base_off,
              base_off + n, n_srs);
  Ctx& cx = *ctx;
  MsmLevels levels;
  levels.n = 0;
  if (n == 0) {
    msm_finish_kernel<Fq><<<1, 256, 0, cx.stream>>>(levels, nullptr, extra, n_extra, out_xyzz, out_affine);
    B2M_CHECK_LAUNCH();
    cx.launches++;
    return;
  }
  const uint32_t B = 1u << (c - 1);
  DBuf<uint32_t> digits(cx, (size_t)W * n), hist(cx, B), offsets(cx, B), cursor(cx, B), sorted(cx, (size_t)W * n);
  hist.zero();
  size_t sp0 = cx.span_begin("msm_sort", (double)n);
  msm_digits_kernel<Fr><<<div_up(n, 256), 256, 0, cx.stream>>>(scalars, mont, n, c, W, digits.p, hist.p);
  B2M_CHECK_LAUNCH();
  cx.launches++;
  exclusive_scan_u32(cx, hist.p, offsets.p, B);
  B2M_CUDA(cudaMemcpyAsync(cursor.p, offsets.p, B * sizeof(uint32_t), cudaMemcpyDeviceToDevice, cx.stream));
  msm_scatter_kernel<<<div_up(n, 256), 256, 0, cx.stream>>>(digits.p, n, W, cursor.p, sorted.p);
  B2M_CHECK_LAUNCH();
  cx.launches++;
  cx.span_end(sp0);
  DBuf<XYZZ<Fq>> buckets(cx, B);
  size_t sp = cx.span_begin("msm_accumulate_kernel", (double)n);
  msm_accumulate_kernel<Fq><<<div_up(B, 128), 128, 0, cx.stream>>>(tables.p, n_srs, base_off, offsets.p, cursor.p, sorted.p, B,
                                                                    buckets.p);
  B2M_CHECK_LAUNCH();
  cx.launches++;
  cx.span_end(sp);
  size_t sp2 = cx.span_begin("msm_reduce", (double)n);

  // hierarchical reduction: W0(A) = L * W0(S) + sum T0, level by level
  std::vector<DBuf<XYZZ<Fq>>> keep;
  const XYZZ<Fq>* cur = buckets.p;
  uint32_t m = B;
  while (m > 1) {
    uint32_t L = m >= (uint32_t)MSM_SEG ? (uint32_t)MSM_SEG : m;
    uint32_t nseg = m / L;
    uint32_t grid = (nseg + MSM_RED_THREADS - 1) / MSM_RED_THREADS;
    DBuf<XYZZ<Fq>> S(cx, nseg), part(cx, grid);
    msm_seg_reduce_kernel<Fq><<<grid, MSM_RED_THREADS, 0, cx.stream>>>(cur, m, L, S.p, part.p);
    B2M_CHECK_LAUNCH();
    cx.launches++;
    B2M_REQUIRE(levels.n < MSM_MAX_LEVELS, B2M_ERR_INVALID_ARG, "too many reduction levels");
    int lg = 0;
    while ((1u << lg) < L) lg++;
    levels.lv[levels.n++] = MsmLevel{part.p, grid, (uint32_t)lg};
    cur = S.p;
    m = nseg;
    keep.push_back(std::move(S));
    keep.push_back(std::move(part));
  }
  msm_finish_kernel<Fq><<<1, 256, 0, cx.stream>>>(levels, cur, extra, n_extra, out_xyzz, out_affine);
  B2M_CHECK_LAUNCH();
  cx.launches++;
  cx.span_end(sp2);
  // the DBufs are stream-ordered: their frees are enqueued behind the kernels above
}

template <class Fr, class Fq>
void Msm<Fr, Fq>::run_small(const Affine<Fq>* bases, const Fr* scalars, bool mont, int n, XYZZ<Fq>* out_xyzz) {
  B2M_REQUIRE(n >= 0 && n <= 4096, B2M_ERR_INVALID_ARG, "run_small: n = %d", n);
  msm_small_kernel<Fr, Fq><<<1, 256, 0, ctx->stream>>>(bases, scalars, mont, n, out_xyzz);
  B2M_CHECK_LAUNCH();
  ctx->launches++;
}

template <class Fr, class Fq>
void Msm<Fr, Fq>::run_host(size_t base_off, const uint64_t* scalars, size_t n, uint64_t* out_xy, int* out_is_inf) {
  Ctx& cx = *ctx;
  DBuf<Fr> sc(cx, n ? n : 1);
  if (n) sc.upload(reinterpret_cast<const Fr*>(scalars), n);
  DBuf<Affine<Fq>> res(cx, 1);
  run(sc.p, false, n, base_off, nullptr, 0, nullptr, res.p);
  Affine<Fq> h;
  res.download(&h, 1);
  memcpy(out_xy, &h, sizeof(h));
  if (out_is_inf) *out_is_inf = h.is_inf() ? 1 : 0;
}

template <class Fr, class Fq>
void Msm<Fr, Fq>::g1_powers_host(Ctx& cx, const uint64_t* g_xy, const uint64_t* beta, size_t n, uint64_t* out) {
  Affine<Fq> g;
  memcpy(&g, g_xy, sizeof(g));
  Fr b;
  memcpy(&b, beta, sizeof(b));
  b = Fr::from_canonical(b);
  DBuf<Affine<Fq>> d(cx, n ? n : 1);
  if (n) {
    g1_powers_kernel<Fr, Fq><<<div_up(n, 64), 64, 0, cx.stream>>>(g, b, n, d.p);
    B2M_CHECK_LAUNCH();
    cx.launches++;
    d.download(reinterpret_cast<Affine<Fq>*>(out), n);
  }
}

}  // namespace b2m
