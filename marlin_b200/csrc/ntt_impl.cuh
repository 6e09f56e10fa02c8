// Kernels and member definitions of Ntt<Fr>; included only by the inst_ntt_*.cu units.
#pragma once
#include "ntt.cuh"
#include "devmem.cuh"

namespace b2m {


// tw[j] = root_N^j
template <class Fr>
__global__ void ntt_table_kernel(Fr* tw, size_t half, Fr root) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= half) return;
  st_fr(tw + j, root.pow_u64(j));
}

// w_n^(+-e) for e in [0, n/2) expressed through the big table; `sh` = max_log - log_n.
template <class Fr>
__device__ __forceinline__ Fr twiddle(const Fr* tw, size_t e_n, int sh, size_t halfN, bool inverse) {
  size_t e = e_n << sh;  // exponent in units of w_N, < N/2
  if (!inverse) return ldg_fr(tw + e);
  if (e == 0) return Fr::one();
  return ldg_fr(tw + (halfN - e)).neg();  // w^-e = -w^(N/2 - e)
}

// One pass: stages [s0, s0 + k) of a 2^log_n DIF transform.  q indexes the n / 2^k
// independent sub-transforms of this pass ("columns"); a tile is `cols` consecutive q.
// LAST: this pass ends at the final stage (lb == 0): store bit-reversed (and scaled if inverse).
template <class Fr, bool LAST>
__global__ void __launch_bounds__(NTT_THREADS)
ntt_pass_kernel(const Fr* src, Fr* dst, const Fr* __restrict__ tw, int log_n, int s0, int k,
                int cols_log, int max_log, bool inverse, Fr n_inv) {
  extern __shared__ uint4 smem_raw[];
  Fr* sm = reinterpret_cast<Fr*>(smem_raw);
  const int R = 1 << k;
  const int C = 1 << cols_log;
  const int lb = log_n - s0 - k;  // low bits below the row bits
  const size_t q0 = (size_t)blockIdx.x << cols_log;
  const int tid = threadIdx.x;
  const int sh = max_log - log_n;
  const size_t halfN = (size_t)1 << (max_log - 1);

  // smem index: row-major [r][cc] when columns are the contiguous dimension (lb > 0), else [cc][r].
  auto sidx = [&](int r, int cc) -> int { return LAST ? (cc << k) + r : (r << cols_log) + cc; };
  auto gidx = [&](int r, int cc) -> size_t {
    size_t q = q0 + cc;
    size_t hi = q >> lb, lo = q & (((size_t)1 << lb) - 1);
    return (hi << (k + lb)) + ((size_t)r << lb) + lo;
  };

  // Tile load through the TMA engine (cp.async.bulk, 1-D): every contiguous run of the tile -- one 2^cols_log-element
  // row (256 B) in the strided passes, one 2^k-element column group in the last pass -- is one bulk copy that
  // completes on a single mbarrier; no thread spends registers or LSU slots on the staging.
  __shared__ __align__(8) unsigned long long tile_bar;
  const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&tile_bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  {
    const int runs = LAST ? C : R;                      // number of contiguous runs in the tile
    const uint32_t run_bytes = (uint32_t)sizeof(Fr) << (LAST ? k : cols_log);
    if (tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(run_bytes * (uint32_t)runs) : "memory");
    for (int run = tid; run < runs; run += NTT_THREADS) {
      const Fr* g = LAST ? src + gidx(0, run) : src + gidx(run, 0);
      const uint32_t dst_addr = (uint32_t)__cvta_generic_to_shared(sm + (LAST ? sidx(0, run) : sidx(run, 0)));
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_addr), "l"(g),
                   "r"(run_bytes), "r"(bar_addr)
                   : "memory");
    }
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
          : "=r"(done)
          : "r"(bar_addr)
          : "memory");
    }
  }

  for (int t = 0; t < k; t++) {
    const int stride_log = k - 1 - t;  // row distance of a butterfly pair = 2^stride_log
    for (int idx = tid; idx < (R / 2) * C; idx += NTT_THREADS) {
      int u, cc;
      if (LAST) { u = idx & (R / 2 - 1); cc = idx >> (k - 1); } else { cc = idx & (C - 1); u = idx >> cols_log; }
      int r_lo = u & ((1 << stride_log) - 1);
      int r0 = ((u >> stride_log) << (stride_log + 1)) + r_lo;
      int r1 = r0 + (1 << stride_log);
      size_t q = q0 + cc;
      size_t lo = q & (((size_t)1 << lb) - 1);
      size_t j = ((size_t)r_lo << lb) + lo;  // i mod half
      Fr w = twiddle(tw, j << (s0 + t), sh, halfN, inverse);
      Fr x = sm[sidx(r0, cc)], y = sm[sidx(r1, cc)];
      sm[sidx(r0, cc)] = x + y;
      sm[sidx(r1, cc)] = (x - y) * w;
    }
    __syncthreads();
  }

  for (int idx = tid; idx < R * C; idx += NTT_THREADS) {
    int r, cc;
    if (LAST) { r = idx & (R - 1); cc = idx >> k; } else { cc = idx & (C - 1); r = idx >> cols_log; }
    Fr v = sm[sidx(r, cc)];
    size_t g = gidx(r, cc);
    if (LAST) {
      size_t rev = __brevll((unsigned long long)g) >> (64 - log_n);
      if (inverse) v = v * n_inv;
      st_fr(dst + rev, v);
    } else {
      st_fr(dst + g, v);
    }
  }
}

// data[i] *= g^i  (coset shift), g^i by per-thread exponentiation of a 16-element run.
template <class Fr>
__global__ void ntt_coset_scale_kernel(Fr* data, size_t n, Fr g) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t i0 = t * 16;
  if (i0 >= n) return;
  Fr p = g.pow_u64(i0);
  for (int k = 0; k < 16 && i0 + k < n; k++) {
    st_fr(data + i0 + k, ld_fr(data + i0 + k) * p);
    p = p * g;
  }
}

template <class Fr>
Ntt<Fr>::Ntt(Ctx& c) : ctx(&c) {
  B2M_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<Fr, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  B2M_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<Fr, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
}

template <class Fr>
Fr Ntt<Fr>::root_of_unity(int log_n) {
  using P = typename Fr::Params;
  Fr r;
  for (int i = 0; i < Fr::N; i++) r.l[i] = P::root(i);
  for (int i = 0; i < P::TWO_ADICITY - log_n; i++) r = r.sqr();
  return r;
}

template <class Fr>
void Ntt<Fr>::ensure_table(int log_n) {
  if (log_n <= table.max_log) return;
  using P = typename Fr::Params;
  B2M_REQUIRE(log_n <= P::TWO_ADICITY, B2M_ERR_DEGREE_TOO_LARGE, "domain 2^%d exceeds the field's 2-adicity %d", log_n,
              P::TWO_ADICITY);
  int ml = log_n < 12 ? 12 : log_n;
  size_t half = (size_t)1 << (ml - 1);
  DBuf<Fr> nb(*ctx, half);
  ntt_table_kernel<Fr><<<div_up(half, 256), 256, 0, ctx->stream>>>(nb.p, half, root_of_unity(ml));
  B2M_CHECK_LAUNCH();
  ctx->launches++;
  tw_buf = std::move(nb);
  table.tw = tw_buf.p;
  table.max_log = ml;
}

template <class Fr>
void Ntt<Fr>::run(Fr* work, Fr* out, int log_n, bool inverse) {
  if (log_n == 0) {
    B2M_CUDA(cudaMemcpyAsync(out, work, sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
    return;
  }
  ensure_table(log_n);
  int passes = (log_n + NTT_MAX_K - 1) / NTT_MAX_K;
  int base = log_n / passes, extra = log_n % passes;
  Fr n_inv = Fr::from_u64((uint64_t)1 << log_n).inverse();
  int s0 = 0;
  size_t sp = ctx->span_begin("ntt", (double)((size_t)1 << log_n));
  for (int p = 0; p < passes; p++) {
    int k = base + (p < extra ? 1 : 0);
    bool last = (p == passes - 1);
    size_t ncols = (size_t)1 << (log_n - k);
    int cols_log = 3;
    while (((size_t)1 << cols_log) > ncols) cols_log--;
    size_t tiles = ncols >> cols_log;
    size_t smem = (sizeof(Fr) << (k + cols_log));
    if (last)
      ntt_pass_kernel<Fr, true><<<(unsigned)tiles, NTT_THREADS, smem, ctx->stream>>>(work, out, table.tw, log_n, s0, k, cols_log,
                                                                                   table.max_log, inverse, n_inv);
    else
      ntt_pass_kernel<Fr, false><<<(unsigned)tiles, NTT_THREADS, smem, ctx->stream>>>(work, work, table.tw, log_n, s0, k, cols_log,
                                                                                    table.max_log, inverse, n_inv);
    B2M_CHECK_LAUNCH();
    ctx->launches++;
    s0 += k;
  }
  ctx->span_end(sp);
}

template <class Fr>
void Ntt<Fr>::coset_scale(Fr* data, size_t n, const Fr& g) {
  ntt_coset_scale_kernel<Fr><<<div_up(div_up(n, 16), 128), 128, 0, ctx->stream>>>(data, n, g);
  B2M_CHECK_LAUNCH();
  ctx->launches++;
}

template <class Fr>
void Ntt<Fr>::run_host(uint64_t* data, unsigned log_n, bool inverse, bool coset) {
  using P = typename Fr::Params;
  B2M_REQUIRE((int)log_n <= P::TWO_ADICITY, B2M_ERR_DEGREE_TOO_LARGE, "2^%u exceeds the 2-adicity of the field", log_n);
  Ctx& cx = *ctx;
  size_t n = (size_t)1 << log_n;
  DBuf<Fr> work(cx, n), out(cx, n);
  work.upload(reinterpret_cast<const Fr*>(data), n);
  Fr g;
  for (int i = 0; i < Fr::N; i++) g.l[i] = P::gen(i);
  if (coset && !inverse) coset_scale(work.p, n, g);
  run(work.p, out.p, (int)log_n, inverse);
  if (coset && inverse) coset_scale(out.p, n, g.inverse());
  out.download(reinterpret_cast<Fr*>(data), n);
}

}  // namespace b2m
