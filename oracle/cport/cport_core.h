#pragma once
/* oracle/cport core (shared by cport.c and prover.cpp) -- C restatement (OpenMP) of the reference's CPU algorithms for the prover hot path:
 * ark-ec `VariableBaseMSM::multi_scalar_mul` and ark-poly radix-2 FFT, over ark-ff style Montgomery
 * fields.  TEST INFRASTRUCTURE: used by tests/ as a fast checker at sizes the Python oracle cannot reach
 * and by bench.py as the CPU baseline ("kind": "port").  The reference itself (Rust, un-vendored crates)
 * cannot be built in this environment; parity of this port is pinned against the Python oracle (tests). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
typedef uint64_t u64;
typedef unsigned __int128 u128;

#define FP f4
#define NL 4
#include "fp_impl.h"
#undef FP
#undef NL
#define FP f6
#define NL 6
#include "fp_impl.h"
#undef FP
#undef NL

#define FQ f6
#define G bls
#include "g1_impl.h"
#undef FQ
#undef G
#define FQ f4
#define G bn
#include "g1_impl.h"
#undef FQ
#undef G

/* ---- field contexts (moduli: SURVEY.md App. C); R, R^2, -p^-1 derived at init --------------------------- */
static f4_ctx BLS_FR, BN_FR, BN_FQ;
static f6_ctx BLS_FQ;
static int inited = 0;
static const u64 BLS_FR_P[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
static const u64 BLS_FQ_P[6] = {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull, 0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};
static const u64 BN_FR_P[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const u64 BN_FQ_P[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const u64 BLS_FR_GEN = 7, BN_FR_GEN = 5;
static const int BLS_FR_S = 32, BN_FR_S = 28;

static u64 neg_inv64(u64 p0) { u64 x = 1; for (int i = 0; i < 6; i++) x *= 2 - p0 * x; return (u64)0 - x; }
/* r = 2^k mod p by repeated doubling */
static void pow2_mod(u64* r, const u64* p, int nl, int k) {
  memset(r, 0, 8 * nl); r[0] = 1;
  for (int i = 0; i < k; i++) {
    u64 cy = 0;
    for (int j = 0; j < nl; j++) { u64 n = (r[j] << 1) | cy; cy = r[j] >> 63; r[j] = n; }
    int ge = cy != 0;
    if (!ge) { ge = 1; for (int j = nl - 1; j >= 0; j--) if (r[j] != p[j]) { ge = r[j] > p[j]; break; } }
    if (ge) { u128 br = 0; for (int j = 0; j < nl; j++) { u128 t = (u128)r[j] - p[j] - (u64)br; r[j] = (u64)t; br = (t >> 64) & 1; } }
  }
}
static void init_all(void) {
  if (inited) return;
  memcpy(BLS_FR.p, BLS_FR_P, 32); BLS_FR.inv = neg_inv64(BLS_FR_P[0]); pow2_mod(BLS_FR.r, BLS_FR_P, 4, 256); pow2_mod(BLS_FR.r2, BLS_FR_P, 4, 512);
  memcpy(BN_FR.p, BN_FR_P, 32); BN_FR.inv = neg_inv64(BN_FR_P[0]); pow2_mod(BN_FR.r, BN_FR_P, 4, 256); pow2_mod(BN_FR.r2, BN_FR_P, 4, 512);
  memcpy(BN_FQ.p, BN_FQ_P, 32); BN_FQ.inv = neg_inv64(BN_FQ_P[0]); pow2_mod(BN_FQ.r, BN_FQ_P, 4, 256); pow2_mod(BN_FQ.r2, BN_FQ_P, 4, 512);
  memcpy(BLS_FQ.p, BLS_FQ_P, 48); BLS_FQ.inv = neg_inv64(BLS_FQ_P[0]); pow2_mod(BLS_FQ.r, BLS_FQ_P, 6, 384); pow2_mod(BLS_FQ.r2, BLS_FQ_P, 6, 768);
  inited = 1;
}
static double now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
/* Processors this process may run on -- NOT omp_get_max_threads(): launchers such as torchrun export OMP_NUM_THREADS=1,
 * which would silently turn the baseline single-threaded.  Every parallel region below names its team size explicitly. */
static int max_threads(void) {
#ifdef _OPENMP
  omp_set_dynamic(0);
  return omp_get_num_procs();
#else
  return 1;
#endif
}

/* ---- radix-2 FFT over Fr (natural order in and out; inverse scales by 1/n) ------------------------------ */
static void fr_root(f4_t* w, const f4_ctx* c, u64 gen, int two_adicity, int log_n) {
  /* TWO_ADIC_ROOT = gen^((p-1)/2^s); w_n = root^(2^(s - log_n)) */
  u64 e[4]; memcpy(e, c->p, 32); e[0] -= 1;
  for (int k = 0; k < two_adicity; k++) { for (int j = 0; j < 4; j++) { e[j] = (e[j] >> 1) | (j < 3 ? e[j + 1] << 63 : 0); } }
  f4_t g = {{gen, 0, 0, 0}}; f4_to_mont(&g, &g, c);
  f4_pow(w, &g, e, 4, c);
  for (int k = 0; k < two_adicity - log_n; k++) f4_sqr(w, w, c);
}
static void fr_fft(f4_t* a, int log_n, int inverse, const f4_ctx* c, u64 gen, int s, int threads) {
  size_t n = (size_t)1 << log_n;
  if (log_n == 0) return;
  f4_t w; fr_root(&w, c, gen, s, log_n);
  if (inverse) f4_inv(&w, &w, c);
  /* twiddle table w^j, j < n/2, built in parallel chunks */
  f4_t* tw = (f4_t*)malloc(sizeof(f4_t) * (n / 2 ? n / 2 : 1));
  size_t half = n / 2;
  #pragma omp parallel num_threads(threads)
  {
    int nt = 1, id = 0;
#ifdef _OPENMP
    nt = omp_get_num_threads(); id = omp_get_thread_num();
#endif
    size_t lo = half * id / nt, hi = half * (id + 1) / nt;
    if (lo < hi) {
      u64 e[1] = {lo}; f4_t cur; f4_pow(&cur, &w, e, 1, c);
      for (size_t j = lo; j < hi; j++) { tw[j] = cur; f4_mul(&cur, &cur, &w, c); }
    }
  }
  /* bit reversal */
  for (size_t i = 0; i < n; i++) {
    size_t r = 0; for (int b = 0; b < log_n; b++) r |= ((i >> b) & 1) << (log_n - 1 - b);
    if (i < r) { f4_t t = a[i]; a[i] = a[r]; a[r] = t; }
  }
  for (int st = 1; st <= log_n; st++) {
    size_t len = (size_t)1 << st, hl = len / 2, step = n / len;
    #pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t k = 0; k < n / 2; k++) {
      size_t blk = k / hl, j = k % hl, i0 = blk * len + j, i1 = i0 + hl;
      f4_t v; f4_mul(&v, &a[i1], &tw[j * step], c);
      f4_t u = a[i0];
      f4_add(&a[i0], &u, &v, c); f4_sub(&a[i1], &u, &v, c);
    }
  }
  if (inverse) {
    f4_t ninv = {{(u64)n, 0, 0, 0}}; f4_to_mont(&ninv, &ninv, c); f4_inv(&ninv, &ninv, c);
    #pragma omp parallel for num_threads(threads)
    for (size_t i = 0; i < n; i++) f4_mul(&a[i], &a[i], &ninv, c);
  }
  free(tw);
}
