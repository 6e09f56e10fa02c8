// Integer-ALU roofline microbenchmark: sustained Fq / Fr Montgomery multiplications per second
// and XYZZ mixed additions per second on the whole chip (no memory traffic).  The MSM and NTT
// kernels are bound by these rates, not by HBM (DESIGN.md "Rooflines").
#include <cstdio>
#include <cuda_runtime.h>
#include "../marlin_b200/csrc/curve.cuh"
using namespace b2m;

template <class F, int ILP>
__global__ void __launch_bounds__(256) mul_kernel(F* io, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  F a[ILP], b = io[t & 1023];
#pragma unroll
  for (int k = 0; k < ILP; k++) { a[k] = io[(t + k) & 1023]; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) a[k] = a[k] * b;
  }
  F s = a[0];
#pragma unroll
  for (int k = 1; k < ILP; k++) s = s + a[k];
  if (s.is_zero()) io[t & 1023] = s;
}
template <class F>
__global__ void __launch_bounds__(128) madd_kernel(Affine<F>* pts, XYZZ<F>* out, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  XYZZ<F> acc = XYZZ<F>::from_affine(pts[t & 255]);
  Affine<F> p = pts[(t + 7) & 255];
  for (int i = 0; i < iters; i++) acc.add_mixed(p, i & 1);
  if (acc.is_inf()) out[0] = acc;
}
template <class K, class... A>
float time_kernel(K k, dim3 g, dim3 b, A... args) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<g, b>>>(args...); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0); k<<<g, b>>>(args...); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}
int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount;
  void* buf; cudaMalloc(&buf, 1 << 20); cudaMemset(buf, 0x5a, 1 << 20);
  void* outb; cudaMalloc(&outb, 4096);
  printf("{\"device\": \"%s\", \"sms\": %d", prop.name, sms);
  const int iters = 2000;
  for (int bps = 1; bps <= 4; bps *= 2) {
    dim3 g(sms * bps), b(256);
    double n = (double)sms * bps * 256 * iters;
    float ms;
    ms = time_kernel(mul_kernel<FqBls, 1>, g, b, (FqBls*)buf, iters); printf(", \"fq_mul_ilp1_bps%d_Gps\": %.2f", bps, n * 1 / ms / 1e6);
    ms = time_kernel(mul_kernel<FqBls, 2>, g, b, (FqBls*)buf, iters); printf(", \"fq_mul_ilp2_bps%d_Gps\": %.2f", bps, n * 2 / ms / 1e6);
    ms = time_kernel(mul_kernel<FrBls, 1>, g, b, (FrBls*)buf, iters); printf(", \"fr_mul_ilp1_bps%d_Gps\": %.2f", bps, n * 1 / ms / 1e6);
    ms = time_kernel(mul_kernel<FrBls, 4>, g, b, (FrBls*)buf, iters); printf(", \"fr_mul_ilp4_bps%d_Gps\": %.2f", bps, n * 4 / ms / 1e6);
  }
  for (int bps = 1; bps <= 4; bps++) {
    dim3 g(sms * bps), b(128);
    double n = (double)sms * bps * 128 * 500;
    float ms = time_kernel(madd_kernel<FqBls>, g, b, (Affine<FqBls>*)buf, (XYZZ<FqBls>*)outb, 500);
    printf(", \"g1_madd_bps%d_Gps\": %.3f", bps, n / ms / 1e6);
  }
  printf("}\n");
  return 0;
}
