// Exclusive prefix sum over uint32 (bucket sizes of the MSM counting sort, acceptance flags of the
// mask-polynomial sampler).  Kernels are `static`: each translation unit that includes this gets its own.
#pragma once
#include "common.cuh"

namespace b2m {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 4;
static __global__ void scan_block_kernel(const uint32_t* in, uint32_t* out, size_t n, uint32_t* block_sums) {
  __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
  size_t base = (size_t)blockIdx.x * SCAN_THREADS * SCAN_ITEMS + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    sum += v[k];
  }
  uint32_t incl = sum;
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t ws = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0, wi = ws;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - ws;  // exclusive
    if (lane == SCAN_THREADS / 32 - 1 && block_sums) block_sums[blockIdx.x] = wi;
  }
  __syncthreads();
  uint32_t excl = incl - sum + warp_sums[wid];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) out[base + k] = excl;
    excl += v[k];
  }
}
static __global__ void scan_add_kernel(uint32_t* out, size_t n, const uint32_t* block_offsets) {
  size_t i = (size_t)blockIdx.x * SCAN_THREADS * SCAN_ITEMS + threadIdx.x;
  uint32_t o = block_offsets[blockIdx.x];
  for (int k = 0; k < SCAN_ITEMS; k++, i += SCAN_THREADS)
    if (i < n) out[i] += o;
}
static void exclusive_scan_u32(Ctx& ctx, const uint32_t* in, uint32_t* out, size_t n) {
  const size_t per = SCAN_THREADS * SCAN_ITEMS;
  size_t blocks = (n + per - 1) / per;
  if (blocks <= 1) {
    scan_block_kernel<<<1, SCAN_THREADS, 0, ctx.stream>>>(in, out, n, nullptr);
    B2M_CHECK_LAUNCH();
    ctx.launches++;
    return;
  }
  DBuf<uint32_t> sums(ctx, blocks), offs(ctx, blocks);
  scan_block_kernel<<<(unsigned)blocks, SCAN_THREADS, 0, ctx.stream>>>(in, out, n, sums.p);
  B2M_CHECK_LAUNCH();
  ctx.launches++;
  exclusive_scan_u32(ctx, sums.p, offs.p, blocks);
  scan_add_kernel<<<(unsigned)blocks, SCAN_THREADS, 0, ctx.stream>>>(out, n, offs.p);
  B2M_CHECK_LAUNCH();
  ctx.launches++;
}

}  // namespace b2m
