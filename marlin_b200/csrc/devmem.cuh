// Vectorised 128-bit global/shared memory access for limb structs (Fr, Fq, Affine, XYZZ).
#pragma once
#include "field.cuh"

namespace b2m {

template <class T>
__device__ __forceinline__ T ld_words(const T* p) {
  T r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) {
    uint4 v = q[i];
    d[4 * i] = v.x; d[4 * i + 1] = v.y; d[4 * i + 2] = v.z; d[4 * i + 3] = v.w;
  }
  return r;
}
// read-only (non-coherent) path
template <class T>
__device__ __forceinline__ T ldg_words(const T* p) {
  T r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) {
    uint4 v = __ldg(q + i);
    d[4 * i] = v.x; d[4 * i + 1] = v.y; d[4 * i + 2] = v.z; d[4 * i + 3] = v.w;
  }
  return r;
}
template <class T>
__device__ __forceinline__ void st_words(T* p, const T& v) {
  uint4* q = reinterpret_cast<uint4*>(p);
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) q[i] = make_uint4(s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]);
}
template <class Fr> __device__ __forceinline__ Fr ld_fr(const Fr* p) { return ld_words(p); }
template <class Fr> __device__ __forceinline__ Fr ldg_fr(const Fr* p) { return ldg_words(p); }
template <class Fr> __device__ __forceinline__ void st_fr(Fr* p, const Fr& r) { st_words(p, r); }

}  // namespace b2m
