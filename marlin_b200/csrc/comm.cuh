// NCCL plumbing for the multi-GPU MSM (one process per GPU; config 5 of BASELINE.json).  NCCL is
// resolved with dlopen at attach time, so single-GPU use has no NCCL dependency and a process that
// already loaded torch's bundled libnccl.so.2 shares that copy (same SONAME).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace b2m {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;

  static NcclApi& get() {
    static NcclApi api;
    if (!api.handle) {
      api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
      B2M_REQUIRE(api.handle != nullptr, B2M_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
      api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle, "ncclGetUniqueId"));
      api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle, "ncclCommInitRank"));
      api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.handle, "ncclAllGather"));
      api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(dlsym(api.handle, "ncclBroadcast"));
      api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
      api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
      B2M_REQUIRE(api.GetUniqueId && api.CommInitRank && api.AllGather && api.Broadcast && api.CommDestroy, B2M_ERR_NCCL, "libnccl lacks a required symbol");
    }
    return api;
  }
};

#define B2M_NCCL(expr)                                                                                           \
  do {                                                                                                           \
    ncclResult_t r__ = (expr);                                                                                   \
    if (r__ != ncclSuccess) {                                                                                    \
      auto& a__ = ::b2m::NcclApi::get();                                                                         \
      throw ::b2m::Error(B2M_ERR_NCCL, ::b2m::fmt("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                  \
                                                  a__.GetErrorString ? a__.GetErrorString(r__) : "nccl error")); \
    }                                                                                                            \
  } while (0)

inline void all_gather_bytes(Ctx& cx, const void* send, void* recv, size_t bytes_per_rank) {
  B2M_NCCL(NcclApi::get().AllGather(send, recv, bytes_per_rank, ncclUint8, static_cast<ncclComm_t>(cx.comm), cx.stream));
}

// in-place broadcast of `bytes` from rank `root` (every rank passes the same buffer size)
inline void broadcast_bytes(Ctx& cx, void* buf, size_t bytes, int root) {
  B2M_NCCL(NcclApi::get().Broadcast(buf, buf, bytes, ncclUint8, root, static_cast<ncclComm_t>(cx.comm), cx.stream));
}

}  // namespace b2m
