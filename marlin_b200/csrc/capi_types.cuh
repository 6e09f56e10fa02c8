// Definitions of the opaque handles of include/b2m.h and the host-buffer wrappers of Level 0.
#pragma once
#include <memory>
#include <string>

#include "common.cuh"
#include "msm.cuh"
#include "ntt.cuh"

namespace b2m {
extern thread_local std::string g_last_error;

template <class F>
int guard(F&& f) {
  try {
    f();
    return B2M_OK;
  } catch (const Error& e) {
    g_last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return B2M_ERR_INVALID_ARG;
  }
}
}  // namespace b2m

struct b2m_ctx {
  b2m::Ctx cx;
  std::unique_ptr<b2m::Ntt<b2m::FrBls>> ntt_bls_;
  std::unique_ptr<b2m::Ntt<b2m::FrBn>> ntt_bn_;
  explicit b2m_ctx(int device) : cx(device) {}
  b2m::Ntt<b2m::FrBls>& ntt_bls() {
    if (!ntt_bls_) ntt_bls_.reset(new b2m::Ntt<b2m::FrBls>(cx));
    return *ntt_bls_;
  }
  b2m::Ntt<b2m::FrBn>& ntt_bn() {
    if (!ntt_bn_) ntt_bn_.reset(new b2m::Ntt<b2m::FrBn>(cx));
    return *ntt_bn_;
  }
};

struct b2m_srs {
  b2m_ctx* ctx;
  int curve;
  size_t n_g, n_gamma;
  std::unique_ptr<b2m::Msm<b2m::FrBls, b2m::FqBls>> bls;
  std::unique_ptr<b2m::Msm<b2m::FrBn, b2m::FqBn>> bn;
  // powers_of_gamma_g (hiding bases), device resident, raw bytes (Affine<Fq>[n_gamma])
  void* gamma_dev = nullptr;

  b2m_srs(b2m_ctx* c, int curve_, const uint64_t* g, size_t ng, const uint64_t* gamma, size_t ngamma, int window_bits)
      : ctx(c), curve(curve_), n_g(ng), n_gamma(ngamma) {
    using namespace b2m;
    if (curve == B2M_CURVE_BLS12_381) {
      bls.reset(new Msm<FrBls, FqBls>(c->cx, reinterpret_cast<const Affine<FqBls>*>(g), ng, window_bits));
      if (ngamma) {
        gamma_dev = c->cx.alloc_bytes(ngamma * sizeof(Affine<FqBls>));
        B2M_CUDA(cudaMemcpyAsync(gamma_dev, gamma, ngamma * sizeof(Affine<FqBls>), cudaMemcpyHostToDevice, c->cx.stream));
      }
    } else {
      bn.reset(new Msm<FrBn, FqBn>(c->cx, reinterpret_cast<const Affine<FqBn>*>(g), ng, window_bits));
      if (ngamma) {
        gamma_dev = c->cx.alloc_bytes(ngamma * sizeof(Affine<FqBn>));
        B2M_CUDA(cudaMemcpyAsync(gamma_dev, gamma, ngamma * sizeof(Affine<FqBn>), cudaMemcpyHostToDevice, c->cx.stream));
      }
    }
    c->cx.sync();
  }
  ~b2m_srs() {
    if (gamma_dev) ctx->cx.free_bytes(gamma_dev);
  }
  int window_bits() const { return bls ? bls->c : bn->c; }
};

