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

// Lifetimes: an index / committer key borrows its SRS, an SRS borrows its context.  Destroying a parent while children are
// alive only marks it; the storage goes when the last child is destroyed, so no destroy order is a use-after-free.
struct b2m_ctx {
  b2m::Ctx cx;
  int children = 0;
  bool dead = false;
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
  int children = 0;
  bool dead = false;
  int curve;
  size_t n_g, n_gamma;
  std::unique_ptr<b2m::Msm<b2m::FrBls, b2m::FqBls>> bls;
  std::unique_ptr<b2m::Msm<b2m::FrBn, b2m::FqBn>> bn;
  // gamma_idx[k] = the power of beta held in slot k of the powers_of_gamma_g (they live in the window
  // tables right after the G1 powers: see Msm::n_extra)
  std::vector<uint64_t> gamma_idx;

  // slot of beta^i * gamma * G, or throws
  size_t gamma_slot(uint64_t i) const {
    for (size_t k = 0; k < gamma_idx.size(); k++)
      if (gamma_idx[k] == i) return k;
    throw b2m::Error(B2M_ERR_INVALID_ARG, b2m::fmt("the SRS holds no power %llu of gamma*G", (unsigned long long)i));
  }

  b2m_srs(b2m_ctx* c, int curve_, const uint64_t* g, size_t ng, const uint64_t* gamma, const uint64_t* gidx, size_t ngamma,
          int window_bits)
      : ctx(c), curve(curve_), n_g(ng), n_gamma(ngamma) {
    using namespace b2m;
    for (size_t k = 0; k < ngamma; k++) gamma_idx.push_back(gidx ? gidx[k] : k);
    if (curve == B2M_CURVE_BLS12_381) {
      bls.reset(new Msm<FrBls, FqBls>(c->cx, reinterpret_cast<const Affine<FqBls>*>(g), ng, reinterpret_cast<const Affine<FqBls>*>(gamma), ngamma,
                                      window_bits));
    } else {
      bn.reset(new Msm<FrBn, FqBn>(c->cx, reinterpret_cast<const Affine<FqBn>*>(g), ng, reinterpret_cast<const Affine<FqBn>*>(gamma), ngamma,
                                  window_bits));
    }
    c->cx.sync();
  }
  int window_bits() const { return bls ? bls->c : bn->c; }
  int affine_levels() const { return bls ? bls->affine_levels : bn->affine_levels; }
  size_t affine_min_refs() const { return bls ? bls->affine_min_refs : bn->affine_min_refs; }
};


// `PC::CommitterKey` after `PC::trim`: a validated view of the device-resident SRS
struct b2m_ck {
  b2m_srs* srs;
  int pc;
  size_t supported_degree, hiding_bound;
  std::vector<uint64_t> bounds;  // enforced degree bounds, sorted
  int64_t max_bound() const { return bounds.empty() ? -1 : (int64_t)bounds.back(); }
  bool enforced(uint64_t b) const {
    for (uint64_t x : bounds)
      if (x == b) return true;
    return false;
  }
};

void b2m_release_ctx(b2m_ctx* ctx);
void b2m_release_srs(b2m_srs* srs);
