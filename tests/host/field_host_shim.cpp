// Host build of marlin_b200/csrc/field.cuh (carry flag emulated) exposed over a tiny C ABI so
// pytest can compare the limb-level Montgomery code against Python big integers without a GPU.
#include "../../marlin_b200/csrc/field.cuh"
using namespace b2m;
template <class F> static void op(int which, const uint32_t* a, const uint32_t* b, uint32_t* r) {
  F x, y, z;
  memcpy(x.l, a, sizeof(x.l));
  memcpy(y.l, b, sizeof(y.l));
  switch (which) {
    case 0: z = x * y; break;
    case 1: z = x + y; break;
    case 2: z = x - y; break;
    case 3: z = x.neg(); break;
    case 4: z = x.inverse(); break;
    case 5: z = x.to_canonical(); break;
    case 6: z = F::from_canonical(x); break;
    case 7: z = x.inverse_fast(); break;
    default: z = F::zero();
  }
  memcpy(r, z.l, sizeof(z.l));
}
extern "C" void field_op(int field, int which, const uint32_t* a, const uint32_t* b, uint32_t* r) {
  switch (field) {
    case 0: op<FrBls>(which, a, b, r); break;
    case 1: op<FqBls>(which, a, b, r); break;
    case 2: op<FrBn>(which, a, b, r); break;
    case 3: op<FqBn>(which, a, b, r); break;
  }
}

#include "../../marlin_b200/csrc/curve.cuh"
// XYZZ ops on the host: pts are affine (x,y) Montgomery limbs; (0,0) = infinity.
// which: 0 = sum of all points via add_mixed (with negate flags), 1 = tree of full adds, 2 = scalar_mul(p[0], k)
template <class Fq> static void cop(int which, const uint32_t* pts, const uint8_t* neg, int n, const uint32_t* k, int klimbs, uint32_t* out) {
  const Affine<Fq>* P = reinterpret_cast<const Affine<Fq>*>(pts);
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
  if (which == 0) {
    for (int i = 0; i < n; i++) acc.add_mixed(P[i], neg[i] != 0);
  } else if (which == 1) {
    XYZZ<Fq> a = XYZZ<Fq>::inf(), b = XYZZ<Fq>::inf();
    for (int i = 0; i < n; i++) { if (i & 1) a.add_mixed(P[i], neg[i] != 0); else b.add_mixed(P[i], neg[i] != 0); }
    a.add(b); acc = a; acc.add(XYZZ<Fq>::inf());
    XYZZ<Fq> z = XYZZ<Fq>::inf(); z.add(acc); acc = z;
  } else {
    acc = scalar_mul<Fq>(P[0], k, klimbs);
  }
  Affine<Fq> r = acc.to_affine();
  memcpy(out, &r, sizeof(r));
}
extern "C" void curve_op(int curve, int which, const uint32_t* pts, const uint8_t* neg, int n, const uint32_t* k, int klimbs, uint32_t* out) {
  if (curve == 0) cop<FqBls>(which, pts, neg, n, k, klimbs, out); else cop<FqBn>(which, pts, neg, n, k, klimbs, out);
}

#include <vector>
#include "../../marlin_b200/csrc/msm_affine.cuh"
// Batched-affine levels on the host, one "thread" after the other: `levels` levels over the bucket-sorted
// references, then every bucket's remaining points are summed with XYZZ additions (what the XYZZ pass does).
template <class Fq>
static void run_levels(const uint32_t* tables, const uint32_t* refs, const uint32_t* off0, uint32_t B, int levels, uint32_t T, uint32_t* out_buckets,
                       int variant, int interleaved) {
  const Affine<Fq>* tab = reinterpret_cast<const Affine<Fq>*>(tables);
  const uint2* sorted = reinterpret_cast<const uint2*>(refs);
  std::vector<uint32_t> off_in(off0, off0 + B + 1), off_out(B + 1);
  std::vector<Affine<Fq>> in, out;
  std::vector<uint2> out_refs;
  for (int l = 0; l < levels; l++) {
    off_out[0] = 0;
    for (uint32_t b = 0; b < B; b++) off_out[b + 1] = off_out[b] + (off_in[b + 1] - off_in[b] + 1) / 2;
    const uint32_t total = off_out[B];
    const uint32_t lane_step = interleaved ? 32u : 1u;
    const uint32_t nthreads = lane_step * ((total + lane_step * T - 1) / (lane_step * T));
    out.assign(total, Affine<Fq>::inf());
    out_refs.assign(total, uint2{0, 0});
    std::vector<Fq> pref((size_t)T * (nthreads ? nthreads : 1)), invs(nthreads ? nthreads : 1);
    std::vector<uint4> meta((size_t)T * (nthreads ? nthreads : 1));
    AffLevel<Fq> A{tab, 0, sorted, in.data(), off_in.data(), off_out.data(), B, out.data(), l == levels - 1 ? out_refs.data() : nullptr,
                   pref.data(), meta.data(), T, nthreads, lane_step, invs.data()};
    for (uint32_t t = 0; t < nthreads; t++) {
      if (l == 0) aff_plan_thread<Fq, true>(A, t); else aff_plan_thread<Fq, false>(A, t);
    }
    for (uint32_t t = 0; t < nthreads; t++) {
      const Affine<Fq>* base = l == 0 ? tab : in.data();
      // variant 0: fused kernel's thread function (odd T: with operand prefetch); 3: software-pipelined; 4 / 5: its split
      // (two-kernel) form with the pipelined / plain addition pass
      if (variant == 3) aff_level_thread_sp<Fq, 0>(A, base, t);
      else if (variant == 4) { aff_level_thread_sp<Fq, 1>(A, base, t); aff_level_thread_sp<Fq, 2>(A, base, t); }
      else if (variant == 5) { aff_level_thread_sp<Fq, 1>(A, base, t); aff_level_thread_sp<Fq, 2, false>(A, base, t); }
      else if (variant == 6) { aff_level_thread_sp<Fq, 3>(A, base, t); invs[t] = invs[t].inverse_fast(); aff_level_thread_sp<Fq, 2, false>(A, base, t); }  // (the device inverts the level's chain products together)
      else if (T & 1) aff_level_thread<Fq, true>(A, base, t);
      else aff_level_thread<Fq, false>(A, base, t);
    }
    in.swap(out);
    off_in = off_out;
  }
  std::vector<XYZZ<Fq>> acc(B, XYZZ<Fq>::inf());
  if (levels == 0) {
    for (uint32_t b = 0; b < B; b++)
      for (uint32_t i = off0[b]; i < off0[b + 1]; i++) acc[b].add_mixed(tab[sorted[i].x & 0x7fffffffu], sorted[i].x >> 31);
  } else {
    for (uint32_t i = 0; i < off_in[B]; i++) acc[out_refs[i].y].add_mixed(in[out_refs[i].x], false);
  }
  Affine<Fq>* ob = reinterpret_cast<Affine<Fq>*>(out_buckets);
  for (uint32_t b = 0; b < B; b++) ob[b] = acc[b].to_affine();
}
extern "C" void affine_levels_host(int curve, const uint32_t* tables, const uint32_t* refs, const uint32_t* off0, uint32_t B, int levels, uint32_t T,
                                   uint32_t* out_buckets, int variant, int interleaved) {
  if (curve == 0) run_levels<FqBls>(tables, refs, off0, B, levels, T, out_buckets, variant, interleaved);
  else run_levels<FqBn>(tables, refs, off0, B, levels, T, out_buckets, variant, interleaved);
}
