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
