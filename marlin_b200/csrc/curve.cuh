// Short-Weierstrass (a = 0) G1 arithmetic in extended Jacobian ("XYZZ") coordinates.
//
// Replaces, for the MSM behind every `PC::commit`/`open` call [R src/lib.rs:172,193,213,292],
// ark-ec 0.3's `GroupProjective::add_assign_mixed` / `double_in_place`
// [U ark-ec models/short_weierstrass_jacobian.rs].  The group law is the same, the
// coordinate system is not (XYZZ saves two field multiplications per mixed addition);
// results are compared only after conversion to the unique affine form.
//
// Point at infinity: ZZ == 0 (XYZZ) / inf flag (affine).  Formulas: EFD
// "madd-2008-s", "add-2008-s", "dbl-2008-s-1", "mdbl-2008-s-1" for y^2 = x^3 + b.
#pragma once
#include "field.cuh"

namespace b2m {

template <class Fq>
struct Affine {
  Fq x, y;  // Montgomery form; (0,0) encodes infinity (never a curve point since b != 0)
  B2M_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
  B2M_HD static Affine inf() { return Affine{Fq::zero(), Fq::zero()}; }
};

template <class Fq>
struct XYZZ {
  Fq X, Y, ZZ, ZZZ;

  B2M_HD static XYZZ inf() { return XYZZ{Fq::zero(), Fq::zero(), Fq::zero(), Fq::zero()}; }
  B2M_HD bool is_inf() const { return ZZ.is_zero(); }
  B2M_HD static XYZZ from_affine(const Affine<Fq>& p) {
    if (p.is_inf()) return inf();
    return XYZZ{p.x, p.y, Fq::one(), Fq::one()};
  }

  // 2 * (affine p)
  B2M_HD static XYZZ dbl_affine(const Affine<Fq>& p) {
    if (p.is_inf() || p.y.is_zero()) return inf();
    Fq U = p.y.dbl();
    Fq V = U.sqr();
    Fq W = U * V;
    Fq S = p.x * V;
    Fq xx = p.x.sqr();
    Fq M = xx.dbl() + xx;
    XYZZ r;
    r.X = M.sqr() - S.dbl();
    r.Y = M * (S - r.X) - W * p.y;
    r.ZZ = V;
    r.ZZZ = W;
    return r;
  }

  B2M_HD XYZZ dbl() const {
    if (is_inf() || Y.is_zero()) return inf();
    Fq U = Y.dbl();
    Fq V = U.sqr();
    Fq W = U * V;
    Fq S = X * V;
    Fq xx = X.sqr();
    Fq M = xx.dbl() + xx;
    XYZZ r;
    r.X = M.sqr() - S.dbl();
    r.Y = M * (S - r.X) - W * Y;
    r.ZZ = V * ZZ;
    r.ZZZ = W * ZZZ;
    return r;
  }

  // this += (affine p), p negated first when `neg`.
  B2M_HD void add_mixed(const Affine<Fq>& p, bool neg = false) {
    if (p.is_inf()) return;
    Fq py = neg ? p.y.neg() : p.y;
    if (is_inf()) {
      X = p.x;
      Y = py;
      ZZ = Fq::one();
      ZZZ = Fq::one();
      return;
    }
    Fq U2 = p.x * ZZ;
    Fq S2 = py * ZZZ;
    Fq Pp = U2 - X;
    Fq R = S2 - Y;
    if (Pp.is_zero()) {
      if (R.is_zero()) {
        *this = dbl_affine(Affine<Fq>{p.x, py});
      } else {
        *this = inf();
      }
      return;
    }
    Fq PP = Pp.sqr();
    Fq PPP = Pp * PP;
    Fq Q = X * PP;
    Fq X3 = R.sqr() - PPP - Q.dbl();
    Fq Y3 = R * (Q - X3) - Y * PPP;
    X = X3;
    Y = Y3;
    ZZ = ZZ * PP;
    ZZZ = ZZZ * PPP;
  }

  // this += o
  B2M_HD void add(const XYZZ& o) {
    if (o.is_inf()) return;
    if (is_inf()) {
      *this = o;
      return;
    }
    Fq U1 = X * o.ZZ;
    Fq U2 = o.X * ZZ;
    Fq S1 = Y * o.ZZZ;
    Fq S2 = o.Y * ZZZ;
    Fq Pp = U2 - U1;
    Fq R = S2 - S1;
    if (Pp.is_zero()) {
      if (R.is_zero()) {
        *this = dbl();
      } else {
        *this = inf();
      }
      return;
    }
    Fq PP = Pp.sqr();
    Fq PPP = Pp * PP;
    Fq Q = U1 * PP;
    Fq X3 = R.sqr() - PPP - Q.dbl();
    Fq Y3 = R * (Q - X3) - S1 * PPP;
    X = X3;
    Y = Y3;
    ZZ = ZZ * o.ZZ * PP;
    ZZZ = ZZZ * o.ZZZ * PPP;
  }

  B2M_HD XYZZ negated() const { return XYZZ{X, Y.neg(), ZZ, ZZZ}; }

  // x = X/ZZ, y = Y/ZZZ with one inversion: ZZ^3 == ZZZ^2, so 1/ZZ = (ZZ/ZZZ)^2.
  B2M_HD Affine<Fq> to_affine() const {
    if (is_inf()) return Affine<Fq>::inf();
    Fq izzz = ZZZ.inverse();
    Fq izz = (ZZ * izzz).sqr();
    return Affine<Fq>{X * izz, Y * izzz};
  }
};

// k * P by double-and-add over a canonical little-endian scalar (nlimbs 32-bit limbs).
template <class Fq>
B2M_HD XYZZ<Fq> scalar_mul(const Affine<Fq>& p, const uint32_t* k, int nlimbs) {
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
  bool started = false;
  for (int i = nlimbs - 1; i >= 0; i--) {
    for (int b = 31; b >= 0; b--) {
      if (started) acc = acc.dbl();
      if ((k[i] >> b) & 1u) {
        acc.add_mixed(p);
        started = true;
      }
    }
  }
  return acc;
}

}  // namespace b2m
