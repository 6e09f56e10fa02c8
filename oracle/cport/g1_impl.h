/* G1 in Jacobian coordinates over the field FQ (prefix), restating ark-ec 0.3
 * `GroupProjective::{double_in_place, add_assign_mixed, add_assign}` [U ark-ec models/short_weierstrass_jacobian.rs]
 * (dbl-2009-l, madd-2007-bl, add-2007-bl; a = 0) and `VariableBaseMSM::multi_scalar_mul`
 * [U ark-ec msm/variable_base.rs]: window c = 3 if n < 32 else ln(n) + 2, unsigned digits, 2^c - 1 buckets
 * per window, one parallel task per window, running-sum reduction, Horner over the windows. */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define Q(name) CAT(FQ, name)
#define GN(name) CAT(G, name)

typedef struct { Q(_t) x, y; } GN(_aff);       /* (0,0) = infinity */
typedef struct { Q(_t) x, y, z; } GN(_jac);     /* z = 0: infinity */

static inline int GN(_aff_is_inf)(const GN(_aff)* p) { return Q(_is_zero)(&p->x) && Q(_is_zero)(&p->y); }
static inline void GN(_jac_set_inf)(GN(_jac)* p) { memset(p, 0, sizeof(*p)); }
static inline int GN(_jac_is_inf)(const GN(_jac)* p) { return Q(_is_zero)(&p->z); }

static void GN(_dbl)(GN(_jac)* p, const Q(_ctx)* c) {
  if (GN(_jac_is_inf)(p)) return;
  Q(_t) a, b, cc, d, e, f, t;
  Q(_sqr)(&a, &p->x, c); Q(_sqr)(&b, &p->y, c); Q(_sqr)(&cc, &b, c);
  Q(_add)(&t, &p->x, &b, c); Q(_sqr)(&t, &t, c); Q(_sub)(&t, &t, &a, c); Q(_sub)(&t, &t, &cc, c); Q(_dbl)(&d, &t, c);
  Q(_dbl)(&e, &a, c); Q(_add)(&e, &e, &a, c);
  Q(_sqr)(&f, &e, c);
  Q(_mul)(&p->z, &p->z, &p->y, c); Q(_dbl)(&p->z, &p->z, c);
  Q(_sub)(&p->x, &f, &d, c); Q(_sub)(&p->x, &p->x, &d, c);
  Q(_sub)(&t, &d, &p->x, c); Q(_mul)(&t, &t, &e, c);
  Q(_dbl)(&cc, &cc, c); Q(_dbl)(&cc, &cc, c); Q(_dbl)(&cc, &cc, c);
  Q(_sub)(&p->y, &t, &cc, c);
}
static void GN(_add_mixed)(GN(_jac)* p, const GN(_aff)* q, const Q(_ctx)* c) {
  if (GN(_aff_is_inf)(q)) return;
  if (GN(_jac_is_inf)(p)) { p->x = q->x; p->y = q->y; memcpy(p->z.l, c->r, sizeof(p->z.l)); return; }
  Q(_t) z1z1, u2, s2, h, hh, i, j, r, v, t;
  Q(_sqr)(&z1z1, &p->z, c); Q(_mul)(&u2, &q->x, &z1z1, c);
  Q(_mul)(&s2, &p->z, &q->y, c); Q(_mul)(&s2, &s2, &z1z1, c);
  if (Q(_eq)(&p->x, &u2)) {
    if (Q(_eq)(&p->y, &s2)) { GN(_dbl)(p, c); return; }
    GN(_jac_set_inf)(p); return;
  }
  Q(_sub)(&h, &u2, &p->x, c); Q(_sqr)(&hh, &h, c); Q(_dbl)(&i, &hh, c); Q(_dbl)(&i, &i, c);
  Q(_mul)(&j, &h, &i, c);
  Q(_sub)(&r, &s2, &p->y, c); Q(_dbl)(&r, &r, c);
  Q(_mul)(&v, &p->x, &i, c);
  Q(_sqr)(&t, &r, c); Q(_sub)(&t, &t, &j, c); Q(_sub)(&t, &t, &v, c); Q(_sub)(&t, &t, &v, c);
  Q(_t) y1j; Q(_mul)(&y1j, &p->y, &j, c); Q(_dbl)(&y1j, &y1j, c);
  p->x = t;
  Q(_sub)(&t, &v, &p->x, c); Q(_mul)(&t, &t, &r, c); Q(_sub)(&p->y, &t, &y1j, c);
  Q(_add)(&t, &p->z, &h, c); Q(_sqr)(&t, &t, c); Q(_sub)(&t, &t, &z1z1, c); Q(_sub)(&p->z, &t, &hh, c);
}
static void GN(_add)(GN(_jac)* p, const GN(_jac)* q, const Q(_ctx)* c) {
  if (GN(_jac_is_inf)(q)) return;
  if (GN(_jac_is_inf)(p)) { *p = *q; return; }
  Q(_t) z1z1, z2z2, u1, u2, s1, s2, h, i, j, r, v, t;
  Q(_sqr)(&z1z1, &p->z, c); Q(_sqr)(&z2z2, &q->z, c);
  Q(_mul)(&u1, &p->x, &z2z2, c); Q(_mul)(&u2, &q->x, &z1z1, c);
  Q(_mul)(&s1, &p->y, &q->z, c); Q(_mul)(&s1, &s1, &z2z2, c);
  Q(_mul)(&s2, &q->y, &p->z, c); Q(_mul)(&s2, &s2, &z1z1, c);
  if (Q(_eq)(&u1, &u2)) {
    if (Q(_eq)(&s1, &s2)) { GN(_dbl)(p, c); return; }
    GN(_jac_set_inf)(p); return;
  }
  Q(_sub)(&h, &u2, &u1, c); Q(_dbl)(&i, &h, c); Q(_sqr)(&i, &i, c); Q(_mul)(&j, &h, &i, c);
  Q(_sub)(&r, &s2, &s1, c); Q(_dbl)(&r, &r, c);
  Q(_mul)(&v, &u1, &i, c);
  Q(_sqr)(&t, &r, c); Q(_sub)(&t, &t, &j, c); Q(_sub)(&t, &t, &v, c); Q(_sub)(&t, &t, &v, c);
  Q(_t) s1j; Q(_mul)(&s1j, &s1, &j, c); Q(_dbl)(&s1j, &s1j, c);
  Q(_t) zz; Q(_add)(&zz, &p->z, &q->z, c); Q(_sqr)(&zz, &zz, c); Q(_sub)(&zz, &zz, &z1z1, c); Q(_sub)(&zz, &zz, &z2z2, c);
  p->x = t;
  Q(_sub)(&t, &v, &p->x, c); Q(_mul)(&t, &t, &r, c); Q(_sub)(&p->y, &t, &s1j, c);
  Q(_mul)(&p->z, &zz, &h, c);
}
static void GN(_to_affine)(GN(_aff)* r, const GN(_jac)* p, const Q(_ctx)* c) {
  if (GN(_jac_is_inf)(p)) { memset(r, 0, sizeof(*r)); return; }
  Q(_t) zi, zi2, zi3; Q(_inv)(&zi, &p->z, c); Q(_sqr)(&zi2, &zi, c); Q(_mul)(&zi3, &zi2, &zi, c);
  Q(_mul)(&r->x, &p->x, &zi2, c); Q(_mul)(&r->y, &p->y, &zi3, c);
}
static int GN(_arkworks_window)(size_t n) {
  if (n < 32) return 3;
  int lg = 63 - __builtin_clzll((unsigned long long)n);
  return lg * 69 / 100 + 2;
}
/* scalars: canonical 4-limb integers (into_repr) */
static void GN(_msm)(GN(_jac)* out, const GN(_aff)* bases, const u64* scalars, size_t n, int scalar_bits, const Q(_ctx)* c, int threads) {
  int w = GN(_arkworks_window)(n);
  int nwin = (scalar_bits + w - 1) / w;
  GN(_jac)* sums = (GN(_jac)*)malloc(sizeof(GN(_jac)) * nwin);
  #pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int wi = 0; wi < nwin; wi++) {
    int start = wi * w;
    size_t nb = ((size_t)1 << w) - 1;
    GN(_jac)* buckets = (GN(_jac)*)calloc(nb, sizeof(GN(_jac)));
    GN(_jac) res; GN(_jac_set_inf)(&res);
    for (size_t i = 0; i < n; i++) {
      const u64* s = scalars + 4 * i;
      if ((s[0] | s[1] | s[2] | s[3]) == 0) continue;
      if (s[0] == 1 && (s[1] | s[2] | s[3]) == 0) { if (start == 0) GN(_add_mixed)(&res, &bases[i], c); continue; }
      int limb = start >> 6, off = start & 63;
      u64 d = s[limb] >> off;
      if (off + w > 64 && limb + 1 < 4) d |= s[limb + 1] << (64 - off);
      d &= ((u64)1 << w) - 1;
      if (d) GN(_add_mixed)(&buckets[d - 1], &bases[i], c);
    }
    GN(_jac) running; GN(_jac_set_inf)(&running);
    for (size_t b = nb; b-- > 0;) { GN(_add)(&running, &buckets[b], c); GN(_add)(&res, &running, c); }
    sums[wi] = res;
    free(buckets);
  }
  GN(_jac) total; GN(_jac_set_inf)(&total);
  for (int wi = nwin - 1; wi >= 1; wi--) {
    GN(_add)(&total, &sums[wi], c);
    for (int k = 0; k < w; k++) GN(_dbl)(&total, c);
  }
  GN(_add)(&total, &sums[0], c);
  *out = total;
  free(sums);
}
#undef Q
#undef GN
#undef CAT
#undef CAT_
