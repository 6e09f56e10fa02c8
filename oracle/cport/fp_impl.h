/* Montgomery field arithmetic over NL 64-bit limbs (CIOS with unsigned __int128), instantiated by
 * including this file with FP (name prefix), NL (limbs) defined.  Restates ark-ff 0.3
 * `Fp256` / `Fp384` arithmetic [U ark-ff fields/models/fp_256.rs, fp_384.rs]; same representation
 * (R = 2^(64 NL), little-endian limbs).  TEST INFRASTRUCTURE (oracle), not product code. */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(FP, name)

typedef struct { u64 l[NL]; } FN(_t);
typedef struct { u64 p[NL]; u64 inv; u64 r[NL]; u64 r2[NL]; } FN(_ctx);

static inline int FN(_is_zero)(const FN(_t)* a) { u64 t = 0; for (int i = 0; i < NL; i++) t |= a->l[i]; return t == 0; }
static inline int FN(_eq)(const FN(_t)* a, const FN(_t)* b) { u64 t = 0; for (int i = 0; i < NL; i++) t |= a->l[i] ^ b->l[i]; return t == 0; }
static inline int FN(_geq_p)(const u64* a, const FN(_ctx)* c) {
  for (int i = NL - 1; i >= 0; i--) { if (a[i] != c->p[i]) return a[i] > c->p[i]; }
  return 1;
}
static inline void FN(_sub_p)(u64* a, const FN(_ctx)* c) {
  u128 br = 0;
  for (int i = 0; i < NL; i++) { u128 t = (u128)a[i] - c->p[i] - (u64)br; a[i] = (u64)t; br = (t >> 64) & 1; }
}
static inline void FN(_add)(FN(_t)* r, const FN(_t)* a, const FN(_t)* b, const FN(_ctx)* c) {
  u128 cy = 0;
  for (int i = 0; i < NL; i++) { u128 t = (u128)a->l[i] + b->l[i] + (u64)cy; r->l[i] = (u64)t; cy = t >> 64; }
  if (cy || FN(_geq_p)(r->l, c)) FN(_sub_p)(r->l, c);
}
static inline void FN(_sub)(FN(_t)* r, const FN(_t)* a, const FN(_t)* b, const FN(_ctx)* c) {
  u128 br = 0;
  for (int i = 0; i < NL; i++) { u128 t = (u128)a->l[i] - b->l[i] - (u64)br; r->l[i] = (u64)t; br = (t >> 64) & 1; }
  if (br) { u128 cy = 0; for (int i = 0; i < NL; i++) { u128 t = (u128)r->l[i] + c->p[i] + (u64)cy; r->l[i] = (u64)t; cy = t >> 64; } }
}
static inline void FN(_dbl)(FN(_t)* r, const FN(_t)* a, const FN(_ctx)* c) { FN(_add)(r, a, a, c); }
static inline void FN(_neg)(FN(_t)* r, const FN(_t)* a, const FN(_ctx)* c) {
  if (FN(_is_zero)(a)) { *r = *a; return; }
  u128 br = 0;
  for (int i = 0; i < NL; i++) { u128 t = (u128)c->p[i] - a->l[i] - (u64)br; r->l[i] = (u64)t; br = (t >> 64) & 1; }
}
static inline void FN(_mul)(FN(_t)* r, const FN(_t)* a, const FN(_t)* b, const FN(_ctx)* c) {
  u64 t[NL + 2];
  for (int i = 0; i < NL + 2; i++) t[i] = 0;
  for (int i = 0; i < NL; i++) {
    u128 cy = 0;
    for (int j = 0; j < NL; j++) { u128 s = (u128)a->l[j] * b->l[i] + t[j] + (u64)cy; t[j] = (u64)s; cy = s >> 64; }
    u128 s = (u128)t[NL] + (u64)cy; t[NL] = (u64)s; t[NL + 1] = (u64)(s >> 64);
    u64 m = t[0] * c->inv;
    cy = ((u128)m * c->p[0] + t[0]) >> 64;
    for (int j = 1; j < NL; j++) { u128 s2 = (u128)m * c->p[j] + t[j] + (u64)cy; t[j - 1] = (u64)s2; cy = s2 >> 64; }
    s = (u128)t[NL] + (u64)cy; t[NL - 1] = (u64)s; t[NL] = t[NL + 1] + (u64)(s >> 64);
  }
  for (int i = 0; i < NL; i++) r->l[i] = t[i];
  if (t[NL] || FN(_geq_p)(r->l, c)) FN(_sub_p)(r->l, c);
}
static inline void FN(_sqr)(FN(_t)* r, const FN(_t)* a, const FN(_ctx)* c) { FN(_mul)(r, a, a, c); }
static void FN(_pow)(FN(_t)* r, const FN(_t)* a, const u64* e, int ne, const FN(_ctx)* c) {
  FN(_t) acc; memcpy(acc.l, c->r, sizeof(acc.l));
  for (int i = ne - 1; i >= 0; i--) for (int b = 63; b >= 0; b--) {
    FN(_sqr)(&acc, &acc, c);
    if ((e[i] >> b) & 1) FN(_mul)(&acc, &acc, a, c);
  }
  *r = acc;
}
static void FN(_inv)(FN(_t)* r, const FN(_t)* a, const FN(_ctx)* c) {  /* a^(p-2) */
  u64 e[NL]; memcpy(e, c->p, sizeof(e));
  u128 br = 2;
  for (int i = 0; i < NL; i++) { u128 t = (u128)e[i] - (u64)br; e[i] = (u64)t; br = (t >> 64) & 1; if (!br) break; }
  FN(_pow)(r, a, e, NL, c);
}
static void FN(_from_mont)(FN(_t)* r, const FN(_t)* a, const FN(_ctx)* c) { FN(_t) one = {{0}}; one.l[0] = 1; FN(_mul)(r, a, &one, c); }
static void FN(_to_mont)(FN(_t)* r, const FN(_t)* a, const FN(_ctx)* c) { FN(_t) r2; memcpy(r2.l, c->r2, sizeof(r2.l)); FN(_mul)(r, a, &r2, c); }
#undef FN
#undef CAT
#undef CAT_
