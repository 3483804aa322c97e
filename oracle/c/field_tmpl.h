/* ORACLE -- test infrastructure only (see oracle/README.md). Never linked into the product.
 *
 * 4x64-bit Montgomery prime-field template, R = 2^256, restating the semantics of
 * pasta_curves 0.4.1 `fields/fp.rs` / `fields/fq.rs` (crate pinned at
 * /root/reference/Cargo.lock:567-579, not vendored): in-memory form = Montgomery
 * little-endian [u64;4]; to_repr = canonical LE bytes; from_u512 = 512-bit LE mod p.
 *
 * Include with:  #define FE(name) fp_##name   and  FE_MOD0..3, FE_INV (= -p^-1 mod 2^64)
 */
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 FE(u128);

typedef struct { uint64_t l[4]; } FE(t);

static const uint64_t FE(MOD)[4] = { FE_MOD0, FE_MOD1, FE_MOD2, FE_MOD3 };

static inline int FE(geq_mod)(const uint64_t a[4]) {
  for (int i = 3; i >= 0; --i) {
    if (a[i] > FE(MOD)[i]) return 1;
    if (a[i] < FE(MOD)[i]) return 0;
  }
  return 1;
}

static inline void FE(sub_mod_raw)(uint64_t a[4]) {
  unsigned __int128 br = 0;
  for (int i = 0; i < 4; ++i) {
    unsigned __int128 d = (unsigned __int128)a[i] - FE(MOD)[i] - (uint64_t)br;
    a[i] = (uint64_t)d;
    br = (d >> 64) & 1;
  }
}

static inline int FE(is_zero)(const FE(t)* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int FE(eq)(const FE(t)* a, const FE(t)* b) {
  return a->l[0] == b->l[0] && a->l[1] == b->l[1] && a->l[2] == b->l[2] && a->l[3] == b->l[3];
}

static inline void FE(add)(FE(t)* r, const FE(t)* a, const FE(t)* b) {
  unsigned __int128 c = 0;
  uint64_t t[4];
  for (int i = 0; i < 4; ++i) { c += (unsigned __int128)a->l[i] + b->l[i]; t[i] = (uint64_t)c; c >>= 64; }
  /* p < 2^255 so no carry out of limb 3 for reduced inputs */
  if (FE(geq_mod)(t)) FE(sub_mod_raw)(t);
  memcpy(r->l, t, 32);
}

static inline void FE(sub)(FE(t)* r, const FE(t)* a, const FE(t)* b) {
  uint64_t t[4];
  unsigned __int128 br = 0;
  for (int i = 0; i < 4; ++i) {
    unsigned __int128 d = (unsigned __int128)a->l[i] - b->l[i] - (uint64_t)br;
    t[i] = (uint64_t)d;
    br = (d >> 64) & 1;
  }
  if (br) {
    unsigned __int128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (unsigned __int128)t[i] + FE(MOD)[i]; t[i] = (uint64_t)c; c >>= 64; }
  }
  memcpy(r->l, t, 32);
}

static inline void FE(neg)(FE(t)* r, const FE(t)* a) {
  FE(t) z = {{0, 0, 0, 0}};
  FE(sub)(r, &z, a);
}

static inline void FE(dbl)(FE(t)* r, const FE(t)* a) { FE(add)(r, a, a); }

/* Montgomery product a*b*R^-1 mod p (coarsely-integrated operand scanning).
 * Accepts a < 2^256 (not necessarily reduced) as long as b < p: result < 2p, then one
 * conditional subtraction. */
static inline void FE(mul)(FE(t)* r, const FE(t)* a, const FE(t)* b) {
  uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
  for (int i = 0; i < 4; ++i) {
    unsigned __int128 acc;
    uint64_t bi = b->l[i], c;
    acc = (unsigned __int128)a->l[0] * bi + t0; t0 = (uint64_t)acc; c = (uint64_t)(acc >> 64);
    acc = (unsigned __int128)a->l[1] * bi + t1 + c; t1 = (uint64_t)acc; c = (uint64_t)(acc >> 64);
    acc = (unsigned __int128)a->l[2] * bi + t2 + c; t2 = (uint64_t)acc; c = (uint64_t)(acc >> 64);
    acc = (unsigned __int128)a->l[3] * bi + t3 + c; t3 = (uint64_t)acc; c = (uint64_t)(acc >> 64);
    acc = (unsigned __int128)t4 + c; t4 = (uint64_t)acc; uint64_t t5 = (uint64_t)(acc >> 64);
    uint64_t m = t0 * FE_INV;
    acc = (unsigned __int128)m * FE_MOD0 + t0; c = (uint64_t)(acc >> 64);
    acc = (unsigned __int128)m * FE_MOD1 + t1 + c; t0 = (uint64_t)acc; c = (uint64_t)(acc >> 64);
    acc = (unsigned __int128)m * FE_MOD2 + t2 + c; t1 = (uint64_t)acc; c = (uint64_t)(acc >> 64);
    acc = (unsigned __int128)m * FE_MOD3 + t3 + c; t2 = (uint64_t)acc; c = (uint64_t)(acc >> 64);
    acc = (unsigned __int128)t4 + c; t3 = (uint64_t)acc; t4 = t5 + (uint64_t)(acc >> 64);
  }
  uint64_t t[4] = {t0, t1, t2, t3};
  if (t4 || FE(geq_mod)(t)) FE(sub_mod_raw)(t);
  memcpy(r->l, t, 32);
}

static inline void FE(sqr)(FE(t)* r, const FE(t)* a) { FE(mul)(r, a, a); }

/* constants filled by FE(init)() */
static FE(t) FE(R1);   /* 1 in Montgomery form */
static FE(t) FE(R2c);  /* R^2 */
static FE(t) FE(R3c);  /* R^3 */

static inline void FE(from_raw)(FE(t)* r, const uint64_t v[4]) {   /* canonical < 2^256 -> Montgomery */
  FE(t) a; memcpy(a.l, v, 32);
  FE(mul)(r, &a, &FE(R2c));
}
static inline void FE(from_u64)(FE(t)* r, uint64_t v) { uint64_t a[4] = {v, 0, 0, 0}; FE(from_raw)(r, a); }

static inline void FE(to_raw)(uint64_t out[4], const FE(t)* a) {     /* Montgomery -> canonical */
  FE(t) one = {{1, 0, 0, 0}}, r;
  FE(mul)(&r, a, &one);
  memcpy(out, r.l, 32);
}
static inline void FE(to_repr)(uint8_t out[32], const FE(t)* a) { uint64_t v[4]; FE(to_raw)(v, a); memcpy(out, v, 32); }

/* U: pasta `from_u512`: (lo + hi*2^256) mod p  ==  lo*R2*R^-1 + hi*R3*R^-1 in Montgomery terms */
static inline void FE(from_u512)(FE(t)* r, const uint64_t w[8]) {
  FE(t) lo, hi, a, b;
  memcpy(lo.l, w, 32); memcpy(hi.l, w + 4, 32);
  FE(mul)(&a, &lo, &FE(R2c));
  FE(mul)(&b, &hi, &FE(R3c));
  FE(add)(r, &a, &b);
}
static inline void FE(from_bytes_wide)(FE(t)* r, const uint8_t b[64]) { uint64_t w[8]; memcpy(w, b, 64); FE(from_u512)(r, w); }

static inline void FE(pow)(FE(t)* r, const FE(t)* a, const uint64_t e[4]) {
  FE(t) acc = FE(R1), base = *a;
  int started = 0;
  for (int i = 255; i >= 0; --i) {
    if (started) FE(sqr)(&acc, &acc);
    if ((e[i >> 6] >> (i & 63)) & 1) { if (started) FE(mul)(&acc, &acc, &base); else { acc = base; started = 1; } }
  }
  if (!started) acc = FE(R1);
  *r = acc;
}
static inline void FE(pow_u64)(FE(t)* r, const FE(t)* a, uint64_t e) { uint64_t ee[4] = {e, 0, 0, 0}; FE(pow)(r, a, ee); }

/* Fermat inversion; 0 -> 0 (callers that need Option semantics test is_zero first) */
static inline void FE(inv)(FE(t)* r, const FE(t)* a) {
  uint64_t e[4] = { FE_MOD0 - 2, FE_MOD1, FE_MOD2, FE_MOD3 };   /* MOD0 ends in ...0001 so no borrow */
  FE(pow)(r, a, e);
}

/* numeric compare on canonical values (pasta `Ord for Fp`), -1/0/1 */
static inline int FE(cmp)(const FE(t)* a, const FE(t)* b) {
  uint64_t x[4], y[4]; FE(to_raw)(x, a); FE(to_raw)(y, b);
  for (int i = 3; i >= 0; --i) { if (x[i] < y[i]) return -1; if (x[i] > y[i]) return 1; }
  return 0;
}

/* ff::BatchInvert: Montgomery trick, zero entries skipped (left zero). scratch: n elements */
static inline void FE(batch_invert)(FE(t)* v, size_t n, FE(t)* scratch) {
  FE(t) acc = FE(R1);
  for (size_t i = 0; i < n; ++i) { scratch[i] = acc; if (!FE(is_zero)(&v[i])) FE(mul)(&acc, &acc, &v[i]); }
  FE(t) ai; FE(inv)(&ai, &acc);
  for (size_t i = n; i-- > 0;) {
    if (FE(is_zero)(&v[i])) continue;
    FE(t) t; FE(mul)(&t, &ai, &scratch[i]);
    FE(mul)(&ai, &ai, &v[i]);
    v[i] = t;
  }
}

static void FE(init)(void) {
  /* R mod p = 2^256 - p * floor(2^256/p); since 2^255 < ... p ~ 2^254: compute by repeated doubling of 1 */
  FE(t) x = {{1, 0, 0, 0}};
  for (int i = 0; i < 256; ++i) FE(add)(&x, &x, &x);   /* x = 2^256 mod p (plain integers, add is mod p) */
  FE(R1) = x;
  for (int i = 0; i < 256; ++i) FE(add)(&x, &x, &x);   /* 2^512 mod p = R^2 */
  FE(R2c) = x;
  FE(mul)(&FE(R3c), &FE(R2c), &FE(R2c));               /* R2*R2/R = R^3 */
}
