// Host-side Pasta field / curve helpers for the prover's control path (transcript challenges,
// rotation points, domain constants, blinds).  Same memory format as the device and as pasta:
// 4 x u64 little-endian Montgomery, R = 2^256.  Product code -- independent of oracle/.
// (U: pasta_curves 0.4.1 fields/{fp,fq}.rs semantics; constants SURVEY App. B.)
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace bzh {

typedef unsigned __int128 u128;

struct FieldConsts {
  uint64_t mod[4];
  uint64_t inv;            // -m^-1 mod 2^64
  uint64_t root_of_unity[4];   // canonical, 2^32-th primitive root  (5^t)
  uint64_t zeta[4];            // canonical primitive cube root named ZETA by pasta
  uint64_t delta[4];           // canonical 5^(2^32)
};

inline const FieldConsts& consts(int field) {
  static const FieldConsts fp = {
      {0x992d30ed00000001ULL, 0x224698fc094cf91bULL, 0x0ULL, 0x4000000000000000ULL},
      0x992d30ecffffffffULL,
      {0xbdad6fabd87ea32fULL, 0xea322bf2b7bb7584ULL, 0x362120830561f81aULL, 0x2bce74deac30ebdaULL},
      {0x1dad5ebdfdfe4ab9ULL, 0x1d1f8bd237ad3149ULL, 0x2caad5dc57aab1b0ULL, 0x12ccca834acdba71ULL},
      {0x6a6ccd20dd7b9ba2ULL, 0xf5e4f3f13eee5636ULL, 0xbd455b7112a5049dULL, 0x0a757d0f0006ab6cULL}};
  static const FieldConsts fq = {
      {0x8c46eb2100000001ULL, 0x224698fc0994a8ddULL, 0x0ULL, 0x4000000000000000ULL},
      0x8c46eb20ffffffffULL,
      {0xa70e2c1102b6d05fULL, 0x9bb97ea3c106f049ULL, 0x9e5c4dfd492ae26eULL, 0x2de6a9b8746d3f58ULL},
      {0x2aa9d2e050aa0e4fULL, 0x0fed467d47c033afULL, 0x511db4d81cf70f5aULL, 0x06819a58283e528eULL},
      {0x8494392472d1683cULL, 0xe3ac3376541d1140ULL, 0x06f0a88e7f7949f8ULL, 0x2237d54423724166ULL}};
  return field == 0 ? fp : fq;
}

struct Fe {
  uint64_t l[4];
  bool operator==(const Fe& o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }
  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
};

// A field "context": all ops are methods so Fp / Fq share the code.
struct Field {
  int id;
  const FieldConsts* c;
  Fe R1, R2, R3;
  explicit Field(int field) : id(field), c(&consts(field)) {
    Fe x = {{1, 0, 0, 0}};
    for (int i = 0; i < 256; ++i) x = add(x, x);
    R1 = x;
    for (int i = 0; i < 256; ++i) x = add(x, x);
    R2 = x;
    R3 = mul(R2, R2);
  }
  bool geq_mod(const uint64_t a[4]) const {
    for (int i = 3; i >= 0; --i) { if (a[i] > c->mod[i]) return true; if (a[i] < c->mod[i]) return false; }
    return true;
  }
  Fe add(const Fe& a, const Fe& b) const {
    Fe r; u128 cy = 0;
    for (int i = 0; i < 4; ++i) { cy += (u128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)cy; cy >>= 64; }
    if (geq_mod(r.l)) sub_mod(r.l);
    return r;
  }
  void sub_mod(uint64_t a[4]) const {
    u128 br = 0;
    for (int i = 0; i < 4; ++i) { u128 d = (u128)a[i] - c->mod[i] - (uint64_t)br; a[i] = (uint64_t)d; br = (d >> 64) & 1; }
  }
  Fe sub(const Fe& a, const Fe& b) const {
    Fe r; u128 br = 0;
    for (int i = 0; i < 4; ++i) { u128 d = (u128)a.l[i] - b.l[i] - (uint64_t)br; r.l[i] = (uint64_t)d; br = (d >> 64) & 1; }
    if (br) { u128 cy = 0; for (int i = 0; i < 4; ++i) { cy += (u128)r.l[i] + c->mod[i]; r.l[i] = (uint64_t)cy; cy >>= 64; } }
    return r;
  }
  Fe neg(const Fe& a) const { Fe z = {{0, 0, 0, 0}}; return sub(z, a); }
  Fe mul(const Fe& a, const Fe& b) const {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
      u128 acc; uint64_t cy = 0;
      for (int j = 0; j < 4; ++j) { acc = (u128)a.l[j] * b.l[i] + t[j] + cy; t[j] = (uint64_t)acc; cy = (uint64_t)(acc >> 64); }
      acc = (u128)t[4] + cy; t[4] = (uint64_t)acc; t[5] = (uint64_t)(acc >> 64);
      uint64_t m = t[0] * c->inv;
      acc = (u128)m * c->mod[0] + t[0]; cy = (uint64_t)(acc >> 64);
      for (int j = 1; j < 4; ++j) { acc = (u128)m * c->mod[j] + t[j] + cy; t[j - 1] = (uint64_t)acc; cy = (uint64_t)(acc >> 64); }
      acc = (u128)t[4] + cy; t[3] = (uint64_t)acc; t[4] = t[5] + (uint64_t)(acc >> 64);
    }
    Fe r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq_mod(r.l)) sub_mod(r.l);
    return r;
  }
  Fe sqr(const Fe& a) const { return mul(a, a); }
  Fe one() const { return R1; }
  Fe zero() const { return Fe{{0, 0, 0, 0}}; }
  Fe from_raw(const uint64_t v[4]) const { Fe a; memcpy(a.l, v, 32); return mul(a, R2); }
  Fe from_u64(uint64_t v) const { uint64_t a[4] = {v, 0, 0, 0}; return from_raw(a); }
  void to_raw(const Fe& a, uint64_t out[4]) const { Fe one_raw = {{1, 0, 0, 0}}; Fe r = mul(a, one_raw); memcpy(out, r.l, 32); }
  void to_repr(const Fe& a, uint8_t out[32]) const { uint64_t v[4]; to_raw(a, v); memcpy(out, v, 32); }
  Fe from_bytes_wide(const uint8_t b[64]) const {
    Fe lo, hi; memcpy(lo.l, b, 32); memcpy(hi.l, b + 32, 32);
    return add(mul(lo, R2), mul(hi, R3));
  }
  Fe pow(const Fe& a, const uint64_t e[4]) const {
    Fe acc = R1;
    for (int i = 255; i >= 0; --i) { acc = sqr(acc); if ((e[i >> 6] >> (i & 63)) & 1) acc = mul(acc, a); }
    return acc;
  }
  Fe pow_u64(const Fe& a, uint64_t e) const { uint64_t ee[4] = {e, 0, 0, 0}; return pow(a, ee); }
  Fe inv(const Fe& a) const {
    uint64_t e[4] = {c->mod[0] - 2, c->mod[1], c->mod[2], c->mod[3]};
    return pow(a, e);
  }
  Fe root_of_unity() const { return from_raw(c->root_of_unity); }
  Fe zeta() const { return from_raw(c->zeta); }
  Fe delta() const { return from_raw(c->delta); }
  // numeric compare of canonical values (pasta `Ord`)
  int cmp(const Fe& a, const Fe& b) const {
    uint64_t x[4], y[4]; to_raw(a, x); to_raw(b, y);
    for (int i = 3; i >= 0; --i) { if (x[i] < y[i]) return -1; if (x[i] > y[i]) return 1; }
    return 0;
  }
};

// ---- y^2 = x^3 + 5 in XYZZ coordinates on the host (final window Horner of the bucket MSM: 255 sequential
// doublings are latency-bound on one GPU thread (~1.1 ms) and ~0.1 ms on a CPU core) ------------------------------
struct HXyzz { Fe x, y, zz, zzz; };
inline bool hx_is_identity(const HXyzz& a) { return a.zz.is_zero(); }
inline HXyzz hx_identity() { HXyzz r; memset(&r, 0, sizeof(r)); return r; }
inline HXyzz hx_dbl(const Field& F, const HXyzz& p) {
  if (hx_is_identity(p)) return p;
  Fe u = F.add(p.y, p.y), v = F.sqr(u), w = F.mul(u, v), s = F.mul(p.x, v), xx = F.sqr(p.x);
  Fe m = F.add(F.add(xx, xx), xx);
  HXyzz r;
  r.x = F.sub(F.sqr(m), F.add(s, s));
  r.y = F.sub(F.mul(m, F.sub(s, r.x)), F.mul(w, p.y));
  r.zz = F.mul(v, p.zz);
  r.zzz = F.mul(w, p.zzz);
  return r;
}
inline HXyzz hx_add(const Field& F, const HXyzz& a, const HXyzz& b) {
  if (hx_is_identity(a)) return b;
  if (hx_is_identity(b)) return a;
  Fe u1 = F.mul(a.x, b.zz), u2 = F.mul(b.x, a.zz), s1 = F.mul(a.y, b.zzz), s2 = F.mul(b.y, a.zzz);
  Fe p = F.sub(u2, u1), r = F.sub(s2, s1);
  if (p.is_zero()) return r.is_zero() ? hx_dbl(F, a) : hx_identity();
  Fe pp = F.sqr(p), ppp = F.mul(p, pp), q = F.mul(u1, pp);
  HXyzz o;
  o.x = F.sub(F.sub(F.sqr(r), ppp), F.add(q, q));
  o.y = F.sub(F.mul(r, F.sub(q, o.x)), F.mul(s1, ppp));
  o.zz = F.mul(F.mul(a.zz, b.zz), pp);
  o.zzz = F.mul(F.mul(a.zzz, b.zzz), ppp);
  return o;
}
// XYZZ -> Jacobian (x, y, z) with z = zzz:  X' = X * ZZ^2, Y' = Y * ZZZ^2   (identity: (0, 1, 0))
inline void hx_to_jac(const Field& F, const HXyzz& a, Fe out[3]) {
  if (hx_is_identity(a)) { out[0] = F.zero(); out[1] = F.one(); out[2] = F.zero(); return; }
  out[0] = F.mul(a.x, F.sqr(a.zz));
  out[1] = F.mul(a.y, F.sqr(a.zzz));
  out[2] = a.zzz;
}

}  // namespace bzh
