// Pasta curves y^2 = x^3 + 5 on the device (U: pasta_curves 0.4.1 curves.rs).
// ABI layouts: affine = x||y (64 B, Montgomery, identity = all zero), Jacobian = x||y||z (96 B,
// identity z = 0).  Accumulation inside kernels uses extended Jacobian "XYZZ" coordinates
// (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2): mixed addition 8M+2S and no inversion anywhere.
// Group results are only ever compared/serialised after to_affine, so any representative of the
// projective class is bit-exact with the reference.
#pragma once
#include "field.cuh"

namespace bz {

template <class BP> struct Affine { Fe<BP> x, y; };
template <class BP> struct Jac { Fe<BP> x, y, z; };
template <class BP> struct Xyzz { Fe<BP> x, y, zz, zzz; };

template <class BP> __device__ __forceinline__ bool aff_is_identity(const Affine<BP>& a) { return fe_is_zero(a.x) && fe_is_zero(a.y); }
template <class BP> __device__ __forceinline__ bool xyzz_is_identity(const Xyzz<BP>& a) { return fe_is_zero(a.zz); }

template <class BP> __device__ __forceinline__ Xyzz<BP> xyzz_identity() {
  Xyzz<BP> r; r.x = fe_zero<BP>(); r.y = fe_zero<BP>(); r.zz = fe_zero<BP>(); r.zzz = fe_zero<BP>(); return r;
}
template <class BP> __device__ __forceinline__ Xyzz<BP> xyzz_from_affine(const Affine<BP>& a) {
  Xyzz<BP> r;
  if (aff_is_identity(a)) return xyzz_identity<BP>();
  r.x = a.x; r.y = a.y; r.zz = fe_one<BP>(); r.zzz = fe_one<BP>();
  return r;
}

// dbl-2008-s-1 for a = 0
template <class BP> __device__ __noinline__ Xyzz<BP> xyzz_dbl(const Xyzz<BP>& p) {
  if (xyzz_is_identity(p)) return p;
  Fe<BP> u = fe_dbl(p.y);
  Fe<BP> v = fe_sqr(u);
  Fe<BP> w = fe_mul(u, v);
  Fe<BP> s = fe_mul(p.x, v);
  Fe<BP> xx = fe_sqr(p.x);
  Fe<BP> m = fe_add(fe_dbl(xx), xx);
  Xyzz<BP> r;
  r.x = fe_sub(fe_sqr(m), fe_dbl(s));
  r.y = fe_sub(fe_mul(m, fe_sub(s, r.x)), fe_mul(w, p.y));
  r.zz = fe_mul(v, p.zz);
  r.zzz = fe_mul(w, p.zzz);
  return r;
}

// mixed addition acc += q (affine), complete: handles identity / doubling / inverse (madd-2008-s)
template <class BP> __device__ __forceinline__ void xyzz_add_mixed(Xyzz<BP>& acc, const Affine<BP>& q) {
  if (aff_is_identity(q)) return;
  if (xyzz_is_identity(acc)) { acc = xyzz_from_affine(q); return; }
  Fe<BP> u2 = fe_mul(q.x, acc.zz);
  Fe<BP> s2 = fe_mul(q.y, acc.zzz);
  Fe<BP> p = fe_sub(u2, acc.x);
  Fe<BP> r = fe_sub(s2, acc.y);
  if (fe_is_zero(p)) {
    if (fe_is_zero(r)) { acc = xyzz_dbl(xyzz_from_affine(q)); }
    else acc = xyzz_identity<BP>();
    return;
  }
  Fe<BP> pp = fe_sqr(p);
  Fe<BP> ppp = fe_mul(p, pp);
  Fe<BP> qq = fe_mul(acc.x, pp);
  Fe<BP> x3 = fe_sub(fe_sub(fe_sqr(r), ppp), fe_dbl(qq));
  acc.y = fe_sub(fe_mul(r, fe_sub(qq, x3)), fe_mul(acc.y, ppp));
  acc.x = x3;
  acc.zz = fe_mul(acc.zz, pp);
  acc.zzz = fe_mul(acc.zzz, ppp);
}

// acc += sign ? -q : q
template <class BP> __device__ __forceinline__ void xyzz_add_mixed_signed(Xyzz<BP>& acc, Affine<BP> q, bool negate) {
  if (negate) q.y = fe_neg(q.y);    // identity (0,0) stays (0,0)
  xyzz_add_mixed(acc, q);
}

// full addition (add-2008-s), complete
template <class BP> __device__ __noinline__ Xyzz<BP> xyzz_add(const Xyzz<BP>& a, const Xyzz<BP>& b) {
  if (xyzz_is_identity(a)) return b;
  if (xyzz_is_identity(b)) return a;
  Fe<BP> u1 = fe_mul(a.x, b.zz);
  Fe<BP> u2 = fe_mul(b.x, a.zz);
  Fe<BP> s1 = fe_mul(a.y, b.zzz);
  Fe<BP> s2 = fe_mul(b.y, a.zzz);
  Fe<BP> p = fe_sub(u2, u1);
  Fe<BP> r = fe_sub(s2, s1);
  if (fe_is_zero(p)) {
    if (fe_is_zero(r)) return xyzz_dbl(a);
    return xyzz_identity<BP>();
  }
  Fe<BP> pp = fe_sqr(p);
  Fe<BP> ppp = fe_mul(p, pp);
  Fe<BP> qq = fe_mul(u1, pp);
  Xyzz<BP> o;
  o.x = fe_sub(fe_sub(fe_sqr(r), ppp), fe_dbl(qq));
  o.y = fe_sub(fe_mul(r, fe_sub(qq, o.x)), fe_mul(s1, ppp));
  o.zz = fe_mul(fe_mul(a.zz, b.zz), pp);
  o.zzz = fe_mul(fe_mul(a.zzz, b.zzz), ppp);
  return o;
}

template <class BP> __device__ __forceinline__ Xyzz<BP> xyzz_neg(const Xyzz<BP>& a) {
  Xyzz<BP> r = a; r.y = fe_neg(a.y); return r;
}

// XYZZ -> Jacobian without inversion: Z = ZZZ (= z^3)  =>  X' = X*ZZ^2, Y' = Y*ZZZ^2
template <class BP> __device__ __forceinline__ Jac<BP> xyzz_to_jac(const Xyzz<BP>& a) {
  Jac<BP> r;
  if (xyzz_is_identity(a)) { r.x = fe_zero<BP>(); r.y = fe_one<BP>(); r.z = fe_zero<BP>(); return r; }
  r.x = fe_mul(a.x, fe_sqr(a.zz));
  r.y = fe_mul(a.y, fe_sqr(a.zzz));
  r.z = a.zzz;
  return r;
}
template <class BP> __device__ __forceinline__ Xyzz<BP> jac_to_xyzz(const Jac<BP>& a) {
  Xyzz<BP> r;
  if (fe_is_zero(a.z)) return xyzz_identity<BP>();
  r.x = a.x; r.y = a.y; r.zz = fe_sqr(a.z); r.zzz = fe_mul(r.zz, a.z);
  return r;
}
// one inversion
template <class BP> __device__ __forceinline__ Affine<BP> xyzz_to_affine(const Xyzz<BP>& a) {
  Affine<BP> r;
  if (xyzz_is_identity(a)) { r.x = fe_zero<BP>(); r.y = fe_zero<BP>(); return r; }
  Fe<BP> zi = fe_inv_gcd(a.zzz);             // 1/z^3 (binary GCD: the callers run this on one lane per point)
  Fe<BP> zi2 = fe_mul(fe_sqr(zi), a.zz);     // z^-6 * z^2 ... = z^-4? no: see below
  // zz = z^2, zzz = z^3: 1/zz = zzz^-2 * zz^2 = z^-6 * z^4 = z^-2
  zi2 = fe_mul(zi2, a.zz);
  r.x = fe_mul(a.x, zi2);
  r.y = fe_mul(a.y, zi);
  return r;
}

// scalar multiple by a small non-negative integer (used by bucket-reduction chunk combine)
template <class BP> __device__ __noinline__ Xyzz<BP> xyzz_mul_u32(const Xyzz<BP>& p, uint32_t k) {
  Xyzz<BP> acc = xyzz_identity<BP>();
  for (int i = 31; i >= 0; --i) {
    acc = xyzz_dbl(acc);
    if ((k >> i) & 1) acc = xyzz_add(acc, p);
  }
  return acc;
}

// [k]P, k canonical 256-bit, double-and-add from the top bit
template <class BP> __device__ Xyzz<BP> xyzz_mul_scalar(const Xyzz<BP>& p, const uint32_t k[8]) {
  Xyzz<BP> acc = xyzz_identity<BP>();
  int top = 255;
  while (top >= 0 && !((k[top >> 5] >> (top & 31)) & 1)) --top;
  for (int i = top; i >= 0; --i) {
    acc = xyzz_dbl(acc);
    if ((k[i >> 5] >> (i & 31)) & 1) acc = xyzz_add(acc, p);
  }
  return acc;
}

template <class BP> __device__ __forceinline__ Affine<BP> aff_load(const Affine<BP>* p) {
  Affine<BP> r; r.x = fe_load(&p->x); r.y = fe_load(&p->y); return r;
}

}  // namespace bz
