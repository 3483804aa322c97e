// Square roots in the Pasta fields and the constants of pasta's hash-to-curve (shared by params.cu -- Params::new --
// and the verifier's point decompression in prover.cu).  U: pasta_curves 0.4.1 src/fields/{fp,fq}.rs `sqrt`,
// src/hashtocurve.rs, src/curves.rs.
#pragma once
#include "field.cuh"

namespace bz {

// ---- per-curve constants of the SWU map and the isogeny (coordinate field, Montgomery form) --------------------------
// iso-curve coefficients as published for the Pasta cycle; (x0, t, u) = kernel abscissa and Velu sums of the rational
// 3-isogeny, output rescaled by (1/9, 1/27).  Pinned by tests/test_gpu_params.py against the reference's KATs.
struct IsoConsts { uint32_t a[8], b[8], z[8], x0[8], t[8], u[8], inv9[8], inv27[8], rou[8], sqrt_exp[8]; };

template <class BP> struct Iso;
template <> struct Iso<FqP> {    // Vesta: coordinates in Fq
  static __host__ __device__ constexpr IsoConsts c() {
    return IsoConsts{
        {0xe5fa2060u, 0xe39dd73cu, 0x41bd984au, 0xa67a4eacu, 0x1c85040eu, 0x4e933438u, 0x203524b5u, 0x287658b7u},
        {0xffffec3du, 0xe28772dcu, 0xab3aedd4u, 0xa6dec34eu, 0xfffffd5au, 0xffffffffu, 0xffffffffu, 0x3fffffffu},
        {0x00000034u, 0x7e67c2b4u, 0xf2324d00u, 0xf6571331u, 0x00000006u, 0x00000000u, 0x00000000u, 0x00000000u},
        {0x06902433u, 0xbb4bd425u, 0x603d3ddbu, 0x45c1d742u, 0x350920fbu, 0x24d65648u, 0x22ea6761u, 0x181a7f51u},
        {0xfacba014u, 0x37584d8cu, 0xe19cd8c0u, 0x098423b8u, 0x05b43403u, 0x761d70d8u, 0x06710757u, 0x3b4ade8bu},
        {0xffffffb1u, 0xb61d70d0u, 0x0b1fe3a1u, 0x6c36ca39u, 0xfffffff5u, 0xffffffffu, 0xffffffffu, 0x3fffffffu},
        {0x71c71c72u, 0xad6517ceu, 0x3b04974du, 0xb24893c6u, 0xaaaaaaaau, 0xaaaaaaaau, 0xaaaaaaaau, 0x2aaaaaaau},
        {0x25ed097cu, 0x41fba4b0u, 0xc4b9f858u, 0xfcf1ec94u, 0x38e38e38u, 0x8e38e38eu, 0xe38e38e3u, 0x38e38e38u},
        {0x8c9942deu, 0x21807742u, 0x21b60494u, 0xcc495789u, 0xb2efbee2u, 0xac2e5d27u, 0x7f2db056u, 0x0b79fa89u},
        {0xc6237590u, 0x04ca546eu, 0x11234c7eu, 0x00000000u, 0x00000000u, 0x00000000u, 0x20000000u, 0x00000000u}};
  }
};
template <> struct Iso<FpP> {    // Pallas: coordinates in Fp
  static __host__ __device__ constexpr IsoConsts c() {
    return IsoConsts{
        {0x77bb08deu, 0x7fc5d290u, 0xcf122108u, 0x93090252u, 0xda1145bbu, 0x49f63ff5u, 0x7137f0dcu, 0x1c6d4f08u},
        {0xffffec3du, 0xf7f22478u, 0x33e1339bu, 0xa6dec354u, 0xfffffd5au, 0xffffffffu, 0xffffffffu, 0x3fffffffu},
        {0x00000034u, 0x1d2df024u, 0xe3a2999bu, 0xf6571331u, 0x00000006u, 0x00000000u, 0x00000000u, 0x00000000u},
        {0x1debddd5u, 0x0cc3fd72u, 0xf2897bb5u, 0xa08dbbc8u, 0x9ae2df0fu, 0x6594ee8eu, 0xa99b686fu, 0x1a12ef53u},
        {0xe4bf01c6u, 0x4cc12a1cu, 0x296a069bu, 0x1d6833aau, 0xf869dabfu, 0x75313ffdu, 0xe3719692u, 0x05af7634u},
        {0xffffffb1u, 0xbb0de6dcu, 0x213f207bu, 0x6c36ca39u, 0xfffffff5u, 0xffffffffu, 0xffffffffu, 0x3fffffffu},
        {0x1c71c71du, 0xc6e037a0u, 0xe8b8fc2bu, 0x130ac6c4u, 0x00000000u, 0x00000000u, 0x00000000u, 0x40000000u},
        {0xb425ed0au, 0xcaaf22d9u, 0x50aca717u, 0xbc707540u, 0xaaaaaaaau, 0xaaaaaaaau, 0xaaaaaaaau, 0x2aaaaaaau},
        {0xbad6dbf0u, 0xa28db849u, 0xd3b539dfu, 0x9083cd03u, 0x9dc8448eu, 0xfba6b9cau, 0x7b89c6dau, 0x3ec92874u},
        {0xcc969876u, 0x04a67c8du, 0x11234c7eu, 0x00000000u, 0x00000000u, 0x00000000u, 0x20000000u, 0x00000000u}};
  }
};
template <class P> __device__ __forceinline__ Fe<P> fe_const(const uint32_t (&l)[8]) {
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.l[i] = l[i];
  return r;
}

// ---- field helpers ---------------------------------------------------------------------------------------------------------
// a^e for a 256-bit exponent (most significant bit first)
template <class P> __device__ Fe<P> fe_pow_limbs(const Fe<P>& a, const uint32_t (&e)[8]) {
  Fe<P> acc = fe_one<P>();
  bool started = false;
  for (int i = 255; i >= 0; --i) {
    if (started) acc = fe_sqr(acc);
    if ((e[i >> 5] >> (i & 31)) & 1) { acc = started ? fe_mul(acc, a) : a; started = true; }
  }
  return acc;
}
// Tonelli-Shanks over the 2^32-smooth part (p - 1 = t * 2^32): returns false for a non-residue, else some root in r
// (the caller fixes the sign, so which of the two roots comes out does not matter)
template <class P> __device__ bool fe_sqrt(const Fe<P>& x, Fe<P>& r) {
  constexpr IsoConsts K = Iso<P>::c();
  if (fe_is_zero(x)) { r = x; return true; }
  const Fe<P> one = fe_one<P>();
  Fe<P> w = fe_pow_limbs(x, K.sqrt_exp);        // x^((t-1)/2)
  r = fe_mul(x, w);                              // x^((t+1)/2)
  Fe<P> tt = fe_mul(r, w);                       // x^t
  Fe<P> c = fe_const<P>(K.rou);
  uint32_t m = 32;
  while (!fe_eq(tt, one)) {
    uint32_t i = 0;
    Fe<P> t2 = tt;
    while (!fe_eq(t2, one)) { t2 = fe_sqr(t2); if (++i == m) return false; }
    Fe<P> b = c;
    for (uint32_t j = 0; j + i + 1 < m; ++j) b = fe_sqr(b);
    m = i;
    c = fe_sqr(b);
    tt = fe_mul(tt, c);
    r = fe_mul(r, b);
  }
  return true;
}
template <class P> __device__ __forceinline__ uint32_t fe_sgn0(const Fe<P>& a) { return fe_from_mont(a).l[0] & 1u; }

}  // namespace bz
