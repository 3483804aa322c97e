// Montgomery multiplication for the Pasta fields as explicit carry chains (mad.lo.cc / madc.hi.cc), so that
// ptxas emits one IMAD.WIDE(.X) per 32x32 product with the accumulate and the carry folded in, instead of the
// IMAD.WIDE + IADD3 + IADD3.X + IMAD.X + MOV soup the 64-bit C formulation compiles to (field.cuh history:
// ~220 FMA-pipe + ~250 ALU-pipe instructions per multiplication).
//
// Operand scanning with two accumulator rows:  E[k] holds limb position k, O[k] holds position k+1, so the
// (lo, hi) halves of a product always land on adjacent registers of ONE row and a whole row is a single carry
// chain.  After every multiplier limb the Montgomery digit q = -E[0] (because -m^-1 = -1 mod 2^32) is folded in with
// the sparse modulus  m = 1 + M1*2^32 + M2*2^64 + M3*2^96 + 2^254 :  three real products, q itself, and q*2^30 as
// two shifts.  The one-limb right shift is free: the rows swap roles, the odd row is read two registers ahead, and
// the single stray limb (old E[1]) is merged by the add.cc that opens the next chain.
//
// Tried and dropped (round 1): making every chain-ending `addc` consume-and-produce a (mathematically zero) carry, so that
// ptxas must emit IADD3.X on the ALU pipe instead of IMAD.X on the fma-heavy pipe (30 -> 24 IMAD.X, heavy-pipe issue
// cycles per multiplication 217 -> 204).  It chains all rows into ONE dependency chain: Shot 3 310 -> 3 171 proofs/s.
// The independent chains below overlap; the pipe imbalance is the cheaper evil.
//
// Every primitive has a host emulation (explicit carry variable) so the exact instruction sequence is unit-tested
// on the CPU against big-integer arithmetic (tests/test_montmul_host.py) before it ever runs on a GPU.
#pragma once
#include <cstdint>

namespace bz {
namespace mm {

#ifdef __CUDA_ARCH__
#define BZ_MM_FN __device__ __forceinline__
BZ_MM_FN uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BZ_MM_FN uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BZ_MM_FN uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
BZ_MM_FN uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
BZ_MM_FN uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
BZ_MM_FN uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
BZ_MM_FN uint32_t xadd_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BZ_MM_FN uint32_t xaddc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BZ_MM_FN uint32_t xaddc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
#else
#define BZ_MM_FN inline
static thread_local uint32_t g_cc = 0;     // emulated carry flag
BZ_MM_FN uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
BZ_MM_FN uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
BZ_MM_FN uint32_t emu_add(uint64_t x, uint64_t c, bool use_cin, bool set_cc) {
  uint64_t s = x + c + (use_cin ? g_cc : 0);
  if (set_cc) g_cc = (uint32_t)(s >> 32);
  return (uint32_t)s;
}
BZ_MM_FN uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_lo(a, b), c, false, true); }
BZ_MM_FN uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_lo(a, b), c, true, true); }
BZ_MM_FN uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_hi(a, b), c, true, true); }
BZ_MM_FN uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_hi(a, b), c, true, false); }
BZ_MM_FN uint32_t xadd_cc(uint32_t a, uint32_t b) { return emu_add(a, b, false, true); }
BZ_MM_FN uint32_t xaddc_cc(uint32_t a, uint32_t b) { return emu_add(a, b, true, true); }
BZ_MM_FN uint32_t xaddc(uint32_t a, uint32_t b) { return emu_add(a, b, true, false); }
#endif

// One multiplier limb:  (E, O) <- ((E, O) + a * bi + q * m) / 2^32  with the roles of the rows swapped on exit
// (the caller alternates the arguments).  `first`: rows are empty.
template <uint32_t M1, uint32_t M2, uint32_t M3, bool FIRST>
BZ_MM_FN void mad_redc(uint32_t (&E)[8], uint32_t (&O)[8], const uint32_t (&a)[8], uint32_t bi) {
  if (FIRST) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      E[j] = mul_lo(a[j], bi); E[j + 1] = mul_hi(a[j], bi);
      O[j] = mul_lo(a[j + 1], bi); O[j + 1] = mul_hi(a[j + 1], bi);
    }
  } else {
    // stray limb: old-even position 1 now sits under E[0]; its carry opens the odd chain
    E[0] = xadd_cc(E[0], O[1]);
    // odd row, read two registers ahead (the one-limb shift of the previous reduction, for free)
    O[0] = madc_lo_cc(a[1], bi, O[2]); O[1] = madc_hi_cc(a[1], bi, O[3]);
    O[2] = madc_lo_cc(a[3], bi, O[4]); O[3] = madc_hi_cc(a[3], bi, O[5]);
    O[4] = madc_lo_cc(a[5], bi, O[6]); O[5] = madc_hi_cc(a[5], bi, O[7]);
    O[6] = madc_lo_cc(a[7], bi, 0u);   O[7] = madc_hi(a[7], bi, 0u);
    // even row
    E[0] = mad_lo_cc(a[0], bi, E[0]);  E[1] = madc_hi_cc(a[0], bi, E[1]);
    E[2] = madc_lo_cc(a[2], bi, E[2]); E[3] = madc_hi_cc(a[2], bi, E[3]);
    E[4] = madc_lo_cc(a[4], bi, E[4]); E[5] = madc_hi_cc(a[4], bi, E[5]);
    E[6] = madc_lo_cc(a[6], bi, E[6]); E[7] = madc_hi_cc(a[6], bi, E[7]);
    O[7] = xaddc(O[7], 0u);
  }
  // Montgomery digit and sparse reduction row
  const uint32_t q = 0u - E[0];
  // odd row: q*M1 at positions (1,2), q*M3 at (3,4), q*2^30 at (7,8)
  O[0] = mad_lo_cc(q, M1, O[0]);  O[1] = madc_hi_cc(q, M1, O[1]);
  O[2] = madc_lo_cc(q, M3, O[2]); O[3] = madc_hi_cc(q, M3, O[3]);
  O[4] = xaddc_cc(O[4], 0u);      O[5] = xaddc_cc(O[5], 0u);
  O[6] = xaddc_cc(O[6], q << 30); O[7] = xaddc(O[7], q >> 2);
  // even row: q*1 at (0,1), q*M2 at (2,3)
  E[0] = xadd_cc(E[0], q);        E[1] = xaddc_cc(E[1], 0u);
  E[2] = madc_lo_cc(q, M2, E[2]); E[3] = madc_hi_cc(q, M2, E[3]);
  E[4] = xaddc_cc(E[4], 0u);      E[5] = xaddc_cc(E[5], 0u);
  E[6] = xaddc_cc(E[6], 0u);      E[7] = xaddc_cc(E[7], 0u);
  O[7] = xaddc(O[7], 0u);
  // E[0] == 0 now
}

// r = a * b / 2^256 mod m, r < 2m (caller does the final conditional subtraction).  b < m; a < 2^256 - m
// (so that the running value fits 9 limbs); fully general 256-bit `a` (from_u512) uses field.cuh's fe_mul_c.
template <uint32_t M1, uint32_t M2, uint32_t M3>
BZ_MM_FN void mont_mul_wide(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
  uint32_t E[8], O[8];
  mad_redc<M1, M2, M3, true>(E, O, a, b[0]);
  mad_redc<M1, M2, M3, false>(O, E, a, b[1]);
  mad_redc<M1, M2, M3, false>(E, O, a, b[2]);
  mad_redc<M1, M2, M3, false>(O, E, a, b[3]);
  mad_redc<M1, M2, M3, false>(E, O, a, b[4]);
  mad_redc<M1, M2, M3, false>(O, E, a, b[5]);
  mad_redc<M1, M2, M3, false>(E, O, a, b[6]);
  mad_redc<M1, M2, M3, false>(O, E, a, b[7]);
  // last call had row-even = O, row-odd = E:  result limb k = E[k] + O[k+1]
  r[0] = xadd_cc(E[0], O[1]);
  r[1] = xaddc_cc(E[1], O[2]); r[2] = xaddc_cc(E[2], O[3]); r[3] = xaddc_cc(E[3], O[4]);
  r[4] = xaddc_cc(E[4], O[5]); r[5] = xaddc_cc(E[5], O[6]); r[6] = xaddc_cc(E[6], O[7]);
  r[7] = xaddc(E[7], 0u);
}

}  // namespace mm
}  // namespace bz
