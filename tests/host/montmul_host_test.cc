// Host harness for battlezips-halo2_b200/csrc/montmul.cuh: runs the EXACT carry-chain instruction sequence that
// ships in the CUDA kernels through the header's CPU emulation of mad.lo.cc / madc.hi.cc / addc.
// stdin: lines "field a_hex b_hex" (64 hex digits, canonical big-endian integers < 2^256); stdout: result hex (< 2m).
#include <cstdio>
#include <cstring>
#include <string>
#include <iostream>
#include "../../battlezips-halo2_b200/csrc/montmul.cuh"

static void parse(const std::string& h, uint32_t (&l)[8]) {
  for (int i = 0; i < 8; ++i) l[7 - i] = (uint32_t)std::stoul(h.substr(8 * i, 8), nullptr, 16);
}
int main() {
  int f; std::string a, b;
  while (std::cin >> f >> a >> b) {
    uint32_t x[8], y[8], r[8];
    parse(a, x); parse(b, y);
    if (f == 0) bz::mm::mont_mul_wide<0x992d30edu, 0x094cf91bu, 0x224698fcu>(r, x, y);
    else bz::mm::mont_mul_wide<0x8c46eb21u, 0x0994a8ddu, 0x224698fcu>(r, x, y);
    for (int i = 7; i >= 0; --i) printf("%08x", r[i]);
    printf("\n");
  }
}
