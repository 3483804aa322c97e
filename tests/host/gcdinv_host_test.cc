// Host run of csrc/gcdinv.h (the same code the device would compile): stdin lines "<field> <a as 64 hex digits>",
// stdout a^-1 mod p as 64 hex digits.  field 0 = Pallas base (Fp), 1 = Pallas scalar (Fq).
#include "../../battlezips-halo2_b200/csrc/gcdinv.h"
#include <cstdio>
#include <cstring>
static const uint32_t MODS[2][8] = {
  {1u, 0x992d30edu, 0x094cf91bu, 0x224698fcu, 0u, 0u, 0u, 0x40000000u},
  {1u, 0x8c46eb21u, 0x0994a8ddu, 0x224698fcu, 0u, 0u, 0u, 0x40000000u}};
int main() {
  int f; char hex[80];
  while (scanf("%d %79s", &f, hex) == 2) {
    uint32_t a[8], p[8], r[8];
    if (strlen(hex) != 64) return 1;
    for (int i = 0; i < 8; ++i) { unsigned v; sscanf(hex + 8 * (7 - i), "%8x", &v); a[i] = v; p[i] = MODS[f][i]; }
    bz::gcdinv::inverse(r, a, p);
    for (int i = 7; i >= 0; --i) printf("%08x", r[i]);
    printf("\n");
  }
  return 0;
}
