// BLAKE2b (RFC 7693) with personalisation, for the Fiat-Shamir transcript on the host
// (U: halo2_proofs 0.2.0 src/transcript.rs uses blake2b_simd with personal "Halo2-Transcript", 64-byte digest).
// Supports update / clone / finalize, which is all Blake2bWrite::squeeze_challenge needs.
#pragma once
#include <cstdint>
#include <cstring>

namespace bzh {

struct Blake2b {
  uint64_t h[8];
  uint64_t t0 = 0, t1 = 0;
  uint8_t buf[128];
  size_t buflen = 0;
  size_t outlen = 64;

  static inline uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
  static const uint64_t* iv() {
    static const uint64_t v[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                  0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    return v;
  }

  explicit Blake2b(const char personal[16], size_t digest_len = 64) {
    outlen = digest_len;
    uint8_t param[64] = {0};
    param[0] = (uint8_t)digest_len;   // digest length
    param[1] = 0;                     // key length
    param[2] = 1;                     // fanout
    param[3] = 1;                     // depth
    memcpy(param + 48, personal, 16);
    for (int i = 0; i < 8; ++i) { uint64_t w; memcpy(&w, param + 8 * i, 8); h[i] = iv()[i] ^ w; }
    memset(buf, 0, sizeof(buf));
  }

  void compress(const uint8_t block[128], bool last) {
    static const uint8_t sigma[12][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
    uint64_t m[16], v[16];
    for (int i = 0; i < 16; ++i) memcpy(&m[i], block + 8 * i, 8);
    for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[i + 8] = iv()[i]; }
    v[12] ^= t0; v[13] ^= t1;
    if (last) v[14] = ~v[14];
#define BZ_G(a, b, c, d, x, y)                                   \
  v[a] = v[a] + v[b] + (x); v[d] = rotr(v[d] ^ v[a], 32);        \
  v[c] = v[c] + v[d];       v[b] = rotr(v[b] ^ v[c], 24);        \
  v[a] = v[a] + v[b] + (y); v[d] = rotr(v[d] ^ v[a], 16);        \
  v[c] = v[c] + v[d];       v[b] = rotr(v[b] ^ v[c], 63);
    for (int r = 0; r < 12; ++r) {
      const uint8_t* s = sigma[r];
      BZ_G(0, 4, 8, 12, m[s[0]], m[s[1]]) BZ_G(1, 5, 9, 13, m[s[2]], m[s[3]])
      BZ_G(2, 6, 10, 14, m[s[4]], m[s[5]]) BZ_G(3, 7, 11, 15, m[s[6]], m[s[7]])
      BZ_G(0, 5, 10, 15, m[s[8]], m[s[9]]) BZ_G(1, 6, 11, 12, m[s[10]], m[s[11]])
      BZ_G(2, 7, 8, 13, m[s[12]], m[s[13]]) BZ_G(3, 4, 9, 14, m[s[14]], m[s[15]])
    }
#undef BZ_G
    for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
  }

  void update(const void* data, size_t len) {
    const uint8_t* p = (const uint8_t*)data;
    while (len > 0) {
      if (buflen == 128) {              // buffer full and more input follows: compress it (not the last block)
        t0 += 128; if (t0 < 128) ++t1;
        compress(buf, false);
        buflen = 0;
      }
      size_t take = 128 - buflen; if (take > len) take = len;
      memcpy(buf + buflen, p, take);
      buflen += take; p += take; len -= take;
    }
  }

  // non-destructive finalize (works on a copy, like `state.clone().finalize()`)
  void finalize(uint8_t* out) const {
    Blake2b c = *this;
    c.t0 += c.buflen; if (c.t0 < c.buflen) ++c.t1;
    memset(c.buf + c.buflen, 0, 128 - c.buflen);
    c.compress(c.buf, true);
    memcpy(out, c.h, outlen);
  }
};

}  // namespace bzh
