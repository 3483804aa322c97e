// `Params::new(k)` on the device (U: halo2_proofs 0.2.0 src/poly/commitment.rs `Params::new`,
// pasta_curves 0.4.1 src/hashtocurve.rs + curves.rs `hash_to_curve`; reference call sites
// /root/reference/benches/shot.rs:58, benches/board.rs:51, src/circuits/shot.rs:915, src/circuits/board.rs:907):
//
//   g[i]       = hash_to_curve("Halo2-Parameters")(0x00 || i as u32 LE)          i < 2^k
//   g_lagrange = best_fft over GROUP elements of g with omega^-1, every point scaled by 2^-k   (EC inverse FFT)
//   w, u       = hasher(0x01), hasher(0x02)
//
// hash_to_curve = BLAKE2b-XMD hash_to_field (two field elements) -> simplified SWU on the iso-curve
// y^2 = x^3 + A x + 1265 (Z = -13) twice -> affine addition on the iso-curve -> 3-isogeny onto y^2 = x^3 + 5.
// One thread per point; the ~2^k * k / 2 scalar multiplications of the group FFT run one butterfly per thread.
// The same entry point (`bz_hash_to_curve`) reproduces the reference's Pallas generator KATs
// (/root/reference/src/utils/constants/fixed_bases/board_commit_v.rs:5-14, board_commit_r.rs:5-14).
#include "../../include/bzhalo2.h"
#include "common.h"
#include "curve.cuh"
#include "sqrt.cuh"
#include <cstring>

namespace bz {

// ---- BLAKE2b-512, unkeyed, all-zero personalisation (what pasta's hash_to_field instantiates) --------------------------
__device__ __forceinline__ uint64_t rotr64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
__constant__ uint8_t c_sigma[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
__constant__ uint64_t c_iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};

__device__ void blake2b_compress(uint64_t h[8], const uint64_t m[16], uint64_t t, bool last) {
  uint64_t v[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[i + 8] = c_iv[i]; }
  v[12] ^= t;
  if (last) v[14] = ~v[14];
#define BZ_G(a, b, c, d, x, y)                                   \
  v[a] = v[a] + v[b] + (x); v[d] = rotr64(v[d] ^ v[a], 32);      \
  v[c] = v[c] + v[d];       v[b] = rotr64(v[b] ^ v[c], 24);      \
  v[a] = v[a] + v[b] + (y); v[d] = rotr64(v[d] ^ v[a], 16);      \
  v[c] = v[c] + v[d];       v[b] = rotr64(v[b] ^ v[c], 63);
  for (int r = 0; r < 12; ++r) {
    const uint8_t* s = c_sigma[r];
    BZ_G(0, 4, 8, 12, m[s[0]], m[s[1]]) BZ_G(1, 5, 9, 13, m[s[2]], m[s[3]])
    BZ_G(2, 6, 10, 14, m[s[4]], m[s[5]]) BZ_G(3, 7, 11, 15, m[s[6]], m[s[7]])
    BZ_G(0, 5, 10, 15, m[s[8]], m[s[9]]) BZ_G(1, 6, 11, 12, m[s[10]], m[s[11]])
    BZ_G(2, 7, 8, 13, m[s[12]], m[s[13]]) BZ_G(3, 4, 9, 14, m[s[14]], m[s[15]])
  }
#undef BZ_G
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
}
// digest (64 B) of `len` bytes at `data` (local memory, zero padded to a multiple of 128 by the caller)
__device__ void blake2b_512(const uint8_t* data, uint32_t len, uint8_t out[64]) {
  uint64_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = c_iv[i];
  h[0] ^= 0x01010040ULL;                 // digest 64, no key, fanout 1, depth 1; salt and personal all zero
  uint32_t off = 0;
  uint64_t m[16];
  while (true) {
    const bool last = len - off <= 128;
    for (int i = 0; i < 16; ++i) {
      uint64_t w = 0;
      for (int j = 7; j >= 0; --j) w = (w << 8) | data[off + 8 * i + j];
      m[i] = w;
    }
    blake2b_compress(h, m, last ? len : off + 128, last);
    if (last) break;
    off += 128;
  }
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) out[8 * i + j] = (uint8_t)(h[i] >> (8 * j));
}

// simplified SWU on y^2 = x^3 + A x + B with Z = -13 (U: hashtocurve.rs map_to_curve_simple_swu), affine out
template <class P> __device__ void swu_map(const Fe<P>& u, Fe<P>& x, Fe<P>& y) {
  constexpr IsoConsts K = Iso<P>::c();
  const Fe<P> a = fe_const<P>(K.a), b = fe_const<P>(K.b), z = fe_const<P>(K.z);
  Fe<P> z_u2 = fe_mul(z, fe_sqr(u));
  Fe<P> ta = fe_add(fe_sqr(z_u2), z_u2);
  Fe<P> num_x1 = fe_mul(b, fe_add(ta, fe_one<P>()));
  Fe<P> div = fe_mul(a, fe_is_zero(ta) ? z : fe_neg(ta));
  Fe<P> x1 = fe_mul(num_x1, fe_inv(div));
  Fe<P> gx1 = fe_add(fe_mul(fe_add(fe_sqr(x1), a), x1), b);
  Fe<P> y1;
  if (fe_sqrt(gx1, y1)) { x = x1; y = y1; }
  else {
    x = fe_mul(z_u2, x1);
    Fe<P> gx2 = fe_add(fe_mul(fe_add(fe_sqr(x), a), x), b);
    fe_sqrt(gx2, y);
  }
  if (fe_sgn0(u) != fe_sgn0(y)) y = fe_neg(y);
}

// hash_to_curve for one message.  `tmpl`: 128 zero bytes || <msg_len message bytes> || 0x00 0x80 0x00 || tail ;
// tail = domain "-" curve_id "_XMD:BLAKE2b_SSWU_RO_" len  (built on the host; the message bytes are patched in here)
constexpr uint32_t H2C_MAX = 512;
template <class BP>
__global__ void hash_to_curve_kernel(const uint8_t* __restrict__ tmpl, uint32_t tmpl_len, uint32_t msg_len, uint32_t tail_len,
                                     const uint8_t* __restrict__ messages, uint32_t index_msgs /* message = prefix byte + LE32 of (first_index + i) */,
                                     uint32_t first_index, uint8_t prefix, Affine<BP>* __restrict__ out, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  constexpr IsoConsts K = Iso<BP>::c();
  uint8_t buf[H2C_MAX];
  const uint32_t padded = (tmpl_len + 127u) & ~127u;
  for (uint32_t j = 0; j < padded; ++j) buf[j] = j < tmpl_len ? tmpl[j] : 0;
  if (index_msgs) {
    const uint32_t v = first_index + i;
    buf[128] = prefix; buf[129] = (uint8_t)v; buf[130] = (uint8_t)(v >> 8); buf[131] = (uint8_t)(v >> 16); buf[132] = (uint8_t)(v >> 24);
  } else {
    for (uint32_t j = 0; j < msg_len; ++j) buf[128 + j] = messages[(size_t)i * msg_len + j];
  }
  uint8_t b0[64], b1[64], b2[64];
  blake2b_512(buf, tmpl_len, b0);
  const uint8_t* tail = tmpl + (tmpl_len - tail_len);
  // b1 = H(b0 || 0x01 || tail), b2 = H((b0 ^ b1) || 0x02 || tail)
  const uint32_t l2 = 64 + 1 + tail_len, p2 = (l2 + 127u) & ~127u;
  for (uint32_t j = 0; j < p2; ++j) buf[j] = j < 64 ? b0[j] : j == 64 ? 1 : j < l2 ? tail[j - 65] : 0;
  blake2b_512(buf, l2, b1);
  for (uint32_t j = 0; j < 64; ++j) buf[j] = b0[j] ^ b1[j];
  buf[64] = 2;
  blake2b_512(buf, l2, b2);
  // field elements: 64 bytes BIG-endian mod p
  Fe<BP> u[2];
  for (int e = 0; e < 2; ++e) {
    const uint8_t* bb = e == 0 ? b1 : b2;
    uint32_t w[16];
    for (int j = 0; j < 16; ++j) {
      const uint8_t* q = bb + 60 - 4 * j;
      w[j] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
    }
    u[e] = fe_from_u512<BP>(w);
  }
  Fe<BP> x1, y1, x2, y2;
  swu_map<BP>(u[0], x1, y1);
  swu_map<BP>(u[1], x2, y2);
  // affine addition on the iso-curve (A != 0: doubling uses 3x^2 + A)
  Affine<BP> res; res.x = fe_zero<BP>(); res.y = fe_zero<BP>();
  bool inf = false;
  Fe<BP> lam;
  if (fe_eq(x1, x2)) {
    if (fe_is_zero(fe_add(y1, y2))) inf = true;
    else { Fe<BP> xx = fe_sqr(x1); lam = fe_mul(fe_add(fe_add(fe_dbl(xx), xx), fe_const<BP>(K.a)), fe_inv(fe_dbl(y1))); }
  } else lam = fe_mul(fe_sub(y2, y1), fe_inv(fe_sub(x2, x1)));
  if (!inf) {
    Fe<BP> x3 = fe_sub(fe_sub(fe_sqr(lam), x1), x2);
    Fe<BP> y3 = fe_sub(fe_mul(lam, fe_sub(x1, x3)), y1);
    // 3-isogeny (Velu) and rescale by (1/9, 1/27)
    Fe<BP> d = fe_sub(x3, fe_const<BP>(K.x0));
    if (!fe_is_zero(d)) {
      Fe<BP> di = fe_inv(d), di2 = fe_sqr(di), di3 = fe_mul(di2, di);
      const Fe<BP> t = fe_const<BP>(K.t), uu = fe_const<BP>(K.u);
      Fe<BP> X = fe_add(fe_add(x3, fe_mul(t, di)), fe_mul(uu, di2));
      Fe<BP> Y = fe_mul(y3, fe_sub(fe_sub(fe_one<BP>(), fe_mul(t, di2)), fe_mul(fe_dbl(uu), di3)));
      res.x = fe_mul(X, fe_const<BP>(K.inv9));
      res.y = fe_mul(Y, fe_const<BP>(K.inv27));
    }
  }
  fe_store(&out[i].x, res.x); fe_store(&out[i].y, res.y);
}

// ---- group FFT ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bitrev(uint32_t x, uint32_t bits) { return bits ? __brev(x) >> (32 - bits) : 0u; }

template <class BP> __device__ __forceinline__ Xyzz<BP> xyzz_load(const Xyzz<BP>* q) {
  Xyzz<BP> v; v.x = fe_load(&q->x); v.y = fe_load(&q->y); v.zz = fe_load(&q->zz); v.zzz = fe_load(&q->zzz); return v;
}
template <class BP> __device__ __forceinline__ void xyzz_store(Xyzz<BP>* q, const Xyzz<BP>& v) {
  fe_store(&q->x, v.x); fe_store(&q->y, v.y); fe_store(&q->zz, v.zz); fe_store(&q->zzz, v.zzz);
}
// a[bitrev(i)] = g[i] as XYZZ
template <class BP>
__global__ void ecfft_load_kernel(const Affine<BP>* __restrict__ g, Xyzz<BP>* __restrict__ a, uint32_t logn) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (1u << logn)) return;
  xyzz_store(a + bitrev(i, logn), xyzz_from_affine(aff_load(g + i)));
}
// twiddles (canonical scalars): tw[i] = omega^i, i < n/2
template <class SP>
__global__ void ecfft_twiddle_kernel(Fe<SP>* __restrict__ tw, Fe<SP> omega, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  fe_store(tw + i, fe_from_mont(fe_pow_u64<SP>(omega, i)));
}
// one radix-2 stage (decimation in time, input bit-reversed): half = 2^(s-1); butterfly (lo, hi = lo + half) with twiddle
// omega^(l * n / (2 half))
template <class BP, class SP>
__global__ void ecfft_stage_kernel(Xyzz<BP>* __restrict__ a, const Fe<SP>* __restrict__ tw, uint32_t logn, uint32_t s) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = 1u << logn, half = 1u << (s - 1);
  if (t >= (n >> 1)) return;
  const uint32_t l = t & (half - 1), lo = ((t >> (s - 1)) << s) | l, hi = lo + half;
  Xyzz<BP> x = xyzz_load(a + lo), y = xyzz_load(a + hi);
  if (l) {
    Fe<SP> w = fe_load(tw + (size_t)l * (n >> s));
    y = xyzz_mul_scalar(y, w.l);
  }
  xyzz_store(a + lo, xyzz_add(x, y));
  xyzz_store(a + hi, xyzz_add(x, xyzz_neg(y)));
}
// every point times one scalar (2^-k), in place
template <class BP>
__global__ void ec_scale_kernel(Xyzz<BP>* __restrict__ a, const uint32_t* __restrict__ k8, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  uint32_t k[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) k[j] = k8[j];
  xyzz_store(a + i, xyzz_mul_scalar(xyzz_load(a + i), k));
}
// XYZZ -> affine with Montgomery's trick over RUN consecutive entries per thread
template <class BP, int RUN>
__global__ void xyzz_normalize_kernel(const Xyzz<BP>* __restrict__ in, Affine<BP>* __restrict__ out, uint64_t total) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t lo = t * RUN;
  if (lo >= total) return;
  uint32_t cnt = (uint32_t)min((uint64_t)RUN, total - lo);
  Fe<BP> pre[RUN];
  Fe<BP> acc = fe_one<BP>();
  for (uint32_t j = 0; j < cnt; ++j) {
    pre[j] = acc;
    Fe<BP> z = fe_load(&in[lo + j].zzz);
    if (!fe_is_zero(z)) acc = fe_mul(acc, z);
  }
  Fe<BP> inv = fe_inv(acc);
  for (int j = (int)cnt - 1; j >= 0; --j) {
    const Xyzz<BP>* q = in + lo + j;
    Fe<BP> zzz = fe_load(&q->zzz);
    Affine<BP> r;
    if (fe_is_zero(zzz)) { r.x = fe_zero<BP>(); r.y = fe_zero<BP>(); }
    else {
      Fe<BP> zi3 = fe_mul(inv, pre[j]);
      inv = fe_mul(inv, zzz);
      Fe<BP> zz = fe_load(&q->zz);
      Fe<BP> zi2 = fe_mul(fe_mul(fe_sqr(zi3), zz), zz);
      r.x = fe_mul(fe_load(&q->x), zi2);
      r.y = fe_mul(fe_load(&q->y), zi3);
    }
    fe_store(&out[lo + j].x, r.x); fe_store(&out[lo + j].y, r.y);
  }
}

// ---- point (de)compression: `Params::write / read`, proof points -------------------------------------------------------
// compressed point (32 B: x little-endian, bit 255 = parity of y) -> affine Montgomery.  status: 0 ok, 1 identity
// encoding (all zero), 2 invalid (x >= p or x^3 + 5 not a square)
template <class BP>
__global__ void decompress_points_kernel(const uint8_t* __restrict__ in, Affine<BP>* __restrict__ out, uint8_t* __restrict__ status, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(in + (size_t)i * 32);
  Fe<BP> x;
  uint32_t any = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { x.l[j] = w[j]; any |= w[j]; }
  const uint32_t sign = x.l[7] >> 31;
  x.l[7] &= 0x7fffffffu;
  Affine<BP> r; r.x = fe_zero<BP>(); r.y = fe_zero<BP>();
  uint8_t st = 0;
  if (!any) st = 1;
  else {
    // canonical check: x < p
    uint32_t t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = x.l[j];
    sub_cc(t[0], mod_limb<BP>(0));
#pragma unroll
    for (int j = 1; j < 8; ++j) subc_cc(t[j], mod_limb<BP>(j));
    const bool below = subc(0u, 0u) != 0u;
    if (!below) st = 2;
    else {
      Fe<BP> xm = fe_to_mont(x);
      Fe<BP> five = fe_zero<BP>(); five.l[0] = 5; five = fe_to_mont(five);
      Fe<BP> y2 = fe_add(fe_mul(fe_sqr(xm), xm), five), y;
      if (!fe_sqrt(y2, y)) st = 2;
      else {
        if (fe_sgn0(y) != sign) y = fe_neg(y);
        if (fe_sgn0(y) != sign) st = 2;             // y = 0 with the sign bit set
        r.x = xm; r.y = y;
      }
    }
  }
  fe_store(&out[i].x, r.x); fe_store(&out[i].y, r.y);
  status[i] = st;
}


// affine Montgomery -> compressed (pasta `to_bytes`): x canonical little-endian, bit 255 = parity of y; identity = zeros
template <class BP>
__global__ void compress_points_kernel(const Affine<BP>* __restrict__ in, uint8_t* __restrict__ out, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Affine<BP> p = aff_load(in + i);
  uint32_t* w = reinterpret_cast<uint32_t*>(out + (size_t)i * 32);
  if (aff_is_identity(p)) {
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = 0;
    return;
  }
  Fe<BP> x = fe_from_mont(p.x);
  x.l[7] |= fe_sgn0(p.y) << 31;
#pragma unroll
  for (int j = 0; j < 8; ++j) w[j] = x.l[j];
}

void decompress_points_run(Ctx* ctx, int curve, const void* d_in, void* d_out_affine, uint8_t* d_status, uint32_t count) {
  if (!count) return;
  if (curve == 0) decompress_points_kernel<FqP><<<(count + 63) / 64, 64, 0, ctx->stream>>>((const uint8_t*)d_in, (Affine<FqP>*)d_out_affine, d_status, count);
  else decompress_points_kernel<FpP><<<(count + 63) / 64, 64, 0, ctx->stream>>>((const uint8_t*)d_in, (Affine<FpP>*)d_out_affine, d_status, count);
  ctx->kernel_launches++;
  BZ_CUDA(cudaGetLastError());
}
void compress_points_run(Ctx* ctx, int curve, const void* d_in_affine, void* d_out, uint32_t count) {
  if (!count) return;
  if (curve == 0) compress_points_kernel<FqP><<<(count + 127) / 128, 128, 0, ctx->stream>>>((const Affine<FqP>*)d_in_affine, (uint8_t*)d_out, count);
  else compress_points_kernel<FpP><<<(count + 127) / 128, 128, 0, ctx->stream>>>((const Affine<FpP>*)d_in_affine, (uint8_t*)d_out, count);
  ctx->kernel_launches++;
  BZ_CUDA(cudaGetLastError());
}

// ---- host drivers ----------------------------------------------------------------------------------------------------------------
static std::vector<uint8_t> h2c_template(int curve, const char* domain, uint32_t msg_len, uint32_t& tail_len) {
  const std::string cid = curve == 0 ? "vesta" : "pallas", dom(domain);
  BZ_CHECK(dom.size() < 256 && 22 + cid.size() + dom.size() < 256, "hash_to_curve: domain prefix too long");
  std::string tail = dom + "-" + cid + "_XMD:BLAKE2b_SSWU_RO_";
  tail.push_back((char)(22 + cid.size() + dom.size()));
  tail_len = (uint32_t)tail.size();
  std::vector<uint8_t> t(128 + msg_len, 0);
  t.push_back(0); t.push_back(128); t.push_back(0);
  t.insert(t.end(), tail.begin(), tail.end());
  BZ_CHECK(t.size() <= H2C_MAX - 128 && 65 + tail_len <= H2C_MAX - 128, "hash_to_curve: message too long");
  return t;
}

template <class BP>
static void hash_to_curve_t(Ctx* ctx, int curve, const char* domain, const uint8_t* d_messages, uint32_t msg_len, bool index_msgs,
                            uint32_t first_index, uint8_t prefix, void* d_out, uint32_t count) {
  uint32_t tail_len = 0;
  std::vector<uint8_t> t = h2c_template(curve, domain, msg_len, tail_len);
  DevBuf d_t; d_t.alloc(t.size());
  BZ_CUDA(cudaMemcpyAsync(d_t.p, t.data(), t.size(), cudaMemcpyHostToDevice, ctx->stream));
  hash_to_curve_kernel<BP><<<(count + 63) / 64, 64, 0, ctx->stream>>>(d_t.as<uint8_t>(), (uint32_t)t.size(), msg_len, tail_len, d_messages,
                                                                        index_msgs ? 1u : 0u, first_index, prefix, (Affine<BP>*)d_out, count);
  ctx->kernel_launches++;
  BZ_CUDA(cudaGetLastError());
  BZ_CUDA(cudaStreamSynchronize(ctx->stream));      // d_t and t go out of scope
}
void hash_to_curve_run(Ctx* ctx, int curve, const char* domain, const uint8_t* d_messages, uint32_t msg_len, bool index_msgs,
                       uint32_t first_index, uint8_t prefix, void* d_out, uint32_t count) {
  if (!count) return;
  if (curve == 0) hash_to_curve_t<FqP>(ctx, curve, domain, d_messages, msg_len, index_msgs, first_index, prefix, d_out, count);
  else hash_to_curve_t<FpP>(ctx, curve, domain, d_messages, msg_len, index_msgs, first_index, prefix, d_out, count);
}

// g_lagrange = 2^-k * group-FFT_{omega^-1}(g)      (d_g, d_out: n affine points on the device)
template <class BP, class SP>
static void ec_ifft_t(Ctx* ctx, int sfield, const void* d_g, void* d_out, uint32_t k) {
  const uint32_t n = 1u << k;
  cudaStream_t st = ctx->stream;
  const bzh::Field& F = ctx->field(sfield);
  bzh::Fe om = F.root_of_unity();
  for (uint32_t i = k; i < 32; ++i) om = F.sqr(om);
  om = F.inv(om);
  bzh::Fe ninv = F.inv(F.from_u64(n));
  uint64_t ninv_raw[4]; F.to_raw(ninv, ninv_raw);
  DevBuf a, tw, d_ninv;
  a.alloc((size_t)n * sizeof(Xyzz<BP>));
  tw.alloc((size_t)std::max(1u, n / 2) * 32);
  d_ninv.alloc(32);
  BZ_CUDA(cudaMemcpyAsync(d_ninv.p, ninv_raw, 32, cudaMemcpyHostToDevice, st));
  Fe<SP> omd; memcpy(omd.l, om.l, 32);
  if (n > 1) ecfft_twiddle_kernel<SP><<<(n / 2 + 127) / 128, 128, 0, st>>>(tw.as<Fe<SP>>(), omd, n / 2);
  ecfft_load_kernel<BP><<<(n + 127) / 128, 128, 0, st>>>((const Affine<BP>*)d_g, a.as<Xyzz<BP>>(), k);
  for (uint32_t s = 1; s <= k; ++s)
    ecfft_stage_kernel<BP, SP><<<(n / 2 + 63) / 64, 64, 0, st>>>(a.as<Xyzz<BP>>(), tw.as<Fe<SP>>(), k, s);
  ec_scale_kernel<BP><<<(n + 63) / 64, 64, 0, st>>>(a.as<Xyzz<BP>>(), d_ninv.as<uint32_t>(), n);
  const uint64_t nthreads = ((uint64_t)n + 15) / 16;
  xyzz_normalize_kernel<BP, 16><<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(a.as<Xyzz<BP>>(), (Affine<BP>*)d_out, n);
  ctx->kernel_launches += 4 + k;
  BZ_CUDA(cudaGetLastError());
  BZ_CUDA(cudaStreamSynchronize(st));
}
void ec_ifft_run(Ctx* ctx, int curve, const void* d_g, void* d_out, uint32_t k) {
  if (curve == 0) ec_ifft_t<FqP, FpP>(ctx, 0, d_g, d_out, k); else ec_ifft_t<FpP, FqP>(ctx, 1, d_g, d_out, k);
}

}  // namespace bz

#define BZ_TRY2(ctx_, ...)                                   \
  if (!(ctx_)) return BZ_ERR_INVALID;                        \
  try {                                                      \
    cudaSetDevice((ctx_)->c.device);                         \
    __VA_ARGS__;                                             \
    return BZ_OK;                                            \
  } catch (const bz::Error& e) {                             \
    (ctx_)->c.last_error = e.what();                         \
    return e.code;                                           \
  } catch (const std::exception& e) {                        \
    (ctx_)->c.last_error = e.what();                         \
    return BZ_ERR_INVALID;                                   \
  }

extern "C" {

// group encoding of pasta_curves (`GroupEncoding::to_bytes / from_bytes`), n points per call, host buffers
__attribute__((visibility("default"))) int bz_points_compress(bz_ctx* ctx, int curve, const void* affine, uint64_t n, void* out32) {
  BZ_TRY2(ctx, {
    BZ_CHECK((curve == 0 || curve == 1) && affine && out32 && n < (1ull << 31), "points_compress: bad arguments");
    bz::DevBuf d_in, d_out; d_in.alloc(std::max<size_t>(1, n * 64)); d_out.alloc(std::max<size_t>(1, n * 32));
    cudaStream_t st = ctx->c.stream;
    BZ_CUDA(cudaMemcpyAsync(d_in.p, affine, n * 64, cudaMemcpyHostToDevice, st));
    bz::compress_points_run(&ctx->c, curve, d_in.p, d_out.p, (uint32_t)n);
    BZ_CUDA(cudaMemcpyAsync(out32, d_out.p, n * 32, cudaMemcpyDeviceToHost, st));
    BZ_CUDA(cudaStreamSynchronize(st));
  });
}
// status[i]: 0 ok, 1 identity encoding, 2 not a curve point (from_bytes returns None)
__attribute__((visibility("default"))) int bz_points_decompress(bz_ctx* ctx, int curve, const void* in32, uint64_t n, void* out_affine, uint8_t* status) {
  BZ_TRY2(ctx, {
    BZ_CHECK((curve == 0 || curve == 1) && in32 && out_affine && status && n < (1ull << 31), "points_decompress: bad arguments");
    bz::DevBuf d_in, d_out, d_st; d_in.alloc(std::max<size_t>(1, n * 32)); d_out.alloc(std::max<size_t>(1, n * 64)); d_st.alloc(std::max<size_t>(1, n));
    cudaStream_t st = ctx->c.stream;
    BZ_CUDA(cudaMemcpyAsync(d_in.p, in32, n * 32, cudaMemcpyHostToDevice, st));
    bz::decompress_points_run(&ctx->c, curve, d_in.p, d_out.p, d_st.as<uint8_t>(), (uint32_t)n);
    BZ_CUDA(cudaMemcpyAsync(out_affine, d_out.p, n * 64, cudaMemcpyDeviceToHost, st));
    BZ_CUDA(cudaMemcpyAsync(status, d_st.p, n, cudaMemcpyDeviceToHost, st));
    BZ_CUDA(cudaStreamSynchronize(st));
  });
}

__attribute__((visibility("default"))) int bz_hash_to_curve(bz_ctx* ctx, int curve, const char* domain_prefix, const void* messages,
                                                            uint32_t msg_len, uint64_t count, void* out_affine) {
  BZ_TRY2(ctx, {
    BZ_CHECK(curve == 0 || curve == 1, "bad curve id");
    BZ_CHECK(domain_prefix && out_affine && (messages || !msg_len || !count), "null argument");
    BZ_CHECK(count < (1ull << 31) && msg_len <= 200, "hash_to_curve: count / message length out of range");
    bz::DevBuf d_m, d_o;
    d_m.alloc(std::max<size_t>(1, count * msg_len)); d_o.alloc(std::max<size_t>(1, count * 64));
    cudaStream_t st = ctx->c.stream;
    if (count * msg_len) BZ_CUDA(cudaMemcpyAsync(d_m.p, messages, count * msg_len, cudaMemcpyHostToDevice, st));
    bz::hash_to_curve_run(&ctx->c, curve, domain_prefix, d_m.as<uint8_t>(), msg_len, false, 0, 0, d_o.p, (uint32_t)count);
    BZ_CUDA(cudaMemcpyAsync(out_affine, d_o.p, count * 64, cudaMemcpyDeviceToHost, st));
    BZ_CUDA(cudaStreamSynchronize(st));
  });
}

__attribute__((visibility("default"))) int bz_params_new(bz_ctx* ctx, uint32_t k, int curve, void* g, void* g_lagrange, void* w, void* u) {
  BZ_TRY2(ctx, {
    BZ_CHECK(curve == 0 || curve == 1, "bad curve id");
    BZ_CHECK(k >= 1 && k <= 24, "k out of range");
    BZ_CHECK(g && g_lagrange && w && u, "null argument");
    const size_t n = (size_t)1 << k;
    bz::DevBuf d_g, d_gl, d_wu;
    d_g.alloc(n * 64); d_gl.alloc(n * 64); d_wu.alloc(128);
    cudaStream_t st = ctx->c.stream;
    bz::hash_to_curve_run(&ctx->c, curve, "Halo2-Parameters", nullptr, 5, true, 0, 0, d_g.p, (uint32_t)n);
    // w = hasher([1]), u = hasher([2]): one-byte messages
    const uint8_t wu_msgs[2] = {1, 2};
    bz::DevBuf d_m; d_m.alloc(2);
    BZ_CUDA(cudaMemcpyAsync(d_m.p, wu_msgs, 2, cudaMemcpyHostToDevice, st));
    bz::hash_to_curve_run(&ctx->c, curve, "Halo2-Parameters", d_m.as<uint8_t>(), 1, false, 0, 0, d_wu.p, 2);
    bz::ec_ifft_run(&ctx->c, curve, d_g.p, d_gl.p, k);
    BZ_CUDA(cudaMemcpyAsync(g, d_g.p, n * 64, cudaMemcpyDeviceToHost, st));
    BZ_CUDA(cudaMemcpyAsync(g_lagrange, d_gl.p, n * 64, cudaMemcpyDeviceToHost, st));
    BZ_CUDA(cudaMemcpyAsync(w, d_wu.p, 64, cudaMemcpyDeviceToHost, st));
    BZ_CUDA(cudaMemcpyAsync(u, (char*)d_wu.p + 64, 64, cudaMemcpyDeviceToHost, st));
    BZ_CUDA(cudaStreamSynchronize(st));
  });
}

}  // extern "C"
