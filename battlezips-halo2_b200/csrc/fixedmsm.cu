// Fixed-base multi-scalar multiplication for `Params::commit` / `Params::commit_lagrange` and the IPA
// L/R terms (U: halo2_proofs 0.2.0 src/poly/commitment.rs, src/poly/commitment/prover.rs).
//
// B200-first design: every commitment of every proof uses one of two fixed bases (g || w || u  or
// g_lagrange || w), so with 180 GB of HBM we tabulate ALL signed-digit multiples
//     T[w][d-1][i] = d * 2^(c*w) * B_i ,  d = 1 .. 2^(c-1),  w < ceil(256/c)
// once per Params (affine, 64 B per entry; a few GB).  A commitment is then a flat sum of at most
// ceil(256/c) * (n+1) table entries: no buckets, no sorting, no doublings, no bucket reduction, and
// zero digits (witness columns are mostly 0/1) cost nothing.  One launch commits a whole batch of
// polynomials: grid = (chunks, #msm).  All arithmetic is IMAD-pipe Montgomery work in XYZZ coordinates.
#include "common.h"
#include "curve.cuh"
#include "fixedmsm.h"
#include <cstdlib>

namespace bz {

// ---- table construction (one-off per Params) -----------------------------------------------------
template <class BP>
__global__ void fb_multiples_kernel(const Affine<BP>* __restrict__ base, uint32_t npts, uint32_t nbk, Xyzz<BP>* __restrict__ scratch) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts) return;
  Affine<BP> p = aff_load(base + i);
  Xyzz<BP> acc = xyzz_identity<BP>();
  for (uint32_t d = 0; d < nbk; ++d) {
    xyzz_add_mixed(acc, p);
    Xyzz<BP>* o = scratch + (size_t)d * npts + i;
    fe_store(&o->x, acc.x); fe_store(&o->y, acc.y); fe_store(&o->zz, acc.zz); fe_store(&o->zzz, acc.zzz);
  }
}

// XYZZ -> affine with Montgomery's trick over RUN consecutive entries per thread
template <class BP, int RUN>
__global__ void fb_to_affine_kernel(const Xyzz<BP>* __restrict__ in, Affine<BP>* __restrict__ out, uint64_t total) {
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
      Fe<BP> zi3 = fe_mul(inv, pre[j]);        // 1/zzz
      inv = fe_mul(inv, zzz);
      Fe<BP> zz = fe_load(&q->zz);
      Fe<BP> zi2 = fe_mul(fe_mul(fe_sqr(zi3), zz), zz);   // zzz^-2 * zz^2 = 1/zz
      r.x = fe_mul(fe_load(&q->x), zi2);
      r.y = fe_mul(fe_load(&q->y), zi3);
    }
    fe_store(&out[lo + j].x, r.x); fe_store(&out[lo + j].y, r.y);
  }
}

// next window's base: 2 * (2^(c-1) * P) -> affine
template <class BP>
__global__ void fb_next_base_kernel(const Xyzz<BP>* __restrict__ scratch_last, Affine<BP>* __restrict__ base, uint32_t npts) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts) return;
  const Xyzz<BP>* q = scratch_last + i;
  Xyzz<BP> v; v.x = fe_load(&q->x); v.y = fe_load(&q->y); v.zz = fe_load(&q->zz); v.zzz = fe_load(&q->zzz);
  Affine<BP> r = xyzz_to_affine(xyzz_dbl(v));
  fe_store(&base[i].x, r.x); fe_store(&base[i].y, r.y);
}

template <class BP>
static void fb_build_t(Ctx* ctx, FixedBase& fb, const void* bases_dev) {
  cudaStream_t st = ctx->stream;
  const uint32_t npts = fb.npts, nbk = fb.nbk, W = fb.W;
  fb.table.alloc((size_t)W * nbk * npts * sizeof(Affine<BP>));
  DevBuf scratch, base;
  scratch.alloc((size_t)nbk * npts * sizeof(Xyzz<BP>));
  base.alloc((size_t)npts * sizeof(Affine<BP>));
  BZ_CUDA(cudaMemcpyAsync(base.p, bases_dev, (size_t)npts * sizeof(Affine<BP>), cudaMemcpyDeviceToDevice, st));
  const uint64_t per_window = (uint64_t)nbk * npts;
  for (uint32_t w = 0; w < W; ++w) {
    fb_multiples_kernel<BP><<<(npts + 63) / 64, 64, 0, st>>>(base.as<Affine<BP>>(), npts, nbk, scratch.as<Xyzz<BP>>());
    uint64_t nthreads = (per_window + 15) / 16;
    fb_to_affine_kernel<BP, 16><<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(
        scratch.as<Xyzz<BP>>(), fb.table.as<Affine<BP>>() + (size_t)w * per_window, per_window);
    if (w + 1 < W)
      fb_next_base_kernel<BP><<<(npts + 63) / 64, 64, 0, st>>>(scratch.as<Xyzz<BP>>() + (size_t)(nbk - 1) * npts, base.as<Affine<BP>>(), npts);
    ctx->kernel_launches += 3;
  }
  BZ_CUDA(cudaGetLastError());
  BZ_CUDA(cudaStreamSynchronize(st));
}

void fixed_base_build(Ctx* ctx, FixedBase& fb, int curve, const void* bases_dev, uint32_t npts, uint32_t c) {
  BZ_CHECK(c >= 4 && c <= 16, "fixed-base window out of range");
  BZ_CHECK((uint64_t)((256 + c - 1) / c) * (1ull << (c - 1)) * npts < (1ull << 31), "fixed-base table too large for 31-bit entry indices");
  fb.curve = curve; fb.npts = npts; fb.c = c; fb.W = (256 + c - 1) / c; fb.nbk = 1u << (c - 1);
  if (curve == 0) fb_build_t<FqP>(ctx, fb, bases_dev); else fb_build_t<FpP>(ctx, fb, bases_dev);
}

// ---- the commitment kernel -----------------------------------------------------------------------
// grid = (chunks, n_msm), block = FB_THREADS.  MSM m sums over points [0, npts): scalars of points
// i < n_main come from main[m] (a polynomial, Montgomery form), the trailing npts - n_main points
// (w, u: blinds / IPA cross terms) from extra[m] (may be null = zero).
//
// Work compaction: witness columns are mostly 0 / 1, so "one thread = one point" leaves 2/3 of every warp idle
// (ncu r1: 11-13 active threads per warp-instruction).  Each round the CTA decodes FB_THREADS scalars into
// signed digits, packs the non-zero ones as table indices into a shared-memory work list (block-wide exclusive
// scan of the per-thread counts, no atomics, deterministic order) and then ALL threads pull entries from the
// list round-robin: every warp runs full until the list is drained.
constexpr int FB_THREADS = 128;
constexpr int FB_MAX_W = 64;       // windows per scalar (c >= 4)

template <class BP, class SP, int MINB>
__global__ void __launch_bounds__(FB_THREADS, MINB) fixed_msm_kernel(const Affine<BP>* __restrict__ table, uint32_t npts, uint32_t c, uint32_t W, uint32_t nbk,
                                 const Fe<SP>* const* __restrict__ main, uint32_t n_main, const Fe<SP>* const* __restrict__ extra,
                                 Xyzz<BP>* __restrict__ partial, unsigned long long* __restrict__ add_counter) {
  extern __shared__ uint32_t fb_smem[];
  uint32_t* list = fb_smem;                                  // FB_THREADS * W entries
  __shared__ uint32_t warp_cnt[FB_THREADS / 32];
  uint32_t my_adds = 0;
  const uint32_t m = blockIdx.y, chunks = gridDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const Fe<SP>* sm = main[m];
  const Fe<SP>* se = extra ? extra[m] : nullptr;
  Xyzz<BP> acc = xyzz_identity<BP>();
  const uint32_t half = 1u << (c - 1), full = 1u << c;
  // this CTA's contiguous point range
  const uint32_t per = (npts + chunks - 1) / chunks;
  const uint32_t p_lo = blockIdx.x * per, p_hi = min(p_lo + per, npts);
  for (uint32_t base = p_lo; base < p_hi; base += FB_THREADS) {
    const uint32_t i = base + tid;
    // ---- decode this thread's scalar into packed entries (kept in registers as a bitmap + digits recomputed) ----
    Fe<SP> s = fe_zero<SP>();
    bool have = false;
    if (i < p_hi) {
      if (i < n_main) { s = fe_load(sm + i); have = true; }
      else if (se) { s = fe_load(se + (i - n_main)); have = true; }
    }
    uint32_t cnt = 0;
    if (have && !fe_is_zero(s)) {
      s = fe_from_mont(s);
      uint32_t carry = 0;
      for (uint32_t w = 0; w < W; ++w) {
        uint32_t bit = w * c, limb = bit >> 5, sh_ = bit & 31;
        uint32_t raw = 0;
        if (limb < 8) {
          raw = s.l[limb] >> sh_;
          if (sh_ + c > 32 && limb + 1 < 8) raw |= s.l[limb + 1] << (32 - sh_);
          raw &= full - 1;
        }
        uint32_t v = raw + carry;
        carry = v > half ? 1u : 0u;
        cnt += (v != 0 && v != full) ? 1u : 0u;      // v == full cannot happen (raw <= full-1, carry makes v<=full; v==full -> d=0)
      }
    } else have = false;
    // ---- block-wide exclusive scan of cnt ----
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += o; }
    if (lane == 31) warp_cnt[wid] = incl;
    __syncthreads();
    uint32_t off = incl - cnt, total = 0;
#pragma unroll
    for (int w2 = 0; w2 < FB_THREADS / 32; ++w2) { uint32_t t = warp_cnt[w2]; if (w2 < (int)wid) off += t; total += t; }
    // ---- write entries ----
    if (have) {
      uint32_t carry = 0;
      for (uint32_t w = 0; w < W; ++w) {
        uint32_t bit = w * c, limb = bit >> 5, sh_ = bit & 31;
        uint32_t raw = 0;
        if (limb < 8) {
          raw = s.l[limb] >> sh_;
          if (sh_ + c > 32 && limb + 1 < 8) raw |= s.l[limb + 1] << (32 - sh_);
          raw &= full - 1;
        }
        uint32_t v = raw + carry;
        bool neg = v > half;
        uint32_t d = neg ? full - v : v;
        carry = neg ? 1u : 0u;
        if (d) list[off++] = (uint32_t)(((size_t)w * nbk + (d - 1)) * npts + i) | (neg ? 0x80000000u : 0u);
      }
    }
    __syncthreads();
    // ---- drain the list: all threads busy ----
    for (uint32_t e = tid; e < total; e += FB_THREADS) {
      uint32_t ent = list[e];
      Affine<BP> pt = aff_load(table + (ent & 0x7fffffffu));
      xyzz_add_mixed_signed(acc, pt, (ent >> 31) != 0);
      ++my_adds;
    }
    __syncthreads();
  }
  if (add_counter) {       // profiling only: exact number of mixed additions this launch performed
    uint32_t tot = __reduce_add_sync(0xffffffffu, my_adds);
    if (lane == 0 && tot) atomicAdd(add_counter, (unsigned long long)tot);
  }
  // ---- CTA tree reduction (the list buffer is reused as XYZZ scratch: FB_THREADS * 128 B) ----
  Xyzz<BP>* sh = reinterpret_cast<Xyzz<BP>*>(fb_smem);
  sh[tid] = acc;
  __syncthreads();
  for (uint32_t d = FB_THREADS >> 1; d > 0; d >>= 1) {
    if (tid < d) sh[tid] = xyzz_add(sh[tid], sh[tid + d]);
    __syncthreads();
  }
  if (tid == 0) {
    Xyzz<BP>* o = partial + (size_t)m * chunks + blockIdx.x;
    fe_store(&o->x, sh[0].x); fe_store(&o->y, sh[0].y); fe_store(&o->zz, sh[0].zz); fe_store(&o->zzz, sh[0].zzz);
  }
}

// fold the chunk partials of each MSM and normalise: one thread per MSM -> affine (64 B), identity = zeros
template <class BP>
__global__ void fixed_msm_finish_kernel(const Xyzz<BP>* __restrict__ partial, uint32_t chunks, uint32_t n_msm, Affine<BP>* __restrict__ out) {
  uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_msm) return;
  Xyzz<BP> acc = xyzz_identity<BP>();
  for (uint32_t j = 0; j < chunks; ++j) {
    const Xyzz<BP>* q = partial + (size_t)m * chunks + j;
    Xyzz<BP> v; v.x = fe_load(&q->x); v.y = fe_load(&q->y); v.zz = fe_load(&q->zz); v.zzz = fe_load(&q->zzz);
    acc = xyzz_add(acc, v);
  }
  Affine<BP> r = xyzz_to_affine(acc);
  fe_store(&out[m].x, r.x); fe_store(&out[m].y, r.y);
}

template <class BP, class SP>
static void fixed_msm_run_t(Ctx* ctx, const FixedBase& fb, const void* const* d_main, uint32_t n_main, const void* const* d_extra,
                            uint32_t n_msm, uint32_t chunks, void* d_out_affine) {
  cudaStream_t st = ctx->stream;
  if (!ctx->counters.p) { ctx->counters.alloc(64); BZ_CUDA(cudaMemsetAsync(ctx->counters.p, 0, 64, st)); }
  ctx->scratch[3].ensure((size_t)n_msm * chunks * sizeof(Xyzz<BP>));
  Xyzz<BP>* partial = ctx->scratch[3].as<Xyzz<BP>>();
  {
    ProfScope p(ctx, PROF_FIXED_MSM);
    const size_t smem = std::max<size_t>((size_t)FB_THREADS * fb.W * 4, (size_t)FB_THREADS * sizeof(Xyzz<BP>));
    static int occ = -1;
    if (occ < 0) { const char* e = getenv("BZ_MSM_OCC"); occ = e ? atoi(e) : 4; }   // measured on B200 (profiles/README.md): 4 CTAs/SM is fastest; IMAD.WIDE chains saturate the fma-heavy pipe
    unsigned long long* cnt = ctx->profiling ? (unsigned long long*)ctx->counters.p : nullptr;
#define BZ_FB_LAUNCH(MINB)                                                                                       \
    fixed_msm_kernel<BP, SP, MINB><<<dim3(chunks, n_msm), FB_THREADS, smem, st>>>(                                 \
        fb.table.as<Affine<BP>>(), fb.npts, fb.c, fb.W, fb.nbk, (const Fe<SP>* const*)d_main, n_main,             \
        (const Fe<SP>* const*)d_extra, partial, cnt)
    if (occ >= 8) BZ_FB_LAUNCH(8); else if (occ >= 6) BZ_FB_LAUNCH(6); else if (occ == 5) BZ_FB_LAUNCH(5); else BZ_FB_LAUNCH(4);
#undef BZ_FB_LAUNCH
  }
  fixed_msm_finish_kernel<BP><<<(n_msm + 31) / 32, 32, 0, st>>>(partial, chunks, n_msm, (Affine<BP>*)d_out_affine);
  ctx->kernel_launches += 2;
  BZ_CUDA(cudaGetLastError());
}

void fixed_msm_run(Ctx* ctx, const FixedBase& fb, const void* const* d_main, uint32_t n_main, const void* const* d_extra,
                   uint32_t n_msm, uint32_t chunks, void* d_out_affine) {
  if (!n_msm) return;
  BZ_CHECK(n_main <= fb.npts, "fixed msm: more scalars than table points");
  if (chunks < 1) chunks = 1;
  if (fb.curve == 0) fixed_msm_run_t<FqP, FpP>(ctx, fb, d_main, n_main, d_extra, n_msm, chunks, d_out_affine);
  else fixed_msm_run_t<FpP, FqP>(ctx, fb, d_main, n_main, d_extra, n_msm, chunks, d_out_affine);
}

}  // namespace bz
