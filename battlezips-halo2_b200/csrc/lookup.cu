// Lookup permutation for domains that do not fit one CTA's shared memory (n > 4096: the scaled Board circuit, SURVEY
// config 5).  Same result as poly.cuh's `lookup_permute_kernel` (U: halo2_proofs 0.2.0 src/plonk/lookup/prover.rs
// `permute_expression_pair`), built from device-wide primitives: CUB radix sort of the 256-bit canonical keys, one
// binary search per distinct input value, two CUB prefix sums, a gather.
#include "common.h"
#include "field.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda/std/tuple>

namespace bz {

struct Key256 { uint64_t l[4]; };      // canonical value, little-endian limbs
struct Key256Decomposer {
  __host__ __device__ ::cuda::std::tuple<uint64_t&, uint64_t&, uint64_t&, uint64_t&> operator()(Key256& k) const {
    return {k.l[3], k.l[2], k.l[1], k.l[0]};     // most significant first
  }
};

__device__ __forceinline__ bool k_less(const Key256& a, const Key256& b) {
  for (int i = 3; i >= 0; --i) { if (a.l[i] != b.l[i]) return a.l[i] < b.l[i]; }
  return false;
}
__device__ __forceinline__ bool k_eq(const Key256& a, const Key256& b) {
  return a.l[0] == b.l[0] && a.l[1] == b.l[1] && a.l[2] == b.l[2] && a.l[3] == b.l[3];
}
template <class P> __device__ __forceinline__ Fe<P> key_to_mont(const Key256& k) {
  Fe<P> v;
#pragma unroll
  for (int i = 0; i < 4; ++i) { v.l[2 * i] = (uint32_t)k.l[i]; v.l[2 * i + 1] = (uint32_t)(k.l[i] >> 32); }
  return fe_to_mont(v);
}

template <class P>
__global__ void lkl_canon_kernel(const Fe<P>* __restrict__ in, Key256* __restrict__ out, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Fe<P> v = fe_from_mont(fe_load(in + i));
  Key256 k;
#pragma unroll
  for (int j = 0; j < 4; ++j) k.l[j] = (uint64_t)v.l[2 * j] | ((uint64_t)v.l[2 * j + 1] << 32);
  out[i] = k;
}
// A' = sorted input; repeat flags; every distinct input takes one copy out of the sorted table
template <class P>
__global__ void lkl_mark_kernel(const Key256* __restrict__ a, const Key256* __restrict__ t, uint32_t count, Fe<P>* __restrict__ aout,
                                uint32_t* __restrict__ repeat, uint32_t* __restrict__ removed, uint32_t* __restrict__ err) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const Key256 k = a[i];
  fe_store(aout + i, key_to_mont<P>(k));
  const bool rep = i > 0 && k_eq(k, a[i - 1]);
  repeat[i] = rep ? 1u : 0u;
  if (rep) return;
  uint32_t lo = 0, hi = count;
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (k_less(t[mid], k)) lo = mid + 1; else hi = mid; }
  if (lo < count && k_eq(t[lo], k)) removed[lo] = 1u; else atomicOr(err, 1u);
}
__global__ void lkl_kept_kernel(uint32_t* __restrict__ removed_to_kept, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) removed_to_kept[i] = removed_to_kept[i] ? 0u : 1u;
}
__global__ void lkl_compact_kernel(const uint32_t* __restrict__ kept, const uint32_t* __restrict__ kept_rank, uint32_t* __restrict__ left, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count && kept[i]) left[kept_rank[i]] = i;
}
template <class P>
__global__ void lkl_gather_kernel(const Key256* __restrict__ a, const Key256* __restrict__ t, const uint32_t* __restrict__ repeat,
                                  const uint32_t* __restrict__ repeat_rank, const uint32_t* __restrict__ kept, const uint32_t* __restrict__ kept_rank,
                                  const uint32_t* __restrict__ left, uint32_t count, Fe<P>* __restrict__ sout, uint32_t* __restrict__ err) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t n_rep = repeat_rank[count - 1] + repeat[count - 1], n_left = kept_rank[count - 1] + kept[count - 1];
  if (n_rep != n_left) { if (i == 0) atomicOr(err, 2u); return; }
  const Key256 k = repeat[i] ? t[left[n_left - 1 - repeat_rank[i]]] : a[i];
  fe_store(sout + i, key_to_mont<P>(k));
}

// cin / ctab / aout / sout: device arrays of `usable` (or more) Fp elements; err: device word (bit 0: input not in table)
void lookup_permute_large_run(Ctx* ctx, const void* cin, const void* ctab, void* aout, void* sout, uint32_t usable, uint32_t* d_err) {
  typedef Fe<FpP> F;
  cudaStream_t st = ctx->stream;
  if (!usable) return;
  size_t sort_bytes = 0, scan_bytes = 0;
  BZ_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, (const Key256*)nullptr, (Key256*)nullptr, (int64_t)usable, Key256Decomposer{}, st));
  BZ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)usable, st));
  const size_t kb = (size_t)usable * sizeof(Key256), wb = ((size_t)usable * 4 + 255) & ~size_t(255);
  DevBuf& buf = ctx->stage[2];
  buf.ensure(4 * kb + 5 * wb + std::max(sort_bytes, scan_bytes) + 256);
  char* p = (char*)buf.p;
  Key256* ka = (Key256*)p; p += kb;
  Key256* kt = (Key256*)p; p += kb;
  Key256* sa = (Key256*)p; p += kb;
  Key256* stb = (Key256*)p; p += kb;
  uint32_t* repeat = (uint32_t*)p; p += wb;
  uint32_t* repeat_rank = (uint32_t*)p; p += wb;
  uint32_t* kept = (uint32_t*)p; p += wb;
  uint32_t* kept_rank = (uint32_t*)p; p += wb;
  uint32_t* left = (uint32_t*)p; p += wb;
  void* temp = p;
  size_t temp_bytes = std::max(sort_bytes, scan_bytes);
  const unsigned blocks = (usable + 127) / 128;
  lkl_canon_kernel<FpP><<<blocks, 128, 0, st>>>((const F*)cin, ka, usable);
  lkl_canon_kernel<FpP><<<blocks, 128, 0, st>>>((const F*)ctab, kt, usable);
  size_t tb = temp_bytes;
  BZ_CUDA(cub::DeviceRadixSort::SortKeys(temp, tb, (const Key256*)ka, sa, (int64_t)usable, Key256Decomposer{}, st));
  tb = temp_bytes;
  BZ_CUDA(cub::DeviceRadixSort::SortKeys(temp, tb, (const Key256*)kt, stb, (int64_t)usable, Key256Decomposer{}, st));
  BZ_CUDA(cudaMemsetAsync(kept, 0, (size_t)usable * 4, st));
  lkl_mark_kernel<FpP><<<blocks, 128, 0, st>>>(sa, stb, usable, (F*)aout, repeat, kept, d_err);
  lkl_kept_kernel<<<blocks, 128, 0, st>>>(kept, usable);
  tb = temp_bytes;
  BZ_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, (const uint32_t*)repeat, repeat_rank, (int64_t)usable, st));
  tb = temp_bytes;
  BZ_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, (const uint32_t*)kept, kept_rank, (int64_t)usable, st));
  lkl_compact_kernel<<<blocks, 128, 0, st>>>(kept, kept_rank, left, usable);
  lkl_gather_kernel<FpP><<<blocks, 128, 0, st>>>(sa, stb, repeat, repeat_rank, kept, kept_rank, left, usable, (F*)sout, d_err);
  ctx->kernel_launches += 10;
  BZ_CUDA(cudaGetLastError());
}

}  // namespace bz
