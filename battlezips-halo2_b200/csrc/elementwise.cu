// Element-wise field / point operations exposed for the parity suite and for the small host-driven
// steps of the prover (ff::Field ops on slices; group add on slices).  One thread per element.
#include "common.h"
#include "curve.cuh"

namespace bz {

template <class P>
__global__ void field_op_kernel(int op, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, Fe<P>* __restrict__ out, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe<P> r;
  if (op == 4) {          // from_u512: a holds 16 limbs per element
    uint32_t w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = a[i * 16 + j];
    r = fe_from_u512<P>(w);
  } else {
    Fe<P> x = fe_load(reinterpret_cast<const Fe<P>*>(a) + i);
    Fe<P> y = (op <= 2) ? fe_load(reinterpret_cast<const Fe<P>*>(b) + i) : fe_zero<P>();
    switch (op) {
      case 0: r = fe_mul(x, y); break;
      case 1: r = fe_add(x, y); break;
      case 2: r = fe_sub(x, y); break;
      case 3: r = fe_inv(x); break;
      case 5: r = fe_from_mont(x); break;
      case 6: r = fe_to_mont(x); break;
      case 7: r = fe_neg(x); break;
      default: r = fe_sqr(x); break;
    }
  }
  fe_store(out + i, r);
}

// op 9: element-wise inversion by binary GCD (gcdinv.h) -- its own kernel, so that field_op_kernel stays the validated code
template <class P>
__global__ void field_inv_gcd_kernel(const Fe<P>* __restrict__ a, Fe<P>* __restrict__ out, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fe_store(out + i, fe_inv_gcd(fe_load(a + i)));
}

// ff::BatchInvert: Montgomery's trick over RUN consecutive elements per thread, zeros skipped (0 -> 0);
// the same values as element-wise inversion at 3 multiplications per element plus one Fermat chain per RUN
template <class P, int RUN>
__global__ void batch_invert_kernel(const Fe<P>* __restrict__ in, Fe<P>* __restrict__ out, uint64_t n) {
  const uint64_t lo = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * RUN;
  if (lo >= n) return;
  const uint32_t cnt = (uint32_t)min((uint64_t)RUN, n - lo);
  Fe<P> pre[RUN], v[RUN];
  Fe<P> acc = fe_one<P>();
  for (uint32_t j = 0; j < cnt; ++j) {
    v[j] = fe_load(in + lo + j);
    pre[j] = acc;
    if (!fe_is_zero(v[j])) acc = fe_mul(acc, v[j]);
  }
  Fe<P> inv = fe_inv(acc);
  for (int j = (int)cnt - 1; j >= 0; --j) {
    Fe<P> r = fe_zero<P>();
    if (!fe_is_zero(v[j])) { r = fe_mul(inv, pre[j]); inv = fe_mul(inv, v[j]); }
    fe_store(out + lo + j, r);
  }
}

// poly::batch_invert_assigned (U: halo2_proofs 0.2.0 src/poly.rs): the witness columns arrive as Assigned<F>
// = numerator / denominator (Zero and Trivial carry denominator one); all denominators of a column are inverted with
// ff::BatchInvert (zeros skipped, so n / 0 evaluates to 0 like Assigned::evaluate) and multiplied into the numerators.
template <class P, int RUN>
__global__ void batch_invert_assigned_kernel(const Fe<P>* num, const Fe<P>* __restrict__ den, Fe<P>* out /* may alias num */, uint64_t n) {
  const uint64_t lo = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * RUN;
  if (lo >= n) return;
  const uint32_t cnt = (uint32_t)min((uint64_t)RUN, n - lo);
  Fe<P> pre[RUN], v[RUN];
  Fe<P> acc = fe_one<P>();
  bool all_one = true;
  for (uint32_t j = 0; j < cnt; ++j) {
    v[j] = fe_load(den + lo + j);
    pre[j] = acc;
    if (!fe_is_zero(v[j]) && !fe_eq(v[j], fe_one<P>())) { acc = fe_mul(acc, v[j]); all_one = false; }
  }
  Fe<P> inv = all_one ? acc : fe_inv(acc);            // a run of Trivial cells (the common case) needs no Fermat chain
  for (int j = (int)cnt - 1; j >= 0; --j) {
    Fe<P> r = fe_zero<P>();
    if (fe_eq(v[j], fe_one<P>())) r = fe_load(num + lo + j);
    else if (!fe_is_zero(v[j])) { r = fe_mul(fe_mul(inv, pre[j]), fe_load(num + lo + j)); inv = fe_mul(inv, v[j]); }
    fe_store(out + lo + j, r);
  }
}
void batch_invert_assigned_run(Ctx* ctx, int field, const void* num, const void* den, void* out, uint64_t n) {
  if (!n) return;
  const uint64_t threads = (n + 7) / 8;
  const unsigned bl = (unsigned)((threads + 63) / 64);
  if (field == 0) batch_invert_assigned_kernel<FpP, 8><<<bl, 64, 0, ctx->stream>>>((const Fe<FpP>*)num, (const Fe<FpP>*)den, (Fe<FpP>*)out, n);
  else batch_invert_assigned_kernel<FqP, 8><<<bl, 64, 0, ctx->stream>>>((const Fe<FqP>*)num, (const Fe<FqP>*)den, (Fe<FqP>*)out, n);
  ctx->kernel_launches++;
  BZ_CUDA(cudaGetLastError());
}

// op 0: a + b (mixed), 1: 2a, 2: a - b, 3: full XYZZ add of (a+a) and b, 4: [k]a with k = low 32 bits of b.x raw
template <class BP>
__global__ void curve_op_kernel(int op, const Affine<BP>* __restrict__ a, const Affine<BP>* __restrict__ b, Affine<BP>* __restrict__ out, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<BP> p = aff_load(a + i), q = aff_load(b + i);
  Xyzz<BP> acc = xyzz_from_affine(p);
  if (op == 0) xyzz_add_mixed(acc, q);
  else if (op == 1) acc = xyzz_dbl(acc);
  else if (op == 2) xyzz_add_mixed_signed(acc, q, true);
  else if (op == 3) acc = xyzz_add(xyzz_dbl(acc), xyzz_from_affine(q));
  else acc = xyzz_mul_u32(acc, q.x.l[0]);
  Affine<BP> r = xyzz_to_affine(acc);
  fe_store(&out[i].x, r.x); fe_store(&out[i].y, r.y);
}

// sum of `count` Jacobian points (the all-gathered partials of a point-range-split MSM) -> one affine point; one thread:
// count is the number of GPUs
template <class BP>
__global__ void jac_sum_kernel(const Jac<BP>* __restrict__ in, uint32_t count, Affine<BP>* __restrict__ out) {
  if (blockIdx.x || threadIdx.x) return;
  Xyzz<BP> acc = xyzz_identity<BP>();
  for (uint32_t i = 0; i < count; ++i) {
    Jac<BP> p; p.x = fe_load(&in[i].x); p.y = fe_load(&in[i].y); p.z = fe_load(&in[i].z);
    acc = xyzz_add(acc, jac_to_xyzz(p));
  }
  Affine<BP> r = xyzz_to_affine(acc);
  fe_store(&out->x, r.x); fe_store(&out->y, r.y);
}
void jac_sum_run(Ctx* ctx, int curve, const void* d_jac, uint32_t count, void* d_out_affine) {
  if (curve == 0) jac_sum_kernel<FqP><<<1, 32, 0, ctx->stream>>>((const Jac<FqP>*)d_jac, count, (Affine<FqP>*)d_out_affine);
  else jac_sum_kernel<FpP><<<1, 32, 0, ctx->stream>>>((const Jac<FpP>*)d_jac, count, (Affine<FpP>*)d_out_affine);
  ctx->kernel_launches++;
  BZ_CUDA(cudaGetLastError());
}

void field_op_run(Ctx* ctx, int field, int op, const void* a, const void* b, void* out, uint64_t n) {
  if (!n) return;
  if (op == 3) {          // inversion of a slice = batch inversion
    const uint64_t threads = (n + 15) / 16;
    const unsigned bl = (unsigned)((threads + 63) / 64);
    if (field == 0) batch_invert_kernel<FpP, 16><<<bl, 64, 0, ctx->stream>>>((const Fe<FpP>*)a, (Fe<FpP>*)out, n);
    else batch_invert_kernel<FqP, 16><<<bl, 64, 0, ctx->stream>>>((const Fe<FqP>*)a, (Fe<FqP>*)out, n);
    ctx->kernel_launches++;
    BZ_CUDA(cudaGetLastError());
    return;
  }
  unsigned blocks = (unsigned)((n + 127) / 128);
  if (op == 9) {
    if (field == 0) field_inv_gcd_kernel<FpP><<<blocks, 128, 0, ctx->stream>>>((const Fe<FpP>*)a, (Fe<FpP>*)out, n);
    else field_inv_gcd_kernel<FqP><<<blocks, 128, 0, ctx->stream>>>((const Fe<FqP>*)a, (Fe<FqP>*)out, n);
    ctx->kernel_launches++;
    BZ_CUDA(cudaGetLastError());
    return;
  }
  if (field == 0) field_op_kernel<FpP><<<blocks, 128, 0, ctx->stream>>>(op, (const uint32_t*)a, (const uint32_t*)b, (Fe<FpP>*)out, n);
  else field_op_kernel<FqP><<<blocks, 128, 0, ctx->stream>>>(op, (const uint32_t*)a, (const uint32_t*)b, (Fe<FqP>*)out, n);
  ctx->kernel_launches++;
  BZ_CUDA(cudaGetLastError());
}
void curve_op_run(Ctx* ctx, int curve, int op, const void* a, const void* b, void* out, uint64_t n) {
  if (!n) return;
  unsigned blocks = (unsigned)((n + 63) / 64);
  if (curve == 0) curve_op_kernel<FqP><<<blocks, 64, 0, ctx->stream>>>(op, (const Affine<FqP>*)a, (const Affine<FqP>*)b, (Affine<FqP>*)out, n);
  else curve_op_kernel<FpP><<<blocks, 64, 0, ctx->stream>>>(op, (const Affine<FpP>*)a, (const Affine<FpP>*)b, (Affine<FpP>*)out, n);
  ctx->kernel_launches++;
  BZ_CUDA(cudaGetLastError());
}

}  // namespace bz
