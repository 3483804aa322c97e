// Fine-grained C-ABI entry points of the prover's polynomial layer (SURVEY 8b "minimum export set"): one call per inner
// loop of halo2_proofs 0.2.0 that a patched plonk/permutation/prover.rs, plonk/lookup/prover.rs, poly/domain.rs,
// poly/multiopen/prover.rs and arithmetic.rs would bind one-for-one.  Host buffers in and out (pasta's in-memory
// Montgomery form), synchronous on the context's stream; the kernels are the ones create_proof runs batch-major
// (poly.cuh), here with a batch of one.  The pk- and Params-bound calls (bz_pk_quotient, bz_ipa_*) live in prover.cu.
#include "prover_impl.h"

namespace bz {
namespace {

template <class P> struct HostField;
template <> struct HostField<FpP> { static const bzh::Field& get(Ctx* c) { return c->fp; } };
template <> struct HostField<FqP> { static const bzh::Field& get(Ctx* c) { return c->fq; } };

template <class P> __global__ void fg_geometric_kernel(Fe<P>* out, Fe<P> base, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fe_store(out + i, fe_pow_u64<P>(base, i));
}
// a[i] *= t[i mod tn]
template <class P> __global__ void fg_mul_periodic_kernel(Fe<P>* a, const Fe<P>* __restrict__ t, uint32_t tn, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fe_store(a + i, fe_mul(fe_load(a + i), fe_load(t + (i & (tn - 1)))));
}
// acc[i] = acc[i] * x + p[i]
template <class P> __global__ void fg_axpy_kernel(Fe<P>* acc, const Fe<P>* __restrict__ p, Fe<P> x, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fe_store(acc + i, fe_add(fe_mul(fe_load(acc + i), x), fe_load(p + i)));
}

template <class P> Fe<P> dev_fe(const void* host32) { Fe<P> r; memcpy(r.l, host32, 32); return r; }

// the grand product z[0] = z0, z[i+1] = z[i] * num[i] / den[i] from num / den already on the device
template <class P>
void running_product(Ctx* C, Regions& reg, Fe<P>* nd /* 4 x n: num, den, prefix, suffix */, uint32_t n, PolyRef zref, bool has_z0, PolyRef z0ref) {
  cudaStream_t st = C->stream;
  Fe<P>* num = nd; Fe<P>* den = nd + n; Fe<P>* pnum = nd + 2 * (size_t)n; Fe<P>* sden = nd + 3 * (size_t)n;
  const uint32_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (ntiles <= 8) {
    product_scan_kernel<P><<<dim3(1, 1), SCAN_THREADS, 0, st>>>(num, pnum, n, n, 0);
    product_scan_kernel<P><<<dim3(1, 1), SCAN_THREADS, 0, st>>>(den, sden, n, n, 1);
  } else {
    BZ_CHECK(ntiles <= (uint32_t)SCAN_TILE, "grand product: domain too large for the two-level scan");
    DevBuf& tmp = C->stage[5];
    tmp.ensure((size_t)4 * ntiles * 32);
    Fe<P>* tot_n = tmp.as<Fe<P>>(); Fe<P>* car_n = tot_n + ntiles; Fe<P>* tot_d = car_n + ntiles; Fe<P>* car_d = tot_d + ntiles;
    product_scan_kernel<P><<<dim3(ntiles, 1), SCAN_THREADS, 0, st>>>(num, pnum, n, n, 0, 1u, nullptr, tot_n);
    product_scan_kernel<P><<<dim3(ntiles, 1), SCAN_THREADS, 0, st>>>(den, sden, n, n, 1, 1u, nullptr, tot_d);
    product_scan_kernel<P><<<dim3(1, 1), SCAN_THREADS, 0, st>>>(tot_n, car_n, ntiles, ntiles, 0);
    product_scan_kernel<P><<<dim3(1, 1), SCAN_THREADS, 0, st>>>(tot_d, car_d, ntiles, ntiles, 0);
    product_scan_kernel<P><<<dim3(ntiles, 1), SCAN_THREADS, 0, st>>>(num, pnum, n, n, 0, 1u, car_n, nullptr);
    product_scan_kernel<P><<<dim3(ntiles, 1), SCAN_THREADS, 0, st>>>(den, sden, n, n, 1, 1u, car_d, nullptr);
    C->kernel_launches += 4;
  }
  grand_product_finish_kernel<P><<<dim3((n + 127) / 128, 1), 128, 0, st>>>(pnum, sden, n, reg, zref, n, z0ref, 0, has_z0 ? 1 : 0);
  C->kernel_launches += 3;
}

template <class P>
void perm_product_t(Ctx* C, uint32_t k, uint32_t ncols, const void* const* values, const void* const* sigmas, const void* beta, const void* gamma,
                    const void* delta_omega0, const void* z0, void* out_z) {
  const bzh::Field& F = HostField<P>::get(C);
  const uint32_t n = 1u << k;
  cudaStream_t st = C->stream;
  // staging: values[ncols][n], sigmas[ncols][n], omega powers [n], z [n] + z0 [1], num/den/prefix/suffix [4][n], consts [2 + ncols]
  DevBuf &dv = C->stage[0], &ds = C->stage[1], &dz = C->stage[2], &dnd = C->stage[3], &dc = C->stage[4];
  dv.ensure((size_t)ncols * n * 32); ds.ensure(((size_t)ncols + 1) * n * 32); dz.ensure(((size_t)n + 1) * 32); dnd.ensure((size_t)4 * n * 32); dc.ensure((size_t)(2 + ncols) * 32);
  for (uint32_t j = 0; j < ncols; ++j) {
    BZ_CUDA(cudaMemcpyAsync(dv.as<Fe<P>>() + (size_t)j * n, values[j], (size_t)n * 32, cudaMemcpyHostToDevice, st));
    BZ_CUDA(cudaMemcpyAsync(ds.as<Fe<P>>() + (size_t)j * n, sigmas[j], (size_t)n * 32, cudaMemcpyHostToDevice, st));
  }
  Fe<P>* omega_pows = ds.as<Fe<P>>() + (size_t)ncols * n;
  bzh::Fe w = F.root_of_unity();
  for (uint32_t i = k; i < 32; ++i) w = F.sqr(w);
  fg_geometric_kernel<P><<<(n + 127) / 128, 128, 0, st>>>(omega_pows, dev_fe<P>(w.l), n);
  // consts: [beta, gamma, beta * delta_omega0 * delta^j ...]
  std::vector<bzh::Fe> cs(2 + ncols);
  memcpy(cs[0].l, beta, 32); memcpy(cs[1].l, gamma, 32);
  bzh::Fe bd; memcpy(bd.l, delta_omega0, 32);
  bd = F.mul(bd, cs[0]);
  for (uint32_t j = 0; j < ncols; ++j) { cs[2 + j] = bd; bd = F.mul(bd, F.delta()); }
  BZ_CUDA(cudaMemcpyAsync(dc.p, cs.data(), cs.size() * 32, cudaMemcpyHostToDevice, st));
  BZ_CUDA(cudaMemcpyAsync(dz.as<Fe<P>>() + n, z0, 32, cudaMemcpyHostToDevice, st));
  Regions reg; memset(&reg, 0, sizeof(reg));
  reg.base[0] = dv.p; reg.base[1] = dz.p; reg.base[2] = dz.as<Fe<P>>() + n;
  PermSetDesc d{}; d.ncols = ncols;
  for (uint32_t j = 0; j < ncols; ++j) { d.val_kind[j] = 0; d.val_slot[j] = j; d.sigma_slot[j] = j; d.bd_const[j] = 2 + j; }
  perm_fraction_kernel<P><<<dim3((n + 127) / 128, 1), 128, 0, st>>>(reg, n, d, ds.as<Fe<P>>(), omega_pows, dc.as<Fe<P>>(), 2 + ncols, 0, 1, dnd.as<Fe<P>>(), dnd.as<Fe<P>>() + n, n);
  C->kernel_launches += 2;
  running_product<P>(C, reg, dnd.as<Fe<P>>(), n, PolyRef{1, 0}, true, PolyRef{2, 0});
  BZ_CUDA(cudaMemcpyAsync(out_z, dz.p, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
  BZ_CUDA(cudaStreamSynchronize(st));
}

template <class P>
void lookup_product_t(Ctx* C, uint32_t k, const void* cin, const void* ctab, const void* pin, const void* ptab, const void* beta, const void* gamma, void* out_z) {
  const uint32_t n = 1u << k;
  cudaStream_t st = C->stream;
  DevBuf &dv = C->stage[0], &dz = C->stage[2], &dnd = C->stage[3], &dc = C->stage[4];
  dv.ensure((size_t)4 * n * 32); dz.ensure(((size_t)n + 1) * 32); dnd.ensure((size_t)4 * n * 32); dc.ensure(64);
  const void* src[4] = {cin, ctab, pin, ptab};
  for (int j = 0; j < 4; ++j) BZ_CUDA(cudaMemcpyAsync(dv.as<Fe<P>>() + (size_t)j * n, src[j], (size_t)n * 32, cudaMemcpyHostToDevice, st));
  BZ_CUDA(cudaMemcpyAsync(dc.p, beta, 32, cudaMemcpyHostToDevice, st));
  BZ_CUDA(cudaMemcpyAsync(dc.as<Fe<P>>() + 1, gamma, 32, cudaMemcpyHostToDevice, st));
  Regions reg; memset(&reg, 0, sizeof(reg));
  reg.base[0] = dv.p; reg.base[1] = dz.p;
  lookup_fraction_kernel<P><<<dim3((n + 127) / 128, 1), 128, 0, st>>>(reg, n, PolyRef{0, 0}, PolyRef{0, 1}, PolyRef{0, 2}, PolyRef{0, 3}, dc.as<Fe<P>>(), 2, 0, 1,
                                                                    dnd.as<Fe<P>>(), dnd.as<Fe<P>>() + n, n);
  C->kernel_launches++;
  running_product<P>(C, reg, dnd.as<Fe<P>>(), n, PolyRef{1, 0}, false, PolyRef{1, 0});
  BZ_CUDA(cudaMemcpyAsync(out_z, dz.p, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
  BZ_CUDA(cudaStreamSynchronize(st));
}

template <class P>
void eval_many_t(Ctx* C, uint64_t n, uint32_t count, const void* const* polys, const void* points, void* out) {
  cudaStream_t st = C->stream;
  BZ_CHECK(n >= 1 && n < (1ull << 31), "eval: bad length");
  const uint32_t splits = (uint32_t)std::max<uint64_t>(1, n / 16384u);
  DevBuf &dp = C->stage[0], &dc = C->stage[4], &dq = C->stage[1], &dout = C->stage[2], &dtmp = C->stage[3];
  dc.ensure((size_t)count * 32); dq.ensure((size_t)count * sizeof(EvalQuery)); dout.ensure((size_t)count * 32); dtmp.ensure((size_t)count * splits * 32);
  // identical host pointers share one upload
  std::map<const void*, uint32_t> slot;
  std::vector<EvalQuery> q(count);
  for (uint32_t i = 0; i < count; ++i) {
    auto it = slot.find(polys[i]);
    if (it == slot.end()) it = slot.emplace(polys[i], (uint32_t)slot.size()).first;
    q[i] = EvalQuery{PolyRef{0, it->second}, i};
  }
  dp.ensure(slot.size() * n * 32);
  for (auto& kv : slot) BZ_CUDA(cudaMemcpyAsync(dp.as<Fe<P>>() + (size_t)kv.second * n, kv.first, n * 32, cudaMemcpyHostToDevice, st));
  BZ_CUDA(cudaMemcpyAsync(dc.p, points, (size_t)count * 32, cudaMemcpyHostToDevice, st));
  BZ_CUDA(cudaMemcpyAsync(dq.p, q.data(), q.size() * sizeof(EvalQuery), cudaMemcpyHostToDevice, st));
  Regions reg; memset(&reg, 0, sizeof(reg));
  reg.base[0] = dp.p;
  if (splits == 1) eval_queries_kernel<P><<<dim3(count, 1), EVALQ_THREADS, 0, st>>>(reg, (uint32_t)n, dq.as<EvalQuery>(), dc.as<Fe<P>>(), count, dout.as<Fe<P>>(), count);
  else {
    eval_queries_kernel<P><<<dim3(count, 1, splits), EVALQ_THREADS, 0, st>>>(reg, (uint32_t)n, dq.as<EvalQuery>(), dc.as<Fe<P>>(), count, dtmp.as<Fe<P>>(), count);
    eval_reduce_kernel<P><<<(count + 127) / 128, 128, 0, st>>>(dtmp.as<Fe<P>>(), splits, dout.as<Fe<P>>(), count);
    C->kernel_launches++;
  }
  C->kernel_launches++;
  BZ_CUDA(cudaMemcpyAsync(out, dout.p, (size_t)count * 32, cudaMemcpyDeviceToHost, st));
  BZ_CUDA(cudaStreamSynchronize(st));
}

template <class P>
void kate_div_t(Ctx* C, uint64_t n, const void* a, const void* point, void* out_q) {
  cudaStream_t st = C->stream;
  BZ_CHECK(n >= 1 && n < (1ull << 31), "kate_division: bad length");
  if (n == 1) return;
  DevBuf &dp = C->stage[0], &dc = C->stage[4], &dk = C->stage[1], &dtmp = C->stage[3];
  dp.ensure(2 * n * 32); dc.ensure(32); dk.ensure(sizeof(KateDesc));
  BZ_CUDA(cudaMemcpyAsync(dp.p, a, n * 32, cudaMemcpyHostToDevice, st));
  BZ_CUDA(cudaMemcpyAsync(dc.p, point, 32, cudaMemcpyHostToDevice, st));
  KateDesc kd{PolyRef{0, 0}, PolyRef{0, 1}, 0};
  BZ_CUDA(cudaMemcpyAsync(dk.p, &kd, sizeof(kd), cudaMemcpyHostToDevice, st));
  Regions reg; memset(&reg, 0, sizeof(reg));
  reg.base[0] = dp.p;
  const uint32_t splits = (uint32_t)std::max<uint64_t>(1, n / 16384u);
  if (splits == 1) kate_division_kernel<P><<<dim3(1, 1), KATE_THREADS, 0, st>>>(reg, (uint32_t)n, dk.as<KateDesc>(), dc.as<Fe<P>>(), 1);
  else {
    dtmp.ensure((size_t)2 * splits * 32);
    Fe<P>* totals = dtmp.as<Fe<P>>(); Fe<P>* carries = totals + splits;
    kate_division_kernel<P><<<dim3(1, 1, splits), KATE_THREADS, 0, st>>>(reg, (uint32_t)n, dk.as<KateDesc>(), dc.as<Fe<P>>(), 1, nullptr, totals);
    kate_carry_kernel<P><<<1, 64, 0, st>>>((uint32_t)n, splits, 1, 1, dk.as<KateDesc>(), dc.as<Fe<P>>(), 1, totals, carries);
    kate_division_kernel<P><<<dim3(1, 1, splits), KATE_THREADS, 0, st>>>(reg, (uint32_t)n, dk.as<KateDesc>(), dc.as<Fe<P>>(), 1, carries, nullptr);
    C->kernel_launches += 2;
  }
  C->kernel_launches++;
  BZ_CUDA(cudaMemcpyAsync(out_q, dp.as<Fe<P>>() + n, (n - 1) * 32, cudaMemcpyDeviceToHost, st));
  BZ_CUDA(cudaStreamSynchronize(st));
}

template <class P>
void axpy_t(Ctx* C, uint64_t n, void* acc, const void* x, const void* poly) {
  cudaStream_t st = C->stream;
  DevBuf &da = C->stage[0], &dp = C->stage[1];
  da.ensure(n * 32 + 32); dp.ensure(n * 32 + 32);
  BZ_CUDA(cudaMemcpyAsync(da.p, acc, n * 32, cudaMemcpyHostToDevice, st));
  BZ_CUDA(cudaMemcpyAsync(dp.p, poly, n * 32, cudaMemcpyHostToDevice, st));
  fg_axpy_kernel<P><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(da.as<Fe<P>>(), dp.as<Fe<P>>(), dev_fe<P>(x), n);
  C->kernel_launches++;
  BZ_CUDA(cudaMemcpyAsync(acc, da.p, n * 32, cudaMemcpyDeviceToHost, st));
  BZ_CUDA(cudaStreamSynchronize(st));
}

template <class P>
void divide_by_vanishing_t(Ctx* C, uint32_t k, uint32_t ext_k, void* a) {
  const bzh::Field& F = HostField<P>::get(C);
  cudaStream_t st = C->stream;
  const uint64_t en = 1ull << ext_k;
  const uint32_t tn = 1u << (ext_k - k);
  bzh::Fe ext_omega = F.root_of_unity();
  for (uint32_t i = ext_k; i < 32; ++i) ext_omega = F.sqr(ext_omega);
  bzh::Fe cur = F.pow_u64(F.zeta(), 1ull << k), step = F.pow_u64(ext_omega, 1ull << k);
  std::vector<bzh::Fe> t(tn);
  for (uint32_t i = 0; i < tn; ++i) { t[i] = F.inv(F.sub(cur, F.one())); cur = F.mul(cur, step); }
  DevBuf &da = C->stage[0], &dt = C->stage[4];
  da.ensure(en * 32); dt.ensure((size_t)tn * 32);
  BZ_CUDA(cudaMemcpyAsync(da.p, a, en * 32, cudaMemcpyHostToDevice, st));
  BZ_CUDA(cudaMemcpyAsync(dt.p, t.data(), (size_t)tn * 32, cudaMemcpyHostToDevice, st));
  fg_mul_periodic_kernel<P><<<(unsigned)((en + 127) / 128), 128, 0, st>>>(da.as<Fe<P>>(), dt.as<Fe<P>>(), tn, en);
  C->kernel_launches++;
  BZ_CUDA(cudaMemcpyAsync(a, da.p, en * 32, cudaMemcpyDeviceToHost, st));
  BZ_CUDA(cudaStreamSynchronize(st));
}

void lookup_permute_fp(Ctx* C, uint32_t k, uint32_t usable, const void* cin, const void* ctab, void* out_in, void* out_tab) {
  typedef Fe<FpP> F;
  const uint32_t n = 1u << k;
  cudaStream_t st = C->stream;
  DevBuf &dv = C->stage[0], &dso = C->stage[1], &derr = C->stage[4], &dd = C->stage[3];
  dv.ensure((size_t)4 * n * 32); dso.ensure((size_t)n * 32); derr.ensure(64); dd.ensure(sizeof(LookupPermDesc));
  BZ_CUDA(cudaMemcpyAsync(dv.p, cin, (size_t)n * 32, cudaMemcpyHostToDevice, st));
  BZ_CUDA(cudaMemcpyAsync(dv.as<F>() + n, ctab, (size_t)n * 32, cudaMemcpyHostToDevice, st));
  BZ_CUDA(cudaMemsetAsync(dv.as<F>() + 2 * (size_t)n, 0, (size_t)2 * n * 32, st));
  BZ_CUDA(cudaMemsetAsync(derr.p, 0, 4, st));
  if (n > LKP_MAX_N) lookup_permute_large_run(C, dv.p, dv.as<F>() + n, dv.as<F>() + 2 * (size_t)n, dv.as<F>() + 3 * (size_t)n, usable, (uint32_t*)derr.p);
  else {
    static PerDeviceOnce once;
    once.run(C->device, [] { cudaFuncSetAttribute(lookup_permute_kernel<FpP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lookup_permute_smem(LKP_MAX_N)); });
    LookupPermDesc d{PolyRef{0, 0}, PolyRef{0, 1}, PolyRef{0, 2}, PolyRef{0, 3}};
    BZ_CUDA(cudaMemcpyAsync(dd.p, &d, sizeof(d), cudaMemcpyHostToDevice, st));
    Regions reg; memset(&reg, 0, sizeof(reg));
    reg.base[0] = dv.p;
    lookup_permute_kernel<FpP><<<dim3(1, 1), LKP_THREADS, lookup_permute_smem(n), st>>>(reg, n, usable, dd.as<LookupPermDesc>(), dso.as<F>(), (uint32_t*)derr.p);
    C->kernel_launches++;
  }
  uint32_t err = 0;
  BZ_CUDA(cudaMemcpyAsync(&err, derr.p, 4, cudaMemcpyDeviceToHost, st));
  BZ_CUDA(cudaMemcpyAsync(out_in, dv.as<F>() + 2 * (size_t)n, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
  BZ_CUDA(cudaMemcpyAsync(out_tab, dv.as<F>() + 3 * (size_t)n, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
  BZ_CUDA(cudaStreamSynchronize(st));
  if (err) throw Error(BZ_ERR_SYNTHESIS, "lookup input not in table (Error::ConstraintSystemFailure)");
}

}  // namespace
}  // namespace bz

using namespace bz;

#define FG_FIELD(field_, CALL)                                        \
  BZ_CHECK((field_) == 0 || (field_) == 1, "bad field id");           \
  if ((field_) == 0) { typedef FpP PP; CALL; } else { typedef FqP PP; CALL; }

extern "C" {

API int bz_perm_product(bz_ctx* ctx, int field, uint32_t k, uint32_t ncols, const void* const* values, const void* const* sigmas, const void* beta,
                        const void* gamma, const void* delta_omega0, const void* z0, void* out_z) {
  if (!ctx) return BZ_ERR_INVALID;
  PV_TRY(ctx, {
    BZ_CHECK(values && sigmas && beta && gamma && delta_omega0 && z0 && out_z, "null argument");
    BZ_CHECK(k >= 1 && k <= 26, "k out of range");
    if (ncols < 1 || ncols > 8) throw Error(BZ_ERR_UNSUPPORTED, "permutation set wider than 8 columns (cs degree > 10)");
    for (uint32_t j = 0; j < ncols; ++j) BZ_CHECK(values[j] && sigmas[j], "null column");
    FG_FIELD(field, perm_product_t<PP>(C, k, ncols, values, sigmas, beta, gamma, delta_omega0, z0, out_z));
  });
}

API int bz_lookup_permute(bz_ctx* ctx, int field, uint32_t k, uint32_t usable_rows, const void* compressed_input, const void* compressed_table,
                          void* out_permuted_input, void* out_permuted_table) {
  if (!ctx) return BZ_ERR_INVALID;
  PV_TRY(ctx, {
    BZ_CHECK(compressed_input && compressed_table && out_permuted_input && out_permuted_table, "null argument");
    BZ_CHECK(k >= 1 && k <= 26 && usable_rows <= (1u << k), "k / usable rows out of range");
    if (field != 0) throw Error(BZ_ERR_UNSUPPORTED, "lookup permutation: only Fp (circuits over pallas::Base) is wired up");
    lookup_permute_fp(C, k, usable_rows, compressed_input, compressed_table, out_permuted_input, out_permuted_table);
  });
}

API int bz_lookup_product(bz_ctx* ctx, int field, uint32_t k, const void* compressed_input, const void* compressed_table, const void* permuted_input,
                          const void* permuted_table, const void* beta, const void* gamma, void* out_z) {
  if (!ctx) return BZ_ERR_INVALID;
  PV_TRY(ctx, {
    BZ_CHECK(compressed_input && compressed_table && permuted_input && permuted_table && beta && gamma && out_z, "null argument");
    BZ_CHECK(k >= 1 && k <= 26, "k out of range");
    FG_FIELD(field, lookup_product_t<PP>(C, k, compressed_input, compressed_table, permuted_input, permuted_table, beta, gamma, out_z));
  });
}

API int bz_divide_by_vanishing(bz_ctx* ctx, int field, uint32_t k, uint32_t extended_k, void* a) {
  if (!ctx) return BZ_ERR_INVALID;
  PV_TRY(ctx, {
    BZ_CHECK(a, "null argument");
    BZ_CHECK(extended_k > k && extended_k <= 30 && extended_k - k <= 16, "divide_by_vanishing_poly needs extended_k > k");
    FG_FIELD(field, divide_by_vanishing_t<PP>(C, k, extended_k, a));
  });
}

API int bz_eval_many(bz_ctx* ctx, int field, uint64_t n, uint32_t count, const void* const* polys, const void* points, void* out) {
  if (!ctx) return BZ_ERR_INVALID;
  PV_TRY(ctx, {
    if (!count) return BZ_OK;
    BZ_CHECK(polys && points && out, "null argument");
    for (uint32_t i = 0; i < count; ++i) BZ_CHECK(polys[i], "null polynomial");
    FG_FIELD(field, eval_many_t<PP>(C, n, count, polys, points, out));
  });
}

API int bz_kate_div(bz_ctx* ctx, int field, uint64_t n, const void* a, const void* point, void* out_q) {
  if (!ctx) return BZ_ERR_INVALID;
  PV_TRY(ctx, {
    BZ_CHECK(a && point && out_q, "null argument");
    FG_FIELD(field, kate_div_t<PP>(C, n, a, point, out_q));
  });
}

API int bz_axpy(bz_ctx* ctx, int field, uint64_t n, void* acc, const void* x, const void* poly) {
  if (!ctx) return BZ_ERR_INVALID;
  PV_TRY(ctx, {
    if (!n) return BZ_OK;
    BZ_CHECK(acc && x && poly, "null argument");
    FG_FIELD(field, axpy_t<PP>(C, n, acc, x, poly));
  });
}

}  // extern "C"
