/* ORACLE -- test infrastructure only (imported by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs; NEVER by the product).
 * ctypes-facing exports of the C restatement of halo2_proofs 0.2.0 arithmetic over pasta 0.4.1.
 * All field elements cross this API in pasta's in-memory form: 32 B little-endian Montgomery
 * ([u64;4], R = 2^256); affine points 64 B x||y (identity = zeros); Jacobian 96 B x||y||z. */
#define ORACLE_THREADS_IMPL
#include "pasta.h"

#define API __attribute__((visibility("default")))

API void orc_init(void) { pasta_init(); }
API void orc_set_threads(int n) { oracle_set_num_threads(n); }
API int  orc_get_threads(void) { return oracle_num_threads(); }

/* field: curve_id 0 = Fp, 1 = Fq */
API void orc_field_mul(int f, const void* a, const void* b, void* r, size_t n) {
  pasta_init();
  for (size_t i = 0; i < n; ++i) {
    if (f == 0) fp_mul((fp_t*)r + i, (const fp_t*)a + i, (const fp_t*)b + i);
    else fq_mul((fq_t*)r + i, (const fq_t*)a + i, (const fq_t*)b + i);
  }
}
API void orc_field_from_u512(int f, const void* wide, void* r, size_t n) {
  pasta_init();
  for (size_t i = 0; i < n; ++i) {
    if (f == 0) fp_from_u512((fp_t*)r + i, (const uint64_t*)wide + 8 * i);
    else fq_from_u512((fq_t*)r + i, (const uint64_t*)wide + 8 * i);
  }
}
API void orc_field_to_repr(int f, const void* a, void* out, size_t n) {
  pasta_init();
  for (size_t i = 0; i < n; ++i) {
    if (f == 0) fp_to_repr((uint8_t*)out + 32 * i, (const fp_t*)a + i);
    else fq_to_repr((uint8_t*)out + 32 * i, (const fq_t*)a + i);
  }
}
API void orc_field_from_repr(int f, const void* in, void* out, size_t n) {
  pasta_init();
  for (size_t i = 0; i < n; ++i) {
    uint64_t v[4]; memcpy(v, (const uint8_t*)in + 32 * i, 32);
    if (f == 0) fp_from_raw((fp_t*)out + i, v); else fq_from_raw((fq_t*)out + i, v);
  }
}
API void orc_field_inv(int f, const void* a, void* r, size_t n) {
  pasta_init();
  for (size_t i = 0; i < n; ++i) {
    if (f == 0) fp_inv((fp_t*)r + i, (const fp_t*)a + i); else fq_inv((fq_t*)r + i, (const fq_t*)a + i);
  }
}

/* curve: 0 = Vesta (scalars Fp, coordinates Fq), 1 = Pallas (scalars Fq, coordinates Fp) */
API void orc_best_multiexp(int curve, const void* scalars, const void* bases, size_t n, void* out_jac) {
  pasta_init();
  if (curve == 0) vesta_best_multiexp((vesta_point*)out_jac, (const fp_t*)scalars, (const vesta_affine*)bases, n);
  else pallas_best_multiexp((pallas_point*)out_jac, (const fq_t*)scalars, (const pallas_affine*)bases, n);
}
API void orc_to_affine(int curve, const void* jac, void* aff, size_t n) {
  pasta_init();
  if (curve == 0) vesta_batch_normalize((vesta_affine*)aff, (const vesta_point*)jac, n);
  else pallas_batch_normalize((pallas_affine*)aff, (const pallas_point*)jac, n);
}
API void orc_point_mul(int curve, const void* aff, const void* scalar, void* out_aff) {
  pasta_init();
  if (curve == 0) { vesta_point p, r; vesta_from_affine(&p, (const vesta_affine*)aff); vesta_mul(&r, &p, (const fp_t*)scalar); vesta_to_affine((vesta_affine*)out_aff, &r); }
  else { pallas_point p, r; pallas_from_affine(&p, (const pallas_affine*)aff); pallas_mul(&r, &p, (const fq_t*)scalar); pallas_to_affine((pallas_affine*)out_aff, &r); }
}
API void orc_point_add(int curve, const void* a_aff, const void* b_aff, void* out_aff) {
  pasta_init();
  if (curve == 0) { vesta_point p; vesta_from_affine(&p, (const vesta_affine*)a_aff); vesta_add_mixed(&p, &p, (const vesta_affine*)b_aff); vesta_to_affine((vesta_affine*)out_aff, &p); }
  else { pallas_point p; pallas_from_affine(&p, (const pallas_affine*)a_aff); pallas_add_mixed(&p, &p, (const pallas_affine*)b_aff); pallas_to_affine((pallas_affine*)out_aff, &p); }
}
/* field: 0 = Fp, 1 = Fq.  In-place best_fft, natural order in/out. */
API void orc_best_fft(int f, void* a, const void* omega, unsigned log_n) {
  pasta_init();
  if (f == 0) fp_fft_best_fft((fp_t*)a, (const fp_t*)omega, log_n);
  else fq_fft_best_fft((fq_t*)a, (const fq_t*)omega, log_n);
}
/* EC best_fft (Params::new's g_lagrange): curve 0 = Vesta points with Fp scalars */
API void orc_best_fft_ec(int curve, void* a_jac, const void* omega, unsigned log_n) {
  pasta_init();
  if (curve == 0) vesta_fft_best_fft((vesta_point*)a_jac, (const fp_t*)omega, log_n);
  else pallas_fft_best_fft((pallas_point*)a_jac, (const fq_t*)omega, log_n);
}
