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

/* ---------------------------------------------------------------------------------------------
 * Vector helpers for the restated prover (oracle/halo2.py).  `parallelize` = halo2_proofs 0.2.0
 * arithmetic.rs::parallelize: chunk = len / num_threads contiguous pieces on the pool.
 * Only Fp/Fq element-wise work; f: 0 = Fp, 1 = Fq.
 * ------------------------------------------------------------------------------------------- */
typedef struct { int f, op; const void *a, *b; void* r; size_t n; int nchunks; } vec_job;

static void vec_task(int t, void* c) {
  vec_job* j = (vec_job*)c;
  size_t lo = j->n * (size_t)t / j->nchunks, hi = j->n * (size_t)(t + 1) / j->nchunks;
#define VEC_LOOP(T, PFX)                                                                             \
  {                                                                                                  \
    const T* a = (const T*)j->a; const T* b = (const T*)j->b; T* r = (T*)j->r;                       \
    for (size_t i = lo; i < hi; ++i) {                                                               \
      switch (j->op) {                                                                               \
        case 0: PFX##_mul(&r[i], &a[i], &b[i]); break;                                               \
        case 1: PFX##_add(&r[i], &a[i], &b[i]); break;                                               \
        case 2: PFX##_sub(&r[i], &a[i], &b[i]); break;                                               \
        case 3: PFX##_mul(&r[i], &a[i], &b[0]); break;   /* scale by scalar b[0] */                  \
        case 4: PFX##_neg(&r[i], &a[i]); break;                                                      \
        case 5: PFX##_add(&r[i], &a[i], &b[0]); break;   /* add scalar */                            \
        default: break;                                                                              \
      }                                                                                              \
    }                                                                                                \
  }
  if (j->f == 0) VEC_LOOP(fp_t, fp) else VEC_LOOP(fq_t, fq)
#undef VEC_LOOP
}

/* op: 0 mul, 1 add, 2 sub, 3 scale (b = one scalar), 4 neg, 5 add scalar */
API void orc_vec_op(int f, int op, const void* a, const void* b, void* r, size_t n) {
  pasta_init();
  vec_job j = { f, op, a, b, r, n, oracle_num_threads() };
  if (n < (size_t)j.nchunks * 4) j.nchunks = 1;
  par_run(j.nchunks, vec_task, &j);
}

/* ff::BatchInvert over a slice (zeros skipped), serial like the reference's `batch_invert()` */
API void orc_batch_invert(int f, void* v, size_t n) {
  pasta_init();
  void* scratch = malloc(32 * (n ? n : 1));
  if (f == 0) fp_batch_invert((fp_t*)v, n, (fp_t*)scratch); else fq_batch_invert((fq_t*)v, n, (fq_t*)scratch);
  free(scratch);
}

/* arithmetic::eval_polynomial (serial Horner from the top coefficient) */
API void orc_eval_polynomial(int f, const void* poly, size_t n, const void* point, void* out) {
  pasta_init();
  if (f == 0) {
    const fp_t* p = (const fp_t*)poly; fp_t acc = {{0, 0, 0, 0}};
    for (size_t i = n; i-- > 0;) { fp_mul(&acc, &acc, (const fp_t*)point); fp_add(&acc, &acc, &p[i]); }
    *(fp_t*)out = acc;
  } else {
    const fq_t* p = (const fq_t*)poly; fq_t acc = {{0, 0, 0, 0}};
    for (size_t i = n; i-- > 0;) { fq_mul(&acc, &acc, (const fq_t*)point); fq_add(&acc, &acc, &p[i]); }
    *(fq_t*)out = acc;
  }
}

/* arithmetic::compute_inner_product (serial) */
API void orc_inner_product(int f, const void* a, const void* b, size_t n, void* out) {
  pasta_init();
  if (f == 0) {
    fp_t acc = {{0, 0, 0, 0}}, t;
    for (size_t i = 0; i < n; ++i) { fp_mul(&t, (const fp_t*)a + i, (const fp_t*)b + i); fp_add(&acc, &acc, &t); }
    *(fp_t*)out = acc;
  } else {
    fq_t acc = {{0, 0, 0, 0}}, t;
    for (size_t i = 0; i < n; ++i) { fq_mul(&t, (const fq_t*)a + i, (const fq_t*)b + i); fq_add(&acc, &acc, &t); }
    *(fq_t*)out = acc;
  }
}

/* arithmetic::kate_division(a, b): quotient of a(X) by (X - b), remainder dropped; out has n-1 coefficients */
API void orc_kate_division(int f, const void* a, size_t n, const void* b, void* out) {
  pasta_init();
  if (n < 2) return;
  if (f == 0) {
    const fp_t* p = (const fp_t*)a; fp_t* q = (fp_t*)out; fp_t tmp = {{0, 0, 0, 0}};
    for (size_t i = n - 1; i-- > 0;) {   /* q[i] = a[i+1] + b*q[i+1] */
      fp_t lead; fp_add(&lead, &p[i + 1], &tmp);
      q[i] = lead;
      fp_mul(&tmp, &lead, (const fp_t*)b);
    }
  } else {
    const fq_t* p = (const fq_t*)a; fq_t* q = (fq_t*)out; fq_t tmp = {{0, 0, 0, 0}};
    for (size_t i = n - 1; i-- > 0;) {
      fq_t lead; fq_add(&lead, &p[i + 1], &tmp);
      q[i] = lead;
      fq_mul(&tmp, &lead, (const fq_t*)b);
    }
  }
}

/* serial running product used by the permutation / lookup grand products:
 * z[0] = z0; z[i] = z[i-1] * frac[i-1] for i in 1..n   (U: plonk/permutation/prover.rs, plonk/lookup/prover.rs) */
API void orc_running_product(int f, const void* z0, const void* frac, void* z, size_t n) {
  pasta_init();
  if (f == 0) {
    fp_t* zz = (fp_t*)z; const fp_t* fr = (const fp_t*)frac;
    zz[0] = *(const fp_t*)z0;
    for (size_t i = 1; i < n; ++i) fp_mul(&zz[i], &zz[i - 1], &fr[i - 1]);
  } else {
    fq_t* zz = (fq_t*)z; const fq_t* fr = (const fq_t*)frac;
    zz[0] = *(const fq_t*)z0;
    for (size_t i = 1; i < n; ++i) fq_mul(&zz[i], &zz[i - 1], &fr[i - 1]);
  }
}

/* powers: out[i] = base^i * first, i < n (serial, as the reference builds `b` / deltaomega tables) */
API void orc_powers(int f, const void* first, const void* base, void* out, size_t n) {
  pasta_init();
  if (f == 0) { fp_t cur = *(const fp_t*)first; for (size_t i = 0; i < n; ++i) { ((fp_t*)out)[i] = cur; fp_mul(&cur, &cur, (const fp_t*)base); } }
  else { fq_t cur = *(const fq_t*)first; for (size_t i = 0; i < n; ++i) { ((fq_t*)out)[i] = cur; fq_mul(&cur, &cur, (const fq_t*)base); } }
}

/* parallel_generator_collapse(g, challenge): g_lo[i] += [challenge] g_hi[i]; affine in/out (Vesta/Pallas) */
typedef struct { int curve; void* g; size_t half; const void* ch; int nchunks; } collapse_job;
static void collapse_task(int t, void* c) {
  collapse_job* j = (collapse_job*)c;
  size_t lo = j->half * (size_t)t / j->nchunks, hi = j->half * (size_t)(t + 1) / j->nchunks;
  if (hi <= lo) return;
  if (j->curve == 0) {
    vesta_affine* g = (vesta_affine*)j->g;
    vesta_point* tmp = (vesta_point*)malloc(sizeof(vesta_point) * (hi - lo));
    for (size_t i = lo; i < hi; ++i) {
      vesta_point h, r; vesta_from_affine(&h, &g[i + j->half]); vesta_mul(&r, &h, (const fp_t*)j->ch);
      vesta_add_mixed(&r, &r, &g[i]); tmp[i - lo] = r;
    }
    vesta_batch_normalize(g + lo, tmp, hi - lo); free(tmp);
  } else {
    pallas_affine* g = (pallas_affine*)j->g;
    pallas_point* tmp = (pallas_point*)malloc(sizeof(pallas_point) * (hi - lo));
    for (size_t i = lo; i < hi; ++i) {
      pallas_point h, r; pallas_from_affine(&h, &g[i + j->half]); pallas_mul(&r, &h, (const fq_t*)j->ch);
      pallas_add_mixed(&r, &r, &g[i]); tmp[i - lo] = r;
    }
    pallas_batch_normalize(g + lo, tmp, hi - lo); free(tmp);
  }
}
API void orc_generator_collapse(int curve, void* g_affine, size_t len, const void* challenge) {
  pasta_init();
  collapse_job j = { curve, g_affine, len / 2, challenge, oracle_num_threads() };
  par_run(j.nchunks, collapse_task, &j);
}

/* sort keys for the lookup argument: canonical big-endian image so memcmp order == numeric order (pasta Ord) */
API void orc_canonical_be(int f, const void* a, void* out, size_t n) {
  pasta_init();
  for (size_t i = 0; i < n; ++i) {
    uint8_t le[32];
    if (f == 0) fp_to_repr(le, (const fp_t*)a + i); else fq_to_repr(le, (const fq_t*)a + i);
    for (int k = 0; k < 32; ++k) ((uint8_t*)out)[32 * i + k] = le[31 - k];
  }
}
