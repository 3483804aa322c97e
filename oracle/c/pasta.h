/* ORACLE -- test infrastructure only.  Instantiates the field / curve / MSM / FFT templates for
 * the Pasta cycle (SURVEY App. B constants): Fp, Fq, Vesta (base Fq, scalar Fp), Pallas (base Fp,
 * scalar Fq). */
#ifndef ORACLE_PASTA_H
#define ORACLE_PASTA_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "threads.h"

/* ---- Fp ---- */
#define FE(name) fp_##name
#define FE_MOD0 0x992d30ed00000001ULL
#define FE_MOD1 0x224698fc094cf91bULL
#define FE_MOD2 0x0000000000000000ULL
#define FE_MOD3 0x4000000000000000ULL
#define FE_INV  0x992d30ecffffffffULL
#include "field_tmpl.h"
#undef FE
#undef FE_MOD0
#undef FE_MOD1
#undef FE_MOD2
#undef FE_MOD3
#undef FE_INV

/* ---- Fq ---- */
#define FE(name) fq_##name
#define FE_MOD0 0x8c46eb2100000001ULL
#define FE_MOD1 0x224698fc0994a8ddULL
#define FE_MOD2 0x0000000000000000ULL
#define FE_MOD3 0x4000000000000000ULL
#define FE_INV  0x8c46eb20ffffffffULL
#include "field_tmpl.h"
#undef FE
#undef FE_MOD0
#undef FE_MOD1
#undef FE_MOD2
#undef FE_MOD3
#undef FE_INV

/* ---- Vesta: base Fq, scalar Fp (the commitment curve of this prover) ---- */
#define EC(name) vesta_##name
#define BF(name) fq_##name
#define SF(name) fp_##name
#include "curve_tmpl.h"
#include "msm_tmpl.h"
#define FFT(name) vesta_fft_##name
#define FFT_ELEM vesta_point
#define FFT_SCALAR fp_t
#define FFT_ADD(r, a, b) vesta_add(r, a, b)
#define FFT_SUB(r, a, b) vesta_sub(r, a, b)
#define FFT_SCALE(r, a, s) vesta_mul(r, a, s)
#include "fft_tmpl.h"
#undef FFT
#undef FFT_ELEM
#undef FFT_SCALAR
#undef FFT_ADD
#undef FFT_SUB
#undef FFT_SCALE
/* field FFT over Fp */
#define FFT(name) fp_fft_##name
#define FFT_ELEM fp_t
#define FFT_SCALAR fp_t
#define FFT_ADD(r, a, b) fp_add(r, a, b)
#define FFT_SUB(r, a, b) fp_sub(r, a, b)
#define FFT_SCALE(r, a, s) fp_mul(r, a, s)
#include "fft_tmpl.h"
#undef FFT
#undef FFT_ELEM
#undef FFT_SCALAR
#undef FFT_ADD
#undef FFT_SUB
#undef FFT_SCALE
#undef EC
#undef BF
#undef SF

/* ---- Pallas: base Fp, scalar Fq ---- */
#define EC(name) pallas_##name
#define BF(name) fp_##name
#define SF(name) fq_##name
#include "curve_tmpl.h"
#include "msm_tmpl.h"
#define FFT(name) pallas_fft_##name
#define FFT_ELEM pallas_point
#define FFT_SCALAR fq_t
#define FFT_ADD(r, a, b) pallas_add(r, a, b)
#define FFT_SUB(r, a, b) pallas_sub(r, a, b)
#define FFT_SCALE(r, a, s) pallas_mul(r, a, s)
#include "fft_tmpl.h"
#undef FFT
#undef FFT_ELEM
#undef FFT_SCALAR
#undef FFT_ADD
#undef FFT_SUB
#undef FFT_SCALE
#define FFT(name) fq_fft_##name
#define FFT_ELEM fq_t
#define FFT_SCALAR fq_t
#define FFT_ADD(r, a, b) fq_add(r, a, b)
#define FFT_SUB(r, a, b) fq_sub(r, a, b)
#define FFT_SCALE(r, a, s) fq_mul(r, a, s)
#include "fft_tmpl.h"
#undef FFT
#undef FFT_ELEM
#undef FFT_SCALAR
#undef FFT_ADD
#undef FFT_SUB
#undef FFT_SCALE
#undef EC
#undef BF
#undef SF

static inline void pasta_init(void) {
  static int done = 0;
  if (done) return;
  fp_init(); fq_init(); done = 1;
}
#endif
