/* ORACLE -- test infrastructure only.
 * Restates halo2_proofs 0.2.0 `arithmetic.rs::{multiexp_serial, best_multiexp}` (U:, SURVEY §8 a2,
 * App. D): per-thread contiguous chunks of len/num_threads; per chunk Pippenger with
 * c = 1 (n<4) | 3 (n<32) | ceil(ln n), segments = 256/c + 1, unsigned c-bit digits of the
 * canonical little-endian repr, 2^c - 1 buckets {None, Affine, Projective}, running-sum
 * reduction, c doublings per segment; chunk results folded in order.
 * Include with EC(name), BF(name), SF(name). */
#include <math.h>

typedef struct { int kind; EC(affine) a; EC(point) p; } EC(bucket);   /* kind 0 none, 1 affine, 2 projective */

static inline size_t EC(get_at)(size_t segment, size_t c, const uint8_t bytes[32]) {
  size_t skip_bits = segment * c, skip_bytes = skip_bits / 8;
  if (skip_bytes >= 32) return 0;
  uint8_t v[8] = {0};
  for (size_t i = 0; i < 8 && skip_bytes + i < 32; ++i) v[i] = bytes[skip_bytes + i];
  uint64_t tmp; memcpy(&tmp, v, 8);
  tmp >>= skip_bits - skip_bytes * 8;
  tmp %= ((uint64_t)1 << c);
  return (size_t)tmp;
}

static void EC(multiexp_serial)(const SF(t)* coeffs, const EC(affine)* bases, size_t n, EC(point)* acc) {
  uint8_t* reprs = (uint8_t*)malloc(32 * (n ? n : 1));
  for (size_t i = 0; i < n; ++i) SF(to_repr)(reprs + 32 * i, &coeffs[i]);
  size_t c;
  if (n < 4) c = 1; else if (n < 32) c = 3; else c = (size_t)ceil(log((double)(uint32_t)n));
  size_t segments = 256 / c + 1, nb = ((size_t)1 << c) - 1;
  EC(bucket)* buckets = (EC(bucket)*)malloc(sizeof(EC(bucket)) * nb);
  for (size_t seg = segments; seg-- > 0;) {
    for (size_t i = 0; i < c; ++i) EC(dbl)(acc, acc);
    for (size_t b = 0; b < nb; ++b) buckets[b].kind = 0;
    for (size_t i = 0; i < n; ++i) {
      size_t d = EC(get_at)(seg, c, reprs + 32 * i);
      if (!d) continue;
      EC(bucket)* bk = &buckets[d - 1];
      if (bk->kind == 0) { bk->a = bases[i]; bk->kind = 1; }
      else if (bk->kind == 1) { EC(from_affine)(&bk->p, &bk->a); EC(add_mixed)(&bk->p, &bk->p, &bases[i]); bk->kind = 2; }
      else EC(add_mixed)(&bk->p, &bk->p, &bases[i]);
    }
    EC(point) running; EC(set_identity)(&running);
    for (size_t b = nb; b-- > 0;) {
      if (buckets[b].kind == 1) EC(add_mixed)(&running, &running, &buckets[b].a);
      else if (buckets[b].kind == 2) EC(add)(&running, &running, &buckets[b].p);
      EC(add)(acc, acc, &running);
    }
  }
  free(buckets); free(reprs);
}

typedef struct { const SF(t)* coeffs; const EC(affine)* bases; size_t n, chunk; EC(point)* results; } EC(msm_job);
static void EC(msm_task)(int t, void* c) {
  EC(msm_job)* j = (EC(msm_job)*)c;
  size_t lo = (size_t)t * j->chunk, hi = lo + j->chunk; if (hi > j->n) hi = j->n;
  EC(set_identity)(&j->results[t]);
  EC(multiexp_serial)(j->coeffs + lo, j->bases + lo, hi - lo, &j->results[t]);
}

static void EC(best_multiexp)(EC(point)* out, const SF(t)* coeffs, const EC(affine)* bases, size_t n) {
  size_t threads = (size_t)oracle_num_threads();
  EC(set_identity)(out);
  if (n > threads) {
    size_t chunk = n / threads, nchunks = (n + chunk - 1) / chunk;
    EC(point)* results = (EC(point)*)malloc(sizeof(EC(point)) * nchunks);
    EC(msm_job) j = { coeffs, bases, n, chunk, results };
    par_run((int)nchunks, EC(msm_task), &j);
    for (size_t i = 0; i < nchunks; ++i) EC(add)(out, out, &results[i]);
    free(results);
  } else {
    EC(multiexp_serial)(coeffs, bases, n, out);
  }
}
