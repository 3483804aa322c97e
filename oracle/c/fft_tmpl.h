/* ORACLE -- test infrastructure only.
 * Restates halo2_proofs 0.2.0 `arithmetic.rs::best_fft` + `recursive_butterfly_arithmetic`
 * (U:): bit-reversal permutation, serial twiddle table of n/2 powers,
 * then either the iterative radix-2 loop (log_n <= log2(threads)) or the recursive split whose
 * top log2(threads) levels fork.  Here the fork is realised as: 2^log_threads independent
 * sub-transforms run in parallel, followed by the top levels with their butterflies chunked over
 * the pool -- the same butterflies in the same dependency order.
 *
 * Include with: FFT(name) prefix, FFT_ELEM element type, FFT_SCALAR scalar type (SF(t)),
 *   FFT_ADD(r,a,b) FFT_SUB(r,a,b) FFT_SCALE(r,a,s)  and SF(name) scalar-field prefix. */

static void FFT(serial_rec)(FFT_ELEM* a, size_t n, size_t twiddle_chunk, const FFT_SCALAR* tw) {
  if (n == 2) {
    FFT_ELEM t = a[1];
    a[1] = a[0];
    FFT_ADD(&a[0], &a[0], &t);
    FFT_SUB(&a[1], &a[1], &t);
    return;
  }
  size_t h = n / 2;
  FFT(serial_rec)(a, h, twiddle_chunk * 2, tw);
  FFT(serial_rec)(a + h, h, twiddle_chunk * 2, tw);
  {
    FFT_ELEM t = a[h];
    a[h] = a[0];
    FFT_ADD(&a[0], &a[0], &t);
    FFT_SUB(&a[h], &a[h], &t);
  }
  for (size_t i = 1; i < h; ++i) {
    FFT_ELEM t;
    FFT_SCALE(&t, &a[h + i], &tw[i * twiddle_chunk]);
    a[h + i] = a[i];
    FFT_ADD(&a[i], &a[i], &t);
    FFT_SUB(&a[h + i], &a[h + i], &t);
  }
}

typedef struct { FFT_ELEM* a; size_t n, sub_n, twiddle_chunk; const FFT_SCALAR* tw; size_t level_n; int nchunks; } FFT(job);

static void FFT(sub_task)(int t, void* c) {
  FFT(job)* j = (FFT(job)*)c;
  FFT(serial_rec)(j->a + (size_t)t * j->sub_n, j->sub_n, j->twiddle_chunk, j->tw);
}
static void FFT(level_task)(int t, void* c) {
  FFT(job)* j = (FFT(job)*)c;
  /* butterflies of one level, all blocks, split into nchunks contiguous ranges of butterfly ids */
  size_t h = j->level_n / 2, total = j->n / 2;
  size_t lo = total * (size_t)t / j->nchunks, hi = total * (size_t)(t + 1) / j->nchunks;
  for (size_t b = lo; b < hi; ++b) {
    size_t blk = b / h, i = b % h;
    FFT_ELEM* x = j->a + blk * j->level_n;
    FFT_ELEM t2;
    if (i == 0) t2 = x[h]; else FFT_SCALE(&t2, &x[h + i], &j->tw[i * j->twiddle_chunk]);
    x[h + i] = x[i];
    FFT_ADD(&x[i], &x[i], &t2);
    FFT_SUB(&x[h + i], &x[h + i], &t2);
  }
}

static void FFT(best_fft)(FFT_ELEM* a, const FFT_SCALAR* omega, unsigned log_n) {
  size_t n = (size_t)1 << log_n;
  int threads = oracle_num_threads();
  unsigned log_threads = 0; while ((2u << log_threads) <= (unsigned)threads) ++log_threads;
  for (size_t k = 0; k < n; ++k) {
    size_t rk = 0, x = k;
    for (unsigned b = 0; b < log_n; ++b) { rk = (rk << 1) | (x & 1); x >>= 1; }
    if (k < rk) { FFT_ELEM t = a[rk]; a[rk] = a[k]; a[k] = t; }
  }
  if (n < 2) return;
  FFT_SCALAR* tw = (FFT_SCALAR*)malloc(sizeof(FFT_SCALAR) * (n / 2));
  { FFT_SCALAR w = SF(R1); for (size_t i = 0; i < n / 2; ++i) { tw[i] = w; SF(mul)(&w, &w, omega); } }
  if (log_n <= log_threads || threads == 1) {
    if (threads == 1) FFT(serial_rec)(a, n, 1, tw);
    else {
      size_t chunk = 2, twiddle_chunk = n / 2;
      for (unsigned s = 0; s < log_n; ++s) {
        for (size_t base = 0; base < n; base += chunk) {
          FFT_ELEM* x = a + base; size_t h = chunk / 2;
          for (size_t i = 0; i < h; ++i) {
            FFT_ELEM t;
            if (i == 0) t = x[h]; else FFT_SCALE(&t, &x[h + i], &tw[i * twiddle_chunk]);
            x[h + i] = x[i];
            FFT_ADD(&x[i], &x[i], &t);
            FFT_SUB(&x[h + i], &x[h + i], &t);
          }
        }
        chunk *= 2; twiddle_chunk /= 2;
      }
    }
  } else {
    FFT(job) j; j.a = a; j.n = n; j.tw = tw; j.nchunks = threads;
    j.sub_n = n >> log_threads; j.twiddle_chunk = (size_t)1 << log_threads;
    par_run(1 << log_threads, FFT(sub_task), &j);
    for (unsigned lvl = log_threads; lvl-- > 0;) {
      j.level_n = n >> lvl; j.twiddle_chunk = (size_t)1 << lvl;
      par_run(threads, FFT(level_task), &j);
    }
  }
  free(tw);
}
