/* ORACLE -- test infrastructure only.
 * Minimal fork-join pool standing in for rayon's `multicore::scope` / `parallelize`
 * (halo2_proofs 0.2.0 `src/multicore.rs`, `src/arithmetic.rs::parallelize`). */
#ifndef ORACLE_THREADS_H
#define ORACLE_THREADS_H
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <stdint.h>
#include <unistd.h>

typedef void (*par_fn)(int task, void* ctx);

int  oracle_num_threads(void);
void oracle_set_num_threads(int n);
/* run fn(0..ntasks-1) across the pool, return when all done */
void par_run(int ntasks, par_fn fn, void* ctx);

#ifdef ORACLE_THREADS_IMPL
static int g_nthreads = 0;
static pthread_t* g_workers = NULL;
static int g_nworkers = 0;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_cv_start = PTHREAD_COND_INITIALIZER, g_cv_done = PTHREAD_COND_INITIALIZER;
static unsigned long g_gen = 0;
static par_fn g_fn; static void* g_ctx; static int g_ntasks; static atomic_int g_next; static int g_active = 0;
static __thread int t_in_pool = 0;

static void drain(void) {
  for (;;) { int t = atomic_fetch_add(&g_next, 1); if (t >= g_ntasks) break; g_fn(t, g_ctx); }
}
static void* worker(void* arg) {
  unsigned long seen = (unsigned long)(uintptr_t)arg; t_in_pool = 1;
  pthread_mutex_lock(&g_mu);
  for (;;) {
    while (g_gen == seen) pthread_cond_wait(&g_cv_start, &g_mu);
    seen = g_gen;
    pthread_mutex_unlock(&g_mu);
    drain();
    pthread_mutex_lock(&g_mu);
    if (--g_active == 0) pthread_cond_signal(&g_cv_done);
  }
  return NULL;
}
int oracle_num_threads(void) {
  if (g_nthreads == 0) {
    const char* e = getenv("ORACLE_NUM_THREADS");
    int n = e ? atoi(e) : (int)sysconf(_SC_NPROCESSORS_ONLN);
    g_nthreads = n > 0 ? n : 1;
  }
  return g_nthreads;
}
void oracle_set_num_threads(int n) { if (n > 0) g_nthreads = n; }
void par_run(int ntasks, par_fn fn, void* ctx) {
  int nt = oracle_num_threads();
  if (ntasks <= 1 || nt <= 1 || t_in_pool) { for (int t = 0; t < ntasks; ++t) fn(t, ctx); return; }
  pthread_mutex_lock(&g_mu);
  while (g_nworkers < nt - 1) {             /* grow lazily; main thread is the nt-th worker */
    g_workers = (pthread_t*)realloc(g_workers, sizeof(pthread_t) * (g_nworkers + 1));
    pthread_create(&g_workers[g_nworkers++], NULL, worker, (void*)(uintptr_t)g_gen);
  }
  g_fn = fn; g_ctx = ctx; g_ntasks = ntasks; atomic_store(&g_next, 0);
  g_active = g_nworkers; ++g_gen;
  pthread_cond_broadcast(&g_cv_start);
  pthread_mutex_unlock(&g_mu);
  t_in_pool = 1; drain(); t_in_pool = 0;
  pthread_mutex_lock(&g_mu);
  while (g_active) pthread_cond_wait(&g_cv_done, &g_mu);
  pthread_mutex_unlock(&g_mu);
}
#endif
#endif
