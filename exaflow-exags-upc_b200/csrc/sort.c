/* sort.c -- stable parallel LSD radix sort used by gs_setup.
 *
 * The reference sorts 24-byte rows twice with a single-threaded merge/8-bit
 * radix hybrid (src/sort_imp.h:446-509) and spends ~0.6 us per DOF doing so
 * (SURVEY.md 3.2).  Here keys and payloads are split into separate arrays
 * (12 bytes per row), digits are 11 bits wide, only the significant key bits
 * are visited, and every pass is spread over the host cores with OpenMP.
 * Stability is what gs_setup relies on: equal labels keep ascending local
 * index, exactly like the reference's stable sorts.
 */
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "exags_internal.h"

#define DIGIT_BITS 11
#define NBINS (1u << DIGIT_BITS)

/* Threads for the setup phase.  gs_setup is a collective bulk phase, so each
   rank takes its share of the node's cores regardless of OMP_NUM_THREADS
   (launchers such as torchrun force that to 1); EXAGS_SETUP_THREADS overrides. */
int xg_setup_threads(void)
{
#ifdef _OPENMP
  static int cached = 0;
  if (!cached) {
    const char *e = getenv("EXAGS_SETUP_THREADS");
    int t = e ? atoi(e) : 0;
    if (t <= 0) {
      const char *lw = getenv("LOCAL_WORLD_SIZE"), *ws = getenv("WORLD_SIZE");
      int ranks = lw ? atoi(lw) : ws ? atoi(ws) : 1;
      if (ranks < 1) ranks = 1;
      t = omp_get_num_procs() / ranks;
    }
    if (t < 1) t = 1;
    if (t > 64) t = 64;
    cached = t;
  }
  return cached;
#else
  return 1;
#endif
}

static int sort_threads(size_t n)
{
#ifdef _OPENMP
  int t = xg_setup_threads();
  size_t cap = n / 65536 + 1;
  if ((size_t)t > cap) t = (int)cap;
  if (t > 64) t = 64;
  return t < 1 ? 1 : t;
#else
  (void)n;
  return 1;
#endif
}

static void radix_pass(const u64 *ki, const u32 *vi, u64 *ko, u32 *vo,
                       size_t n, unsigned shift, int nt, size_t *hist)
{
  /* hist is nt * NBINS counters */
  memset(hist, 0, (size_t)nt * NBINS * sizeof(size_t));
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
  {
#ifdef _OPENMP
    int t = omp_get_thread_num();
#else
    int t = 0;
#endif
    size_t lo = n * (size_t)t / (size_t)nt, hi = n * (size_t)(t + 1) / (size_t)nt;
    size_t *h = hist + (size_t)t * NBINS;
    for (size_t i = lo; i < hi; ++i) ++h[(ki[i] >> shift) & (NBINS - 1)];
#ifdef _OPENMP
#pragma omp barrier
#pragma omp single
#endif
    {
      size_t run = 0;
      for (unsigned d = 0; d < NBINS; ++d)
        for (int s = 0; s < nt; ++s) {
          size_t c = hist[(size_t)s * NBINS + d];
          hist[(size_t)s * NBINS + d] = run;
          run += c;
        }
    }
    for (size_t i = lo; i < hi; ++i) {
      size_t o = h[(ki[i] >> shift) & (NBINS - 1)]++;
      ko[o] = ki[i];
      vo[o] = vi[i];
    }
  }
}

void xsort_u64_u32(u64 *key, u32 *val, size_t n, unsigned bits, u64 *tmpk, u32 *tmpv)
{
  if (n < 2 || bits == 0) return;
  int nt = sort_threads(n);
  size_t *hist = xnew(size_t, (size_t)nt * NBINS);
  u64 *ka = key, *kb = tmpk;
  u32 *va = val, *vb = tmpv;
  for (unsigned shift = 0; shift < bits; shift += DIGIT_BITS) {
    radix_pass(ka, va, kb, vb, n, shift, nt, hist);
    u64 *tk = ka; ka = kb; kb = tk;
    u32 *tv = va; va = vb; vb = tv;
  }
  if (ka != key) {
    memcpy(key, ka, n * sizeof(u64));
    memcpy(val, va, n * sizeof(u32));
  }
  free(hist);
}

static unsigned bits_of(u64 maxkey)
{
  unsigned b = 0;
  while (maxkey) { ++b; maxkey >>= 1; }
  return b ? b : 1;
}

void xsort_perm_u64(u32 *perm, const u64 *key, size_t n, u64 maxkey)
{
  if (n < 2) return;
  u64 *k = xnew(u64, n), *tk = xnew(u64, n);
  u32 *tv = xnew(u32, n);
  for (size_t i = 0; i < n; ++i) k[i] = key[perm[i]];
  xsort_u64_u32(k, perm, n, bits_of(maxkey), tk, tv);
  free(k); free(tk); free(tv);
}

void xsort_perm_u32(u32 *perm, const u32 *key, size_t n, u32 maxkey)
{
  if (n < 2) return;
  u64 *k = xnew(u64, n), *tk = xnew(u64, n);
  u32 *tv = xnew(u32, n);
  for (size_t i = 0; i < n; ++i) k[i] = key[perm[i]];
  xsort_u64_u32(k, perm, n, bits_of(maxkey), tk, tv);
  free(k); free(tk); free(tv);
}
