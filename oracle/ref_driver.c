/* ref_driver.c -- runs the reference library itself (compiled by
 * build_ref.sh from /root/reference/src through the UPC shim) on a job given
 * by the test harness: np ranks as forked processes, gs_setup, a list of
 * gs/gs_vec/gs_many cases, optional timing with the reference's own protocol
 * (src/gs.upc:1094-1111: barrier, 2 warm-ups, N timed calls, mean, max over
 * ranks).  TEST INFRASTRUCTURE ONLY; part of oracle/_ref/libexags_ref.so. */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "c99.h"
#include "name.h"
#include "fail.h"
#include "types.h"
#define gs_op gs_op_t
#include "gs_defs.h"
#include "comm.h"
#include "mem.h"
#include "gs.h"

void shim_arena_create(size_t bytes);
void shim_arena_destroy(void);
void *shim_arena_raw(size_t bytes);
int shim_run(int threads, void (*fn)(int, void *), void *arg);
void shim_barrier(void);

/* layout mirror of the reference's opaque handle head (src/gs.upc:1161-1167),
   only to read the three local map pointers for the golden fixtures */
struct gs_data_head { comm_ptr comm; const uint *map_local[2]; const uint *flagged_primaries; };

struct ref_case { int mode, vn, dom, op, transpose; };

struct job {
  int np, unique, method, ncase, reps;
  const struct ref_case *cases;
  long long **ids; unsigned *n;
  char **data;            /* [ncase*np] arena copies */
  unsigned *maps; size_t maps_cap; size_t *map_len;   /* arena; [np][3] */
  double *times;          /* arena; [ncase][np] */
};

static size_t map_length(const uint *m, int flat)
{
  size_t k = 0;
  if (flat) { while (m[k] != (uint)-1) ++k; return k + 1; }
  while (m[k] != (uint)-1) { while (m[k] != (uint)-1) ++k; ++k; }
  return k + 1;
}

static void rank_main(int r, void *arg)
{
  struct job *J = arg;
  comm_ptr cp;
  static const unsigned esz[5] = { 8, 4, 4, 8, 8 };
  comm_init();
  comm_world(&cp);
  struct gs_data *g = gs_setup((const slong *)J->ids[r], J->n[r], cp, J->unique, (gs_method)J->method, 0);
  if (J->maps) {
    const struct gs_data_head *h = (const struct gs_data_head *)g;
    const uint *m[3] = { h->map_local[0], h->map_local[1], h->flagged_primaries };
    unsigned *out = J->maps + (size_t)r * J->maps_cap;
    size_t o = 0;
    for (int w = 0; w < 3; ++w) {
      size_t len = map_length(m[w], w == 2);
      if (o + len > J->maps_cap) { fprintf(stderr, "ref_driver: map buffer too small\n"); _Exit(3); }
      memcpy(out + o, m[w], len * sizeof(uint));
      J->map_len[r * 3 + w] = len; o += len;
    }
  }
  for (int c = 0; c < J->ncase; ++c) {
    const struct ref_case *k = &J->cases[c];
    char *u = J->data[c * J->np + r];
    void *tab[64];
    if (k->mode == 2) for (int v = 0; v < k->vn; ++v) tab[v] = u + (size_t)v * J->n[r] * esz[k->dom];
    int total = J->reps > 0 ? J->reps + 2 : 1;
    double t0 = 0, t1 = 0;
    for (int it = 0; it < total; ++it) {
      if (J->reps > 0 && it == 2) { comm_barrier(cp); comm_time(&t0); }
      if (k->mode == 0) gs(u, (gs_dom)k->dom, (gs_op)k->op, k->transpose, g, 0);
      else if (k->mode == 1) gs_vec(u, k->vn, (gs_dom)k->dom, (gs_op)k->op, k->transpose, g, 0);
      else gs_many((void *const *)tab, k->vn, (gs_dom)k->dom, (gs_op)k->op, k->transpose, g, 0);
    }
    if (J->reps > 0) { comm_time(&t1); J->times[c * J->np + r] = (t1 - t0) / J->reps; }
  }
  comm_barrier(cp);
  gs_free(g);
}

/* ids: np pointers; data: ncase*np pointers (case-major), each vn*n[r] elements, updated
   in place; maps/map_len may be NULL; reps > 0 times each case instead of running it once
   (data is then garbage); times: ncase doubles (max over ranks of the mean per call).   */
int ref_job(int np, long long *const *ids, const unsigned *n, int unique, int method,
            int ncase, const struct ref_case *cases, void *const *data,
            unsigned *maps, size_t maps_cap, size_t *map_len, int reps, double *times,
            size_t arena_bytes)
{
  static const unsigned esz[5] = { 8, 4, 4, 8, 8 };
  struct job J;
  memset(&J, 0, sizeof J);
  shim_arena_create(arena_bytes);
  J.np = np; J.unique = unique; J.method = method; J.ncase = ncase; J.reps = reps;
  struct ref_case *cs = shim_arena_raw((size_t)(ncase + 1) * sizeof *cs);
  memcpy(cs, cases, (size_t)ncase * sizeof *cs);
  J.cases = cs;
  J.ids = shim_arena_raw((size_t)np * sizeof *J.ids);
  J.n = shim_arena_raw((size_t)np * sizeof *J.n);
  for (int r = 0; r < np; ++r) {
    J.n[r] = n[r];
    J.ids[r] = shim_arena_raw(((size_t)n[r] + 1) * sizeof(long long));
    memcpy(J.ids[r], ids[r], (size_t)n[r] * sizeof(long long));
  }
  J.data = shim_arena_raw((size_t)(ncase * np + 1) * sizeof *J.data);
  for (int c = 0; c < ncase; ++c)
    for (int r = 0; r < np; ++r) {
      size_t bytes = (size_t)(cases[c].mode == 0 ? 1 : cases[c].vn) * n[r] * esz[cases[c].dom];
      J.data[c * np + r] = shim_arena_raw(bytes + 8);
      memcpy(J.data[c * np + r], data[c * np + r], bytes);
    }
  if (maps) {
    J.maps = shim_arena_raw((size_t)np * maps_cap * sizeof(unsigned));
    J.maps_cap = maps_cap;
    J.map_len = shim_arena_raw((size_t)np * 3 * sizeof(size_t));
  }
  J.times = shim_arena_raw((size_t)(ncase * np + 1) * sizeof(double));
  struct job *Js = shim_arena_raw(sizeof J);
  *Js = J;
  int bad = shim_run(np, rank_main, Js);
  if (!bad) {
    for (int c = 0; c < ncase; ++c)
      for (int r = 0; r < np; ++r) {
        size_t bytes = (size_t)(cases[c].mode == 0 ? 1 : cases[c].vn) * n[r] * esz[cases[c].dom];
        if (reps <= 0) memcpy(data[c * np + r], J.data[c * np + r], bytes);
      }
    if (maps) {
      memcpy(maps, J.maps, (size_t)np * maps_cap * sizeof(unsigned));
      memcpy(map_len, J.map_len, (size_t)np * 3 * sizeof(size_t));
    }
    if (times)
      for (int c = 0; c < ncase; ++c) {
        double t = 0;
        for (int r = 0; r < np; ++r) if (J.times[c * np + r] > t) t = J.times[c * np + r];
        times[c] = t;
      }
  }
  shim_arena_destroy();
  return bad;
}
