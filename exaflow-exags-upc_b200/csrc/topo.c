/* topo.c -- local label topology and the local group maps.
 *
 * Restates what nonzero_ids() (src/gs.upc:87-112), local_map()
 * (src/gs.upc:329-357) and flagged_primaries_map() (src/gs.upc:359-370)
 * establish, with a different data layout:
 *
 *   reference: one 24-byte {id,i,flag,primary} row per nonzero label, two
 *              stable struct sorts, then three sentinel-terminated uint
 *              streams walked serially by the kernels;
 *   here:      split key/payload arrays, one radix sort on (|id|,flag) and a
 *              group-level sort on the primary index, then CSR arrays
 *              (offsets + members) that go to HBM as they are.
 *
 * Ordering contract kept bit-for-bit, because it fixes the floating-point
 * summation order: groups ascend by primary index; inside a group unflagged
 * members come first, each class by ascending local index; the primary is the
 * first member.
 */
#include <stdlib.h>
#include <string.h>
#include "exags_internal.h"

void xg_topo_free(struct xg_topo *t)
{
  free(t->id); free(t->idx); free(t->flag); free(t->gstart);
  memset(t, 0, sizeof *t);
}

void xg_hmap_free(struct xg_hmap *m)
{
  free(m->off); free(m->idx); free(m->nunf);
  memset(m, 0, sizeof *m);
}

static unsigned nbits64(u64 v)
{
  unsigned b = 0;
  while (v) { ++b; v >>= 1; }
  return b ? b : 1;
}

/* The setup loops below run over 10^8 rows; each is cut into XG_CHUNKS
   contiguous chunks handled in two passes (count per chunk, exclusive prefix,
   fill at the chunk's offset), so the output order is the serial order and the
   chunks can be spread over the host cores.                                 */
#define XG_CHUNKS 256
#define CHUNK_LO(n, c) ((size_t)(n) * (size_t)(c) / XG_CHUNKS)
#ifdef _OPENMP
#define PAR_CHUNKS _Pragma("omp parallel for schedule(dynamic, 1) num_threads(xg_setup_threads())")
#else
#define PAR_CHUNKS
#endif

static size_t prefix_chunks(size_t *cnt)       /* exclusive prefix in place; returns the total */
{
  size_t run = 0;
  for (int c = 0; c < XG_CHUNKS; ++c) { size_t v = cnt[c]; cnt[c] = run; run += v; }
  return run;
}

void xg_topo_build(struct xg_topo *t, const i64 *id, size_t n)
{
  memset(t, 0, sizeof *t);
  if (n >= (size_t)XG_NONE)
    xfail("gs_setup: local size %lu does not fit the 32-bit index type", (unsigned long)n);
  size_t cnt[XG_CHUNKS], cnt2[XG_CHUNKS];
  u64 cmax[XG_CHUNKS];

  /* rows for nonzero labels: key = |id|*2 + flagged, payload = index */
  PAR_CHUNKS
  for (int c = 0; c < XG_CHUNKS; ++c) {
    size_t k = 0;
    for (size_t i = CHUNK_LO(n, c); i < CHUNK_LO(n, c + 1); ++i) k += (id[i] != 0);
    cnt[c] = k;
  }
  const size_t nnz = prefix_chunks(cnt);
  u64 *key = xnew(u64, nnz ? nnz : 1), *tk = xnew(u64, nnz ? nnz : 1);
  u32 *val = xnew(u32, nnz ? nnz : 1), *tv = xnew(u32, nnz ? nnz : 1);
  PAR_CHUNKS
  for (int c = 0; c < XG_CHUNKS; ++c) {
    size_t r = cnt[c];
    u64 mx = 0;
    for (size_t i = CHUNK_LO(n, c); i < CHUNK_LO(n, c + 1); ++i) {
      i64 v = id[i];
      if (v == 0) continue;
      u64 a = v < 0 ? (u64)(-(v + 1)) + 1u : (u64)v;
      u64 k = (a << 1) | (u64)(v < 0);
      key[r] = k; val[r] = (u32)i; ++r;
      if (a > mx) mx = a;
    }
    cmax[c] = mx;
  }
  u64 maxabs = 0;
  for (int c = 0; c < XG_CHUNKS; ++c) if (cmax[c] > maxabs) maxabs = cmax[c];
  if (maxabs >> 62) xfail("gs_setup: |id| = %llu is too large", (unsigned long long)maxabs);
  /* (|id|, flag) ascending; stable => ascending index inside each class */
  xsort_u64_u32(key, val, nnz, nbits64((maxabs << 1) | 1u), tk, tv);
  free(tk); free(tv);

  /* label groups in label order: row j heads a group when its label differs
     from the row before */
#define IS_HEAD(j) ((j) == 0 || (key[j] >> 1) != (key[(j) - 1] >> 1))
  PAR_CHUNKS
  for (int c = 0; c < XG_CHUNKS; ++c) {
    size_t k = 0;
    for (size_t j = CHUNK_LO(nnz, c); j < CHUNK_LO(nnz, c + 1); ++j) k += IS_HEAD(j);
    cnt[c] = k;
  }
  const size_t ngrp = prefix_chunks(cnt);
  u32 *gs0 = xnew(u32, ngrp + 1);
  /* order the groups by primary index.  Primaries are distinct local indices,
     so one table over the indices replaces a sort: slot[primary] = group, then
     the occupied slots are read off in ascending order                       */
  u32 *slot = xnew(u32, n ? n : 1);
  PAR_CHUNKS
  for (int c = 0; c < XG_CHUNKS; ++c)
    for (size_t i = CHUNK_LO(n, c); i < CHUNK_LO(n, c + 1); ++i) slot[i] = XG_NONE;
  PAR_CHUNKS
  for (int c = 0; c < XG_CHUNKS; ++c) {
    size_t g = cnt[c];
    for (size_t j = CHUNK_LO(nnz, c); j < CHUNK_LO(nnz, c + 1); ++j)
      if (IS_HEAD(j)) { gs0[g] = (u32)j; slot[val[j]] = (u32)g; ++g; }
  }
#undef IS_HEAD
  gs0[ngrp] = (u32)nnz;
  u32 *order = xnew(u32, ngrp ? ngrp : 1);
  PAR_CHUNKS
  for (int c = 0; c < XG_CHUNKS; ++c) {
    size_t k = 0;
    for (size_t i = CHUNK_LO(n, c); i < CHUNK_LO(n, c + 1); ++i) k += (slot[i] != XG_NONE);
    cnt2[c] = k;
  }
  prefix_chunks(cnt2);
  PAR_CHUNKS
  for (int c = 0; c < XG_CHUNKS; ++c) {
    size_t o = cnt2[c];
    for (size_t i = CHUNK_LO(n, c); i < CHUNK_LO(n, c + 1); ++i)
      if (slot[i] != XG_NONE) order[o++] = slot[i];
  }
  free(slot);

  t->nnz = nnz; t->ngrp = ngrp;
  t->id = xnew(u64, ngrp ? ngrp : 1);
  t->idx = xnew(u32, nnz ? nnz : 1);
  t->flag = xnew(unsigned char, nnz ? nnz : 1);
  t->gstart = xnew(u32, ngrp + 1);
  /* row offset of every group in the new order */
  PAR_CHUNKS
  for (int c = 0; c < XG_CHUNKS; ++c) {
    size_t k = 0;
    for (size_t g = CHUNK_LO(ngrp, c); g < CHUNK_LO(ngrp, c + 1); ++g) k += gs0[order[g] + 1] - gs0[order[g]];
    cnt[c] = k;
  }
  prefix_chunks(cnt);
  PAR_CHUNKS
  for (int c = 0; c < XG_CHUNKS; ++c) {
    u32 o = (u32)cnt[c];
    for (size_t g = CHUNK_LO(ngrp, c); g < CHUNK_LO(ngrp, c + 1); ++g) {
      const u32 s = order[g], a = gs0[s], b = gs0[s + 1];
      t->gstart[g] = o;
      t->id[g] = key[a] >> 1;
      for (u32 j = a; j < b; ++j, ++o) {
        t->idx[o] = val[j];
        t->flag[o] = (unsigned char)(key[j] & 1u);
      }
    }
  }
  t->gstart[ngrp] = (u32)nnz;
  free(order); free(gs0); free(key); free(val);
}

/* CSR of the groups with at least two members (src/gs.upc:329-357 emits a
   group only when it has a second participant).  all != 0: every member
   (map_local[1]); all == 0: unflagged members only (map_local[0]).
   nunf is filled only for the `all` map and only if some label is flagged. */
void xg_topo_local_map(const struct xg_topo *t, int all, struct xg_hmap *m)
{
  memset(m, 0, sizeof *m);
  size_t cg[XG_CHUNKS], ci[XG_CHUNKS], cf[XG_CHUNKS];
  PAR_CHUNKS
  for (int c = 0; c < XG_CHUNKS; ++c) {
    size_t f = 0;
    for (size_t j = CHUNK_LO(t->nnz, c); j < CHUNK_LO(t->nnz, c + 1) && !f; ++j) f = t->flag[j];
    cf[c] = f;
  }
  int any_flag = 0;
  for (int c = 0; c < XG_CHUNKS; ++c) any_flag |= cf[c] != 0;
  PAR_CHUNKS
  for (int c = 0; c < XG_CHUNKS; ++c) {
    size_t ng = 0, ni = 0;
    for (size_t g = CHUNK_LO(t->ngrp, c); g < CHUNK_LO(t->ngrp, c + 1); ++g) {
      u32 a = t->gstart[g], b = t->gstart[g + 1], k = 1;
      if (all) k = b - a;
      else for (u32 j = a + 1; j < b && !t->flag[j]; ++j) ++k;
      if (k >= 2) { ++ng; ni += k; }
    }
    cg[c] = ng; ci[c] = ni;
  }
  const size_t ng = prefix_chunks(cg), ni = prefix_chunks(ci);
  if (ni >= (size_t)XG_NONE) xfail("gs_setup: local map too large");
  m->ng = (u32)ng; m->ni = (u32)ni;
  m->off = xnew(u32, ng + 1);
  m->idx = xnew(u32, ni ? ni : 1);
  if (all && any_flag) m->nunf = xnew(u32, ng ? ng : 1);
  PAR_CHUNKS
  for (int c = 0; c < XG_CHUNKS; ++c) {
    size_t og = cg[c], oi = ci[c];
    for (size_t g = CHUNK_LO(t->ngrp, c); g < CHUNK_LO(t->ngrp, c + 1); ++g) {
      u32 a = t->gstart[g], b = t->gstart[g + 1], k = 1, unf = 0;
      if (any_flag) for (u32 j = a; j < b && !t->flag[j]; ++j) ++unf;
      if (all) k = b - a;
      else for (u32 j = a + 1; j < b && !t->flag[j]; ++j) ++k;
      if (k < 2) continue;
      m->off[og] = (u32)oi;
      if (m->nunf) m->nunf[og] = unf;
      memcpy(m->idx + oi, t->idx + a, (size_t)k * sizeof(u32));
      oi += k; ++og;
    }
  }
  m->off[ng] = (u32)ni;
}

/* indices of primaries whose label is flagged (src/gs.upc:359-370); when
   singles_only != 0 keep only those that head no multi-member group.       */
u32 *xg_topo_flagged_primaries(const struct xg_topo *t, int singles_only, u32 *count)
{
  size_t c = 0;
  for (size_t g = 0; g < t->ngrp; ++g) {
    u32 a = t->gstart[g], b = t->gstart[g + 1];
    if (t->flag[a] && (!singles_only || b - a == 1)) ++c;
  }
  u32 *out = xnew(u32, c ? c : 1);
  size_t o = 0;
  for (size_t g = 0; g < t->ngrp; ++g) {
    u32 a = t->gstart[g], b = t->gstart[g + 1];
    if (t->flag[a] && (!singles_only || b - a == 1)) out[o++] = t->idx[a];
  }
  *count = (u32)c;
  return out;
}

/* ------------------------------------------------------------------ */
void xg_sell_free(struct xg_sell *s)
{
  free(s->mask); free(s->fmask); free(s->idx); free(s->win_min); free(s->win_max);
  xg_hmap_free(&s->rest);
  memset(s, 0, sizeof *s);
}

#include "sell_window.h"
#define sell_width xg_sell_width

/* EXAGS_SELL_PAIRS: 0 deals every class in plain primary order, 1 chains the
   pairs only, 2 (default) every width class (sell_window.h)                */
int xg_sell_pairs_enabled(void)
{
  const char *e = getenv("EXAGS_SELL_PAIRS");
  const int v = e ? atoi(e) : 2;
  /* EXAGS_SELL_LINE_SHIFT (4..6): line size the ordering works with, in bits
     8.. of the value (sell_window.h).  Default 5 = 32 elements: the line of a
     4-byte field and two lines of an 8-byte one.  Measured on M7
     (profiles/r2_knobs_line_shift_compress.txt): float 0.1959 -> 0.1922 ms,
     gs_vec(3,double) 0.9046 -> 0.8957 ms, double unchanged (0.3182 -> 0.3177). */
  const char *l = getenv("EXAGS_SELL_LINE_SHIFT");
  const int ls = l ? atoi(l) : 5;
  return (v < 0 ? 0 : v > 2 ? 2 : v) | ((ls >= 4 && ls <= 6 && ls != 4) ? ls << 8 : 0);
}

void xg_sell_build(struct xg_sell *s, const struct xg_hmap *m)
{
  const u32 WSLOTS = XG_SELL_WB * 256u;            /* member slots per window */
  memset(s, 0, sizeof *s);
  /* long groups go to the residual CSR */
  size_t nlong = 0, ilong = 0, nshort = 0;
  for (u32 g = 0; g < m->ng; ++g) {
    u32 sz = m->off[g + 1] - m->off[g];
    if (sz > XG_SELL_WMAX) { ++nlong; ilong += sz; } else ++nshort;
  }
  s->rest.ng = (u32)nlong; s->rest.ni = (u32)ilong;
  s->rest.off = xnew(u32, nlong + 1);
  s->rest.idx = xnew(u32, ilong ? ilong : 1);
  if (m->nunf) s->rest.nunf = xnew(u32, nlong ? nlong : 1);
  u32 *sg = xnew(u32, nshort ? nshort : 1);        /* short groups, primary order */
  {
    size_t og = 0, oi = 0, os = 0;
    for (u32 g = 0; g < m->ng; ++g) {
      u32 a = m->off[g], sz = m->off[g + 1] - a;
      if (sz > XG_SELL_WMAX) {
        if (s->rest.nunf) s->rest.nunf[og] = m->nunf[g];
        s->rest.off[og++] = (u32)oi;
        memcpy(s->rest.idx + oi, m->idx + a, (size_t)sz * sizeof(u32));
        oi += sz;
      } else sg[os++] = g;
    }
    s->rest.off[nlong] = (u32)oi;
  }
  /* windows: greedy runs of consecutive groups whose padded widths fill
     WSLOTS; power-of-two widths dealt in descending order always pack, so a
     window is full exactly when its widths add up to WSLOTS */
  size_t nwin = 0;
  for (size_t k = 0; k < nshort;) {
    u32 used = 0;
    while (k < nshort) {
      u32 w = sell_width(m->off[sg[k] + 1] - m->off[sg[k]]);
      if (used + w > WSLOTS) break;
      used += w; ++k;
    }
    ++nwin;
  }
  u32 *wfirst = xnew(u32, nwin + 1);
  {
    size_t w = 0;
    for (size_t k = 0; k < nshort;) {
      u32 used = 0;
      wfirst[w++] = (u32)k;
      while (k < nshort) {
        u32 wd = sell_width(m->off[sg[k] + 1] - m->off[sg[k]]);
        if (used + wd > WSLOTS) break;
        used += wd; ++k;
      }
    }
    wfirst[nwin] = (u32)nshort;
  }
  if ((nwin * XG_SELL_WB) >> 23) xfail("gs_setup: map too large for the sliced layout");
  s->nblock = (u32)(nwin * XG_SELL_WB);
  s->nidx = (size_t)s->nblock * 256;
  s->nwin = (u32)nwin;
  s->win_min = xnew(u32, nwin ? nwin : 1);
  s->win_max = xnew(u32, nwin ? nwin : 1);
  s->mask = xnew0(unsigned char, (size_t)s->nblock * 32 + 1);
  if (m->nunf) s->fmask = xnew0(unsigned char, (size_t)s->nblock * 32 + 1);
  s->idx = xnew(u32, s->nidx ? s->nidx : 1);
  int bad = 0;
  const int chain = xg_sell_pairs_enabled();
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64) num_threads(xg_setup_threads()) reduction(| : bad)
#endif
  for (size_t w = 0; w < nwin; ++w)
    bad |= xg_sell_fill_window(w, wfirst[w], wfirst[w + 1], sg, m->off, m->idx, m->nunf,
                               s->idx, s->mask, s->fmask, s->win_min, s->win_max, chain);
  if (bad) xfail("gs_setup: internal error packing the sliced layout");
  free(wfirst); free(sg);
}
