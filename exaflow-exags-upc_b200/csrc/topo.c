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

void xg_topo_build(struct xg_topo *t, const i64 *id, size_t n)
{
  memset(t, 0, sizeof *t);
  if (n >= (size_t)XG_NONE)
    xfail("gs_setup: local size %lu does not fit the 32-bit index type", (unsigned long)n);

  /* rows for nonzero labels: key = |id|*2 + flagged, payload = index */
  size_t nnz = 0;
  for (size_t i = 0; i < n; ++i) nnz += (id[i] != 0);
  u64 *key = xnew(u64, nnz), *tk = xnew(u64, nnz);
  u32 *val = xnew(u32, nnz), *tv = xnew(u32, nnz);
  u64 maxkey = 0;
  size_t r = 0;
  for (size_t i = 0; i < n; ++i) {
    i64 v = id[i];
    if (v == 0) continue;
    u64 a = v < 0 ? (u64)(-(v + 1)) + 1u : (u64)v;
    if (a >> 62) xfail("gs_setup: |id| = %llu is too large", (unsigned long long)a);
    u64 k = (a << 1) | (u64)(v < 0);
    key[r] = k; val[r] = (u32)i; ++r;
    if (k > maxkey) maxkey = k;
  }
  /* (|id|, flag) ascending; stable => ascending index inside each class */
  xsort_u64_u32(key, val, nnz, nbits64(maxkey), tk, tv);
  free(tk); free(tv);

  /* label groups in label order */
  size_t ngrp = 0;
  for (size_t j = 0; j < nnz; ++j)
    ngrp += (j == 0 || (key[j] >> 1) != (key[j - 1] >> 1));
  u32 *gs0 = xnew(u32, ngrp + 1);
  u32 *prim = xnew(u32, ngrp ? ngrp : 1);
  u32 maxprim = 0;
  {
    size_t g = 0;
    for (size_t j = 0; j < nnz; ++j)
      if (j == 0 || (key[j] >> 1) != (key[j - 1] >> 1)) {
        gs0[g] = (u32)j; prim[g] = val[j];
        if (val[j] > maxprim) maxprim = val[j];
        ++g;
      }
    gs0[ngrp] = (u32)nnz;
  }
  /* order the groups by primary index (primaries are distinct) */
  u32 *order = xnew(u32, ngrp ? ngrp : 1);
  for (size_t g = 0; g < ngrp; ++g) order[g] = (u32)g;
  xsort_perm_u32(order, prim, ngrp, maxprim);
  free(prim);

  t->nnz = nnz; t->ngrp = ngrp;
  t->id = xnew(u64, ngrp ? ngrp : 1);
  t->idx = xnew(u32, nnz ? nnz : 1);
  t->flag = xnew(unsigned char, nnz ? nnz : 1);
  t->gstart = xnew(u32, ngrp + 1);
  {
    u32 run = 0;
    for (size_t g = 0; g < ngrp; ++g) {
      u32 s = order[g];
      t->gstart[g] = run;
      run += gs0[s + 1] - gs0[s];
    }
    t->gstart[ngrp] = run;
  }
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(xg_setup_threads())
#endif
  for (size_t g = 0; g < ngrp; ++g) {
    u32 s = order[g], a = gs0[s], b = gs0[s + 1], o = t->gstart[g];
    t->id[g] = key[a] >> 1;
    for (u32 j = a; j < b; ++j, ++o) {
      t->idx[o] = val[j];
      t->flag[o] = (unsigned char)(key[j] & 1u);
    }
  }
  free(order); free(gs0); free(key); free(val);
}

/* CSR of the groups with at least two members (src/gs.upc:329-357 emits a
   group only when it has a second participant).  all != 0: every member
   (map_local[1]); all == 0: unflagged members only (map_local[0]).
   nunf is filled only for the `all` map and only if some label is flagged. */
void xg_topo_local_map(const struct xg_topo *t, int all, struct xg_hmap *m)
{
  memset(m, 0, sizeof *m);
  int any_flag = 0;
  for (size_t j = 0; j < t->nnz && !any_flag; ++j) any_flag = t->flag[j];
  size_t ng = 0, ni = 0;
  for (size_t g = 0; g < t->ngrp; ++g) {
    u32 a = t->gstart[g], b = t->gstart[g + 1], cnt = 1;
    if (all) cnt = b - a;
    else for (u32 j = a + 1; j < b && !t->flag[j]; ++j) ++cnt;
    if (cnt >= 2) { ++ng; ni += cnt; }
  }
  if (ni >= (size_t)XG_NONE) xfail("gs_setup: local map too large");
  m->ng = (u32)ng; m->ni = (u32)ni;
  m->off = xnew(u32, ng + 1);
  m->idx = xnew(u32, ni ? ni : 1);
  if (all && any_flag) m->nunf = xnew(u32, ng ? ng : 1);
  size_t og = 0, oi = 0;
  for (size_t g = 0; g < t->ngrp; ++g) {
    u32 a = t->gstart[g], b = t->gstart[g + 1], cnt = 1, unf = 0;
    for (u32 j = a; j < b && !t->flag[j]; ++j) ++unf;
    if (all) cnt = b - a;
    else for (u32 j = a + 1; j < b && !t->flag[j]; ++j) ++cnt;
    if (cnt < 2) continue;
    m->off[og] = (u32)oi;
    if (m->nunf) m->nunf[og] = unf;
    memcpy(m->idx + oi, t->idx + a, (size_t)cnt * sizeof(u32));
    oi += cnt; ++og;
  }
  m->off[ng] = (u32)oi;
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

static inline u32 sell_width(u32 size) { return size <= 2 ? 2u : size <= 4 ? 4u : 8u; }

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
  size_t *wfirst = xnew(size_t, nwin + 1);
  {
    size_t w = 0;
    for (size_t k = 0; k < nshort;) {
      u32 used = 0;
      wfirst[w++] = k;
      while (k < nshort) {
        u32 wd = sell_width(m->off[sg[k] + 1] - m->off[sg[k]]);
        if (used + wd > WSLOTS) break;
        used += wd; ++k;
      }
    }
    wfirst[nwin] = nshort;
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
#ifdef _OPENMP
#pragma omp parallel num_threads(xg_setup_threads())
#endif
  {
    unsigned char fill[XG_SELL_WB * 32];
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
    for (size_t w = 0; w < nwin; ++w) {
      u32 *base = s->idx + w * WSLOTS;
      unsigned char *mk = s->mask + w * XG_SELL_WB * 32;
      for (u32 z = 0; z < WSLOTS; ++z) base[z] = XG_NONE;
      memset(fill, 0, sizeof fill);
      {
        u32 lo = XG_NONE, hi = 0;
        for (size_t k = wfirst[w]; k < wfirst[w + 1]; ++k)
          for (u32 q = m->off[sg[k]]; q < m->off[sg[k] + 1]; ++q) {
            if (m->idx[q] < lo) lo = m->idx[q];
            if (m->idx[q] > hi) hi = m->idx[q];
          }
        s->win_min[w] = lo; s->win_max[w] = hi;
      }
      /* descending width, stable in primary order; rows of equal width are
         dealt lane after lane, block after block */
      for (u32 wd = 8; wd >= 2; wd >>= 1) {
        u32 pos = 0;                                /* lane cursor over the window */
        for (size_t k = wfirst[w]; k < wfirst[w + 1]; ++k) {
          u32 g = sg[k], a = m->off[g], sz = m->off[g + 1] - a;
          if (sell_width(sz) != wd) continue;
          /* next lane (block-major, then lane) with room; fills are multiples
             of wd here, and lanes before the cursor are at least as full */
          u32 tries = 0;
          while (fill[pos] + wd > 8) {
            pos = (pos + 1) % (XG_SELL_WB * 32);
            if (++tries > XG_SELL_WB * 32) xfail("gs_setup: internal error packing the sliced layout");
          }
          u32 *slot = base + (size_t)pos * 8 + fill[pos];
          for (u32 q = 0; q < sz; ++q) slot[q] = m->idx[a + q];
          mk[pos] |= (unsigned char)(1u << fill[pos]);
          if (s->fmask) {                            /* members [nunf, sz) are flagged */
            u32 unf = m->nunf[g];
            s->fmask[w * XG_SELL_WB * 32 + pos] |= (unsigned char)((((1u << sz) - 1u) & ~((1u << unf) - 1u)) << fill[pos]);
          }
          fill[pos] = (unsigned char)(fill[pos] + wd);
          /* row-major dealing: move on to the next lane; wrap to start a new
             row of this width */
          pos = (pos + 1) % (XG_SELL_WB * 32);
        }
      }
    }
  }
  free(wfirst); free(sg);
}
