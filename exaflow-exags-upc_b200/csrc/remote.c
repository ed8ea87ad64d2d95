/* remote.c -- discovery of labels shared between ranks and the pairwise /
 * all-reduce exchange plans.
 *
 * Restates, on the host in C as the north star prescribes:
 *   unique_ids + shared_ids + shared_ids_aux   src/gs.upc:117-244
 *   make_topology_unique (owner choice)         src/gs.upc:253-322
 *   pw_comm_setup + pw_map_setup                src/gs.upc:505-570
 *   allreduce_map_setup                         src/gs.upc:1031-1050
 *
 * The reference routes rows with a log-P crystal router over UPC puts
 * (src/sarray_transfer.upc:137-180); on one node every rank reaches every
 * other directly, so the same rendezvous ("work rank" = id mod np) is done
 * with one all-to-all over the host shared-memory segment each way.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "exags_internal.h"
#include "gs_plan.h"

struct uq_row { u64 id; u32 src_if; u32 pad; };                 /* to work rank   */
struct wk_row { u64 id, ord; u32 p2, i1f, i2f, pad; };          /* back to source */

static inline u32 unflag(u32 v) { return (~v < v) ? ~v : v; }   /* src/gs.upc:176 */
static inline int is_flagged(u32 v) { return ~v < v; }

static unsigned nbits64(u64 v) { unsigned b = 0; while (v) { ++b; v >>= 1; } return b ? b : 1; }

/* ------------------------------------------------------------------ */
void xg_shared_free(struct xg_shared *s)
{
  free(s->row);
  memset(s, 0, sizeof *s);
}

void xg_cand_free(struct xg_cand *c) { free(c->id); free(c->src_if); memset(c, 0, sizeof *c); }

/* Candidate labels of this rank, chosen on the host from the host copy of the
   topology (EXAGS_SETUP=host, gs_unique, setup with unique=1).  The device
   passes of setup_cuda.cu (xcu_setup_candidates) apply the same two filters
   in HBM and bring back only the candidates.                               */
static void shared_candidates_host(struct xg_cand *out, const struct xg_topo *t, struct xg_world *w)
{
  const int np = xg_world_np(w), me = xg_world_rank(w);
  memset(out, 0, sizeof *out);

  /* Which of my labels can another rank hold at all?  The reference ships every
     local label to its work rank (src/gs.upc:118-131); on a spectral-element
     mesh well under 1 % of them are shared.  Every rank publishes a presence
     bitmap over a hash of its labels; a label whose bit is clear in all the
     other ranks' bitmaps is on no other rank and cannot produce a shared row,
     so it stays home.  The filter only errs towards sending (hash collisions),
     so the rows the work ranks pair up -- and everything derived from them --
     are exactly the reference's.                                            */
  unsigned char *cand = xnew(unsigned char, t->ngrp ? t->ngrp : 1);
  const int nthr = xg_setup_threads();
  (void)nthr;
  {
    /* Stage 1, coarse: the label space [0, global max] is cut into XG_BINS
       equal bins and every rank publishes which bins it has labels in (8 KB).
       A label in a bin no other rank occupies cannot be shared.  With the
       lexicographic numbering of a partitioned mesh this already leaves little
       more than the interface labels, and the table stays in L1 -- the hashed
       bitmap below, whose random accesses miss the caches, then sees only
       those.  A shared label passes on both of its ranks (each sees the other's
       bin), so the filter never loses one.                                   */
    enum { XG_BINS = 65536 };
    u64 lmax = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(max : lmax) num_threads(nthr)
#endif
    for (size_t g = 0; g < t->ngrp; ++g) if (t->id[g] > lmax) lmax = t->id[g];
    u64 *all_max = xnew(u64, np), gmax = 0;
    xg_host_allgather(w, &lmax, sizeof lmax, all_max);
    for (int p = 0; p < np; ++p) if (all_max[p] > gmax) gmax = all_max[p];
    free(all_max);
    unsigned shift = 0;
    while ((gmax >> shift) >= (u64)XG_BINS) ++shift;
    u64 *bins = xnew0(u64, XG_BINS / 64), *all_bins = xnew(u64, (size_t)np * (XG_BINS / 64));
    /* few distinct words, hit over and over: plain loop per thread, merged */
#ifdef _OPENMP
#pragma omp parallel num_threads(nthr)
#endif
    {
      u64 *loc = xnew0(u64, XG_BINS / 64);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
      for (size_t g = 0; g < t->ngrp; ++g) { const u64 b = t->id[g] >> shift; loc[b >> 6] |= (u64)1 << (b & 63); }
#ifdef _OPENMP
#pragma omp critical
#endif
      for (size_t k = 0; k < XG_BINS / 64; ++k) bins[k] |= loc[k];
      free(loc);
    }
    xg_host_allgather(w, bins, (XG_BINS / 64) * sizeof(u64), all_bins);
    memset(bins, 0, (XG_BINS / 64) * sizeof(u64));
    for (int p = 0; p < np; ++p)
      if (p != me) for (size_t k = 0; k < XG_BINS / 64; ++k) bins[k] |= all_bins[(size_t)p * (XG_BINS / 64) + k];
    free(all_bins);
    u64 npass = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+ : npass) num_threads(nthr)
#endif
    for (size_t g = 0; g < t->ngrp; ++g) {
      const u64 b = t->id[g] >> shift;
      cand[g] = (unsigned char)((bins[b >> 6] >> (b & 63)) & 1u);
      npass += cand[g];
    }
    free(bins);

    /* Stage 2, fine: presence bitmap over a hash of the labels that passed */
    u64 mine_n = npass, *all_n = xnew(u64, np), maxn = 1;
    xg_host_allgather(w, &mine_n, sizeof mine_n, all_n);
    for (int p = 0; p < np; ++p) if (all_n[p] > maxn) maxn = all_n[p];
    free(all_n);
    unsigned bits = 20;
    while (bits < 34 && ((u64)1 << bits) < 16 * maxn) ++bits;       /* <= 1/16 full: ~6 % false positives */
    const size_t words = (size_t)1 << (bits - 6);
    u64 *mine = xg_host_bitmap_new(w, words), *others = xnew(u64, words);
#define LABEL_BIT(id_) (((id_) * 0x9E3779B97F4A7C15ull) >> (64 - bits))
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthr)
#endif
    for (size_t g = 0; g < t->ngrp; ++g) {
      if (!cand[g]) continue;
      const u64 h = LABEL_BIT(t->id[g]);
      __atomic_fetch_or(&mine[h >> 6], (u64)1 << (h & 63), __ATOMIC_RELAXED);
    }
    xg_host_bitmap_or_others(w, mine, words, others);   /* releases mine */
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthr)
#endif
    for (size_t g = 0; g < t->ngrp; ++g) {
      if (!cand[g]) continue;
      const u64 h = LABEL_BIT(t->id[g]);
      cand[g] = (unsigned char)((others[h >> 6] >> (h & 63)) & 1u);
    }
#undef LABEL_BIT
    free(others);
  }
  size_t nc = 0;
  for (size_t g = 0; g < t->ngrp; ++g) nc += cand[g];
  out->n = nc;
  out->id = xnew(u64, nc ? nc : 1); out->src_if = xnew(u32, nc ? nc : 1);
  nc = 0;
  for (size_t g = 0; g < t->ngrp; ++g) {
    if (!cand[g]) continue;
    const u32 a = t->gstart[g];
    out->id[nc] = t->id[g];
    out->src_if[nc] = t->flag[a] ? ~t->idx[a] : t->idx[a];
    ++nc;
  }
  free(cand);
}

void xg_shared_discover(struct xg_shared *out, const struct xg_topo *t, struct xg_world *w)
{
  memset(out, 0, sizeof *out);
  if (xg_world_np(w) == 1) return;
  struct xg_cand c;
  shared_candidates_host(&c, t, w);
  xg_shared_pair(out, &c, w);
  xg_cand_free(&c);
}

/* The rendezvous of shared_ids (src/gs.upc:190-244) on the candidate labels:
   one row per candidate goes to its work rank id % np, which pairs the ranks
   holding the same label and sends the pair rows back.                      */
void xg_shared_pair(struct xg_shared *out, const struct xg_cand *c, struct xg_world *w)
{
  const int np = xg_world_np(w), me = xg_world_rank(w);
  memset(out, 0, sizeof *out);
  if (np == 1) return;
  const char *prof_env = getenv("EXAGS_SETUP_PROFILE");
  const int prof = prof_env && atoi(prof_env) >= 2 && me == 0;
  double tp[8]; int ntp = 0;
#define TICK() do { if (prof) exags_comm_time(&tp[ntp++]); } while (0)
  TICK();
  /* one row per candidate label, grouped by work rank id % np (src/gs.upc:118-131) */
  size_t *cnt = xnew0(size_t, np), *pos = xnew(size_t, np);
  const size_t nsend = c->n;
  for (size_t k = 0; k < nsend; ++k) ++cnt[c->id[k] % (u64)np];
  { size_t run = 0; for (int p = 0; p < np; ++p) { pos[p] = run; run += cnt[p]; } }
  struct uq_row *send = xnew(struct uq_row, nsend ? nsend : 1);
  for (size_t k = 0; k < nsend; ++k) {
    struct uq_row *r = send + pos[c->id[k] % (u64)np]++;
    r->id = c->id[k];
    r->src_if = c->src_if[k];
    r->pad = 0;
  }
  TICK();
  size_t *rcnt = xnew(size_t, np), nrecv = 0;
  struct uq_row *un = (struct uq_row *)xg_host_alltoallv(w, send, sizeof *send, cnt, rcnt, &nrecv);
  free(send);
  TICK();

  /* work rank: group by id with unflagged rows ahead of flagged ones
     (src/gs.upc:204).  Rows arrive ordered by source rank. */
  u32 *src = xnew(u32, nrecv ? nrecv : 1);
  { size_t o = 0; for (int p = 0; p < np; ++p) for (size_t k = 0; k < rcnt[p]; ++k) src[o++] = (u32)p; }
  u32 *perm = xnew(u32, nrecv ? nrecv : 1);
  {
    size_t o = 0;
    for (size_t k = 0; k < nrecv; ++k) if (!is_flagged(un[k].src_if)) perm[o++] = (u32)k;
    for (size_t k = 0; k < nrecv; ++k) if (is_flagged(un[k].src_if)) perm[o++] = (u32)k;
  }
  u64 *ids = xnew(u64, nrecv ? nrecv : 1), maxid = 0;
  for (size_t k = 0; k < nrecv; ++k) { ids[k] = un[k].id; if (ids[k] > maxid) maxid = ids[k]; }
  xsort_perm_u64(perm, ids, nrecv, maxid);
  free(ids);

  /* count shared labels (src/gs.upc:206-212): an unflagged first row with at
     least one more row behind it */
  u64 n_shared = 0;
  size_t nwork = 0;
  for (size_t a = 0; a < nrecv;) {
    size_t b = a + 1;
    while (b < nrecv && un[perm[b]].id == un[perm[a]].id) ++b;
    if (b - a >= 2 && !is_flagged(un[perm[a]].src_if)) {
      ++n_shared;
      size_t nu = 0;
      for (size_t k = a; k < b; ++k) nu += !is_flagged(un[perm[k]].src_if);
      /* pairs with at least one unflagged side, both directions */
      size_t L = b - a;
      nwork += 2 * (nu * (nu - 1) / 2 + nu * (L - nu));
    }
    a = b;
  }
  /* global ordinal of each shared label (comm_scan, src/gs.upc:213) */
  u64 *all_ns = xnew(u64, np);
  xg_host_allgather(w, &n_shared, sizeof n_shared, all_ns);
  u64 ord0 = 0, total = 0;
  for (int p = 0; p < np; ++p) { if (p < me) ord0 += all_ns[p]; total += all_ns[p]; }
  free(all_ns);

  /* emit the pair rows (src/gs.upc:218-236), then group them by destination */
  struct wk_row *wk = xnew(struct wk_row, nwork ? nwork : 1);
  u32 *dst = xnew(u32, nwork ? nwork : 1);
  size_t nw = 0;
  u64 ord = ord0;
  for (size_t a = 0; a < nrecv;) {
    size_t b = a + 1;
    while (b < nrecv && un[perm[b]].id == un[perm[a]].id) ++b;
    if (b - a >= 2 && !is_flagged(un[perm[a]].src_if)) {
      for (size_t x = a; x < b; ++x) {
        const struct uq_row *rx = un + perm[x];
        if (is_flagged(rx->src_if)) break;
        for (size_t y = x + 1; y < b; ++y) {
          const struct uq_row *ry = un + perm[y];
          u32 p1 = src[perm[x]], p2 = src[perm[y]];
          wk[nw].id = rx->id; wk[nw].ord = ord; wk[nw].p2 = p2; wk[nw].i1f = rx->src_if; wk[nw].i2f = ry->src_if; wk[nw].pad = 0; dst[nw] = p1; ++nw;
          wk[nw].id = rx->id; wk[nw].ord = ord; wk[nw].p2 = p1; wk[nw].i1f = ry->src_if; wk[nw].i2f = rx->src_if; wk[nw].pad = 0; dst[nw] = p2; ++nw;
        }
      }
      ++ord;
    }
    a = b;
  }
  free(perm); free(src); free(un);

  memset(cnt, 0, (size_t)np * sizeof *cnt);
  for (size_t k = 0; k < nw; ++k) ++cnt[dst[k]];
  { size_t run = 0; for (int p = 0; p < np; ++p) { pos[p] = run; run += cnt[p]; } }
  struct wk_row *wsend = xnew(struct wk_row, nw ? nw : 1);
  for (size_t k = 0; k < nw; ++k) wsend[pos[dst[k]]++] = wk[k];
  free(wk); free(dst);
  size_t nback = 0;
  struct wk_row *back = (struct wk_row *)xg_host_alltoallv(w, wsend, sizeof *wsend, cnt, rcnt, &nback);
  free(wsend); free(cnt); free(pos); free(rcnt);

  /* translate to shared rows (src/gs.upc:163-188) */
  out->n = nback;
  out->total_shared = total;
  out->row = xnew(struct xg_shared_row, nback ? nback : 1);
  for (size_t k = 0; k < nback; ++k) {
    struct xg_shared_row *s = out->row + k;
    s->id = back[k].id; s->ord = back[k].ord;
    s->i = unflag(back[k].i1f); s->p = back[k].p2; s->ri = unflag(back[k].i2f);
    s->flags = (is_flagged(back[k].i1f) ? XG_FLAGS_LOCAL : 0u) | (is_flagged(back[k].i2f) ? XG_FLAGS_REMOTE : 0u);
  }
  free(back);
  TICK();
  if (prof)
    printf("shared discovery (rank 0): %lu candidate rows %.3f s, to work ranks %.3f s, "
           "pairing + return %.3f s\n", (unsigned long)nsend, tp[1] - tp[0], tp[2] - tp[1], tp[3] - tp[2]);
#undef TICK
}

/* sort shared rows by (p, id): the globally consistent message order
   (src/gs.upc:510-511) */
static void sort_rows_p_id(struct xg_shared *s)
{
  size_t n = s->n;
  if (n < 2) return;
  u32 *perm = xnew(u32, n);
  u64 *key = xnew(u64, n), maxk = 0;
  for (size_t k = 0; k < n; ++k) { perm[k] = (u32)k; key[k] = s->row[k].id; if (key[k] > maxk) maxk = key[k]; }
  xsort_perm_u64(perm, key, n, maxk);
  maxk = 0;
  for (size_t k = 0; k < n; ++k) { key[k] = s->row[k].p; if (key[k] > maxk) maxk = key[k]; }
  xsort_perm_u64(perm, key, n, maxk);
  struct xg_shared_row *r2 = xnew(struct xg_shared_row, n);
  for (size_t k = 0; k < n; ++k) r2[k] = s->row[perm[k]];
  free(s->row); s->row = r2;
  free(perm); free(key);
}

/* ------------------------------------------------------------------ */
/* gs_unique: which local labels end up flagged (src/gs.upc:253-322).
   newflag[i] = 1 marks local index i as flagged afterwards.             */
void xg_unique_flags(unsigned char *newflag, size_t n, const struct xg_topo *t,
                     struct xg_shared *sh, int me)
{
  memset(newflag, 0, n);
  /* keep existing flags; flag every local non-primary (src/gs.upc:262-271) */
  for (size_t g = 0; g < t->ngrp; ++g) {
    u32 a = t->gstart[g], b = t->gstart[g + 1];
    newflag[t->idx[a]] = t->flag[a];
    for (u32 j = a + 1; j < b; ++j) newflag[t->idx[j]] = 1;
  }
  /* owner among the ranks sharing a label: the (id mod count)-th unflagged
     sharer in rank order (src/gs.upc:279-302) */
  sort_rows_p_id(sh);               /* by p, then group by i below */
  size_t ns = sh->n;
  if (!ns) return;
  u32 *perm = xnew(u32, ns), *key = xnew(u32, ns), maxi = 0;
  for (size_t k = 0; k < ns; ++k) { perm[k] = (u32)k; key[k] = sh->row[k].i; if (key[k] > maxi) maxi = key[k]; }
  xsort_perm_u32(perm, key, ns, maxi);   /* stable: same i stays ordered by p */
  free(key);
  for (size_t a = 0; a < ns;) {
    size_t b = a;
    const struct xg_shared_row *first = sh->row + perm[a];
    u32 lt = 0, gt = 0;
    while (b < ns && sh->row[perm[b]].i == first->i) {
      const struct xg_shared_row *r = sh->row + perm[b];
      if (!(r->flags & XG_FLAGS_REMOTE)) { if (r->p < (u32)me) ++lt; else ++gt; }
      ++b;
    }
    int mine = 0;
    if (!(first->flags & XG_FLAGS_LOCAL)) mine = (first->id % (u64)(lt + gt + 1)) == lt;
    if (!mine) newflag[first->i] = 1;
    a = b;
  }
  free(perm);
}

/* ------------------------------------------------------------------ */
void xg_plan_free(struct xg_plan *pl)
{
  free(pl->bprim); free(pl->bgrp);
  for (int d = 0; d < 2; ++d) {
    free(pl->soff[d]); free(pl->slot[d]);
    free(pl->nb_rank[d]); free(pl->nb_start[d]); free(pl->nb_size[d]); free(pl->nb_peer_start[d]);
    free(pl->snb[d]);
  }
  free(pl->sym_rank);
  free(pl->ord); free(pl->pflag);
  memset(pl, 0, sizeof *pl);
}

/* Exchange plan for this rank.  Boundary primaries are the distinct local
   indices that appear in a shared row.  For each direction d the slots are
   numbered over the rows sorted by (p,id) that pass the mask
   (d==0: remote unflagged, d==1: local unflagged), src/gs.upc:505-535.     */
void xg_plan_build(struct xg_plan *pl, struct xg_shared *sh, const struct xg_topo *t)
{
  memset(pl, 0, sizeof *pl);
  pl->total_shared = sh->total_shared;
  size_t ns = sh->n;
  sort_rows_p_id(sh);

  /* boundary primaries, ascending */
  u32 *perm = xnew(u32, ns ? ns : 1), *key = xnew(u32, ns ? ns : 1), maxi = 0;
  for (size_t k = 0; k < ns; ++k) { perm[k] = (u32)k; key[k] = sh->row[k].i; if (key[k] > maxi) maxi = key[k]; }
  xsort_perm_u32(perm, key, ns, maxi);           /* by i, ties keep (p,id) order */
  free(key);
  u32 nb = 0;
  for (size_t k = 0; k < ns; ++k) nb += (k == 0 || sh->row[perm[k]].i != sh->row[perm[k - 1]].i);
  pl->nb = nb;
  pl->bprim = xnew(u32, nb ? nb : 1);
  pl->bgrp = xnew(u32, nb ? nb : 1);
  pl->ord = xnew(u32, nb ? nb : 1);
  pl->pflag = xnew(unsigned char, nb ? nb : 1);
  u32 *bof = xnew(u32, ns ? ns : 1);             /* boundary number of each row */
  {
    u32 b = 0;
    for (size_t k = 0; k < ns; ++k) {
      const struct xg_shared_row *r = sh->row + perm[k];
      if (k && r->i != sh->row[perm[k - 1]].i) ++b;
      if (k == 0 || r->i != sh->row[perm[k - 1]].i) {
        pl->bprim[b] = r->i;
        if (r->ord >> 32) xfail("gs_setup: more than 2^32 shared labels");
        pl->ord[b] = (u32)r->ord;
        pl->pflag[b] = (r->flags & XG_FLAGS_LOCAL) ? 1 : 0;
      }
      bof[perm[k]] = b;
    }
  }
  /* topology group of each boundary primary (both lists ascend by primary);
     with t == NULL the caller fills bgrp from the device topology */
  if (t) {
    size_t g = 0;
    for (u32 b = 0; b < nb; ++b) {
      while (g < t->ngrp && t->idx[t->gstart[g]] < pl->bprim[b]) ++g;
      if (g >= t->ngrp || t->idx[t->gstart[g]] != pl->bprim[b])
        xfail("gs_setup: shared label's primary index %u is not a local primary", pl->bprim[b]);
      pl->bgrp[b] = (u32)g;
    }
  }

  for (int d = 0; d < 2; ++d) {
    const unsigned mask = d == 0 ? XG_FLAGS_REMOTE : XG_FLAGS_LOCAL;
    /* slot numbers in (p,id) order; neighbour table */
    u32 *bi = xnew(u32, ns ? ns : 1);
    u32 count = 0, nn = 0, lastp = XG_NONE;
    for (size_t k = 0; k < ns; ++k) {
      const struct xg_shared_row *r = sh->row + k;
      if (r->flags & mask) { bi[k] = XG_NONE; continue; }
      bi[k] = count++;
      if (r->p != lastp) { lastp = r->p; ++nn; }
    }
    pl->total[d] = count;
    pl->nnb[d] = nn;
    pl->nb_rank[d] = xnew(u32, nn ? nn : 1);
    pl->nb_start[d] = xnew(u32, nn ? nn : 1);
    pl->nb_size[d] = xnew0(u32, nn ? nn : 1);
    nn = 0; lastp = XG_NONE;
    for (size_t k = 0; k < ns; ++k) {
      const struct xg_shared_row *r = sh->row + k;
      if (bi[k] == XG_NONE) continue;
      if (r->p != lastp) { lastp = r->p; pl->nb_rank[d][nn] = r->p; pl->nb_start[d][nn] = bi[k]; ++nn; }
      ++pl->nb_size[d][nn - 1];
    }
    /* per boundary primary: its slots, ascending remote rank (src/gs.upc:546-570) */
    pl->soff[d] = xnew0(u32, (size_t)nb + 1);
    pl->slot[d] = xnew(u32, count ? count : 1);
    for (size_t k = 0; k < ns; ++k) if (bi[k] != XG_NONE) ++pl->soff[d][bof[k] + 1];
    for (u32 b = 0; b < nb; ++b) pl->soff[d][b + 1] += pl->soff[d][b];
    u32 *fill = xnew(u32, nb ? nb : 1);
    memcpy(fill, pl->soff[d], (size_t)nb * sizeof(u32));
    for (size_t k = 0; k < ns; ++k)         /* rows are in (p,id) order */
      if (bi[k] != XG_NONE) pl->slot[d][fill[bof[k]]++] = bi[k];
    free(fill); free(bi);
    /* neighbour number of each slot-list entry: a neighbour's slots are one
       contiguous range of the numbering */
    pl->snb[d] = xnew(unsigned char, count ? count : 1);
    for (u32 s = 0; s < count; ++s) {
      u32 v = pl->slot[d][s], lo = 0, hi = nn;
      while (hi - lo > 1) { u32 mid = (lo + hi) / 2; if (pl->nb_start[d][mid] <= v) lo = mid; else hi = mid; }
      pl->snb[d][s] = (unsigned char)lo;
    }
  }
  /* union of both directions' neighbours, ascending */
  {
    u32 a = 0, b = 0, o = 0;
    pl->sym_rank = xnew(u32, (size_t)pl->nnb[0] + pl->nnb[1] + 1);
    while (a < pl->nnb[0] || b < pl->nnb[1]) {
      u32 ra = a < pl->nnb[0] ? pl->nb_rank[0][a] : XG_NONE, rb = b < pl->nnb[1] ? pl->nb_rank[1][b] : XG_NONE;
      u32 r = ra < rb ? ra : rb;
      pl->sym_rank[o++] = r;
      if (ra == r) ++a;
      if (rb == r) ++b;
    }
    pl->nsym = o;
  }
  free(bof); free(perm);
  (void)nbits64;
}

/* Tell every neighbour where its message starts in my numbering, so a
   receiver can pull its segment straight out of the sender's buffer.  A
   message I receive in direction d was numbered by its sender in direction d
   too: slots of direction 0 carry "remote unflagged" rows on the receiving
   side and "local unflagged" rows on the sending side of the same transfer,
   i.e. my recv table of direction rd pairs with the peer's send table of
   direction 1^rd (src/gs.upc:479, 578-584).                                  */
void xg_plan_exchange_offsets(struct xg_plan *pl, struct xg_world *w)
{
  const int np = xg_world_np(w), me = xg_world_rank(w);
  /* row: {from, to, direction, start} */
  struct row { u32 from, to, d, start; };
  size_t *cnt = xnew0(size_t, np), *pos = xnew(size_t, np), *rcnt = xnew(size_t, np), nrecv = 0;
  for (int d = 0; d < 2; ++d) for (u32 k = 0; k < pl->nnb[d]; ++k) ++cnt[pl->nb_rank[d][k]];
  size_t tot = 0; for (int p = 0; p < np; ++p) { pos[p] = tot; tot += cnt[p]; }
  struct row *send = xnew(struct row, tot ? tot : 1);
  for (int d = 0; d < 2; ++d)
    for (u32 k = 0; k < pl->nnb[d]; ++k) {
      struct row *r = send + pos[pl->nb_rank[d][k]]++;
      r->from = (u32)me; r->to = pl->nb_rank[d][k]; r->d = (u32)d; r->start = pl->nb_start[d][k];
    }
  struct row *got = (struct row *)xg_host_alltoallv(w, send, sizeof *send, cnt, rcnt, &nrecv);
  for (int d = 0; d < 2; ++d) {
    pl->nb_peer_start[d] = xnew0(u32, pl->nnb[d] ? pl->nnb[d] : 1);
    for (u32 k = 0; k < pl->nnb[d]; ++k) {
      int found = 0;
      for (size_t j = 0; j < nrecv; ++j)
        if (got[j].from == pl->nb_rank[d][k] && got[j].d == (u32)(1 ^ d)) { pl->nb_peer_start[d][k] = got[j].start; found = 1; break; }
      if (!found) xfail("gs_setup: neighbour %u has no matching message table", pl->nb_rank[d][k]);
    }
  }
  free(send); free(got); free(cnt); free(pos); free(rcnt);
  u64 mine = pl->total[0] > pl->total[1] ? pl->total[0] : pl->total[1], *all = xnew(u64, np);
  if (pl->total_shared > mine) mine = pl->total_shared;
  xg_host_allgather(w, &mine, sizeof mine, all);
  pl->max_total = 0;
  for (int p = 0; p < np; ++p) if (all[p] > pl->max_total) pl->max_total = all[p];
  free(all);
}
