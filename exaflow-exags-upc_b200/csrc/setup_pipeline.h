// setup_pipeline.h -- gs_setup's label sort, group discovery and map
// construction as data-parallel passes (one item per thread, plus scans, one
// stable radix sort and two reductions).
//
// Replaces nonzero_ids / local_map / flagged_primaries_map
// (src/gs.upc:83-112,329-370 of the reference) and reproduces, array for
// array, what this library's host builders in topo.c produce -- the ordering
// contract is the same: groups ascend by primary index; inside a group
// unflagged members come first, each class by ascending local index.
//
// The passes are written once against a small backend interface `B`:
//
//   T   *alloc<T>(count)              device allocation
//   void release(p)
//   void zero(p, bytes)               p[0..bytes) = 0
//   void fill_ff(p, bytes)            p[0..bytes) = 0xFF
//   void for_each(n, f)               f(i) for every i in [0,n), any order, in parallel
//   void sort_pairs(k, k2, v, v2, n, bits)   stable, ascending on the low `bits` bits of the
//                                     keys; k/v are swapped with k2/v2 as needed so that the
//                                     result is in k/v on return
//   void scan(p, n)                   exclusive prefix sum in place
//   u64  max_u64(p, n), sum_u32(p, n) reductions, result on the host
//   u32  get(p)                       one element to the host
//
// setup_cuda.cu instantiates it with the CUDA backend (kernels + CUB's device
// radix sort / scan / reduce); tests/c/setup_emul.cpp instantiates the same
// source with a sequential host backend, so the index logic is checked against
// topo.c on machines without a GPU.  That emulation is test code: nothing in
// the library calls it.
#ifndef EXAGS_SETUP_PIPELINE_H
#define EXAGS_SETUP_PIPELINE_H

#include "exags_internal.h"
#include "sell_window.h"

#if defined(__CUDACC__)
#define XG_DEV __device__
#else
#define XG_DEV
#endif

namespace xgsetup {

enum { OK = 0, ERR_ID_TOO_LARGE = 1, ERR_MAP_TOO_LARGE = 2, ERR_SELL_TOO_LARGE = 3, ERR_SELL_PACK = 4 };

static inline unsigned nbits64(u64 v)
{
  unsigned b = 0;
  while (v) { ++b; v >>= 1; }
  return b ? b : 1;
}

// ---------------------------------------------------------------------------
// topology: rows (label, flag, index) sorted, groups ordered by primary index
// ---------------------------------------------------------------------------
// d_id: n labels on the device, idbytes 4 or 8.
template <class B>
int topo_build(B &be, const void *d_id, int idbytes, size_t n, xcu_topo *t, u64 *maxabs_out)
{
  t->n = n; t->nnz = 0; t->ngrp = 0; t->any_flag = 0;
  t->id = 0; t->idx = 0; t->flag = 0; t->gstart = 0; t->rowgrp = 0;
  u64 *key = be.template alloc<u64>(n + 1), *key2 = be.template alloc<u64>(n + 1);
  u32 *val = be.template alloc<u32>(n + 1), *val2 = be.template alloc<u32>(n + 1);
  u32 *cnt = be.template alloc<u32>(4);
  be.zero(cnt, 4 * sizeof(u32));

  // key = |id|*2 + flagged, payload = index; unlabelled entries (id 0) get
  // key 0 and therefore sort in front of every labelled row
  {
    const int *id32 = (const int *)d_id;
    const i64 *id64 = (const i64 *)d_id;
    const int wide = idbytes == 8;
    be.for_each(n, [=] XG_DEV(size_t i) {
      const i64 v = wide ? id64[i] : (i64)id32[i];
      const u64 a = v < 0 ? (u64)(-(v + 1)) + 1u : (u64)v;
      key[i] = (a << 1) | (u64)(v < 0);
      val[i] = (u32)i;
    });
  }
  const u64 maxkey = n ? be.max_u64(key, n) : 0;
  if (maxabs_out) *maxabs_out = maxkey >> 1;
  if ((maxkey >> 1) >> 62) {
    be.release(key); be.release(key2); be.release(val); be.release(val2); be.release(cnt);
    return ERR_ID_TOO_LARGE;
  }
  // (|id|, flag) ascending; stable => ascending index inside each class
  be.sort_pairs(key, key2, val, val2, n, nbits64(maxkey));
  be.for_each(n, [=] XG_DEV(size_t j) {
    if (key[j] == 0 && (j + 1 == n || key[j + 1] != 0)) cnt[0] = (u32)(j + 1);
  });
  const size_t nzero = be.get(cnt);
  const size_t nnz = n - nzero;
  const u64 *K = key + nzero;
  const u32 *V = val + nzero;

  // label groups in label order: row j heads a group when its label differs
  // from the row before.  After the scan hs[j] = heads in [0,j).
  u32 *hs = be.template alloc<u32>(nnz + 1);
  be.for_each(nnz + 1, [=] XG_DEV(size_t j) {
    hs[j] = (j < nnz && (j == 0 || (K[j] >> 1) != (K[j - 1] >> 1))) ? 1u : 0u;
  });
  be.scan(hs, nnz + 1);
  const size_t ngrp = be.get(hs + nnz);

  // order the groups by primary index: primaries are distinct local indices,
  // so slot[primary] = group, and the occupied slots read in ascending order
  // give the new group order
  u32 *gs0 = be.template alloc<u32>(ngrp + 1);
  u32 *slot = be.template alloc<u32>(n + 1);
  be.fill_ff(slot, (n + 1) * sizeof(u32));
  be.for_each(nnz, [=] XG_DEV(size_t j) {
    if (hs[j + 1] != hs[j]) { const u32 g = hs[j]; gs0[g] = (u32)j; slot[V[j]] = g; }
    if (K[j] & 1u) cnt[1] = 1u;
  });
  be.for_each(1, [=] XG_DEV(size_t) { gs0[ngrp] = (u32)nnz; });
  u32 *os = be.template alloc<u32>(n + 1);
  be.for_each(n + 1, [=] XG_DEV(size_t i) { os[i] = (i < n && slot[i] != XG_NONE) ? 1u : 0u; });
  be.scan(os, n + 1);
  u32 *inv = be.template alloc<u32>(ngrp + 1);
  u32 *gstart = be.template alloc<u32>(ngrp + 1);
  be.for_each(n, [=] XG_DEV(size_t i) {
    const u32 gl = slot[i];
    if (gl == XG_NONE) return;
    const u32 k = os[i];
    inv[gl] = k;
    gstart[k] = gs0[gl + 1] - gs0[gl];
  });
  be.for_each(1, [=] XG_DEV(size_t) { gstart[ngrp] = 0; });
  be.scan(gstart, ngrp + 1);
  be.release(os); be.release(slot);

  u64 *gid = be.template alloc<u64>(ngrp + 1);
  u32 *idx = be.template alloc<u32>(nnz + 1);
  u32 *rowgrp = be.template alloc<u32>(nnz + 1);
  unsigned char *flag = be.template alloc<unsigned char>(nnz + 1);
  be.for_each(nnz, [=] XG_DEV(size_t j) {
    const u32 gl = hs[j + 1] - 1u;              // label-order group of row j
    const u32 k = inv[gl];
    const u32 pos = gstart[k] + ((u32)j - gs0[gl]);
    idx[pos] = V[j];
    flag[pos] = (unsigned char)(K[j] & 1u);
    rowgrp[pos] = k;
    if (hs[j + 1] != hs[j]) gid[k] = K[j] >> 1;
  });
  t->any_flag = be.get(cnt + 1) != 0;
  be.release(inv); be.release(gs0); be.release(hs);
  be.release(key); be.release(key2); be.release(val); be.release(val2); be.release(cnt);
  t->nnz = nnz; t->ngrp = ngrp;
  t->id = gid; t->idx = idx; t->flag = flag; t->gstart = gstart; t->rowgrp = rowgrp;
  return OK;
}

template <class B>
void topo_release(B &be, xcu_topo *t)
{
  be.release(t->id); be.release(t->idx); be.release(t->flag); be.release(t->gstart); be.release(t->rowgrp);
  t->id = 0; t->idx = 0; t->flag = 0; t->gstart = 0; t->rowgrp = 0;
}

// ---------------------------------------------------------------------------
// local maps: CSR of every group with >= 2 members (map_local[1] with the
// unflagged count per group standing in for map_local[0], src/gs.upc:329-357),
// the flagged-primary lists (src/gs.upc:359-370) and the sector count
// ---------------------------------------------------------------------------
// Compacts in ascending order: sc[0..n] holds 0/1 marks (sc[n] = 0) and is
// consumed; out[rank of i among the marked] = src[at[i]].  Returns the count.
template <class B>
u32 compact_marked(B &be, u32 *sc, size_t n, const u32 *src, const u32 *at, u32 **out_p)
{
  be.scan(sc, n + 1);
  const u32 cnt = be.get(sc + n);
  u32 *out = be.template alloc<u32>((size_t)cnt + 1);
  be.for_each(n, [=] XG_DEV(size_t i) { if (sc[i + 1] != sc[i]) out[sc[i]] = src[at[i]]; });
  *out_p = out;
  return cnt;
}

// CSR of the groups k with sel[k] & bit, every member, in group order.  sel is
// a per-group byte of selection bits computed by the caller.  With want_nunf
// the unflagged member count of each group is recorded too.
template <class B>
int csr_of(B &be, const xcu_topo *t, const unsigned char *sel, unsigned bit, int want_nunf, xcu_dmap *m)
{
  const size_t ngrp = t->ngrp, nnz = t->nnz;
  const u32 *gstart = t->gstart, *tidx = t->idx, *rowgrp = t->rowgrp;
  const unsigned char *flag = t->flag;
  m->ng = m->ni = 0; m->off = m->idx = m->nunf = 0;
  u32 *ks = be.template alloc<u32>(ngrp + 1);    // kept groups before k
  u32 *is = be.template alloc<u32>(ngrp + 1);    // their members before k
  be.for_each(ngrp + 1, [=] XG_DEV(size_t k) {
    const int keep = k < ngrp && (sel[k] & bit);
    ks[k] = keep ? 1u : 0u;
    is[k] = keep ? gstart[k + 1] - gstart[k] : 0u;
  });
  // the member total fits 32 bits because nnz does (n < 2^32 - 1)
  be.scan(ks, ngrp + 1);
  be.scan(is, ngrp + 1);
  const u32 ng = be.get(ks + ngrp), ni = be.get(is + ngrp);
  if ((size_t)ni >= (size_t)XG_NONE) { be.release(ks); be.release(is); return ERR_MAP_TOO_LARGE; }
  u32 *off = be.template alloc<u32>((size_t)ng + 1);
  u32 *idx = be.template alloc<u32>((size_t)ni + 1);
  u32 *nunf = want_nunf ? be.template alloc<u32>((size_t)ng + 1) : 0;
  be.for_each(ngrp, [=] XG_DEV(size_t k) {
    if (!(sel[k] & bit)) return;
    const u32 a = gstart[k], b = gstart[k + 1], g = ks[k];
    off[g] = is[k];
    if (nunf) {
      u32 unf = 0;
      for (u32 j = a; j < b && !flag[j]; ++j) ++unf;
      nunf[g] = unf;
    }
  });
  be.for_each(1, [=] XG_DEV(size_t) { off[ng] = ni; });
  be.for_each(nnz, [=] XG_DEV(size_t pos) {
    const u32 k = rowgrp[pos];
    if (!(sel[k] & bit)) return;
    idx[is[k] + ((u32)pos - gstart[k])] = tidx[pos];
  });
  be.release(ks); be.release(is);
  m->ng = ng; m->ni = ni; m->off = off; m->idx = idx; m->nunf = nunf;
  return OK;
}

// selection bits of a group
enum { SEL_ALL2 = 1,      // >= 2 members                                 (hall)
       SEL_INT2 = 2,      // >= 2 members and not a boundary group        (interior map)
       SEL_BND = 4,       // boundary group, any size                     (boundary CSR)
       SEL_BND2 = 8,      // boundary group with >= 2 members             (boundary sliced map)
       SEL_FP = 16,       // primary flagged                              (flagged_primaries)
       SEL_FS = 32,       // primary flagged, single member, not boundary (gs_init's part of flagged_primaries)
       SEL_INTM = 64 };   // SEL_INT2 or SEL_FS: the interior map the kernels run on.  A flagged single is a
                          // one-member group there: with no unflagged member it reduces to the identity, which
                          // is what gs_init writes over it (src/gs.upc:1182) -- no separate init pass.

// bgrp: nb boundary groups (device array, may be null when nb == 0)
template <class B>
unsigned char *select_groups(B &be, const xcu_topo *t, const u32 *bgrp, u32 nb)
{
  const size_t ngrp = t->ngrp;
  const u32 *gstart = t->gstart;
  const unsigned char *flag = t->flag;
  unsigned char *sel = be.template alloc<unsigned char>(ngrp + 1);
  be.zero(sel, ngrp + 1);
  be.for_each(nb, [=] XG_DEV(size_t b) { sel[bgrp[b]] = SEL_BND; });
  be.for_each(ngrp, [=] XG_DEV(size_t k) {
    const u32 a = gstart[k], sz = gstart[k + 1] - a;
    const unsigned bnd = sel[k] & SEL_BND;
    unsigned s = bnd;
    if (sz >= 2) s |= SEL_ALL2 | (bnd ? SEL_BND2 : (SEL_INT2 | SEL_INTM));
    if (flag[a]) s |= SEL_FP | ((sz == 1 && !bnd) ? (SEL_FS | SEL_INTM) : 0u);
    sel[k] = (unsigned char)s;
  });
  return sel;
}

template <class B>
u32 list_of(B &be, const xcu_topo *t, const unsigned char *sel, unsigned bit, u32 **out)
{
  const size_t ngrp = t->ngrp;
  u32 *sc = be.template alloc<u32>(ngrp + 1);
  be.for_each(ngrp + 1, [=] XG_DEV(size_t k) { sc[k] = (k < ngrp && (sel[k] & bit)) ? 1u : 0u; });
  const u32 cnt = compact_marked(be, sc, ngrp, t->idx, t->gstart, out);
  be.release(sc);
  return cnt;
}

template <class B>
u64 count_sectors(B &be, size_t n, const xcu_dmap *m)
{
  // distinct 32-byte sectors of a double field that hold a member: members of
  // different groups are distinct indices, so mark and count
  if (!m->ni) return 0;
  const u32 *idx = m->idx;
  const size_t nsec = (n * 8) / 32 + 1;
  u32 *mark = be.template alloc<u32>(nsec);
  be.zero(mark, nsec * sizeof(u32));
  be.for_each(m->ni, [=] XG_DEV(size_t q) { mark[((size_t)idx[q] * 8) / 32] = 1u; });
  const u64 c = be.sum_u32(mark, nsec);
  be.release(mark);
  return c;
}

// One rank: hall is also the interior map, unless there are flagged singles:
// then `in` is hall plus those as one-member groups (in->off stays null otherwise).
template <class B>
int local_maps(B &be, const xcu_topo *t, xcu_dmap *hall, xcu_dmap *in, u32 **fp, u32 *nfp, u32 **fs, u32 *nfs,
               u64 *sectors)
{
  *fp = *fs = 0; *nfp = *nfs = 0; *sectors = 0;
  in->ng = in->ni = 0; in->off = in->idx = in->nunf = 0;
  unsigned char *sel = select_groups(be, t, (const u32 *)0, 0);
  int rc = csr_of(be, t, sel, SEL_ALL2, t->any_flag, hall);
  if (rc) { be.release(sel); return rc; }
  if (t->any_flag) {
    *nfp = list_of(be, t, sel, SEL_FP, fp);
    *nfs = list_of(be, t, sel, SEL_FS, fs);
    if (*nfs) rc = csr_of(be, t, sel, SEL_INTM, 1, in);
    if (rc) { be.release(sel); return rc; }
  }
  be.release(sel);
  *sectors = count_sectors(be, t->n, hall);
  return OK;
}

// Several ranks: the exchange plan has named the boundary groups (bgrp,
// ascending).  hall as above; in = interior groups with >= 2 members plus the
// interior flagged singles as one-member groups; bnd =
// every boundary group with all its members and unflagged counts; bnd2 = the
// boundary groups with >= 2 members (built only for maps without flagged
// labels, where the boundary takes the sliced layout too); fs = flagged single
// primaries outside the boundary.  Restates the split of split_groups()
// (gs_host.c) -- the interior/boundary form of local_setup, src/gs.upc:1210-1217.
template <class B>
int split_maps(B &be, const xcu_topo *t, const u32 *bgrp, u32 nb, xcu_dmap *hall, xcu_dmap *in,
               xcu_dmap *bnd, xcu_dmap *bnd2, u32 **fp, u32 *nfp, u32 **fs, u32 *nfs, u64 *sectors)
{
  *fp = *fs = 0; *nfp = *nfs = 0; *sectors = 0;
  bnd2->ng = bnd2->ni = 0; bnd2->off = bnd2->idx = bnd2->nunf = 0;
  unsigned char *sel = select_groups(be, t, bgrp, nb);
  int rc = csr_of(be, t, sel, SEL_ALL2, t->any_flag, hall);
  if (!rc) rc = csr_of(be, t, sel, SEL_INTM, t->any_flag, in);
  if (!rc) rc = csr_of(be, t, sel, SEL_BND, 1, bnd);
  if (!rc && !t->any_flag) rc = csr_of(be, t, sel, SEL_BND2, 0, bnd2);
  if (rc) { be.release(sel); return rc; }
  if (t->any_flag) {
    *nfp = list_of(be, t, sel, SEL_FP, fp);
    *nfs = list_of(be, t, sel, SEL_FS, fs);
  }
  be.release(sel);
  *sectors = count_sectors(be, t->n, hall);
  return OK;
}

// ---------------------------------------------------------------------------
// which labels can another rank hold?  (the filters in front of unique_ids /
// shared_ids, src/gs.upc:117-131,190-244; host form in remote.c)
// ---------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define XG_ATOMIC_OR32(p, v) atomicOr((p), (v))
#else
#define XG_ATOMIC_OR32(p, v) (*(p) |= (v))
#endif
enum { CAND_BINS = 65536 };          // coarse label-range bins (one occupancy bit each)

// bins: CAND_BINS / 32 zeroed words; bit b is set when a label with (id >> shift) == b is present.
// The words are little-endian halves of the host's 64-bit bitmap words.
template <class B>
void cand_bins(B &be, const xcu_topo *t, unsigned shift, u32 *bins)
{
  const u64 *id = t->id;
  be.for_each(t->ngrp, [=] XG_DEV(size_t g) {
    const u64 b = id[g] >> shift;
    const u32 m = 1u << (b & 31u);
    if (!(bins[b >> 5] & m)) XG_ATOMIC_OR32(&bins[b >> 5], m);      // few distinct words: test before the atomic
  });
}

// c32[g] = 1 when label g lies in a bin some OTHER rank occupies (others = OR of their bins)
template <class B>
void cand_from_bins(B &be, const xcu_topo *t, unsigned shift, const u32 *others, u32 *c32)
{
  const u64 *id = t->id;
  const size_t ngrp = t->ngrp;
  be.for_each(ngrp + 1, [=] XG_DEV(size_t g) {
    if (g == ngrp) { c32[g] = 0; return; }
    const u64 b = id[g] >> shift;
    c32[g] = (others[b >> 5] >> (b & 31u)) & 1u;
  });
}

#define XG_LABEL_HASH(id_, bits_) (((id_) * 0x9E3779B97F4A7C15ull) >> (64 - (bits_)))

// presence bitmap (2^bits bits, zeroed) over a hash of the labels that passed the bins
template <class B>
void cand_presence(B &be, const xcu_topo *t, const u32 *c32, unsigned bits, u32 *bitmap)
{
  const u64 *id = t->id;
  be.for_each(t->ngrp, [=] XG_DEV(size_t g) {
    if (!c32[g]) return;
    const u64 h = XG_LABEL_HASH(id[g], bits);
    XG_ATOMIC_OR32(&bitmap[h >> 5], 1u << (h & 31u));
  });
}

// keep the candidates whose hash bit is set in the OR of the other ranks' presence bitmaps
template <class B>
void cand_refine(B &be, const xcu_topo *t, u32 *c32, unsigned bits, const u32 *others)
{
  const u64 *id = t->id;
  be.for_each(t->ngrp, [=] XG_DEV(size_t g) {
    if (!c32[g]) return;
    const u64 h = XG_LABEL_HASH(id[g], bits);
    c32[g] = (others[h >> 5] >> (h & 31u)) & 1u;
  });
}

// the candidates in group order: |label| and the primary's index, complemented when the primary is
// flagged (struct uq_row of remote.c).  c32 (ngrp + 1 entries, last 0) is consumed.
template <class B>
u32 cand_compact(B &be, const xcu_topo *t, u32 *c32, u64 **id_out, u32 **sif_out)
{
  const size_t ngrp = t->ngrp;
  be.scan(c32, ngrp + 1);
  const u32 cnt = be.get(c32 + ngrp);
  u64 *oid = be.template alloc<u64>((size_t)cnt + 1);
  u32 *osf = be.template alloc<u32>((size_t)cnt + 1);
  const u64 *id = t->id;
  const u32 *gstart = t->gstart, *idx = t->idx;
  const unsigned char *flag = t->flag;
  be.for_each(ngrp, [=] XG_DEV(size_t g) {
    if (c32[g + 1] == c32[g]) return;
    const u32 o = c32[g], a = gstart[g];
    oid[o] = id[g];
    osf[o] = flag[a] ? ~idx[a] : idx[a];
  });
  *id_out = oid; *sif_out = osf;
  return cnt;
}

// topology group headed by each of the nb primaries bprim (ascending, like the groups' own primaries);
// XG_NONE where a primary heads no group
template <class B>
void groups_of_primaries(B &be, const xcu_topo *t, const u32 *bprim, u32 nb, u32 *bgrp)
{
  const size_t ngrp = t->ngrp;
  const u32 *gstart = t->gstart, *idx = t->idx;
  be.for_each(nb, [=] XG_DEV(size_t b) {
    const u32 want = bprim[b];
    size_t lo = 0, hi = ngrp;                      // first group whose primary is >= want
    while (lo < hi) { const size_t mid = lo + (hi - lo) / 2; if (idx[gstart[mid]] < want) lo = mid + 1; else hi = mid; }
    bgrp[b] = (lo < ngrp && idx[gstart[lo]] == want) ? (u32)lo : XG_NONE;
  });
}

template <class B>
void dmap_release(B &be, xcu_dmap *m)
{
  be.release(m->off); be.release(m->idx); be.release(m->nunf);
  m->off = m->idx = m->nunf = 0; m->ng = m->ni = 0;
}

// ---------------------------------------------------------------------------
// sliced-ELL layout of a CSR map (struct xg_sell, exags_internal.h)
// ---------------------------------------------------------------------------
template <class B>
int sell_build(B &be, const xcu_dmap *m, xcu_dsell *s, int chain = 1)
{
  const u32 ng = m->ng;
  const u32 *off = m->off, *idx = m->idx, *nunf = m->nunf;
  s->nblock = s->nwin = 0; s->nidx = 0;
  s->mask = s->fmask = 0; s->idx = s->win_min = s->win_max = 0;
  s->rest.ng = s->rest.ni = 0; s->rest.off = s->rest.idx = s->rest.nunf = 0;
  if ((size_t)m->ni >= ((size_t)1 << 30)) return ERR_SELL_TOO_LARGE;   // padded widths are scanned in 32 bits

  // short groups (<= 8 members) go to the sliced layout, the others to a CSR
  u32 *ss = be.template alloc<u32>((size_t)ng + 1);   // short groups before g
  u32 *ls = be.template alloc<u32>((size_t)ng + 1);   // members of long groups before g
  be.for_each((size_t)ng + 1, [=] XG_DEV(size_t g) {
    const u32 sz = g < ng ? off[g + 1] - off[g] : 0u;
    ss[g] = (g < ng && sz <= XG_SELL_WMAX) ? 1u : 0u;
    ls[g] = sz > XG_SELL_WMAX ? sz : 0u;
  });
  be.scan(ss, (size_t)ng + 1);
  be.scan(ls, (size_t)ng + 1);
  const u32 nshort = be.get(ss + ng), nlong = ng - nshort, ilong = be.get(ls + ng);
  if (nlong) {
    u32 *roff = be.template alloc<u32>((size_t)nlong + 1);
    u32 *ridx = be.template alloc<u32>((size_t)ilong + 1);
    u32 *rnunf = nunf ? be.template alloc<u32>((size_t)nlong + 1) : 0;
    be.for_each(ng, [=] XG_DEV(size_t g) {
      const u32 a = off[g], sz = off[g + 1] - a;
      if (sz <= XG_SELL_WMAX) return;
      const u32 r = (u32)g - ss[g], o = ls[g];
      roff[r] = o;
      if (rnunf) rnunf[r] = nunf[g];
      for (u32 q = 0; q < sz; ++q) ridx[o + q] = idx[a + q];
    });
    be.for_each(1, [=] XG_DEV(size_t) { roff[nlong] = ilong; });
    s->rest.ng = nlong; s->rest.ni = ilong;
    s->rest.off = roff; s->rest.idx = ridx; s->rest.nunf = rnunf;
  }
  be.release(ls);
  if (!nshort) { be.release(ss); return OK; }

  // short groups in primary order and the running sum of their padded widths
  u32 *sg = be.template alloc<u32>((size_t)nshort + 1);
  u32 *P = be.template alloc<u32>((size_t)nshort + 1);
  be.for_each(ng, [=] XG_DEV(size_t g) {
    const u32 sz = off[g + 1] - off[g];
    if (sz > XG_SELL_WMAX) return;
    sg[ss[g]] = (u32)g;
    P[ss[g]] = xg_sell_width(sz);
  });
  be.for_each(1, [=] XG_DEV(size_t) { P[nshort] = 0; });
  be.scan(P, (size_t)nshort + 1);
  be.release(ss);

  // windows: greedy runs of consecutive groups whose padded widths fit
  // XG_SELL_WSLOTS.  nxt[k] = first group of the window after one starting at
  // k; the chain from 0 is then followed by one thread (a window holds at
  // least 256 groups, so the chain is short).
  u32 *nxt = be.template alloc<u32>((size_t)nshort + 1);
  be.for_each(nshort, [=] XG_DEV(size_t k) {
    // largest e in [k, nshort] with P[e] - P[k] <= WSLOTS; widths lie in [2,8]
    const u32 lim = P[k] + XG_SELL_WSLOTS;
    size_t lo = k + XG_SELL_WSLOTS / 8, hi = k + XG_SELL_WSLOTS / 2;   // P[lo] fits, P[hi+1] cannot
    if (lo > nshort) lo = nshort;
    if (hi > nshort) hi = nshort;
    while (lo < hi) {
      const size_t mid = lo + (hi - lo + 1) / 2;
      if (P[mid] <= lim) lo = mid; else hi = mid - 1;
    }
    nxt[k] = (u32)lo;
  });
  const size_t wcap = (size_t)nshort / (XG_SELL_WSLOTS / 8) + 2;
  u32 *wfirst = be.template alloc<u32>(wcap + 1);
  u32 *cnt = be.template alloc<u32>(4);
  be.zero(cnt, 4 * sizeof(u32));
  be.for_each(1, [=] XG_DEV(size_t) {
    u32 k = 0, w = 0;
    wfirst[0] = 0;
    while (k < nshort && w < wcap) { k = nxt[k]; wfirst[++w] = k; }
    cnt[0] = w;
  });
  const u32 nwin = be.get(cnt);
  be.release(nxt); be.release(P);
  if (((size_t)nwin * XG_SELL_WB) >> 23) {
    be.release(sg); be.release(wfirst); be.release(cnt);
    return ERR_SELL_TOO_LARGE;
  }
  const u32 nblock = nwin * XG_SELL_WB;
  const size_t nidx = (size_t)nblock * 256;
  u32 *oidx = be.template alloc<u32>(nidx + 1);
  unsigned char *mask = be.template alloc<unsigned char>((size_t)nblock * 32 + 1);
  unsigned char *fmask = nunf ? be.template alloc<unsigned char>((size_t)nblock * 32 + 1) : 0;
  u32 *wmin = be.template alloc<u32>((size_t)nwin + 1), *wmax = be.template alloc<u32>((size_t)nwin + 1);
  be.zero(mask + (size_t)nblock * 32, 1);
  if (fmask) be.zero(fmask + (size_t)nblock * 32, 1);
  be.for_each(nwin, [=] XG_DEV(size_t w) {
    if (xg_sell_fill_window(w, wfirst[w], wfirst[w + 1], sg, off, idx, nunf, oidx, mask, fmask, wmin, wmax, chain))
      cnt[1] = 1u;
  });
  const u32 bad = be.get(cnt + 1);
  be.release(sg); be.release(wfirst); be.release(cnt);
  s->nblock = nblock; s->nwin = nwin; s->nidx = nidx;
  s->mask = mask; s->fmask = fmask; s->idx = oidx; s->win_min = wmin; s->win_max = wmax;
  return bad ? ERR_SELL_PACK : OK;
}

template <class B>
void sell_release(B &be, xcu_dsell *s)
{
  be.release(s->mask); be.release(s->fmask); be.release(s->idx); be.release(s->win_min); be.release(s->win_max);
  dmap_release(be, &s->rest);
  s->mask = s->fmask = 0; s->idx = s->win_min = s->win_max = 0;
}

}  // namespace xgsetup
#endif
