// setup_emul.cpp -- TEST CODE.  Instantiates the device gs_setup passes of
// csrc/setup_pipeline.h with a sequential host backend and compares every
// array they produce with the host builders of csrc/topo.c (xg_topo_build,
// xg_topo_local_map, xg_topo_flagged_primaries, xg_sell_build).  This checks
// the index logic of the passes on machines without a GPU; the CUDA backend
// (kernel launch, CUB sort / scan / reduce) is exercised by the -m gpu tests.
// Nothing in the library links or calls this file.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <set>
#include <vector>
#include "setup_pipeline.h"

namespace {
struct HostBackend {
  bool reverse = false;       // run for_each back to front: passes must not depend on item order
  template <class T> T *alloc(size_t count) { return (T *)malloc((count ? count : 1) * sizeof(T)); }
  void release(void *p) { free(p); }
  void zero(void *p, size_t bytes) { memset(p, 0, bytes); }
  void fill_ff(void *p, size_t bytes) { memset(p, 0xFF, bytes); }
  template <class F> void for_each(size_t n, F f)
  {
    if (reverse) for (size_t i = n; i-- > 0;) f(i);
    else for (size_t i = 0; i < n; ++i) f(i);
  }
  void sort_pairs(u64 *&k, u64 *&k2, u32 *&v, u32 *&v2, size_t n, unsigned bits)
  {
    const u64 m = bits >= 64 ? ~(u64)0 : (((u64)1 << bits) - 1);
    std::vector<size_t> p(n);
    std::iota(p.begin(), p.end(), (size_t)0);
    std::stable_sort(p.begin(), p.end(), [&](size_t a, size_t b) { return (k[a] & m) < (k[b] & m); });
    for (size_t i = 0; i < n; ++i) { k2[i] = k[p[i]]; v2[i] = v[p[i]]; }
    std::swap(k, k2); std::swap(v, v2);     // result must be reachable through k / v
  }
  void scan(u32 *p, size_t n) { u32 run = 0; for (size_t i = 0; i < n; ++i) { u32 x = p[i]; p[i] = run; run += x; } }
  u64 max_u64(const u64 *p, size_t n) { u64 m = 0; for (size_t i = 0; i < n; ++i) m = std::max(m, p[i]); return m; }
  u64 sum_u32(const u32 *p, size_t n) { u64 s = 0; for (size_t i = 0; i < n; ++i) s += p[i]; return s; }
  u32 get(const u32 *p) { return *p; }
};

template <class T> bool same(const T *a, const T *b, size_t n) { return n == 0 || memcmp(a, b, n * sizeof(T)) == 0; }

u64 host_sectors(const xg_hmap *m)
{
  if (!m->ni) return 0;
  u32 maxi = 0;
  for (u32 k = 0; k < m->ni; ++k) maxi = std::max(maxi, m->idx[k]);
  std::vector<unsigned char> mark(((size_t)maxi * 8) / 32 + 1, 0);
  for (u32 k = 0; k < m->ni; ++k) mark[((size_t)m->idx[k] * 8) / 32] = 1;
  u64 c = 0;
  for (unsigned char x : mark) c += x;
  return c;
}

int compare_sell(const xcu_dsell &d, const xg_sell &h)
{
  if (d.nblock != h.nblock || d.nwin != h.nwin || d.nidx != h.nidx) return 40;
  if (h.nblock) {
    if (!same(d.idx, h.idx, h.nidx)) return 41;
    if (!same(d.mask, h.mask, (size_t)h.nblock * 32 + 1)) return 42;
    if ((d.fmask != 0) != (h.fmask != 0)) return 43;
    if (h.fmask && !same(d.fmask, h.fmask, (size_t)h.nblock * 32 + 1)) return 44;
    if (!same(d.win_min, h.win_min, h.nwin) || !same(d.win_max, h.win_max, h.nwin)) return 45;
  }
  if (d.rest.ng != h.rest.ng || d.rest.ni != h.rest.ni) return 46;
  if (h.rest.ng) {
    if (!same(d.rest.off, h.rest.off, (size_t)h.rest.ng + 1) || !same(d.rest.idx, h.rest.idx, h.rest.ni)) return 47;
    if ((d.rest.nunf != 0) != (h.rest.nunf != 0)) return 48;
    if (h.rest.nunf && !same(d.rest.nunf, h.rest.nunf, h.rest.ng)) return 49;
  }
  return 0;
}
}  // namespace

// The interior/boundary split: bgrp = nb group numbers (ascending) standing in
// for the boundary groups an exchange plan would name.  Expected output from
// plain loops over the host topology (the rule of split_groups(), gs_host.c).
extern "C" int setup_emul_split(const void *id, int idbytes, size_t n, const u32 *bgrp, u32 nb, int reverse)
{
  std::vector<i64> wide(n ? n : 1);
  for (size_t i = 0; i < n; ++i) wide[i] = idbytes == 8 ? ((const i64 *)id)[i] : (i64)((const int *)id)[i];
  xg_topo ht;
  xg_topo_build(&ht, wide.data(), n);
  bool any_flag = false;
  for (size_t j = 0; j < ht.nnz; ++j) any_flag |= ht.flag[j] != 0;
  std::vector<u32> in_off, in_idx, in_unf, b_off, b_idx, b_unf, b2_off, b2_idx, fs;
  {
    u32 bb = 0;
    for (size_t g = 0; g < ht.ngrp; ++g) {
      const u32 a = ht.gstart[g], e = ht.gstart[g + 1];
      u32 unf = 0;
      for (u32 j = a; j < e && !ht.flag[j]; ++j) ++unf;
      if (bb < nb && bgrp[bb] == g) {
        ++bb;
        b_off.push_back((u32)b_idx.size()); b_unf.push_back(unf);
        b_idx.insert(b_idx.end(), ht.idx + a, ht.idx + e);
        if (e - a >= 2) { b2_off.push_back((u32)b2_idx.size()); b2_idx.insert(b2_idx.end(), ht.idx + a, ht.idx + e); }
      } else {
        if (e - a >= 2 || ht.flag[a]) {        // a flagged single is a one-member group of the interior map
          in_off.push_back((u32)in_idx.size()); in_unf.push_back(unf);
          in_idx.insert(in_idx.end(), ht.idx + a, ht.idx + e);
        }
        if (e - a < 2 && ht.flag[a]) fs.push_back(ht.idx[a]);
      }
    }
    in_off.push_back((u32)in_idx.size()); b_off.push_back((u32)b_idx.size()); b2_off.push_back((u32)b2_idx.size());
  }
  HostBackend be;
  be.reverse = reverse != 0;
  xcu_topo dt;
  int out = 0;
  if (xgsetup::topo_build(be, id, idbytes, n, &dt, (u64 *)0)) { xg_topo_free(&ht); return 10; }
  xcu_dmap hall, in, bnd, bnd2;
  u32 *dfp = 0, *dfs = 0, dnfp = 0, dnfs = 0;
  u64 sectors = 0;
  u32 *dbgrp = be.alloc<u32>((size_t)nb + 1);
  if (nb) memcpy(dbgrp, bgrp, (size_t)nb * sizeof(u32));
  if (xgsetup::split_maps(be, &dt, dbgrp, nb, &hall, &in, &bnd, &bnd2, &dfp, &dnfp, &dfs, &dnfs, &sectors)) out = 50;
  else if (in.ng + 1 != in_off.size() || in.ni != in_idx.size()) out = 51;
  else if (!same(in.off, in_off.data(), in_off.size()) || !same(in.idx, in_idx.data(), in_idx.size())) out = 52;
  else if ((in.nunf != 0) != any_flag || (any_flag && !same(in.nunf, in_unf.data(), in_unf.size()))) out = 53;
  else if (bnd.ng != nb || bnd.ni != b_idx.size()) out = 54;
  else if (!same(bnd.off, b_off.data(), b_off.size()) || !same(bnd.idx, b_idx.data(), b_idx.size())) out = 55;
  else if (!bnd.nunf || !same(bnd.nunf, b_unf.data(), b_unf.size())) out = 56;
  else if (!any_flag && (bnd2.ng + 1 != b2_off.size() || bnd2.ni != b2_idx.size() ||
                         !same(bnd2.off, b2_off.data(), b2_off.size()) || !same(bnd2.idx, b2_idx.data(), b2_idx.size()))) out = 57;
  else if (any_flag && (bnd2.ng || bnd2.off)) out = 58;
  else if (dnfs != fs.size() || (any_flag && !same(dfs, fs.data(), fs.size()))) out = 59;
  if (!out) {
    xg_hmap hh;
    xg_topo_local_map(&ht, 1, &hh);
    if (hall.ng != hh.ng || hall.ni != hh.ni || !same(hall.idx, hh.idx, hh.ni) || sectors != host_sectors(&hh)) out = 60;
    xg_hmap_free(&hh);
  }
  if (out != 50) {
    xgsetup::dmap_release(be, &hall); xgsetup::dmap_release(be, &in);
    xgsetup::dmap_release(be, &bnd); xgsetup::dmap_release(be, &bnd2);
  }
  be.release(dfp); be.release(dfs); be.release(dbgrp);
  xgsetup::topo_release(be, &dt);
  xg_topo_free(&ht);
  return out;
}

// Trips through L1 the field loads of one gs() make with this layout: distinct
// 128-byte lines (of a double field) per load instruction of a warp, summed
// over blocks and slots, honouring the phase bits.  A planning figure for the
// pair ordering of sell_window.h, not a test.
extern "C" unsigned long long setup_emul_line_visits(const void *id, int idbytes, size_t n)
{
  std::vector<i64> wide(n ? n : 1);
  for (size_t i = 0; i < n; ++i) wide[i] = idbytes == 8 ? ((const i64 *)id)[i] : (i64)((const int *)id)[i];
  xg_topo ht; xg_topo_build(&ht, wide.data(), n);
  xg_hmap hh; xg_topo_local_map(&ht, 1, &hh);
  xg_sell hs; xg_sell_build(&hs, &hh);
  unsigned long long visits = 0;
  const char *lsenv = getenv("EXAGS_EMUL_LINE_SHIFT");        // 4: doubles (default), 5: 4-byte fields
  const unsigned lshift = lsenv ? (unsigned)atoi(lsenv) : 4u;
  for (u32 b = 0; b < hs.nblock; ++b)
    for (int j = 0; j < 8; ++j) {
      u32 lines[32]; int nl = 0;
      for (int lane = 0; lane < 32; ++lane) {
        const size_t L = (size_t)b * 32 + lane;
        const unsigned ph = hs.mask[L] & 0xAAu;
        const int slot = ((ph >> ((j | 1))) & 1u) ? (j ^ 1) : j;
        const u32 v = hs.idx[L * 8 + slot];
        if (v == XG_NONE) continue;
        const u32 line = v >> lshift;
        bool seen = false;
        for (int q = 0; q < nl; ++q) if (lines[q] == line) { seen = true; break; }
        if (!seen) lines[nl++] = line;
      }
      visits += nl;
    }
  xg_sell_free(&hs); xg_hmap_free(&hh); xg_topo_free(&ht);
  return visits;
}

// The sliced layout decodes back to the map it was built from: every group of
// at most 8 members appears exactly once, members in order, inside a lane, in
// the window that holds its primary-order neighbours; longer groups are in the
// residual CSR; phase bits (odd mask bits) sit only on pairs.  0 = ok.
extern "C" int setup_emul_sell_roundtrip(const void *id, int idbytes, size_t n)
{
  std::vector<i64> wide(n ? n : 1);
  for (size_t i = 0; i < n; ++i) wide[i] = idbytes == 8 ? ((const i64 *)id)[i] : (i64)((const int *)id)[i];
  xg_topo ht; xg_topo_build(&ht, wide.data(), n);
  xg_hmap hh; xg_topo_local_map(&ht, 1, &hh);
  xg_sell hs; xg_sell_build(&hs, &hh);
  int out = 0;
  // groups found in the layout: (primary, first slot position)
  std::vector<std::pair<u32, size_t>> found;
  for (size_t L = 0; L < (size_t)hs.nblock * 32 && !out; ++L) {
    const unsigned mk = hs.mask[L], start = mk & 0x55u, ph = mk & 0xAAu;
    const u32 *sl = hs.idx + L * 8;
    for (int j = 0; j < 8; ++j) {
      if (sl[j] == XG_NONE) { if ((start >> j) & 1u) out = 70; continue; }
      if (j == 0 && !(start & 1u)) out = 71;                 // a used lane starts with a group
      if ((start >> j) & 1u) found.push_back({sl[j], L * 8 + j});
    }
    for (int j = 0; j < 8; j += 2)
      if ((ph >> (j + 1)) & 1u) {          // phase only where both slots hold members of one group
        if (sl[j] == XG_NONE || sl[j + 1] == XG_NONE) out = 72;
      }
  }
  std::sort(found.begin(), found.end());
  size_t f = 0, r = 0;
  for (u32 g = 0; g < hh.ng && !out; ++g) {
    const u32 a = hh.off[g], sz = hh.off[g + 1] - a;
    if (sz > XG_SELL_WMAX) {
      if (r >= hs.rest.ng || hs.rest.off[r + 1] - hs.rest.off[r] != sz ||
          memcmp(hs.rest.idx + hs.rest.off[r], hh.idx + a, sz * sizeof(u32))) out = 73;
      ++r;
      continue;
    }
    if (f >= found.size() || found[f].first != hh.idx[a]) { out = 74; break; }
    const size_t pos = found[f].second;
    if ((pos & 7) + sz > 8) { out = 75; break; }               // inside one lane
    if (memcmp(hs.idx + pos, hh.idx + a, sz * sizeof(u32))) { out = 76; break; }
    // the slot after the group is free, a new group, or the lane's end
    if ((pos & 7) + sz < 8 && hs.idx[pos + sz] != XG_NONE && !((hs.mask[pos >> 3] >> ((pos & 7) + sz)) & 1u)) { out = 77; break; }
    if (hs.win_min[pos / XG_SELL_WSLOTS] > hh.idx[a] || hs.win_max[pos / XG_SELL_WSLOTS] < hh.idx[a + sz - 1]) { out = 78; break; }
    ++f;
  }
  if (!out && (f != found.size() || r != hs.rest.ng)) out = 79;
  xg_sell_free(&hs); xg_hmap_free(&hh); xg_topo_free(&ht);
  return out;
}

// group count of the host topology (so a test can pick boundary groups)
extern "C" size_t setup_emul_ngroups(const void *id, int idbytes, size_t n)
{
  std::vector<i64> wide(n ? n : 1);
  for (size_t i = 0; i < n; ++i) wide[i] = idbytes == 8 ? ((const i64 *)id)[i] : (i64)((const int *)id)[i];
  xg_topo ht;
  xg_topo_build(&ht, wide.data(), n);
  const size_t g = ht.ngrp;
  xg_topo_free(&ht);
  return g;
}

// Returns 0 when every array matches, else a stage code (10s topology, 20s
// local map, 30s flagged lists / sectors, 40s sliced layout).
extern "C" int setup_emul_compare(const void *id, int idbytes, size_t n, int reverse)
{
  std::vector<i64> wide(n ? n : 1);
  for (size_t i = 0; i < n; ++i) wide[i] = idbytes == 8 ? ((const i64 *)id)[i] : (i64)((const int *)id)[i];
  xg_topo ht;
  xg_topo_build(&ht, wide.data(), n);
  xg_hmap hh;
  xg_topo_local_map(&ht, 1, &hh);
  u32 hnfp = 0, hnfs = 0;
  u32 *hfp = xg_topo_flagged_primaries(&ht, 0, &hnfp), *hfs = xg_topo_flagged_primaries(&ht, 1, &hnfs);
  xg_sell hs;
  xg_sell_build(&hs, &hh);

  HostBackend be;
  be.reverse = reverse != 0;
  xcu_topo dt;
  u64 maxabs = 0;
  int rc = xgsetup::topo_build(be, id, idbytes, n, &dt, &maxabs), out = 0;
  xcu_dmap dh; memset(&dh, 0, sizeof dh);
  xcu_dmap din; memset(&din, 0, sizeof din);
  xcu_dsell ds; memset(&ds, 0, sizeof ds);
  u32 *dfp = 0, *dfs = 0, dnfp = 0, dnfs = 0;
  u64 sectors = 0;
  if (rc) { out = 10; goto done; }
  if (dt.nnz != ht.nnz || dt.ngrp != ht.ngrp) { out = 11; goto done; }
  if (!same(dt.id, ht.id, ht.ngrp)) { out = 12; goto done; }
  if (!same(dt.idx, ht.idx, ht.nnz)) { out = 13; goto done; }
  if (!same(dt.flag, ht.flag, ht.nnz)) { out = 14; goto done; }
  if (!same(dt.gstart, ht.gstart, ht.ngrp + 1)) { out = 15; goto done; }
  for (size_t g = 0; g < ht.ngrp; ++g)
    for (u32 j = ht.gstart[g]; j < ht.gstart[g + 1]; ++j) if (dt.rowgrp[j] != g) { out = 16; goto done; }

  rc = xgsetup::local_maps(be, &dt, &dh, &din, &dfp, &dnfp, &dfs, &dnfs, &sectors);
  if (rc) { out = 20; goto done; }
  if ((din.off != 0) != (dnfs != 0)) { out = 25; goto done; }
  if (din.off) {
    // hall plus the flagged singles as one-member groups, in group order
    std::vector<u32> eo, ei, eu;
    for (size_t g = 0; g < ht.ngrp; ++g) {
      const u32 a = ht.gstart[g], e = ht.gstart[g + 1];
      if (e - a < 2 && !ht.flag[a]) continue;
      u32 unf = 0;
      for (u32 j = a; j < e && !ht.flag[j]; ++j) ++unf;
      eo.push_back((u32)ei.size()); eu.push_back(unf);
      ei.insert(ei.end(), ht.idx + a, ht.idx + e);
    }
    eo.push_back((u32)ei.size());
    if (din.ng + 1 != eo.size() || din.ni != ei.size() || !din.nunf) { out = 26; goto done; }
    if (!same(din.off, eo.data(), eo.size()) || !same(din.idx, ei.data(), ei.size()) ||
        !same(din.nunf, eu.data(), eu.size())) { out = 27; goto done; }
    // and its sliced layout against the host builder's on the same map
    xg_hmap hin; memset(&hin, 0, sizeof hin);
    hin.ng = din.ng; hin.ni = din.ni; hin.off = eo.data(); hin.idx = ei.data(); hin.nunf = eu.data();
    xg_sell hs2; xg_sell_build(&hs2, &hin);
    xcu_dsell ds2; memset(&ds2, 0, sizeof ds2);
    rc = xgsetup::sell_build(be, &din, &ds2, xg_sell_pairs_enabled());
    const int c2 = rc ? 39 : compare_sell(ds2, hs2);
    xgsetup::sell_release(be, &ds2); xg_sell_free(&hs2);
    if (c2) { out = 100 + c2; goto done; }
  }
  if (dh.ng != hh.ng || dh.ni != hh.ni) { out = 21; goto done; }
  if (!same(dh.off, hh.off, (size_t)hh.ng + 1) || !same(dh.idx, hh.idx, hh.ni)) { out = 22; goto done; }
  if ((dh.nunf != 0) != (hh.nunf != 0)) { out = 23; goto done; }
  if (hh.nunf && !same(dh.nunf, hh.nunf, hh.ng)) { out = 24; goto done; }
  if (dt.any_flag) {
    if (dnfp != hnfp || !same(dfp, hfp, hnfp)) { out = 30; goto done; }
    if (dnfs != hnfs || !same(dfs, hfs, hnfs)) { out = 31; goto done; }
  } else if (hnfp || hnfs) { out = 32; goto done; }
  if (sectors != host_sectors(&hh)) { out = 33; goto done; }

  rc = xgsetup::sell_build(be, &dh, &ds, xg_sell_pairs_enabled());
  if (rc) { out = 39; goto done; }
  out = compare_sell(ds, hs);
done:
  xgsetup::sell_release(be, &ds);
  xgsetup::dmap_release(be, &dh);
  xgsetup::dmap_release(be, &din);
  be.release(dfp); be.release(dfs);
  xgsetup::topo_release(be, &dt);
  xg_sell_free(&hs); xg_hmap_free(&hh); free(hfp); free(hfs); xg_topo_free(&ht);
  return out;
}

// The sliced layout itself, for layout studies in Python: returns the number of
// blocks; idx_out (cap_blocks * 256 uints) and mask_out (cap_blocks * 32 bytes)
// are filled when they are large enough.
extern "C" size_t setup_emul_sell_dump(const void *id, int idbytes, size_t n, u32 *idx_out,
                                       unsigned char *mask_out, size_t cap_blocks)
{
  std::vector<i64> wide(n ? n : 1);
  for (size_t i = 0; i < n; ++i) wide[i] = idbytes == 8 ? ((const i64 *)id)[i] : (i64)((const int *)id)[i];
  xg_topo ht; xg_topo_build(&ht, wide.data(), n);
  xg_hmap hh; xg_topo_local_map(&ht, 1, &hh);
  xg_sell hs; xg_sell_build(&hs, &hh);
  const size_t nb = hs.nblock;
  if (nb <= cap_blocks && idx_out && mask_out) {
    memcpy(idx_out, hs.idx, nb * 256 * sizeof(u32));
    memcpy(mask_out, hs.mask, nb * 32);
  }
  xg_sell_free(&hs); xg_hmap_free(&hh); xg_topo_free(&ht);
  return nb;
}

// The candidate filters of the np > 1 setup (setup_pipeline.h: cand_bins, cand_from_bins,
// cand_presence, cand_refine, cand_compact, groups_of_primaries) for two label sets standing in
// for two ranks.  What must hold whatever the hash does: a label both ranks hold is a candidate on
// rank A (the filters only err towards sending); every candidate row is (|label|, primary index,
// complemented when the primary is flagged) of its group, in group order; and every primary is
// found again by the group lookup.  Returns 0, or a code naming the first violated property.
extern "C" int setup_emul_candidates(const void *idA, size_t nA, const void *idB, size_t nB, int idbytes, int reverse)
{
  HostBackend be;
  be.reverse = reverse != 0;
  xcu_topo ta, tb;
  if (xgsetup::topo_build(be, idA, idbytes, nA, &ta, (u64 *)0)) return 10;
  if (xgsetup::topo_build(be, idB, idbytes, nB, &tb, (u64 *)0)) return 11;
  u64 gmax = 0;
  for (size_t g = 0; g < ta.ngrp; ++g) gmax = std::max(gmax, ta.id[g]);
  for (size_t g = 0; g < tb.ngrp; ++g) gmax = std::max(gmax, tb.id[g]);
  unsigned shift = 0;
  while ((gmax >> shift) >= (u64)xgsetup::CAND_BINS) ++shift;
  const size_t bw = xgsetup::CAND_BINS / 32;
  std::vector<u32> binsA(bw, 0), binsB(bw, 0);
  xgsetup::cand_bins(be, &ta, shift, binsA.data());
  xgsetup::cand_bins(be, &tb, shift, binsB.data());
  std::vector<u32> cA(ta.ngrp + 1), cB(tb.ngrp + 1);
  xgsetup::cand_from_bins(be, &ta, shift, binsB.data(), cA.data());
  xgsetup::cand_from_bins(be, &tb, shift, binsA.data(), cB.data());
  const u64 maxn = std::max<u64>(1, std::max(be.sum_u32(cA.data(), ta.ngrp), be.sum_u32(cB.data(), tb.ngrp)));
  unsigned bits = 20;
  while (bits < 34 && ((u64)1 << bits) < 16 * maxn) ++bits;
  const size_t words = (size_t)1 << (bits - 5);
  std::vector<u32> bmA(words, 0), bmB(words, 0);
  xgsetup::cand_presence(be, &ta, cA.data(), bits, bmA.data());
  xgsetup::cand_presence(be, &tb, cB.data(), bits, bmB.data());
  xgsetup::cand_refine(be, &ta, cA.data(), bits, bmB.data());
  std::vector<u32> flagged(cA.begin(), cA.end());     // cand_compact consumes its input
  u64 *cid = 0; u32 *csf = 0;
  const u32 cnt = xgsetup::cand_compact(be, &ta, cA.data(), &cid, &csf);
  int out = 0;
  std::set<u64> inB(tb.id, tb.id + tb.ngrp);
  u32 o = 0;
  for (size_t g = 0; g < ta.ngrp && !out; ++g) {
    const bool is_cand = flagged[g] != 0;
    if (inB.count(ta.id[g]) && !is_cand) out = 20;                 // a shared label was dropped
    if (!is_cand) continue;
    if (o >= cnt) { out = 21; break; }
    const u32 a = ta.gstart[g];
    if (cid[o] != ta.id[g] || csf[o] != (ta.flag[a] ? ~ta.idx[a] : ta.idx[a])) out = 22;
    ++o;
  }
  if (!out && o != cnt) out = 23;
  if (!out && ta.ngrp) {
    std::vector<u32> prim(ta.ngrp), grp(ta.ngrp);
    for (size_t g = 0; g < ta.ngrp; ++g) prim[g] = ta.idx[ta.gstart[g]];
    xgsetup::groups_of_primaries(be, &ta, prim.data(), (u32)ta.ngrp, grp.data());
    for (size_t g = 0; g < ta.ngrp; ++g) if (grp[g] != (u32)g) { out = 24; break; }
    u32 none_in = (u32)nA, none_out = 0;                           // an index that heads no group
    for (u32 i = 0; i < (u32)nA; ++i) {
      bool heads = false;
      for (size_t g = 0; g < ta.ngrp; ++g) if (prim[g] == i) { heads = true; break; }
      if (!heads) { none_in = i; break; }
    }
    if (none_in < (u32)nA) {
      xgsetup::groups_of_primaries(be, &ta, &none_in, 1, &none_out);
      if (none_out != XG_NONE) out = 25;
    }
  }
  be.release(cid); be.release(csf);
  xgsetup::topo_release(be, &ta); xgsetup::topo_release(be, &tb);
  return out;
}
