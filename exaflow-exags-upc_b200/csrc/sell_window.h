/* sell_window.h -- the layout of one window of the sliced-ELL map, written once
 * and run in two places: by the host builder (topo.c, one window per OpenMP
 * iteration) and by the device builder (setup_cuda.cu, one window per
 * thread).  Plain C that is also valid CUDA C++.
 *
 * A window is XG_SELL_WB blocks x 32 lanes x 8 member slots.  The groups
 * sg[first .. last) (consecutive in primary order, every one with at most 8
 * members) are dealt by descending padded width {8,4,2}, stable in primary
 * order inside a width class, lane after lane, block after block; a lane that
 * has no room for the class is skipped.  See exags_internal.h (struct xg_sell)
 * for why the layout looks like this.
 */
#ifndef EXAGS_SELL_WINDOW_H
#define EXAGS_SELL_WINDOW_H

#if defined(__CUDACC__)
#define XG_HD __host__ __device__
#else
#define XG_HD
#endif

#define XG_SELL_LANES (XG_SELL_WB * 32u)          /* lanes per window        */
#define XG_SELL_WSLOTS (XG_SELL_WB * 256u)        /* member slots per window */

XG_HD static inline u32 xg_sell_width(u32 size) { return size <= 2 ? 2u : size <= 4 ? 4u : 8u; }

/* Dealing order of the groups of one width class of a window, and a phase bit
   for each.  Written for pairs (groups of exactly two members), where it
   matters most; wider groups are ordered by their first two members in the same
   way, and their phase applies to every complete slot pair of the group.

   A lane loads its 8 slots with 8 load instructions; instruction j of a warp
   therefore touches slot j of 32 different lanes, and costs one trip through
   L1 per distinct 128-byte line among those 32 addresses.  On a hexahedral
   spectral-element mesh the pairs across an x-face are {last node of a row of
   element e, first node of the same row of element e+1}: dealt in primary
   order, the 32 primaries of an instruction lie in 32 different rows, one or
   two useful values per line.  Two things put more of them into one line:

   twins    two pairs whose primaries share a line and whose partners share a
            line (rows r and r+1 of the same face) go to neighbouring lanes;
   chains   the pair that continues the same row in the next element --
            its primary lies in the line of this pair's partner -- comes next,
            with the opposite phase.  A lane of phase 1 loads its pair as
            (partner, primary) instead of (primary, partner), so the partner of
            one unit and the primary of the next, which share a line, are
            fetched by the same instruction.  The values are put back in slot
            order before the reduction, so nothing else changes.

   Purely an ordering of lanes inside a window plus a load schedule: every
   group keeps its members, their order and its window.  Meshes without this
   structure get chains of length one, i.e. the plain primary order.

   chain: 0 switches the reordering off, 1 applies it to pairs only, 2 to every
   width class (EXAGS_SELL_PAIRS, for A/B runs).
   ord[t] = position (k - first) of the t-th group to deal, ph[t] its phase;
   returns the number of groups of the class.  Lines are taken as 16 elements (128 bytes of
   doubles; for 4-byte fields that is half a line, which still shares).  Bits
   8.. of `chain` may carry another line size as a shift (5: 32 elements, the
   line of a 4-byte field; EXAGS_SELL_LINE_SHIFT, an experiment knob): a unit
   then takes up to 2^(shift-3) - 1 twins, which for doubles is the same set of
   lines per instruction (rows r..r+3 are two full lines either way).          */
XG_HD static inline u32 xg_sell_class_order(u32 first, u32 last, const u32 *sg, const u32 *off,
                                            const u32 *idx, u32 wd, unsigned short *ord,
                                            unsigned char *ph, int chain)
{
  unsigned short ck[XG_SELL_WSLOTS / 2];          /* the class in primary order: k - first */
  u32 P[XG_SELL_WSLOTS / 2], Q[XG_SELL_WSLOTS / 2];
  u32 used[XG_SELL_WSLOTS / 64];
  u32 M = 0, ne = 0;
  for (u32 k = first; k < last; ++k) {
    const u32 a = off[sg[k]];
    if (xg_sell_width(off[sg[k] + 1] - a) != wd) continue;
    if (M == XG_SELL_WSLOTS / 2) break;           /* cannot happen: >= 2 slots each */
    /* a one-member group (a flagged single) has no second member to order by */
    ck[M] = (unsigned short)(k - first); P[M] = idx[a]; Q[M] = off[sg[k] + 1] - a > 1 ? idx[a + 1] : idx[a]; ++M;
  }
  const unsigned ls = (chain >> 8) ? (unsigned)(chain >> 8) : 4u;
  const u32 max_twins = (1u << (ls - 3u)) - 1u;
  chain &= 0xff;
  if (!chain || (chain == 1 && wd != 2)) {        /* plain primary order, no phases */
    for (u32 s = 0; s < M; ++s) { ord[s] = ck[s]; ph[s] = 0; }
    return M;
  }
  for (u32 z = 0; z < XG_SELL_WSLOTS / 64; ++z) used[z] = 0;
#define XG_USED(s_) ((used[(s_) >> 5] >> ((s_) & 31u)) & 1u)
#define XG_MARK(s_) (used[(s_) >> 5] |= 1u << ((s_) & 31u))
  for (u32 s = 0; s < M; ++s) {
    if (XG_USED(s)) continue;
    u32 cur = s;
    unsigned char phase = 0;
    for (;;) {
      XG_MARK(cur); ord[ne] = ck[cur]; ph[ne] = phase; ++ne;
      /* twins: the next pairs in primary order that share both lines */
      for (u32 tw = cur + 1; tw < M && tw <= cur + max_twins; ++tw) {
        if (XG_USED(tw) || (P[tw] >> ls) != (P[cur] >> ls) || (Q[tw] >> ls) != (Q[cur] >> ls)) break;
        XG_MARK(tw); ord[ne] = ck[tw]; ph[ne] = phase; ++ne;
      }
      /* successor: first unused pair whose primary lies in the partner's line
         (P ascends: binary search for the start of the line) */
      const u32 line = Q[cur] >> ls;
      u32 lo = 0, hi = M;
      while (lo < hi) { const u32 mid = lo + (hi - lo) / 2; if ((P[mid] >> ls) < line) lo = mid + 1; else hi = mid; }
      while (lo < M && (P[lo] >> ls) == line && XG_USED(lo)) ++lo;
      if (lo < M && (P[lo] >> ls) == line) { cur = lo; phase ^= 1; } else break;
    }
  }
#undef XG_USED
#undef XG_MARK
  return ne;
}

/* off/idx/nunf: CSR of the map the groups come from (nunf may be NULL);
   out_idx, mask, fmask (may be NULL), win_min, win_max: the arrays of struct
   xg_sell, indexed from their start.  Returns 0, or 1 when the groups handed
   in do not fit the window (a bug in the caller's partition).              */
XG_HD static inline int xg_sell_fill_window(size_t w, u32 first, u32 last, const u32 *sg,
                                            const u32 *off, const u32 *idx, const u32 *nunf,
                                            u32 *out_idx, unsigned char *mask, unsigned char *fmask,
                                            u32 *win_min, u32 *win_max, int chain)
{
  unsigned char fill[XG_SELL_LANES];
  u32 *base = out_idx + w * XG_SELL_WSLOTS;
  unsigned char *mk = mask + w * XG_SELL_LANES;
  unsigned char *fk = fmask ? fmask + w * XG_SELL_LANES : 0;
  for (u32 z = 0; z < XG_SELL_WSLOTS; ++z) base[z] = XG_NONE;
  for (u32 z = 0; z < XG_SELL_LANES; ++z) { fill[z] = 0; mk[z] = 0; }
  if (fk) for (u32 z = 0; z < XG_SELL_LANES; ++z) fk[z] = 0;
  {
    u32 lo = XG_NONE, hi = 0;
    for (u32 k = first; k < last; ++k)
      for (u32 q = off[sg[k]]; q < off[sg[k] + 1]; ++q) {
        if (idx[q] < lo) lo = idx[q];
        if (idx[q] > hi) hi = idx[q];
      }
    win_min[w] = lo; win_max[w] = hi;
  }
  /* Groups are dealt, class by class, in an order that lets one load
     instruction of the warp find several of its 32 members in the same
     128-byte line, see xg_sell_class_order().  ordc/phc list the class in
     dealing order with the phase of each group.                             */
  unsigned short ordc[XG_SELL_WSLOTS / 2];
  unsigned char phc[XG_SELL_WSLOTS / 2];
  for (u32 wd = 8; wd >= 2; wd >>= 1) {
    u32 pos = 0;                                  /* lane cursor over the window */
    const u32 count = xg_sell_class_order(first, last, sg, off, idx, wd, ordc, phc, chain);
    for (u32 t = 0; t < count; ++t) {
      const u32 k = first + ordc[t];
      const u32 g = sg[k], a = off[g], sz = off[g + 1] - a;
      if (xg_sell_width(sz) != wd) continue;
      /* next lane (block-major, then lane) with room; fills are multiples of
         wd here, and lanes before the cursor are at least as full */
      u32 tries = 0;
      while (fill[pos] + wd > 8) {
        pos = (pos + 1) % XG_SELL_LANES;
        if (++tries > XG_SELL_LANES) return 1;
      }
      u32 *slot = base + (size_t)pos * 8 + fill[pos];
      for (u32 q = 0; q < sz; ++q) slot[q] = idx[a + q];
      mk[pos] = (unsigned char)(mk[pos] | (1u << fill[pos]));
      /* a group never starts at an odd slot, so the odd bits of the start mask
         are free: bit 2j+1 carries the phase of slots (2j, 2j+1).  Only slot
         pairs that hold two members of the group get one. */
      if (phc[t])
        for (u32 q = 0; q + 1 < sz; q += 2) mk[pos] = (unsigned char)(mk[pos] | (2u << (fill[pos] + q)));
      if (fk) {                                   /* members [nunf, sz) are flagged */
        const u32 unf = nunf[g];
        fk[pos] = (unsigned char)(fk[pos] | ((((1u << sz) - 1u) & ~((1u << unf) - 1u)) << fill[pos]));
      }
      fill[pos] = (unsigned char)(fill[pos] + wd);
      /* row-major dealing: on to the next lane; wrapping starts a new row of
         this width */
      pos = (pos + 1) % XG_SELL_LANES;
    }
  }
  return 0;
}

#endif
