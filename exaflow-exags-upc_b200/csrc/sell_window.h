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

/* off/idx/nunf: CSR of the map the groups come from (nunf may be NULL);
   out_idx, mask, fmask (may be NULL), win_min, win_max: the arrays of struct
   xg_sell, indexed from their start.  Returns 0, or 1 when the groups handed
   in do not fit the window (a bug in the caller's partition).              */
XG_HD static inline int xg_sell_fill_window(size_t w, u32 first, u32 last, const u32 *sg,
                                            const u32 *off, const u32 *idx, const u32 *nunf,
                                            u32 *out_idx, unsigned char *mask, unsigned char *fmask,
                                            u32 *win_min, u32 *win_max)
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
  for (u32 wd = 8; wd >= 2; wd >>= 1) {
    u32 pos = 0;                                  /* lane cursor over the window */
    for (u32 k = first; k < last; ++k) {
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
