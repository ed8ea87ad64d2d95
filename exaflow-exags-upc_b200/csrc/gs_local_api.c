/* gs_local_api.c -- the public gs_local.h entry points
 * (src/gs_local.h:23-41, bodies src/gs_local.c:187-331) on the GPU.
 *
 * These take the reference's sentinel map format
 *     [dst, src, src, ..., (uint)-1] ... (uint)-1
 * as a host array.  The map is parsed into CSR, uploaded, and the generic
 * CSR gather/scatter kernels run on it; out/in may be host or device
 * pointers (host data is staged through HBM for exactly the index range the
 * map touches).  gs() itself never comes through here -- it uses the CSR
 * maps built once by gs_setup -- so this is compatibility surface, not the
 * hot path.
 */
#include <stdlib.h>
#include <string.h>
#include "exags_internal.h"

struct parsed {
  struct xcu_csr c;
  u32 *d_off, *d_idx;
  u32 max_dst, max_src;   /* largest index used on either side, XG_NONE if none */
};

static void parse_map(struct parsed *p, const exags_uint *map)
{
  memset(p, 0, sizeof *p);
  p->max_dst = p->max_src = XG_NONE;
  size_t ng = 0, ni = 0;
  for (const exags_uint *q = map; *q != XG_NONE;) {
    ++ng;
    while (*q != XG_NONE) { ++q; ++ni; }
    ++q;
  }
  u32 *off = xnew(u32, ng + 1), *idx = xnew(u32, ni ? ni : 1);
  size_t g = 0, o = 0;
  for (const exags_uint *q = map; *q != XG_NONE;) {
    off[g++] = (u32)o;
    if (p->max_dst == XG_NONE || *q > p->max_dst) p->max_dst = *q;
    idx[o++] = *q++;
    while (*q != XG_NONE) {
      if (p->max_src == XG_NONE || *q > p->max_src) p->max_src = *q;
      idx[o++] = *q++;
    }
    ++q;
  }
  off[ng] = (u32)o;
  p->c.ng = (u32)ng; p->c.ni = (u32)ni;
  if (ng) {
    p->d_off = (u32 *)xcu_malloc((ng + 1) * sizeof(u32));
    p->d_idx = (u32 *)xcu_malloc((ni ? ni : 1) * sizeof(u32));
    xcu_h2d(p->d_off, off, (ng + 1) * sizeof(u32), 0);
    xcu_h2d(p->d_idx, idx, ni * sizeof(u32), 0);
    xcu_device_sync();
    p->c.off = p->d_off; p->c.idx = p->d_idx;
  }
  free(off); free(idx);
}

static void parsed_free(struct parsed *p) { xcu_free(p->d_off); xcu_free(p->d_idx); }

/* one side of a kernel call: user pointers + how to stage them */
struct side {
  int mode; unsigned vn; size_t extent;   /* elements per component */
  void *user[XG_MAXV]; void *dev[XG_MAXV]; int host; int is_many;
  void *stage;
};

static void side_open(struct side *s, void *ptr, int mode, unsigned vn, size_t extent,
                      unsigned esz, int copy_in)
{
  memset(s, 0, sizeof *s);
  s->mode = mode; s->vn = vn; s->extent = extent; s->is_many = mode == XM_MANY;
  if (s->is_many) {
    if (vn > XG_MAXV) xfail("gs_local: at most %d arrays per call", XG_MAXV);
    for (unsigned k = 0; k < vn; ++k) s->user[k] = ((void *const *)ptr)[k];
  } else s->user[0] = ptr;
  s->host = !xcu_is_device_ptr(s->user[0]);
  size_t each = s->is_many ? extent * esz : extent * esz * (mode == XM_VEC ? vn : 1);
  unsigned parts = s->is_many ? vn : 1;
  if (!s->host) { for (unsigned k = 0; k < parts; ++k) s->dev[k] = s->user[k]; return; }
  s->stage = xcu_malloc(each * parts);
  for (unsigned k = 0; k < parts; ++k) {
    s->dev[k] = (char *)s->stage + each * k;
    if (copy_in) xcu_h2d(s->dev[k], s->user[k], each, 0);
  }
}

static void side_close(struct side *s, unsigned esz, int copy_out)
{
  if (!s->host) return;
  size_t each = s->is_many ? s->extent * esz : s->extent * esz * (s->mode == XM_VEC ? s->vn : 1);
  unsigned parts = s->is_many ? s->vn : 1;
  if (copy_out) for (unsigned k = 0; k < parts; ++k) xcu_d2h(s->user[k], s->dev[k], each, 0);
  xcu_device_sync();
  xcu_free(s->stage);
}

static void side_field(const struct side *s, struct xcu_field *f)
{
  memset(f, 0, sizeof *f);
  f->vn = s->vn; f->tuple = s->vn; f->k0 = 0;
  for (unsigned k = 0; k < XG_MAXV; ++k) f->p[k] = s->dev[k];
}

static void run_gather(void *out, int omode, const void *in, int imode, unsigned vn,
                       const exags_uint *map, int dom, int op)
{
  int xd = xg_dom_of(dom);
  if (xd < 0) xfail("gs_gather: bad domain %d", dom);
  xg_world_get();
  struct parsed p; parse_map(&p, map);
  if (p.c.ng) {
    unsigned esz = xg_dom_bytes(xd);
    struct side so, si; struct xcu_field fo, fi;
    int same = (out == in && omode == imode);
    size_t eo = (size_t)p.max_dst + 1, ei = p.max_src == XG_NONE ? 0 : (size_t)p.max_src + 1;
    if (same) { if (ei > eo) eo = ei; }
    side_open(&so, out, omode, omode == XM_PLAIN ? 1 : vn, eo, esz, 1);
    if (same) si = so; else side_open(&si, (void *)in, imode, imode == XM_PLAIN ? 1 : vn, ei, esz, 1);
    side_field(&so, &fo); side_field(&si, &fi);
    if (omode == XM_PLAIN) fi.vn = 1;
    fo.vn = fi.vn = (omode == XM_PLAIN && imode == XM_PLAIN) ? 1 : vn;
    fo.tuple = omode == XM_VEC ? vn : 1; fi.tuple = imode == XM_VEC ? vn : 1;
    xcu_csr_gather(0, &p.c, &fo, omode, &fi, imode, xd, op);
    side_close(&so, esz, 1);
    if (!same) side_close(&si, esz, 0);
  }
  parsed_free(&p);
}

static void run_scatter(void *out, int omode, const void *in, int imode, unsigned vn,
                        const exags_uint *map, int dom)
{
  int xd = xg_dom_of(dom);
  if (xd < 0) xfail("gs_scatter: bad domain %d", dom);
  xg_world_get();
  struct parsed p; parse_map(&p, map);
  if (p.c.ng) {
    unsigned esz = xg_dom_bytes(xd);
    struct side so, si; struct xcu_field fo, fi;
    int same = (out == in && omode == imode);
    /* scatter reads in[first], writes out[rest] */
    size_t ei = (size_t)p.max_dst + 1, eo = p.max_src == XG_NONE ? 0 : (size_t)p.max_src + 1;
    if (same) { if (ei > eo) eo = ei; }
    side_open(&so, out, omode, omode == XM_PLAIN ? 1 : vn, eo, esz, 1);
    if (same) si = so; else side_open(&si, (void *)in, imode, imode == XM_PLAIN ? 1 : vn, ei, esz, 1);
    side_field(&so, &fo); side_field(&si, &fi);
    fo.vn = fi.vn = (omode == XM_PLAIN && imode == XM_PLAIN) ? 1 : vn;
    fo.tuple = omode == XM_VEC ? vn : 1; fi.tuple = imode == XM_VEC ? vn : 1;
    xcu_csr_scatter(0, &p.c, &fo, omode, &fi, imode, xd);
    side_close(&so, esz, 1);
    if (!same) side_close(&si, esz, 0);
  }
  parsed_free(&p);
}

static void run_init(void *out, int omode, unsigned vn, const exags_uint *map, int dom, int op)
{
  int xd = xg_dom_of(dom);
  if (xd < 0) xfail("gs_init: bad domain %d", dom);
  xg_world_get();
  size_t cnt = 0; u32 maxi = 0;
  for (const exags_uint *q = map; *q != XG_NONE; ++q) { ++cnt; if (*q > maxi) maxi = *q; }
  if (!cnt) return;
  unsigned esz = xg_dom_bytes(xd);
  u32 *d_list = (u32 *)xcu_malloc(cnt * sizeof(u32));
  xcu_h2d(d_list, map, cnt * sizeof(u32), 0);
  struct side so; struct xcu_field fo;
  side_open(&so, out, omode, omode == XM_PLAIN ? 1 : vn, (size_t)maxi + 1, esz, 1);
  side_field(&so, &fo);
  fo.tuple = omode == XM_VEC ? vn : 1;
  xcu_init_list(0, d_list, (u32)cnt, &fo, omode, xd, op);
  side_close(&so, esz, 1);
  xcu_device_sync();
  xcu_free(d_list);
}

void exags_gs_gather(void *out, const void *in, const unsigned vn, const exags_uint *map, gs_dom dom, gs_op_t op)
{ (void)vn; run_gather(out, XM_PLAIN, in, XM_PLAIN, 1, map, (int)dom, (int)op); }
void exags_gs_scatter(void *out, const void *in, const unsigned vn, const exags_uint *map, gs_dom dom)
{ (void)vn; run_scatter(out, XM_PLAIN, in, XM_PLAIN, 1, map, (int)dom); }
void exags_gs_init(void *out, const unsigned vn, const exags_uint *map, gs_dom dom, gs_op_t op)
{ (void)vn; run_init(out, XM_PLAIN, 1, map, (int)dom, (int)op); }
void exags_gs_gather_vec(void *out, const void *in, const unsigned vn, const exags_uint *map, gs_dom dom, gs_op_t op)
{ run_gather(out, XM_VEC, in, XM_VEC, vn, map, (int)dom, (int)op); }
void exags_gs_scatter_vec(void *out, const void *in, const unsigned vn, const exags_uint *map, gs_dom dom)
{ run_scatter(out, XM_VEC, in, XM_VEC, vn, map, (int)dom); }
void exags_gs_init_vec(void *out, const unsigned vn, const exags_uint *map, gs_dom dom, gs_op_t op)
{ run_init(out, XM_VEC, vn, map, (int)dom, (int)op); }
void exags_gs_gather_many(void *out, const void *in, const unsigned vn, const exags_uint *map, gs_dom dom, gs_op_t op)
{ run_gather(out, XM_MANY, in, XM_MANY, vn, map, (int)dom, (int)op); }
void exags_gs_scatter_many(void *out, const void *in, const unsigned vn, const exags_uint *map, gs_dom dom)
{ run_scatter(out, XM_MANY, in, XM_MANY, vn, map, (int)dom); }
void exags_gs_init_many(void *out, const unsigned vn, const exags_uint *map, gs_dom dom, gs_op_t op)
{ run_init(out, XM_MANY, vn, map, (int)dom, (int)op); }
void exags_gs_gather_vec_to_many(void *out, const void *in, const unsigned vn, const exags_uint *map, gs_dom dom, gs_op_t op)
{ run_gather(out, XM_MANY, in, XM_VEC, vn, map, (int)dom, (int)op); }
void exags_gs_scatter_many_to_vec(void *out, const void *in, const unsigned vn, const exags_uint *map, gs_dom dom)
{ run_scatter(out, XM_VEC, in, XM_MANY, vn, map, (int)dom); }
void exags_gs_scatter_vec_to_many(void *out, const void *in, const unsigned vn, const exags_uint *map, gs_dom dom)
{ run_scatter(out, XM_MANY, in, XM_VEC, vn, map, (int)dom); }

void exags_gs_gather_array(void *out, const void *in, exags_uint n, gs_dom dom, gs_op_t op)
{
  int xd = xg_dom_of((int)dom);
  if (xd < 0) xfail("gs_gather_array: bad domain %d", (int)dom);
  if (!n) return;
  xg_world_get();
  unsigned esz = xg_dom_bytes(xd);
  int ho = !xcu_is_device_ptr(out), hi = !xcu_is_device_ptr(in);
  void *dout = out; const void *din = in;
  if (ho) { dout = xcu_malloc((size_t)n * esz); xcu_h2d(dout, out, (size_t)n * esz, 0); }
  if (hi) { void *t = xcu_malloc((size_t)n * esz); xcu_h2d(t, in, (size_t)n * esz, 0); din = t; }
  xcu_array_op(0, dout, din, n, xd, (int)op);
  if (ho) xcu_d2h(out, dout, (size_t)n * esz, 0);
  xcu_device_sync();
  if (ho) xcu_free(dout);
  if (hi) xcu_free((void *)din);
}

void exags_gs_init_array(void *out, exags_uint n, gs_dom dom, gs_op_t op)
{
  int xd = xg_dom_of((int)dom);
  if (xd < 0) xfail("gs_init_array: bad domain %d", (int)dom);
  if (!n) return;
  xg_world_get();
  unsigned esz = xg_dom_bytes(xd);
  int ho = !xcu_is_device_ptr(out);
  void *dout = ho ? xcu_malloc((size_t)n * esz) : out;
  xcu_fill_identity(0, dout, n, xd, (int)op);
  if (ho) xcu_d2h(out, dout, (size_t)n * esz, 0);
  xcu_device_sync();
  if (ho) xcu_free(dout);
}
