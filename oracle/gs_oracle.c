/*
 * gs_oracle.c -- CPU restatement of the reference gs_op path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under exaflow-exags-upc_b200/ may link,
 * import or call this file; it exists so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg can check (and time beside) the CUDA path.
 *
 * It follows the reference algorithm step by step, in the reference's own
 * data formats (struct rows + sentinel-terminated uint maps), with all np
 * ranks of a job simulated inside one process:
 *
 *   nonzero_ids              src/gs.upc:87-112
 *   unique_ids / shared_ids  src/gs.upc:117-244   (rendezvous on id % np)
 *   make_topology_unique     src/gs.upc:253-322
 *   local_map                src/gs.upc:329-357
 *   flagged_primaries_map    src/gs.upc:359-370
 *   pw_comm_setup/map_setup  src/gs.upc:505-570
 *   allreduce_map_setup      src/gs.upc:1031-1050
 *   gather/scatter/init      src/gs_local.c:60-158,256-331
 *   GS_DO_* and identities   src/gs_defs.h:33-52
 *   gs_aux / pw_exec / allreduce_exec   src/gs.upc:470-500,1006-1026,1169-1185
 *
 * Parity pin: tests/test_oracle_pin.py checks this file against (a) the
 * known answers the survey obtained by running the reference (SURVEY.md 3.4,
 * 8c), (b) the reference's own tests/gs/gs_test.upc assertions, and (c) when
 * /root/reference is present, the reference sources themselves compiled by
 * oracle/build_ref.sh into oracle/_ref/.
 */
#include <float.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned int uint;
typedef unsigned long long ulong64;
typedef long long slong64;
#define NIL ((uint)-1)

enum { D_DOUBLE, D_FLOAT, D_INT, D_LONG };
enum { O_ADD, O_MUL, O_MIN, O_MAX, O_BPR };
enum { M_PLAIN, M_VEC, M_MANY };
enum { METHOD_AUTO, METHOD_PAIRWISE, METHOD_CRYSTAL, METHOD_ALLREDUCE };
#define FLAGS_LOCAL 1u
#define FLAGS_REMOTE 2u

static const unsigned dom_size[4] = { sizeof(double), sizeof(float), sizeof(int), sizeof(long long) };

static void *zalloc(size_t n) { void *p = calloc(n ? n : 1, 1); if (!p) { fprintf(stderr, "oracle: out of memory\n"); exit(1); } return p; }

/* ---- rows ---------------------------------------------------------------- */
struct nz_row { ulong64 id; uint i, flag, primary; };                       /* gs.upc:83-85 */
struct sh_row { ulong64 id; uint i, p, ri, bi; unsigned flags; ulong64 ord; }; /* gs.upc:152-154 */
struct pr_row { ulong64 id, ord; uint i; unsigned flag; };                  /* gs.upc:156-158 */
struct un_row { ulong64 id; uint proc, src_if; };                           /* gs.upc:117 */

struct pw_comm { uint n, *p, *size, total; };                               /* gs.upc:395-401 */

struct rank_data {
  uint n;
  struct nz_row *nz; size_t nnz;
  struct sh_row *sh; size_t nsh;
  struct pr_row *pr; size_t npr;
  uint *map_local[2], *flagged_primaries;       /* gs.upc:1161-1167 */
  size_t len_local[2], len_fp;
  struct pw_comm comm[2]; uint *pw_map[2]; size_t len_pw[2];   /* gs.upc:403-407 */
  uint *ar_to[2], *ar_from[2];                  /* gs.upc:1001-1004 */
  char *buf; size_t buf_cap;
};

typedef struct ogs {
  int np, method;
  ulong64 total_shared;
  struct rank_data *r;
} ogs;

/* ---- stable sorting: qsort with the original position as last key -------- */
#define DEF_CMP(name, T, expr_lt, expr_gt) \
  static int name(const void *a_, const void *b_) { \
    const T *a = *(T *const *)a_, *b = *(T *const *)b_; \
    if (expr_lt) return -1; \
    if (expr_gt) return 1; \
    return a < b ? -1 : a > b; }

static void stable_sort(void *base, size_t n, size_t size, int (*cmp)(const void *, const void *))
{
  if (n < 2) return;
  char **ptr = zalloc(n * sizeof *ptr), *tmp = zalloc(n * size);
  for (size_t k = 0; k < n; ++k) ptr[k] = (char *)base + k * size;
  qsort(ptr, n, sizeof *ptr, cmp);
  for (size_t k = 0; k < n; ++k) memcpy(tmp + k * size, ptr[k], size);
  memcpy(base, tmp, n * size);
  free(ptr); free(tmp);
}

DEF_CMP(nz_by_id_flag, struct nz_row, a->id < b->id || (a->id == b->id && a->flag < b->flag),
        a->id > b->id || (a->id == b->id && a->flag > b->flag))
DEF_CMP(nz_by_primary, struct nz_row, a->primary < b->primary, a->primary > b->primary)
DEF_CMP(nz_by_i, struct nz_row, a->i < b->i, a->i > b->i)
DEF_CMP(un_by_id_srcif, struct un_row, a->id < b->id || (a->id == b->id && a->src_if < b->src_if),
        a->id > b->id || (a->id == b->id && a->src_if > b->src_if))
DEF_CMP(sh_by_id, struct sh_row, a->id < b->id, a->id > b->id)
DEF_CMP(sh_by_i, struct sh_row, a->i < b->i, a->i > b->i)
DEF_CMP(sh_by_i_p, struct sh_row, a->i < b->i || (a->i == b->i && a->p < b->p),
        a->i > b->i || (a->i == b->i && a->p > b->p))
DEF_CMP(sh_by_p_id, struct sh_row, a->p < b->p || (a->p == b->p && a->id < b->id),
        a->p > b->p || (a->p == b->p && a->id > b->id))
DEF_CMP(pr_by_i, struct pr_row, a->i < b->i, a->i > b->i)
DEF_CMP(pr_by_id, struct pr_row, a->id < b->id, a->id > b->id)

/* ---- nonzero_ids (gs.upc:87-112) ----------------------------------------- */
static void nonzero_ids(struct rank_data *r, const slong64 *id, uint n)
{
  struct nz_row *nz = zalloc((size_t)n * sizeof *nz);
  size_t m = 0;
  for (uint i = 0; i < n; ++i) {
    slong64 v = id[i], a = v < 0 ? -v : v;
    if (v == 0) continue;
    nz[m].i = i; nz[m].id = (ulong64)a; nz[m].flag = v != a; ++m;
  }
  stable_sort(nz, m, sizeof *nz, nz_by_id_flag);
  ulong64 last = (ulong64)-1; uint primary = NIL;
  for (size_t k = 0; k < m; ++k) {
    if (nz[k].id != last) primary = nz[k].i;
    nz[k].primary = primary; last = nz[k].id;
  }
  stable_sort(nz, m, sizeof *nz, nz_by_primary);
  r->nz = nz; r->nnz = m; r->n = n;
}

/* ---- shared_ids (gs.upc:190-244) with all ranks in one address space ------ */
static void shared_ids(ogs *h)
{
  int np = h->np;
  size_t total_un = 0;
  for (int p = 0; p < np; ++p)
    for (size_t k = 0; k < h->r[p].nnz; ++k) total_un += h->r[p].nz[k].i == h->r[p].nz[k].primary;
  ulong64 ordinal = 0;
  size_t *cap = zalloc(np * sizeof *cap);
  /* work rank w handles ids with id % np == w, in rank order (comm_scan) */
  for (int w = 0; w < np; ++w) {
    struct un_row *un = zalloc(total_un * sizeof *un);
    size_t m = 0;
    for (int p = 0; p < np; ++p)      /* arrival order: by source rank */
      for (size_t k = 0; k < h->r[p].nnz; ++k) {
        const struct nz_row *z = &h->r[p].nz[k];
        if (z->i != z->primary || z->id % (ulong64)np != (ulong64)w) continue;
        un[m].id = z->id; un[m].proc = (uint)p; un[m].src_if = z->flag ? ~z->i : z->i; ++m;
      }
    stable_sort(un, m, sizeof *un, un_by_id_srcif);
    ulong64 last = (ulong64)-1;
    for (size_t a = 0; a < m; ++a) {
      uint i1f = un[a].src_if;
      if (~i1f < i1f) continue;                     /* flagged: never first of a pair */
      for (size_t b = a + 1; b < m && un[b].id == un[a].id; ++b) {
        if (un[a].id != last) { last = un[a].id; ++ordinal; }
        ulong64 ord = ordinal - 1;
        uint i2f = un[b].src_if;
        for (int dir = 0; dir < 2; ++dir) {
          uint p1 = dir ? un[b].proc : un[a].proc, p2 = dir ? un[a].proc : un[b].proc;
          uint s1 = dir ? i2f : i1f, s2 = dir ? i1f : i2f;
          struct rank_data *r = &h->r[p1];
          if (r->nsh == cap[p1]) { cap[p1] = cap[p1] * 2 + 16; r->sh = realloc(r->sh, cap[p1] * sizeof *r->sh); }
          struct sh_row *s = &r->sh[r->nsh++];
          uint i1 = ~s1 < s1 ? ~s1 : s1, i2 = ~s2 < s2 ? ~s2 : s2;   /* gs.upc:176 */
          s->id = un[a].id; s->i = i1; s->p = p2; s->ri = i2; s->bi = NIL; s->ord = ord;
          s->flags = ((s2 ^ i2) & FLAGS_REMOTE) | ((s1 ^ i1) & FLAGS_LOCAL);
        }
      }
    }
    free(un);
  }
  h->total_shared = ordinal;
  free(cap);
  /* shared_ids_aux (gs.upc:163-188): primary_shared rows, one per id, by i */
  for (int p = 0; p < np; ++p) {
    struct rank_data *r = &h->r[p];
    stable_sort(r->sh, r->nsh, sizeof *r->sh, sh_by_id);
    r->pr = zalloc(r->nsh * sizeof *r->pr); r->npr = 0;
    ulong64 last = (ulong64)-1;
    for (size_t k = 0; k < r->nsh; ++k)
      if (r->sh[k].id != last) {
        last = r->sh[k].id;
        struct pr_row *q = &r->pr[r->npr++];
        q->id = last; q->ord = r->sh[k].ord; q->i = r->sh[k].i; q->flag = r->sh[k].flags & FLAGS_LOCAL;
      }
    stable_sort(r->pr, r->npr, sizeof *r->pr, pr_by_i);
  }
}

/* ---- make_topology_unique (gs.upc:253-322) -------------------------------- */
static void make_unique(struct rank_data *r, slong64 *id, uint pid)
{
  stable_sort(r->nz, r->nnz, sizeof *r->nz, nz_by_i);
  for (size_t k = 0; k < r->nnz; ++k)
    if (r->nz[k].i != r->nz[k].primary) { if (id) id[r->nz[k].i] = -(slong64)r->nz[k].id; r->nz[k].flag = 1; }
  stable_sort(r->nz, r->nnz, sizeof *r->nz, nz_by_primary);

  stable_sort(r->sh, r->nsh, sizeof *r->sh, sh_by_i_p);
  struct sh_row *keep = zalloc((r->nsh + 1) * sizeof *keep); size_t nk = 0;
  size_t z = 0;
  for (size_t a = 0; a < r->nsh;) {
    size_t b = a; uint i = r->sh[a].i, lt = 0, gt = 0, owner; int mine = 0;
    while (r->nz[z].i != i) ++z;     /* nz is ordered by primary; heads are primaries */
    for (; b < r->nsh && r->sh[b].i == i && r->sh[b].p < pid; ++b) if (!(r->sh[b].flags & FLAGS_REMOTE)) ++lt;
    for (; b < r->nsh && r->sh[b].i == i; ++b) if (!(r->sh[b].flags & FLAGS_REMOTE)) ++gt;
    if (!(r->sh[a].flags & FLAGS_LOCAL)) {
      owner = (uint)(r->sh[a].id % (lt + gt + 1));
      if (owner == lt) mine = 1; else if (owner > lt) --owner;
    } else owner = (uint)(r->sh[a].id % (lt + gt));
    if (!mine) {
      if (id) id[i] = -(slong64)r->sh[a].id;
      r->nz[z].flag = 1;
      for (size_t k = a; k < b; ++k)
        if (!(r->sh[k].flags & FLAGS_REMOTE) && !(owner--)) { keep[nk] = r->sh[k]; keep[nk].flags = FLAGS_LOCAL; ++nk; break; }
    } else
      for (size_t k = a; k < b; ++k) { keep[nk] = r->sh[k]; keep[nk].flags = FLAGS_REMOTE; ++nk; }
    a = b;
  }
  free(r->sh); r->sh = keep; r->nsh = nk;
  /* primary flags follow (gs.upc:310-321) */
  stable_sort(r->sh, r->nsh, sizeof *r->sh, sh_by_id);
  stable_sort(r->pr, r->npr, sizeof *r->pr, pr_by_id);
  size_t q = 0;
  for (size_t a = 0; a < r->nsh;) {
    size_t b = a; while (b < r->nsh && r->sh[b].i == r->sh[a].i) ++b;
    while (q < r->npr && r->pr[q].id != r->sh[a].id) ++q;    /* ids that lost all rows drop out */
    r->pr[q].flag = r->sh[a].flags & FLAGS_LOCAL;
    ++q;
    a = b;
  }
  /* drop primary rows whose id has no shared row left */
  {
    size_t w = 0, s = 0;
    for (size_t k = 0; k < r->npr; ++k) {
      while (s < r->nsh && r->sh[s].id < r->pr[k].id) ++s;
      if (s < r->nsh && r->sh[s].id == r->pr[k].id) r->pr[w++] = r->pr[k];
    }
    r->npr = w;
  }
  stable_sort(r->pr, r->npr, sizeof *r->pr, pr_by_i);
}

/* ---- local_map / flagged_primaries_map (gs.upc:329-370) ------------------- */
static uint *local_map(const struct rank_data *r, int ignore_flagged, size_t *len)
{
  const struct nz_row *nz = r->nz; size_t m = r->nnz;
  uint *map = zalloc((2 * m + 2) * sizeof *map), *p = map;
  for (size_t a = 0; a < m;) {
    size_t b = a + 1; int any = 0;
    *p++ = nz[a].i;
    for (; b < m && nz[b].id == nz[a].id && (!ignore_flagged || nz[b].flag == 0); ++b) { any = 1; *p++ = nz[b].i; }
    if (any) *p++ = NIL; else --p;
    a = b;
  }
  *p++ = NIL;
  *len = (size_t)(p - map);
  return map;
}

static uint *flagged_primaries_map(const struct rank_data *r, size_t *len)
{
  uint *map = zalloc((r->nnz + 1) * sizeof *map), *p = map;
  for (size_t k = 0; k < r->nnz; ++k)
    if (r->nz[k].i == r->nz[k].primary && r->nz[k].flag == 1) *p++ = r->nz[k].i;
  *p++ = NIL;
  *len = (size_t)(p - map);
  return map;
}

/* ---- pairwise setup (gs.upc:505-570) -------------------------------------- */
static void pw_comm_setup(struct pw_comm *c, struct rank_data *r, unsigned mask)
{
  uint n = 0, count = 0, lp = NIL;
  stable_sort(r->sh, r->nsh, sizeof *r->sh, sh_by_p_id);
  for (size_t k = 0; k < r->nsh; ++k) {
    struct sh_row *s = &r->sh[k];
    if (s->flags & mask) { s->bi = NIL; continue; }
    s->bi = count++;
    if (s->p != lp) { lp = s->p; ++n; }
  }
  c->n = n; c->total = count;
  c->p = zalloc(2 * (size_t)n * sizeof *c->p); c->size = c->p + n;
  n = 0; lp = NIL; count = 0;
  for (size_t k = 0; k < r->nsh; ++k) {
    struct sh_row *s = &r->sh[k];
    if (s->flags & mask) continue;
    if (s->p != lp) { lp = s->p; if (n) c->size[n - 1] = count; count = 0; c->p[n++] = lp; }
    ++count;
  }
  if (n) c->size[n - 1] = count;
}

static uint *pw_map_setup(struct rank_data *r, size_t *len)
{
  stable_sort(r->sh, r->nsh, sizeof *r->sh, sh_by_i);
  uint *map = zalloc((3 * r->nsh + 1) * sizeof *map), *p = map;
  for (size_t k = 0; k < r->nsh;) {
    uint i = r->sh[k].i;
    if (r->sh[k].bi == NIL) { ++k; continue; }
    *p++ = i; *p++ = r->sh[k].bi;
    for (++k; k < r->nsh && r->sh[k].i == i; ++k) if (r->sh[k].bi != NIL) *p++ = r->sh[k].bi;
    *p++ = NIL;
  }
  *p++ = NIL;
  *len = (size_t)(p - map);
  return map;
}

/* ---- all-reduce setup (gs.upc:1031-1050) ---------------------------------- */
static uint *allreduce_map(const struct rank_data *r, unsigned mask, int to_buf)
{
  uint *map = zalloc((3 * r->npr + 1) * sizeof *map), *m = map;
  for (size_t k = 0; k < r->npr; ++k)
    if ((r->pr[k].flag & mask) == 0) {
      if (to_buf) { *m++ = r->pr[k].i; *m++ = (uint)r->pr[k].ord; }
      else        { *m++ = (uint)r->pr[k].ord; *m++ = r->pr[k].i; }
      *m++ = NIL;
    }
  *m = NIL;
  return map;
}

/* ---- kernels (gs_local.c, gs_defs.h) -------------------------------------- */
#define DO_add(a, b) a += b
#define DO_mul(a, b) a *= b
#define DO_min(a, b) if (b < a) a = b
#define DO_max(a, b) if (b > a) a = b
#define DO_bpr(a, b) \
  do if (b != 0) { uint a_ = (uint)a; uint b_ = (uint)b; \
       if (a_ == 0) { a = b_; break; } \
       for (;;) { if (a_ < b_) b_ >>= 1; else if (b_ < a_) a_ >>= 1; else break; } \
       a = a_; \
     } while (0)

#define FOR_T(M) M(double, double) M(float, float) M(int, int) M(long long, llong)
#define FOR_OP(T, N, M) M(T, N, add) M(T, N, mul) M(T, N, min) M(T, N, max) M(T, N, bpr)

#define DEF_IDENT(T, N, lo, hi) static const T ident_##N[5] = { 0, 1, hi, lo, 0 };
DEF_IDENT(double, double, -DBL_MAX, DBL_MAX)
DEF_IDENT(float, float, -FLT_MAX, FLT_MAX)
DEF_IDENT(int, int, INT_MIN, INT_MAX)
DEF_IDENT(long long, llong, LLONG_MIN, LLONG_MAX)

/* gather with strides: out[i*os] OP= in[j*is]   (gs_local.c:60-71, 295-306) */
#define DEF_GATHER(T, N, OP) \
  static void gather_##N##_##OP(T *out, unsigned os, const T *in, unsigned is, const uint *map) { \
    uint i, j; \
    while ((i = *map++) != NIL) { \
      T t = out[(size_t)i * os]; \
      j = *map++; do { T q = in[(size_t)j * is]; DO_##OP(t, q); } while ((j = *map++) != NIL); \
      out[(size_t)i * os] = t; } }
#define DEF_GATHERS(T, N) FOR_OP(T, N, DEF_GATHER)
FOR_T(DEF_GATHERS)

/* scatter with strides (gs_local.c:76-87) */
#define DEF_SCATTER(T, N) \
  static void scatter_##N(T *out, unsigned os, const T *in, unsigned is, const uint *map) { \
    uint i, j; \
    while ((i = *map++) != NIL) { T t = in[(size_t)i * is]; j = *map++; do out[(size_t)j * os] = t; while ((j = *map++) != NIL); } }
FOR_T(DEF_SCATTER)

#define DEF_INIT(T, N) \
  static void init_##N(T *out, unsigned os, const uint *map, int op) { \
    uint i; const T e = ident_##N[op]; while ((i = *map++) != NIL) out[(size_t)i * os] = e; }
FOR_T(DEF_INIT)

#define CASE_OP(T, N, OP) case O_##OP: gather_##N##_##OP((T *)out, os, (const T *)in, is, map); break;
#define O_add O_ADD
#define O_mul O_MUL
#define O_min O_MIN
#define O_max O_MAX
#define O_bpr O_BPR
static void gather_any(void *out, unsigned os, const void *in, unsigned is, const uint *map, int dom, int op)
{
  switch (dom) {
    case D_DOUBLE: switch (op) { FOR_OP(double, double, CASE_OP) } break;
    case D_FLOAT:  switch (op) { FOR_OP(float, float, CASE_OP) } break;
    case D_INT:    switch (op) { FOR_OP(int, int, CASE_OP) } break;
    case D_LONG:   switch (op) { FOR_OP(long long, llong, CASE_OP) } break;
  }
}
static void scatter_any(void *out, unsigned os, const void *in, unsigned is, const uint *map, int dom)
{
  switch (dom) {
    case D_DOUBLE: scatter_double(out, os, in, is, map); break;
    case D_FLOAT:  scatter_float(out, os, in, is, map); break;
    case D_INT:    scatter_int(out, os, in, is, map); break;
    case D_LONG:   scatter_llong(out, os, in, is, map); break;
  }
}
static void init_any(void *out, unsigned os, const uint *map, int dom, int op)
{
  switch (dom) {
    case D_DOUBLE: init_double(out, os, map, op); break;
    case D_FLOAT:  init_float(out, os, map, op); break;
    case D_INT:    init_int(out, os, map, op); break;
    case D_LONG:   init_llong(out, os, map, op); break;
  }
}

/* A user field seen component by component: component k of element i lives
   at base_k + i*stride elements.  plain: 1 comp; vec: base+k, stride vn
   (gs_local.c:114-158 loop over the tuple); many: separate arrays
   (gs_local.c:256-288 loop over k). */
struct view { char *base[64]; unsigned stride, vn; };
static void make_view(struct view *v, void *u, int mode, unsigned vn, int dom)
{
  v->vn = mode == M_PLAIN ? 1 : vn;
  if (v->vn > 64) { fprintf(stderr, "oracle: vn too large\n"); exit(1); }
  for (unsigned k = 0; k < v->vn; ++k)
    v->base[k] = mode == M_MANY ? (char *)((void **)u)[k] : (char *)u + (size_t)k * dom_size[dom];
  v->stride = mode == M_VEC ? vn : 1;
}

/* ---- setup ---------------------------------------------------------------- */
ogs *ogs_setup(int np, const long long *const *ids, const unsigned *n, int unique, int method)
{
  ogs *h = zalloc(sizeof *h);
  h->np = np; h->method = method;
  h->r = zalloc((size_t)np * sizeof *h->r);
  for (int p = 0; p < np; ++p) nonzero_ids(&h->r[p], ids[p], n[p]);
  shared_ids(h);
  if (unique) for (int p = 0; p < np; ++p) make_unique(&h->r[p], 0, (uint)p);
  for (int p = 0; p < np; ++p) {
    struct rank_data *r = &h->r[p];
    r->map_local[0] = local_map(r, 1, &r->len_local[0]);
    r->map_local[1] = local_map(r, 0, &r->len_local[1]);
    r->flagged_primaries = flagged_primaries_map(r, &r->len_fp);
    /* pw_setup_aux (gs.upc:572-592) */
    pw_comm_setup(&r->comm[0], r, FLAGS_REMOTE); r->pw_map[0] = pw_map_setup(r, &r->len_pw[0]);
    pw_comm_setup(&r->comm[1], r, FLAGS_LOCAL);  r->pw_map[1] = pw_map_setup(r, &r->len_pw[1]);
    /* allreduce_setup_aux (gs.upc:1052-1068) */
    r->ar_to[0] = allreduce_map(r, 1, 1); r->ar_from[0] = allreduce_map(r, 0, 0);
    r->ar_to[1] = allreduce_map(r, 0, 1); r->ar_from[1] = allreduce_map(r, 1, 0);
  }
  return h;
}

void ogs_free(ogs *h)
{
  if (!h) return;
  for (int p = 0; p < h->np; ++p) {
    struct rank_data *r = &h->r[p];
    free(r->nz); free(r->sh); free(r->pr); free(r->map_local[0]); free(r->map_local[1]);
    free(r->flagged_primaries); free(r->comm[0].p); free(r->comm[1].p); free(r->pw_map[0]); free(r->pw_map[1]);
    for (int t = 0; t < 2; ++t) { free(r->ar_to[t]); free(r->ar_from[t]); }
    free(r->buf);
  }
  free(h->r); free(h);
}

unsigned long long ogs_total_shared(const ogs *h) { return h->total_shared; }

/* which: 0 map_local[0], 1 map_local[1], 2 flagged_primaries, 3 pw map[0] (recv), 4 pw map[1] (send) */
const unsigned *ogs_map(const ogs *h, int rank, int which, size_t *len)
{
  const struct rank_data *r = &h->r[rank];
  switch (which) {
    case 0: case 1: *len = r->len_local[which]; return r->map_local[which];
    case 2: *len = r->len_fp; return r->flagged_primaries;
    case 3: case 4: *len = r->len_pw[which - 3]; return r->pw_map[which - 3];
  }
  *len = 0; return 0;
}

/* neighbour table of direction d: fills p[] and size[] (cap entries), returns count */
unsigned ogs_neighbours(const ogs *h, int rank, int d, unsigned *p, unsigned *size, unsigned cap)
{
  const struct pw_comm *c = &h->r[rank].comm[d];
  for (unsigned k = 0; k < c->n && k < cap; ++k) { p[k] = c->p[k]; size[k] = c->size[k]; }
  return c->n;
}

void ogs_unique(int np, long long *const *ids, const unsigned *n)
{
  ogs *h = zalloc(sizeof *h);
  h->np = np; h->r = zalloc((size_t)np * sizeof *h->r);
  for (int p = 0; p < np; ++p) nonzero_ids(&h->r[p], ids[p], n[p]);
  shared_ids(h);
  for (int p = 0; p < np; ++p) make_unique(&h->r[p], ids[p], (uint)p);
  for (int p = 0; p < np; ++p) { free(h->r[p].nz); free(h->r[p].sh); free(h->r[p].pr); }
  free(h->r); free(h);
}

/* ---- execution (gs_aux, gs.upc:1169-1185) --------------------------------- */
static void reserve(struct rank_data *r, size_t bytes)
{
  if (r->buf_cap < bytes) { r->buf = realloc(r->buf, bytes); r->buf_cap = bytes; }
}

void ogs_exec(ogs *h, void *const *u, int mode, unsigned vn, int dom, int op, int transpose)
{
  const int np = h->np, t = transpose != 0;
  const unsigned esz = dom_size[dom], nv = mode == M_PLAIN ? 1 : vn;
  struct view *v = zalloc((size_t)np * sizeof *v);
  for (int p = 0; p < np; ++p) make_view(&v[p], u[p], mode, vn, dom);

  /* local gather + init (gs.upc:1181-1182).  Ranks are the reference's SPMD
     threads: each phase runs them side by side on the host cores (OpenMP),
     with the implicit barrier between phases standing in for the put/flag
     handshake of pw_exec_sends/recvs. */
#pragma omp parallel for schedule(static)
  for (int p = 0; p < np; ++p) {
    struct rank_data *r = &h->r[p];
    for (unsigned k = 0; k < nv; ++k) {
      gather_any(v[p].base[k], v[p].stride, v[p].base[k], v[p].stride, r->map_local[0 ^ t], dom, op);
      if (!t) init_any(v[p].base[k], v[p].stride, r->flagged_primaries, dom, op);
    }
  }

  if (np > 1 && h->method != METHOD_ALLREDUCE) {
    /* pw_exec (gs.upc:470-500): pack, deliver, unpack */
    const int recv = 0 ^ t, send = 1 ^ t;
    char **sendbuf = zalloc((size_t)np * sizeof *sendbuf);
#pragma omp parallel for schedule(static)
    for (int p = 0; p < np; ++p) {
      struct rank_data *r = &h->r[p];
      sendbuf[p] = zalloc((size_t)r->comm[send].total * nv * esz);
      for (unsigned k = 0; k < nv; ++k)      /* buffer layout [slot][vn] (gs_local.c:308-318) */
        scatter_any(sendbuf[p] + (size_t)k * esz, nv, v[p].base[k], v[p].stride, r->pw_map[send], dom);
    }
#pragma omp parallel for schedule(static)
    for (int p = 0; p < np; ++p) {
      struct rank_data *r = &h->r[p];
      reserve(r, (size_t)r->comm[recv].total * nv * esz + 1);
      size_t acc = 0;
      for (uint m = 0; m < r->comm[recv].n; ++m) {
        uint q = r->comm[recv].p[m]; size_t len = (size_t)r->comm[recv].size[m] * nv * esz;
        /* find my segment in q's send buffer */
        const struct pw_comm *cs = &h->r[q].comm[send];
        size_t off = 0; uint s = 0;
        for (; s < cs->n && cs->p[s] != (uint)p; ++s) off += (size_t)cs->size[s] * nv * esz;
        if (s == cs->n || (size_t)cs->size[s] * nv * esz != len) { fprintf(stderr, "oracle: pairwise message mismatch %d<-%u\n", p, q); exit(1); }
        memcpy(r->buf + acc, sendbuf[q] + off, len);
        acc += len;
      }
    }
#pragma omp parallel for schedule(static)
    for (int p = 0; p < np; ++p) {
      struct rank_data *r = &h->r[p];
      for (unsigned k = 0; k < nv; ++k)
        gather_any(v[p].base[k], v[p].stride, r->buf + (size_t)k * esz, nv, r->pw_map[recv], dom, op);
    }
    for (int p = 0; p < np; ++p) free(sendbuf[p]);
    free(sendbuf);
  } else if (np > 1) {
    /* allreduce_exec (gs.upc:1006-1026); combine in rank order */
    size_t gvn = (size_t)h->total_shared * nv;
    char *acc = zalloc(gvn * esz + 1), *mine = zalloc(gvn * esz + 1);
    uint *all = zalloc((gvn + 1) * sizeof *all);
    for (size_t k = 0; k < gvn; ++k) all[k] = (uint)k;
    all[gvn] = NIL;
    for (int p = 0; p < np; ++p) {
      struct rank_data *r = &h->r[p];
      init_any(mine, 1, all, dom, op);
      for (unsigned k = 0; k < nv; ++k)
        scatter_any(mine + (size_t)k * esz, nv, v[p].base[k], v[p].stride, r->ar_to[t], dom);
      if (p == 0) memcpy(acc, mine, gvn * esz);
      else {
        /* element-wise acc OP= mine (gs_gather_array, gs_local.c:187-194) */
        uint pair[4] = { 0, 0, NIL, NIL };
        for (size_t k = 0; k < gvn; ++k) gather_any(acc + k * esz, 1, mine + k * esz, 1, pair, dom, op);
      }
    }
    for (int p = 0; p < np; ++p) {
      struct rank_data *r = &h->r[p];
      for (unsigned k = 0; k < nv; ++k)
        scatter_any(v[p].base[k], v[p].stride, acc + (size_t)k * esz, nv, r->ar_from[t], dom);
    }
    free(acc); free(mine); free(all);
  }

  /* local scatter (gs.upc:1184) */
#pragma omp parallel for schedule(static)
  for (int p = 0; p < np; ++p) {
    struct rank_data *r = &h->r[p];
    for (unsigned k = 0; k < nv; ++k)
      scatter_any(v[p].base[k], v[p].stride, v[p].base[k], v[p].stride, r->map_local[1 ^ t], dom);
  }
  free(v);
}

/* single-rank convenience used by the bench's cpu_baseline leg */
void ogs_exec1(ogs *h, void *u, int mode, unsigned vn, int dom, int op, int transpose)
{
  void *uu[1] = { u };
  ogs_exec(h, uu, mode, vn, dom, op, transpose);
}

/* ---- the public gs_local.h entry points (src/gs_local.h:23-41) -------------
   One dispatcher for the fourteen of them, restated on the strided kernels
   above exactly as src/gs_local.c composes them:
     gs_gather_array / gs_init_array    src/gs_local.c:187-201 (n in `vn`, no map)
     gs_gather / gs_scatter / gs_init   :206-230  (vn ignored)
     *_vec                              :114-158,235-251  (tuples of vn)
     *_many                             :256-288  (vn arrays, plain kernel on each)
     gs_gather_vec_to_many              :295-307  out = vn arrays, in = [slot][vn]
     gs_scatter_many_to_vec             :309-319  out = [slot][vn], in = vn arrays
     gs_scatter_vec_to_many             :321-331  out = vn arrays, in = [slot][vn]
   tests/test_oracle_pin.py pins it to the reference's compiled gs_local.c.     */
enum { L_GATHER_ARRAY, L_INIT_ARRAY, L_GATHER, L_SCATTER, L_INIT, L_GATHER_VEC, L_SCATTER_VEC, L_INIT_VEC,
       L_GATHER_MANY, L_SCATTER_MANY, L_INIT_MANY, L_GATHER_VEC_TO_MANY, L_SCATTER_MANY_TO_VEC,
       L_SCATTER_VEC_TO_MANY };
void ogs_local(int which, void *out, const void *in, unsigned vn, const unsigned *map, int dom, int op)
{
  const unsigned sz = dom_size[dom];
  void *const *po = (void *const *)out;
  const void *const *pi = (const void *const *)in;
  switch (which) {
    case L_GATHER_ARRAY: {            /* dense: element i of out OP= element i of in */
      uint *idm = (uint *)zalloc(((size_t)3 * vn + 1) * sizeof(uint));
      for (uint i = 0; i < vn; ++i) { idm[3 * i] = i; idm[3 * i + 1] = i; idm[3 * i + 2] = NIL; }
      idm[3 * (size_t)vn] = NIL;
      gather_any(out, 1, in, 1, idm, dom, op);
      free(idm);
      break;
    }
    case L_INIT_ARRAY: {
      uint *idm = (uint *)zalloc(((size_t)vn + 1) * sizeof(uint));
      for (uint i = 0; i < vn; ++i) idm[i] = i;
      idm[vn] = NIL;
      init_any(out, 1, idm, dom, op);
      free(idm);
      break;
    }
    case L_GATHER:  gather_any(out, 1, in, 1, map, dom, op); break;
    case L_SCATTER: scatter_any(out, 1, in, 1, map, dom); break;
    case L_INIT:    init_any(out, 1, map, dom, op); break;
    case L_GATHER_VEC:
      for (unsigned k = 0; k < vn; ++k) gather_any((char *)out + k * sz, vn, (const char *)in + k * sz, vn, map, dom, op);
      break;
    case L_SCATTER_VEC:
      for (unsigned k = 0; k < vn; ++k) scatter_any((char *)out + k * sz, vn, (const char *)in + k * sz, vn, map, dom);
      break;
    case L_INIT_VEC:
      for (unsigned k = 0; k < vn; ++k) init_any((char *)out + k * sz, vn, map, dom, op);
      break;
    case L_GATHER_MANY:  for (unsigned k = 0; k < vn; ++k) gather_any(po[k], 1, pi[k], 1, map, dom, op); break;
    case L_SCATTER_MANY: for (unsigned k = 0; k < vn; ++k) scatter_any(po[k], 1, pi[k], 1, map, dom); break;
    case L_INIT_MANY:    for (unsigned k = 0; k < vn; ++k) init_any(po[k], 1, map, dom, op); break;
    case L_GATHER_VEC_TO_MANY:
      for (unsigned k = 0; k < vn; ++k) gather_any(po[k], 1, (const char *)in + k * sz, vn, map, dom, op);
      break;
    case L_SCATTER_MANY_TO_VEC:
      for (unsigned k = 0; k < vn; ++k) scatter_any((char *)out + k * sz, vn, pi[k], 1, map, dom);
      break;
    case L_SCATTER_VEC_TO_MANY:
      for (unsigned k = 0; k < vn; ++k) scatter_any(po[k], 1, (const char *)in + k * sz, vn, map, dom);
      break;
    default: fprintf(stderr, "oracle: ogs_local: bad entry %d\n", which); exit(1);
  }
}
