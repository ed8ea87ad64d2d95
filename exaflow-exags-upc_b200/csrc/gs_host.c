/* gs_host.c -- gs_setup / gs / gs_vec / gs_many / gs_free / gs_unique, the
 * Fortran bindings and the public gs_local.h entry points.
 *
 * Host driver of the hot path.  The reference's gs_aux (src/gs.upc:1169-1185)
 * runs four serial passes per call: local gather, init, remote exec, local
 * scatter.  Here one call is
 *
 *   np == 1:  one fused kernel over the sliced-ELL map in HBM (flagged
 *             singles are one-member groups of that map: no init pass);
 *   np  > 1:  groups are split at setup into interior (no remote sharer) and
 *             boundary.  Boundary work -- reduce, pack with peer stores into the
 *             neighbours' mailboxes over NVLink (or NCCL grouped send/recv, or
 *             CUDA-IPC pulls between ranks sharing a GPU), flag wait,
 *             unpack+scatter -- runs on a high-priority stream while the fused
 *             interior kernel runs on the caller's stream; an event joins the
 *             two.  The number of the exchange is device state, so the same
 *             launches can be replayed from a CUDA graph.
 *   host u:   chunked H2D / window waves / D2H pipeline; pageable arrays pass
 *             through pinned bounce rings filled by the host cores.
 *
 * There is no CPU execution path: without a CUDA device gs() fails loudly.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "exags_internal.h"
#include "gs_plan.h"
#include "sell_window.h"

struct dev_csr { struct xcu_csr c; u32 *off, *idx, *nunf; };
#define XG_RING 3                     /* slots of each pinned bounce ring (pageable host arrays) */

struct gs_data {
  comm_ptr comm;
  struct xg_world *w;
  u32 n;
  int method;                 /* method in use (never gs_auto) */
  int on_device;
  /* host copies, kept for exags_gs_dump_map / exags_gs_info */
  struct xg_hmap hall;        /* every group with >= 2 members */
  u32 *hfp; u32 nfp;          /* all flagged primaries (src/gs.upc:359-370) */
  struct xg_plan plan;        /* np > 1 */
  u64 sectors_f64;            /* distinct 32-byte sectors of a double field touched */
  /* device-built handles (np == 1): hall / hfp stay in HBM and are fetched
     into the host copies above only when the introspection calls ask        */
  int dev_built;
  struct xcu_dmap d_hall;
  u32 *d_hfp;
  /* device */
  struct dev_csr d_int;       /* interior groups (all groups when np == 1): CSR form,
                                 or only the long-group residue when d_sell is used */
  struct xcu_sell d_sell;     /* interior groups, sliced-ELL form (symmetric maps) */
  void *d_sell_mem[3];
  size_t sell_idx_entries;
  /* pipelined host-pointer path (np == 1, symmetric map): chunk schedule */
  u32 nwin, *win_min, *win_max;       /* host copies from the sliced layout */
  u32 *h_bidx; size_t h_nbidx;        /* boundary groups' member indices (np > 1) */
  unsigned char *pipe_bnd;            /* [chunks] chunk holds a boundary member: final only after the unpack */
  unsigned pipe_chunks;               /* 0: not available */
  u32 *pipe_run;                      /* [chunks] windows runnable once chunks 0..k are on the device */
  u32 *pipe_out_wave;                 /* [chunks] wave after which chunk j is final */
  void **pipe_ev;                     /* [2*chunks] events */
  /* pageable caller arrays on the pipelined path: pinned bounce rings */
  char *ring_in[XG_RING], *ring_out[XG_RING];
  size_t ring_bytes;
  void *ring_ev_in[XG_RING], *ring_ev_out[XG_RING];
  u32 nfs;                    /* flagged primaries heading no group, not boundary: one-member groups
                                 of the interior map (their reduction is the identity gs_init writes) */
  struct xcu_bnd d_bnd;       /* boundary primaries */
  void *d_bnd_mem[20];
  /* push transport (pairwise over peer stores): this handle's mailbox and the
     device tables that say where each neighbour's data area is */
  struct xg_mbox mbox;
  const unsigned char *d_snb[2];
  u32 push_nnbmax;
  /* symmetric maps: boundary groups in sliced-ELL form too, and the local
     primary behind every send slot, in slot order */
  int bnd_fast;
  struct xcu_sell d_bsell;
  struct dev_csr d_brest;
  const u32 *d_sprim[2];
  size_t handle_bytes;
};

/* ------------------------------------------------------------------ */
static void *upload(const void *src, size_t bytes)
{
  void *d = xcu_malloc(bytes);
  xcu_h2d(d, src, bytes, 0);
  return d;
}

static void upload_csr(struct dev_csr *d, const struct xg_hmap *m)
{
  memset(d, 0, sizeof *d);
  d->c.ng = m->ng; d->c.ni = m->ni;
  if (!m->ng) return;
  d->off = (u32 *)upload(m->off, ((size_t)m->ng + 1) * sizeof(u32));
  d->idx = (u32 *)upload(m->idx, (size_t)m->ni * sizeof(u32));
  if (m->nunf) d->nunf = (u32 *)upload(m->nunf, (size_t)m->ng * sizeof(u32));
  d->c.off = d->off; d->c.idx = d->idx; d->c.nunf = d->nunf;
}

static u64 count_sectors(const struct xg_hmap *m, unsigned elem_bytes)
{
  /* distinct 32-byte sectors of a T[n] field holding a participating DOF;
     members of different groups are distinct indices, so mark and count */
  u64 cnt = 0;
  if (!m->ni) return 0;
  u32 maxi = 0;
  const int nthr = xg_setup_threads();
  (void)nthr;
#ifdef _OPENMP
#pragma omp parallel for reduction(max : maxi) num_threads(nthr)
#endif
  for (u32 k = 0; k < m->ni; ++k) if (m->idx[k] > maxi) maxi = m->idx[k];
  size_t nsec = ((size_t)maxi * elem_bytes) / 32 + 1;
  unsigned char *mark = xnew0(unsigned char, nsec);
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthr)
#endif
  for (u32 k = 0; k < m->ni; ++k) mark[((size_t)m->idx[k] * elem_bytes) / 32] = 1;   /* same value from every writer */
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : cnt) num_threads(nthr)
#endif
  for (size_t s = 0; s < nsec; ++s) cnt += mark[s];
  free(mark);
  return cnt;
}

/* ------------------------------------------------------------------ */
/*  exchange (np > 1)                                                  */
/* ------------------------------------------------------------------ */
static void exchange_pairwise(struct gs_data *g, unsigned vn, int xd, int transpose,
                              void *sendbuf, void *recvbuf, void *stream)
{
  const struct xg_plan *pl = &g->plan;
  const int sd = 1 ^ transpose, rd = 0 ^ transpose;   /* src/gs.upc:479 */
  const size_t unit = (size_t)vn * xg_dom_bytes(xd);
  if (!pl->nnb[sd] && !pl->nnb[rd]) return;
  xg_nccl_group_start(g->w);
  for (u32 k = 0; k < pl->nnb[rd]; ++k)
    xg_nccl_recv(g->w, (char *)recvbuf + (size_t)pl->nb_start[rd][k] * unit,
                 (size_t)pl->nb_size[rd][k] * vn, xd, (int)pl->nb_rank[rd][k], stream);
  for (u32 k = 0; k < pl->nnb[sd]; ++k)
    xg_nccl_send(g->w, (char *)sendbuf + (size_t)pl->nb_start[sd][k] * unit,
                 (size_t)pl->nb_size[sd][k] * vn, xd, (int)pl->nb_rank[sd][k], stream);
  xg_nccl_group_end(g->w);
}

/* Pull transport (ranks sharing a GPU, or EXAGS_EXCHANGE=ipc): senders pack
   into an IPC-exported buffer, a host barrier says "packed", receivers copy
   their segments out of the peers' buffers with the copy engine, a second
   barrier says "pulled".  No kernel ever waits on another rank.  It costs two
   host round trips per call, so NCCL is the default wherever it can run.    */
static void exchange_ipc_pairwise(struct gs_data *g, unsigned vn, int xd, int transpose,
                                  void *recvbuf, void *stream)
{
  const struct xg_plan *pl = &g->plan;
  const int rd = 0 ^ transpose;
  const size_t unit = (size_t)vn * xg_dom_bytes(xd);
  xcu_stream_sync(stream);
  xg_host_barrier(g->w);
  for (u32 k = 0; k < pl->nnb[rd]; ++k)
    xcu_d2d((char *)recvbuf + (size_t)pl->nb_start[rd][k] * unit,
            (const char *)xg_ipc_peer(g->w, (int)pl->nb_rank[rd][k]) + (size_t)pl->nb_peer_start[rd][k] * unit,
            (size_t)pl->nb_size[rd][k] * unit, stream);
  xcu_stream_sync(stream);
  xg_host_barrier(g->w);
}

/* all-reduce on the pull transport: every rank folds all ranks' dense
   vectors in rank order (identical result everywhere) */
static void exchange_ipc_allreduce(struct gs_data *g, void *acc, size_t count, int xd, int xop, void *stream)
{
  struct xg_world *w = g->w;
  const int np = xg_world_np(w);
  const size_t bytes = count * xg_dom_bytes(xd);
  void *tmp = xg_world_scratch(w, 1, bytes ? bytes : 1);
  xcu_stream_sync(stream);
  xg_host_barrier(w);
  xcu_d2d(acc, xg_ipc_peer(w, 0), bytes, stream);
  for (int p = 1; p < np; ++p) {
    xcu_d2d(tmp, xg_ipc_peer(w, p), bytes, stream);
    xcu_array_op(stream, acc, tmp, count, xd, xop);
  }
  xcu_stream_sync(stream);
  xg_host_barrier(w);
}

/* Push transport: the exchange is part of the pack kernel (peer stores over
   NVLink + a flag per neighbour), then a one-CTA wait on my own flags.  Device
   tables, rebuilt whenever the mailbox is (re)created:
     u64 base[parity][d][nnbmax]   neighbour k's data area for that parity
     u64 sig[nsym]                 my flag inside each neighbour's mailbox
     u32 delta[d][nnbmax]          neighbour's slot number - mine
     u32 wait[nsym]                neighbour ranks (= flag numbers in my mailbox)
     u32 start[d][nnbmax+1]        first send slot of each neighbour              */
static void push_prepare(struct gs_data *g, size_t data_bytes)
{
  struct xg_world *w = g->w;
  const struct xg_plan *pl = &g->plan;
  struct xg_mbox *mb = &g->mbox;
  if (!xg_mbox_ensure(w, mb, pl->sym_rank, pl->nsym, data_bytes) && mb->d_tab) return;
  const u32 nm = pl->nnb[0] > pl->nnb[1] ? pl->nnb[0] : pl->nnb[1], M = nm ? nm : 1, ns = pl->nsym;
  const size_t n64 = 4 * (size_t)M + ns, n32 = 2 * (size_t)M + ns + 2 * ((size_t)M + 1);
  u64 *t64 = xnew0(u64, n64 + (n32 + 1) / 2 + 1);
  u32 *t32 = (u32 *)(t64 + n64);
  const int me = xg_world_rank(w);
  for (int par = 0; par < 2; ++par)
    for (int d = 0; d < 2; ++d)
      for (u32 k = 0; k < pl->nnb[d]; ++k)
        t64[((size_t)par * 2 + d) * M + k] =
          (u64)(uintptr_t)((char *)mb->peer[pl->nb_rank[d][k]] + XG_MBOX_HDR + (size_t)par * mb->data_bytes);
  for (u32 i = 0; i < ns; ++i)
    t64[4 * (size_t)M + i] = (u64)(uintptr_t)((char *)mb->peer[pl->sym_rank[i]] + 4 * (size_t)me);
  for (int d = 0; d < 2; ++d)
    for (u32 k = 0; k < pl->nnb[d]; ++k)
      t32[(size_t)d * M + k] = pl->nb_peer_start[d][k] - pl->nb_start[d][k];
  for (u32 i = 0; i < ns; ++i) t32[2 * (size_t)M + i] = pl->sym_rank[i];
  for (int d = 0; d < 2; ++d) {
    u32 *st = t32 + 2 * (size_t)M + ns + (size_t)d * (M + 1);
    for (u32 k = 0; k <= M; ++k) st[k] = k < pl->nnb[d] ? pl->nb_start[d][k] : pl->total[d];
  }
  if (mb->d_tab) { xcu_device_sync(); xcu_free(mb->d_tab); }
  mb->d_tab = upload(t64, (n64 + (n32 + 1) / 2 + 1) * sizeof(u64));
  xcu_device_sync();
  free(t64);
  g->push_nnbmax = M;
}

static double push_timeout(void)
{
  static double t = -1;
  if (t < 0) { const char *v = getenv("EXAGS_PUSH_TIMEOUT_S"); t = v && atof(v) > 0 ? atof(v) : 60.0; }
  return t;
}

static void push_check(struct gs_data *g, const char *when)
{
  int *st = xg_world_push_status(g->w);
  if (st && *st)
    xfail("gs: push exchange timed out waiting for rank %d (%s); is every rank calling gs collectively?", *st - 1, when);
}

/* ------------------------------------------------------------------ */
/*  the operator                                                       */
/* ------------------------------------------------------------------ */
static void check_dom_op(int dom, int op, int *xd, const char *who)
{
  *xd = xg_dom_of(dom);
  if (*xd < 0) xfail("%s: datatype %d not in valid range", who, dom);
  if (op < 0 || op >= XO_N) xfail("%s: op %d not in valid range", who, op);
}

/* ---- boundary half of one call (np > 1): pack [+ push] ... exchange, unpack ---- */
struct bnd_ctx {
  int active, allred, push;
  void *sendbuf, *recvbuf;
  struct xcu_push pp;
  const u32 *push_wait;
};

static void field_of(struct xcu_field *f, void *const *ptrs, int mode, unsigned vn, unsigned k0,
                     unsigned step, int buffer_side)
{
  memset(f, 0, sizeof *f);
  f->vn = (vn - k0 < step) ? vn - k0 : step; f->tuple = vn;
  f->k0 = (buffer_side && mode == XM_MANY) ? k0 : 0;     /* component offset inside the [slot][vn] buffers */
  if (mode == XM_MANY) for (unsigned k = 0; k < f->vn; ++k) f->p[k] = ptrs[k0 + k];
  else f->p[0] = ptrs[0];
}

/* Decide whether this rank takes part in an exchange and get the buffers.
   The pull and push transports have collective steps (host barriers, mailbox
   growth): there a rank takes part whenever any rank has something to send. */
static void bnd_prepare(struct gs_data *g, struct bnd_ctx *c, unsigned vn, int xd, int xop, int transpose)
{
  struct xg_world *w = g->w;
  const struct xg_plan *pl = &g->plan;
  const unsigned esz = xg_dom_bytes(xd);
  memset(c, 0, sizeof *c);
  c->allred = g->method == gs_all_reduce;
  /* gs_bpr has no NCCL reduction (allreduce_exec goes through comm_allreduce,
     which knows every op, src/gs.upc:1006-1026): such calls take the pairwise
     plan, which every handle has and which gives the same values */
  if (c->allred && xop == XO_BPR && !xg_world_ipc(w)) c->allred = 0;
  c->push = xg_world_push(w) && !c->allred;
  c->active = xg_world_np(w) > 1 &&
    (c->allred || ((c->push || xg_world_ipc(w)) ? pl->max_total > 0 : pl->nb > 0));
  if (!c->active) return;
  if (c->push) {
    push_check(g, "before this call");
    push_prepare(g, (size_t)pl->max_total * vn * esz);
    /* nothing below depends on how many exchanges this mailbox has seen: the
       exchange number (and with it the parity of the data area) is read from
       the mailbox header by the kernels, so the same launches can be replayed
       from a CUDA graph */
    struct xg_mbox *mb = &g->mbox;
    const u32 M = g->push_nnbmax;
    const int sd = 1 ^ transpose;
    const u64 *t64 = (const u64 *)mb->d_tab;
    const u32 *t32 = (const u32 *)(t64 + 4 * (size_t)M + pl->nsym);
    c->pp.base = (const unsigned long long *)(t64 + (size_t)sd * M);
    c->pp.par_stride = 2 * M;
    c->pp.sig = (const unsigned long long *)(t64 + 4 * (size_t)M);
    c->pp.delta = t32 + (size_t)sd * M;
    c->pp.snb = g->d_snb[sd];
    c->pp.nsig = pl->nsym;
    c->pp.depoch = (u32 *)((char *)mb->mine + XG_MBOX_EPOCH); c->pp.epoch_imm = 0;
    c->pp.counter = (u32 *)((char *)mb->mine + XG_MBOX_COUNTER);
    c->push_wait = t32 + 2 * (size_t)M;
    c->pp.start = t32 + 2 * (size_t)M + pl->nsym + (size_t)sd * (M + 1);
    c->pp.nnb = pl->nnb[sd];
    c->recvbuf = (char *)mb->mine + XG_MBOX_HDR;          /* parity 0; the unpack adds the parity */
  } else if (xg_world_ipc(w)) {
    c->sendbuf = xg_ipc_buffer(w, (size_t)pl->max_total * vn * 8);
    c->recvbuf = xg_world_scratch(w, c->allred ? 3 : 1, (size_t)pl->max_total * vn * esz);
  } else if (c->allred) {
    c->sendbuf = xg_world_scratch(w, 0, (size_t)pl->total_shared * vn * esz);
    c->recvbuf = c->sendbuf;
  } else {
    c->sendbuf = xg_world_scratch(w, 0, (size_t)pl->total[1 ^ transpose] * vn * esz);
    c->recvbuf = xg_world_scratch(w, 1, (size_t)pl->total[0 ^ transpose] * vn * esz);
  }
}

/* gather + pack (with the push transport this is also the send) */
static void bnd_pack_phase(struct gs_data *g, struct bnd_ctx *c, void *const *ptrs, int mode, unsigned vn,
                           int xd, int xop, int transpose, void *scomm, int phase)
{
  const unsigned step = (mode == XM_MANY) ? XG_MAXV : vn;
  if (c->allred) xcu_fill_identity(scomm, c->sendbuf, (size_t)g->plan.total_shared * vn, xd, xop);
  if (g->bnd_fast && !c->allred) {
    /* symmetric map: reduce the boundary groups into their primaries with the
       sliced kernel (gather-only form; the unpack writes the members), then
       fill the send slots in slot order */
    const int sd = 1 ^ transpose;
    for (unsigned k0 = 0; k0 < vn && (phase & 1); k0 += step) {
      struct xcu_field f; field_of(&f, ptrs, mode, vn, k0, step, 0);
      if (g->d_bsell.nblock) xcu_gather_sell(scomm, &g->d_bsell, &f, mode, xd, xop);
      xcu_fused(scomm, &g->d_brest.c, &f, mode, xd, xop, transpose);
    }
    for (unsigned k0 = 0; k0 < vn && (phase & 2); k0 += step) {
      struct xcu_field f; field_of(&f, ptrs, mode, vn, k0, step, 1);
      struct xcu_push p1 = c->pp;
      if (k0 + step < vn) p1.nsig = 0;
      xcu_pack_slots(scomm, g->d_sprim[sd], g->plan.total[sd], &f, mode, xd, c->sendbuf, vn, c->push ? &p1 : 0);
    }
    return;
  }
  if (!(phase & 1)) return;
  for (unsigned k0 = 0; k0 < vn; k0 += step) {
    struct xcu_field f; field_of(&f, ptrs, mode, vn, k0, step, 1);
    if (c->push) {
      struct xcu_push p1 = c->pp;
      if (k0 + step < vn) p1.nsig = 0;        /* only the last launch signals */
      xcu_bnd_pack_push(scomm, &g->d_bnd, &f, mode, xd, xop, transpose, vn, &p1);
    } else
      xcu_bnd_pack(scomm, &g->d_bnd, &f, mode, xd, xop, transpose, c->sendbuf, vn, c->allred);
  }
}

static void bnd_pack(struct gs_data *g, struct bnd_ctx *c, void *const *ptrs, int mode, unsigned vn,
                     int xd, int xop, int transpose, void *scomm)
{ bnd_pack_phase(g, c, ptrs, mode, vn, xd, xop, transpose, scomm, 3); }

/* exchange (or wait for the neighbours' stores) + unpack and scatter */
static void bnd_unpack(struct gs_data *g, struct bnd_ctx *c, void *const *ptrs, int mode, unsigned vn,
                       int xd, int xop, int transpose, void *scomm, const void *wgt)
{
  struct xg_world *w = g->w;
  const unsigned step = (mode == XM_MANY) ? XG_MAXV : vn;
  if (c->push)
    xcu_push_wait(scomm, (const u32 *)g->mbox.mine, c->push_wait, g->plan.nsym, c->pp.depoch, 0,
                  push_timeout(), xg_world_push_status(w));
  else if (xg_world_ipc(w)) {
    if (c->allred) exchange_ipc_allreduce(g, c->recvbuf, (size_t)g->plan.total_shared * vn, xd, xop, scomm);
    else exchange_ipc_pairwise(g, vn, xd, transpose, c->recvbuf, scomm);
  } else if (c->allred) xg_nccl_allreduce(w, c->sendbuf, (size_t)g->plan.total_shared * vn, xd, xop, scomm);
  else exchange_pairwise(g, vn, xd, transpose, c->sendbuf, c->recvbuf, scomm);
  for (unsigned k0 = 0; k0 < vn; k0 += step) {
    struct xcu_field f; field_of(&f, ptrs, mode, vn, k0, step, 1);
    f.w = wgt;
    xcu_bnd_unpack(scomm, &g->d_bnd, &f, mode, xd, xop, transpose, c->recvbuf, vn, c->allred,
                   c->push ? c->pp.depoch : 0, c->push ? g->mbox.data_bytes : 0);
  }
}

/* EXAGS_TRACE=1: run the stages of every call one after the other on the
   caller's stream with timing events around them and print the per-stage
   means at gs_free -- the cost of each stage alone, without the overlap.    */
static int g_trace = -1;
static double g_trace_ms[5]; static unsigned long g_trace_calls;
static void interior_enqueue(struct gs_data *g, void *const *ptrs, int mode, unsigned vn,
                             int xd, int xop, int transpose, void *stream, const void *wgt);

static void gs_enqueue_traced(struct gs_data *g, void *const *ptrs, int mode, unsigned vn,
                              int xd, int xop, int transpose, void *stream, const void *wgt)
{
  static void *ev[6];
  if (!ev[0]) for (int i = 0; i < 6; ++i) ev[i] = xcu_event_create_timed();
  struct bnd_ctx bc;
  bnd_prepare(g, &bc, vn, xd, xop, transpose);
  xcu_event_record(ev[0], stream);
  if (bc.active) bnd_pack_phase(g, &bc, ptrs, mode, vn, xd, xop, transpose, stream, 1);
  xcu_event_record(ev[1], stream);
  if (bc.active) bnd_pack_phase(g, &bc, ptrs, mode, vn, xd, xop, transpose, stream, 2);
  xcu_event_record(ev[2], stream);
  interior_enqueue(g, ptrs, mode, vn, xd, xop, transpose, stream, wgt);
  xcu_event_record(ev[3], stream);
  if (bc.active) bnd_unpack(g, &bc, ptrs, mode, vn, xd, xop, transpose, stream, wgt);
  xcu_event_record(ev[4], stream);
  xcu_stream_sync(stream);
  /* the first launches of a kernel instantiation include module loading */
  static unsigned char seen[XD_N][XO_N][3];
  if (seen[xd][xop][mode] < 3) { ++seen[xd][xop][mode]; return; }
  for (int i = 0; i < 4; ++i) g_trace_ms[i] += xcu_event_ms(ev[i], ev[i + 1]);
  ++g_trace_calls;
}

static void trace_report(struct gs_data *g)
{
  if (g_trace <= 0 || !g_trace_calls) return;
  printf("[exags trace rank %d] %lu calls: boundary reduce %.4f ms, pack(+send) %.4f ms, interior %.4f ms, wait+unpack %.4f ms "
         "(stages serialised; boundary groups %u in %u sliced blocks, send slots %u, interior sliced blocks %u)\n",
         xg_world_rank(g->w), g_trace_calls, g_trace_ms[0] / g_trace_calls, g_trace_ms[1] / g_trace_calls,
         g_trace_ms[2] / g_trace_calls, g_trace_ms[3] / g_trace_calls, g->plan.nb, g->d_bsell.nblock, g->plan.total[1], g->d_sell.nblock);
  fflush(stdout);
  g_trace_calls = 0; memset(g_trace_ms, 0, sizeof g_trace_ms);
}

static void interior_enqueue(struct gs_data *g, void *const *ptrs, int mode, unsigned vn,
                             int xd, int xop, int transpose, void *stream, const void *wgt)
{
  const unsigned step = (mode == XM_MANY) ? XG_MAXV : vn;
  for (unsigned k0 = 0; k0 < vn; k0 += step) {
    struct xcu_field f; field_of(&f, ptrs, mode, vn, k0, step, 0);
    f.w = wgt;      /* weighted gs: the stores of the group results carry the weight */
    if (g->d_sell.nblock) xcu_fused_sell(stream, &g->d_sell, &f, mode, xd, xop, transpose);
    xcu_fused(stream, &g->d_int.c, &f, mode, xd, xop, transpose);
  }
}

/* Enqueue one gs on device-resident data.  ptrs: vn device pointers for
   mode many, else ptrs[0].  Work is ordered on `stream`.                   */
static void gs_enqueue(struct gs_data *g, void *const *ptrs, int mode, unsigned vn,
                       int xd, int xop, int transpose, void *stream, const void *wgt)
{
  struct xg_world *w = g->w;
  struct bnd_ctx bc;
  void *scomm = xg_world_stream(w, 1);
  if (g_trace < 0) { const char *v = getenv("EXAGS_TRACE"); g_trace = v ? atoi(v) : 0; }
  if (g_trace > 0) { gs_enqueue_traced(g, ptrs, mode, vn, xd, xop, transpose, stream, wgt); return; }
  bnd_prepare(g, &bc, vn, xd, xop, transpose);
  if (bc.active) {
    /* fork: boundary stream starts after whatever precedes us on `stream` */
    xcu_event_record(xg_world_event(w, 0), stream);
    xcu_stream_wait_event(scomm, xg_world_event(w, 0));
    bnd_pack(g, &bc, ptrs, mode, vn, xd, xop, transpose, scomm);
  }
  /* interior: fused gather-scatter on the caller's stream */
  interior_enqueue(g, ptrs, mode, vn, xd, xop, transpose, stream, wgt);
  if (bc.active) {
    bnd_unpack(g, &bc, ptrs, mode, vn, xd, xop, transpose, scomm, wgt);
    /* join */
    xcu_event_record(xg_world_event(w, 1), scomm);
    xcu_stream_wait_event(stream, xg_world_event(w, 1));
  }
}

/* Host-pointer fast path (np == 1, symmetric sliced-ELL map): the field is cut
   into contiguous chunks; chunk k goes up on the copy-in stream, the windows
   whose members all lie in chunks 0..k run as wave k, and chunk j comes back
   on the copy-out stream as soon as every window that touches it has run.
   Windows ascend by their first member, so both sets are prefixes of the
   window list.  PCIe then carries H2D and D2H at the same time and the
   kernels hide under the copies; any map the schedule cannot cover (flagged
   labels, long groups, ranks > 1) takes the plain stage-run-copy path.       */
static void pipe_plan(struct gs_data *g)
{
  if (g->pipe_chunks || !g->nwin) return;
  const size_t n = g->n;
  size_t chunk = 24u << 20;                       /* bytes of a double field per chunk */
  { const char *v = getenv("EXAGS_PIPE_CHUNK_BYTES"); if (v && atol(v) >= 256) chunk = (size_t)atol(v); }
  unsigned C = (unsigned)(((size_t)n * 8 + chunk - 1) / chunk);
  if (C < 2) return;
  if (C > 128) C = 128;
  const size_t per = (n + C - 1) / C;
  g->pipe_run = xnew(u32, C); g->pipe_out_wave = xnew(u32, C);
  u32 r = 0, pm = 0;
  for (unsigned k = 0; k < C; ++k) {
    size_t end = (size_t)(k + 1) * per; if (end > n) end = n;     /* chunk k = [k*per, end) */
    while (r < g->nwin && (g->win_max[r] > pm ? g->win_max[r] : pm) < end) { if (g->win_max[r] > pm) pm = g->win_max[r]; ++r; }
    if (k == C - 1) r = g->nwin;
    g->pipe_run[k] = r;
  }
  u32 s0 = 0;
  for (unsigned j = 0; j < C; ++j) {
    size_t end = (size_t)(j + 1) * per; if (end > n) end = n;
    while (s0 < g->nwin && g->win_min[s0] < end) ++s0;            /* windows [0,s0) may touch chunk j */
    unsigned k = j;
    while (k < C - 1 && g->pipe_run[k] < s0) ++k;
    g->pipe_out_wave[j] = k;
  }
  g->pipe_bnd = xnew0(unsigned char, C);
  for (size_t k = 0; k < g->h_nbidx; ++k) g->pipe_bnd[g->h_bidx[k] / per] = 1;
  g->pipe_ev = xnew(void *, 2 * (size_t)C);
  for (unsigned k = 0; k < 2 * C; ++k) g->pipe_ev[k] = xcu_event_create();
  g->pipe_chunks = C;
}

/* memcpy spread over this rank's share of the host cores */
static void par_memcpy(void *dst, const void *src, size_t bytes)
{
  const int nthr = xg_setup_threads();
  const size_t piece = (size_t)1 << 20;
  if (nthr <= 1 || bytes < 4 * piece) { memcpy(dst, src, bytes); return; }
  const long np = (long)((bytes + piece - 1) / piece);
  (void)nthr;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthr)
#endif
  for (long i = 0; i < np; ++i) {
    const size_t o = (size_t)i * piece, len = bytes - o < piece ? bytes - o : piece;
    memcpy((char *)dst + o, (const char *)src + o, len);
  }
}

/* chunks on their way out of the pinned ring into the caller's pageable array, oldest first */
struct ring_out_q { unsigned head, count, chunk[XG_RING]; };

static void ring_ensure(struct gs_data *g, size_t bytes)
{
  if (g->ring_bytes >= bytes) return;
  xcu_device_sync();
  for (int k = 0; k < XG_RING; ++k) {
    xcu_host_free(g->ring_in[k]); xcu_host_free(g->ring_out[k]);
    g->ring_in[k] = (char *)xcu_host_alloc(bytes); g->ring_out[k] = (char *)xcu_host_alloc(bytes);
    if (!g->ring_ev_in[k]) { g->ring_ev_in[k] = xcu_event_create(); g->ring_ev_out[k] = xcu_event_create(); }
  }
  g->ring_bytes = bytes;
}

/* pageable: the caller's arrays are ordinary host memory (the reference's
   calling convention, src/gs.upc:1187-1191).  A copy engine cannot read them,
   and the driver's own staging of a cudaMemcpyAsync on pageable memory is
   serial and slow (115 ms per call at M7 against 18 ms from pinned memory,
   profiles/r2_bench_N1_b_pageable.json).  Chunks therefore pass through two
   small rings of pinned buffers, filled and emptied by this rank's host cores
   while the copy engines and the kernels work on the neighbouring chunks.   */
static void gs_run_host_pipelined(struct gs_data *g, void *const *ptrs, int mode, unsigned vn,
                                  int xd, int xop, int transpose, int pageable)
{
  struct xg_world *w = g->w;
  const unsigned C = g->pipe_chunks;
  const size_t n = g->n, esz = xg_dom_bytes(xd), per = (n + C - 1) / C;
  const unsigned parts = mode == XM_MANY ? vn : 1;
  const size_t tuple = mode == XM_VEC ? vn : 1;           /* elements per index */
  const size_t each = n * esz * tuple;
  const size_t cbytes = per * esz * tuple;                /* one chunk of one part */
  char *stage = (char *)xg_world_scratch(w, 2, each * parts);
  void *s_in = xg_world_stream(w, 2), *s_run = xg_world_stream(w, 0), *s_out = xg_world_stream(w, 3);
  void *scomm = xg_world_stream(w, 1);
  struct xcu_field f; memset(&f, 0, sizeof f);
  f.vn = vn; f.tuple = vn; f.k0 = 0;
  void *dptr[XG_MAXV];
  if (mode == XM_MANY) { if (vn > XG_MAXV) xfail("gs: internal: pipelined path with %u arrays", vn); for (unsigned k = 0; k < vn; ++k) dptr[k] = f.p[k] = stage + each * k; }
  else dptr[0] = f.p[0] = stage;
  if (pageable) ring_ensure(g, cbytes * parts);
  struct ring_out_q q; memset(&q, 0, sizeof q);
  unsigned nout = 0;
  /* np > 1: the boundary groups run once the whole field is up (their members
     sit in the first and last layers of a slab); the chunks they touch go
     back after the unpack, every other chunk as soon as its windows are done */
  struct bnd_ctx bc;
  bnd_prepare(g, &bc, vn, xd, xop, transpose);
  u32 done = 0;
#define CHUNK_BYTES(j) ((((size_t)(j) * per + per < n ? (size_t)(j) * per + per : n) - (size_t)(j) * per) * esz * tuple)
#define RING_DRAIN_ONE() do {                                                              \
    const unsigned j_ = q.chunk[q.head], sl_ = (nout - q.count) % XG_RING;                 \
    xcu_event_sync(g->ring_ev_out[sl_]);                                                   \
    for (unsigned p = 0; p < parts; ++p)                                                   \
      par_memcpy((char *)ptrs[p] + (size_t)j_ * per * esz * tuple, g->ring_out[sl_] + cbytes * p, CHUNK_BYTES(j_)); \
    q.head = (q.head + 1) % XG_RING; --q.count;                                            \
  } while (0)
#define CHUNK_OUT(j) do {                                                                  \
    const size_t olo_ = (size_t)(j) * per, ob_ = CHUNK_BYTES(j);                           \
    if (!pageable) {                                                                       \
      for (unsigned p = 0; p < parts; ++p)                                                 \
        xcu_d2h((char *)ptrs[p] + olo_ * esz * tuple, stage + each * p + olo_ * esz * tuple, ob_, s_out); \
    } else {                                                                               \
      if (q.count == XG_RING) RING_DRAIN_ONE();                                            \
      const unsigned sl_ = nout % XG_RING;                                                 \
      for (unsigned p = 0; p < parts; ++p)                                                 \
        xcu_d2h(g->ring_out[sl_] + cbytes * p, stage + each * p + olo_ * esz * tuple, ob_, s_out); \
      xcu_event_record(g->ring_ev_out[sl_], s_out);                                        \
      q.chunk[(q.head + q.count) % XG_RING] = (unsigned)(j); ++q.count; ++nout;            \
    }                                                                                      \
  } while (0)
  for (unsigned k = 0; k < C; ++k) {
    size_t lo = (size_t)k * per, hi = lo + per < n ? lo + per : n;
    if (pageable) {
      const unsigned sl = k % XG_RING;
      if (k >= XG_RING) xcu_event_sync(g->ring_ev_in[sl]);       /* its previous upload has left the slot */
      for (unsigned p = 0; p < parts; ++p)
        par_memcpy(g->ring_in[sl] + cbytes * p, (const char *)ptrs[p] + lo * esz * tuple, (hi - lo) * esz * tuple);
      for (unsigned p = 0; p < parts; ++p)
        xcu_h2d(stage + each * p + lo * esz * tuple, g->ring_in[sl] + cbytes * p, (hi - lo) * esz * tuple, s_in);
      xcu_event_record(g->ring_ev_in[sl], s_in);
    } else
      for (unsigned p = 0; p < parts; ++p)
        xcu_h2d(stage + each * p + lo * esz * tuple, (const char *)ptrs[p] + lo * esz * tuple, (hi - lo) * esz * tuple, s_in);
    xcu_event_record(g->pipe_ev[k], s_in);
    xcu_stream_wait_event(s_run, g->pipe_ev[k]);
    if (g->pipe_run[k] > done) {
      struct xcu_sell m = g->d_sell;
      m.idx += (size_t)done * XG_SELL_WB * 256; m.mask += (size_t)done * XG_SELL_WB * 32;
      if (m.fmask) m.fmask += (size_t)done * XG_SELL_WB * 32;
      m.nblock = (g->pipe_run[k] - done) * XG_SELL_WB;
      xcu_fused_sell(s_run, &m, &f, mode, xd, xop, transpose);
      done = g->pipe_run[k];
    }
    xcu_event_record(g->pipe_ev[C + k], s_run);
    if (k == C - 1 && bc.active) {
      xcu_stream_wait_event(scomm, g->pipe_ev[k]);
      bnd_pack(g, &bc, dptr, mode, vn, xd, xop, transpose, scomm);
      bnd_unpack(g, &bc, dptr, mode, vn, xd, xop, transpose, scomm, 0);
      xcu_event_record(xg_world_event(w, 1), scomm);
    }
    for (unsigned j = 0; j < C; ++j) {
      if (g->pipe_out_wave[j] != k || (bc.active && g->pipe_bnd[j])) continue;
      xcu_stream_wait_event(s_out, g->pipe_ev[C + k]);
      CHUNK_OUT(j);
    }
    /* hand finished chunks back while the next ones are in flight */
    while (pageable && q.count && xcu_event_done(g->ring_ev_out[(nout - q.count) % XG_RING])) RING_DRAIN_ONE();
  }
  if (bc.active) {
    xcu_stream_wait_event(s_out, xg_world_event(w, 1));
    xcu_stream_wait_event(s_out, g->pipe_ev[2 * C - 1]);
    for (unsigned j = 0; j < C; ++j) {
      if (!g->pipe_bnd[j]) continue;
      CHUNK_OUT(j);
    }
  }
  while (q.count) RING_DRAIN_ONE();
#undef CHUNK_OUT
#undef RING_DRAIN_ONE
#undef CHUNK_BYTES
  xcu_stream_sync(s_out);
  xcu_stream_sync(s_run);
  if (bc.active) xcu_stream_sync(scomm);
  push_check(g, "during this call");
}

/* Pageable host arrays.  The reference's u is an ordinary host array
   (src/gs.upc:1187-1191; Nek passes Fortran arrays from static storage).  A
   cudaMemcpyAsync on pageable memory is staged by the driver through its own
   small pinned buffers and does not overlap; page-locking the caller's array
   once (cudaHostRegister) lifts it to the speed of pinned memory for every
   later call.  A registration must not outlive the array -- memory that is
   freed while registered and mapped again at the same address would be copied
   from its old pages -- so the library never registers on its own by default:
     EXAGS_HOST_REGISTER=1   register an array the second time gs() sees it
                             (same address and size; >= 1 MB), keep it;
     EXAGS_HOST_REGISTER=2   register on first sight;
     exags_host_register / exags_host_unregister (include/exags_cuda.h) for
                             callers that manage it themselves.
   Callers that keep their fields for the length of the run (Nek5000 does) can
   switch it on safely.                                                       */
#define XG_HOSTREG_MAX 64
static struct { void *p; size_t bytes; unsigned seen; int registered; } g_hostreg[XG_HOSTREG_MAX];
static int g_hostreg_n = 0, g_hostreg_mode = 0;

static void hostreg_note(void *p, size_t bytes)
{
  { const char *e = getenv("EXAGS_HOST_REGISTER"); g_hostreg_mode = e ? atoi(e) : 0; }
  if (g_hostreg_mode <= 0 || bytes < ((size_t)1 << 20) || xcu_ptr_kind(p) != 0) return;
  int k = 0;
  while (k < g_hostreg_n && !(g_hostreg[k].p == p && g_hostreg[k].bytes == bytes)) ++k;
  if (k == g_hostreg_n) {
    if (g_hostreg_n == XG_HOSTREG_MAX) return;
    g_hostreg[k].p = p; g_hostreg[k].bytes = bytes; g_hostreg[k].seen = 0; g_hostreg[k].registered = 0;
    ++g_hostreg_n;
  }
  if (++g_hostreg[k].seen >= (g_hostreg_mode >= 2 ? 1u : 2u) && !g_hostreg[k].registered)
    g_hostreg[k].registered = xcu_host_register(p, bytes) ? 1 : -1;     /* -1: refused, do not retry */
}

int exags_host_register(void *p, size_t bytes) { xg_world_get(); return xcu_host_register(p, bytes); }
void exags_host_unregister(void *p)
{
  for (int k = 0; k < g_hostreg_n; ++k) if (g_hostreg[k].p == p) g_hostreg[k] = g_hostreg[--g_hostreg_n];
  xcu_host_unregister(p);
}

static void gs_run(void *u, int mode, unsigned vn, int dom, int op, unsigned transpose,
                   struct gs_data *g, void *stream, int blocking, const char *who)
{
  int xd;
  if (!g) xfail("%s: null handle", who);
  check_dom_op(dom, op, &xd, who);
  if (!g->on_device)
    xfail("%s: no CUDA device -- this library has no CPU execution path", who);
  if (vn == 0) return;
  transpose = transpose != 0;
  struct xg_world *w = g->w;
  if (blocking) stream = xg_world_stream(w, 0);   /* async callers own `stream` (0 = legacy default) */
  const size_t esz = xg_dom_bytes(xd), n = g->n;

  /* the stream-ordered form is for device-resident data by contract
     (include/exags_cuda.h): no allocation and no pointer query per call */
  void *ptrs_few[XG_MAXV];
  void **ptrs = vn <= XG_MAXV ? ptrs_few : xnew(void *, vn);
  int host = 0;
  if (mode == XM_MANY) {
    void *const *tab = (void *const *)u;
    for (unsigned k = 0; k < vn; ++k) ptrs[k] = tab[k];
    if (blocking) {
      host = !xcu_is_device_ptr(ptrs[0]);
      for (unsigned k = 1; k < vn; ++k)
        if ((!xcu_is_device_ptr(ptrs[k])) != host)
          xfail("%s: gs_many arrays must be all host or all device", who);
    }
  } else {
    ptrs[0] = u;
    if (blocking) host = !xcu_is_device_ptr(u);
  }

  if (!host) {
    gs_enqueue(g, ptrs, mode, vn, xd, op, (int)transpose, stream, 0);
    if (blocking) { xcu_stream_sync(stream); push_check(g, "during this call"); }
    if (ptrs != ptrs_few) free(ptrs);
    return;
  }
  /* host data: stage through HBM (still GPU compute) */
  if (mode == XM_MANY) for (unsigned k = 0; k < vn; ++k) hostreg_note(ptrs[k], n * esz);
  else hostreg_note(u, n * esz * vn);
  if (g->nwin && !(mode == XM_MANY && vn > XG_MAXV)) {
    const char *pp = getenv("EXAGS_HOST_PIPELINE");
    if (!(pp && atoi(pp) == 0)) pipe_plan(g);
    if (g->pipe_chunks) {
      int pageable = 0;
      for (unsigned k = 0; k < (mode == XM_MANY ? vn : 1u); ++k) pageable |= xcu_ptr_kind(ptrs[k]) == 0;
      gs_run_host_pipelined(g, ptrs, mode, vn, xd, op, (int)transpose, pageable);
      if (ptrs != ptrs_few) free(ptrs);
      return;
    }
  }
  size_t each = n * esz, total = each * vn;
  char *stage = (char *)xg_world_scratch(w, 2, total);
  void **dptr = xnew(void *, vn);
  if (mode == XM_MANY) {
    for (unsigned k = 0; k < vn; ++k) { dptr[k] = stage + each * k; xcu_h2d(dptr[k], ptrs[k], each, stream); }
  } else {
    dptr[0] = stage; xcu_h2d(stage, u, total, stream);
  }
  gs_enqueue(g, dptr, mode, vn, xd, op, (int)transpose, stream, 0);
  if (mode == XM_MANY) for (unsigned k = 0; k < vn; ++k) xcu_d2h(ptrs[k], dptr[k], each, stream);
  else xcu_d2h(u, stage, total, stream);
  xcu_stream_sync(stream);
  push_check(g, "during this call");
  free(dptr);
  if (ptrs != ptrs_few) free(ptrs);
}

void exags_gs(void *u, gs_dom dom, gs_op_t op, unsigned transpose, struct gs_data *gsh, buffer *buf)
{ (void)buf; gs_run(u, XM_PLAIN, 1, (int)dom, (int)op, transpose, gsh, 0, 1, "gs"); }

void exags_gs_vec(void *u, unsigned vn, gs_dom dom, gs_op_t op, unsigned transpose,
                  struct gs_data *gsh, buffer *buf)
{ (void)buf; gs_run(u, XM_VEC, vn, (int)dom, (int)op, transpose, gsh, 0, 1, "gs_vec"); }

void exags_gs_many(void *const *u, unsigned vn, gs_dom dom, gs_op_t op, unsigned transpose,
                   struct gs_data *gsh, buffer *buf)
{ (void)buf; gs_run((void *)u, XM_MANY, vn, (int)dom, (int)op, transpose, gsh, 0, 1, "gs_many"); }

void exags_gs_async(void *u, int mode, unsigned vn, int dom, int op, unsigned transpose,
                    struct gs_data *gsh, void *stream)
{
  if (mode < XM_PLAIN || mode > XM_MANY) xfail("gs_async: bad mode %d", mode);
  if (mode == XM_PLAIN) vn = 1;
  gs_run(u, mode, vn, dom, op, transpose, gsh, stream, 0, "gs_async");
}

/* Weighted gs (extension, include/exags_cuda.h): what Nek issues
   as dssum followed by a multiplication with the inverse multiplicity, in the
   one pass over the field that gs makes anyway.  Device vectors, float/double
   fields -- plain, T[n][vn] (one weight per node for all its components, as
   for a velocity field) or vn arrays -- transpose 0.                          */
static void gs_run_weighted(void *u, int mode, unsigned vn, const void *wgt, int dom, int op, struct gs_data *g,
                            void *stream, int blocking, const char *who)
{
  int xd;
  if (!g) xfail("%s: null handle", who);
  check_dom_op(dom, op, &xd, who);
  if (xd != XD_F64 && xd != XD_F32) xfail("%s: floating-point domains only (got %d)", who, dom);
  if (op > XO_MAX) xfail("%s: op %d not in valid range", who, op);
  if (!g->on_device)
    xfail("%s: no CUDA device -- this library has no CPU execution path", who);
  if (!g->n || !vn) return;
  void *ptrs_few[XG_MAXV];
  void **ptrs = vn <= XG_MAXV || mode != XM_MANY ? ptrs_few : xnew(void *, vn);
  if (mode == XM_MANY) for (unsigned k = 0; k < vn; ++k) ptrs[k] = ((void *const *)u)[k];
  else ptrs[0] = u;
  if (blocking) {
    if (!xcu_is_device_ptr(wgt)) xfail("%s: u and w must be device vectors", who);
    for (unsigned k = 0; k < (mode == XM_MANY ? vn : 1u); ++k)
      if (!xcu_is_device_ptr(ptrs[k])) xfail("%s: u and w must be device vectors", who);
    stream = xg_world_stream(g->w, 0);
  }
  gs_enqueue(g, ptrs, mode, vn, xd, op, 0, stream, wgt);
  if (blocking) { xcu_stream_sync(stream); push_check(g, "during this call"); }
  if (ptrs != ptrs_few) free(ptrs);
}

void exags_gs_weighted(void *u, const void *w, int dom, int op, struct gs_data *gsh)
{ gs_run_weighted(u, XM_PLAIN, 1, w, dom, op, gsh, 0, 1, "gs_weighted"); }
void exags_gs_weighted_async(void *u, const void *w, int dom, int op, struct gs_data *gsh, void *stream)
{ gs_run_weighted(u, XM_PLAIN, 1, w, dom, op, gsh, stream, 0, "gs_weighted_async"); }
void exags_gs_weighted_vec(void *u, unsigned vn, const void *w, int dom, int op, struct gs_data *gsh)
{ gs_run_weighted(u, XM_VEC, vn, w, dom, op, gsh, 0, 1, "gs_weighted_vec"); }
void exags_gs_weighted_vec_async(void *u, unsigned vn, const void *w, int dom, int op, struct gs_data *gsh, void *stream)
{ gs_run_weighted(u, XM_VEC, vn, w, dom, op, gsh, stream, 0, "gs_weighted_vec_async"); }
void exags_gs_weighted_many(void *const *u, unsigned vn, const void *w, int dom, int op, struct gs_data *gsh)
{ gs_run_weighted((void *)u, XM_MANY, vn, w, dom, op, gsh, 0, 1, "gs_weighted_many"); }
void exags_gs_weighted_many_async(void *const *u, unsigned vn, const void *w, int dom, int op, struct gs_data *gsh, void *stream)
{ gs_run_weighted((void *)u, XM_MANY, vn, w, dom, op, gsh, stream, 0, "gs_weighted_many_async"); }

/* the Nek flavour shares the operator; only the id width differs */
void gslib_gs(void *u, int dom, int op, unsigned t, struct gs_data *g, buffer *b)
{ (void)b; gs_run(u, XM_PLAIN, 1, dom, op, t, g, 0, 1, "gs"); }
void gslib_gs_vec(void *u, unsigned vn, int dom, int op, unsigned t, struct gs_data *g, buffer *b)
{ (void)b; gs_run(u, XM_VEC, vn, dom, op, t, g, 0, 1, "gs_vec"); }
void gslib_gs_many(void *const *u, unsigned vn, int dom, int op, unsigned t, struct gs_data *g, buffer *b)
{ (void)b; gs_run((void *)u, XM_MANY, vn, dom, op, t, g, 0, 1, "gs_many"); }

/* ------------------------------------------------------------------ */
/*  setup                                                              */
/* ------------------------------------------------------------------ */
/* the exchange plan's per-boundary-primary tables, into d_bnd_mem[m...]      */
static int upload_boundary_plan(struct gs_data *g, int m)
{
  const struct xg_plan *pl = &g->plan;
  struct xcu_bnd *d = &g->d_bnd;
  const u32 nb = pl->nb;
#define UP(dst, src, cnt) do { g->d_bnd_mem[m] = upload((src), (size_t)(cnt) * sizeof(*(src))); (dst) = g->d_bnd_mem[m]; ++m; } while (0)
  for (int dd = 0; dd < 2; ++dd) {
    UP(d->soff[dd], pl->soff[dd], nb + 1);
    UP(d->slot[dd], pl->slot[dd], pl->total[dd] ? pl->total[dd] : 1);
  }
  UP(d->ord, pl->ord, nb);
  UP(d->pflag, pl->pflag, nb);
  for (int dd = 0; dd < 2; ++dd) UP(g->d_snb[dd], pl->snb[dd], pl->total[dd] ? pl->total[dd] : 1);
#undef UP
  return m;
}

/* the local primary behind every send slot, in slot order (symmetric maps) */
static int upload_send_primaries(struct gs_data *g, int m)
{
  const struct xg_plan *pl = &g->plan;
  for (int dd = 0; dd < 2; ++dd) {
    u32 *sp = xnew(u32, pl->total[dd] ? pl->total[dd] : 1);
    for (u32 b = 0; b < pl->nb; ++b)
      for (u32 q = pl->soff[dd][b]; q < pl->soff[dd][b + 1]; ++q) sp[pl->slot[dd][q]] = pl->bprim[b];
    g->d_bnd_mem[m] = upload(sp, (size_t)(pl->total[dd] ? pl->total[dd] : 1) * sizeof(u32));
    g->d_sprim[dd] = (const u32 *)g->d_bnd_mem[m]; ++m;
    free(sp);
  }
  return m;
}

static void split_groups(struct gs_data *g, const struct xg_topo *t)
{
  /* interior CSR = groups (>= 2 members) whose primary is not a boundary
     primary; boundary CSR = all members of each boundary primary's group */
  const struct xg_plan *pl = &g->plan;
  const u32 nb = pl->nb;
  int any_flag = g->hall.nunf != 0;
  struct xg_hmap in; memset(&in, 0, sizeof in);
  /* two chunked passes over the label groups (count, prefix, fill): the output
     keeps the serial order and the chunks spread over the host cores */
  enum { NCH = 256 };
  size_t cg[NCH], ci[NCH], cf[NCH], cb[NCH], cbi[NCH];
  const int nthr = xg_setup_threads();
  (void)nthr;
#define GLO(c) ((size_t)t->ngrp * (size_t)(c) / NCH)
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthr)
#endif
  for (int c = 0; c < NCH; ++c) {
    size_t ng_ = 0, ni_ = 0, nf_ = 0, bni_ = 0;
    /* first boundary primary whose group lies in this chunk (bgrp ascends) */
    u32 lo = 0, hi = nb;
    while (lo < hi) { u32 mid = lo + (hi - lo) / 2; if (pl->bgrp[mid] < GLO(c)) lo = mid + 1; else hi = mid; }
    u32 bb = lo;
    for (size_t gi = GLO(c); gi < GLO(c + 1); ++gi) {
      u32 a = t->gstart[gi], e = t->gstart[gi + 1];
      if (bb < nb && pl->bgrp[bb] == gi) { ++bb; bni_ += e - a; continue; }
      if (e - a >= 2 || t->flag[a]) { ++ng_; ni_ += e - a; }   /* a flagged single is a one-member group */
      if (e - a < 2 && t->flag[a]) ++nf_;
    }
    cg[c] = ng_; ci[c] = ni_; cf[c] = nf_; cb[c] = lo; cbi[c] = bni_;
  }
  size_t ng = 0, ni = 0, nfs = 0, bni = 0;
  for (int c = 0; c < NCH; ++c) {
    size_t v;
    v = cg[c]; cg[c] = ng; ng += v;
    v = ci[c]; ci[c] = ni; ni += v;
    v = cf[c]; cf[c] = nfs; nfs += v;
    v = cbi[c]; cbi[c] = bni; bni += v;
  }
  in.ng = (u32)ng; in.ni = (u32)ni;
  in.off = xnew(u32, ng + 1);
  in.idx = xnew(u32, ni ? ni : 1);
  if (any_flag) in.nunf = xnew(u32, ng ? ng : 1);
  u32 *bgoff = xnew(u32, (size_t)nb + 1), *bgidx = xnew(u32, bni ? bni : 1), *bnunf = xnew(u32, nb ? nb : 1);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthr)
#endif
  for (int c = 0; c < NCH; ++c) {
    u32 b = (u32)cb[c]; size_t og = cg[c], oi = ci[c], ob = cbi[c];
    for (size_t gi = GLO(c); gi < GLO(c + 1); ++gi) {
      u32 a = t->gstart[gi], e = t->gstart[gi + 1], unf = e - a;
      if (any_flag) { unf = 0; for (u32 j = a; j < e && !t->flag[j]; ++j) ++unf; }
      if (b < nb && pl->bgrp[b] == gi) {
        bgoff[b] = (u32)ob; bnunf[b] = unf;
        memcpy(bgidx + ob, t->idx + a, (size_t)(e - a) * sizeof(u32));
        ob += e - a; ++b;
        continue;
      }
      if (e - a >= 2 || t->flag[a]) {
        in.off[og] = (u32)oi;
        if (in.nunf) in.nunf[og] = unf;
        memcpy(in.idx + oi, t->idx + a, (size_t)(e - a) * sizeof(u32));
        oi += e - a; ++og;
      }
    }
  }
#undef GLO
  in.off[ng] = (u32)ni;
  bgoff[nb] = (u32)bni;
  g->nfs = (u32)nfs;
  /* symmetric maps take the sliced-ELL layout (EXAGS_LAYOUT=csr keeps CSR) */
  struct xg_sell sell; memset(&sell, 0, sizeof sell);
  const char *lay = getenv("EXAGS_LAYOUT");
  int use_sell = in.ng && !(lay && strcmp(lay, "csr") == 0);
  if (use_sell) {
    xg_sell_build(&sell, &in);
    g->sell_idx_entries = sell.nidx;
  }
  if (use_sell && !sell.rest.ng) {
    /* keep what the pipelined host path needs to schedule windows by chunk */
    g->nwin = sell.nwin;
    g->win_min = sell.win_min; g->win_max = sell.win_max;
    sell.win_min = sell.win_max = 0;
    if (nb && bni) { g->h_bidx = xnew(u32, bni); memcpy(g->h_bidx, bgidx, bni * sizeof(u32)); g->h_nbidx = bni; }
  }
  if (g->on_device && use_sell) {
    upload_csr(&g->d_int, &sell.rest);
        g->d_sell.nblock = sell.nblock;
    g->d_sell_mem[0] = upload(sell.mask, (size_t)sell.nblock * 32 + 1);
    g->d_sell_mem[1] = upload(sell.idx, (sell.nidx ? sell.nidx : 1) * sizeof(u32));
    g->d_sell.mask = (const unsigned char *)g->d_sell_mem[0]; g->d_sell.idx = (const u32 *)g->d_sell_mem[1];
    if (sell.fmask) {
      g->d_sell_mem[2] = upload(sell.fmask, (size_t)sell.nblock * 32 + 1);
      g->d_sell.fmask = (const unsigned char *)g->d_sell_mem[2];
    }
  }
  if (g->on_device) {
    if (!use_sell) upload_csr(&g->d_int, &in);
    if (nb) {
      int m = 0;
      struct xcu_bnd *d = &g->d_bnd;
      d->nb = nb;
#define UP(dst, src, cnt) do { g->d_bnd_mem[m] = upload((src), (size_t)(cnt) * sizeof(*(src))); (dst) = g->d_bnd_mem[m]; ++m; } while (0)
      UP(d->goff, bgoff, nb + 1);
      UP(d->gidx, bgidx, bni ? bni : 1);
      UP(d->nunf, bnunf, nb);
      m = upload_boundary_plan(g, m);
      const char *bf = getenv("EXAGS_BND_FAST");
      if (!any_flag && !(bf && atoi(bf) == 0)) {
        /* boundary groups with >= 2 local members, sliced like the interior */
        struct xg_hmap bh; memset(&bh, 0, sizeof bh);
        for (u32 b = 0; b < nb; ++b) if (bgoff[b + 1] - bgoff[b] >= 2) { ++bh.ng; bh.ni += bgoff[b + 1] - bgoff[b]; }
        bh.off = xnew(u32, (size_t)bh.ng + 1); bh.idx = xnew(u32, bh.ni ? bh.ni : 1);
        { u32 og = 0, oi = 0;
          for (u32 b = 0; b < nb; ++b) {
            u32 sz = bgoff[b + 1] - bgoff[b];
            if (sz < 2) continue;
            bh.off[og++] = oi; memcpy(bh.idx + oi, bgidx + bgoff[b], (size_t)sz * sizeof(u32)); oi += sz;
          }
          bh.off[bh.ng] = oi; }
        struct xg_sell bs; memset(&bs, 0, sizeof bs);
        if (bh.ng) {
          xg_sell_build(&bs, &bh);
          upload_csr(&g->d_brest, &bs.rest);
          g->d_bsell.nblock = bs.nblock;
          if (bs.nblock) {
            UP(g->d_bsell.mask, bs.mask, (size_t)bs.nblock * 32 + 1);
            UP(g->d_bsell.idx, bs.idx, bs.nidx ? bs.nidx : 1);
          }
        }
        xg_sell_free(&bs); xg_hmap_free(&bh);
        m = upload_send_primaries(g, m);
        g->bnd_fast = 1;
      }
#undef UP
    }
    xcu_device_sync();
  } else {
    g->d_int.c.ng = in.ng; g->d_int.c.ni = in.ni;
  }
  g->handle_bytes += ((size_t)ng + 1 + ni - nfs + nb + 1 + bni + nb) * sizeof(u32);   /* ng, ni count the flagged singles too */
  xg_sell_free(&sell);
  xg_hmap_free(&in); free(bgoff); free(bgidx); free(bnunf);
}

/* time the exchange step alone, the way dry_run_time does
   (src/gs.upc:1094-1111): 2 warm-ups, 10 timed, mean; max over ranks */
static double time_exchange(struct gs_data *g, int method)
{
  struct xg_world *w = g->w;
  int save = g->method;
  g->method = method;
  const struct xg_plan *pl = &g->plan;
  void *s = xg_world_stream(w, 1);
  size_t nsend = method == gs_all_reduce ? (size_t)pl->total_shared : pl->total[1];
  size_t nrecv = method == gs_all_reduce ? 0 : pl->total[0];
  if (method == gs_all_reduce && xg_world_ipc(w)) nrecv = nsend;
  void *sb = xg_world_scratch(w, 0, nsend * 8);
  void *rb = xg_world_scratch(w, method == gs_all_reduce && xg_world_ipc(w) ? 3 : 1, nrecv * 8);
  xcu_memset(sb, 0, nsend * 8, s);
  double t0 = 0, t1 = 0;
  const int push = xg_world_push(w) && method != gs_all_reduce;
  if (xg_world_ipc(w) && !push) sb = xg_ipc_buffer(w, (size_t)pl->max_total * 8);
  for (int it = 0; it < 12; ++it) {
    if (it == 2) { xcu_stream_sync(s); xg_host_barrier(w); exags_comm_time(&t0); }
    if (push) {
      /* the transport alone: copy each neighbour's segment into its mailbox,
         raise the flags, wait for mine */
      push_prepare(g, (size_t)pl->max_total * 8);
      struct xg_mbox *mb = &g->mbox;
      const u32 M = g->push_nnbmax, epoch = ++mb->epoch, par = epoch & 1u;
      const u64 *t64 = (const u64 *)mb->d_tab;
      struct xcu_push pp; memset(&pp, 0, sizeof pp);
      pp.sig = (const unsigned long long *)(t64 + 4 * (size_t)M);
      pp.nsig = pl->nsym; pp.epoch_imm = epoch;          /* copies go by the copy engine: parity known here */
      pp.depoch = (u32 *)((char *)mb->mine + XG_MBOX_EPOCH);
      for (u32 k = 0; k < pl->nnb[1]; ++k)
        xcu_d2d((char *)mb->peer[pl->nb_rank[1][k]] + XG_MBOX_HDR + (size_t)par * mb->data_bytes + (size_t)pl->nb_peer_start[1][k] * 8,
                (const char *)sb + (size_t)pl->nb_start[1][k] * 8, (size_t)pl->nb_size[1][k] * 8, s);
      xcu_push_signal(s, &pp);
      xcu_push_wait(s, (const u32 *)mb->mine, (const u32 *)(t64 + 4 * (size_t)M + pl->nsym) + 2 * (size_t)M,
                    pl->nsym, 0, epoch, push_timeout(), xg_world_push_status(w));
    }
    else if (xg_world_ipc(w)) {
      if (method == gs_all_reduce) exchange_ipc_allreduce(g, rb, nsend, XD_F64, XO_ADD, s);
      else exchange_ipc_pairwise(g, 1, XD_F64, 0, rb, s);
    }
    else if (method == gs_all_reduce) xg_nccl_allreduce(w, sb, nsend, XD_F64, XO_ADD, s);
    else exchange_pairwise(g, 1, XD_F64, 0, sb, rb, s);
  }
  xcu_stream_sync(s);
  exags_comm_time(&t1);
  double t = (t1 - t0) / 10, buf;
  exags_comm_allreduce(g->comm, gs_double, gs_max, &t, 1, &buf);
  g->method = save;
  return t;
}

/* gs_setup's label sort and group discovery run on the device whenever there
   is one (EXAGS_SETUP=host keeps the host builders of topo.c; both produce the
   same arrays).  With one rank the maps are built there too and never visit
   the host; with several ranks the topology comes back for the shared-label
   discovery and the exchange plan, which are host code.                      */
static int setup_on_device(const struct gs_data *g)
{
  const char *e = getenv("EXAGS_SETUP");
  return g->on_device && !(e && strcmp(e, "host") == 0);
}

static i64 label_at(const void *id, int idbytes, size_t i)
{
  return idbytes == 8 ? ((const i64 *)id)[i] : (i64)((const int *)id)[i];
}

static void topo_any(struct gs_data *g, struct xg_topo *topo, const void *id, int idbytes, u32 n)
{
  if (setup_on_device(g)) {
    struct xcu_topo dt;
    xcu_setup_begin();
    xcu_setup_topo(id, idbytes, n, &dt);
    xcu_setup_topo_to_host(&dt, topo);
    xcu_setup_topo_free(&dt);
    xcu_setup_end();
  } else if (idbytes == 8) xg_topo_build(topo, (const i64 *)id, n);
  else {
    i64 *wide = xnew(i64, n ? n : 1);
    for (u32 i = 0; i < n; ++i) wide[i] = ((const int *)id)[i];
    xg_topo_build(topo, wide, n);
    free(wide);
  }
}

/* the sliced layout of the interior map `in`, as the handle's d_sell / d_int,
   plus what the pipelined host-pointer path keeps on the host               */
static void adopt_interior(struct gs_data *g, const struct xcu_dmap *in, const struct xcu_dmap *bnd)
{
  struct xcu_dsell ds;
  if (xcu_setup_sell(in, &ds)) {
    /* beyond the sliced layout's size limits (2^30 member indices or 2^23
       blocks): the thread-per-group CSR kernel (k_fused) serves the whole map */
    struct xcu_dmap own;
    xcu_dmap_clone(in, &own);
    g->d_int.c.ng = own.ng; g->d_int.c.ni = own.ni;
    g->d_int.off = own.off; g->d_int.idx = own.idx; g->d_int.nunf = own.nunf;
    g->d_int.c.off = own.off; g->d_int.c.idx = own.idx; g->d_int.c.nunf = own.nunf;
    return;
  }
  g->d_sell.nblock = ds.nblock;
  g->d_sell.mask = ds.mask; g->d_sell.idx = ds.idx; g->d_sell.fmask = ds.fmask;
  g->d_sell_mem[0] = ds.mask; g->d_sell_mem[1] = ds.idx; g->d_sell_mem[2] = ds.fmask;
  g->sell_idx_entries = ds.nidx;
  g->d_int.c.ng = ds.rest.ng; g->d_int.c.ni = ds.rest.ni;
  g->d_int.off = ds.rest.off; g->d_int.idx = ds.rest.idx; g->d_int.nunf = ds.rest.nunf;
  g->d_int.c.off = ds.rest.off; g->d_int.c.idx = ds.rest.idx; g->d_int.c.nunf = ds.rest.nunf;
  if (ds.nwin && !ds.rest.ng) {
    /* what the pipelined host-pointer path needs to schedule windows by chunk */
    g->nwin = ds.nwin;
    g->win_min = xnew(u32, ds.nwin); g->win_max = xnew(u32, ds.nwin);
    xcu_d2h(g->win_min, ds.win_min, (size_t)ds.nwin * sizeof(u32), 0);
    xcu_d2h(g->win_max, ds.win_max, (size_t)ds.nwin * sizeof(u32), 0);
    if (bnd && bnd->ng && bnd->ni) {
      g->h_bidx = xnew(u32, bnd->ni); g->h_nbidx = bnd->ni;
      xcu_d2h(g->h_bidx, bnd->idx, (size_t)bnd->ni * sizeof(u32), 0);
    }
    xcu_device_sync();
  }
  xcu_free(ds.win_min); xcu_free(ds.win_max);
}

/* one rank, device present: maps built in HBM (setup_cuda.cu)               */
static void setup_device_np1(struct gs_data *g, const void *id, int idbytes, u32 n, int prof)
{
  double tp[6]; int ntp = 0;
#define TICK() do { if (prof) { xcu_device_sync(); exags_comm_time(&tp[ntp++]); } } while (0)
  struct xcu_topo dt;
  struct xcu_dmap in;
  u32 *d_fs = 0;
  TICK();
  xcu_setup_begin();
  xcu_setup_topo(id, idbytes, n, &dt);
  TICK();
  xcu_setup_local(&dt, &g->d_hall, &in, &g->d_hfp, &g->nfp, &d_fs, &g->nfs, &g->sectors_f64);
  xcu_setup_topo_free(&dt);
  TICK();
  g->dev_built = 1;
  xcu_free(d_fs);
  g->hall.ng = g->d_hall.ng; g->hall.ni = g->d_hall.ni;      /* counts now, arrays on demand */
  /* with one rank the interior map is hall itself (plus the flagged singles,
     if any); hall stays with the handle for the introspection calls */
  adopt_interior(g, in.off ? &in : &g->d_hall, 0);
  if (in.off) xcu_dmap_free(&in);
  TICK();
  xcu_setup_end();
  g->handle_bytes = sizeof *g + ((size_t)g->d_hall.ng + 1 + g->d_hall.ni + g->nfs + 1) * sizeof(u32);
  TICK();
  if (prof)
    printf("gs_setup profile (device, n=%u): label sort+groups %.3f s, local map+sectors %.3f s, "
           "sliced layout %.3f s, finish %.3f s\n", n, tp[1] - tp[0], tp[2] - tp[1], tp[3] - tp[2], tp[4] - tp[3]);
#undef TICK
}

/* several ranks, device present: the exchange plan (host) has named the
   boundary groups; the interior/boundary split and both sliced layouts are
   built in HBM from the device topology.  Same device structures as
   split_groups() below.                                                     */
static int upload_boundary_plan(struct gs_data *g, int m);
static int upload_send_primaries(struct gs_data *g, int m);
static void split_groups_device(struct gs_data *g, const struct xcu_topo *dt)
{
  const struct xg_plan *pl = &g->plan;
  const u32 nb = pl->nb;
  struct xcu_dmap in, bnd, bnd2;
  u32 *d_fs = 0;
  xcu_setup_split(dt, pl->bgrp, nb, &g->d_hall, &in, &bnd, &bnd2, &g->d_hfp, &g->nfp, &d_fs, &g->nfs,
                  &g->sectors_f64);
  g->dev_built = 1;
  xcu_free(d_fs);
  g->hall.ng = g->d_hall.ng; g->hall.ni = g->d_hall.ni;
  const u32 ing = in.ng, ini = in.ni, bni = bnd.ni;
  adopt_interior(g, &in, &bnd);
  xcu_dmap_free(&in);
  if (nb) {
    int m = 0;
    struct xcu_bnd *d = &g->d_bnd;
    d->nb = nb;
    g->d_bnd_mem[m] = bnd.off; d->goff = bnd.off; ++m;
    g->d_bnd_mem[m] = bnd.idx; d->gidx = bnd.idx; ++m;
    g->d_bnd_mem[m] = bnd.nunf; d->nunf = bnd.nunf; ++m;
    m = upload_boundary_plan(g, m);
    const char *bf = getenv("EXAGS_BND_FAST");
    if (!dt->any_flag && !(bf && atoi(bf) == 0)) {
      if (bnd2.ng) {
        struct xcu_dsell bs;
        if (xcu_setup_sell(&bnd2, &bs)) xfail("gs_setup: boundary map too large for the sliced layout");
        g->d_brest.c.ng = bs.rest.ng; g->d_brest.c.ni = bs.rest.ni;
        g->d_brest.off = bs.rest.off; g->d_brest.idx = bs.rest.idx; g->d_brest.nunf = bs.rest.nunf;
        g->d_brest.c.off = bs.rest.off; g->d_brest.c.idx = bs.rest.idx; g->d_brest.c.nunf = bs.rest.nunf;
        g->d_bsell.nblock = bs.nblock;
        g->d_bsell.mask = bs.mask; g->d_bsell.idx = bs.idx;
        g->d_bnd_mem[m++] = bs.mask; g->d_bnd_mem[m++] = bs.idx;
        xcu_free(bs.win_min); xcu_free(bs.win_max); xcu_free(bs.fmask);
      }
      m = upload_send_primaries(g, m);
      g->bnd_fast = 1;
    }
  } else xcu_dmap_free(&bnd);
  xcu_dmap_free(&bnd2);
  xcu_device_sync();
  g->handle_bytes += ((size_t)ing + 1 + ini - g->nfs + nb + 1 + bni + nb) * sizeof(u32);
}

/* host copies of a device-built handle's group map, for the introspection calls */
static void fetch_host_maps(struct gs_data *g)
{
  if (!g->dev_built || g->hall.off) return;
  xcu_dmap_to_host(&g->d_hall, &g->hall);
  g->hfp = xnew(u32, g->nfp ? g->nfp : 1);
  if (g->nfp) { xcu_d2h(g->hfp, g->d_hfp, (size_t)g->nfp * sizeof(u32), 0); xcu_device_sync(); }
}

static struct gs_data *setup_core(const void *id_in, int idbytes, u32 n, const comm_ptr comm, int unique,
                                  int method, int verbose)
{
  if (!comm) xfail("gs_setup: null communicator");
  if (method < gs_auto || method > gs_all_reduce) xfail("gs_setup: bad method %d", method);
  struct gs_data *g = xnew0(struct gs_data, 1);
  exags_comm_dup(&g->comm, comm);
  struct xg_world *w = (struct xg_world *)comm->priv;
  g->w = w; g->n = n;
  g->on_device = xg_world_has_cuda(w);
  if (!g->on_device) {
    const char *ho = getenv("EXAGS_HOST_SETUP_ONLY");
    if (!(ho && atoi(ho)))
      xfail("gs_setup: no CUDA device available; this library has no CPU path "
            "(set EXAGS_HOST_SETUP_ONLY=1 only to inspect the host-side maps)");
  }
  const int np = comm->np;
  /* labels may live in HBM already (extension); the passes that read them on
     the host get a staged copy */
  void *id_staged = 0;
  if (g->on_device && n && xcu_is_device_ptr(id_in) && (unique || !setup_on_device(g))) {
    id_staged = xmalloc_((size_t)n * (size_t)idbytes, __FILE__, __LINE__);
    xcu_d2h(id_staged, id_in, (size_t)n * (size_t)idbytes, 0);
    xcu_device_sync();
    id_in = id_staged;
  }
  const void *id = id_in;
  i64 *id_own = 0;
  struct xg_topo topo;
  struct xg_shared sh;
  const char *prof_env = getenv("EXAGS_SETUP_PROFILE");
  const int prof = prof_env && atoi(prof_env) && comm->id == 0;
  const char *lay = getenv("EXAGS_LAYOUT");
  const int dev_maps = np == 1 && setup_on_device(g) && !(lay && strcmp(lay, "csr") == 0);
  double tp[8]; int ntp = 0;
#define TICK() do { if (prof) exags_comm_time(&tp[ntp++]); } while (0)
  TICK();
  if (unique) {
    /* same outcome as gs_unique followed by a plain setup (src/gs.h) */
    topo_any(g, &topo, id, idbytes, n);
    xg_shared_discover(&sh, &topo, w);
    unsigned char *nf = xnew(unsigned char, n ? n : 1);
    xg_unique_flags(nf, n, &topo, &sh, comm->id);
    id_own = xnew(i64, n ? n : 1);
    for (u32 i = 0; i < n; ++i) {
      i64 v = label_at(id_in, idbytes, i);
      i64 a = v < 0 ? -v : v;
      id_own[i] = nf[i] ? -a : a;
    }
    free(nf);
    xg_topo_free(&topo); xg_shared_free(&sh);
    id = id_own; idbytes = 8;
  }
  if (dev_maps) {
    setup_device_np1(g, id, idbytes, n, prof);
    free(id_own); free(id_staged);
    g->method = method == gs_all_reduce ? gs_all_reduce : gs_pairwise;
    if (verbose && comm->id == 0) {
      printf("gs_setup: %ld unique labels shared\n", 0L);
      printf("   handle bytes (avg, min, max): %g %u %u\n", (double)g->handle_bytes,
             (unsigned)g->handle_bytes, (unsigned)g->handle_bytes);
      printf("   buffer bytes (avg, min, max): %g %u %u\n", 0.0, 0u, 0u);
    }
    fflush(stdout);
    return g;
  }
  const int dev_split = np > 1 && setup_on_device(g) && !(lay && strcmp(lay, "csr") == 0);
  if (dev_split) {
    /* topology on the device, and it stays there: the labels another rank may
       hold are picked in HBM, only those candidates come to the host for the
       rendezvous on id % np (shared_ids, src/gs.upc:190-244) and the exchange
       plan, and the plan's boundary primaries are matched to their groups on
       the device again */
    struct xcu_topo dt;
    struct xg_cand cand;
    memset(&topo, 0, sizeof topo);
    xcu_setup_begin();
    xcu_setup_topo(id, idbytes, n, &dt);
    TICK();
    xcu_setup_candidates(&dt, w, &cand);
    xg_shared_pair(&sh, &cand, w);
    xg_cand_free(&cand);
    TICK();
    g->handle_bytes = sizeof *g;
    xg_plan_build(&g->plan, &sh, 0);
    xcu_setup_bgrp(&dt, g->plan.bprim, g->plan.nb, g->plan.bgrp);
    for (u32 b = 0; b < g->plan.nb; ++b)
      if (g->plan.bgrp[b] == XG_NONE)
        xfail("gs_setup: shared label's primary index %u is not a local primary", g->plan.bprim[b]);
    TICK();
    split_groups_device(g, &dt);
    xcu_setup_topo_free(&dt);
    xcu_setup_end();
    TICK();
    if (prof)
      printf("gs_setup profile (rank 0, n=%u, device): label sort+groups %.3f s, shared discovery %.3f s, "
             "exchange plan %.3f s, split + sliced layouts %.3f s\n", n,
             tp[1] - tp[0], tp[2] - tp[1], tp[3] - tp[2], tp[4] - tp[3]);
  } else {
  topo_any(g, &topo, id, idbytes, n);
  xg_shared_discover(&sh, &topo, w);

  TICK();
  xg_topo_local_map(&topo, 1, &g->hall);
  g->hfp = xg_topo_flagged_primaries(&topo, 0, &g->nfp);
  TICK();
  g->sectors_f64 = count_sectors(&g->hall, 8);
  g->handle_bytes = sizeof *g;
  TICK();
  xg_plan_build(&g->plan, &sh, &topo);
  TICK();
  split_groups(g, &topo);
  TICK();
  }
  if (prof && !dev_split)
    printf("gs_setup profile (rank 0, n=%u): label sort+groups(+shared discovery) %.3f s, local map %.3f s, "
           "sector count %.3f s, exchange plan %.3f s, split + sliced layout + upload %.3f s\n", n,
           tp[1] - tp[0], tp[2] - tp[1], tp[3] - tp[2], tp[4] - tp[3], tp[5] - tp[4]);
#undef TICK
  /* the NCCL communicator is created here, collectively, never lazily inside
     a ncclGroupStart/End bracket */
  if (np > 1) xg_plan_exchange_offsets(&g->plan, w);
  if (np > 1 && g->on_device && !xg_world_ipc(w) &&
      (!xg_world_push(w) || method == gs_all_reduce || method == gs_auto))
    (void)xg_world_nccl(w);       /* the push transport needs NCCL only for the all-reduce method */
  xg_shared_free(&sh);
  xg_topo_free(&topo);
  free(id_own); free(id_staged);

  if (verbose && comm->id == 0)
    printf("gs_setup: %ld unique labels shared\n", (long)g->plan.total_shared);

  /* method: crystal-router requests run the direct pairwise exchange -- on
     one NVSwitch domain every peer is one hop away, so log-P routing has
     nothing to save */
  g->method = method == gs_all_reduce ? gs_all_reduce : gs_pairwise;
  if (method == gs_auto && np > 1 && g->on_device) {
    double tp = time_exchange(g, gs_pairwise);
    const char *name = "pairwise";
    if (comm->id == 0) printf("   pairwise times (avg, min, max): %g %g %g\n", tp, tp, tp);
    if (g->plan.total_shared < 100000) {    /* src/gs.upc:1146 */
      double ta = time_exchange(g, gs_all_reduce);
      if (comm->id == 0) printf("   all reduce                    : %g %g %g\n", ta, ta, ta);
      if (ta < tp) { g->method = gs_all_reduce; name = "allreduce"; }
    }
    if (comm->id == 0) printf("   used all_to_all method: %s\n", name);
  }
  if (g->method == gs_all_reduce && np > 1 && g->plan.total_shared > 0xFFFFFFFFull)
    xfail("gs_setup: too many shared labels for the all-reduce method");

  if (verbose) {
    double avg[2], tmp[2];
    int mn[2], mx[2], ti[2];
    size_t bufvals = g->method == gs_all_reduce ? 2 * (size_t)g->plan.total_shared
                                                : (size_t)g->plan.total[0] + g->plan.total[1];
    avg[0] = (double)g->handle_bytes / np; avg[1] = 8.0 * (double)bufvals / np;
    mn[0] = mx[0] = (int)g->handle_bytes; mn[1] = mx[1] = (int)(8 * bufvals);
    exags_comm_allreduce(g->comm, gs_double, gs_add, avg, 2, tmp);
    exags_comm_allreduce(g->comm, gs_int, gs_min, mn, 2, ti);
    exags_comm_allreduce(g->comm, gs_int, gs_max, mx, 2, ti);
    if (comm->id == 0) {
      printf("   handle bytes (avg, min, max): %g %u %u\n", avg[0], (unsigned)mn[0], (unsigned)mx[0]);
      printf("   buffer bytes (avg, min, max): %g %u %u\n", avg[1], (unsigned)mn[1], (unsigned)mx[1]);
    }
  }
  fflush(stdout);
  return g;
}

struct gs_data *gslib_gs_setup(const long long *id, exags_uint n, const comm_ptr comm,
                               int unique, int method, int verbose)
{
  return setup_core(id, 8, n, comm, unique, method, verbose);
}

struct gs_data *exags_gs_setup(const int *id, exags_uint n, const comm_ptr comm, int unique,
                               gs_method method, int verbose)
{
  return setup_core(id, 4, n, comm, unique, (int)method, verbose);
}

void exags_gs_free(struct gs_data *g)
{
  if (!g) return;
  trace_report(g);
  if (g->on_device) {
    xcu_device_sync();
    xcu_free(g->d_int.off); xcu_free(g->d_int.idx); xcu_free(g->d_int.nunf);
    for (int m = 0; m < 3; ++m) xcu_free(g->d_sell_mem[m]);
    for (int m = 0; m < 20; ++m) xcu_free(g->d_bnd_mem[m]);
    xcu_free(g->d_brest.off); xcu_free(g->d_brest.idx); xcu_free(g->d_brest.nunf);
    xg_mbox_free(g->w, &g->mbox);
    if (g->dev_built) { xcu_dmap_free(&g->d_hall); xcu_free(g->d_hfp); }
    for (unsigned k = 0; g->pipe_ev && k < 2 * g->pipe_chunks; ++k) xcu_event_destroy(g->pipe_ev[k]);
    for (int k = 0; k < XG_RING; ++k) {
      xcu_host_free(g->ring_in[k]); xcu_host_free(g->ring_out[k]);
      xcu_event_destroy(g->ring_ev_in[k]); xcu_event_destroy(g->ring_ev_out[k]);
    }
  }
  xg_hmap_free(&g->hall);
  free(g->hfp); free(g->win_min); free(g->win_max); free(g->pipe_run); free(g->pipe_out_wave); free(g->pipe_ev);
  free(g->h_bidx); free(g->pipe_bnd);
  xg_plan_free(&g->plan);
  exags_comm_free(&g->comm);
  free(g);
}
void gslib_gs_free(struct gs_data *g) { exags_gs_free(g); }

static void unique_core(i64 *id, u32 n, const comm_ptr comm)
{
  if (!comm) xfail("gs_unique: null communicator");
  struct xg_world *w = (struct xg_world *)comm->priv;
  struct xg_topo topo; struct xg_shared sh;
  struct gs_data probe; memset(&probe, 0, sizeof probe);
  probe.on_device = xg_world_has_cuda(w);
  topo_any(&probe, &topo, id, 8, n);
  xg_shared_discover(&sh, &topo, w);
  unsigned char *nf = xnew(unsigned char, n ? n : 1);
  xg_unique_flags(nf, n, &topo, &sh, comm->id);
  for (u32 i = 0; i < n; ++i)
    if (nf[i] && id[i] > 0) id[i] = -id[i];
  free(nf);
  xg_shared_free(&sh); xg_topo_free(&topo);
}

void gslib_gs_unique(long long *id, exags_uint n, const comm_ptr comm) { unique_core((i64 *)id, n, comm); }

void exags_gs_unique(int *id, exags_uint n, const comm_ptr comm)
{
  i64 *wide = xnew(i64, n ? n : 1);
  for (exags_uint i = 0; i < n; ++i) wide[i] = id[i];
  unique_core(wide, n, comm);
  for (exags_uint i = 0; i < n; ++i) id[i] = (int)wide[i];
  free(wide);
}

/* ------------------------------------------------------------------ */
/*  introspection                                                      */
/* ------------------------------------------------------------------ */
void exags_gs_info(const struct gs_data *g, unsigned long long what[16])
{
  memset(what, 0, 16 * sizeof what[0]);
  what[0] = g->n; what[1] = g->hall.ng; what[2] = g->hall.ni; what[3] = g->nfp;
  what[4] = g->plan.nnb[0] > g->plan.nnb[1] ? g->plan.nnb[0] : g->plan.nnb[1];
  what[5] = g->plan.total[1]; what[6] = g->plan.total[0];
  what[7] = g->plan.total_shared; what[8] = (unsigned long long)g->method;
  what[9] = g->sectors_f64; what[10] = g->plan.nb; what[11] = g->d_int.c.ng;
  what[12] = g->d_int.c.ni; what[13] = g->nfs; what[14] = g->handle_bytes;
  what[15] = g->sell_idx_entries;
}

size_t exags_gs_dump_map(const struct gs_data *g, int which, exags_uint *out, size_t cap)
{
  size_t o = 0;
  fetch_host_maps((struct gs_data *)g);
#define PUT(v) do { if (o < cap) out[o] = (v); ++o; } while (0)
  if (which == 0 || which == 1) {
    const struct xg_hmap *m = &g->hall;
    for (u32 gi = 0; gi < m->ng; ++gi) {
      u32 a = m->off[gi], e = m->off[gi + 1];
      u32 cnt = e - a;
      if (which == 0 && m->nunf) {
        /* unflagged members only; the head is the primary even when flagged,
           but then nothing follows and the group is dropped */
        u32 unf = m->nunf[gi];
        cnt = unf >= 1 ? unf : 1;
      }
      if (cnt < 2) continue;
      for (u32 j = 0; j < cnt; ++j) PUT(m->idx[a + j]);
      PUT(XG_NONE);
    }
    PUT(XG_NONE);
  } else if (which == 2) {
    for (u32 k = 0; k < g->nfp; ++k) PUT(g->hfp[k]);
    PUT(XG_NONE);
  } else if (which == 3 || which == 4) {
    const struct xg_plan *pl = &g->plan;
    int d = which - 3;
    for (u32 b = 0; b < pl->nb; ++b) {
      u32 a = pl->soff[d] ? pl->soff[d][b] : 0, e = pl->soff[d] ? pl->soff[d][b + 1] : 0;
      if (a == e) continue;
      PUT(pl->bprim[b]);
      for (u32 s = a; s < e; ++s) PUT(pl->slot[d][s]);
      PUT(XG_NONE);
    }
    PUT(XG_NONE);
  }
#undef PUT
  return o;
}

unsigned long long exags_kernel_launches(void) { return xcu_launch_count(); }

void *exags_device_malloc(size_t bytes) { xg_world_get(); return xcu_malloc(bytes); }
void exags_device_free(void *p) { xcu_free(p); }
void exags_memcpy_h2d(void *d, const void *s, size_t n) { xcu_h2d(d, s, n, 0); xcu_device_sync(); }
void exags_memcpy_d2h(void *d, const void *s, size_t n) { xcu_d2h(d, s, n, 0); xcu_device_sync(); }
void exags_device_sync(void) { xcu_device_sync(); }
void *exags_host_malloc_pinned(size_t bytes) { xg_world_get(); return xcu_host_alloc(bytes); }
void exags_host_free_pinned(void *p) { xcu_host_free(p); }

/* ------------------------------------------------------------------ */
/*  Fortran bindings (src/gs.upc:1309-1418)                            */
/* ------------------------------------------------------------------ */
static struct gs_data **fgs_info = 0;
static int fgs_max = 0, fgs_n = 0;

static int fgs_new(struct gs_data *g)
{
  if (fgs_n == fgs_max) {
    fgs_max += fgs_max / 2 + 1;
    fgs_info = xgrow(struct gs_data *, fgs_info, fgs_max);
  }
  fgs_info[fgs_n] = g;
  return fgs_n++;
}

static void fgs_check(int handle, int dom, int op, int with_parms, const char *func, unsigned line)
{
  if (with_parms) {
    if (dom < 1 || dom > 3) exags_fail(1, __FILE__, line, "%s: datatype %d not in valid range 1-3", func, dom);
    if (op < 1 || op > 4) exags_fail(1, __FILE__, line, "%s: op %d not in valid range 1-4", func, op);
  }
  if (handle < 0 || handle >= fgs_n || !fgs_info[handle])
    exags_fail(1, __FILE__, line, "%s: invalid handle", func);
}

/* Fortran dom 1..3 = {double, sint, slong}; slong is 32-bit in the default
   flavour and 64-bit in the Nek flavour (src/gs.upc:1352, src/types.h:60-76) */
static int fdom(int dom, int wide) { return dom == 1 ? gs_double : dom == 2 ? gs_int : (wide ? 4 : gs_int); }
static size_t fdom_bytes(int dom, int wide) { return dom == 1 ? 8 : dom == 2 ? 4 : (wide ? 8 : 4); }

#define FGS_BODY_SETUP(IDT, SETUP, METHOD)                                    \
  exags_comm_init();                                                          \
  *handle = fgs_new(SETUP((const IDT *)id, (exags_uint)*n, exags_glb_comm, 0, METHOD, 1))

void exags_fgs_setup(int *handle, const int id[], const int *n, const int *comm, const int *np)
{ (void)comm; (void)np; FGS_BODY_SETUP(int, exags_gs_setup, gs_auto); }
void exags_fgs_setup_pick(int *handle, const int id[], const int *n, const int *comm, const int *np, const int *method)
{ (void)comm; (void)np; FGS_BODY_SETUP(int, exags_gs_setup, (gs_method)*method); }
void fgslib_gs_setup_(int *handle, const long long id[], const int *n, const int *comm, const int *np)
{ (void)comm; (void)np; FGS_BODY_SETUP(long long, gslib_gs_setup, gs_auto); }
void fgslib_gs_setup_pick_(int *handle, const long long id[], const int *n, const int *comm, const int *np, const int *method)
{ (void)comm; (void)np; FGS_BODY_SETUP(long long, gslib_gs_setup, *method); }

static void fgs_op(int wide, const int *handle, void *u, const int *dom, const int *op, const int *transpose)
{
  fgs_check(*handle, *dom, *op, 1, "gs_op", __LINE__);
  gs_run(u, XM_PLAIN, 1, fdom(*dom, wide), *op - 1, *transpose != 0, fgs_info[*handle], 0, 1, "gs_op");
}
static void fgs_op_vec(int wide, const int *handle, void *u, const int *n, const int *dom, const int *op, const int *transpose)
{
  fgs_check(*handle, *dom, *op, 1, "gs_op_vec", __LINE__);
  gs_run(u, XM_VEC, (unsigned)*n, fdom(*dom, wide), *op - 1, *transpose != 0, fgs_info[*handle], 0, 1, "gs_op_vec");
}
static void fgs_op_many(int wide, const int *handle, void *u1, void *u2, void *u3, void *u4, void *u5, void *u6,
                        const int *n, const int *dom, const int *op, const int *transpose)
{
  void *uu[6] = { u1, u2, u3, u4, u5, u6 };
  fgs_check(*handle, *dom, *op, 1, "gs_op_many", __LINE__);
  if (*n < 0 || *n > 6) xfail("gs_op_many: %d arrays, at most 6", *n);
  gs_run((void *)uu, XM_MANY, (unsigned)*n, fdom(*dom, wide), *op - 1, *transpose != 0, fgs_info[*handle], 0, 1, "gs_op_many");
}
static void fgs_op_fields(int wide, const int *handle, void *u, const int *stride, const int *n,
                          const int *dom, const int *op, const int *transpose)
{
  fgs_check(*handle, *dom, *op, 1, "gs_op_fields", __LINE__);
  if (*n <= 0) return;
  void **p = xnew(void *, *n);
  size_t offset = (size_t)*stride * fdom_bytes(*dom, wide);
  for (int k = 0; k < *n; ++k) p[k] = (char *)u + offset * (size_t)k;
  gs_run((void *)p, XM_MANY, (unsigned)*n, fdom(*dom, wide), *op - 1, *transpose != 0, fgs_info[*handle], 0, 1, "gs_op_fields");
  free(p);
}
static void fgs_release(const int *handle)
{
  fgs_check(*handle, 0, 0, 0, "gs_free", __LINE__);
  exags_gs_free(fgs_info[*handle]);
  fgs_info[*handle] = 0;
}

void exags_fgs(const int *h, void *u, const int *d, const int *o, const int *t) { fgs_op(0, h, u, d, o, t); }
void exags_fgs_vec(const int *h, void *u, const int *n, const int *d, const int *o, const int *t) { fgs_op_vec(0, h, u, n, d, o, t); }
void exags_fgs_many(const int *h, void *u1, void *u2, void *u3, void *u4, void *u5, void *u6, const int *n, const int *d, const int *o, const int *t)
{ fgs_op_many(0, h, u1, u2, u3, u4, u5, u6, n, d, o, t); }
void exags_fgs_fields(const int *h, void *u, const int *s, const int *n, const int *d, const int *o, const int *t) { fgs_op_fields(0, h, u, s, n, d, o, t); }
void exags_fgs_free(const int *h) { fgs_release(h); }

void fgslib_gs_op_(const int *h, void *u, const int *d, const int *o, const int *t) { fgs_op(1, h, u, d, o, t); }
void fgslib_gs_op_vec_(const int *h, void *u, const int *n, const int *d, const int *o, const int *t) { fgs_op_vec(1, h, u, n, d, o, t); }
void fgslib_gs_op_many_(const int *h, void *u1, void *u2, void *u3, void *u4, void *u5, void *u6, const int *n, const int *d, const int *o, const int *t)
{ fgs_op_many(1, h, u1, u2, u3, u4, u5, u6, n, d, o, t); }
void fgslib_gs_op_fields_(const int *h, void *u, const int *s, const int *n, const int *d, const int *o, const int *t) { fgs_op_fields(1, h, u, s, n, d, o, t); }
void fgslib_gs_free_(const int *h) { fgs_release(h); }

/* default-flavour Fortran spellings: bare and with trailing underscore
   (FORTRAN_NAME, src/name.h:32-41) */
#define FALIAS(name, target) \
  extern __typeof__(target) name __attribute__((alias(#target)))
FALIAS(gs_setup_, exags_fgs_setup);
FALIAS(gs_setup_pick_, exags_fgs_setup_pick);
FALIAS(gs_op_, exags_fgs);
FALIAS(gs_op_vec_, exags_fgs_vec);
FALIAS(gs_op_many_, exags_fgs_many);
FALIAS(gs_op_fields_, exags_fgs_fields);
FALIAS(gs_free_, exags_fgs_free);
FALIAS(gs_setup_pick, exags_fgs_setup_pick);
FALIAS(gs_op, exags_fgs);
FALIAS(gs_op_vec, exags_fgs_vec);
FALIAS(gs_op_many, exags_fgs_many);
FALIAS(gs_op_fields, exags_fgs_fields);
