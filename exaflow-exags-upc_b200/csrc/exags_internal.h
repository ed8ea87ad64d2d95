/* exags_internal.h -- private declarations shared by the host C code and the
 * CUDA translation unit.  Nothing in here is ABI. */
#ifndef EXAGS_INTERNAL_H
#define EXAGS_INTERNAL_H

#include <stddef.h>
#include <stdint.h>

#define EXAGS_NO_INT_MACROS 1
#define EXAGS_FORTRAN_GS_OP 1   /* keep the name gs_op free for the Fortran entry point */
#include "../../include/exags.h"
#include "../../include/exags_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t  i64;

#define XG_NONE ((u32)0xFFFFFFFFu)   /* the reference's (uint)-1 sentinel */
#define XG_MAXV 8                    /* component pointers passed per launch */

/* internal storage type of a gs_dom value; both ABI flavours map onto it
   (gs_long and gs_long_long are both 64-bit on LP64, src/gs_defs.h:18-23) */
enum { XD_F64 = 0, XD_F32 = 1, XD_I32 = 2, XD_I64 = 3, XD_N = 4 };
enum { XO_ADD = 0, XO_MUL = 1, XO_MIN = 2, XO_MAX = 3, XO_BPR = 4, XO_N = 5 };
enum { XM_PLAIN = 0, XM_VEC = 1, XM_MANY = 2 };

static inline int xg_dom_of(int dom)      /* ABI dom -> storage type, -1 bad */
{
  static const int t[5] = { XD_F64, XD_F32, XD_I32, XD_I64, XD_I64 };
  return (dom >= 0 && dom < 5) ? t[dom] : -1;
}
static inline unsigned xg_dom_bytes(int xd)
{
  static const unsigned s[4] = { 8, 4, 4, 8 };
  return s[xd];
}

#define xfail(...) exags_fail(1, __FILE__, __LINE__, __VA_ARGS__)

/* ---- memory ------------------------------------------------------------ */
void *xmalloc_(size_t bytes, const char *file, unsigned line);
void *xcalloc_(size_t bytes, const char *file, unsigned line);
void *xrealloc_(void *p, size_t bytes, const char *file, unsigned line);
#define xnew(T, n)        ((T *)xmalloc_((size_t)(n) * sizeof(T), __FILE__, __LINE__))
#define xnew0(T, n)       ((T *)xcalloc_((size_t)(n) * sizeof(T), __FILE__, __LINE__))
#define xgrow(T, p, n)    ((T *)xrealloc_((p), (size_t)(n) * sizeof(T), __FILE__, __LINE__))

/* ---- sorting (sort.c) --------------------------------------------------- */
int xg_setup_threads(void);
/* Stable LSD radix sort of (key,val) pairs on the low `bits` bits of key.
   tmpk/tmpv are scratch of n entries; result ends up back in key/val.      */
void xsort_u64_u32(u64 *key, u32 *val, size_t n, unsigned bits,
                   u64 *tmpk, u32 *tmpv);
/* Stable sort of a permutation: perm is reordered so key[perm[.]] ascends.  */
void xsort_perm_u32(u32 *perm, const u32 *key, size_t n, u32 maxkey);
void xsort_perm_u64(u32 *perm, const u64 *key, size_t n, u64 maxkey);

/* ---- local topology (topo.c) -------------------------------------------- */
/* One row per nonzero local label, ordered by (primary, flag, index): the
   order the reference establishes in nonzero_ids(), src/gs.upc:87-112.      */
struct xg_topo {
  size_t nnz;         /* rows */
  u64   *id;          /* |label| of each group                [ngrp] */
  u32   *idx;         /* local index i                        [nnz] */
  unsigned char *flag;/* label was negative                   [nnz] */
  size_t ngrp;        /* distinct labels on this rank (singletons included) */
  u32   *gstart;      /* row offset of each label group       [ngrp+1] */
};
void xg_topo_build(struct xg_topo *t, const i64 *id, size_t n);
void xg_topo_free(struct xg_topo *t);

/* host CSR of gather/scatter groups (only groups with >= 2 members) plus the
   list of flagged primaries; replaces the three sentinel streams built by
   local_setup(), src/gs.upc:1210-1217.                                      */
struct xg_hmap {
  u32  ng, ni;
  u32 *off;           /* [ng+1] */
  u32 *idx;           /* [ni] members: primary first, unflagged before flagged */
  u32 *nunf;          /* [ng] unflagged members per group (NULL: all unflagged) */
};
void xg_hmap_free(struct xg_hmap *m);

/* Sliced-ELL view of a symmetric CSR map (after SELL-C-sigma, Kreutzer et
   al. 2014, C = 32), shaped for one warp per block and one CTA per window:

   block   32 lanes x 8 member slots = 256 indices (1 KB).  A lane's 8 slots
           are contiguous (two 128-bit loads) and hold whole groups back to
           back; bit j of the lane's start mask says "slot j begins a group".
           Unused slots hold XG_NONE.  Whatever the group sizes, every lane
           has 8 independent field loads in flight.
   window  XG_SELL_WB consecutive blocks = the consecutive groups (in primary
           order) that fill them, stably sorted by descending size and dealt
           row-major over the lanes, so lanes next to each other hold groups
           next to each other in the field and groups of all sizes that touch
           the same stretch of the field meet in the same CTA's L1.

   No per-group offsets remain: 4 bytes per member + 1 mask byte per lane.
   Groups with more than 8 members stay in a residual CSR.                   */
#define XG_SELL_WMAX 8
#ifndef XG_SELL_WB
#define XG_SELL_WB   8     /* blocks per window (a multiple of XG_SELL_CW)            */
#endif
#ifndef XG_SELL_CW
#define XG_SELL_CW   8     /* blocks per CTA of the sliced kernels = warps per CTA.
                              Measured: a window split over two CTAs loses the L1
                              reuse the chain order creates (profiles/r1_sweep10) */
#endif
/* resident CTAs per SM are quoted for 8-warp CTAs; scaled for other CTA sizes */
#define XG_MINB(m) (((m) * 8 + XG_SELL_CW - 1) / XG_SELL_CW)
struct xg_sell {
  u32  nblock;        /* multiple of XG_SELL_WB                              */
  unsigned char *mask;/* [nblock*32] start mask of each lane                 */
  unsigned char *fmask;/* [nblock*32] "slot holds a flagged member" mask, NULL for
                         symmetric maps (src/gs.upc:329-357: map_local[0] leaves
                         flagged members out, map_local[1] keeps them)        */
  u32 *idx;           /* [nblock*256]                                        */
  size_t nidx;
  u32  nwin;          /* windows = nblock / XG_SELL_WB                       */
  u32 *win_min, *win_max; /* [nwin] smallest / largest member index per window */
  struct xg_hmap rest;/* groups with more than XG_SELL_WMAX members          */
};
void xg_sell_build(struct xg_sell *s, const struct xg_hmap *m);
int  xg_sell_pairs_enabled(void);
void xg_sell_free(struct xg_sell *s);
void xg_topo_local_map(const struct xg_topo *t, int all, struct xg_hmap *m);
u32 *xg_topo_flagged_primaries(const struct xg_topo *t, int singles_only, u32 *count);

/* ---- world / communicator (comm.c) -------------------------------------- */
struct xg_world;
struct xg_world *xg_world_get(void);           /* creates on first use */
int   xg_world_rank(const struct xg_world *w);
int   xg_world_np(const struct xg_world *w);
int   xg_world_has_cuda(const struct xg_world *w);
void *xg_world_stream(struct xg_world *w, int which);   /* 0 main, 1 comm, 2 copy-in, 3 copy-out */
void *xg_world_event(struct xg_world *w, int which);
void *xg_world_nccl(struct xg_world *w);       /* ncclComm_t, created lazily */
/* grow-only device scratch areas owned by the world (the analogue of the
   reference's file-static buffer, src/gs.upc:35): 0 send, 1 recv, 2 stage   */
void *xg_world_scratch(struct xg_world *w, int which, size_t bytes);

/* exchange transport between ranks' GPUs: 0 = NCCL send/recv (default),
   1 = CUDA-IPC pulls (chosen when two ranks share a device, or EXAGS_EXCHANGE=ipc) */
int   xg_world_ipc(const struct xg_world *w);
/* collective: every rank's exported buffer holds at least `bytes` (same value
   on all ranks); returns this rank's buffer */
void *xg_ipc_buffer(struct xg_world *w, size_t bytes);
void *xg_ipc_peer(struct xg_world *w, int rank);   /* mapped pointer to rank's buffer */

/* Push transport (default between distinct NVLink-connected GPUs, or
   EXAGS_EXCHANGE=push): the boundary pack kernel stores each value straight
   into the receiving rank's mailbox over NVLink and the last CTA to finish
   raises a flag there; the receiver's stream waits on its own flags, then
   unpacks.  No NCCL call and no host round trip on the data path -- the
   B200 counterpart of pw_exec_sends/pw_exec_recvs (src/gs.upc:409-468),
   whose upc_memput + polled flag it restates with peer stores.  A mailbox
   belongs to one gs handle; two data areas alternate by call parity.       */
#define XG_MBOX_HDR 4096u            /* flags u32[np] at 0, CTA counter at 512, exchange number at 520 */
#define XG_MBOX_COUNTER 512u
#define XG_MBOX_EPOCH 520u
struct xg_mbox {
  void  *mine;                       /* header + 2 data areas                  */
  size_t data_bytes;                 /* bytes of one data area                 */
  void **peer;                       /* [np] mapped mailboxes of my neighbours */
  u32    epoch;                      /* exchanges done on this mailbox         */
  void  *d_tab;                      /* device tables built by the handle      */
};
int   xg_world_push(const struct xg_world *w);
int  *xg_world_push_status(struct xg_world *w);   /* pinned+mapped int, 0 = ok */
/* collective over the ranks of w: make every mailbox at least data_bytes per
   area (same value everywhere) and map the neighbours'; returns 1 when the
   mailbox was (re)created, in which case epoch restarts at 0                 */
int   xg_mbox_ensure(struct xg_world *w, struct xg_mbox *mb, const u32 *nbr, u32 nnbr, size_t data_bytes);
void  xg_mbox_free(struct xg_world *w, struct xg_mbox *mb);

/* host-side collectives over the rendezvous segment (np>1) */
void  xg_host_barrier(struct xg_world *w);
/* every rank contributes `bytes` bytes; out receives np*bytes, rank-major   */
void  xg_host_allgather(struct xg_world *w, const void *in, size_t bytes, void *out);
u64  *xg_host_bitmap_new(struct xg_world *w, size_t words);
void  xg_host_bitmap_or_others(struct xg_world *w, u64 *mine, size_t words, u64 *others);
/* rows of `rowbytes` bytes, already grouped by destination: sendcnt[p] rows
   go to rank p.  Returns a malloc'd array of received rows ordered by source
   rank; *nrecv rows, recvcnt[p] (caller array of np) from each source.      */
void *xg_host_alltoallv(struct xg_world *w, const void *rows, size_t rowbytes,
                        const size_t *sendcnt, size_t *recvcnt, size_t *nrecv);

/* ---- CUDA side (gs_cuda.cu), thin C ABI used by the host C code ---------- */
int    xcu_ptr_kind(const void *p);                 /* 0 pageable host, 1 pinned/registered host, 2 device */
int    xcu_host_register(void *p, size_t bytes);    /* 1 on success */
void   xcu_host_unregister(void *p);
int    xcu_device_count(void);
int    xcu_set_device(int dev);
void  *xcu_malloc(size_t bytes);
void   xcu_free(void *p);
void   xcu_h2d(void *dst, const void *src, size_t bytes, void *stream);
void   xcu_d2h(void *dst, const void *src, size_t bytes, void *stream);
void   xcu_d2d(void *dst, const void *src, size_t bytes, void *stream);
void   xcu_memset(void *dst, int byte, size_t bytes, void *stream);
int    xcu_is_device_ptr(const void *p);
void  *xcu_stream_create(int high_priority);
void   xcu_stream_destroy(void *s);
void   xcu_stream_sync(void *s);
void  *xcu_event_create(void);
void  *xcu_event_create_timed(void);
void   xcu_event_destroy(void *ev);
float  xcu_event_ms(void *a, void *b);
void   xcu_event_sync(void *ev);
int    xcu_event_done(void *ev);                    /* 1 when the event has completed */
void   xcu_event_record(void *ev, void *stream);
void   xcu_stream_wait_event(void *stream, void *ev);
void   xcu_device_sync(void);
void  *xcu_host_alloc(size_t bytes);
void   xcu_host_free(void *p);
unsigned long long xcu_launch_count(void);
void   xcu_ipc_export(void *handle64, void *devptr);
void  *xcu_ipc_open(const void *handle64);
void   xcu_ipc_close(void *p);
void  *xcu_ipc_try_open(const void *handle64);
void   xcu_device_uuid(int dev, char out16[16]);
int    xcu_peer_ok(int dev, const char uuid16[16]);   /* can dev store into the device with that uuid? */
void  *xcu_host_alloc_mapped(size_t bytes);

/* a field as the kernels see it */
struct xcu_field {
  void    *p[XG_MAXV];   /* plain/vec: p[0]; many: one pointer per component */
  unsigned vn;           /* components handled by this launch               */
  unsigned tuple;        /* vec: elements per tuple in memory (>= vn)       */
  unsigned k0;           /* vec: first component handled by this launch     */
  const void *w;         /* weights T[n] of the weighted gs (exags_cuda.h), plain
                            float/double fields only; NULL otherwise          */
};

/* device CSR */
struct xcu_csr {
  u32 ng, ni;
  const u32 *off, *idx, *nunf;   /* device pointers; nunf may be NULL */
};

/* device SELL map */
struct xcu_sell {
  u32 nblock;
  const unsigned char *mask;
  const u32 *idx;
  const unsigned char *fmask;   /* NULL: no flagged members */
};
void xcu_fused_sell(void *stream, const struct xcu_sell *m, const struct xcu_field *f,
                    int mode, int xdom, int xop, int transpose);
void xcu_gather_sell(void *stream, const struct xcu_sell *m, const struct xcu_field *f,
                     int mode, int xdom, int xop);

/* boundary description on the device (np>1) */
struct xcu_bnd {
  u32 nb;                        /* boundary primaries */
  const u32 *goff, *gidx, *nunf; /* members of each (>=1), unflagged count   */
  const u32 *soff[2], *slot[2];  /* pairwise buffer slots per direction      */
  const u32 *ord;                /* all-reduce slot per boundary primary     */
  const unsigned char *pflag;    /* primary flagged?                          */
};

/* push transport: where the pack kernel stores (device tables of one handle,
   one direction, one parity) and whom it signals                            */
struct xcu_push {
  const unsigned long long *base;  /* [2 parities][par_stride] peer data area of each neighbour, as an address */
  u32  par_stride;                 /* entries between the two parities of `base`    */
  const u32 *delta;                /* [neighbours] peer slot - my slot (mod 2^32)   */
  const unsigned char *snb;        /* [slot entries] neighbour number of each entry */
  const unsigned long long *sig;   /* [nsig] address of my flag in each neighbour   */
  u32  nsig;                       /* 0: do not signal (nor advance the epoch) from this launch */
  /* The exchange number lives in device memory (a word of my own mailbox
     header) so that a call can be captured in a CUDA graph and replayed: this
     exchange is number *depoch + 1; the launch that signals stores it back.
     epoch_imm != 0 (the transport timing of gs_setup, which copies with the
     copy engine and so needs the parity on the host) names the number instead
     and the signal kernel stores it into *depoch.                            */
  u32 *depoch;
  u32  epoch_imm;
  u32 *counter;                    /* CTA completion counter in my mailbox          */
  const u32 *start;                /* [nnb+1] first send slot of each neighbour (slot-ordered pack) */
  u32  nnb;
};
/* slot-ordered pack (symmetric maps): slot s of the send numbering carries
   u[sprim[s]], which already holds its group's local total; with pp != NULL
   the value goes straight into the neighbour's mailbox (consecutive slots of a
   neighbour are consecutive there, so the NVLink stores coalesce)            */
void xcu_pack_slots(void *stream, const u32 *sprim, u32 nslots, const struct xcu_field *f,
                    int mode, int xdom, void *sendbuf, unsigned bufvn, const struct xcu_push *pp);
void xcu_bnd_pack_push(void *stream, const struct xcu_bnd *b, const struct xcu_field *f,
                       int mode, int xdom, int xop, int transpose, unsigned bufvn,
                       const struct xcu_push *pp);
void xcu_push_signal(void *stream, const struct xcu_push *pp);
/* wait until the flags of `ranks` in my mailbox reach the current exchange
   number (*depoch, or epoch_imm when non-zero)                               */
void xcu_push_wait(void *stream, const u32 *flags, const u32 *ranks, u32 n, const u32 *depoch, u32 epoch_imm,
                   double timeout_s, int *status);

/* fused local gather(+init)+scatter over every group of m                  */
void xcu_fused(void *stream, const struct xcu_csr *m, const struct xcu_field *f,
               int mode, int xdom, int xop, int transpose);
/* out[list[j]] = identity(op) for the flagged primaries that head no group */
void xcu_init_list(void *stream, const u32 *list, u32 n, const struct xcu_field *f,
                   int mode, int xdom, int xop);
/* boundary: local gather (+init) -> u[primary] and -> sendbuf slots (dir) */
void xcu_bnd_pack(void *stream, const struct xcu_bnd *b, const struct xcu_field *f,
                  int mode, int xdom, int xop, int transpose, void *sendbuf,
                  unsigned bufvn, int allreduce);
/* boundary: fold recvbuf slots into u[primary], then scatter to members    */
/* depoch != NULL (push transport): recvbuf is the data area of parity 0 and
   the kernel adds (*depoch & 1) * par_bytes, the area of the current exchange */
void xcu_bnd_unpack(void *stream, const struct xcu_bnd *b, const struct xcu_field *f,
                    int mode, int xdom, int xop, int transpose, const void *recvbuf,
                    unsigned bufvn, int allreduce, const u32 *depoch, size_t par_bytes);
unsigned xcu_dot_parts(void);
void xcu_dot(void *stream, const double *v, const double *w, size_t n, double *scratch, double *out);
void xcu_fill_identity(void *stream, void *out, size_t n, int xdom, int xop);
void xcu_array_op(void *stream, void *out, const void *in, size_t n, int xdom, int xop);
/* generic CSR gather / scatter between two (possibly different) fields:
   used by the public gs_local.h entry points.                              */
void xcu_csr_gather(void *stream, const struct xcu_csr *m, const struct xcu_field *out,
                    int omode, const struct xcu_field *in, int imode,
                    int xdom, int xop);
void xcu_csr_scatter(void *stream, const struct xcu_csr *m, const struct xcu_field *out,
                     int omode, const struct xcu_field *in, int imode, int xdom);

/* ---- gs_setup on the device (setup_cuda.cu, setup_pipeline.h) ------------- */
/* The label sort, group discovery and map construction of gs_setup as device
   passes (replaces nonzero_ids / local_map / flagged_primaries_map,
   src/gs.upc:83-112,329-370, and this library's own host builders in topo.c,
   whose output it reproduces array for array).  All pointers below are device
   pointers.                                                                  */
struct xcu_topo {               /* struct xg_topo in HBM                       */
  size_t n, nnz, ngrp;
  u64 *id; u32 *idx; unsigned char *flag; u32 *gstart;
  u32 *rowgrp;                  /* [nnz] group of each row                      */
  int any_flag;
};
struct xcu_dmap {               /* struct xg_hmap in HBM                       */
  u32 ng, ni;
  u32 *off, *idx, *nunf;
};
struct xcu_dsell {              /* struct xg_sell in HBM                       */
  u32 nblock, nwin;
  size_t nidx;
  unsigned char *mask, *fmask;
  u32 *idx, *win_min, *win_max;
  struct xcu_dmap rest;
};
/* id: host or device array of n labels, idbytes 4 (int) or 8 (long long)     */
void xcu_setup_topo(const void *id, int idbytes, size_t n, struct xcu_topo *t);
void xcu_setup_topo_free(struct xcu_topo *t);
void xcu_setup_topo_to_host(const struct xcu_topo *dt, struct xg_topo *ht);
struct xg_cand; struct xg_world;
/* np > 1: candidate labels chosen on the device (collective over the ranks of w) */
void xcu_setup_candidates(const struct xcu_topo *t, struct xg_world *w, struct xg_cand *out);
void xcu_setup_bgrp(const struct xcu_topo *t, const u32 *bprim, u32 nb, u32 *bgrp);
/* hall: every group with >= 2 members; fp: all flagged primaries; fs: flagged
   primaries that head no multi-member group; sectors: distinct 32-byte sectors
   of a double field that hold a member of hall                               */
void xcu_setup_local(const struct xcu_topo *t, struct xcu_dmap *hall, struct xcu_dmap *in, u32 **fp, u32 *nfp,
                     u32 **fs, u32 *nfs, u64 *sectors);   /* in: hall + fs as one-member groups, or off == NULL: same as hall */
/* several ranks: bgrp = the nb boundary groups named by the exchange plan
   (host array, ascending).  in = interior groups with >= 2 members, bnd = every
   boundary group (all members, unflagged counts), bnd2 = boundary groups with
   >= 2 members (only for maps without flagged labels), fs excludes the boundary */
void xcu_setup_split(const struct xcu_topo *t, const u32 *bgrp_host, u32 nb, struct xcu_dmap *hall,
                     struct xcu_dmap *in, struct xcu_dmap *bnd, struct xcu_dmap *bnd2,
                     u32 **fp, u32 *nfp, u32 **fs, u32 *nfs, u64 *sectors);
int  xcu_setup_sell(const struct xcu_dmap *m, struct xcu_dsell *s);   /* 1: too large for the sliced layout */
void xcu_dmap_clone(const struct xcu_dmap *m, struct xcu_dmap *out);
void xcu_dmap_free(struct xcu_dmap *m);
void xcu_dmap_to_host(const struct xcu_dmap *m, struct xg_hmap *h);
void xcu_note_launches(unsigned long long k);
void xcu_setup_begin(void);   /* brackets the device passes of one gs_setup */
void xcu_setup_end(void);

#ifdef __cplusplus
}
#endif
#endif
