/* gs_plan.h -- host-side description of what a rank exchanges with its
 * neighbours (the counterpart of struct shared_id / pw_data / allreduce_data,
 * src/gs.upc:145-158,395-407,1001-1004). */
#ifndef EXAGS_GS_PLAN_H
#define EXAGS_GS_PLAN_H
#include "exags_internal.h"

#define XG_FLAGS_LOCAL  1u   /* this rank's copy of the label is flagged  */
#define XG_FLAGS_REMOTE 2u   /* the remote rank's copy is flagged         */

struct xg_shared_row {
  u64 id, ord;        /* label, global ordinal among shared labels        */
  u32 i, p, ri;       /* local primary index, remote rank, remote primary */
  unsigned flags;
};
struct xg_shared {
  size_t n;
  struct xg_shared_row *row;
  u64 total_shared;
};

struct xg_plan {
  u64  total_shared;
  u32  nb;            /* boundary primaries                                */
  u32 *bprim;         /* [nb] local index, ascending                       */
  u32 *bgrp;          /* [nb] topology group headed by that primary        */
  u32 *ord;           /* [nb] all-reduce slot                              */
  unsigned char *pflag; /* [nb] primary flagged                            */
  /* pairwise, per direction d (0: rows with remote unflagged = default recv,
     1: rows with local unflagged = default send; swapped under transpose) */
  u32  total[2];      /* slots                                             */
  u32 *soff[2];       /* [nb+1] slot list offsets per boundary primary     */
  u32 *slot[2];       /* [total] slots, ascending remote rank              */
  u32  nnb[2];        /* neighbour ranks                                   */
  u32 *nb_rank[2], *nb_start[2], *nb_size[2];
  /* for the pull (CUDA-IPC) transport: where my message starts in the
     neighbour's send numbering of the same direction, and the largest slot
     count of any rank/direction (uniform buffer sizing) */
  u32 *nb_peer_start[2];
  u64  max_total;
  /* for the push transport: neighbour number (index into nb_rank[d]) of each
     entry of slot[d], and the union of both directions' neighbours -- the
     same set on either side of every pair, so signals and waits match up   */
  unsigned char *snb[2];
  u32  nsym, *sym_rank;
};

/* labels of this rank that some other rank may hold too: |label| and the
   primary's local index (complemented when the primary is flagged)          */
struct xg_cand { size_t n; u64 *id; u32 *src_if; };
void xg_cand_free(struct xg_cand *c);
void xg_shared_pair(struct xg_shared *out, const struct xg_cand *c, struct xg_world *w);
void xg_shared_discover(struct xg_shared *out, const struct xg_topo *t, struct xg_world *w);
void xg_shared_free(struct xg_shared *s);
void xg_unique_flags(unsigned char *newflag, size_t n, const struct xg_topo *t,
                     struct xg_shared *sh, int me);
void xg_plan_build(struct xg_plan *pl, struct xg_shared *sh, const struct xg_topo *t);
void xg_plan_free(struct xg_plan *pl);
void xg_plan_exchange_offsets(struct xg_plan *pl, struct xg_world *w);

/* NCCL wrappers (comm.c) */
void xg_nccl_group_start(struct xg_world *w);
void xg_nccl_group_end(struct xg_world *w);
void xg_nccl_send(struct xg_world *w, const void *buf, size_t count, int xd, int peer, void *stream);
void xg_nccl_recv(struct xg_world *w, void *buf, size_t count, int xd, int peer, void *stream);
void xg_nccl_allreduce(struct xg_world *w, void *buf, size_t count, int xd, int xop, void *stream);

#endif
