/*
 * exags_cuda.h -- B200 extensions to the reference API (SURVEY.md 8b
 * "New-build addenda").  Nothing here exists in the reference; the blocking
 * calls in exags.h are the drop-in surface.  Still a plain C ABI: a CUDA
 * stream is passed as an opaque void* (a cudaStream_t).
 */
#ifndef EXAGS_CUDA_H
#define EXAGS_CUDA_H

#include "exags.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Stream-ordered gs on device-resident u: enqueues the kernels (and, for
   np>1, the halo exchange) and returns without synchronising the host.
   `stream` is a cudaStream_t; the work is ordered after what is already in
   that stream and later work in that stream is ordered after it.
   mode: 0 plain (u = T*), 1 vec (u = T[n][vn]), 2 many (u = host table of vn
   device pointers).
   CUDA graphs: the call makes no allocation and no host synchronisation, so it
   can be captured and replayed -- with several ranks too when the halo exchange
   is the push transport (the number of the exchange is device state, see
   DESIGN.md 5).  Capture after one ordinary call with the same handle and the
   largest vn * sizeof(T) the graph will use: that call sizes the handle's
   mailbox, and a later call that needs a larger one re-creates it, which would
   leave an older graph pointing at freed memory.  Every rank must replay the
   same number of exchanges, as with the plain calls.                        */
void exags_gs_async(void *u, int mode, unsigned vn, int dom, int op,
                    unsigned transpose, struct gs_data *gsh, void *stream);

/* Weighted gs: u <- w .* gs(u, op) in one pass over the field -- the pair
   "dssum, then multiply by the inverse multiplicity" that Nek5000 issues around
   the reference's gs, without the second pass.  Every value gs
   writes as the result of a group (to each member of a group of two or more
   DOFs, counted over all ranks) is multiplied by w at that member's index;
   DOFs that share their label with nobody are not touched, so the call equals
   gs() followed by u[i] *= w[i] whenever w[i] == 1 off the shared nodes, as an
   inverse-multiplicity weight is.  u and w are device vectors T[n] of the same
   floating-point dom (gs_double or gs_float); plain fields, transpose 0.
   exags_gs_weighted blocks like gs(); the _async form is ordered on `stream`
   like exags_gs_async.                                                       */
void exags_gs_weighted(void *u, const void *w, int dom, int op, struct gs_data *gsh);
void exags_gs_weighted_async(void *u, const void *w, int dom, int op,
                             struct gs_data *gsh, void *stream);
/* The same for gs_vec fields T[n][vn] and gs_many fields (host table of vn
   device pointers): one weight per node, w[i], applies to all vn components of
   node i -- the velocity field of a Nek step.  Equal, bit for bit, to vn
   weighted plain calls.                                                      */
void exags_gs_weighted_vec(void *u, unsigned vn, const void *w, int dom, int op, struct gs_data *gsh);
void exags_gs_weighted_vec_async(void *u, unsigned vn, const void *w, int dom, int op,
                                 struct gs_data *gsh, void *stream);
void exags_gs_weighted_many(void *const *u, unsigned vn, const void *w, int dom, int op, struct gs_data *gsh);
void exags_gs_weighted_many_async(void *const *u, unsigned vn, const void *w, int dom, int op,
                                  struct gs_data *gsh, void *stream);

/* Stream-ordered comm_dot (src/comm.upc:954-959 on device vectors): *d_out =
   sum over all ranks of sum_i v[i]*w[i], left in device memory and ordered on
   `stream`; nothing comes back to the host.  v, w: 16-byte aligned device
   vectors of n doubles; d_out: one device double.  The summation order is
   fixed (same bits on every call); with np > 1 the ranks' partial sums are
   combined by ncclAllReduce on the same stream.                              */
void exags_comm_dot_async(const comm_ptr comm, const double *v, const double *w,
                          exags_uint n, double *d_out, void *stream);

/* Device selection (default: LOCAL_RANK mod device count, chosen in
   comm_world).  Must be called before the first comm_world/comm_init.      */
void exags_set_device(int device);
int  exags_get_device(void);

/* How ranks exchange halo values: "single" (one rank), "push" (the pack kernel
   stores into the neighbours' mailboxes over NVLink / peer access -- default
   between distinct GPUs), "nccl" (grouped ncclSend/ncclRecv; EXAGS_EXCHANGE=nccl
   or no peer access), "ipc" (CUDA-IPC pulls with host barriers, for ranks that
   share a GPU).  EXAGS_EXCHANGE=push|nccl|ipc overrides the choice.          */
const char *exags_exchange_transport(void);

/* Introspection used by the parity tests and the bench.                    */
/* Number of library kernels launched since process start.                  */
unsigned long long exags_kernel_launches(void);
/* Handle geometry: what[0]=n, [1]=groups G, [2]=member indices S,
   [3]=flagged primaries, [4]=neighbour ranks, [5]=send slots, [6]=recv slots,
   [7]=total_shared (low 32 bits), [8]=method in use, [9]=touched 32B sectors
   of a double field, [10]=boundary groups, [11]=interior groups held in CSR
   form, [12]=their member indices, [13]=flagged primaries heading no group,
   [14]=handle bytes, [15]=entries of the sliced-ELL index array              */
void exags_gs_info(const struct gs_data *gsh, unsigned long long what[16]);
/* Write local map `which` (0 = unflagged members only, 1 = all members,
   2 = flagged primaries, 3 = pairwise recv map, 4 = pairwise send map) in the
   reference's sentinel format into out (capacity cap uints); returns the
   length needed.  Lets tests compare setup against src/gs.upc:329-370,546-570. */
size_t exags_gs_dump_map(const struct gs_data *gsh, int which,
                         exags_uint *out, size_t cap);

/* Device memory helpers so a C or Fortran caller needs no CUDA headers.    */
void *exags_device_malloc(size_t bytes);
void  exags_device_free(void *p);
void  exags_memcpy_h2d(void *dst, const void *src, size_t bytes);
void  exags_memcpy_d2h(void *dst, const void *src, size_t bytes);
void  exags_device_sync(void);
void *exags_host_malloc_pinned(size_t bytes);
/* Page-lock / release an ordinary host array that gs() will be called on, so
   that its copies run at link speed (cudaHostRegister).  The array must be
   unregistered before it is freed.  EXAGS_HOST_REGISTER=1|2 makes gs() do this
   itself for arrays it sees repeatedly.  Returns 1 on success. */
int   exags_host_register(void *p, size_t bytes);
void  exags_host_unregister(void *p);
void  exags_host_free_pinned(void *p);

#ifdef __cplusplus
}
#endif
#endif
