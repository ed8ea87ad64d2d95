/*
 * exags.h -- public C ABI of the B200-native ExaGS gather-scatter library.
 *
 * This is the drop-in boundary for the gs_op hot path of
 * EPCCed/exaflow-exags-upc.  Every entry point below names the reference
 * interface it replaces (paths relative to the reference tree).  The ABI is
 * plain C: pointers, sizes and small enums; no CUDA or torch types appear in
 * any signature.  `u` arguments may be device (HBM) pointers -- the fast
 * path -- or ordinary host pointers, which are staged through HBM; in both
 * cases every reduction runs in the sm_100a kernels (there is no CPU path).
 *
 * Two build flavours of the reference exist and both are exported from the
 * one shared object, told apart by their symbol prefix exactly as the
 * reference's PREFIX macro does (src/name.h:20-27, configure.ac:57-61):
 *
 *   exags_*   default build:  slong == int   (32-bit global ids)
 *   gslib_*   Nek5000 build:  slong == long long (GLOBAL_LONG_LONG)
 *
 * Define EXAGS_NEK (or GLOBAL_LONG_LONG) before including this header to get
 * the gslib_ flavour behind the short names; define EXAGS_SHORT_NAMES to get
 * the reference's unprefixed spellings (gs, gs_setup, comm_world, ...).
 */
#ifndef EXAGS_H
#define EXAGS_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(GLOBAL_LONG_LONG) && !defined(EXAGS_NEK)
#define EXAGS_NEK 1
#endif

/* ---- integer widths (src/types.h:52-76) -------------------------------- */
typedef int                exags_sint;   /* local index type, "sint"        */
typedef unsigned int       exags_uint;   /* "uint"; map sentinel is (uint)-1 */
#ifdef EXAGS_NEK
typedef long long          exags_slong;  /* global id type under Nek flags  */
typedef unsigned long long exags_ulong;
#else
typedef int                exags_slong;  /* global id type, default flags   */
typedef unsigned int       exags_ulong;
#endif

/* ---- domains and operations (src/gs_defs.h:18-31, 59-79) --------------- */
/* Numeric values are ABI: Fortran and C callers pass them as plain ints.   */
typedef enum {
  gs_double = 0,
  gs_float  = 1,
  gs_int    = 2,
  gs_long   = 3,
#ifdef EXAGS_NEK
  gs_long_long = 4,
  gs_dom_n  = 5
#else
  gs_dom_n  = 4
#endif
} gs_dom;

#define gs_sint gs_int
#ifdef EXAGS_NEK
#define gs_slong gs_long_long
#else
#define gs_slong gs_int
#endif

#ifndef EXAGS_NO_GS_OP_ENUM
/* the reference renames this enum to gs_op_t inside gs.upc (src/gs.upc:12)
   because Fortran's entry point is also called gs_op; we keep both names. */
typedef enum {
  gs_add = 0,   /* a += b                      (src/gs_defs.h:33) */
  gs_mul = 1,   /* a *= b                      (src/gs_defs.h:34) */
  gs_min = 2,   /* if (b < a) a = b            (src/gs_defs.h:35) */
  gs_max = 3,   /* if (b > a) a = b            (src/gs_defs.h:36) */
  gs_bpr = 4,   /* binary-prefix on unsigned   (src/gs_defs.h:37-42) */
  gs_op_n = 5
} gs_op_t;
#ifndef EXAGS_FORTRAN_GS_OP
typedef gs_op_t gs_op;
#endif
#endif

/* exchange method (src/gs.h:127) */
typedef enum {
  gs_auto = 0, gs_pairwise = 1, gs_crystal_router = 2, gs_all_reduce = 3
} gs_method;

/* ---- scratch buffer (src/mem.h:111,158: struct array / buffer) --------- */
struct array { void *ptr; size_t n, max; };
typedef struct array buffer;

/* ---- communicator (src/comm.h:128-154) --------------------------------- */
/* The first three members are read directly by reference callers
   (tests/gs/gs_test.upc:31 uses cp->np, cp->id); everything after them is
   private to the library: rank<->GPU binding, host rendezvous segment, NCCL
   communicator and CUDA streams replace the UPC shared-buffer directories. */
struct comm {
  int h;
  int id, np;
  void *priv;       /* library-owned world state                      */
  int  *ref_count;  /* shared by comm_dup copies (src/comm.upc:256-284) */
};
typedef struct comm *comm_ptr;

struct gs_data;     /* opaque handle (src/gs.upc:1161-1167) */

/* ======================================================================
 *  gather-scatter operator                       replaces src/gs.h:129-138
 * ====================================================================== */

/* gs(): u[i] <- OP over all (p,j) with |id_p[j]| == |id[i]|.
   Replaces gs() src/gs.upc:1187-1191.  Blocking: on return u holds the
   result (device work is complete and, for host u, copied back).          */
void exags_gs(void *u, gs_dom dom, gs_op_t op, unsigned transpose,
              struct gs_data *gsh, buffer *buf);
/* u is T[n][vn], tuples contiguous.   Replaces gs_vec() src/gs.upc:1193-1197 */
void exags_gs_vec(void *u, unsigned vn, gs_dom dom, gs_op_t op,
                  unsigned transpose, struct gs_data *gsh, buffer *buf);
/* u is an array of vn pointers to T[n] (the pointer table itself is in host
   memory; each entry may be host or device).
   Replaces gs_many() src/gs.upc:1199-1203                                  */
void exags_gs_many(void *const *u, unsigned vn, gs_dom dom, gs_op_t op,
                   unsigned transpose, struct gs_data *gsh, buffer *buf);
/* Replaces gs_setup() src/gs.upc:1262-1269 (default flags: 32-bit ids).    */
struct gs_data *exags_gs_setup(const int *id, exags_uint n, const comm_ptr comm,
                               int unique, gs_method method, int verbose);
/* Replaces gs_free() src/gs.upc:1271-1279 */
void exags_gs_free(struct gs_data *gsh);
/* Replaces gs_unique() src/gs.upc:1281-1290: negates non-owner ids in place */
void exags_gs_unique(int *id, exags_uint n, const comm_ptr comm);

/* Nek5000 flavour (-DPREFIX=gslib_ -DGLOBAL_LONG_LONG): 64-bit ids.        */
void gslib_gs(void *u, int dom, int op, unsigned transpose,
              struct gs_data *gsh, buffer *buf);
void gslib_gs_vec(void *u, unsigned vn, int dom, int op,
                  unsigned transpose, struct gs_data *gsh, buffer *buf);
void gslib_gs_many(void *const *u, unsigned vn, int dom, int op,
                   unsigned transpose, struct gs_data *gsh, buffer *buf);
struct gs_data *gslib_gs_setup(const long long *id, exags_uint n,
                               const comm_ptr comm, int unique, int method,
                               int verbose);
void gslib_gs_free(struct gs_data *gsh);
void gslib_gs_unique(long long *id, exags_uint n, const comm_ptr comm);

/* ======================================================================
 *  local kernels on sentinel maps              replaces src/gs_local.h:23-41
 *  map format: [dst, src, src, ..., (uint)-1]*, one extra (uint)-1 at the
 *  end (src/gs.upc:329-357).  `map` is a host array; out/in may be host or
 *  device pointers.  All of them run on the GPU.
 * ====================================================================== */
void exags_gs_gather_array(void *out, const void *in, exags_uint n,
                           gs_dom dom, gs_op_t op);            /* gs_local.c:187 */
void exags_gs_init_array(void *out, exags_uint n, gs_dom dom, gs_op_t op);
                                                               /* gs_local.c:196 */
void exags_gs_gather(void *out, const void *in, const unsigned vn,
                     const exags_uint *map, gs_dom dom, gs_op_t op);  /* :206 */
void exags_gs_scatter(void *out, const void *in, const unsigned vn,
                      const exags_uint *map, gs_dom dom);              /* :216 */
void exags_gs_init(void *out, const unsigned vn, const exags_uint *map,
                   gs_dom dom, gs_op_t op);                            /* :224 */
void exags_gs_gather_vec(void *out, const void *in, const unsigned vn,
                         const exags_uint *map, gs_dom dom, gs_op_t op); /* :235 */
void exags_gs_scatter_vec(void *out, const void *in, const unsigned vn,
                          const exags_uint *map, gs_dom dom);          /* :132 */
void exags_gs_init_vec(void *out, const unsigned vn, const exags_uint *map,
                       gs_dom dom, gs_op_t op);                        /* :245 */
void exags_gs_gather_many(void *out, const void *in, const unsigned vn,
                          const exags_uint *map, gs_dom dom, gs_op_t op); /* :256 */
void exags_gs_scatter_many(void *out, const void *in, const unsigned vn,
                           const exags_uint *map, gs_dom dom);         /* :268 */
void exags_gs_init_many(void *out, const unsigned vn, const exags_uint *map,
                        gs_dom dom, gs_op_t op);                       /* :279 */
void exags_gs_gather_vec_to_many(void *out, const void *in, const unsigned vn,
                                 const exags_uint *map, gs_dom dom, gs_op_t op);
                                                                       /* :295 */
void exags_gs_scatter_many_to_vec(void *out, const void *in, const unsigned vn,
                                  const exags_uint *map, gs_dom dom);  /* :308 */
void exags_gs_scatter_vec_to_many(void *out, const void *in, const unsigned vn,
                                  const exags_uint *map, gs_dom dom);  /* :320 */

/* ======================================================================
 *  communicator                               replaces src/comm.h:167-206
 *  Rank identity comes from the launcher's environment (RANK/WORLD_SIZE/
 *  LOCAL_RANK as torchrun sets them, or EXAGS_RANK/EXAGS_NP); each rank
 *  binds GPU (LOCAL_RANK mod #GPUs).  One node only.
 * ====================================================================== */
extern comm_ptr exags_glb_comm;                       /* comm.h:156 glb_comm */
void exags_comm_init(void);                           /* comm.upc:35  */
void exags_comm_finalize(void);                       /* comm.upc:55  */
void exags_comm_world(comm_ptr *cpp);                 /* comm.upc:223 */
void exags_comm_dup(comm_ptr *cpp, const comm_ptr cp);/* comm.upc:256 */
void exags_comm_free(comm_ptr *cpp);                  /* comm.upc:287 */
void exags_comm_np(const comm_ptr cp, int *np);       /* comm.upc:360 */
void exags_comm_id(const comm_ptr cp, int *id);       /* comm.upc:368 */
void exags_comm_time(double *tm);                     /* comm.upc:411 */
void exags_comm_barrier(const comm_ptr cp);           /* comm.upc:420 */
void exags_comm_bcast(const comm_ptr cp, void *p, size_t n, exags_uint root);
                                                      /* comm.upc:427 */
/* v <- OP over ranks, element-wise; buf is scratch of the same size (unused).
   v may be a host vector (the reference's contract) or a device vector.
   Replaces comm_allreduce() src/comm.upc:831-950                           */
void exags_comm_allreduce(const comm_ptr cp, gs_dom dom, gs_op_t op,
                          void *v, exags_uint vn, void *buf);
/* scan[0..vn) = OP over ranks < id (exclusive), scan[vn..2vn) = total.
   Replaces comm_scan() src/comm.upc:706-801                                */
void exags_comm_scan(void *scan, const comm_ptr cp, gs_dom dom, gs_op_t op,
                     const void *v, exags_uint vn, void *buffer);
/* sum over ranks of v.w; v and w both host vectors, or both 16-byte aligned
   device vectors (then the local part is an HBM-bound reduction kernel).   */
double exags_comm_dot(const comm_ptr cp, double *v, double *w, exags_uint n);
                                                      /* comm.upc:954 */
/* fold a local host array with op, then across ranks (comm.upc:961-984)    */
double    exags_comm_reduce__double(const comm_ptr cp, gs_op_t op, const double *in, exags_uint n);
float     exags_comm_reduce__float(const comm_ptr cp, gs_op_t op, const float *in, exags_uint n);
int       exags_comm_reduce__int(const comm_ptr cp, gs_op_t op, const int *in, exags_uint n);
long      exags_comm_reduce__long(const comm_ptr cp, gs_op_t op, const long *in, exags_uint n);
long long exags_comm_reduce__long_long(const comm_ptr cp, gs_op_t op, const long long *in, exags_uint n);
/* Element type tags for the Fortran-facing reductions (the reference hands
   out upc_type_t values, src/comm.h:72, src/comm.upc:376-400; callers only
   ever obtain them from these four functions).                             */
typedef int comm_type;
void exags_comm_type_int(comm_type *ct);              /* comm.upc:376 */
void exags_comm_type_int8(comm_type *ct);             /* comm.upc:382 */
void exags_comm_type_real(comm_type *ct);             /* comm.upc:389 */
void exags_comm_type_dp(comm_type *ct);               /* comm.upc:396 */
void exags_comm_tag_ub(const comm_ptr cp, int *ub);   /* comm.upc:403 (MPI only: 0) */
void exags_comm_allreduce_cdom(const comm_ptr cp, comm_type cdom, gs_op_t op,
                               void *v, exags_uint vn, void *buf);  /* comm.upc:804 */
/* Fortran bindings of the communicator (src/comm.upc:1031-1148) are exported
   as comm_init_, comm_finalize_, comm_world_, comm_free_, comm_np_, comm_id_,
   comm_type_int_, comm_type_int8_, comm_type_real_, comm_type_dp_,
   comm_tag_ub_, comm_time_, comm_barrier_, comm_bcast_, comm_allreduce_add_,
   comm_allreduce_min_, comm_allreduce_max_, comm_allreduce_mul_ (and the same
   with the fgslib_ prefix for a Nek build).                                */

/* ======================================================================
 *  errors                                       replaces src/fail.h:32-38
 *  fail() prints "ERROR (proc NNNN, file:line): msg" to stdout and exits
 *  (src/fail.upc:19-63).  CUDA and NCCL errors are routed here too.
 * ====================================================================== */
void exags_fail(int status, const char *file, unsigned line,
                const char *fmt, ...);
void exags_die(int status);
void exags_diagnostic(const char *prefix, const char *file, unsigned line,
                      const char *fmt, ...);
extern exags_uint exags_comm_gbl_id, exags_comm_gbl_np;   /* comm.h:121-126 */

/* ======================================================================
 *  Fortran bindings                         replaces src/gs.upc:1309-1418
 *  All arguments by reference; handles are 0-based ints; dom 1..3 =
 *  {double, sint, slong}; op 1..4 = {add, mul, min, max}.  Exported both
 *  bare and with a trailing underscore (UNDERSCORE, src/name.h:32-41), and
 *  with the Nek prefix fgslib_ (64-bit ids).
 * ====================================================================== */
void exags_fgs_setup(int *handle, const int id[], const int *n,
                     const int *comm, const int *np);
void exags_fgs_setup_pick(int *handle, const int id[], const int *n,
                          const int *comm, const int *np, const int *method);
void exags_fgs(const int *handle, void *u, const int *dom, const int *op,
               const int *transpose);
void exags_fgs_vec(const int *handle, void *u, const int *n, const int *dom,
                   const int *op, const int *transpose);
void exags_fgs_many(const int *handle, void *u1, void *u2, void *u3, void *u4,
                    void *u5, void *u6, const int *n, const int *dom,
                    const int *op, const int *transpose);
void exags_fgs_fields(const int *handle, void *u, const int *stride,
                      const int *n, const int *dom, const int *op,
                      const int *transpose);
void exags_fgs_free(const int *handle);
/* Nek flavour, as Nek5000's Fortran links them */
void fgslib_gs_setup_(int *handle, const long long id[], const int *n,
                      const int *comm, const int *np);
void fgslib_gs_setup_pick_(int *handle, const long long id[], const int *n,
                           const int *comm, const int *np, const int *method);
void fgslib_gs_op_(const int *handle, void *u, const int *dom, const int *op,
                   const int *transpose);
void fgslib_gs_op_vec_(const int *handle, void *u, const int *n,
                       const int *dom, const int *op, const int *transpose);
void fgslib_gs_op_many_(const int *handle, void *u1, void *u2, void *u3,
                        void *u4, void *u5, void *u6, const int *n,
                        const int *dom, const int *op, const int *transpose);
void fgslib_gs_op_fields_(const int *handle, void *u, const int *stride,
                          const int *n, const int *dom, const int *op,
                          const int *transpose);
void fgslib_gs_free_(const int *handle);

/* ---- reference spellings ---------------------------------------------- */
#ifdef EXAGS_SHORT_NAMES
# ifdef EXAGS_NEK
#  define EXAGS_PFX(x) gslib_##x
# else
#  define EXAGS_PFX(x) exags_##x
# endif
# define gs         EXAGS_PFX(gs)
# define gs_vec     EXAGS_PFX(gs_vec)
# define gs_many    EXAGS_PFX(gs_many)
# define gs_setup   EXAGS_PFX(gs_setup)
# define gs_free    EXAGS_PFX(gs_free)
# define gs_unique  EXAGS_PFX(gs_unique)
# define comm_init      exags_comm_init
# define comm_finalize  exags_comm_finalize
# define comm_world     exags_comm_world
# define comm_dup       exags_comm_dup
# define comm_free      exags_comm_free
# define comm_np        exags_comm_np
# define comm_id        exags_comm_id
# define comm_time      exags_comm_time
# define comm_barrier   exags_comm_barrier
# define comm_bcast     exags_comm_bcast
# define comm_allreduce exags_comm_allreduce
# define comm_scan      exags_comm_scan
# define comm_dot       exags_comm_dot
# define comm_reduce_double    exags_comm_reduce__double
# define comm_reduce_float     exags_comm_reduce__float
# define comm_reduce_int       exags_comm_reduce__int
# define comm_reduce_long      exags_comm_reduce__long
# define comm_reduce_long_long exags_comm_reduce__long_long
# define comm_reduce_sint      exags_comm_reduce__int        /* comm.h:216-219 */
# ifdef EXAGS_NEK
#  define comm_reduce_slong    exags_comm_reduce__long_long
# else
#  define comm_reduce_slong    exags_comm_reduce__int
# endif
# define comm_type_int  exags_comm_type_int
# define comm_type_int8 exags_comm_type_int8
# define comm_type_real exags_comm_type_real
# define comm_type_dp   exags_comm_type_dp
# define comm_tag_ub    exags_comm_tag_ub
# define comm_allreduce_cdom exags_comm_allreduce_cdom
# define glb_comm       exags_glb_comm
# define fail           exags_fail
# define die            exags_die
# define diagnostic     exags_diagnostic
# define gs_gather_array        exags_gs_gather_array
# define gs_init_array          exags_gs_init_array
# define gs_gather              exags_gs_gather
# define gs_scatter             exags_gs_scatter
# define gs_init                exags_gs_init
# define gs_gather_vec          exags_gs_gather_vec
# define gs_scatter_vec         exags_gs_scatter_vec
# define gs_init_vec            exags_gs_init_vec
# define gs_gather_many         exags_gs_gather_many
# define gs_scatter_many        exags_gs_scatter_many
# define gs_init_many           exags_gs_init_many
# define gs_gather_vec_to_many  exags_gs_gather_vec_to_many
# define gs_scatter_many_to_vec exags_gs_scatter_many_to_vec
# define gs_scatter_vec_to_many exags_gs_scatter_vec_to_many
# ifndef EXAGS_NO_INT_MACROS
/* glibc's <sys/types.h> has its own `uint` and `ulong` typedefs: take it in
   before the names become macros, so that a system header included after this
   one cannot re-declare them through the macros (the reference's types.h has
   the same hazard and relies on callers including system headers first) */
#  include <sys/types.h>
#  include <stdlib.h>
#  define sint  exags_sint
#  define uint  exags_uint
#  define slong exags_slong
#  define ulong exags_ulong
# endif
#endif

#ifdef __cplusplus
}
#endif
#endif /* EXAGS_H */
