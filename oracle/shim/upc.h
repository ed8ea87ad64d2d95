/* upc.h -- a C stand-in for the UPC runtime, written for the test oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  It lets the reference's own .upc sources
 * (gs.upc, comm.upc, crystal.upc, sarray_transfer.upc, fail.upc) be compiled
 * as C -- after build_ref.sh has removed the `shared`/`strict` qualifiers
 * from temporary copies -- so that the oracle and the golden fixtures are
 * pinned to the reference code itself.  UPC threads become fork()ed processes
 * over one MAP_SHARED arena created before the fork: a stripped shared pointer
 * is then an ordinary pointer valid in every "thread", file-scope statics stay
 * private per thread as in UPC, and upc_memput/get are memcpy plus a fence.
 * Nothing here is taken from any UPC implementation.
 */
#ifndef EXAGS_SHIM_UPC_H
#define EXAGS_SHIM_UPC_H
#include <stddef.h>
#include <string.h>

#define __UPC__ 1
#define EXAGS_SHIM_MAXTHREADS 256

extern int shim_mythread, shim_threads;
#define MYTHREAD shim_mythread
#define THREADS  shim_threads

typedef int upc_type_t;
typedef int upc_op_t;
typedef int upc_flag_t;
typedef struct shim_lock upc_lock_t;
enum { UPC_CHAR, UPC_UCHAR, UPC_SHORT, UPC_USHORT, UPC_INT, UPC_UINT, UPC_LONG, UPC_ULONG,
       UPC_INT64, UPC_UINT64, UPC_FLOAT, UPC_DOUBLE, UPC_LDOUBLE };
enum { UPC_ADD = 1, UPC_MULT = 2, UPC_AND = 4, UPC_OR = 8, UPC_XOR = 16, UPC_LOGAND = 32,
       UPC_LOGOR = 64, UPC_MIN = 128, UPC_MAX = 256 };
#define UPC_IN_ALLSYNC 1
#define UPC_IN_MYSYNC 2
#define UPC_IN_NOSYNC 4
#define UPC_OUT_ALLSYNC 8
#define UPC_OUT_MYSYNC 16
#define UPC_OUT_NOSYNC 32

void  shim_barrier(void);
void *shim_all_alloc(size_t nblocks, size_t nbytes);
void *shim_alloc(size_t nbytes);
void  shim_poll(void);
void  shim_all_broadcast(void *dst, const void *src, size_t nbytes);
void  shim_unsupported(const char *what);

#define upc_barrier            shim_barrier()
#define upc_fence              __sync_synchronize()
#define upc_poll()             shim_poll()
#define upc_all_alloc(nb, sz)  shim_all_alloc((nb), (sz))
#define upc_alloc(sz)          shim_alloc((sz))
#define upc_free(p)            ((void)(p))
#define upc_all_free(p)        ((void)(p))

static inline void shim_copy(void *d, const void *s, size_t n)
{ __sync_synchronize(); memcpy(d, s, n); __sync_synchronize(); }
#define upc_memput(d, s, n)      shim_copy((void *)(d), (const void *)(s), (n))
#define upc_memget(d, s, n)      shim_copy((void *)(d), (const void *)(s), (n))
#define upc_memcpy(d, s, n)      shim_copy((void *)(d), (const void *)(s), (n))
#define upc_memput_nbi(d, s, n)  shim_copy((void *)(d), (const void *)(s), (n))
#define upc_memget_nbi(d, s, n)  shim_copy((void *)(d), (const void *)(s), (n))
#define upc_memset(d, c, n)      memset((void *)(d), (c), (n))
#define upc_synci()              __sync_synchronize()
#define upc_synci_attempt()      (__sync_synchronize(), 1)
#define upc_sync(h)              __sync_synchronize()
#define upc_sync_attempt(h)      (__sync_synchronize(), 1)

#define upc_all_broadcast(dst, src, n, flags) shim_all_broadcast((void *)(dst), (const void *)(src), (n))
/* the UPC-collective fast paths are never taken: build_ref.sh routes
   comm_scan/comm_allreduce to their hand-rolled put/flag versions */
#define upc_all_reduceD(...)        shim_unsupported("upc_all_reduceD")
#define upc_all_reduceF(...)        shim_unsupported("upc_all_reduceF")
#define upc_all_reduceI(...)        shim_unsupported("upc_all_reduceI")
#define upc_all_reduceL(...)        shim_unsupported("upc_all_reduceL")
#define upc_all_prefix_reduceD(...) shim_unsupported("upc_all_prefix_reduceD")
#define upc_all_prefix_reduceF(...) shim_unsupported("upc_all_prefix_reduceF")
#define upc_all_prefix_reduceI(...) shim_unsupported("upc_all_prefix_reduceI")
#define upc_all_prefix_reduceL(...) shim_unsupported("upc_all_prefix_reduceL")

upc_lock_t *shim_lock_alloc(void);
void shim_lock(upc_lock_t *l);
void shim_unlock(upc_lock_t *l);
#define upc_global_lock_alloc() shim_lock_alloc()
#define upc_global_lock_free(l) ((void)(l))
#define upc_lock_free(l)        ((void)(l))
#define upc_lock(l)             shim_lock(l)
#define upc_unlock(l)           shim_unlock(l)
#endif
