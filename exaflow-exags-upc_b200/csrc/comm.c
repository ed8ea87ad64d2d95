/* comm.c -- communicator, rank identity, host rendezvous and NCCL bootstrap.
 *
 * Replaces the reference's UPC "comm" layer (src/comm.h:128-206,
 * src/comm.upc:35-424,706-959) for the gs hot path.  The reference keeps
 * rank identity in MYTHREAD/THREADS and moves data with upc_memput + polled
 * flags in shared directories.  Here:
 *
 *   - rank identity comes from the launcher environment (one process per
 *     GPU: RANK/WORLD_SIZE/LOCAL_RANK as torchrun sets them);
 *   - host-side setup traffic (the id rendezvous of gs_setup, comm_scan,
 *     comm_allreduce on host vectors) goes through a POSIX shared-memory
 *     rendezvous on the node -- the direct analogue of the PGAS segment,
 *     and it works without a GPU, which is how the multi-rank host logic is
 *     tested;
 *   - per-call halo traffic goes through NCCL send/recv over NVLink on a
 *     dedicated high-priority stream (gs_host.c); NCCL is dlopen()ed on first
 *     multi-rank use so single-GPU users never need it.
 *
 * One node only (the north star is one 8xB200 box).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <stdatomic.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <time.h>
#include <unistd.h>
#include <float.h>
#include <limits.h>
#include <stdint.h>
#include "exags_internal.h"

#define XW_MAXNP   64
#define XW_SMALL   4096            /* per-rank inline area for small collectives */
#define XW_MAGIC   0x45584147u     /* "EXAG" */

struct xw_ctl {
  _Atomic unsigned magic;
  unsigned np;
  _Atomic unsigned attached;
  _Atomic unsigned bar_count;
  _Atomic unsigned bar_sense;
  /* who is attached: a rank that waits long looks whether the others are still
     alive, and a late rank can tell a segment left behind by a killed run
     (its owner, rank 0, is gone) from the one this launch created */
  int pid[XW_MAXNP];
  unsigned long long owner_start;     /* start time of rank 0's process (/proc/<pid>/stat field 22) */
  unsigned long long pub_bytes[XW_MAXNP];
  char small[XW_MAXNP][XW_SMALL];
};

struct xg_world {
  int rank, np, local_rank, device, has_cuda;
  /* host rendezvous */
  char key[64];
  struct xw_ctl *ctl;
  unsigned sense;
  unsigned long long epoch;
  /* device side */
  void *stream[4];
  void *event[4];
  void *scratch[4];
  void *dot_scratch;                 /* partial sums of comm_dot_async */
  size_t scratch_bytes[4];
  /* nccl */
  void *nccl_lib;
  void *nccl_comm;
  /* cuda-ipc transport */
  int use_ipc;
  void *ipc_mine; size_t ipc_bytes;
  void **ipc_peer;
  /* push transport */
  int use_push;
  int *push_status;
};

static struct xg_world *g_world = 0;
static int g_device_override = -1;
comm_ptr exags_glb_comm = 0;

static const char *env_first(const char *const *names)
{
  for (; *names; ++names) {
    const char *v = getenv(*names);
    if (v && *v) return v;
  }
  return 0;
}

/* ------------------------------------------------------------------ */
/*  host rendezvous                                                    */
/* ------------------------------------------------------------------ */
static void relax(unsigned *spins)
{
  if (++*spins < 200) { __asm__ __volatile__("pause"); return; }
  if (*spins < 2000) { sched_yield(); return; }
  struct timespec ts = { 0, 50000 };
  nanosleep(&ts, 0);
}

/* start time of a process in clock ticks since boot (0 when it does not exist) */
static unsigned long long proc_start_time(int pid)
{
  unsigned long long start = 0;
  char path[64], line[1024];
  snprintf(path, sizeof path, "/proc/%d/stat", pid);
  FILE *f = fopen(path, "r");
  if (!f) return 0;
  if (fgets(line, sizeof line, f)) {
    char *p = strrchr(line, ')');
    int field = 2;
    while (p && *p && field < 22) { if (*p == ' ') ++field; ++p; }
    if (p) start = strtoull(p, 0, 10);
  }
  fclose(f);
  return start;
}

/* 0 when the process has ended (a zombie waiting to be reaped counts as ended) */
static int proc_alive(int pid)
{
  char path[64], line[512];
  snprintf(path, sizeof path, "/proc/%d/stat", pid);
  FILE *f = fopen(path, "r");
  if (!f) return kill(pid, 0) == 0 || errno != ESRCH;     /* no /proc: fall back to the signal probe */
  int alive = 1;
  if (fgets(line, sizeof line, f)) {
    const char *p = strrchr(line, ')');
    if (p && p[1] == ' ' && (p[2] == 'Z' || p[2] == 'X')) alive = 0;
  }
  fclose(f);
  return alive;
}

static double host_timeout(void)
{
  static double t = -1;
  if (t < 0) { const char *v = getenv("EXAGS_HOST_TIMEOUT_S"); t = v && atof(v) > 0 ? atof(v) : 1800.0; }
  return t;
}

/* Called now and then from a wait loop of the rendezvous: a peer that died
   (exags_fail, a signal) ends the wait with the reference's fail() instead of a
   hang, and no wait lasts longer than EXAGS_HOST_TIMEOUT_S (default 1800 s).  */
static void wait_check(struct xg_world *w, double t0, const char *what)
{
  double now; exags_comm_time(&now);
  for (int p = 0; p < w->np; ++p) {
    const int pid = w->ctl->pid[p];
    if (p != w->rank && pid > 0 && !proc_alive(pid))
      xfail("comm: rank %d (pid %d) is gone while rank %d waits in %s", p, pid, w->rank, what);
  }
  if (now - t0 > host_timeout())
    xfail("comm: rank %d waited more than %.0f s in %s (EXAGS_HOST_TIMEOUT_S)", w->rank, host_timeout(), what);
}

static void make_key(struct xg_world *w)
{
  const char *k = getenv("EXAGS_JOB_KEY");
  if (k && *k) { snprintf(w->key, sizeof w->key, "%.48s", k); return; }
  /* every rank of one launch has the same parent (torchrun agent, mpirun,
     a test harness): key on that parent's pid and start time */
  const unsigned long long start = proc_start_time((int)getppid());
  const char *port = getenv("MASTER_PORT");
  /* an elastic launcher restarts its workers under the same agent: the restart
     count keeps a new attempt off the segment a failed one may have left */
  const char *attempt = getenv("TORCHELASTIC_RESTART_COUNT");
  snprintf(w->key, sizeof w->key, "%d-%llu-%.8s-%.4s", (int)getppid(), start, port ? port : "0",
           attempt ? attempt : "0");
}

static void ctl_name(const struct xg_world *w, char *out, size_t cap)
{ snprintf(out, cap, "/exags-%s-ctl", w->key); }
static void pub_name(const struct xg_world *w, int rank, unsigned long long epoch, char *out, size_t cap)
{ snprintf(out, cap, "/exags-%s-r%d-e%llu", w->key, rank, epoch); }

static void rendezvous_attach(struct xg_world *w)
{
  char name[128];
  make_key(w);
  ctl_name(w, name, sizeof name);
  int fd = -1;
  if (w->rank == 0) {
    shm_unlink(name);
    fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) xfail("comm: cannot create rendezvous segment %s: %s", name, strerror(errno));
    if (ftruncate(fd, (off_t)sizeof(struct xw_ctl)) != 0)
      xfail("comm: cannot size rendezvous segment: %s", strerror(errno));
  } else {
    /* A segment of this name may be the leftover of a killed run with the same
       key (only atexit unlinks): it is this launch's only if its owner -- the
       process that initialised it -- is alive and is the process it says it
       is.  Anything else is skipped until rank 0 has replaced it.            */
    unsigned spins = 0; double t0, now; exags_comm_time(&t0);
    for (;;) {
      fd = shm_open(name, O_RDWR, 0600);
      if (fd >= 0) {
        struct stat st;
        if (fstat(fd, &st) == 0 && (size_t)st.st_size >= sizeof(struct xw_ctl)) {
          struct xw_ctl *c = (struct xw_ctl *)mmap(0, sizeof(struct xw_ctl), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
          if (c != MAP_FAILED) {
            const int owner = c->pid[0];
            if (atomic_load(&c->magic) == XW_MAGIC && owner > 0 && c->owner_start &&
                proc_start_time(owner) == c->owner_start) { w->ctl = c; close(fd); break; }
            munmap(c, sizeof(struct xw_ctl));
          }
        }
        close(fd); fd = -1;
      }
      relax(&spins);
      exags_comm_time(&now);
      if (now - t0 > 120.0) xfail("comm: rank %d timed out waiting for rank 0's rendezvous segment %s", w->rank, name);
    }
  }
  if (w->rank == 0) {
    w->ctl = (struct xw_ctl *)mmap(0, sizeof(struct xw_ctl), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (w->ctl == MAP_FAILED) xfail("comm: mmap of rendezvous segment failed: %s", strerror(errno));
    w->ctl->np = (unsigned)w->np;
    w->ctl->pid[0] = (int)getpid();
    w->ctl->owner_start = proc_start_time((int)getpid());
    atomic_store(&w->ctl->attached, 1u);
    atomic_store(&w->ctl->bar_count, 0u);
    atomic_store(&w->ctl->bar_sense, 0u);
    atomic_store(&w->ctl->magic, XW_MAGIC);
  } else {
    if (w->ctl->np != (unsigned)w->np)
      xfail("comm: rendezvous segment is for %u ranks, this job has %d", w->ctl->np, w->np);
    w->ctl->pid[w->rank] = (int)getpid();
    atomic_fetch_add(&w->ctl->attached, 1u);
  }
  unsigned spins = 0; double t0; exags_comm_time(&t0);
  while (atomic_load(&w->ctl->attached) < (unsigned)w->np) {
    relax(&spins);
    if ((spins & 0x3fffu) == 0) wait_check(w, t0, "comm_init (not every rank has attached)");
  }
  w->sense = 0;
}

void xg_host_barrier(struct xg_world *w)
{
  if (w->np == 1) return;
  unsigned mine = (w->sense ^= 1u);
  if (atomic_fetch_add(&w->ctl->bar_count, 1u) == (unsigned)w->np - 1u) {
    atomic_store(&w->ctl->bar_count, 0u);
    atomic_store(&w->ctl->bar_sense, mine);
  } else {
    unsigned spins = 0; double t0 = 0;
    while (atomic_load(&w->ctl->bar_sense) != mine) {
      relax(&spins);
      if ((spins & 0x3fffu) == 0) {               /* about once a second once the wait sleeps */
        if (t0 == 0) exags_comm_time(&t0);
        wait_check(w, t0, "a host barrier");
      }
    }
  }
}

/* publish a block of bytes for this epoch; returns a writable mapping */
static void *pub_create(struct xg_world *w, size_t bytes)
{
  char name[128];
  pub_name(w, w->rank, w->epoch, name, sizeof name);
  shm_unlink(name);
  int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0) xfail("comm: cannot create exchange segment %s: %s", name, strerror(errno));
  size_t sz = bytes ? bytes : 1;
  if (ftruncate(fd, (off_t)sz) != 0) xfail("comm: cannot size exchange segment (%lu bytes): %s", (unsigned long)sz, strerror(errno));
  void *p = mmap(0, sz, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) xfail("comm: mmap of exchange segment failed: %s", strerror(errno));
  w->ctl->pub_bytes[w->rank] = bytes;
  return p;
}

static const void *pub_peer(struct xg_world *w, int peer, size_t *bytes)
{
  char name[128];
  pub_name(w, peer, w->epoch, name, sizeof name);
  int fd = shm_open(name, O_RDONLY, 0600);
  if (fd < 0) xfail("comm: cannot open exchange segment %s: %s", name, strerror(errno));
  size_t sz = (size_t)w->ctl->pub_bytes[peer];
  void *p = mmap(0, sz ? sz : 1, PROT_READ, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) xfail("comm: mmap of peer exchange segment failed: %s", strerror(errno));
  *bytes = sz;
  return p;
}

static void pub_destroy(struct xg_world *w, void *mine, size_t bytes)
{
  char name[128];
  munmap(mine, bytes ? bytes : 1);
  pub_name(w, w->rank, w->epoch, name, sizeof name);
  shm_unlink(name);
  ++w->epoch;
}

void xg_host_allgather(struct xg_world *w, const void *in, size_t bytes, void *out)
{
  if (w->np == 1) { memcpy(out, in, bytes); return; }
  if (bytes <= XW_SMALL) {
    memcpy(w->ctl->small[w->rank], in, bytes);
    xg_host_barrier(w);
    for (int p = 0; p < w->np; ++p) memcpy((char *)out + (size_t)p * bytes, w->ctl->small[p], bytes);
    xg_host_barrier(w);
    return;
  }
  void *mine = pub_create(w, bytes);
  memcpy(mine, in, bytes);
  xg_host_barrier(w);
  for (int p = 0; p < w->np; ++p) {
    size_t pb;
    const void *q = pub_peer(w, p, &pb);
    memcpy((char *)out + (size_t)p * bytes, q, bytes);
    munmap((void *)q, pb ? pb : 1);
  }
  xg_host_barrier(w);
  pub_destroy(w, mine, bytes);
}

/* Label-presence bitmaps of gs_setup.  xg_host_bitmap_new returns this rank's
   bitmap inside a fresh shared segment (zero-filled by the OS, so there is
   nothing to clear or copy); xg_host_bitmap_or_others then fills others[k] with
   the OR of every OTHER rank's word k and releases the segment.  `words` is the
   same on every rank.                                                        */
u64 *xg_host_bitmap_new(struct xg_world *w, size_t words)
{
  if (w->np == 1) return xnew0(u64, words);
  return (u64 *)pub_create(w, words * sizeof(u64));
}

void xg_host_bitmap_or_others(struct xg_world *w, u64 *mine, size_t words, u64 *others)
{
  memset(others, 0, words * sizeof(u64));
  if (w->np == 1) { free(mine); return; }
  xg_host_barrier(w);
  for (int p = 0; p < w->np; ++p) {
    if (p == w->rank) continue;
    size_t pb;
    const u64 *q = (const u64 *)pub_peer(w, p, &pb);
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(xg_setup_threads())
#endif
    for (size_t k = 0; k < words; ++k) others[k] |= q[k];
    munmap((void *)q, pb ? pb : 1);
  }
  xg_host_barrier(w);
  pub_destroy(w, mine, words * sizeof(u64));
}

void *xg_host_alltoallv(struct xg_world *w, const void *rows, size_t rowbytes,
                        const size_t *sendcnt, size_t *recvcnt, size_t *nrecv)
{
  int np = w->np;
  size_t total = 0;
  for (int p = 0; p < np; ++p) total += sendcnt[p];
  if (np == 1) {
    void *out = xmalloc_(total * rowbytes, __FILE__, __LINE__);
    memcpy(out, rows, total * rowbytes);
    recvcnt[0] = total; *nrecv = total;
    return out;
  }
  unsigned long long *mine_cnt = xnew(unsigned long long, np);
  unsigned long long *all_cnt = xnew(unsigned long long, (size_t)np * np);
  for (int p = 0; p < np; ++p) mine_cnt[p] = sendcnt[p];
  xg_host_allgather(w, mine_cnt, (size_t)np * sizeof *mine_cnt, all_cnt);
  size_t got = 0;
  for (int p = 0; p < np; ++p) { recvcnt[p] = (size_t)all_cnt[(size_t)p * np + w->rank]; got += recvcnt[p]; }
  void *mine = pub_create(w, total * rowbytes);
  memcpy(mine, rows, total * rowbytes);
  xg_host_barrier(w);
  char *out = (char *)xmalloc_(got * rowbytes, __FILE__, __LINE__);
  size_t o = 0;
  for (int p = 0; p < np; ++p) {
    size_t skip = 0, pb;
    for (int q = 0; q < w->rank; ++q) skip += (size_t)all_cnt[(size_t)p * np + q];
    if (!recvcnt[p]) continue;
    const char *src = (const char *)pub_peer(w, p, &pb);
    memcpy(out + o * rowbytes, src + skip * rowbytes, recvcnt[p] * rowbytes);
    munmap((void *)src, pb ? pb : 1);
    o += recvcnt[p];
  }
  xg_host_barrier(w);
  pub_destroy(w, mine, total * rowbytes);
  free(mine_cnt); free(all_cnt);
  *nrecv = got;
  return out;
}

/* ------------------------------------------------------------------ */
/*  NCCL via dlopen                                                    */
/* ------------------------------------------------------------------ */
typedef struct { char internal[128]; } xg_nccl_uid;
static struct {
  int (*GetUniqueId)(xg_nccl_uid *);
  int (*CommInitRank)(void **, int, xg_nccl_uid, int);
  int (*CommDestroy)(void *);
  int (*Send)(const void *, size_t, int, int, void *, void *);
  int (*Recv)(void *, size_t, int, int, void *, void *);
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, void *);
  int (*GroupStart)(void);
  int (*GroupEnd)(void);
  const char *(*GetErrorString)(int);
} nc;

#define NC(call)                                                              \
  do {                                                                        \
    int r_ = (call);                                                          \
    if (r_ != 0)                                                              \
      xfail("NCCL error: %s (%s)", nc.GetErrorString ? nc.GetErrorString(r_) : "?", #call); \
  } while (0)

static void nccl_load(struct xg_world *w)
{
  if (w->nccl_lib) return;
  const char *names[] = { getenv("EXAGS_NCCL_LIB"), "libnccl.so.2", "libnccl.so", 0 };
  for (int i = 0; i < 3 && !w->nccl_lib; ++i)
    if (names[i] && *names[i]) w->nccl_lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!w->nccl_lib) xfail("comm: np=%d needs NCCL but libnccl.so.2 could not be loaded: %s", w->np, dlerror());
#define SYM(field, name)                                                      \
  do {                                                                        \
    *(void **)(&nc.field) = dlsym(w->nccl_lib, name);                         \
    if (!nc.field) xfail("comm: %s missing from NCCL library", name);         \
  } while (0)
  SYM(GetErrorString, "ncclGetErrorString");
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(AllReduce, "ncclAllReduce");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
#undef SYM
}

void *xg_world_nccl(struct xg_world *w)
{
  if (w->nccl_comm || w->np == 1) return w->nccl_comm;
  if (!w->has_cuda) xfail("comm: NCCL requested without a CUDA device");
  nccl_load(w);
  xg_nccl_uid uid, *all = xnew(xg_nccl_uid, w->np);
  memset(&uid, 0, sizeof uid);
  if (w->rank == 0) NC(nc.GetUniqueId(&uid));
  xg_host_allgather(w, &uid, sizeof uid, all);
  NC(nc.CommInitRank(&w->nccl_comm, w->np, all[0], w->rank));
  free(all);
  return w->nccl_comm;
}

/* nccl dtype / op codes (nccl.h: ncclInt32=2, ncclInt64=4, ncclFloat32=7,
   ncclFloat64=8; ncclSum=0, ncclProd=1, ncclMax=2, ncclMin=3) */
static int nccl_dtype(int xd)
{
  static const int t[4] = { 8, 7, 2, 4 };
  return t[xd];
}
static int nccl_op(int xop)
{
  switch (xop) {
    case XO_ADD: return 0;
    case XO_MUL: return 1;
    case XO_MAX: return 2;
    case XO_MIN: return 3;
  }
  xfail("gs: op %d has no NCCL reduction", xop);
  return 0;
}

void xg_nccl_group_start(struct xg_world *w) { (void)xg_world_nccl(w); NC(nc.GroupStart()); }
void xg_nccl_group_end(struct xg_world *w)   { (void)w; NC(nc.GroupEnd()); }
/* the communicator is created (and the library loaded) on first use: fetch it
   before the function pointer is read */
void xg_nccl_send(struct xg_world *w, const void *buf, size_t count, int xd, int peer, void *stream)
{ void *c = xg_world_nccl(w); NC(nc.Send(buf, count, nccl_dtype(xd), peer, c, stream)); }
void xg_nccl_recv(struct xg_world *w, void *buf, size_t count, int xd, int peer, void *stream)
{ void *c = xg_world_nccl(w); NC(nc.Recv(buf, count, nccl_dtype(xd), peer, c, stream)); }
void xg_nccl_allreduce(struct xg_world *w, void *buf, size_t count, int xd, int xop, void *stream)
{ void *c = xg_world_nccl(w); NC(nc.AllReduce(buf, buf, count, nccl_dtype(xd), nccl_op(xop), c, stream)); }

/* ------------------------------------------------------------------ */
/*  world                                                              */
/* ------------------------------------------------------------------ */
static void world_atexit(void)
{
  struct xg_world *w = g_world;
  if (!w) return;
  if (w->np > 1 && w->ctl && w->rank == 0) {
    char name[128];
    ctl_name(w, name, sizeof name);
    shm_unlink(name);
  }
}

struct xg_world *xg_world_get(void)
{
  if (g_world) return g_world;
  struct xg_world *w = xnew0(struct xg_world, 1);
  static const char *const rk[] = { "EXAGS_RANK", "RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID", 0 };
  static const char *const sz[] = { "EXAGS_NP", "WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS", 0 };
  static const char *const lr[] = { "EXAGS_LOCAL_RANK", "LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID", 0 };
  const char *v;
  w->rank = (v = env_first(rk)) ? atoi(v) : 0;
  w->np = (v = env_first(sz)) ? atoi(v) : 1;
  w->local_rank = (v = env_first(lr)) ? atoi(v) : w->rank;
  if (w->np < 1 || w->rank < 0 || w->rank >= w->np || w->np > XW_MAXNP)
    xfail("comm: bad rank/size %d/%d from the environment (max %d ranks)", w->rank, w->np, XW_MAXNP);
  exags_comm_gbl_id = (exags_uint)w->rank;
  exags_comm_gbl_np = (exags_uint)w->np;

  const char *host_only = getenv("EXAGS_HOST_SETUP_ONLY");
  int ndev = (host_only && atoi(host_only)) ? 0 : xcu_device_count();
  w->has_cuda = ndev > 0;
  w->device = -1;
  if (w->has_cuda) {
    w->device = g_device_override >= 0 ? g_device_override : w->local_rank % ndev;
    xcu_set_device(w->device);
    w->stream[0] = xcu_stream_create(0);
    w->stream[1] = xcu_stream_create(1);
    w->stream[2] = xcu_stream_create(1);   /* host->device copies of the pipelined host path */
    w->stream[3] = xcu_stream_create(1);   /* device->host copies */
    for (int i = 0; i < 4; ++i) w->event[i] = xcu_event_create();
  }
  if (w->np > 1) rendezvous_attach(w);
  if (w->np > 1 && w->has_cuda) {
    /* NCCL cannot place two ranks on one device: fall back to CUDA-IPC pulls */
    char mine[16], *all = xnew(char, 16 * (size_t)w->np);
    xcu_device_uuid(w->device, mine);
    xg_host_allgather(w, mine, 16, all);
    for (int p = 0; p < w->np; ++p)
      for (int q = p + 1; q < w->np; ++q)
        if (memcmp(all + 16 * p, all + 16 * q, 16) == 0) w->use_ipc = 1;
    /* push transport: every rank must be able to store into every other
       rank's device (NVLink / PCIe peer access); all ranks take the same
       decision */
    char ok = 1, *oks = xnew(char, w->np);
    for (int p = 0; p < w->np; ++p)
      if (p != w->rank && memcmp(all + 16 * p, mine, 16) != 0 && !xcu_peer_ok(w->device, all + 16 * p)) ok = 0;
    xg_host_allgather(w, &ok, 1, oks);
    for (int p = 0; p < w->np; ++p) ok &= oks[p];
    if (ok) {
      /* and CUDA IPC must work between the processes (it does not across some
         container boundaries): every rank maps a probe buffer of the next one */
      char h[64], *hs = xnew(char, 64 * (size_t)w->np);
      void *probe = xcu_malloc(256);
      xcu_ipc_export(h, probe);
      xg_host_allgather(w, h, 64, hs);
      void *q = xcu_ipc_try_open(hs + 64 * (size_t)((w->rank + 1) % w->np));
      ok = q != 0;
      if (q) xcu_ipc_close(q);
      xg_host_allgather(w, &ok, 1, oks);
      ok = 1;
      for (int p = 0; p < w->np; ++p) ok &= oks[p];
      xg_host_barrier(w);
      xcu_free(probe);
      free(hs);
    }
    free(oks);
    free(all);
    const char *ex = getenv("EXAGS_EXCHANGE");
    w->use_push = ok && !w->use_ipc;
    if (ex && strcmp(ex, "ipc") == 0) { w->use_ipc = 1; w->use_push = 0; }
    if (ex && strcmp(ex, "nccl") == 0) { w->use_ipc = 0; w->use_push = 0; }
    if (ex && strcmp(ex, "push") == 0) {
      /* on a shared device (tests) the kernels of different ranks time-slice;
         the all-reduce method then still needs the pull transport */
      if (!ok) xfail("comm: EXAGS_EXCHANGE=push but the devices of this job have no peer access");
      w->use_push = 1;
    }
    w->ipc_peer = xnew0(void *, w->np);
    if (w->use_push) { w->push_status = (int *)xcu_host_alloc_mapped(sizeof(int)); *w->push_status = 0; }
  }
  g_world = w;
  atexit(world_atexit);
  return w;
}

int xg_world_rank(const struct xg_world *w) { return w->rank; }
int xg_world_np(const struct xg_world *w) { return w->np; }
int xg_world_has_cuda(const struct xg_world *w) { return w->has_cuda; }
void *xg_world_stream(struct xg_world *w, int which) { return w->stream[which]; }
void *xg_world_event(struct xg_world *w, int which) { return w->event[which]; }

void *xg_world_scratch(struct xg_world *w, int which, size_t bytes)
{
  if (w->scratch_bytes[which] < bytes) {
    if (w->scratch[which]) { xcu_device_sync(); xcu_free(w->scratch[which]); }
    size_t want = bytes + bytes / 4 + 256;
    w->scratch[which] = xcu_malloc(want);
    w->scratch_bytes[which] = want;
  }
  return w->scratch[which];
}

int xg_world_ipc(const struct xg_world *w) { return w->use_ipc; }

void *xg_ipc_buffer(struct xg_world *w, size_t bytes)
{
  if (w->ipc_bytes >= bytes && w->ipc_mine) return w->ipc_mine;
  /* `bytes` is identical on every rank, so every rank grows in the same call */
  xcu_device_sync();
  xg_host_barrier(w);
  for (int p = 0; p < w->np; ++p) if (w->ipc_peer[p]) { xcu_ipc_close(w->ipc_peer[p]); w->ipc_peer[p] = 0; }
  xg_host_barrier(w);
  if (w->ipc_mine) xcu_free(w->ipc_mine);
  size_t want = bytes + bytes / 2 + 4096;
  w->ipc_mine = xcu_malloc(want);
  w->ipc_bytes = want;
  char h[64], *all = xnew(char, 64 * (size_t)w->np);
  xcu_ipc_export(h, w->ipc_mine);
  xg_host_allgather(w, h, 64, all);
  for (int p = 0; p < w->np; ++p)
    if (p != w->rank) w->ipc_peer[p] = xcu_ipc_open(all + 64 * p);
  free(all);
  xg_host_barrier(w);
  return w->ipc_mine;
}

void *xg_ipc_peer(struct xg_world *w, int rank) { return rank == w->rank ? w->ipc_mine : w->ipc_peer[rank]; }

int xg_world_push(const struct xg_world *w) { return w->use_push; }
int *xg_world_push_status(struct xg_world *w) { return w->push_status; }

static void mbox_close(struct xg_world *w, struct xg_mbox *mb)
{
  if (!mb->mine) return;
  /* nobody may still be storing into a mailbox that goes away */
  xcu_device_sync();
  xg_host_barrier(w);
  for (int p = 0; p < w->np; ++p) if (mb->peer && mb->peer[p]) { xcu_ipc_close(mb->peer[p]); mb->peer[p] = 0; }
  xg_host_barrier(w);
  xcu_free(mb->mine);
  mb->mine = 0; mb->data_bytes = 0;
}

int xg_mbox_ensure(struct xg_world *w, struct xg_mbox *mb, const u32 *nbr, u32 nnbr, size_t data_bytes)
{
  if (mb->mine && mb->data_bytes >= data_bytes) return 0;
  /* data_bytes is the same on every rank, so every rank is here in the same call */
  mbox_close(w, mb);
  if (!mb->peer) mb->peer = xnew0(void *, w->np);
  size_t want = (data_bytes + data_bytes / 4 + 255) & ~(size_t)255;
  mb->mine = xcu_malloc(XG_MBOX_HDR + 2 * want);
  mb->data_bytes = want;
  xcu_memset(mb->mine, 0, XG_MBOX_HDR, 0);
  xcu_device_sync();
  char h[64], *all = xnew(char, 64 * (size_t)w->np);
  xcu_ipc_export(h, mb->mine);
  xg_host_allgather(w, h, 64, all);
  for (u32 k = 0; k < nnbr; ++k)
    if ((int)nbr[k] != w->rank) mb->peer[nbr[k]] = xcu_ipc_open(all + 64 * nbr[k]);
  free(all);
  mb->epoch = 0;
  xg_host_barrier(w);
  return 1;
}

void xg_mbox_free(struct xg_world *w, struct xg_mbox *mb)
{
  mbox_close(w, mb);
  free(mb->peer); mb->peer = 0;
  if (mb->d_tab) { xcu_free(mb->d_tab); mb->d_tab = 0; }
}

const char *exags_exchange_transport(void)
{
  struct xg_world *w = xg_world_get();
  if (w->np == 1) return "single";
  return w->use_push ? "push" : w->use_ipc ? "ipc" : "nccl";
}

void exags_set_device(int device) { g_device_override = device; }
int exags_get_device(void) { return g_world ? g_world->device : g_device_override; }

/* ------------------------------------------------------------------ */
/*  public communicator API                                            */
/* ------------------------------------------------------------------ */
void exags_comm_world(comm_ptr *cpp)
{
  if (!cpp) return;
  struct xg_world *w = xg_world_get();
  comm_ptr cp = xnew0(struct comm, 1);
  cp->h = 0;
  cp->id = w->rank;
  cp->np = w->np;
  cp->priv = w;
  cp->ref_count = xnew0(int, 1);
  *cpp = cp;
}

void exags_comm_init(void)
{
  if (!exags_glb_comm) exags_comm_world(&exags_glb_comm);
}

void exags_comm_finalize(void) {}

void exags_comm_dup(comm_ptr *cpp, const comm_ptr cp)
{
  if (!cpp || !cp) return;
  comm_ptr d = xnew(struct comm, 1);
  *d = *cp;
  ++*d->ref_count;
  *cpp = d;
}

void exags_comm_free(comm_ptr *cpp)
{
  if (!cpp || !*cpp) return;
  comm_ptr cp = *cpp;
  if (*cp->ref_count == 0) free(cp->ref_count);
  else --*cp->ref_count;
  free(cp);
  *cpp = 0;
}

void exags_comm_np(const comm_ptr cp, int *np) { if (cp && np) *np = cp->np; }
void exags_comm_id(const comm_ptr cp, int *id) { if (cp && id) *id = cp->id; }

void exags_comm_time(double *tm)
{
  if (!tm) return;
  struct timeval tv;
  gettimeofday(&tv, 0);
  *tm = (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}

void exags_comm_barrier(const comm_ptr cp)
{
  if (!cp) return;
  xg_host_barrier((struct xg_world *)cp->priv);
}

void exags_comm_bcast(const comm_ptr cp, void *p, size_t n, exags_uint root)
{
  if (!cp || !p || cp->np == 1) return;
  struct xg_world *w = (struct xg_world *)cp->priv;
  char *all = (char *)xmalloc_(n * (size_t)cp->np, __FILE__, __LINE__);
  xg_host_allgather(w, p, n, all);
  memcpy(p, all + (size_t)root * n, n);
  free(all);
}

/* host-side element-wise reductions, rank order 0..np-1 on every rank so all
   ranks get bit-identical results */
#define HOST_COMBINE(T, OPC)                                                  \
  do {                                                                        \
    T *a_ = (T *)acc; const T *b_ = (const T *)in;                            \
    for (size_t k_ = 0; k_ < n; ++k_) { T x = a_[k_], y = b_[k_]; OPC; a_[k_] = x; } \
  } while (0)

static void host_combine(void *acc, const void *in, size_t n, int xd, int xop)
{
#define BY_OP(T)                                                              \
  switch (xop) {                                                              \
    case XO_ADD: HOST_COMBINE(T, x += y); break;                              \
    case XO_MUL: HOST_COMBINE(T, x *= y); break;                              \
    case XO_MIN: HOST_COMBINE(T, if (y < x) x = y); break;                    \
    case XO_MAX: HOST_COMBINE(T, if (y > x) x = y); break;                    \
    default: xfail("comm: op %d not supported in host reductions", xop);      \
  }
  switch (xd) {
    case XD_F64: BY_OP(double); break;
    case XD_F32: BY_OP(float); break;
    case XD_I32: BY_OP(int); break;
    case XD_I64: BY_OP(long long); break;
    default: xfail("comm: bad domain %d", xd);
  }
#undef BY_OP
}

static void host_identity(void *out, size_t n, int xd, int xop)
{
#define FILL(T, lo, hi)                                                       \
  do {                                                                        \
    T e = xop == XO_MUL ? (T)1 : xop == XO_MIN ? (T)(hi) : xop == XO_MAX ? (T)(lo) : (T)0; \
    for (size_t k_ = 0; k_ < n; ++k_) ((T *)out)[k_] = e;                     \
  } while (0)
  switch (xd) {
    case XD_F64: FILL(double, -DBL_MAX, DBL_MAX); break;
    case XD_F32: FILL(float, -FLT_MAX, FLT_MAX); break;
    case XD_I32: FILL(int, INT_MIN, INT_MAX); break;
    case XD_I64: FILL(long long, LLONG_MIN, LLONG_MAX); break;
    default: xfail("comm: bad domain %d", xd);
  }
#undef FILL
}

/* v may be a host vector (the reference's contract) or a device vector: small
   device vectors are folded on the host in rank order like host ones (bit-
   identical on every rank), large ones go through ncclAllReduce in place.     */
void exags_comm_allreduce(const comm_ptr cp, gs_dom dom, gs_op_t op, void *v, exags_uint vn, void *buf)
{
  (void)buf;
  if (!cp || !vn || cp->np == 1) return;
  int xd = xg_dom_of((int)dom);
  if (xd < 0) xfail("comm_allreduce: bad domain %d", (int)dom);
  struct xg_world *w = (struct xg_world *)cp->priv;
  size_t bytes = (size_t)vn * xg_dom_bytes(xd);
  const int on_dev = w->has_cuda && xcu_is_device_ptr(v);
  if (on_dev && bytes > (1u << 20) && !w->use_ipc && (int)op <= XO_MAX) {
    void *s = w->stream[1];
    xcu_device_sync();
    xg_nccl_allreduce(w, v, vn, xd, (int)op, s);
    xcu_stream_sync(s);
    return;
  }
  char *all = (char *)xmalloc_(bytes * ((size_t)cp->np + 1), __FILE__, __LINE__);
  char *mine = all + bytes * (size_t)cp->np;
  if (on_dev) { xcu_device_sync(); xcu_d2h(mine, v, bytes, 0); xcu_device_sync(); }
  else memcpy(mine, v, bytes);
  xg_host_allgather(w, mine, bytes, all);
  memcpy(mine, all, bytes);
  for (int p = 1; p < cp->np; ++p) host_combine(mine, all + (size_t)p * bytes, vn, xd, (int)op);
  if (on_dev) { xcu_h2d(v, mine, bytes, 0); xcu_device_sync(); }
  else memcpy(v, mine, bytes);
  free(all);
}

void exags_comm_scan(void *scan, const comm_ptr cp, gs_dom dom, gs_op_t op,
                     const void *v, exags_uint vn, void *buffer)
{
  (void)buffer;
  if (!cp || !vn) return;
  int xd = xg_dom_of((int)dom);
  if (xd < 0) xfail("comm_scan: bad domain %d", (int)dom);
  struct xg_world *w = (struct xg_world *)cp->priv;
  size_t bytes = (size_t)vn * xg_dom_bytes(xd);
  /* v and scan may each be a host vector (the reference's contract) or a device
     vector: scans carry a few counters per rank, so device vectors are folded
     on the host in rank order like host ones */
  const int v_dev = w->has_cuda && xcu_is_device_ptr(v);
  const int s_dev = w->has_cuda && xcu_is_device_ptr(scan);
  char *all = (char *)xmalloc_(bytes * ((size_t)cp->np + 3), __FILE__, __LINE__);
  char *mine = all + bytes * (size_t)cp->np, *out = mine + bytes;
  if (v_dev) { xcu_device_sync(); xcu_d2h(mine, v, bytes, 0); xcu_device_sync(); }
  else memcpy(mine, v, bytes);
  xg_host_allgather(w, mine, bytes, all);
  char *excl = s_dev ? out : (char *)scan, *tot = excl + bytes;
  host_identity(excl, vn, xd, (int)op);
  host_identity(tot, vn, xd, (int)op);
  for (int p = 0; p < cp->np; ++p) {
    if (p < cp->id) host_combine(excl, all + (size_t)p * bytes, vn, xd, (int)op);
    host_combine(tot, all + (size_t)p * bytes, vn, xd, (int)op);
  }
  if (s_dev) { xcu_h2d(scan, out, 2 * bytes, 0); xcu_device_sync(); }
  free(all);
}

double exags_comm_dot(const comm_ptr cp, double *v, double *w, exags_uint n)
{
  double s = 0, b;
  struct xg_world *wd = cp ? (struct xg_world *)cp->priv : 0;
  if (wd && wd->has_cuda && n && xcu_is_device_ptr(v)) {
    /* device vectors: HBM-bound two-pass reduction in a fixed order */
    if (!xcu_is_device_ptr(w)) xfail("comm_dot: v is a device vector, w is not");
    if (((uintptr_t)v | (uintptr_t)w) & 15u) xfail("comm_dot: device vectors must be 16-byte aligned");
    double *scr = (double *)xg_world_scratch(wd, 3, ((size_t)xcu_dot_parts() + 1) * sizeof(double));
    void *st = wd->stream[0];
    xcu_dot(st, v, w, n, scr + 1, scr);
    xcu_d2h(&s, scr, sizeof s, st);
    xcu_stream_sync(st);
  } else
    for (exags_uint i = 0; i < n; ++i) s += v[i] * w[i];
  exags_comm_allreduce(cp, gs_double, gs_add, &s, 1, &b);
  return s;
}

/* Stream-ordered comm_dot (extension, include/exags_cuda.h): v, w and the
   result are device memory, nothing returns to the host, the work is ordered on
   `stream` -- so a CG iteration can chain gs_async, the dot and the vector
   updates that consume it without a host round trip (and capture them in a
   CUDA graph when np == 1).  Across ranks the partial sums are combined by
   ncclAllReduce on the same stream.                                         */
void exags_comm_dot_async(const comm_ptr cp, const double *v, const double *w, exags_uint n,
                          double *d_out, void *stream)
{
  struct xg_world *wd = cp ? (struct xg_world *)cp->priv : 0;
  if (!wd || !wd->has_cuda) xfail("comm_dot_async: no CUDA device");
  if (((uintptr_t)v | (uintptr_t)w) & 15u) xfail("comm_dot_async: device vectors must be 16-byte aligned");
  if (!wd->dot_scratch) wd->dot_scratch = xcu_malloc(((size_t)xcu_dot_parts() + 1) * sizeof(double));
  xcu_dot(stream, v, w, n, (double *)wd->dot_scratch, d_out);
  if (wd->np > 1) {
    if (wd->use_ipc) xfail("comm_dot_async: ranks sharing one GPU have no stream-ordered all-reduce; use comm_dot");
    xg_nccl_allreduce(wd, d_out, 1, XD_F64, XO_ADD, stream);
  }
}

/* T comm_reduce__T(cp, op, in, n): fold a local host array, then across ranks
   (src/comm.upc:961-984).                                              */
#define DEFINE_REDUCE(T, NAME, DOM)                                           \
  T exags_comm_reduce__##NAME(const comm_ptr cp, gs_op_t op, const T *in, exags_uint n) \
  {                                                                           \
    T accum, buf;                                                             \
    host_identity(&accum, 1, xg_dom_of(DOM), (int)op);                        \
    for (exags_uint i = 0; i < n; ++i) host_combine(&accum, in + i, 1, xg_dom_of(DOM), (int)op); \
    exags_comm_allreduce(cp, (gs_dom)DOM, op, &accum, 1, &buf);               \
    return accum;                                                             \
  }
DEFINE_REDUCE(double, double, gs_double)
DEFINE_REDUCE(float, float, gs_float)
DEFINE_REDUCE(int, int, gs_int)
DEFINE_REDUCE(long, long, gs_long)
DEFINE_REDUCE(long long, long_long, 4)
#undef DEFINE_REDUCE

/* ------------------------------------------------------------------ */
/*  element type tags and Fortran bindings (src/comm.upc:376-407,       */
/*  804-827, 1031-1148)                                                 */
/* ------------------------------------------------------------------ */
enum { XCT_INT = 4, XCT_INT8 = 8, XCT_REAL = 10, XCT_DP = 11 };

void exags_comm_type_int(comm_type *ct)  { if (ct) *ct = XCT_INT; }
void exags_comm_type_int8(comm_type *ct) { if (ct) *ct = XCT_INT8; }
void exags_comm_type_real(comm_type *ct) { if (ct) *ct = XCT_REAL; }
void exags_comm_type_dp(comm_type *ct)   { if (ct) *ct = XCT_DP; }
void exags_comm_tag_ub(const comm_ptr cp, int *ub) { if (cp && ub) *ub = 0; }

void exags_comm_allreduce_cdom(const comm_ptr cp, comm_type cdom, gs_op_t op, void *v,
                               exags_uint vn, void *buf)
{
  if (!cp || !v || !buf) return;
  gs_dom dom;
  switch (cdom) {
    case XCT_INT:  dom = gs_int; break;
    case XCT_INT8: dom = gs_long; break;
    case XCT_REAL: dom = gs_float; break;
    case XCT_DP:   dom = gs_double; break;
    default: xfail("comm_allreduce_cdom: cannot identify cdom P=%u.", (unsigned)cdom); return;
  }
  exags_comm_allreduce(cp, dom, op, v, vn, buf);
}

#define FCOMM(name) name##_
#define FCOMM2(ret, name, args, body)                                         \
  ret FCOMM(name) args body                                                   \
  extern __typeof__(FCOMM(name)) fgslib_##name##_ __attribute__((alias(#name "_")));

FCOMM2(void, comm_init, (void), { exags_comm_init(); })
FCOMM2(void, comm_finalize, (void), { exags_comm_finalize(); })
FCOMM2(void, comm_world, (comm_ptr *cpp), { exags_comm_world(cpp); })
FCOMM2(void, comm_free, (comm_ptr *cpp), { exags_comm_free(cpp); })
FCOMM2(void, comm_np, (const comm_ptr *cpp, int *np), { exags_comm_np(*cpp, np); })
FCOMM2(void, comm_id, (const comm_ptr *cpp, int *id), { exags_comm_id(*cpp, id); })
FCOMM2(void, comm_type_int, (comm_type *ct), { exags_comm_type_int(ct); })
FCOMM2(void, comm_type_int8, (comm_type *ct), { exags_comm_type_int8(ct); })
FCOMM2(void, comm_type_real, (comm_type *ct), { exags_comm_type_real(ct); })
FCOMM2(void, comm_type_dp, (comm_type *ct), { exags_comm_type_dp(ct); })
FCOMM2(void, comm_tag_ub, (const comm_ptr *cpp, int *ub), { exags_comm_tag_ub(*cpp, ub); })
FCOMM2(void, comm_time, (double *tm), { exags_comm_time(tm); })
FCOMM2(void, comm_barrier, (const comm_ptr *cpp), { exags_comm_barrier(*cpp); })
FCOMM2(void, comm_bcast, (const comm_ptr *cpp, void *p, size_t *n, exags_uint *root), { exags_comm_bcast(*cpp, p, *n, *root); })
FCOMM2(void, comm_allreduce_add, (const comm_ptr *cpp, comm_type *ct, void *v, exags_uint *vn, void *buf),
       { exags_comm_allreduce_cdom(*cpp, *ct, gs_add, v, *vn, buf); })
FCOMM2(void, comm_allreduce_min, (const comm_ptr *cpp, comm_type *ct, void *v, exags_uint *vn, void *buf),
       { exags_comm_allreduce_cdom(*cpp, *ct, gs_min, v, *vn, buf); })
FCOMM2(void, comm_allreduce_max, (const comm_ptr *cpp, comm_type *ct, void *v, exags_uint *vn, void *buf),
       { exags_comm_allreduce_cdom(*cpp, *ct, gs_max, v, *vn, buf); })
FCOMM2(void, comm_allreduce_mul, (const comm_ptr *cpp, comm_type *ct, void *v, exags_uint *vn, void *buf),
       { exags_comm_allreduce_cdom(*cpp, *ct, gs_mul, v, *vn, buf); })
