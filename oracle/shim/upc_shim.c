/* upc_shim.c -- runtime of the UPC stand-in (see upc.h): a shared arena,
 * a sense-reversing barrier, bump allocation and fork()-based "threads".
 * TEST INFRASTRUCTURE ONLY. */
#define _GNU_SOURCE
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/prctl.h>
#include <sys/wait.h>
#include <unistd.h>
#include "upc.h"

int shim_mythread = 0, shim_threads = 1;

struct shim_lock { volatile int held; };
struct arena_hdr {
  volatile unsigned bar_count, bar_sense;
  volatile size_t brk;
  void *volatile bcast;
  size_t size;
};
static struct arena_hdr *A = 0;
static unsigned my_sense = 0;

void shim_unsupported(const char *what)
{ fprintf(stderr, "upc shim: %s is not provided\n", what); abort(); }

void shim_poll(void)
{
  static unsigned n = 0;
  __sync_synchronize();
  if ((++n & 63u) == 0) sched_yield(); else __asm__ __volatile__("pause");
}

void shim_arena_create(size_t bytes)
{
  void *p = mmap(0, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (p == MAP_FAILED) { perror("upc shim: mmap"); abort(); }
  A = (struct arena_hdr *)p;
  A->bar_count = 0; A->bar_sense = 0; A->brk = 4096; A->bcast = 0; A->size = bytes;
  my_sense = 0;
}

void shim_arena_destroy(void)
{ if (A) munmap((void *)A, A->size); A = 0; }

void *shim_arena_raw(size_t bytes)      /* usable before the fork as well */
{
  size_t need = (bytes + 63u) & ~(size_t)63u;
  size_t at = __sync_fetch_and_add(&A->brk, need);
  if (at + need > A->size) { fprintf(stderr, "upc shim: arena exhausted (%zu + %zu > %zu)\n", at, need, A->size); abort(); }
  return (char *)A + at;        /* anonymous mappings are zero-filled */
}

void *shim_alloc(size_t nbytes) { return shim_arena_raw(nbytes ? nbytes : 1); }

void shim_barrier(void)
{
  if (shim_threads == 1) return;
  unsigned mine = (my_sense ^= 1u);
  __sync_synchronize();
  if (__sync_fetch_and_add(&A->bar_count, 1u) == (unsigned)shim_threads - 1u) {
    A->bar_count = 0;
    __sync_synchronize();
    A->bar_sense = mine;
  } else {
    unsigned n = 0;
    while (A->bar_sense != mine) { if ((++n & 63u) == 0) sched_yield(); else __asm__ __volatile__("pause"); }
  }
  __sync_synchronize();
}

void *shim_all_alloc(size_t nblocks, size_t nbytes)
{
  shim_barrier();
  if (shim_mythread == 0) A->bcast = shim_arena_raw(nblocks * nbytes);
  shim_barrier();
  void *p = A->bcast;
  shim_barrier();
  return p;
}

void shim_all_broadcast(void *dst, const void *src, size_t nbytes)
{
  shim_barrier();
  char *mine = (char *)dst + (size_t)shim_mythread * nbytes;
  if ((const void *)mine != src) memmove(mine, src, nbytes);
  shim_barrier();
}

upc_lock_t *shim_lock_alloc(void) { return (upc_lock_t *)shim_arena_raw(sizeof(struct shim_lock)); }
void shim_lock(upc_lock_t *l) { while (__sync_lock_test_and_set(&l->held, 1)) shim_poll(); }
void shim_unlock(upc_lock_t *l) { __sync_lock_release(&l->held); }

/* run fn(rank, arg) as `threads` forked processes; returns 0 if all exited 0 */
int shim_run(int threads, void (*fn)(int, void *), void *arg)
{
  pid_t *pid = (pid_t *)calloc((size_t)threads, sizeof *pid);
  int bad = 0;
  fflush(stdout); fflush(stderr);
  for (int r = 0; r < threads; ++r) {
    pid[r] = fork();
    if (pid[r] < 0) { perror("upc shim: fork"); abort(); }
    if (pid[r] == 0) {
      prctl(PR_SET_PDEATHSIG, SIGKILL);     /* never outlive the harness */
      alarm(1800);                          /* nor spin forever on a lost flag */
      shim_mythread = r; shim_threads = threads; my_sense = 0;
      fn(r, arg);
      fflush(stdout);
      _exit(0);
    }
  }
  /* a thread that dies (arena exhausted, fail()) leaves the others spinning
     on a barrier: reap in any order and take the rest down with it */
  for (int left = threads; left > 0; --left) {
    int st = 0;
    pid_t p = waitpid(-1, &st, 0);
    if (p < 0) { bad = 1; break; }
    int mine = 0;
    for (int r = 0; r < threads; ++r) if (pid[r] == p) { pid[r] = 0; mine = 1; }
    if (!mine) { ++left; continue; }
    if ((!WIFEXITED(st) || WEXITSTATUS(st) != 0) && !bad) {
      bad = 1;
      for (int r = 0; r < threads; ++r) if (pid[r] > 0) kill(pid[r], SIGKILL);
    }
  }
  free(pid);
  return bad;
}
