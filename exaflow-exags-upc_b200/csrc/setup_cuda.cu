// setup_cuda.cu -- gs_setup's label sort, group discovery and map construction
// on the device: the CUDA backend of setup_pipeline.h plus the thin C ABI the
// host C code calls it through.
//
// The passes themselves (setup_pipeline.h) are this library's own kernels, one
// item per thread; the stable radix sort of the (label, index) rows, the
// prefix sums and the two reductions are CUB device primitives from the CUDA
// toolkit -- library code, used here as gs_setup's sort (the reference's own is
// sort_imp.h's radix/merge hybrid, src/sort_imp.h:446-509), not on the per-call
// path.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "setup_pipeline.h"
#include "gs_plan.h"

#define CU(call)                                                              \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess)                                                    \
      exags_fail(1, __FILE__, __LINE__, "CUDA error: %s (%s)",                \
                 cudaGetErrorString(e_), #call);                              \
  } while (0)

template <class F>
__global__ void __launch_bounds__(256) k_setup_for_each(size_t n, F f)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}

namespace {

// One backend per process, created on first use and never destroyed (its
// members would outlive the CUDA context at exit).  Work arrays come from a
// stream-ordered memory pool of the library's own: between begin() and end() the pool
// keeps what the passes free, so the ~40 arrays of one gs_setup reuse each
// other's memory instead of going to the driver; end() hands everything that
// is not part of the finished handle back.  Arrays that survive (the maps of
// the handle) are ordinary allocations of the same pool and are released with
// cudaFree like any other.
struct CudaBackend {
  cudaStream_t st = 0;
  void *tmp = nullptr;            // CUB scratch, grow-only until end()
  size_t tmp_bytes = 0;
  u64 *res = nullptr;             // reduction result (device) + pinned mirror
  u64 *res_h = nullptr;
  unsigned long long launches = 0;
  int sms = 148;
  int dev = -1;
  int depth = 0;
  cudaMemPool_t pool = nullptr;   // the library's own pool (not the device's default one)
  bool pool_own = false;

  void begin()
  {
    if (depth++) return;
    int d = 0;
    CU(cudaGetDevice(&d));
    if (d != dev) {
      /* a memory pool and a stream of the library's own, per device: gs_setup
         neither trims nor re-tunes the application's default pool, and its
         passes do not run on the legacy default stream.  The stream is a
         blocking one, so labels the caller produced on the default stream are
         ordered before the first pass without an explicit synchronisation.   */
      if (res) { cudaFree(res); res = nullptr; }
      if (pool_own) { cudaStreamDestroy(st); cudaMemPoolDestroy(pool); pool_own = false; }
      dev = d;
      CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      cudaMemPoolProps props;
      memset(&props, 0, sizeof props);
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      CU(cudaMemPoolCreate(&pool, &props));
      unsigned long long keep = ~0ull;             // freed blocks stay for the passes that follow
      CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
      CU(cudaStreamCreateWithFlags(&st, cudaStreamDefault));
      pool_own = true;
      if (!res_h) CU(cudaMallocHost(&res_h, 64));
    }
    if (!res) CU(cudaMalloc(&res, 64));
  }
  void end()
  {
    if (--depth) return;
    CU(cudaStreamSynchronize(st));
    if (tmp) { CU(cudaFreeAsync(tmp, st)); tmp = nullptr; tmp_bytes = 0; }
    CU(cudaStreamSynchronize(st));
    CU(cudaMemPoolTrimTo(pool, 0));                // what the handle keeps stays allocated; the rest goes back
    xcu_note_launches(launches);
    launches = 0;
  }

  template <class T> T *alloc(size_t count)
  {
    void *p = nullptr;
    CU(cudaMallocFromPoolAsync(&p, (count ? count : 1) * sizeof(T), pool, st));
    return (T *)p;
  }
  void release(void *p)
  {
    if (p) CU(cudaFreeAsync(p, st));
  }
  void zero(void *p, size_t bytes) { CU(cudaMemsetAsync(p, 0, bytes, st)); }
  void fill_ff(void *p, size_t bytes) { CU(cudaMemsetAsync(p, 0xFF, bytes, st)); }

  template <class F> void for_each(size_t n, F f)
  {
    if (!n) return;
    const unsigned threads = 256;
    size_t blocks = (n + threads - 1) / threads;
    const size_t cap = (size_t)sms * 32;        // grid-stride beyond 32 CTAs per SM
    if (blocks > cap) blocks = cap;
    k_setup_for_each<<<(unsigned)blocks, threads, 0, st>>>(n, f);
    CU(cudaGetLastError());
    ++launches;
  }
  void need_tmp(size_t bytes)
  {
    if (bytes <= tmp_bytes) return;
    if (tmp) CU(cudaFreeAsync(tmp, st));
    CU(cudaMallocFromPoolAsync(&tmp, bytes, pool, st));
    tmp_bytes = bytes;
  }
  void sort_pairs(u64 *&k, u64 *&k2, u32 *&v, u32 *&v2, size_t n, unsigned bits)
  {
    if (n < 2) return;
    cub::DoubleBuffer<u64> dk(k, k2);
    cub::DoubleBuffer<u32> dv(v, v2);
    size_t bytes = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, (unsigned long long)n, 0, (int)bits, st));
    need_tmp(bytes);
    CU(cub::DeviceRadixSort::SortPairs(tmp, bytes, dk, dv, (unsigned long long)n, 0, (int)bits, st));
    ++launches;
    if (dk.Current() != k) { u64 *t = k; k = k2; k2 = t; }
    if (dv.Current() != v) { u32 *t = v; v = v2; v2 = t; }
  }
  void scan(u32 *p, size_t n)
  {
    if (!n) return;
    size_t bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, bytes, p, p, (unsigned long long)n, st));
    need_tmp(bytes);
    CU(cub::DeviceScan::ExclusiveSum(tmp, bytes, p, p, (unsigned long long)n, st));
    ++launches;
  }
  u64 max_u64(const u64 *p, size_t n)
  {
    size_t bytes = 0;
    CU(cub::DeviceReduce::Max(nullptr, bytes, p, res, (unsigned long long)n, st));
    need_tmp(bytes);
    CU(cub::DeviceReduce::Max(tmp, bytes, p, res, (unsigned long long)n, st));
    ++launches;
    CU(cudaMemcpyAsync(res_h, res, sizeof(u64), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return res_h[0];
  }
  u64 sum_u32(const u32 *p, size_t n)
  {
    size_t bytes = 0;
    CU(cub::DeviceReduce::Sum(nullptr, bytes, p, res, (unsigned long long)n, st));
    need_tmp(bytes);
    CU(cub::DeviceReduce::Sum(tmp, bytes, p, res, (unsigned long long)n, st));
    ++launches;
    CU(cudaMemcpyAsync(res_h, res, sizeof(u64), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return res_h[0];
  }
  u32 get(const u32 *p)
  {
    CU(cudaMemcpyAsync(res_h, p, sizeof(u32), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return *(const u32 *)res_h;
  }
};

CudaBackend &backend()
{
  static CudaBackend *be = new CudaBackend;
  return *be;
}
struct Scope {                    // begin()/end() around one C-ABI call (nests)
  CudaBackend &be;
  Scope() : be(backend()) { be.begin(); }
  ~Scope() { be.end(); }
};

void check(int rc, const char *what)
{
  switch (rc) {
    case xgsetup::OK: return;
    case xgsetup::ERR_ID_TOO_LARGE: exags_fail(1, __FILE__, __LINE__, "gs_setup: |id| is too large (%s)", what); break;
    case xgsetup::ERR_MAP_TOO_LARGE: exags_fail(1, __FILE__, __LINE__, "gs_setup: local map too large (%s)", what); break;
    case xgsetup::ERR_SELL_TOO_LARGE: exags_fail(1, __FILE__, __LINE__, "gs_setup: map too large for the sliced layout (%s)", what); break;
    default: exags_fail(1, __FILE__, __LINE__, "gs_setup: internal error packing the sliced layout (%s)", what);
  }
}

bool is_device_ptr(const void *p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

}  // namespace

extern "C" {

void xcu_setup_topo(const void *id, int idbytes, size_t n, struct xcu_topo *t)
{
  if (n >= (size_t)XG_NONE)
    exags_fail(1, __FILE__, __LINE__, "gs_setup: local size %lu does not fit the 32-bit index type", (unsigned long)n);
  Scope sc; CudaBackend &be = sc.be;
  const void *d_id = id;
  void *own = nullptr;
  if (n && !is_device_ptr(id)) {
    own = be.alloc<char>(n * (size_t)idbytes);
    CU(cudaMemcpyAsync(own, id, n * (size_t)idbytes, cudaMemcpyHostToDevice, be.st));
    d_id = own;
  }
  u64 maxabs = 0;
  const int rc = xgsetup::topo_build(be, d_id, idbytes, n, t, &maxabs);
  if (own) be.release(own);
  if (rc == xgsetup::ERR_ID_TOO_LARGE)
    exags_fail(1, __FILE__, __LINE__, "gs_setup: |id| = %llu is too large", (unsigned long long)maxabs);
  check(rc, "topology");
  CU(cudaStreamSynchronize(be.st));
}

void xcu_setup_topo_free(struct xcu_topo *t)
{
  Scope sc; CudaBackend &be = sc.be;
  xgsetup::topo_release(be, t);
}

void xcu_setup_topo_to_host(const struct xcu_topo *dt, struct xg_topo *ht)
{
  memset(ht, 0, sizeof *ht);
  const size_t nnz = dt->nnz, ngrp = dt->ngrp;
  ht->nnz = nnz; ht->ngrp = ngrp;
  ht->id = xnew(u64, ngrp ? ngrp : 1);
  ht->idx = xnew(u32, nnz ? nnz : 1);
  ht->flag = xnew(unsigned char, nnz ? nnz : 1);
  ht->gstart = xnew(u32, ngrp + 1);
  if (ngrp) CU(cudaMemcpy(ht->id, dt->id, ngrp * sizeof(u64), cudaMemcpyDeviceToHost));
  if (nnz) CU(cudaMemcpy(ht->idx, dt->idx, nnz * sizeof(u32), cudaMemcpyDeviceToHost));
  if (nnz) CU(cudaMemcpy(ht->flag, dt->flag, nnz, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(ht->gstart, dt->gstart, (ngrp + 1) * sizeof(u32), cudaMemcpyDeviceToHost));
}

void xcu_setup_local(const struct xcu_topo *t, struct xcu_dmap *hall, struct xcu_dmap *in, u32 **fp, u32 *nfp,
                     u32 **fs, u32 *nfs, u64 *sectors)
{
  Scope sc; CudaBackend &be = sc.be;
  check(xgsetup::local_maps(be, t, hall, in, fp, nfp, fs, nfs, sectors), "local maps");
  CU(cudaStreamSynchronize(be.st));
}

void xcu_setup_split(const struct xcu_topo *t, const u32 *bgrp_host, u32 nb, struct xcu_dmap *hall,
                     struct xcu_dmap *in, struct xcu_dmap *bnd, struct xcu_dmap *bnd2,
                     u32 **fp, u32 *nfp, u32 **fs, u32 *nfs, u64 *sectors)
{
  Scope sc; CudaBackend &be = sc.be;
  u32 *bgrp = be.alloc<u32>((size_t)nb + 1);
  if (nb) CU(cudaMemcpyAsync(bgrp, bgrp_host, (size_t)nb * sizeof(u32), cudaMemcpyHostToDevice, be.st));
  check(xgsetup::split_maps(be, t, bgrp, nb, hall, in, bnd, bnd2, fp, nfp, fs, nfs, sectors), "interior/boundary split");
  be.release(bgrp);
  CU(cudaStreamSynchronize(be.st));
}

/* np > 1: the labels of this rank that another rank may hold too, chosen in
   HBM (setup_pipeline.h) with the two filters of remote.c -- the ranks trade
   an 8 KB range-occupancy bitmap and a hashed presence bitmap through the host
   rendezvous -- so that only the candidates (the interface labels of a
   partitioned mesh, well under 1 % of the topology) come back to the host.  */
void xcu_setup_candidates(const struct xcu_topo *t, struct xg_world *w, struct xg_cand *out)
{
  Scope sc; CudaBackend &be = sc.be;
  const int np = xg_world_np(w), me = xg_world_rank(w);
  memset(out, 0, sizeof *out);
  const size_t ngrp = t->ngrp;
  u64 lmax = ngrp ? be.max_u64(t->id, ngrp) : 0, gmax = 0;
  u64 *all = xnew(u64, (size_t)np > 2 ? (size_t)np : 2);
  xg_host_allgather(w, &lmax, sizeof lmax, all);
  for (int p = 0; p < np; ++p) if (all[p] > gmax) gmax = all[p];
  unsigned shift = 0;
  while ((gmax >> shift) >= (u64)xgsetup::CAND_BINS) ++shift;

  /* stage 1: coarse bins */
  const size_t bw = xgsetup::CAND_BINS / 64;
  u32 *bins = be.alloc<u32>(2 * bw);
  be.zero(bins, 2 * bw * sizeof(u32));
  xgsetup::cand_bins(be, t, shift, bins);
  u64 *hb = xnew(u64, bw), *allb = xnew(u64, (size_t)np * bw);
  CU(cudaMemcpyAsync(hb, bins, bw * sizeof(u64), cudaMemcpyDeviceToHost, be.st));
  CU(cudaStreamSynchronize(be.st));
  xg_host_allgather(w, hb, bw * sizeof(u64), allb);
  memset(hb, 0, bw * sizeof(u64));
  for (int p = 0; p < np; ++p)
    if (p != me) for (size_t k = 0; k < bw; ++k) hb[k] |= allb[(size_t)p * bw + k];
  CU(cudaMemcpyAsync(bins, hb, bw * sizeof(u64), cudaMemcpyHostToDevice, be.st));
  u32 *c32 = be.alloc<u32>(ngrp + 1);
  xgsetup::cand_from_bins(be, t, shift, bins, c32);
  u64 npass = ngrp ? be.sum_u32(c32, ngrp) : 0, maxn = 1;
  free(hb); free(allb);
  be.release(bins);

  /* stage 2: hashed presence bitmap of what passed */
  xg_host_allgather(w, &npass, sizeof npass, all);
  for (int p = 0; p < np; ++p) if (all[p] > maxn) maxn = all[p];
  free(all);
  unsigned bits = 20;
  while (bits < 34 && ((u64)1 << bits) < 16 * maxn) ++bits;          /* <= 1/16 full: ~6 % false positives */
  const size_t words = (size_t)1 << (bits - 6);
  u32 *bm = be.alloc<u32>(2 * words);
  be.zero(bm, 2 * words * sizeof(u32));
  xgsetup::cand_presence(be, t, c32, bits, bm);
  u64 *mine = xg_host_bitmap_new(w, words), *others = xnew(u64, words);
  CU(cudaMemcpyAsync(mine, bm, words * sizeof(u64), cudaMemcpyDeviceToHost, be.st));
  CU(cudaStreamSynchronize(be.st));
  xg_host_bitmap_or_others(w, mine, words, others);                   /* releases mine */
  CU(cudaMemcpyAsync(bm, others, words * sizeof(u64), cudaMemcpyHostToDevice, be.st));
  xgsetup::cand_refine(be, t, c32, bits, bm);
  CU(cudaStreamSynchronize(be.st));
  free(others);
  be.release(bm);

  u64 *d_id = nullptr; u32 *d_sif = nullptr;
  const u32 cnt = xgsetup::cand_compact(be, t, c32, &d_id, &d_sif);
  be.release(c32);
  out->n = cnt;
  out->id = xnew(u64, cnt ? cnt : 1); out->src_if = xnew(u32, cnt ? cnt : 1);
  if (cnt) {
    CU(cudaMemcpyAsync(out->id, d_id, (size_t)cnt * sizeof(u64), cudaMemcpyDeviceToHost, be.st));
    CU(cudaMemcpyAsync(out->src_if, d_sif, (size_t)cnt * sizeof(u32), cudaMemcpyDeviceToHost, be.st));
  }
  CU(cudaStreamSynchronize(be.st));
  be.release(d_id); be.release(d_sif);
}

/* topology group of each boundary primary the exchange plan named (host arrays) */
void xcu_setup_bgrp(const struct xcu_topo *t, const u32 *bprim, u32 nb, u32 *bgrp)
{
  if (!nb) return;
  Scope sc; CudaBackend &be = sc.be;
  u32 *d_p = be.alloc<u32>(nb), *d_g = be.alloc<u32>(nb);
  CU(cudaMemcpyAsync(d_p, bprim, (size_t)nb * sizeof(u32), cudaMemcpyHostToDevice, be.st));
  xgsetup::groups_of_primaries(be, t, d_p, nb, d_g);
  CU(cudaMemcpyAsync(bgrp, d_g, (size_t)nb * sizeof(u32), cudaMemcpyDeviceToHost, be.st));
  CU(cudaStreamSynchronize(be.st));
  be.release(d_p); be.release(d_g);
}

/* returns 1 when the map is too large for the sliced layout (2^30 members or
   2^23 blocks; s is then empty and the caller keeps the map in CSR form), else 0 */
int xcu_setup_sell(const struct xcu_dmap *m, struct xcu_dsell *s)
{
  Scope sc; CudaBackend &be = sc.be;
  /* EXAGS_SELL_MAX_MEMBERS lowers the limit so that the CSR fallback can be exercised on a small map (tests) */
  const char *lim = getenv("EXAGS_SELL_MAX_MEMBERS");
  int rc = (lim && (size_t)m->ni >= (size_t)atoll(lim)) ? (int)xgsetup::ERR_SELL_TOO_LARGE : (int)xgsetup::OK;
  if (rc == xgsetup::OK) rc = xgsetup::sell_build(be, m, s, xg_sell_pairs_enabled());
  else memset(s, 0, sizeof *s);
  if (rc == xgsetup::ERR_SELL_TOO_LARGE) {
    xgsetup::sell_release(be, s);
    memset(s, 0, sizeof *s);
    CU(cudaStreamSynchronize(be.st));
    return 1;
  }
  check(rc, "sliced layout");
  CU(cudaStreamSynchronize(be.st));
  return 0;
}

/* a copy of a device CSR map that the handle owns */
void xcu_dmap_clone(const struct xcu_dmap *m, struct xcu_dmap *out)
{
  memset(out, 0, sizeof *out);
  out->ng = m->ng; out->ni = m->ni;
  if (!m->ng) return;
  out->off = (u32 *)xcu_malloc(((size_t)m->ng + 1) * sizeof(u32));
  out->idx = (u32 *)xcu_malloc(((size_t)m->ni + 1) * sizeof(u32));
  CU(cudaMemcpy(out->off, m->off, ((size_t)m->ng + 1) * sizeof(u32), cudaMemcpyDeviceToDevice));
  CU(cudaMemcpy(out->idx, m->idx, (size_t)m->ni * sizeof(u32), cudaMemcpyDeviceToDevice));
  if (m->nunf) {
    out->nunf = (u32 *)xcu_malloc(((size_t)m->ng + 1) * sizeof(u32));
    CU(cudaMemcpy(out->nunf, m->nunf, (size_t)m->ng * sizeof(u32), cudaMemcpyDeviceToDevice));
  }
}

void xcu_dmap_free(struct xcu_dmap *m)
{
  Scope sc; CudaBackend &be = sc.be;
  xgsetup::dmap_release(be, m);
}

/* brackets one gs_setup: the work arrays of all its passes share the pool    */
void xcu_setup_begin(void) { backend().begin(); }
void xcu_setup_end(void) { backend().end(); }

void xcu_dmap_to_host(const struct xcu_dmap *m, struct xg_hmap *h)
{
  memset(h, 0, sizeof *h);
  h->ng = m->ng; h->ni = m->ni;
  h->off = xnew(u32, (size_t)m->ng + 1);
  h->idx = xnew(u32, m->ni ? m->ni : 1);
  if (m->off) CU(cudaMemcpy(h->off, m->off, ((size_t)m->ng + 1) * sizeof(u32), cudaMemcpyDeviceToHost));
  else h->off[0] = 0;
  if (m->ni) CU(cudaMemcpy(h->idx, m->idx, (size_t)m->ni * sizeof(u32), cudaMemcpyDeviceToHost));
  if (m->nunf) {
    h->nunf = xnew(u32, m->ng ? m->ng : 1);
    if (m->ng) CU(cudaMemcpy(h->nunf, m->nunf, (size_t)m->ng * sizeof(u32), cudaMemcpyDeviceToHost));
  }
}

}  // extern "C"
