// gs_cuda.cu -- sm_100a kernels for the gs_op hot path and the thin C ABI the
// host C code calls them through.
//
// What the kernels replace (reference paths relative to /root/reference):
//   k_fused_sell   the same fused pass over the sliced-ELL layout (the hot
//                  kernel: one warp per 1 KB block of member indices, one CTA
//                  per window; variants: KU = 3 components per pass, vec3
//                  tuples, FLG flag masks, GONLY gather only, WGT weights)
//   k_pack_slots / k_push_wait   pw_exec_sends / pw_exec_recvs as peer stores
//                  over NVLink and a flag wait (src/gs.upc:409-468)
//   k_fused        gather_T_OP + init_T + scatter_T run back to back by
//                  gs_aux (src/gs.upc:1181-1184, src/gs_local.c:60-97) --
//                  one pass over a CSR map instead of two passes over a
//                  sentinel stream, each group reduced left to right by one
//                  thread so the association order is the reference's.
//   k_bnd_pack     local gather of boundary groups + scatter_to_buf
//                  (src/gs.upc:490, src/gs_local.c:76-87,308-318)
//   k_bnd_unpack   gather_from_buf + local scatter (src/gs.upc:499,1184)
//   k_fill/k_array gs_init_array / gs_gather_array (src/gs_local.c:187-201)
//   k_csr_gather / k_csr_scatter   the public gs_local.h entry points
//
// The path is HBM-bound integer/floating "a OP= b" work on irregular index
// sets: no tensor cores, no GEMM reshaping.  See DESIGN.md for the roofline.
#include <cuda_runtime.h>
#include <cfloat>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include "exags_internal.h"

// ---------------------------------------------------------------------------
// error funnel: every CUDA failure ends in the reference's fail() contract
// ---------------------------------------------------------------------------
#define CU(call)                                                              \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess)                                                    \
      exags_fail(1, __FILE__, __LINE__, "CUDA error: %s (%s)",                \
                 cudaGetErrorString(e_), #call);                              \
  } while (0)

static unsigned long long g_launches = 0;
#define LAUNCHED() do { ++g_launches; CU(cudaGetLastError()); } while (0)
extern "C" void xcu_note_launches(unsigned long long k) { g_launches += k; }

// ---------------------------------------------------------------------------
// monoid definitions (src/gs_defs.h:33-52)
// ---------------------------------------------------------------------------
template <typename T> struct Lim;
template <> struct Lim<double>    { static __host__ __device__ double    hi() { return DBL_MAX; }  static __host__ __device__ double    lo() { return -DBL_MAX; } };
template <> struct Lim<float>     { static __host__ __device__ float     hi() { return FLT_MAX; }  static __host__ __device__ float     lo() { return -FLT_MAX; } };
template <> struct Lim<int>       { static __host__ __device__ int       hi() { return INT_MAX; }  static __host__ __device__ int       lo() { return INT_MIN; } };
template <> struct Lim<long long> { static __host__ __device__ long long hi() { return LLONG_MAX; } static __host__ __device__ long long lo() { return LLONG_MIN; } };

template <typename T, int OP>
__host__ __device__ __forceinline__ T ident()
{
  if (OP == XO_MUL) return (T)1;
  if (OP == XO_MIN) return Lim<T>::hi();
  if (OP == XO_MAX) return Lim<T>::lo();
  return (T)0;   // add, bpr
}

// a OP= b with the reference's exact comparison forms (not fmin/fmax)
template <typename T, int OP>
__device__ __forceinline__ T apply(T a, T b)
{
  if (OP == XO_ADD) return a + b;
  if (OP == XO_MUL) return a * b;
  if (OP == XO_MIN) return (b < a) ? b : a;
  if (OP == XO_MAX) return (b > a) ? b : a;
  // binary-prefix (src/gs_defs.h:37-42): longest common leading bit string.
  // Like the reference's macro it is instantiated for every domain through a
  // conversion to unsigned (meaningful for non-negative values below 2^32).
  if (b != (T)0) {
    unsigned a_ = (unsigned)a, b_ = (unsigned)b;
    if (a_ == 0u) return (T)b_;
    for (;;) { if (a_ < b_) b_ >>= 1; else if (b_ < a_) a_ >>= 1; else break; }
    return (T)a_;
  }
  return a;
}

// ---------------------------------------------------------------------------
// field addressing: plain T[n], vec T[n][tuple], many T*[vn]
// ---------------------------------------------------------------------------
template <typename T, int MODE>
struct Fld {
  T *p[XG_MAXV];
  unsigned vn, tuple, k0;
  // component pointer picked with constant indices only: a run-time index
  // into the kernel-parameter array would make the compiler copy it to local
  // memory and fetch a pointer from there for every access
  __device__ __forceinline__ T *comp(unsigned k) const
  {
    switch (k) {
      case 0: return p[0]; case 1: return p[1]; case 2: return p[2]; case 3: return p[3];
      case 4: return p[4]; case 5: return p[5]; case 6: return p[6]; default: return p[7];
    }
  }
  __device__ __forceinline__ T *at(u32 i, unsigned k) const
  {
    if (MODE == XM_PLAIN) return p[0] + i;
    if (MODE == XM_VEC)   return p[0] + (size_t)i * tuple + k0 + k;
    return comp(k) + i;
  }
  // base of component k and element stride, for loops that visit many indices
  __device__ __forceinline__ T *base(unsigned k) const
  {
    if (MODE == XM_PLAIN) return p[0];
    if (MODE == XM_VEC)   return p[0] + k0 + k;
    return comp(k);
  }
  __device__ __forceinline__ size_t stride() const { return MODE == XM_VEC ? (size_t)tuple : (size_t)1; }
};

template <typename T, int MODE>
static Fld<T, MODE> make_fld(const xcu_field *f)
{
  Fld<T, MODE> r;
  for (int k = 0; k < XG_MAXV; ++k) r.p[k] = (T *)f->p[k];
  r.vn = f->vn; r.tuple = f->tuple; r.k0 = f->k0;
  return r;
}

// runtime-mode variant for the non-hot public gs_local.h kernels
template <typename T>
struct FldR {
  T *p[XG_MAXV];
  unsigned vn, tuple, k0; int mode;
  __device__ __forceinline__ T *at(u32 i, unsigned k) const
  {
    if (mode == XM_PLAIN) return p[0] + i;
    if (mode == XM_VEC)   return p[0] + (size_t)i * tuple + k0 + k;
    return p[k] + i;
  }
};
template <typename T>
static FldR<T> make_fldr(const xcu_field *f, int mode)
{
  FldR<T> r;
  for (int k = 0; k < XG_MAXV; ++k) r.p[k] = (T *)f->p[k];
  r.vn = f->vn; r.tuple = f->tuple; r.k0 = f->k0; r.mode = mode;
  return r;
}

// ---------------------------------------------------------------------------
// one group: reduce members [0,nred) left to right, write members [0,nwr)
// ---------------------------------------------------------------------------
#define GRP_REG 8   // members kept in registers; SEM hex meshes have 2/4/8

template <typename T, int OP, int MODE>
__device__ __forceinline__ void group_fused(const Fld<T, MODE> &f, const u32 *__restrict__ idx,
                                            u32 nall, u32 nred, u32 nwr)
{
  if (nall <= GRP_REG) {
    u32 id[GRP_REG];
#pragma unroll
    for (int j = 0; j < GRP_REG; ++j) id[j] = (j < (int)nall) ? __ldg(idx + j) : 0u;
    for (unsigned k = 0; k < f.vn; ++k) {
      T x[GRP_REG];
#pragma unroll
      for (int j = 0; j < GRP_REG; ++j) if (j < (int)nred) x[j] = *f.at(id[j], k);
      T v = nred ? x[0] : ident<T, OP>();
#pragma unroll
      for (int j = 1; j < GRP_REG; ++j) if (j < (int)nred) v = apply<T, OP>(v, x[j]);
#pragma unroll
      for (int j = 0; j < GRP_REG; ++j) if (j < (int)nwr) *f.at(id[j], k) = v;
    }
  } else {
    for (unsigned k = 0; k < f.vn; ++k) {
      T v = nred ? *f.at(__ldg(idx), k) : ident<T, OP>();
      for (u32 j = 1; j < nred; ++j) v = apply<T, OP>(v, *f.at(__ldg(idx + j), k));
      for (u32 j = 0; j < nwr; ++j) *f.at(__ldg(idx + j), k) = v;
    }
  }
}

// reference semantics of one gs() on a group with nall members of which the
// first nunf are unflagged (src/gs.upc:1181-1184):
//   transpose==0: reduce the unflagged members (identity if none: the primary
//                 is a flagged primary and gs_init overwrites it), write all
//   transpose==1: reduce all members, write the primary and the unflagged
__device__ __forceinline__ void group_counts(u32 nall, u32 nunf, int transpose, u32 &nred, u32 &nwr)
{
  if (transpose) { nred = nall; nwr = nunf > 1u ? nunf : 1u; }
  else           { nred = nunf; nwr = nall; }
}

template <typename T, int OP, int MODE>
__global__ void __launch_bounds__(256)
k_fused(xcu_csr m, Fld<T, MODE> f, int transpose)
{
  u32 g = blockIdx.x * 256u + threadIdx.x;
  if (g >= m.ng) return;
  u32 b = __ldg(m.off + g), e = __ldg(m.off + g + 1);
  u32 nall = e - b, nunf = m.nunf ? __ldg(m.nunf + g) : nall, nred, nwr;
  group_counts(nall, nunf, transpose, nred, nwr);
  group_fused<T, OP, MODE>(f, m.idx + b, nall, nred, nwr);
}


// ---------------------------------------------------------------------------
// k_fused_sell: the hot kernel.  One warp per 1 KB block of the sliced-ELL
// map (struct xg_sell), one CTA per window of 8 blocks.  Lane l owns 8
// consecutive member slots, fetched with two coalesced 128-bit loads, plus one
// start-mask byte; there is no offset array and no dependent offset->index
// load, and every lane keeps 8 independent field loads in flight whether the
// mesh's groups have 2, 4 or 8 members (faces, edges, corners of hex
// elements).  The 8 slots are then reduced as a segmented left-to-right scan
// (forward pass) and the segment totals are copied back over the segment
// (backward pass): each group is summed by one thread in the reference's
// association order (src/gs_local.c:60-71) and its result is written to every
// member (src/gs_local.c:76-87) -- gather and scatter fused in one pass over
// the field.  Both passes are select-based, so lanes with different group
// layouts do not diverge.
// ---------------------------------------------------------------------------
// KU = components handled per pass: 1 for plain fields; 3 for vec/many so the
// loads of all three components of a velocity field are in flight together
// (BASELINE config 3) and the index block is read once for all of them.
// one block of one lane: 8 predicated loads per component, segmented scan, 8
// predicated stores
// GONLY: gather only -- the group total is written to the group's first slot
// (the primary) and the other members are left for the unpack kernel, which
// has to write them anyway once the neighbours' contributions are in.
// FLG: the map has flagged members (bit j of flg: slot j is one).  As
// gs_aux does with map_local[0] / map_local[1] (src/gs.upc:1181-1184):
//   transpose == 0: flagged members do not contribute (a group with none
//                   left reduces to the identity, which is what gs_init writes
//                   over a flagged primary) and every member is written;
//   transpose == 1: every member contributes, flagged members other than
//                   the primary are not written.
// ONE: all components fit one pass (vn <= KU), so component numbers are
// compile-time constants and gs_many's pointer table is read with constant
// indices.
// Phase (csrc/sell_window.h, xg_sell_pair_order): the odd bits of a lane's
// start mask are never group starts; bit 2j+1 set means "fetch the pair in
// slots (2j, 2j+1) partner first".  swap_pairs exchanges bits / values 2j and
// 2j+1 where that bit is set, so that load and store instruction j of the
// warp touches the partner in such lanes and the primary in the others --
// the two then share a 128-byte line by construction of the layout.  The
// values go back to slot order before the reduction: results do not depend on
// the phase.
__device__ __forceinline__ unsigned swap_pair_bits(unsigned m, unsigned ph)
{
  const unsigned hi = ph & 0xAAu, lo = hi >> 1, both = hi | lo;
  return (m & ~both) | ((m & lo) << 1) | ((m & hi) >> 1);
}

// WGT: weighted gs (exags_gs_weighted, include/exags_cuda.h) -- the value stored
// to a member is the group's result times wgt[member]; the weights are fetched
// together with the field so that both loads are in flight at once.
template <typename T, int OP, int MODE, int KU, bool GONLY = false, bool FLG = false, bool ONE = false,
          int WGT = 0>
__device__ __forceinline__ void sell_lane(const Fld<T, MODE> &f, const uint4 a, const uint4 b,
                                          const unsigned startph, const unsigned flg = 0u,
                                          const int transpose = 0, const T *__restrict__ wgt = nullptr)
{
  static_assert(!WGT || !GONLY, "weights: full gather-scatter only");
  const u32 id0[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
  const unsigned start = startph & 0x55u, ph = startph & 0xAAu;
  unsigned mem = 0u;                          // slots that hold a member
#pragma unroll
  for (int j = 0; j < 8; ++j) if (id0[j] != XG_NONE) mem |= 1u << j;
  unsigned in = mem, out = mem;               // slots that contribute / are written
  if (FLG) {
    if (transpose) {
      // a one-member group (a flagged single: a start bit with the next slot
      // empty) only exists for gs_init, which runs with transpose == 0: skip it
      mem &= ~(start & ~(mem >> 1));
      in = mem; out = mem & (~flg | start);
    } else in = mem & ~flg;
  }
  if (GONLY) out &= start;
  // everything below addresses memory in "fetch order": id[], lld, lout
  u32 id[8];
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    const bool sw = (ph >> (j + 1)) & 1u;
    id[j] = sw ? id0[j + 1] : id0[j];
    id[j + 1] = sw ? id0[j] : id0[j + 1];
  }
  // a phase bit is only ever set on a complete pair: without flags both of its
  // slots contribute and both are written, so the masks need no swapping
  const unsigned lout = (FLG || GONLY) ? swap_pair_bits(out, ph) : out;
  // Every member is loaded, also one whose value does not contribute (a flagged
  // member with transpose == 0): it is written afterwards, and an 8-byte store
  // into a 32-byte sector that is not in L2 makes the memory system read the
  // sector anyway -- as a load ahead of the store that read is overlapped like
  // every other one (unique=1 handle, t=0, M7: 0.375 -> see profiles/r2_configs).
  const unsigned lld = FLG ? swap_pair_bits(mem, ph) : mem;
  const size_t es = f.stride();
  for (unsigned k0 = 0; k0 < (ONE ? 1u : f.vn); k0 += KU) {
    T x[KU][8];
    T *pk[KU];
#pragma unroll
    for (int kk = 0; kk < KU; ++kk) pk[kk] = f.base(k0 + kk);
#pragma unroll
    for (int kk = 0; kk < KU; ++kk)
      if (KU == 1 || k0 + kk < f.vn) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if ((lld >> j) & 1u) x[kk][j] = pk[kk][id[j] * es];
      }
    // WGT == 1: the weights are fetched together with the field (both loads in
    // flight at once); WGT == 2: after the reduction, just ahead of the stores --
    // they then occupy no registers across it, which is what the three-component
    // float kernels need to keep four CTAs on an SM without spilling.
    // A one-member group (a flagged single: start bit, next slot empty) is no
    // group result -- gs_init's identity goes there unweighted.  Such a lane
    // slot has no phase, so slot order and fetch order agree for it.
    const unsigned lw = WGT ? (FLG ? lout & ~(start & ~(mem >> 1)) : lout) : 0u;
    T wv[WGT ? 8 : 1];
    if (WGT == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        wv[WGT ? j : 0] = (T)1;
        if ((lw >> j) & 1u) wv[WGT ? j : 0] = __ldg(wgt + id[j]);
      }
    }
#pragma unroll
    for (int kk = 0; kk < KU; ++kk) {
      // fetch order -> slot order
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const bool sw = (ph >> (j + 1)) & 1u;
        const T p0 = x[kk][j], p1 = x[kk][j + 1];
        x[kk][j] = sw ? p1 : p0; x[kk][j + 1] = sw ? p0 : p1;
      }
      // a group whose first member does not contribute (all its members are
      // flagged, transpose == 0) starts from the identity; the other slots that
      // do not contribute are skipped by the scan below
      if (FLG) {
#pragma unroll
        for (int j = 0; j < 8; j += 2)
          if (((start & ~in) >> j) & 1u) x[kk][j] = ident<T, OP>();
      }
      // forward: running reduction, restarted where a group begins
#pragma unroll
      for (int j = 1; j < 8; ++j) {
        const T cont = ((in >> j) & 1u) ? apply<T, OP>(x[kk][j - 1], x[kk][j]) : x[kk][j - 1];
        x[kk][j] = ((start >> j) & 1u) ? x[kk][j] : cont;
      }
      // backward: the last slot of a group holds its total; spread it
#pragma unroll
      for (int j = 6; j >= 0; --j) x[kk][j] = ((start >> (j + 1)) & 1u) ? x[kk][j] : x[kk][j + 1];
      // slot order -> fetch order
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const bool sw = (ph >> (j + 1)) & 1u;
        const T p0 = x[kk][j], p1 = x[kk][j + 1];
        x[kk][j] = sw ? p1 : p0; x[kk][j + 1] = sw ? p0 : p1;
      }
    }
    if (WGT == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        wv[WGT ? j : 0] = (T)1;
        if ((lw >> j) & 1u) wv[WGT ? j : 0] = __ldg(wgt + id[j]);
      }
    }
#pragma unroll
    for (int kk = 0; kk < KU; ++kk)
      if (KU == 1 || k0 + kk < f.vn) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if ((lout >> j) & 1u) pk[kk][id[j] * es] = WGT ? x[kk][j] * wv[WGT ? j : 0] : x[kk][j];
      }
  }
}

template <typename T, int OP, int MODE, int KU, int MINB, bool GONLY = false, bool FLG = false, bool ONE = false>
__global__ void __launch_bounds__(32 * XG_SELL_CW, XG_MINB(MINB))
k_fused_sell(xcu_sell m, Fld<T, MODE> f, int transpose)
{
  const u32 c = blockIdx.x * (u32)XG_SELL_CW + (threadIdx.x >> 5);
  const unsigned lane = threadIdx.x & 31u;
  const uint4 *p = reinterpret_cast<const uint4 *>(m.idx + ((size_t)c * 32u + lane) * 8u);
  const uint4 a = __ldg(p), b = __ldg(p + 1);
  const unsigned start = __ldg(m.mask + (size_t)c * 32u + lane);
  const unsigned flg = FLG ? __ldg(m.fmask + (size_t)c * 32u + lane) : 0u;
  sell_lane<T, OP, MODE, KU, GONLY, FLG, ONE>(f, a, b, start, flg, transpose);
}

// Weighted form of the two local kernels (exags_gs_weighted[_vec|_many]: Nek's
// dssum followed by a multiplication with the inverse multiplicity, fused):
// float/double fields.  One weight per member is read (it applies to every
// component of a vector field), against a separate pass over the whole field
// (read u, read w, write u).
template <typename T, int OP, int MODE, int KU, int MINB, bool FLG, bool ONE, int WL = 1>
__global__ void __launch_bounds__(32 * XG_SELL_CW, XG_MINB(MINB))
k_fused_sell_w(xcu_sell m, Fld<T, MODE> f, const T *__restrict__ wgt, int transpose)
{
  const u32 c = blockIdx.x * (u32)XG_SELL_CW + (threadIdx.x >> 5);
  const unsigned lane = threadIdx.x & 31u;
  const uint4 *p = reinterpret_cast<const uint4 *>(m.idx + ((size_t)c * 32u + lane) * 8u);
  const uint4 a = __ldg(p), b = __ldg(p + 1);
  const unsigned start = __ldg(m.mask + (size_t)c * 32u + lane);
  const unsigned flg = FLG ? __ldg(m.fmask + (size_t)c * 32u + lane) : 0u;
  sell_lane<T, OP, MODE, KU, false, FLG, ONE, WL>(f, a, b, start, flg, transpose, wgt);
}

template <typename T, int OP, int MODE>
__global__ void __launch_bounds__(256)
k_fused_w(xcu_csr m, Fld<T, MODE> f, const T *__restrict__ wgt, int transpose)
{
  u32 g = blockIdx.x * 256u + threadIdx.x;
  if (g >= m.ng) return;
  u32 b = __ldg(m.off + g), e = __ldg(m.off + g + 1);
  u32 nall = e - b, nunf = m.nunf ? __ldg(m.nunf + g) : nall, nred, nwr;
  group_counts(nall, nunf, transpose, nred, nwr);
  const u32 *idx = m.idx + b;
  for (unsigned k = 0; k < f.vn; ++k) {
    T v = nred ? *f.at(__ldg(idx), k) : ident<T, OP>();
    for (u32 j = 1; j < nred; ++j) v = apply<T, OP>(v, *f.at(__ldg(idx + j), k));
    if (nall == 1u) { *f.at(__ldg(idx), k) = v; continue; }    // a flagged single: the identity, unweighted
    for (u32 j = 0; j < nwr; ++j) { const u32 i = __ldg(idx + j); *f.at(i, k) = v * __ldg(wgt + i); }
  }
}

// gs_vec with 3-tuples (a velocity field, BASELINE config 3): a tuple is 3
// consecutive elements, so its address is aligned to one element only.  Read
// it as one aligned pair plus one single -- pair first when the index is even,
// single first when it is odd -- instead of three element loads: two trips
// through L1 per member instead of three.
template <typename T> struct Pair;
template <> struct Pair<double>    { typedef double2 type; };
template <> struct Pair<float>     { typedef float2 type; };
template <> struct Pair<int>       { typedef int2 type; };
template <> struct Pair<long long> { typedef longlong2 type; };

template <typename T, int OP>
__device__ __forceinline__ void vec3_lane(T *__restrict__ u, const uint4 a, const uint4 b, const unsigned startph)
{
  typedef typename Pair<T>::type T2;
  const unsigned start = startph & 0x55u, ph = startph & 0xAAu;   // see swap_pair_bits
  const u32 id0[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
  u32 id[8];                                   // fetch order
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    const bool sw = (ph >> (j + 1)) & 1u;
    id[j] = sw ? id0[j + 1] : id0[j];
    id[j + 1] = sw ? id0[j] : id0[j + 1];
  }
  T x[3][8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (id[j] != XG_NONE) {
      const size_t base = (size_t)id[j] * 3u;
      const bool odd = id[j] & 1u;
      const T2 pr = *reinterpret_cast<const T2 *>(u + base + (odd ? 1 : 0));
      const T sg = u[base + (odd ? 0 : 2)];
      x[0][j] = odd ? sg : pr.x; x[1][j] = odd ? pr.x : pr.y; x[2][j] = odd ? pr.y : sg;
    }
#pragma unroll
  for (int kk = 0; kk < 3; ++kk) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {           // fetch order -> slot order
      const bool sw = (ph >> (j + 1)) & 1u;
      const T p0 = x[kk][j], p1 = x[kk][j + 1];
      x[kk][j] = sw ? p1 : p0; x[kk][j + 1] = sw ? p0 : p1;
    }
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      const T cont = (id0[j] != XG_NONE) ? apply<T, OP>(x[kk][j - 1], x[kk][j]) : x[kk][j - 1];
      x[kk][j] = ((start >> j) & 1u) ? x[kk][j] : cont;
    }
#pragma unroll
    for (int j = 6; j >= 0; --j) x[kk][j] = ((start >> (j + 1)) & 1u) ? x[kk][j] : x[kk][j + 1];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {           // slot order -> fetch order
      const bool sw = (ph >> (j + 1)) & 1u;
      const T p0 = x[kk][j], p1 = x[kk][j + 1];
      x[kk][j] = sw ? p1 : p0; x[kk][j + 1] = sw ? p0 : p1;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (id[j] != XG_NONE) {
      const size_t base = (size_t)id[j] * 3u;
      const bool odd = id[j] & 1u;
      T2 pr;
      pr.x = odd ? x[1][j] : x[0][j]; pr.y = odd ? x[2][j] : x[1][j];
      *reinterpret_cast<T2 *>(u + base + (odd ? 1 : 0)) = pr;
      u[base + (odd ? 0 : 2)] = odd ? x[0][j] : x[2][j];
    }
}

template <typename T, int OP, int MINB>
__global__ void __launch_bounds__(32 * XG_SELL_CW, XG_MINB(MINB))
k_fused_sell_vec3(xcu_sell m, T *__restrict__ u)
{
  const u32 c = blockIdx.x * (u32)XG_SELL_CW + (threadIdx.x >> 5);
  const unsigned lane = threadIdx.x & 31u;
  const uint4 *p = reinterpret_cast<const uint4 *>(m.idx + ((size_t)c * 32u + lane) * 8u);
  const uint4 a = __ldg(p), b = __ldg(p + 1);
  const unsigned startph = __ldg(m.mask + (size_t)c * 32u + lane);
  vec3_lane<T, OP>(u, a, b, startph);
}

// A persistent form of these kernels -- CTAs that stay resident and walk the
// windows round robin while one thread streams the index blocks of the windows
// ahead into a shared-memory ring with TMA bulk copies (cp.async.bulk +
// mbarrier), so that no warp waits a memory latency for its indices -- was
// built and measured in round 2 and is not kept: 0.402 ms against 0.321 ms with
// a 2-stage ring, 0.548 ms with 4 stages.  The ring takes its shared memory
// out of the L1 (17 KB x 6 CTAs per stage pair), the L1 hit rate of the field
// accesses falls from 42 % to 30 %, and the field's lines -- which this layout
// arranges to be shared between the loads of a warp and of its neighbours --
// are what the L1 is needed for (profiles/r2_knobs_pipe_line_shift.txt,
// profiles/r2_ncu_persistent_tma_kernels_summary.txt).

template <typename T, int OP, int MODE>
__global__ void __launch_bounds__(256)
k_init_list(const u32 *__restrict__ list, u32 n, Fld<T, MODE> f)
{
  u32 j = blockIdx.x * 256u + threadIdx.x;
  if (j >= n) return;
  u32 i = __ldg(list + j);
  for (unsigned k = 0; k < f.vn; ++k) *f.at(i, k) = ident<T, OP>();
}

// boundary primaries: gather -> u[primary] and -> send slots.
// PUSH: the send slots live in the neighbours' mailboxes; each value is stored
// straight into the receiving GPU's memory over NVLink (the upc_memput of
// pw_exec_sends, src/gs.upc:437-468, done by the thread that produced the
// value), and the last CTA to finish raises this rank's flag in every
// neighbour's mailbox (the "flag" put of the same function).
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{ asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{ unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// The number of the exchange this launch belongs to (struct xcu_push): read
// by every thread before any CTA of the launch can have advanced it -- the
// last CTA to finish does that, and it finishes after every CTA has started.
__device__ __forceinline__ u32 push_epoch(const xcu_push &pp)
{
  return pp.epoch_imm ? pp.epoch_imm : *reinterpret_cast<volatile u32 *>(pp.depoch) + 1u;
}

__device__ __forceinline__ void push_signal_all(const xcu_push &pp, u32 epoch)
{
  *pp.depoch = epoch;
  for (u32 i = 0; i < pp.nsig; ++i)
    st_release_sys(reinterpret_cast<unsigned *>(pp.sig[i]), epoch);
}

template <typename T, int OP, int MODE, bool PUSH>
__global__ void __launch_bounds__(256)
k_bnd_pack(xcu_bnd bd, Fld<T, MODE> f, int transpose, T *__restrict__ sendbuf,
           unsigned bufvn, int allreduce, xcu_push pp)
{
  const u32 epoch = PUSH ? push_epoch(pp) : 0u;
  const unsigned long long *pbase = PUSH ? pp.base + (size_t)(epoch & 1u) * pp.par_stride : nullptr;
  for (u32 b = blockIdx.x * 256u + threadIdx.x; b < bd.nb; b += gridDim.x * 256u) {
    u32 gb = __ldg(bd.goff + b), ge = __ldg(bd.goff + b + 1);
    u32 nall = ge - gb, nunf = __ldg(bd.nunf + b), nred, nwr;
    group_counts(nall, nunf, transpose, nred, nwr);
    const u32 *idx = bd.gidx + gb;
    u32 prim = __ldg(idx);
    const int sd = 1 ^ transpose;                 // send map (src/gs.upc:479)
    u32 sb = 0, se = 0, ord = 0; bool contributes = true;
    if (allreduce) {
      ord = __ldg(bd.ord + b);
      // map_to_buf (src/gs.upc:1059,1063): flagged primaries stay out unless transposed
      contributes = transpose || !bd.pflag[b];
    } else {
      sb = __ldg(bd.soff[sd] + b); se = __ldg(bd.soff[sd] + b + 1);
    }
    for (unsigned k = 0; k < f.vn; ++k) {
      T v = nred ? *f.at(prim, k) : ident<T, OP>();
      for (u32 j = 1; j < nred; ++j) v = apply<T, OP>(v, *f.at(__ldg(idx + j), k));
      *f.at(prim, k) = v;
      if (PUSH) {
        for (u32 s = sb; s < se; ++s) {
          const unsigned nbk = __ldg(pp.snb + s);
          T *dst = reinterpret_cast<T *>(__ldg(pbase + nbk));
          const u32 slot = __ldg(bd.slot[sd] + s) + __ldg(pp.delta + nbk);
          dst[(size_t)slot * bufvn + f.k0 + k] = v;
        }
      } else if (allreduce) {
        if (contributes) sendbuf[(size_t)ord * bufvn + f.k0 + k] = v;
      } else {
        for (u32 s = sb; s < se; ++s)
          sendbuf[(size_t)__ldg(bd.slot[sd] + s) * bufvn + f.k0 + k] = v;
      }
    }
  }
  if (PUSH && pp.nsig) {
    __threadfence_system();                       // my peer stores are out
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned done = atomicAdd(pp.counter, 1u);
      if (done == gridDim.x - 1u) {               // every CTA's stores are out
        *pp.counter = 0u;
        __threadfence_system();
        push_signal_all(pp, epoch);
      }
    }
  }
}

// Slot-ordered pack for symmetric maps: the boundary groups have been reduced
// (and scattered) by the sliced-ELL kernel, so u[primary] is the group total
// and slot s only has to carry u[sprim[s]] -- scatter_to_buf of
// src/gs.upc:490 read the other way round.  Consecutive threads fill
// consecutive slots, so the stores into a neighbour's mailbox are contiguous.
template <typename T, int MODE, bool PUSH>
__global__ void __launch_bounds__(256)
k_pack_slots(const u32 *__restrict__ sprim, u32 nslots, Fld<T, MODE> f, T *__restrict__ sendbuf,
             unsigned bufvn, xcu_push pp)
{
  const u32 epoch = PUSH ? push_epoch(pp) : 0u;
  const unsigned long long *pbase = PUSH ? pp.base + (size_t)(epoch & 1u) * pp.par_stride : nullptr;
  for (u32 s = blockIdx.x * 256u + threadIdx.x; s < nslots; s += gridDim.x * 256u) {
    const u32 prim = __ldg(sprim + s);
    T *dst = sendbuf;
    u32 slot = s;
    if (PUSH) {
      u32 lo = 0, hi = pp.nnb;                    // neighbour whose slot range holds s
      while (hi - lo > 1u) { const u32 mid = (lo + hi) >> 1; if (__ldg(pp.start + mid) <= s) lo = mid; else hi = mid; }
      dst = reinterpret_cast<T *>(__ldg(pbase + lo));
      slot = s + __ldg(pp.delta + lo);
    }
    for (unsigned k = 0; k < f.vn; ++k)
      dst[(size_t)slot * bufvn + f.k0 + k] = *f.at(prim, k);
  }
  if (PUSH && pp.nsig) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned done = atomicAdd(pp.counter, 1u);
      if (done == gridDim.x - 1u) {
        *pp.counter = 0u;
        __threadfence_system();
        push_signal_all(pp, epoch);
      }
    }
  }
}

// signal without data (a rank whose boundary groups send nothing this call)
__global__ void k_push_signal(xcu_push pp)
{
  if (threadIdx.x == 0) { const u32 epoch = push_epoch(pp); __threadfence_system(); push_signal_all(pp, epoch); }
}

// The receiving side of pw_exec_recvs' poll loop (src/gs.upc:409-435): one
// thread per neighbour waits until that neighbour's flag in my own mailbox has
// reached this call's epoch.  The flags are local memory written by the peer
// over NVLink; the unpack kernel follows in stream order.  A bounded wait: a
// peer that never arrives is reported through *status instead of hanging the
// GPU.
__global__ void k_push_wait(const unsigned *__restrict__ flags, const u32 *__restrict__ ranks,
                            u32 n, const u32 *depoch, u32 epoch_imm, unsigned long long timeout_ns, int *status)
{
  const u32 t = threadIdx.x;
  if (t >= n) return;
  // my own pack (or signal) kernel precedes this one in stream order and has
  // stored this exchange's number
  const u32 epoch = epoch_imm ? epoch_imm : *reinterpret_cast<const volatile u32 *>(depoch);
  const unsigned *f = flags + ranks[t];
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while ((int)(ld_acquire_sys(f) - epoch) < 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > timeout_ns) { *status = 1 + (int)ranks[t]; break; }
    __nanosleep(100);
  }
}

// boundary primaries: fold received slots in ascending remote rank, scatter
template <typename T, int OP, int MODE>
__global__ void __launch_bounds__(256)
k_bnd_unpack(xcu_bnd bd, Fld<T, MODE> f, int transpose, const T *__restrict__ recvbuf,
             unsigned bufvn, int allreduce, const u32 *depoch, size_t par_bytes)
{
  // push transport: the data area of this exchange's parity
  if (depoch) recvbuf = reinterpret_cast<const T *>(reinterpret_cast<const char *>(recvbuf) +
                                                    (size_t)(*reinterpret_cast<const volatile u32 *>(depoch) & 1u) * par_bytes);
  for (u32 b = blockIdx.x * 256u + threadIdx.x; b < bd.nb; b += gridDim.x * 256u) {
    u32 gb = __ldg(bd.goff + b), ge = __ldg(bd.goff + b + 1);
    u32 nall = ge - gb, nunf = __ldg(bd.nunf + b), nred, nwr;
    group_counts(nall, nunf, transpose, nred, nwr);
    const u32 *idx = bd.gidx + gb;
    u32 prim = __ldg(idx);
    const int rd = 0 ^ transpose;                 // recv map (src/gs.upc:479)
    u32 rb = 0, re = 0, ord = 0; bool takes = true;
    if (allreduce) {
      ord = __ldg(bd.ord + b);
      // map_from_buf (src/gs.upc:1060,1064)
      takes = !transpose || !bd.pflag[b];
    } else {
      rb = __ldg(bd.soff[rd] + b); re = __ldg(bd.soff[rd] + b + 1);
    }
    for (unsigned k = 0; k < f.vn; ++k) {
      T v = *f.at(prim, k);
      if (allreduce) {
        if (takes) v = recvbuf[(size_t)ord * bufvn + f.k0 + k];
      } else {
        for (u32 s = rb; s < re; ++s)
          v = apply<T, OP>(v, recvbuf[(size_t)__ldg(bd.slot[rd] + s) * bufvn + f.k0 + k]);
      }
      for (u32 j = 0; j < nwr; ++j) *f.at(__ldg(idx + j), k) = v;
    }
  }
}

// weighted form (float/double fields): what the unpack writes to a member
// is the group's result times the member's weight
template <typename T, int OP, int MODE>
__global__ void __launch_bounds__(256)
k_bnd_unpack_w(xcu_bnd bd, Fld<T, MODE> f, const T *__restrict__ wgt, int transpose,
               const T *__restrict__ recvbuf, unsigned bufvn, int allreduce, const u32 *depoch, size_t par_bytes)
{
  if (depoch) recvbuf = reinterpret_cast<const T *>(reinterpret_cast<const char *>(recvbuf) +
                                                    (size_t)(*reinterpret_cast<const volatile u32 *>(depoch) & 1u) * par_bytes);
  for (u32 b = blockIdx.x * 256u + threadIdx.x; b < bd.nb; b += gridDim.x * 256u) {
    u32 gb = __ldg(bd.goff + b), ge = __ldg(bd.goff + b + 1);
    u32 nall = ge - gb, nunf = __ldg(bd.nunf + b), nred, nwr;
    group_counts(nall, nunf, transpose, nred, nwr);
    const u32 *idx = bd.gidx + gb;
    const int rd = 0 ^ transpose;
    for (unsigned k = 0; k < f.vn; ++k) {
      T v = *f.at(__ldg(idx), k);
      if (allreduce) {
        if (!transpose || !bd.pflag[b]) v = recvbuf[(size_t)__ldg(bd.ord + b) * bufvn + f.k0 + k];
      } else {
        const u32 rb = __ldg(bd.soff[rd] + b), re = __ldg(bd.soff[rd] + b + 1);
        for (u32 s = rb; s < re; ++s) v = apply<T, OP>(v, recvbuf[(size_t)__ldg(bd.slot[rd] + s) * bufvn + f.k0 + k]);
      }
      for (u32 j = 0; j < nwr; ++j) { const u32 i = __ldg(idx + j); *f.at(i, k) = v * __ldg(wgt + i); }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_fill(T *__restrict__ out, size_t n, T v)
{
  size_t i = (size_t)blockIdx.x * 256u + threadIdx.x;
  size_t step = (size_t)gridDim.x * 256u;
  for (; i < n; i += step) out[i] = v;
}

template <typename T, int OP>
__global__ void __launch_bounds__(256) k_array(T *__restrict__ out, const T *__restrict__ in, size_t n)
{
  size_t i = (size_t)blockIdx.x * 256u + threadIdx.x;
  size_t step = (size_t)gridDim.x * 256u;
  for (; i < n; i += step) out[i] = apply<T, OP>(out[i], in[i]);
}

// public gs_local.h kernels on a CSR built from a sentinel map:
// member 0 of each group indexes `out`, the rest index `in`.
template <typename T, int OP>
__global__ void __launch_bounds__(256)
k_csr_gather(xcu_csr m, FldR<T> out, FldR<T> in)
{
  u32 g = blockIdx.x * 256u + threadIdx.x;
  if (g >= m.ng) return;
  u32 b = __ldg(m.off + g), e = __ldg(m.off + g + 1);
  u32 dst = __ldg(m.idx + b);
  for (unsigned k = 0; k < out.vn; ++k) {
    T v = *out.at(dst, k);
    for (u32 j = b + 1; j < e; ++j) v = apply<T, OP>(v, *in.at(__ldg(m.idx + j), k));
    *out.at(dst, k) = v;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_csr_scatter(xcu_csr m, FldR<T> out, FldR<T> in)
{
  u32 g = blockIdx.x * 256u + threadIdx.x;
  if (g >= m.ng) return;
  u32 b = __ldg(m.off + g), e = __ldg(m.off + g + 1);
  u32 src = __ldg(m.idx + b);
  for (unsigned k = 0; k < out.vn; ++k) {
    T v = *in.at(src, k);
    for (u32 j = b + 1; j < e; ++j) *out.at(__ldg(m.idx + j), k) = v;
  }
}

// comm_dot on device vectors (src/comm.upc:954-959, src/tensor.c:12-17): the
// other per-iteration reduction a Nek solver issues next to gs.  Two passes,
// both in a fixed order (a CTA's threads stride the vector, a tree inside the
// CTA, then one CTA folds the partials left to right), so the result does not
// depend on scheduling; it is a different association than the reference's
// serial loop, hence equal only up to rounding.  HBM-bound: 16 B per element.
#define DOT_THREADS 256
#define DOT_UNROLL 4          // independent 16-byte load pairs in flight per thread
__global__ void __launch_bounds__(DOT_THREADS)
k_dot_partial(const double *__restrict__ v, const double *__restrict__ w, size_t n, double *__restrict__ part)
{
  __shared__ double sh[DOT_THREADS];
  const size_t n2 = n / 2, stride = (size_t)gridDim.x * DOT_THREADS;
  const double2 *v2 = reinterpret_cast<const double2 *>(v), *w2 = reinterpret_cast<const double2 *>(w);
  double acc[DOT_UNROLL];
#pragma unroll
  for (int q = 0; q < DOT_UNROLL; ++q) acc[q] = 0;
  size_t i = (size_t)blockIdx.x * DOT_THREADS + threadIdx.x;
  for (; i + (DOT_UNROLL - 1) * stride < n2; i += DOT_UNROLL * stride) {
    double2 a[DOT_UNROLL], b[DOT_UNROLL];
#pragma unroll
    for (int q = 0; q < DOT_UNROLL; ++q) { a[q] = __ldcs(v2 + i + q * stride); b[q] = __ldcs(w2 + i + q * stride); }
#pragma unroll
    for (int q = 0; q < DOT_UNROLL; ++q) acc[q] = fma(a[q].y, b[q].y, fma(a[q].x, b[q].x, acc[q]));
  }
  for (; i < n2; i += stride) {
    const double2 a = __ldcs(v2 + i), b = __ldcs(w2 + i);
    acc[0] = fma(a.y, b.y, fma(a.x, b.x, acc[0]));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 1)) acc[0] = fma(v[n - 1], w[n - 1], acc[0]);
  double t = acc[0];
#pragma unroll
  for (int q = 1; q < DOT_UNROLL; ++q) t += acc[q];
  sh[threadIdx.x] = t;
  __syncthreads();
  for (unsigned s = DOT_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

// one CTA folds the partials: strided sums per thread, then the same tree as
// above -- a fixed order again.  (A single thread adding the 1184 partials one
// after the other took 51 us, profiles/r2_ncu_all_kernels_h_table.txt.)
__global__ void __launch_bounds__(DOT_THREADS)
k_dot_final(const double *__restrict__ part, unsigned nparts, double *__restrict__ out)
{
  __shared__ double sh[DOT_THREADS];
  double acc = 0;
  for (unsigned i = threadIdx.x; i < nparts; i += DOT_THREADS) acc += part[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (unsigned s = DOT_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

// ---------------------------------------------------------------------------
// dispatch helpers
// ---------------------------------------------------------------------------
static inline unsigned blocks_for(size_t n) { return (unsigned)((n + 255u) / 256u); }

/* Boundary kernels run on the high-priority stream beside the interior
   kernel and have its whole duration to finish.  A full-width grid of them
   would take every SM away from the bandwidth-bound interior kernel for the
   length of their dependent-load chains, so they get one CTA per SM (grid-
   stride) until the boundary is big enough to need more: at most ~16 groups
   per thread. */
static inline unsigned bnd_blocks(size_t nb)
{
  unsigned want = blocks_for(nb), narrow = (unsigned)((nb + 4095u) / 4096u);
  static int sm = 0, wide = -1;
  if (wide < 0) { const char *v = getenv("EXAGS_BND_WIDE"); wide = v ? atoi(v) : 0; }
  if (wide) return want;
  if (!sm) { int d = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, d); if (sm <= 0) sm = 148; }
  if (narrow < (unsigned)sm) narrow = (unsigned)sm;
  return narrow < want ? narrow : want;
}

#define FOR_OP(T, MODE, X)                                                    \
  switch (xop) {                                                              \
    case XO_ADD: X(T, XO_ADD, MODE); break;                                   \
    case XO_MUL: X(T, XO_MUL, MODE); break;                                   \
    case XO_MIN: X(T, XO_MIN, MODE); break;                                   \
    case XO_MAX: X(T, XO_MAX, MODE); break;                                   \
    case XO_BPR: X(T, XO_BPR, MODE); break;                                   \
    default: exags_fail(1, __FILE__, __LINE__, "gs: bad op %d", xop);         \
  }
#define FOR_DOM(MODE, X)                                                      \
  switch (xdom) {                                                             \
    case XD_F64: FOR_OP(double, MODE, X); break;                              \
    case XD_F32: FOR_OP(float, MODE, X); break;                               \
    case XD_I32: FOR_OP(int, MODE, X); break;                                 \
    case XD_I64: FOR_OP(long long, MODE, X); break;                           \
    default: exags_fail(1, __FILE__, __LINE__, "gs: bad domain %d", xdom);    \
  }
#define FOR_ALL(X)                                                            \
  switch (mode) {                                                             \
    case XM_PLAIN: FOR_DOM(XM_PLAIN, X); break;                               \
    case XM_VEC:   FOR_DOM(XM_VEC, X); break;                                 \
    case XM_MANY:  FOR_DOM(XM_MANY, X); break;                                \
    default: exags_fail(1, __FILE__, __LINE__, "gs: bad mode %d", mode);      \
  }

/* weighted gs: float/double only */
#define FOR_OP_W(T, X)                                                        \
  switch (xop) {                                                              \
    case XO_ADD: X(T, XO_ADD); break;                                         \
    case XO_MUL: X(T, XO_MUL); break;                                         \
    case XO_MIN: X(T, XO_MIN); break;                                         \
    case XO_MAX: X(T, XO_MAX); break;                                         \
    default: exags_fail(1, __FILE__, __LINE__, "gs_weighted: bad op %d", xop); \
  }
#define FOR_DOM_W(X)                                                          \
  switch (xdom) {                                                             \
    case XD_F64: FOR_OP_W(double, X); break;                                  \
    case XD_F32: FOR_OP_W(float, X); break;                                   \
    default: exags_fail(1, __FILE__, __LINE__, "gs_weighted: floating-point domains only (got %d)", xdom); \
  }

extern "C" {

unsigned long long xcu_launch_count(void) { return g_launches; }

void xcu_fused(void *stream, const xcu_csr *m, const xcu_field *f, int mode, int xdom, int xop,
               int transpose)
{
  if (!m->ng) return;
  cudaStream_t st = (cudaStream_t)stream;
  if (f->w) {
#define X(T, OP)                                                              \
  do {                                                                        \
    if (mode == XM_PLAIN) k_fused_w<T, OP, XM_PLAIN><<<blocks_for(m->ng), 256, 0, st>>>(*m, make_fld<T, XM_PLAIN>(f), (const T *)f->w, transpose); \
    else if (mode == XM_VEC) k_fused_w<T, OP, XM_VEC><<<blocks_for(m->ng), 256, 0, st>>>(*m, make_fld<T, XM_VEC>(f), (const T *)f->w, transpose); \
    else k_fused_w<T, OP, XM_MANY><<<blocks_for(m->ng), 256, 0, st>>>(*m, make_fld<T, XM_MANY>(f), (const T *)f->w, transpose); \
  } while (0)
    FOR_DOM_W(X)
#undef X
    LAUNCHED();
    return;
  }
#define X(T, OP, MODE)                                                        \
  k_fused<T, OP, MODE><<<blocks_for(m->ng), 256, 0, st>>>(*m, make_fld<T, MODE>(f), transpose)
  FOR_ALL(X)
#undef X
  LAUNCHED();
}


void xcu_fused_sell(void *stream, const xcu_sell *m, const xcu_field *f, int mode, int xdom, int xop,
                    int transpose)
{
  if (!m->nblock) return;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = m->nblock / XG_SELL_CW;
  const unsigned nthr = 32 * XG_SELL_CW;
  if (f->w) {
    /* weighted gs: float/double fields (checked by the caller) */
    if (mode == XM_PLAIN || f->vn == 1) {
      xcu_field f1 = *f;
      f1.vn = 1; f1.tuple = 1;
      if (m->fmask) {
#define X(T, OP) k_fused_sell_w<T, OP, XM_PLAIN, 1, 4, true, false><<<grid, nthr, 0, st>>>(*m, make_fld<T, XM_PLAIN>(&f1), (const T *)f->w, transpose)
        FOR_DOM_W(X)
#undef X
      } else {
#define X(T, OP) k_fused_sell_w<T, OP, XM_PLAIN, 1, 4, false, false><<<grid, nthr, 0, st>>>(*m, make_fld<T, XM_PLAIN>(&f1), (const T *)f->w, 0)
        FOR_DOM_W(X)
#undef X
      }
    } else if (mode == XM_VEC) {
      if (m->fmask) {
#define X(T, OP) k_fused_sell_w<T, OP, XM_VEC, 1, 4, true, false><<<grid, nthr, 0, st>>>(*m, make_fld<T, XM_VEC>(f), (const T *)f->w, transpose)
        FOR_DOM_W(X)
#undef X
      } else {
#define X(T, OP) k_fused_sell_w<T, OP, XM_VEC, 3, (sizeof(T) == 8 ? 2 : 4), false, false><<<grid, nthr, 0, st>>>(*m, make_fld<T, XM_VEC>(f), (const T *)f->w, 0)
        FOR_DOM_W(X)
#undef X
      }
    } else {
      if (m->fmask) {
#define X(T, OP) k_fused_sell_w<T, OP, XM_MANY, 1, 4, true, false><<<grid, nthr, 0, st>>>(*m, make_fld<T, XM_MANY>(f), (const T *)f->w, transpose)
        FOR_DOM_W(X)
#undef X
      } else {
        /* up to three arrays: component numbers are compile-time constants (ONE), so the
           pointer table is read with constant indices -- see Fld::comp */
        if (f->vn <= 3) {
#define X(T, OP) k_fused_sell_w<T, OP, XM_MANY, 3, (sizeof(T) == 8 ? 2 : 4), false, true, (sizeof(T) == 8 ? 1 : 2)><<<grid, nthr, 0, st>>>(*m, make_fld<T, XM_MANY>(f), (const T *)f->w, 0)
          FOR_DOM_W(X)
#undef X
        } else {
#define X(T, OP) k_fused_sell_w<T, OP, XM_MANY, 3, (sizeof(T) == 8 ? 2 : 3), false, false><<<grid, nthr, 0, st>>>(*m, make_fld<T, XM_MANY>(f), (const T *)f->w, 0)
          FOR_DOM_W(X)
#undef X
        }
      }
    }
    LAUNCHED();
    return;
  }
  if (m->fmask) {
    /* maps with flagged members: same kernel with the flag mask */
    if (mode == XM_PLAIN || f->vn == 1) {
      xcu_field f1 = *f;
      f1.vn = 1; f1.tuple = 1;
      mode = XM_PLAIN;
#define X(T, OP, MODE)                                                        \
  k_fused_sell<T, OP, XM_PLAIN, 1, 5, false, true><<<grid, nthr, 0, st>>>(*m, make_fld<T, XM_PLAIN>(&f1), transpose)
      FOR_DOM(XM_PLAIN, X)
#undef X
    } else if (mode == XM_VEC) {
#define X(T, OP, MODE)                                                        \
  k_fused_sell<T, OP, XM_VEC, 1, 5, false, true><<<grid, nthr, 0, st>>>(*m, make_fld<T, XM_VEC>(f), transpose)
      FOR_DOM(XM_VEC, X)
#undef X
    } else {
#define X(T, OP, MODE)                                                        \
  k_fused_sell<T, OP, XM_MANY, 1, 5, false, true><<<grid, nthr, 0, st>>>(*m, make_fld<T, XM_MANY>(f), transpose)
      FOR_DOM(XM_MANY, X)
#undef X
    }
    LAUNCHED();
    return;
  }
  if (mode == XM_PLAIN || f->vn == 1) {
    /* one component: a vec/many field with vn == 1 is addressed like a plain one */
    xcu_field f1 = *f;
    f1.vn = 1; f1.tuple = 1;
    mode = XM_PLAIN;
    /* 6 resident CTAs per SM: 8 (32 registers, no spills) was measured in round 2 and is
       no faster for 4-byte elements and 3 % slower for 8-byte ones (profiles/r2_knobs_occupancy_line_shift.txt) */
#define X(T, OP, MODE)                                                        \
  k_fused_sell<T, OP, XM_PLAIN, 1, 6><<<grid, nthr, 0, st>>>(*m, make_fld<T, XM_PLAIN>(&f1), 0)
    FOR_DOM(XM_PLAIN, X)
#undef X
  } else if (mode == XM_VEC && f->vn == 3 && f->tuple == 3 && f->k0 == 0 &&
             ((uintptr_t)f->p[0] % (2 * (xdom == XD_F64 || xdom == XD_I64 ? 8 : 4))) == 0) {
#define X(T, OP, MODE)                                                        \
  k_fused_sell_vec3<T, OP, (sizeof(T) == 8 ? 3 : 6)><<<grid, nthr, 0, st>>>(*m, (T *)f->p[0])
    FOR_DOM(XM_VEC, X)
#undef X
  } else if (mode == XM_VEC) {
#define X(T, OP, MODE)                                                        \
  k_fused_sell<T, OP, XM_VEC, 3, (sizeof(T) == 8 ? 3 : 4)><<<grid, nthr, 0, st>>>(*m, make_fld<T, XM_VEC>(f), 0)
    FOR_DOM(XM_VEC, X)
#undef X
  } else if (f->vn <= 3) {
#define X(T, OP, MODE)                                                        \
  k_fused_sell<T, OP, XM_MANY, 3, (sizeof(T) == 8 ? 2 : 4), false, false, true><<<grid, nthr, 0, st>>>(*m, make_fld<T, XM_MANY>(f), 0)
    FOR_DOM(XM_MANY, X)
#undef X
  } else {
#define X(T, OP, MODE)                                                        \
  k_fused_sell<T, OP, XM_MANY, 3, (sizeof(T) == 8 ? 3 : 4)><<<grid, nthr, 0, st>>>(*m, make_fld<T, XM_MANY>(f), 0)
    FOR_DOM(XM_MANY, X)
#undef X
  }
  LAUNCHED();
}

/* gather-only form for the boundary groups of symmetric maps */
void xcu_gather_sell(void *stream, const xcu_sell *m, const xcu_field *f, int mode, int xdom, int xop)
{
  if (!m->nblock) return;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = m->nblock / XG_SELL_CW;
  const unsigned nthr = 32 * XG_SELL_CW;
  if (mode == XM_PLAIN || f->vn == 1) {
    xcu_field f1 = *f;
    f1.vn = 1; f1.tuple = 1;
    mode = XM_PLAIN;
#define X(T, OP, MODE)                                                        \
  k_fused_sell<T, OP, XM_PLAIN, 1, 6, true><<<grid, nthr, 0, st>>>(*m, make_fld<T, XM_PLAIN>(&f1), 0)
    FOR_DOM(XM_PLAIN, X)
#undef X
  } else if (mode == XM_VEC) {
#define X(T, OP, MODE)                                                        \
  k_fused_sell<T, OP, XM_VEC, 3, (sizeof(T) == 8 ? 3 : 4), true><<<grid, nthr, 0, st>>>(*m, make_fld<T, XM_VEC>(f), 0)
    FOR_DOM(XM_VEC, X)
#undef X
  } else {
#define X(T, OP, MODE)                                                        \
  k_fused_sell<T, OP, XM_MANY, 3, (sizeof(T) == 8 ? 3 : 4), true><<<grid, nthr, 0, st>>>(*m, make_fld<T, XM_MANY>(f), 0)
    FOR_DOM(XM_MANY, X)
#undef X
  }
  LAUNCHED();
}

void xcu_init_list(void *stream, const u32 *list, u32 n, const xcu_field *f, int mode, int xdom,
                   int xop)
{
  if (!n) return;
  cudaStream_t st = (cudaStream_t)stream;
#define X(T, OP, MODE)                                                        \
  k_init_list<T, OP, MODE><<<blocks_for(n), 256, 0, st>>>(list, n, make_fld<T, MODE>(f))
  FOR_ALL(X)
#undef X
  LAUNCHED();
}

void xcu_bnd_pack(void *stream, const xcu_bnd *b, const xcu_field *f, int mode, int xdom, int xop,
                  int transpose, void *sendbuf, unsigned bufvn, int allreduce)
{
  if (!b->nb) return;
  cudaStream_t st = (cudaStream_t)stream;
  xcu_push none; memset(&none, 0, sizeof none);
#define X(T, OP, MODE)                                                        \
  k_bnd_pack<T, OP, MODE, false><<<bnd_blocks(b->nb), 256, 0, st>>>(          \
      *b, make_fld<T, MODE>(f), transpose, (T *)sendbuf, bufvn, allreduce, none)
  FOR_ALL(X)
#undef X
  LAUNCHED();
}

void xcu_bnd_pack_push(void *stream, const xcu_bnd *b, const xcu_field *f, int mode, int xdom,
                       int xop, int transpose, unsigned bufvn, const xcu_push *pp)
{
  cudaStream_t st = (cudaStream_t)stream;
  if (!b->nb) { if (pp->nsig) xcu_push_signal(stream, pp); return; }
#define X(T, OP, MODE)                                                        \
  k_bnd_pack<T, OP, MODE, true><<<bnd_blocks(b->nb), 256, 0, st>>>(           \
      *b, make_fld<T, MODE>(f), transpose, (T *)0, bufvn, 0, *pp)
  FOR_ALL(X)
#undef X
  LAUNCHED();
}

void xcu_pack_slots(void *stream, const u32 *sprim, u32 nslots, const xcu_field *f, int mode,
                    int xdom, void *sendbuf, unsigned bufvn, const xcu_push *pp)
{
  cudaStream_t st = (cudaStream_t)stream;
  if (!nslots) { if (pp && pp->nsig) xcu_push_signal(stream, pp); return; }
  xcu_push none; memset(&none, 0, sizeof none);
  const unsigned nb = bnd_blocks(nslots);
  /* a copy: 8- and 4-byte elements are enough */
#define Y(T, MODE)                                                            \
  do {                                                                        \
    if (pp) k_pack_slots<T, MODE, true><<<nb, 256, 0, st>>>(sprim, nslots, make_fld<T, MODE>(f), (T *)0, bufvn, *pp); \
    else k_pack_slots<T, MODE, false><<<nb, 256, 0, st>>>(sprim, nslots, make_fld<T, MODE>(f), (T *)sendbuf, bufvn, none); \
  } while (0)
#define YM(T)                                                                 \
  switch (mode) {                                                             \
    case XM_PLAIN: Y(T, XM_PLAIN); break;                                     \
    case XM_VEC:   Y(T, XM_VEC); break;                                       \
    case XM_MANY:  Y(T, XM_MANY); break;                                      \
    default: exags_fail(1, __FILE__, __LINE__, "gs: bad mode %d", mode);      \
  }
  switch (xdom) {
    case XD_F64: case XD_I64: YM(long long); break;
    case XD_F32: case XD_I32: YM(int); break;
    default: exags_fail(1, __FILE__, __LINE__, "gs: bad domain %d", xdom);
  }
#undef YM
#undef Y
  LAUNCHED();
}

void xcu_push_signal(void *stream, const xcu_push *pp)
{
  if (!pp->nsig) return;
  k_push_signal<<<1, 32, 0, (cudaStream_t)stream>>>(*pp);
  LAUNCHED();
}

void xcu_push_wait(void *stream, const u32 *flags, const u32 *ranks, u32 n, const u32 *depoch, u32 epoch_imm,
                   double timeout_s, int *status)
{
  if (!n) return;
  k_push_wait<<<1, (n + 31u) & ~31u, 0, (cudaStream_t)stream>>>(
      flags, ranks, n, depoch, epoch_imm, (unsigned long long)(timeout_s * 1e9), status);
  LAUNCHED();
}

void xcu_bnd_unpack(void *stream, const xcu_bnd *b, const xcu_field *f, int mode, int xdom,
                    int xop, int transpose, const void *recvbuf, unsigned bufvn, int allreduce,
                    const u32 *depoch, size_t par_bytes)
{
  if (!b->nb) return;
  cudaStream_t st = (cudaStream_t)stream;
  if (f->w) {
#define X(T, OP)                                                              \
  do {                                                                        \
    if (mode == XM_PLAIN) k_bnd_unpack_w<T, OP, XM_PLAIN><<<bnd_blocks(b->nb), 256, 0, st>>>(*b, make_fld<T, XM_PLAIN>(f), (const T *)f->w, transpose, (const T *)recvbuf, bufvn, allreduce, depoch, par_bytes); \
    else if (mode == XM_VEC) k_bnd_unpack_w<T, OP, XM_VEC><<<bnd_blocks(b->nb), 256, 0, st>>>(*b, make_fld<T, XM_VEC>(f), (const T *)f->w, transpose, (const T *)recvbuf, bufvn, allreduce, depoch, par_bytes); \
    else k_bnd_unpack_w<T, OP, XM_MANY><<<bnd_blocks(b->nb), 256, 0, st>>>(*b, make_fld<T, XM_MANY>(f), (const T *)f->w, transpose, (const T *)recvbuf, bufvn, allreduce, depoch, par_bytes); \
  } while (0)
    FOR_DOM_W(X)
#undef X
    LAUNCHED();
    return;
  }
#define X(T, OP, MODE)                                                        \
  k_bnd_unpack<T, OP, MODE><<<bnd_blocks(b->nb), 256, 0, st>>>(               \
      *b, make_fld<T, MODE>(f), transpose, (const T *)recvbuf, bufvn, allreduce, depoch, par_bytes)
  FOR_ALL(X)
#undef X
  LAUNCHED();
}

void xcu_fill_identity(void *stream, void *out, size_t n, int xdom, int xop)
{
  if (!n) return;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned nb = blocks_for(n); if (nb > 148u * 8u) nb = 148u * 8u;
  const int mode = XM_PLAIN;
#define X(T, OP, MODE) k_fill<T><<<nb, 256, 0, st>>>((T *)out, n, ident<T, OP>())
  FOR_ALL(X)
#undef X
  LAUNCHED();
}

void xcu_array_op(void *stream, void *out, const void *in, size_t n, int xdom, int xop)
{
  if (!n) return;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned nb = blocks_for(n); if (nb > 148u * 8u) nb = 148u * 8u;
  const int mode = XM_PLAIN;
#define X(T, OP, MODE) k_array<T, OP><<<nb, 256, 0, st>>>((T *)out, (const T *)in, n)
  FOR_ALL(X)
#undef X
  LAUNCHED();
}

/* out (device, 1 double) = sum v[i]*w[i]; scratch holds at least xcu_dot_parts() doubles.
   v and w must be 16-byte aligned device pointers (checked by the caller).   */
unsigned xcu_dot_parts(void) { return 148u * 8u; }
void xcu_dot(void *stream, const double *v, const double *w, size_t n, double *scratch, double *out)
{
  cudaStream_t st = (cudaStream_t)stream;
  unsigned nb = (unsigned)((n / 2 + DOT_THREADS - 1) / DOT_THREADS);
  if (nb > xcu_dot_parts()) nb = xcu_dot_parts();
  if (nb == 0) nb = 1;
  k_dot_partial<<<nb, DOT_THREADS, 0, st>>>(v, w, n, scratch);
  LAUNCHED();
  k_dot_final<<<1, DOT_THREADS, 0, st>>>(scratch, nb, out);
  LAUNCHED();
}

void xcu_csr_gather(void *stream, const xcu_csr *m, const xcu_field *out, int omode,
                    const xcu_field *in, int imode, int xdom, int xop)
{
  if (!m->ng) return;
  cudaStream_t st = (cudaStream_t)stream;
  const int mode = XM_PLAIN;
#define X(T, OP, MODE)                                                        \
  k_csr_gather<T, OP><<<blocks_for(m->ng), 256, 0, st>>>(*m, make_fldr<T>(out, omode), make_fldr<T>(in, imode))
  FOR_ALL(X)
#undef X
  LAUNCHED();
}

void xcu_csr_scatter(void *stream, const xcu_csr *m, const xcu_field *out, int omode,
                     const xcu_field *in, int imode, int xdom)
{
  if (!m->ng) return;
  cudaStream_t st = (cudaStream_t)stream;
  switch (xdom) {
    case XD_F64: case XD_I64:
      k_csr_scatter<long long><<<blocks_for(m->ng), 256, 0, st>>>(*m, make_fldr<long long>(out, omode), make_fldr<long long>(in, imode));
      break;
    case XD_F32: case XD_I32:
      k_csr_scatter<int><<<blocks_for(m->ng), 256, 0, st>>>(*m, make_fldr<int>(out, omode), make_fldr<int>(in, imode));
      break;
    default: exags_fail(1, __FILE__, __LINE__, "gs: bad domain %d", xdom);
  }
  LAUNCHED();
}

// ---------------------------------------------------------------------------
// runtime plumbing
// ---------------------------------------------------------------------------
int xcu_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int xcu_set_device(int dev)
{
  CU(cudaSetDevice(dev));
  CU(cudaFree(0));
  return 0;
}

void *xcu_malloc(size_t bytes)
{
  void *p = 0;
  CU(cudaMalloc(&p, bytes ? bytes : 1));
  return p;
}
void xcu_free(void *p) { if (p) CU(cudaFree(p)); }
void xcu_h2d(void *d, const void *s, size_t n, void *st)
{ if (n) CU(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, (cudaStream_t)st)); }
void xcu_d2h(void *d, const void *s, size_t n, void *st)
{ if (n) CU(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, (cudaStream_t)st)); }
void xcu_d2d(void *d, const void *s, size_t n, void *st)
{ if (n) CU(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, (cudaStream_t)st)); }
void xcu_memset(void *d, int byte, size_t n, void *st)
{ if (n) CU(cudaMemsetAsync(d, byte, n, (cudaStream_t)st)); }

int xcu_is_device_ptr(const void *p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return 0; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

/* 0: pageable host memory, 1: pinned or registered host memory, 2: device / managed */
int xcu_ptr_kind(const void *p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return 2;
  return a.type == cudaMemoryTypeHost ? 1 : 0;
}
/* page-lock a caller's array so that copies run at link speed and overlap
   (non-fatal: 0 when the driver refuses, e.g. over the locked-memory limit)   */
int xcu_host_register(void *p, size_t bytes)
{
  if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return 0; }
  return 1;
}
void xcu_host_unregister(void *p) { if (cudaHostUnregister(p) != cudaSuccess) cudaGetLastError(); }

void *xcu_stream_create(int high_priority)
{
  cudaStream_t s;
  int lo = 0, hi = 0;
  CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  /* the main stream is a blocking stream: it orders itself against the
     legacy default stream, so a blocking gs() on a device pointer sees work
     the caller queued there (e.g. torch ops) without an explicit sync */
  CU(cudaStreamCreateWithPriority(&s, high_priority ? cudaStreamNonBlocking : cudaStreamDefault,
                                  high_priority ? hi : lo));
  return (void *)s;
}
void xcu_stream_destroy(void *s) { if (s) CU(cudaStreamDestroy((cudaStream_t)s)); }
void xcu_stream_sync(void *s) { CU(cudaStreamSynchronize((cudaStream_t)s)); }
void *xcu_event_create(void)
{
  cudaEvent_t e;
  CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return (void *)e;
}
void *xcu_event_create_timed(void)
{
  cudaEvent_t e;
  CU(cudaEventCreate(&e));
  return (void *)e;
}
void xcu_event_destroy(void *ev) { if (ev) CU(cudaEventDestroy((cudaEvent_t)ev)); }
float xcu_event_ms(void *a, void *b)
{
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return ms;
}
void xcu_event_sync(void *ev) { CU(cudaEventSynchronize((cudaEvent_t)ev)); }
int xcu_event_done(void *ev)
{
  cudaError_t e = cudaEventQuery((cudaEvent_t)ev);
  if (e == cudaSuccess) return 1;
  if (e != cudaErrorNotReady) CU(e);
  return 0;
}
void xcu_event_record(void *ev, void *st) { CU(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)st)); }
void xcu_stream_wait_event(void *st, void *ev) { CU(cudaStreamWaitEvent((cudaStream_t)st, (cudaEvent_t)ev, 0)); }
void xcu_device_sync(void) { CU(cudaDeviceSynchronize()); }
/* CUDA IPC, for ranks that share a GPU (NCCL refuses two ranks on one device) */
void xcu_ipc_export(void *handle64, void *devptr)
{
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, devptr));
  static_assert(sizeof(h) == 64, "ipc handle size");
  memcpy(handle64, &h, 64);
}
void *xcu_ipc_open(const void *handle64)
{
  cudaIpcMemHandle_t h;
  void *p = 0;
  memcpy(&h, handle64, 64);
  CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  return p;
}
void xcu_ipc_close(void *p) { if (p) CU(cudaIpcCloseMemHandle(p)); }
/* non-fatal open, for the start-up probe of the push transport */
void *xcu_ipc_try_open(const void *handle64)
{
  cudaIpcMemHandle_t h;
  void *p = 0;
  memcpy(&h, handle64, 64);
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return 0; }
  return p;
}
void xcu_device_uuid(int dev, char out16[16])
{
  cudaDeviceProp pr;
  CU(cudaGetDeviceProperties(&pr, dev));
  memcpy(out16, &pr.uuid, 16);
}

/* can `dev` store into the device whose uuid is given?  (same device: yes) */
int xcu_peer_ok(int dev, const char uuid16[16])
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  for (int d = 0; d < n; ++d) {
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, d) != cudaSuccess) { cudaGetLastError(); continue; }
    if (memcmp(&pr.uuid, uuid16, 16) != 0) continue;
    if (d == dev) return 1;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, dev, d) != cudaSuccess) { cudaGetLastError(); return 0; }
    return can;
  }
  return 0;   /* not visible to this process */
}

void *xcu_host_alloc_mapped(size_t bytes)
{
  void *p = 0;
  CU(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocMapped | cudaHostAllocPortable));
  return p;
}

void *xcu_host_alloc(size_t bytes)
{
  void *p = 0;
  CU(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault));
  return p;
}
void xcu_host_free(void *p) { if (p) CU(cudaFreeHost(p)); }

} // extern "C"
