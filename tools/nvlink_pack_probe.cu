// tools/nvlink_pack_probe.cu -- the push transport's pack kernel (k_pack_slots<PUSH>, the literal
// counterpart of pw_exec_sends, src/gs.upc:437-468) in ONE process with two GPUs, so that Nsight
// Compute can count its NVLink bytes (ncu cannot follow the multi-process CUDA-IPC job).  The
// library's own kernel is compiled in from csrc/gs_cuda.cu; device 0 packs u[sprim[s]] into a
// "mailbox" that lives in device 1's memory (peer access), exactly as a rank stores into its
// neighbour's IPC-mapped mailbox.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Iinclude \
//        -Iexaflow-exags-upc_b200/csrc tools/nvlink_pack_probe.cu -o tools/nvlink_pack_probe
//   tools/nvlink_pack_probe [nslots]          (201601 = M7 slab interface, 4363087 = config 5 at P=8)
//   ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum -k regex:k_pack_slots tools/nvlink_pack_probe
#include <cstdarg>
#include <vector>
#include "gs_cuda.cu"

extern "C" void exags_fail(int status, const char *file, unsigned line, const char *fmt, ...)
{
  va_list ap; va_start(ap, fmt);
  fprintf(stderr, "ERROR (%s:%u): ", file, line); vfprintf(stderr, fmt, ap); fprintf(stderr, "\n");
  va_end(ap);
  exit(status);
}

int main(int argc, char **argv)
{
  const u32 nslots = argc > 1 ? (u32)atol(argv[1]) : 201601u;
  const size_t n = 100663296;                    // M7 field
  int ndev = 0; cudaGetDeviceCount(&ndev);
  if (ndev < 2) { printf("needs 2 GPUs\n"); return 0; }
  char *peer_box = nullptr;
  CU(cudaSetDevice(1));
  CU(cudaMalloc(&peer_box, 4096 + 2 * (size_t)nslots * 8));
  CU(cudaMemset(peer_box, 0, 4096));
  CU(cudaSetDevice(0));
  CU(cudaDeviceEnablePeerAccess(1, 0));
  double *u; u32 *sprim; char *my_box;
  CU(cudaMalloc(&u, n * 8)); CU(cudaMemset(u, 0, n * 8));
  CU(cudaMalloc(&my_box, 4096)); CU(cudaMemset(my_box, 0, 4096));
  // the send slots of a z-slab interface: the top layer of nodes of the top layer of elements, in slot (= id) order
  std::vector<u32> h(nslots);
  for (u32 s = 0; s < nslots; ++s) h[s] = (u32)((n - 1) - (size_t)(nslots - 1 - s) * ((n / 2) / nslots));
  CU(cudaMalloc(&sprim, (size_t)nslots * 4));
  CU(cudaMemcpy(sprim, h.data(), (size_t)nslots * 4, cudaMemcpyHostToDevice));
  // device tables of the push transport (gs_host.c, push_prepare): one neighbour
  unsigned long long hb[2] = { (unsigned long long)(uintptr_t)(peer_box + 4096),
                               (unsigned long long)(uintptr_t)(peer_box + 4096 + (size_t)nslots * 8) };
  unsigned long long hs[1] = { (unsigned long long)(uintptr_t)peer_box };
  u32 hd[1] = { 0 }, hst[2] = { 0, nslots };
  unsigned long long *d_base, *d_sig; u32 *d_delta, *d_start;
  CU(cudaMalloc(&d_base, 16)); CU(cudaMalloc(&d_sig, 8)); CU(cudaMalloc(&d_delta, 4)); CU(cudaMalloc(&d_start, 8));
  CU(cudaMemcpy(d_base, hb, 16, cudaMemcpyHostToDevice)); CU(cudaMemcpy(d_sig, hs, 8, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_delta, hd, 4, cudaMemcpyHostToDevice)); CU(cudaMemcpy(d_start, hst, 8, cudaMemcpyHostToDevice));
  xcu_push pp; memset(&pp, 0, sizeof pp);
  pp.base = d_base; pp.par_stride = 1; pp.delta = d_delta; pp.sig = d_sig; pp.nsig = 1;
  pp.depoch = (u32 *)(my_box + XG_MBOX_EPOCH); pp.counter = (u32 *)(my_box + XG_MBOX_COUNTER);
  pp.start = d_start; pp.nnb = 1;
  xcu_field f; memset(&f, 0, sizeof f);
  f.p[0] = u; f.vn = 1; f.tuple = 1;
  cudaStream_t st; CU(cudaStreamCreate(&st));
  for (int it = 0; it < 3; ++it) xcu_pack_slots(st, sprim, nslots, &f, XM_PLAIN, XD_F64, nullptr, 1, &pp);
  cudaEvent_t e0, e1; CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  CU(cudaEventRecord(e0, st));
  const int reps = 50;
  for (int it = 0; it < reps; ++it) xcu_pack_slots(st, sprim, nslots, &f, XM_PLAIN, XD_F64, nullptr, 1, &pp);
  CU(cudaEventRecord(e1, st));
  CU(cudaStreamSynchronize(st));
  float ms = 0; CU(cudaEventElapsedTime(&ms, e0, e1));
  const double us = 1e3 * ms / reps, bytes = 8.0 * nslots;
  unsigned flag = 0, epoch = 0;
  CU(cudaMemcpy(&flag, peer_box, 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(&epoch, my_box + XG_MBOX_EPOCH, 4, cudaMemcpyDeviceToHost));
  printf("k_pack_slots<PUSH>: %u slots (%.2f MB payload) device 0 -> device 1: %.1f us per launch, %.1f GB/s payload "
         "(%.1f %% of 900 GB/s per direction); peer flag %u, exchange number %u\n",
         nslots, bytes / 1e6, us, bytes / us / 1e3, 100.0 * bytes / us / 1e3 / 900.0, flag, epoch);
  return flag == epoch && epoch == (unsigned)(reps + 3) ? 0 : 1;
}
