// microbench.cu -- what HBM delivers for the access shapes gs_op produces.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_rmw8(double *u, size_t n) { size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) u[i] += 1.0; }
__global__ void k_rmw16(double2 *u, size_t n2) { size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n2) { double2 v = u[i]; v.x += 1; v.y += 1; u[i] = v; } }
// one double per 32-byte sector
__global__ void k_sector(double *u, size_t nsec) { size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (s < nsec) u[s * 4] += 1.0; }
// x-face shape: element e, row r (64 rows of 8): pair (e, a=7) with (e+1, a=0)
__global__ void k_xface(double *u, size_t nelem) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t e = t / 64, r = t % 64;
  if (e + 1 >= nelem) return;
  double *a = u + e * 512 + r * 8 + 7, *b = u + (e + 1) * 512 + r * 8;
  double v = *a + *b; *a = v; *b = v;
}
// both sectors of every row, each touched by a different thread far apart in the grid (2 passes in one launch)
__global__ void k_copy(const double *in, double *out, size_t n) { size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = in[i]; }

int main() {
  const size_t n = 100663296;
  double *u, *w; CK(cudaMalloc(&u, n * 8)); CK(cudaMalloc(&w, n * 8)); CK(cudaMemset(u, 0, n * 8));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
#define RUN(name, bytes, launch) \
  for (int it = 0; it < 3; ++it) { launch; } cudaEventRecord(e0); for (int it = 0; it < 10; ++it) { launch; } cudaEventRecord(e1); \
  CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); ms /= 10; \
  printf("%-28s %.4f ms  %.0f GB/s\n", name, ms, (bytes) / ms / 1e6);
  RUN("copy 8B (r+w 1.61GB)", 2.0 * n * 8, (k_copy<<<(n + 255) / 256, 256>>>(u, w, n)));
  RUN("rmw 8B/thread", 2.0 * n * 8, (k_rmw8<<<(n + 255) / 256, 256>>>(u, n)));
  RUN("rmw 16B/thread", 2.0 * n * 8, (k_rmw16<<<(n / 2 + 255) / 256, 256>>>((double2 *)u, n / 2)));
  RUN("rmw 1 double per sector", 2.0 * n * 8, (k_sector<<<(n / 4 + 255) / 256, 256>>>(u, n / 4)));
  RUN("x-face pairs (all sectors)", 2.0 * n * 8, (k_xface<<<(n / 512 * 64 + 255) / 256, 256>>>(u, n / 512)));
  return 0;
}
