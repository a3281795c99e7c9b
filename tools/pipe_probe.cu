// Micro-probe: do the FP64 and XU (MUFU / F2F) pipes of sm_100a overlap?  nvcc -arch=sm_100a -O3 pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ND, int NM, int NF, int NC>
__global__ void k(double* out, float* fout, int iters, double a, double b, float fa) {
  double x[8]; float y[8]; float z[8];
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x + i; y[i] = 0.001f * threadIdx.x + i; z[i] = y[i]; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int j = 0; j < ND; ++j) x[(u + j) & 7] = fma(x[(u + j) & 7], a, b);
#pragma unroll
      for (int j = 0; j < NM; ++j) y[(u + j) & 7] = __sinf(y[(u + j) & 7]);
#pragma unroll
      for (int j = 0; j < NF; ++j) z[(u + j) & 7] = fmaf(z[(u + j) & 7], fa, 0.5f);
#pragma unroll
      for (int j = 0; j < NC; ++j) x[(u + j) & 7] += (double)z[(u + j) & 7];
    }
  }
  double s = 0; float t = 0;
  for (int i = 0; i < 8; ++i) { s += x[i]; t += y[i] + z[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s; fout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int ND, int NM, int NF, int NC> void run(const char* name) {
  double* o; float* f; cudaMalloc(&o, 148 * 8 * 256 * 8); cudaMalloc(&f, 148 * 8 * 256 * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 2000;
  k<ND, NM, NF, NC><<<148 * 4, 256>>>(o, f, 10, 0.999, 1e-7, 0.999f);
  cudaEventRecord(a); k<ND, NM, NF, NC><<<148 * 4, 256>>>(o, f, iters, 0.999, 1e-7, 0.999f); cudaEventRecord(b);
  cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b);
  // cycles per SMSP per loop iteration (8 unrolled groups): 8 warps per SMSP (4 CTAs x 8 warps / 4)
  double cyc = ms * 1e-3 * 1.965e9 / iters / 8.0 / 8.0;   // per warp per group
  printf("%-28s ND=%d NM=%d NF=%d NC=%d  %.3f ms  %.2f SMSP-cycles per warp-group\n", name, ND, NM, NF, NC, ms, cyc);
  cudaFree(o); cudaFree(f);
}
int main() {
  run<8, 0, 0, 0>("dfma only");
  run<0, 2, 0, 0>("mufu only");
  run<8, 2, 0, 0>("dfma + mufu");
  run<8, 1, 0, 0>("dfma + 1 mufu");
  run<0, 0, 8, 0>("ffma only");
  run<8, 0, 8, 0>("dfma + ffma");
  run<8, 0, 16, 0>("dfma + 16 ffma");
  run<0, 0, 0, 2>("cvt+dadd only");
  run<8, 0, 0, 2>("dfma + cvt+dadd");
  return 0;
}
