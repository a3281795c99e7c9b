// DFMA throughput with 1 / 2 / 3 register operands, 8 warps per SMSP.  nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void k(double* out, int iters, double a, double b) {
  double x[8], y[8], z[8];
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x + i; y[i] = 1.0 + 1e-9 * (threadIdx.x + i); z[i] = 1e-7 * (i + threadIdx.x); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (KIND == 0) x[i] = fma(x[i], a, b);
        if (KIND == 1) x[i] = fma(x[i], y[i], b);
        if (KIND == 2) x[i] = fma(x[i], y[i], z[i]);
        if (KIND == 3) x[i] = fma(x[i], y[(i + 1) & 7], z[(i + 3) & 7]);
        if (KIND == 4) x[i] = x[i] * y[i];
        if (KIND == 5) x[i] = x[i] + z[i];
      }
    }
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += x[i] + y[i] + z[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND> void run(const char* name) {
  double* o; cudaMalloc(&o, 148 * 4 * 256 * 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 4000;
  k<KIND><<<148 * 4, 256>>>(o, 10, 0.999, 1e-7);
  cudaEventRecord(a); k<KIND><<<148 * 4, 256>>>(o, iters, 0.999, 1e-7); cudaEventRecord(b);
  cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b);
  double cyc = ms * 1e-3 * 1.965e9 / iters / 32.0 / 8.0;   // per warp per instruction, 8 warps per SMSP
  printf("%-34s %.3f ms  %.3f SMSP-cycles per warp-instruction\n", name, ms, cyc);
}
int main() {
  run<0>("DFMA R, R, c, c");
  run<1>("DFMA R, R, R, c");
  run<2>("DFMA R, R, R, R (same idx)");
  run<3>("DFMA R, R, R, R (mixed idx)");
  run<4>("DMUL R, R, R");
  run<5>("DADD R, R, R");
  return 0;
}
