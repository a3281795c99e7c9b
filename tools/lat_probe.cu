// Latency probe: dependent chains, one warp per SMSP.  nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int CH, int KIND>
__global__ void k(double* out, int iters, double a, double b) {
  double x[CH]; float y[CH];
  for (int i = 0; i < CH; ++i) { x[i] = threadIdx.x + i + 1.0; y[i] = 0.1f * (threadIdx.x + i + 1); }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        if (KIND == 0) x[i] = fma(x[i], a, b);
        if (KIND == 1) y[i] = __sinf(y[i]);
        if (KIND == 2) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[i])); x[i] = r; }
        if (KIND == 3) y[i] = fmaf(y[i], 0.999f, 0.5f);
        if (KIND == 4) x[i] = x[i] * a;
        if (KIND == 5) x[i] = x[i] + b;
        if (KIND == 6) { y[i] = (float)x[i]; x[i] = (double)y[i] + b; }
      }
    }
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < CH; ++i) s += x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[4096] = (double)(t1 - t0) / (iters * 16.0);
}
template <int CH, int KIND> void run(const char* name, int threads) {
  double* o; cudaMalloc(&o, 8 * 8192);
  k<CH, KIND><<<1, threads>>>(o, 2000, 0.999, 1e-7);
  cudaDeviceSynchronize();
  double c; cudaMemcpy(&c, o + 4096, 8, cudaMemcpyDeviceToHost);
  printf("%-10s chains=%d warps/SMSP=%d : %.2f cycles per step of all chains (%.2f per op)\n", name, CH, threads / 128 ? threads / 128 : 1, c, c / CH);
  cudaFree(o);
}
int main() {
  run<1, 0>("DFMA", 128); run<2, 0>("DFMA", 128); run<4, 0>("DFMA", 128); run<8, 0>("DFMA", 128);
  run<4, 0>("DFMA", 256); run<4, 0>("DFMA", 512);
  run<1, 4>("DMUL", 128); run<1, 5>("DADD", 128);
  run<1, 1>("MUFU.SIN", 128); run<4, 1>("MUFU.SIN", 128);
  run<1, 2>("RCP64H", 128); run<4, 2>("RCP64H", 128);
  run<1, 3>("FFMA", 128); run<1, 6>("cvt pair", 128);
  return 0;
}
