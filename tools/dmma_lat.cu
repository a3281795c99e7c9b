// DMMA m8n8k4 f64 latency on B200: one warp per SM sub-partition, NC independent accumulator chains.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_lat tools/dmma_lat.cu && tools/dmma_lat
#include <cstdio>
#include <cuda_runtime.h>
template <int NC>
__global__ void k(double* out, long long* cyc, int iters, double a0, double b0) {
  double c[NC][2];
  for (int i = 0; i < NC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  const long long t1 = clock64();
  double s = 0; for (int i = 0; i < NC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NC> void run(double* out, long long* cyc, int threads) {
  const int iters = 2048;
  k<NC><<<148, threads>>>(out, cyc, iters, 0.999999, 1e-7);
  k<NC><<<148, threads>>>(out, cyc, iters, 0.999999, 1e-7);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("warps/SM=%2d chains=%d : %.1f cycles per step (%.1f per DMMA)\n", threads / 32, NC, (double)h / iters, (double)h / iters / NC);
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  run<1>(out, cyc, 32); run<2>(out, cyc, 32); run<4>(out, cyc, 32); run<8>(out, cyc, 32);
  run<1>(out, cyc, 128); run<2>(out, cyc, 128); run<4>(out, cyc, 128);
  run<1>(out, cyc, 384); run<2>(out, cyc, 384);
  return 0;
}
