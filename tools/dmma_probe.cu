// DMMA (mma.sync m8n8k4 / m16n8k4 / m16n8k8 f64) issue rate on B200 vs DFMA: is the fp64 tensor path worth using for
// K3's trailing updates?  Each warp keeps NC independent accumulator tiles.
#include <cstdio>
#include <cuda_runtime.h>
template <int NC>
__global__ void __launch_bounds__(256) dmma884(double* out, int iters, double a0, double b0) {
  double c[NC][2];
  for (int i = 0; i < NC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0; for (int i = 0; i < NC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NC>
__global__ void __launch_bounds__(256) dmma1688(double* out, int iters, double a0, double b0) {
  double c[NC][4];
  for (int i = 0; i < NC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  double a[4] = {a0, a0 + 1e-9, a0 + 2e-9, a0 + threadIdx.x * 1e-9}, b[2] = {b0, b0 * 0.5};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NC; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
  }
  double s = 0; for (int i = 0; i < NC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) dfma(double* out, int iters, double a0, double b0) {
  double x[8]; for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  const double y = a0 + 1e-12 * threadIdx.x;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(x[i], y, b0);
  double s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  const int grid = 148 * 8, iters = 4096;
  const double warps = (double)grid * 8;
  float ms = timeit([&] { dfma<<<grid, 256>>>(out, iters, 0.999999, 1e-7); });
  printf("DFMA            : %.3f ms  %.2f TFLOP/s\n", ms, 2.0 * 64 * iters * grid * 256 / ms / 1e9);
#define RUN884(NC) ms = timeit([&] { dmma884<NC><<<grid, 256>>>(out, iters, 0.999999, 1e-7); }); \
  printf("DMMA m8n8k4  NC=%d: %.3f ms  %.2f TFLOP/s  %.2f SMSP-cycles per mma\n", NC, ms, 512.0 * NC * iters * warps / ms / 1e9, \
         ms * 1e-3 * 1.965e9 / (NC * (double)iters * warps / (148 * 4)));
  RUN884(1) RUN884(2) RUN884(4) RUN884(8)
#define RUN1688(NC) ms = timeit([&] { dmma1688<NC><<<grid, 256>>>(out, iters, 0.999999, 1e-7); }); \
  printf("DMMA m16n8k8 NC=%d: %.3f ms  %.2f TFLOP/s  %.2f SMSP-cycles per mma\n", NC, ms, 2048.0 * NC * iters * warps / ms / 1e9, \
         ms * 1e-3 * 1.965e9 / (NC * (double)iters * warps / (148 * 4)));
  RUN1688(1) RUN1688(2) RUN1688(4)
  return 0;
}
