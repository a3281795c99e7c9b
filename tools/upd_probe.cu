// Microbenchmark of K3's trailing-update step: a TT x TT register tile, per k-step 2*TT shared loads + TT*TT DFMAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o upd_probe upd_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int TT = 6;
template <int MODE>   // 0: as K3 (stride TT*TT doubles per tile row)  1: padded stride (TT*TT + 2)  2: no shared loads (registers)
__global__ void __launch_bounds__(256, 2) upd(double* out, long long* cyc, int iters, int nt) {
  extern __shared__ double pbuf[];
  const int tid = threadIdx.x;
  int J = 0, rem = tid;
  while (J < nt && rem >= nt - J) { rem -= nt - J; ++J; }
  const int I = J + rem;
  const int stride = MODE == 1 ? TT * TT + 2 : TT * TT;
  for (int i = tid; i < (nt + 1) * stride; i += blockDim.x) pbuf[i] = 1e-3 * (i % 17);
  __syncthreads();
  double a[TT][TT];
#pragma unroll
  for (int r = 0; r < TT; ++r)
#pragma unroll
    for (int c = 0; c < TT; ++c) a[r][c] = r + c + tid;
  const int Ic = I < nt ? I : 0, Jc = J < nt ? J : 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const double2* pi = reinterpret_cast<const double2*>(pbuf + Ic * stride);
    const double2* pj = reinterpret_cast<const double2*>(pbuf + Jc * stride);
#pragma unroll
    for (int k = 0; k < TT; ++k) {
      double Li[TT], Lk[TT];
#pragma unroll
      for (int r = 0; r < TT; r += 2) {
        double2 u, w;
        if (MODE == 2) { u = make_double2(a[r][0] * 1e-9, a[r + 1][0] * 1e-9); w = make_double2(a[0][r] * 1e-9, a[0][r + 1] * 1e-9); }
        else { u = pi[(k * TT + r) / 2]; w = pj[(k * TT + r) / 2]; }
        Li[r] = u.x; Li[r + 1] = u.y; Lk[r] = w.x; Lk[r + 1] = w.y;
      }
#pragma unroll
      for (int r = 0; r < TT; ++r)
#pragma unroll
        for (int c = 0; c < TT; ++c) a[r][c] = fma(-Li[r], Lk[c], a[r][c]);
    }
    if (MODE != 2) asm volatile("" ::: "memory");
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int r = 0; r < TT; ++r)
#pragma unroll
    for (int c = 0; c < TT; ++c) s += a[r][c];
  out[blockIdx.x * blockDim.x + tid] = s;
  if (tid == 230) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int grid, int nt) {
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * grid * 256); cudaMalloc(&cyc, sizeof(long long) * grid);
  const int iters = 2000, smem = (nt + 1) * (TT * TT + 2) * 8;
  upd<MODE><<<grid, 256, smem>>>(out, cyc, iters, nt);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  upd<MODE><<<grid, 256, smem>>>(out, cyc, iters, nt);
  cudaEventRecord(b); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s grid %4d: %8.1f cycles per update (thread 230, clock64), %8.1f from event time; err %s\n", name, grid,
         (double)h / iters, ms * 1e-3 * 1.965e9 / iters, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int grid : {148, 296}) {
    run<0>("K3 layout (stride 36)", grid, 21);
    run<1>("padded (stride 38)", grid, 21);
    run<2>("no shared loads", grid, 21);
  }
  return 0;
}
