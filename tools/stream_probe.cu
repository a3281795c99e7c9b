// Streaming-rate probe for column-block reads of a row-major [S, T] fp64 matrix (design input for K6).
// CTA = CB adjacent columns x a slab of rows; thread (c = tid % CB, rl = tid / CB); U loads in flight per thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int CB, int U, int V>   // V = doubles per thread load (1 or 2)
__global__ void __launch_bounds__(256) probe(const double* __restrict__ A, int64_t S, int64_t T, double* out) {
  const int tid = threadIdx.x;
  constexpr int TPR = CB / V;               // threads per row segment
  constexpr int RPI = 256 / TPR;            // rows per iteration
  const int c = (tid % TPR) * V, rl = tid / TPR;
  const int64_t c0 = (int64_t)blockIdx.x * CB;
  const int64_t rows_per = (S + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = blockIdx.y * rows_per, r1 = min(S, r0 + rows_per);
  double acc = 0;
  if (c0 + c < T) {
    const double* col = A + c0 + c;
    for (int64_t r = r0 + rl; r < r1; r += (int64_t)RPI * U) {
      double x[U][V];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t rr = r + (int64_t)u * RPI;
        if (V == 2) {
          double2 v = rr < r1 ? __ldcs(reinterpret_cast<const double2*>(col + rr * T)) : make_double2(0, 0);
          x[u][0] = v.x; x[u][V - 1] = v.y;
        } else {
          x[u][0] = rr < r1 ? __ldcs(col + rr * T) : 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int v = 0; v < V; ++v) acc += x[u][v];
    }
  }
  if (acc == 123.456) out[0] = acc;
}
template <int CB, int U, int V>
void run(const double* A, int64_t S, int64_t T, double* out, int ctas_per_sm) {
  const int64_t ncb = (T + CB - 1) / CB;
  int64_t split = (148 * ctas_per_sm) / ncb;
  if (split < 1) split = 1;
  dim3 grid(ncb, split);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  probe<CB, U, V><<<grid, 256>>>(A, S, T, out);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) probe<CB, U, V><<<grid, 256>>>(A, S, T, out);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
  printf("CB=%3d U=%d V=%d grid=(%lld,%lld) %.3f ms  %.0f GB/s\n", CB, U, V, (long long)ncb, (long long)split, ms, S * T * 8.0 / ms / 1e6);
}
int main() {
  const int64_t S = 100000, T = 1000;
  double *A, *out; cudaMalloc(&A, S * T * 8); cudaMalloc(&out, 8); cudaMemset(A, 0, S * T * 8);
  for (int cps : {4, 8}) {
    printf("-- %d CTAs per SM\n", cps);
    run<8, 4, 1>(A, S, T, out, cps); run<8, 8, 1>(A, S, T, out, cps); run<8, 16, 1>(A, S, T, out, cps);
    run<16, 8, 1>(A, S, T, out, cps); run<16, 8, 2>(A, S, T, out, cps);
    run<32, 8, 1>(A, S, T, out, cps); run<32, 8, 2>(A, S, T, out, cps);
    run<64, 8, 2>(A, S, T, out, cps); run<128, 8, 2>(A, S, T, out, cps); run<128, 4, 2>(A, S, T, out, cps);
  }
  return 0;
}
