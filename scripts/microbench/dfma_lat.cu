// FP64 pipe microbenchmark for sm_100a: dependent-issue latency and issue interval of DFMA, and how warps per
// SM sub-partition add up.  One block per SM; W warps per block; each warp runs C independent DFMA chains.
// Prints cycles per DFMA per warp, and DFMA per clock per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void chain_kernel(double *out, int iters, double a, double b, long long *cyc) {
  double x[C];
#pragma unroll
  for (int k = 0; k < C; ++k) x[k] = threadIdx.x * 1e-3 + k;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int k = 0; k < C; ++k) x[k] = fma(x[k], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < C; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int C>
void run(int warps, double *out, long long *cyc) {
  const int iters = 2048;
  chain_kernel<C><<<148, warps * 32>>>(out, 16, 0.999999, 1e-9, cyc);
  chain_kernel<C><<<148, warps * 32>>>(out, iters, 0.999999, 1e-9, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  const double per_warp = (double)iters * 16 * C;       // DFMA per warp
  const double wps = warps / 4.0;                       // warps per SM sub-partition (blocks spread evenly)
  printf("chains %d warps/SM %2d: %.2f cycles per DFMA per warp, %.3f DFMA/clk/SMSP\n", C, warps, c / per_warp,
         per_warp * (wps < 1 ? 1 : wps) / c);
}

int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMalloc(&cyc, 148 * sizeof(long long));
  for (int w : {1, 4, 8, 12, 16, 32}) { run<1>(w, out, cyc); run<2>(w, out, cyc); run<4>(w, out, cyc); run<8>(w, out, cyc); }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
