// Front-end microbenchmark 2: straight-line FFMA code with a forward branch every 16 instructions that
// jumps over 4 instructions (taken when skip != 0).  Macro-unrolled so that the code size is exact.
#include <cstdio>
#include <cuda_runtime.h>
#define F4 a0 = fmaf(a0, x, c); a1 = fmaf(a1, x, c); a2 = fmaf(a2, x, c); a3 = fmaf(a3, x, c);
#define F16 F4 a4 = fmaf(a4, x, c); a5 = fmaf(a5, x, c); a6 = fmaf(a6, x, c); a7 = fmaf(a7, x, c); F4 a4 = fmaf(a4, x, c); a5 = fmaf(a5, x, c); a6 = fmaf(a6, x, c); a7 = fmaf(a7, x, c);
#define BLK F16 if (!skip) { a0 = a0 * a1 + a2; a3 = a3 * a4 + a5; a6 = a6 * a7 + a0; a1 = a1 * a3 + a6; }
#define B4 BLK BLK BLK BLK
#define B16 B4 B4 B4 B4
#define B64 B16 B16 B16 B16
#define B256 B64 B64 B64 B64

template <int KIND>
__global__ void __launch_bounds__(512, 1) k(float *out, int iters, int skip, long long *cyc) {
  float a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  const float x = 1.0001f, c = 0.5f;
  long long start = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (KIND == 0) { B64 }            // 64 blocks * 20 instr = 1280 instr = 20 KB
    if (KIND == 1) { B256 B256 }      // 512 blocks = 10240 instr = 160 KB
  }
  long long stop = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0) cyc[blockIdx.x] = stop - start;
}

template <int KIND>
void run(int warps, int skip, float *out, long long *cyc) {
  const int blocks = KIND == 0 ? 64 : 512;
  const int iters = (1 << 16) / blocks * 4;
  k<KIND><<<148, warps * 32>>>(out, 2, skip, cyc);
  k<KIND><<<148, warps * 32>>>(out, iters, skip, cyc);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double per_block = skip ? 17.0 : 21.0;     // executed instructions per block (16 + branch [+4])
  printf("code=%3d KB skip=%d warps/SM=%2d : %.3f executed warp-instr/clk/SMSP\n", blocks * 21 * 16 / 1024, skip, warps,
         (double)iters * blocks * per_block * warps / avg / 4.0);
}

int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 148 * 512 * sizeof(float)); cudaMalloc(&cyc, 148 * sizeof(long long));
  for (int warps : {4, 8, 16})
    for (int skip : {0, 1}) { run<0>(warps, skip, out, cyc); run<1>(warps, skip, out, cyc); }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
