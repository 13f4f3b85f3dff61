// Instruction-fetch microbenchmark for sm_100a: straight-line FFMA bodies of N instructions executed in a loop
// by W warps per SM, (a) all warps aligned, (b) warps phase-shifted by a start delay, (c) with a
// forward branch (never-taken "slow path" that is jumped over) every 16 instructions.
// Prints warp-instructions per clock per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

template <int N, bool BR, int SYNC>
__global__ void __launch_bounds__(512, 1) body_kernel(float *out, int iters, int delay, int never, long long *cyc, int drift) {
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 0.001f + k;
  const float x = 1.0001f, c = 0.5f;
  const int warp = threadIdx.x >> 5;
  // phase shift
  long long t0 = clock64();
  while (clock64() - t0 < (long long)delay * warp) {}
  __syncwarp();
  long long start = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      a[i & 7] = fmaf(a[i & 7], x, c);
      if (SYNC && (i % SYNC) == SYNC - 1) {
        __syncthreads();
        __nanosleep((warp & 1) ? drift : 0);
      }
      if (BR && (i & 15) == 15) {
        if (never) {          // uniform, false at run time: the 4-instruction block below is jumped over
          a[0] = a[0] * a[1] + a[2]; a[3] = a[3] * a[4] + a[5]; a[6] = a[6] * a[7] + a[0]; a[1] += a[3];
        }
      }
    }
  }
  long long stop = clock64();
  float s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = stop - start;
}

template <int N, bool BR, int SYNC>
void run(int warps, int delay, float *out, long long *cyc, int drift = 0) {
  const int total = 1 << 22;               // instructions per warp overall
  const int iters = total / N;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  body_kernel<N, BR, SYNC><<<148, warps * 32>>>(out, 2, delay, 0, cyc, drift);
  cudaEventRecord(e0);
  body_kernel<N, BR, SYNC><<<148, warps * 32>>>(out, iters, delay, 0, cyc, drift);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double instr = (double)iters * N * warps;     // per SM
  printf("drift=%4d sync=%4d N=%6d (%4d KB) br=%d warps/SM=%2d delay=%5d : %.3f warp-instr/clk/SMSP  (%.2f ms)\n", drift, SYNC, N, N * 16 / 1024, (int)BR, warps, delay,
         instr / avg / 4.0, ms);
}

int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 148 * 512 * sizeof(float)); cudaMalloc(&cyc, 148 * sizeof(long long));
  for (int warps : {8, 16}) {
    for (int drift : {0, 100, 500, 2000}) {
      run<8192, false, 2048>(warps, 3000, out, cyc, drift);
      run<8192, false, 256>(warps, 3000, out, cyc, drift);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
