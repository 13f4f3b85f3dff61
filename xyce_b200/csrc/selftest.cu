// xyce_b200 -- diagnostics: run the fast-variant elementary functions (xb_fastmath.h) on arrays so that
// tests can measure their error in ulps against a correctly rounded host result.
#include "ctx.h"
#include "xb_fastmath.h"

namespace {
__global__ void fastmath_kernel(int which, int n, const double *a, const double *b, double *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  switch (which) {
    case 0: out[i] = xb::fm::exp(a[i]); break;
    case 1: out[i] = xb::fm::log(a[i]); break;
    case 3: out[i] = xb::fm::sqrt(a[i]); break;
    default: out[i] = xb::fm::div(a[i], b[i]); break;
  }
}
}  // namespace

extern "C" int xgpu_selftest_fastmath(xgpu_ctx *ctx, int which, int n, const double *h_a, const double *h_b, double *h_out) {
  if (!ctx || n <= 0 || !h_a || !h_out || which < 0 || which > 3 || (which == 2 && !h_b)) return 1;
  double *d = nullptr;
  if (cudaSetDevice(ctx->device) != cudaSuccess || cudaMalloc((void **)&d, 3 * (size_t)n * sizeof(double)) != cudaSuccess)
    return xg_fail(ctx, 11, "device allocation failed in selftest");
  cudaMemcpyAsync(d, h_a, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (h_b) cudaMemcpyAsync(d + n, h_b, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  fastmath_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(which, n, d, d + n, d + 2 * (size_t)n);
  ++ctx->launches;
  cudaMemcpyAsync(h_out, d + 2 * (size_t)n, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  const cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  return e == cudaSuccess ? 0 : xg_fail(ctx, 100 + (int)e, cudaGetErrorString(e));
}
