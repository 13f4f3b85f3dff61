#include "pdl.cuh"
#include "vecops.cuh"
namespace xb {
namespace vec {
namespace {
__global__ void fill_k(double *d, double v, int n) {
  xb::pdl_wait(); int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) d[i] = v; }
__global__ void axpby_k(double *dst, double a, const double *x, double b, const double *y, int n) {
  xb::pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = a * x[i] + b * y[i];
}
__global__ void solw_k(double *dst, double rel, double ab, const double *a, const double *b, int n) {
  xb::pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = rel * fmax(fabs(a[i]), fabs(b[i])) + ab;
}
__global__ void absw_k(double *dst, double rel, double ab, const double *a, int n) {
  xb::pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = rel * fabs(a[i]) + ab;
}
template <int MODE>
__device__ __forceinline__ double term(const double *x, const double *w, int i) {
  if (MODE == kSumSq) return x[i] * x[i];
  if (MODE == kMaxAbs) return fabs(x[i]);
  if (MODE == kWMaxAbs) return fabs(x[i] / w[i]);
  const double t = x[i] / w[i];
  return t * t;
}
template <int MODE>
__device__ __forceinline__ double comb(double a, double b) {
  if (MODE == kMaxAbs || MODE == kWMaxAbs) return (a != a || b != b) ? (a + b) : fmax(a, b);   // NaN propagates
  return a + b;
}
template <int MODE>
__global__ void __launch_bounds__(256) reduce_k(const double *x, const double *w, int n, double *out) {
  xb::pdl_wait();
  __shared__ double sh[256];
  double acc = 0.0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) acc = comb<MODE>(acc, term<MODE>(x, w, i));
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = comb<MODE>(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}
template <int MODE>
__global__ void __launch_bounds__(256) reduce_final_k(double *part, int m) {
  xb::pdl_wait();
  __shared__ double sh[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < m; i += 256) acc = comb<MODE>(acc, part[i]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = comb<MODE>(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) part[0] = sh[0];
}
template <int MODE>
void reduce_dev_t(const double *x, const double *w, int n, double *scratch, cudaStream_t s) {
  const int blocks = n < 256 * 1024 ? (n + 255) / 256 : 1024;
  xb::launch_pdl(reduce_k<MODE>, dim3(blocks > 0 ? blocks : 1), dim3(256), 0, s, x, w, n, scratch);
  xb::launch_pdl(reduce_final_k<MODE>, dim3(1), dim3(256), 0, s, scratch, blocks > 0 ? blocks : 1);
}
template <int MODE>
double reduce_t(const double *x, const double *w, int n, double *scratch, cudaStream_t s) {
  reduce_dev_t<MODE>(x, w, n, scratch, s);
  double h = 0.0;
  cudaMemcpyAsync(&h, scratch, sizeof(double), cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
  return h;
}
constexpr int kResBlocks = 1024;
__global__ void __launch_bounds__(256) residual_norms_k(ResidualArgs a, double *part) {
  xb::pdl_wait();
  __shared__ double sh[3][256];
  __shared__ int shf;
  if (threadIdx.x == 0) shf = 1;
  __syncthreads();
  double s2 = 0.0, mx = 0.0, wm = 0.0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < a.n; i += gridDim.x * 256) {
    double r;
    if (a.form == 0) {             // OneStep::obtainResidual
      r = 1.0 * a.q[i] + -1.0 * a.qh0[i];
      const double t = a.fs * a.f[i] + (-a.fs) * a.b[i];
      r = a.inv_h * r + 1.0 * t;
      if (a.order2) r = 1.0 * r + 0.5 * a.qh2[i];
      r = -1.0 * r + 0.0 * r;
      if (a.limiter) { r = 1.0 * r + a.qlim_coef * a.qlim[i]; r = 1.0 * r + a.fs * a.flim[i]; }
    } else if (a.form == 1) {      // Gear12::obtainResidual
      r = a.a0 * a.q[i] + a.a1 * a.qh0[i];
      if (a.order2) r = 1.0 * r + a.a2 * a.qh1[i];
      const double t = 1.0 * a.f[i] + -1.0 * a.b[i];
      r = a.inv_h * r + 1.0 * t;
      r = -1.0 * r + 0.0 * r;
      if (a.limiter) { r = 1.0 * r + a.qlim_coef * a.qlim[i]; r = 1.0 * r + 1.0 * a.flim[i]; }
    } else {                       // NoTimeIntegration::obtainResidual
      r = 1.0 * a.f[i] + -1.0 * a.b[i];
      r = -1.0 * r + 0.0 * r;
      if (a.limiter) r = 1.0 * r + 1.0 * a.flim[i];
    }
    a.rhs[i] = r;
    s2 = comb<kSumSq>(s2, r * r);
    mx = comb<kMaxAbs>(mx, fabs(r));
    wm = comb<kWMaxAbs>(wm, fabs(a.dx[i] / a.w[i]));
  }
  int ok = 1;
  for (int k = 0; k < a.nflag_arrays; ++k)
    for (int i = blockIdx.x * 256 + threadIdx.x; i < a.flag_n[k]; i += gridDim.x * 256) if (a.flags[k][i] == 0) ok = 0;
  if (!ok) shf = 0;        // every writer stores the same value
  sh[0][threadIdx.x] = s2; sh[1][threadIdx.x] = mx; sh[2][threadIdx.x] = wm;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sh[0][threadIdx.x] = comb<kSumSq>(sh[0][threadIdx.x], sh[0][threadIdx.x + s]);
      sh[1][threadIdx.x] = comb<kMaxAbs>(sh[1][threadIdx.x], sh[1][threadIdx.x + s]);
      sh[2][threadIdx.x] = comb<kWMaxAbs>(sh[2][threadIdx.x], sh[2][threadIdx.x + s]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[blockIdx.x] = sh[0][0]; part[kResBlocks + blockIdx.x] = sh[1][0]; part[2 * kResBlocks + blockIdx.x] = sh[2][0];
    part[3 * kResBlocks + blockIdx.x] = (double)shf;
  }
}
__global__ void __launch_bounds__(256) residual_norms_final_k(const double *part, int m, double *out4) {
  xb::pdl_wait();
  __shared__ double sh[4][256];
  double s2 = 0.0, mx = 0.0, wm = 0.0, ok = 1.0;
  for (int i = threadIdx.x; i < m; i += 256) {
    s2 = comb<kSumSq>(s2, part[i]);
    mx = comb<kMaxAbs>(mx, part[kResBlocks + i]);
    wm = comb<kWMaxAbs>(wm, part[2 * kResBlocks + i]);
    ok = fmin(ok, part[3 * kResBlocks + i]);
  }
  sh[0][threadIdx.x] = s2; sh[1][threadIdx.x] = mx; sh[2][threadIdx.x] = wm; sh[3][threadIdx.x] = ok;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sh[0][threadIdx.x] = comb<kSumSq>(sh[0][threadIdx.x], sh[0][threadIdx.x + s]);
      sh[1][threadIdx.x] = comb<kMaxAbs>(sh[1][threadIdx.x], sh[1][threadIdx.x + s]);
      sh[2][threadIdx.x] = comb<kWMaxAbs>(sh[2][threadIdx.x], sh[2][threadIdx.x + s]);
      sh[3][threadIdx.x] = fmin(sh[3][threadIdx.x], sh[3][threadIdx.x + s]);
    }
    __syncthreads();
  }
  if (threadIdx.x < 4) out4[threadIdx.x] = sh[threadIdx.x][0];
}
__global__ void spmv_add_k(int nrows, const int *rows, const int *ptr, const int *col, const double *val,
                           const double *x, double *y) {
  xb::pdl_wait();
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  double acc = 0.0;
  for (int k = ptr[r]; k < ptr[r + 1]; ++k) acc += val[k] * x[col[k]];
  y[rows[r]] += acc;
}
__global__ void scatter_add_k(int n, const int *pos, const double *v, double *vals) {
  xb::pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) vals[pos[i]] += v[i];
}
inline int nb(int n) { return (n + 255) / 256; }
}  // namespace

void fill(double *d, double v, int n, cudaStream_t s) { if (n > 0) xb::launch_pdl(fill_k, dim3(nb(n)), dim3(256), 0, s, d, v, n); }
void axpby(double *dst, double a, const double *x, double b, const double *y, int n, cudaStream_t s) {
  if (n > 0) xb::launch_pdl(axpby_k, dim3(nb(n)), dim3(256), 0, s, dst, a, x, b, y, n);
}
void sol_weights(double *dst, double rel, double ab, const double *a, const double *b, int n, cudaStream_t s) {
  if (n > 0) xb::launch_pdl(solw_k, dim3(nb(n)), dim3(256), 0, s, dst, rel, ab, a, b, n);
}
void abs_weights(double *dst, double rel, double ab, const double *a, int n, cudaStream_t s) {
  if (n > 0) xb::launch_pdl(absw_k, dim3(nb(n)), dim3(256), 0, s, dst, rel, ab, a, n);
}
double reduce(Reduce mode, const double *x, const double *w, int n, double *scratch, cudaStream_t s) {
  switch (mode) {
    case kSumSq: return reduce_t<kSumSq>(x, w, n, scratch, s);
    case kMaxAbs: return reduce_t<kMaxAbs>(x, w, n, scratch, s);
    case kWMaxAbs: return reduce_t<kWMaxAbs>(x, w, n, scratch, s);
    default: return reduce_t<kWSumSq>(x, w, n, scratch, s);
  }
}
void reduce_dev(Reduce mode, const double *x, const double *w, int n, double *scratch, cudaStream_t s) {
  switch (mode) {
    case kSumSq: reduce_dev_t<kSumSq>(x, w, n, scratch, s); break;
    case kMaxAbs: reduce_dev_t<kMaxAbs>(x, w, n, scratch, s); break;
    case kWMaxAbs: reduce_dev_t<kWMaxAbs>(x, w, n, scratch, s); break;
    default: reduce_dev_t<kWSumSq>(x, w, n, scratch, s); break;
  }
}
void residual_norms(const ResidualArgs &a, double *scratch, double *out4, cudaStream_t s) {
  int work = a.n;
  for (int k = 0; k < a.nflag_arrays; ++k) work = work > a.flag_n[k] ? work : a.flag_n[k];
  int blocks = work < 256 * kResBlocks ? (work + 255) / 256 : kResBlocks;
  if (blocks < 1) blocks = 1;
  xb::launch_pdl(residual_norms_k, dim3(blocks), dim3(256), 0, s, a, scratch);
  xb::launch_pdl(residual_norms_final_k, dim3(1), dim3(256), 0, s, scratch, blocks, out4);
}
void spmv_add(int nrows, const int *rows, const int *ptr, const int *col, const double *val, const double *x,
              double *y, cudaStream_t s) {
  if (nrows > 0) xb::launch_pdl(spmv_add_k, dim3(nb(nrows)), dim3(256), 0, s, nrows, rows, ptr, col, val, x, y);
}
void scatter_add(int n, const int *pos, const double *v, double *vals, cudaStream_t s) {
  if (n > 0) xb::launch_pdl(scatter_add_k, dim3(nb(n)), dim3(256), 0, s, n, pos, v, vals);
}
}  // namespace vec
}  // namespace xb
