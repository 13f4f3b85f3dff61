// xyce_b200 -- assembly kernels (sm_100a).  HBM-bandwidth bound: per destination the kernel
// reads 4 bytes of map + 8 bytes per contribution per plane and writes 8 bytes per plane.
#include "assembly.cuh"

namespace xb {
namespace {

constexpr int kMaxPlanes = 4;
struct PlaneSet {
  const double *in[kMaxPlanes];
  double *out[kMaxPlanes];
};

template <int NP>
__global__ void __launch_bounds__(256) gather_short_kernel(GatherMapDev m, PlaneSet ps, bool accumulate) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= m.ndst) return;
  const int64_t b = m.ptr[d], e = m.ptr[d + 1];
  if (e - b > kLongThreshold) return;   // handled by gather_long_kernel
  double acc[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) acc[p] = accumulate ? ps.out[p][d] : 0.0;
  for (int64_t k = b; k < e; ++k) {
    const int32_t s = __ldg(m.src + k);
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[p] += __ldg(ps.in[p] + s);
  }
#pragma unroll
  for (int p = 0; p < NP; ++p) ps.out[p][d] = acc[p];
}

// Long destinations (supply rails): stage 1, one block per chunk of kChunk sources -- every thread
// sums its strided subsequence in index order, then a fixed-shape shared-memory tree combines the
// 256 partials; stage 2, one thread per long destination adds its chunk partials in chunk order.
// Fixed shapes and orders => bitwise reproducible, no atomics.
template <int NP>
__global__ void __launch_bounds__(256) gather_chunk_kernel(GatherMapDev m, PlaneSet ps) {
  __shared__ double sh[NP][256];
  const int c = blockIdx.x;
  const int d = m.long_dst[m.chunk_dst_slot[c]];
  const int64_t b = m.chunk_begin[c];
  const int64_t e = min(b + (int64_t)kChunk, m.ptr[d + 1]);
  double acc[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) acc[p] = 0.0;
  for (int64_t k = b + threadIdx.x; k < e; k += 256) {
    const int32_t s = __ldg(m.src + k);
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[p] += __ldg(ps.in[p] + s);
  }
#pragma unroll
  for (int p = 0; p < NP; ++p) sh[p][threadIdx.x] = acc[p];
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) {
#pragma unroll
      for (int p = 0; p < NP; ++p) sh[p][threadIdx.x] += sh[p][threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x < NP) m.partials[(size_t)threadIdx.x * m.nchunks + c] = sh[threadIdx.x][0];
}

template <int NP>
__global__ void __launch_bounds__(256) gather_long_finish_kernel(GatherMapDev m, PlaneSet ps, bool accumulate) {
  __shared__ double sh[NP][256];
  const int j = blockIdx.x;
  const int d = m.long_dst[j];
  double acc[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) acc[p] = 0.0;
  for (int c = m.long_chunk_ptr[j] + threadIdx.x; c < m.long_chunk_ptr[j + 1]; c += 256) {
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[p] += m.partials[(size_t)p * m.nchunks + c];
  }
#pragma unroll
  for (int p = 0; p < NP; ++p) sh[p][threadIdx.x] = acc[p];
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) {
#pragma unroll
      for (int p = 0; p < NP; ++p) sh[p][threadIdx.x] += sh[p][threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x < NP) ps.out[threadIdx.x][d] = (accumulate ? ps.out[threadIdx.x][d] : 0.0) + sh[threadIdx.x][0];
}

// ---- fused vector + matrix assembly: the same three stages, one launch each for both maps ----
// (threads / blocks beyond the vector map's range work on the matrix map)
template <int NP>
__device__ __forceinline__ void short_body(const GatherMapDev &m, const PlaneSet &ps, int d, bool accumulate) {
  const int64_t b = m.ptr[d], e = m.ptr[d + 1];
  if (e - b > kLongThreshold) return;
  double acc[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) acc[p] = accumulate ? ps.out[p][d] : 0.0;
  for (int64_t k = b; k < e; ++k) {
    const int32_t s = __ldg(m.src + k);
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[p] += __ldg(ps.in[p] + s);
  }
#pragma unroll
  for (int p = 0; p < NP; ++p) ps.out[p][d] = acc[p];
}

template <int NP>
__device__ __forceinline__ void chunk_body(const GatherMapDev &m, const PlaneSet &ps, int c, double (*sh)[256]) {
  const int d = m.long_dst[m.chunk_dst_slot[c]];
  const int64_t b = m.chunk_begin[c];
  const int64_t e = min(b + (int64_t)kChunk, m.ptr[d + 1]);
  double acc[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) acc[p] = 0.0;
  for (int64_t k = b + threadIdx.x; k < e; k += 256) {
    const int32_t s = __ldg(m.src + k);
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[p] += __ldg(ps.in[p] + s);
  }
#pragma unroll
  for (int p = 0; p < NP; ++p) sh[p][threadIdx.x] = acc[p];
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) {
#pragma unroll
      for (int p = 0; p < NP; ++p) sh[p][threadIdx.x] += sh[p][threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x < NP) m.partials[(size_t)threadIdx.x * m.nchunks + c] = sh[threadIdx.x][0];
}

template <int NP>
__device__ __forceinline__ void finish_body(const GatherMapDev &m, const PlaneSet &ps, int j, bool accumulate, double (*sh)[256]) {
  const int d = m.long_dst[j];
  double acc[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) acc[p] = 0.0;
  for (int c = m.long_chunk_ptr[j] + threadIdx.x; c < m.long_chunk_ptr[j + 1]; c += 256) {
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[p] += m.partials[(size_t)p * m.nchunks + c];
  }
#pragma unroll
  for (int p = 0; p < NP; ++p) sh[p][threadIdx.x] = acc[p];
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) {
#pragma unroll
      for (int p = 0; p < NP; ++p) sh[p][threadIdx.x] += sh[p][threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x < NP) ps.out[threadIdx.x][d] = (accumulate ? ps.out[threadIdx.x][d] : 0.0) + sh[threadIdx.x][0];
}

__global__ void __launch_bounds__(256) fused_short_kernel(GatherMapDev mv, PlaneSet pv, GatherMapDev mm, PlaneSet pm,
                                                          int vec_blocks, bool accumulate) {
  if ((int)blockIdx.x < vec_blocks) {
    const int d = blockIdx.x * 256 + threadIdx.x;
    if (d < mv.ndst) short_body<4>(mv, pv, d, accumulate);
  } else {
    const int d = (blockIdx.x - vec_blocks) * 256 + threadIdx.x;
    if (d < mm.ndst) short_body<2>(mm, pm, d, accumulate);
  }
}
__global__ void __launch_bounds__(256) fused_chunk_kernel(GatherMapDev mv, PlaneSet pv, GatherMapDev mm, PlaneSet pm) {
  __shared__ double sh[4][256];
  if ((int)blockIdx.x < mv.nchunks) chunk_body<4>(mv, pv, blockIdx.x, sh);
  else chunk_body<2>(mm, pm, blockIdx.x - mv.nchunks, sh);
}
__global__ void __launch_bounds__(256) fused_finish_kernel(GatherMapDev mv, PlaneSet pv, GatherMapDev mm, PlaneSet pm, bool accumulate) {
  __shared__ double sh[4][256];
  if ((int)blockIdx.x < mv.nlong) finish_body<4>(mv, pv, blockIdx.x, accumulate, sh);
  else finish_body<2>(mm, pm, blockIdx.x - mv.nlong, accumulate, sh);
}

__global__ void __launch_bounds__(256) linear_combo_kernel(int64_t nnz, double a, const double *__restrict__ A,
                                                           double b, const double *__restrict__ B,
                                                           double *__restrict__ J) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nnz) J[k] = a * A[k] + b * B[k];
}

__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <int NP>
void launch_np(const GatherMapDev &m, const PlaneSet &ps, bool accumulate, cudaStream_t stream) {
  if (m.ndst > 0) gather_short_kernel<NP><<<(m.ndst + 255) / 256, 256, 0, stream>>>(m, ps, accumulate);
  if (m.nlong > 0) {
    gather_chunk_kernel<NP><<<m.nchunks, 256, 0, stream>>>(m, ps);
    gather_long_finish_kernel<NP><<<m.nlong, 256, 0, stream>>>(m, ps, accumulate);
  }
}

}  // namespace

void launch_gather(const GatherMapDev &m, int nplanes, const double *const *planes, int64_t, double *const *dst,
                   bool accumulate, cudaStream_t stream) {
  PlaneSet ps{};
  for (int p = 0; p < nplanes; ++p) { ps.in[p] = planes[p]; ps.out[p] = dst[p]; }
  switch (nplanes) {
    case 1: launch_np<1>(m, ps, accumulate, stream); break;
    case 2: launch_np<2>(m, ps, accumulate, stream); break;
    case 4: launch_np<4>(m, ps, accumulate, stream); break;
    default: break;
  }
}

int launch_gather_fused(const GatherMapDev &mv, const double *const *vplanes, double *const *vdst, const GatherMapDev &mm,
                        const double *const *mplanes, double *const *mdst, bool accumulate, cudaStream_t stream) {
  PlaneSet pv{}, pm{};
  for (int p = 0; p < 4; ++p) { pv.in[p] = vplanes[p]; pv.out[p] = vdst[p]; }
  for (int p = 0; p < 2; ++p) { pm.in[p] = mplanes[p]; pm.out[p] = mdst[p]; }
  int launches = 0;
  const int vb = (mv.ndst + 255) / 256, mb = (mm.ndst + 255) / 256;
  if (vb + mb > 0) { fused_short_kernel<<<vb + mb, 256, 0, stream>>>(mv, pv, mm, pm, vb, accumulate); ++launches; }
  if (mv.nchunks + mm.nchunks > 0) {
    fused_chunk_kernel<<<mv.nchunks + mm.nchunks, 256, 0, stream>>>(mv, pv, mm, pm);
    fused_finish_kernel<<<mv.nlong + mm.nlong, 256, 0, stream>>>(mv, pv, mm, pm, accumulate);
    launches += 2;
  }
  return launches;
}

cudaError_t measure_fp64_peak(cudaStream_t stream, double *tflops) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 8, threads = 256, iters = 4096;
  double *out = nullptr;
  cudaError_t e = cudaMalloc((void **)&out, (size_t)blocks * threads * sizeof(double));
  if (e != cudaSuccess) return e;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a, stream);
    fp64_peak_kernel<<<blocks, threads, 0, stream>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(b, stream);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  e = cudaGetLastError();
  cudaEventDestroy(a); cudaEventDestroy(b);
  cudaFree(out);
  *tflops = 2.0 * 8.0 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
  return e;
}

void launch_linear_combo(int64_t nnz, double a, const double *A, double b, const double *B, double *J,
                         cudaStream_t stream) {
  if (nnz > 0) linear_combo_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, stream>>>(nnz, a, A, b, B, J);
}

}  // namespace xb
