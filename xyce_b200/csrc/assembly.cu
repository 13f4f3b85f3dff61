// xyce_b200 -- assembly kernels (sm_100a).  HBM-bandwidth bound: per destination the kernel
// reads 4 bytes of map + 8 bytes per contribution per plane and writes 8 bytes per plane.
#include "pdl.cuh"
#include "assembly.cuh"

namespace xb {
namespace {

constexpr int kMaxPlanes = 4;
struct PlaneSet {
  const double *in[kMaxPlanes];
  double *out[kMaxPlanes];
};

// ---- per-destination and per-chunk bodies ----
// One thread sums U destinations (d0, d0 + 256, ...): all ELL slots are fetched first, then all plane
// elements, so U * ell_w * NP independent loads are in flight per thread instead of a dependent chain per
// destination; the additions then run in list order per destination.
template <int NP> struct ShortUnroll { static constexpr int U = NP >= 4 ? 2 : 4; };

template <int NP>
__device__ __forceinline__ void short_body(const GatherMapDev &m, const double *const (&in)[NP], double *const (&out)[NP],
                                           int d0, bool accumulate) {
  constexpr int U = ShortUnroll<NP>::U;
  const int hi = m.d_hi < 0 ? m.ndst : m.d_hi;
  int32_t s[U][kEllMax];
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const int d = d0 + j * 256;
#pragma unroll
    for (int k = 0; k < kEllMax; ++k) s[j][k] = (k < m.ell_w && d < hi) ? __ldg(m.ell + (size_t)k * m.ndst + d) : -1;
  }
  double val[U][kEllMax][NP];
#pragma unroll
  for (int j = 0; j < U; ++j)
#pragma unroll
    for (int k = 0; k < kEllMax; ++k)
#pragma unroll
      for (int p = 0; p < NP; ++p) val[j][k][p] = (s[j][k] >= 0) ? __ldg(in[p] + s[j][k]) : 0.0;
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const int d = d0 + j * 256;
    if (d >= hi || s[j][0] == kEllLong) continue;          // out of the window / handled by the chunk blocks
    const bool tail = (s[j][1] == kEllTail) || (s[j][kEllMax - 1] == kEllTail);      // the flag sits in slot ell_w - 1
    double acc[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[p] = accumulate ? out[p][d] : 0.0;
#pragma unroll
    for (int k = 0; k < kEllMax; ++k)
      if (s[j][k] >= 0) {
#pragma unroll
        for (int p = 0; p < NP; ++p) acc[p] += val[j][k][p];
      }
    if (tail) {                             // more than ell_w sources: the rest in list order
      const int64_t b = m.ptr[d] + m.ell_w - 1, e = m.ptr[d + 1];
      for (int64_t k = b; k < e; ++k) {
        const int32_t t = __ldg(m.src + k);
#pragma unroll
        for (int p = 0; p < NP; ++p) acc[p] += __ldg(in[p] + t);
      }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) out[p][d] = acc[p];
  }
}

template <int NP>
__device__ __forceinline__ void chunk_body(const GatherMapDev &m, const PlaneSet &ps, int c, double (*sh)[256]) {
  const int d = m.long_dst[m.chunk_dst_slot[c]];
  const int64_t b = m.chunk_begin[c];
  const int64_t e = min(b + (int64_t)kChunk, m.ptr[d + 1]);
  // the thread's kChunk / 256 strided sources: indices first, then all plane elements (independent loads),
  // then the additions in index order
  constexpr int R = kChunk / 256;
  int32_t si[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { const int64_t k = b + threadIdx.x + r * 256; si[r] = (k < e) ? __ldg(m.src + k) : -1; }
  double v[R][NP];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int p = 0; p < NP; ++p) v[r][p] = (si[r] >= 0) ? __ldg(ps.in[p] + si[r]) : 0.0;
  double acc[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) acc[p] = 0.0;
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (si[r] >= 0) {
#pragma unroll
      for (int p = 0; p < NP; ++p) acc[p] += v[r][p];
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
  // thread 0 publishes ALL NP partials itself: its __threadfence() in chunk_then_finish then orders every one of
  // them before the ticket (stores by other threads would not be covered by that fence)
  if (threadIdx.x == 0) {
#pragma unroll
    for (int p = 0; p < NP; ++p) m.partials[(size_t)p * m.nchunks + c] = sh[p][0];
  }
}

// ---- one-launch assembly -------------------------------------------------------------------------------
// Grid = [chunk blocks of map A][chunk blocks of map B][short blocks of A][short blocks of B].  A chunk block
// publishes its partials (__threadfence), takes a ticket on its destination's counter, and the block that
// draws the last ticket runs the finish stage for that destination -- the same fixed-order sum of the chunk
// partials whichever block ends up doing it, so the result stays bitwise reproducible.
template <int NP>
__device__ __forceinline__ void chunk_then_finish(const GatherMapDev &m, const PlaneSet &ps, int c, bool accumulate,
                                                  double (*sh)[256]) {
  __shared__ int last;
  chunk_body<NP>(m, ps, c, sh);
  const int j = m.chunk_dst_slot[c];
  if (threadIdx.x == 0) {
    __threadfence();
    const int need = m.long_chunk_ptr[j + 1] - m.long_chunk_ptr[j];
    last = (atomicAdd(m.done + j, 1) == need - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x == 0) m.done[j] = 0;          // ready for the next launch
  const int d = m.long_dst[j];
  double acc[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) acc[p] = 0.0;
  for (int cc = m.long_chunk_ptr[j] + threadIdx.x; cc < m.long_chunk_ptr[j + 1]; cc += 256) {
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[p] += __ldcg(m.partials + (size_t)p * m.nchunks + cc);
  }
  __syncthreads();
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
  if (threadIdx.x == 0) {
#pragma unroll
    for (int p = 0; p < NP; ++p) ps.out[p][d] = (accumulate ? ps.out[p][d] : 0.0) + sh[p][0];      // static plane indices
  }
}

// resident blocks per SM: 3 (85 registers) measured 1.3 % faster on the config-2 step than 2 (98 registers, the compiler's
// choice); 4 (64 registers) spills and is slower (72.0 / 72.9 / 73.8 us per step)
#ifndef XB_ASM_MINBLOCKS
#define XB_ASM_MINBLOCKS 3
#endif
template <int NPA, int NPB>
__global__ void __launch_bounds__(256, XB_ASM_MINBLOCKS) assemble_kernel(GatherMapDev ma, PlaneSet pa, GatherMapDev mb, PlaneSet pb,
                                                       int short_a, bool accumulate) {
  __shared__ double sh[(NPA > NPB ? NPA : NPB)][256];
  // launched with programmatic stream serialization: the grid may start while the evaluation kernel that
  // produces the planes is still draining; wait here until that kernel has completed and flushed
  asm volatile("griddepcontrol.wait;" ::: "memory");
  int b = blockIdx.x;
  const int ca = ma.with_chunks ? ma.nchunks : 0, cb = mb.with_chunks ? mb.nchunks : 0;
  if (b < ca) { chunk_then_finish<NPA>(ma, pa, b, accumulate, sh); return; }
  b -= ca;
  if (b < cb) { chunk_then_finish<NPB>(mb, pb, b, accumulate, sh); return; }
  b -= cb;
  // plane pointers into registers (static indices only: no local-memory copy of the parameter structs)
  if (b < short_a) {
    const double *in[NPA]; double *out[NPA];
#pragma unroll
    for (int p = 0; p < NPA; ++p) { in[p] = pa.in[p]; out[p] = pa.out[p]; }
    short_body<NPA>(ma, in, out, ma.d_lo + b * (256 * ShortUnroll<NPA>::U) + threadIdx.x, accumulate);
  } else {
    const double *in[NPB]; double *out[NPB];
#pragma unroll
    for (int p = 0; p < NPB; ++p) { in[p] = pb.in[p]; out[p] = pb.out[p]; }
    short_body<NPB>(mb, in, out, mb.d_lo + (b - short_a) * (256 * ShortUnroll<NPB>::U) + threadIdx.x, accumulate);
  }
}

__global__ void __launch_bounds__(256) linear_combo_kernel(int64_t nnz, double a, const double *__restrict__ A,
                                                           double b, const double *__restrict__ B,
                                                           double *__restrict__ J) {
  xb::pdl_wait();
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

// 256-thread launch that may overlap the tail of the previous kernel in the stream (programmatic dependent launch);
// the kernel itself waits for its producer with griddepcontrol.wait before touching memory.
template <class... KArgs, class... Args>
void launch_dependent(void (*kernel)(KArgs...), int blocks, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <int NP>
void launch_np(const GatherMapDev &m, const PlaneSet &ps, bool accumulate, cudaStream_t stream) {
  const int per = 256 * ShortUnroll<NP>::U;
  const int sb = ((m.d_hi < 0 ? m.ndst : m.d_hi) - m.d_lo + per - 1) / per;
  const int ch = m.with_chunks ? m.nchunks : 0;
  GatherMapDev none{};
  if (ch + sb > 0) launch_dependent(assemble_kernel<NP, 1>, ch + sb, stream, m, ps, none, PlaneSet{}, sb, accumulate);
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
  const int pv_ = 256 * ShortUnroll<4>::U, pm_ = 256 * ShortUnroll<2>::U;
  const int vb = ((mv.d_hi < 0 ? mv.ndst : mv.d_hi) - mv.d_lo + pv_ - 1) / pv_, mb = ((mm.d_hi < 0 ? mm.ndst : mm.d_hi) - mm.d_lo + pm_ - 1) / pm_;
  const int blocks = (mv.with_chunks ? mv.nchunks : 0) + (mm.with_chunks ? mm.nchunks : 0) + vb + mb;
  if (blocks == 0) return 0;
  launch_dependent(assemble_kernel<4, 2>, blocks, stream, mv, pv, mm, pm, vb, accumulate);
  return 1;
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
  if (nnz > 0) xb::launch_pdl(linear_combo_kernel, dim3((unsigned)((nnz + 255) / 256)), dim3(256), 0, stream, nnz, a, A, b, B, J);
}

}  // namespace xb
