// xyce_b200 -- deterministic, atomic-free assembly of device contributions into the
// DAE vectors (F, Q, dFdxdVp, dQdxdVp) and CSR matrices (dF/dx, dQ/dx).
//
// Replaces the reference's scattered "+=" through raw pointers (Master::loadDAEVectors /
// loadDAEMatrices, e.g. N_DEV_MOSFET_B4.C:10691-10776, :11038-11278; address rule
// N_LAS_EpetraMatrix.C:658-664; offsets from N_TOP_Indexor.C:149-214) by a gather:
// every destination (vector row / CSR nonzero) owns the list of contribution-plane
// elements that land on it, ordered device-type-major, instance-minor -- the reference's
// accumulation order (Core/N_DEV_DeviceMgr.C:4238-4248) -- and sums them in that fixed
// order.  Destinations with a very long list (supply rails) are summed by one block with a
// fixed-shape tree, so results are bitwise reproducible run to run.  The whole assembly is ONE launch:
// chunk blocks come first in the grid; the chunk block that finishes last for its destination (integer
// ticket counter, the only atomic in the path -- it carries no data) adds the chunk partials in chunk order.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

namespace xb {

// CSR-by-destination gather map.  src[] indexes one contribution plane.
struct GatherMapHost {
  std::vector<int64_t> ptr;   // [ndst+1]
  std::vector<int32_t> src;   // [total]
  std::vector<int32_t> long_dst;   // destinations with more than kLongThreshold sources
};

struct GatherMapDev {
  int ndst = 0;
  int nlong = 0;
  int nchunks = 0;          // total chunks over all long destinations
  int64_t total = 0;
  int64_t *ptr = nullptr;
  int32_t *src = nullptr;
  int32_t *long_dst = nullptr;
  // long destinations are cut into chunks of kChunk sources: chunk c covers src[chunk_begin[c] ..
  // chunk_begin[c]+kChunk) clipped to its destination's end; destination j owns chunks
  // [long_chunk_ptr[j], long_chunk_ptr[j+1]).
  int64_t *chunk_begin = nullptr;
  int32_t *chunk_dst_slot = nullptr;   // index into long_dst
  int32_t *long_chunk_ptr = nullptr;   // [nlong+1]
  double *partials = nullptr;          // [4][nchunks]
  // ELL view of the short destinations: ell[k * ndst + d] = k-th source of destination d for k < ell_w (coalesced,
  // all slots loadable at once: one dependent-load level less than ptr -> src -> plane).  -1 = no such source;
  // last slot kEllTail = more sources follow, continue in src[] from ptr[d] + ell_w - 1; slot 0 kEllLong = long
  // destination (left to the chunk blocks).
  int ell_w = 0;
  int32_t *ell = nullptr;
  int32_t *done = nullptr;             // [nlong] chunk blocks finished so far (ticket counter; zero between launches)
  // window of this launch over the short destinations: [d_lo, d_hi) (d_hi < 0 = ndst); with_chunks = 0 leaves the long
  // destinations to another launch (pipelined host path: first the destinations fed by the first half of the instances)
  int d_lo = 0, d_hi = -1, with_chunks = 1;
};

constexpr int kLongThreshold = 96;
constexpr int kChunk = 1024;
constexpr int kEllMax = 4;
constexpr int32_t kEllTail = -2, kEllLong = -3;

// Sum `nplanes` planes through one map.  dst[p][d] = (accumulate ? dst[p][d] : 0) + sum_k plane[p][src[k]].
// Long destinations are skipped by the short kernel and handled by the block kernel.
void launch_gather(const GatherMapDev &m, int nplanes, const double *const *planes, int64_t plane_stride,
                   double *const *dst, bool accumulate, cudaStream_t stream);

// Vector planes (4, map mv) and matrix planes (2, map mm) in the same launch.  Same sums in the same
// order as two launch_gather calls; returns the number of kernel launches.
// (windows: set d_lo / d_hi / with_chunks in copies of the maps)
int launch_gather_fused(const GatherMapDev &mv, const double *const *vplanes, double *const *vdst, const GatherMapDev &mm,
                        const double *const *mplanes, double *const *mdst, bool accumulate, cudaStream_t stream);

// J = qscalar * dQdx + fscalar * dFdx over one CSR pattern (N_LAS_EpetraMatrix.C:629-648 linearCombo
// as used by OneStep::obtainJacobian, N_TIA_OneStep.C:490-495).
void launch_linear_combo(int64_t nnz, double a, const double *A, double b, const double *B, double *J,
                         cudaStream_t stream);

// Dependent-chain DFMA microbenchmark (8 independent chains per thread); returns TFLOP/s.
cudaError_t measure_fp64_peak(cudaStream_t stream, double *tflops);

}  // namespace xb
