// xyce_b200 -- sparse LU on the GPU (sm_100a): numeric refactorization and triangular solves on a
// fixed pattern / pivot sequence (klu_refactor + klu_solve semantics; reference call sites
// Amesos_Klu::NumericFactorization / Solve in N_LAS_AmesosSolver.C:363, :396).
//
// Parallelism comes from the block triangular form: every diagonal block is independent during
// refactorization, and blocks of one dependency level are independent during the solve.  One warp
// owns one block: lanes run over the entries of a column (warp-synchronous, no atomics), columns are
// processed in pivot order -- the Gilbert-Peierls left-looking update without the pivot search.
// The dense column work vector lives in a per-block slice of a global array; for blocks of at most
// kSmemRows rows it is staged in shared memory instead.
// HBM traffic per refactor: 8*nnz(A) read + 12*nnz(L+U) read/write (SURVEY.md 8d unit U3).
#include <cuda_runtime.h>
#include "lu.h"

namespace xb {
namespace lu {

namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kSmemRows = 512;     // rows staged in shared memory per warp (4 KB)

__device__ __forceinline__ bool bad_pivot(double p) { return p == 0.0 || !(fabs(p) <= 1.7976931348623157e308); }

__global__ void __launch_bounds__(32 * kWarpsPerCta) lu_refactor_kernel(LuDev d, const double *__restrict__ A) {
  __shared__ double sx[kWarpsPerCta][kSmemRows];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kWarpsPerCta + warp;
  if (b >= d.nblocks) return;
  const int k0 = d.block_ptr[b], k1 = d.block_ptr[b + 1];
  const int nb = k1 - k0;
  if (nb == 1) {                               // 1x1 block: the pivot is the matrix entry itself
    if (lane == 0) {
      double p = 0.0;
      for (int q = d.acol_ptr[k0]; q < d.acol_ptr[k0 + 1]; ++q) p += A[d.acol_src[q]];
      d.Ux[d.Up[k0 + 1] - 1] = p;
      if (bad_pivot(p)) *d.status = 1;
    }
    return;
  }
  double *x = (nb <= kSmemRows) ? (&sx[warp][0] - k0) : d.work;   // x[k0..k1) addressed by position
  for (int k = k0; k < k1; ++k) {
    const int ub = d.Up[k], ue = d.Up[k + 1] - 1;       // off-diagonal U entries [ub, ue), pivot at ue
    const int lb = d.Lp[k], le = d.Lp[k + 1];
    // clear the column pattern, then scatter A(:,k)
    for (int q = ub + lane; q < ue; q += 32) x[d.Ui[q]] = 0.0;
    for (int q = lb + lane; q < le; q += 32) x[d.Li[q]] = 0.0;
    if (lane == 0) x[k] = 0.0;
    __syncwarp();
    for (int q = d.acol_ptr[k] + lane; q < d.acol_ptr[k + 1]; q += 32) x[d.acol_row[q]] = A[d.acol_src[q]];
    __syncwarp();
    // left-looking updates in ascending pivot order
    for (int q = ub; q < ue; ++q) {
      const int i = d.Ui[q];
      const double u = x[i];
      if (lane == 0) d.Ux[q] = u;
      for (int t = d.Lp[i] + lane; t < d.Lp[i + 1]; t += 32) x[d.Li[t]] -= d.Lx[t] * u;
      __syncwarp();
    }
    const double pivot = x[k];
    if (lane == 0) {
      d.Ux[ue] = pivot;
      if (bad_pivot(pivot)) *d.status = 1;
    }
    for (int q = lb + lane; q < le; q += 32) d.Lx[q] = x[d.Li[q]] / pivot;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) lu_permute_rhs_kernel(LuDev d, const double *__restrict__ rhs) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < d.n) d.work[t] = rhs[d.row_perm[t]];
}

// Off-diagonal pull of one level: y[r] -= sum_q A[offr_src[q]] * y[offr_col[q]].  Short rows: one warp per
// row (lane-strided partial sums, shuffle tree).  Long rows (supply rails with ~1e6 entries): one block per
// 4096-entry chunk with a fixed-shape tree, then one block per row over the chunk partials.  Fixed shapes
// and orders, no atomics.
__global__ void __launch_bounds__(256) lu_pull_short_kernel(LuDev d, const double *__restrict__ A, int first, int count) {
  const int w = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= count) return;
  const int r = d.pull_short_rows[first + w];
  double acc = 0.0;
  for (int q = d.offr_ptr[r] + lane; q < d.offr_ptr[r + 1]; q += 32) acc += A[d.offr_src[q]] * d.work[d.offr_col[q]];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) d.work[r] -= acc;
}

__device__ __forceinline__ double block_tree_sum(double v, double *sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  return sh[0];
}

__global__ void __launch_bounds__(256) lu_pull_chunk_kernel(LuDev d, const double *__restrict__ A, int first_chunk) {
  __shared__ double sh[256];
  const int c = first_chunk + blockIdx.x;
  const int r = d.pull_long_rows[d.pull_chunk_row_slot[c]];
  const int b = d.pull_chunk_begin[c];
  const int e = min(b + 4096, d.offr_ptr[r + 1]);
  double acc = 0.0;
  for (int q = b + threadIdx.x; q < e; q += 256) acc += A[d.offr_src[q]] * d.work[d.offr_col[q]];
  const double t = block_tree_sum(acc, sh);
  if (threadIdx.x == 0) d.pull_partials[c] = t;
}

__global__ void __launch_bounds__(256) lu_pull_finish_kernel(LuDev d, int first_slot) {
  __shared__ double sh[256];
  const int slot = first_slot + blockIdx.x;
  const int r = d.pull_long_rows[slot];
  double acc = 0.0;
  for (int c = d.pull_long_chunk_ptr[slot] + threadIdx.x; c < d.pull_long_chunk_ptr[slot + 1]; c += 256) acc += d.pull_partials[c];
  const double t = block_tree_sum(acc, sh);
  if (threadIdx.x == 0) d.work[r] -= t;
}

// One level of the block back-substitution: every block of the level pulls the contributions of the
// already-solved later blocks into its right-hand side, then does L and U solves inside the block.
__global__ void __launch_bounds__(32 * kWarpsPerCta) lu_solve_level_kernel(LuDev d, const double *__restrict__ A,
                                                                          int first, int count,
                                                                          double *__restrict__ xout) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.x * kWarpsPerCta + warp;
  if (idx >= count) return;
  const int b = d.level_blocks[first + idx];
  const int k0 = d.block_ptr[b], k1 = d.block_ptr[b + 1];
  double *y = d.work;
  // forward substitution with unit-lower L
  for (int k = k0; k < k1; ++k) {
    const double yk = y[k];
    for (int q = d.Lp[k] + lane; q < d.Lp[k + 1]; q += 32) y[d.Li[q]] -= d.Lx[q] * yk;
    __syncwarp();
  }
  // backward substitution with U (pivot stored last in each column)
  for (int k = k1 - 1; k >= k0; --k) {
    const int ue = d.Up[k + 1] - 1;
    const double yk = y[k] / d.Ux[ue];
    __syncwarp();
    if (lane == 0) { y[k] = yk; xout[d.col_perm[k]] = yk; }
    for (int q = d.Up[k] + lane; q < ue; q += 32) y[d.Ui[q]] -= d.Ux[q] * yk;
    __syncwarp();
  }
}

template <class T>
cudaError_t up(T **dst, const std::vector<T> &v) {
  cudaError_t e = cudaMalloc((void **)dst, (v.empty() ? 1 : v.size()) * sizeof(T));
  if (e != cudaSuccess) return e;
  if (!v.empty()) e = cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return e;
}

}  // namespace

void free_plan(LuDev &d) {
  cudaFree(d.row_perm); cudaFree(d.col_perm); cudaFree(d.block_ptr); cudaFree(d.Lp); cudaFree(d.Li);
  cudaFree(d.Up); cudaFree(d.Ui); cudaFree(d.Lx); cudaFree(d.Ux); cudaFree(d.acol_ptr); cudaFree(d.acol_row);
  cudaFree(d.acol_src); cudaFree(d.offr_ptr); cudaFree(d.offr_col); cudaFree(d.offr_src);
  cudaFree(d.level_blocks); cudaFree(d.work); cudaFree(d.status);
  cudaFree(d.pull_short_rows); cudaFree(d.pull_long_rows); cudaFree(d.pull_chunk_row_slot); cudaFree(d.pull_chunk_begin);
  cudaFree(d.pull_long_chunk_ptr); cudaFree(d.pull_partials);
  d = LuDev();
}

cudaError_t upload_plan(const LuPlan &p, LuDev &d) {
  free_plan(d);
  d.n = p.n;
  d.nblocks = (int)p.block_ptr.size() - 1;
  d.level_ptr = p.level_ptr;
  d.nlevels = (int)p.level_ptr.size() - 1;
  cudaError_t e;
#define UP(f) if ((e = up(&d.f, p.f)) != cudaSuccess) return e;
  UP(row_perm) UP(col_perm) UP(block_ptr) UP(Lp) UP(Li) UP(Up) UP(Ui) UP(Lx) UP(Ux)
  UP(acol_ptr) UP(acol_row) UP(acol_src) UP(offr_ptr) UP(offr_col) UP(offr_src) UP(level_blocks)
  UP(pull_short_rows) UP(pull_long_rows) UP(pull_chunk_row_slot) UP(pull_chunk_begin) UP(pull_long_chunk_ptr)
#undef UP
  d.pull_short_ptr = p.pull_short_ptr; d.pull_long_ptr = p.pull_long_ptr; d.pull_chunk_ptr = p.pull_chunk_ptr;
  if ((e = cudaMalloc((void **)&d.pull_partials, (p.pull_chunk_begin.size() + 1) * sizeof(double))) != cudaSuccess) return e;
  if ((e = cudaMalloc((void **)&d.work, (size_t)(p.n > 0 ? p.n : 1) * sizeof(double))) != cudaSuccess) return e;
  if ((e = cudaMalloc((void **)&d.status, sizeof(int))) != cudaSuccess) return e;
  return cudaMemset(d.status, 0, sizeof(int));
}

int launch_refactor(const LuDev &d, const double *A, cudaStream_t s) {
  cudaMemsetAsync(d.status, 0, sizeof(int), s);
  const int ctas = (d.nblocks + kWarpsPerCta - 1) / kWarpsPerCta;
  lu_refactor_kernel<<<ctas, 32 * kWarpsPerCta, 0, s>>>(d, A);
  return 1;
}

int launch_solve(const LuDev &d, const double *A, const double *rhs, double *x, cudaStream_t s) {
  lu_permute_rhs_kernel<<<(d.n + 255) / 256, 256, 0, s>>>(d, rhs);
  int launches = 1;
  for (int l = 0; l < d.nlevels; ++l) {
    const int first = d.level_ptr[l], count = d.level_ptr[l + 1] - first;
    if (count <= 0) continue;
    const int ns = d.pull_short_ptr[l + 1] - d.pull_short_ptr[l];
    if (ns > 0) { lu_pull_short_kernel<<<(ns * 32 + 255) / 256, 256, 0, s>>>(d, A, d.pull_short_ptr[l], ns); ++launches; }
    const int nc = d.pull_chunk_ptr[l + 1] - d.pull_chunk_ptr[l], nl = d.pull_long_ptr[l + 1] - d.pull_long_ptr[l];
    if (nl > 0) {
      lu_pull_chunk_kernel<<<nc, 256, 0, s>>>(d, A, d.pull_chunk_ptr[l]);
      lu_pull_finish_kernel<<<nl, 256, 0, s>>>(d, d.pull_long_ptr[l]);
      launches += 2;
    }
    lu_solve_level_kernel<<<(count + kWarpsPerCta - 1) / kWarpsPerCta, 32 * kWarpsPerCta, 0, s>>>(d, A, first, count, x);
    ++launches;
  }
  return launches;
}

}  // namespace lu
}  // namespace xb
