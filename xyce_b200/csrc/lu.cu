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
#include "pdl.cuh"
#include <cuda_runtime.h>
#include "lu.h"

namespace xb {
namespace lu {

namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kSmemRows = 512;     // rows staged in shared memory per warp (4 KB)

__device__ __forceinline__ bool bad_pivot(double p) { return p == 0.0 || !(fabs(p) <= 1.7976931348623157e308); }

// a / b for the factor scaling and the triangular solves: reciprocal seed (MUFU.RCP64H) + one Newton step + quotient +
// one residual correction -- <= 1 ulp for normal operands, a dependent chain of ~70 cycles where IEEE division is
// ~250.  The division sits on the critical path of every column (pivot -> L column -> next column's updates), which
// is what bounds the refactorization of chain-like blocks.  EVERY LU kernel uses this one routine, so the
// warp-per-block, staged, batched and large-block paths stay bitwise identical to each other; zero / non-finite
// pivots are reported separately (bad_pivot), pivots outside the normal range are not supported.
__device__ __forceinline__ double lu_div(double a, double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  r = fma(fma(-b, r, 1.0), r, r);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}

__global__ void __launch_bounds__(32 * kWarpsPerCta) lu_refactor_kernel(LuView d, const double *__restrict__ A) {
  xb::pdl_wait();
  __shared__ double sx[kWarpsPerCta][kSmemRows];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kWarpsPerCta + warp;
  if (b >= d.nblocks) return;
  if (d.block_big[b]) return;                  // large blocks: lu_big_* kernels; staged small blocks: lu_*_staged_kernel
  const int k0 = d.block_ptr[b], k1 = d.block_ptr[b + 1];
  const int nb = k1 - k0;
  if (nb == 1) {                               // 1x1 block: the pivot is the matrix entry itself
    if (lane == 0) {
      double p = 0.0;
      for (int q = d.acol_ptr[k0]; q < d.acol_ptr[k0 + 1]; ++q) p += A[d.acol_src[q]];
      d.Ux[d.Up[k0 + 1] - 1] = p;
      if (bad_pivot(p)) atomicOr(d.status, 1);
    }
    return;
  }
  double *x = (nb <= kSmemRows) ? (&sx[warp][0] - k0) : d.work;   // x[k0..k1) addressed by position
  bool weak = false;
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
      if (bad_pivot(pivot)) atomicOr(d.status, 1);
    }
    for (int q = lb + lane; q < le; q += 32) {
      const double c = x[d.Li[q]];
      if (d.pivot_check && fabs(pivot) < d.pivot_tol * fabs(c)) weak = true;
      d.Lx[q] = lu_div(c, pivot);
    }
    __syncwarp();
  }
  if (weak) atomicOr(d.status, 4);
}

__global__ void __launch_bounds__(256) lu_permute_rhs_kernel(LuView d, const double *__restrict__ rhs) {
  xb::pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < d.n) d.work[t] = d.row_scale ? rhs[d.row_perm[t]] / d.row_scale[t] : rhs[d.row_perm[t]];
}
// imported plans with row scaling: As = diag(1 / row_scale) A, entry by entry (KLU's SCALE_DIV)
__global__ void __launch_bounds__(256) lu_scale_values_kernel(LuView d, const double *__restrict__ A) {
  xb::pdl_wait();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < d.nnz_a) d.As[k] = A[k] / d.row_scale[d.nz_rowpos[k]];
}

// Off-diagonal pull of one level: y[r] -= sum_q A[offr_src[q]] * y[offr_col[q]].  Short rows: one warp per
// row (lane-strided partial sums, shuffle tree).  Long rows (supply rails with ~1e6 entries): one block per
// 4096-entry chunk with a fixed-shape tree, then one block per row over the chunk partials.  Fixed shapes
// and orders, no atomics.
__global__ void __launch_bounds__(256) lu_pull_short_kernel(LuView d, const double *__restrict__ A, int first, int count) {
  xb::pdl_wait();
  const int w = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= count) return;
  const int r = d.pull_short_rows[first + w];
  double acc = 0.0;
  for (int q = d.offr_ptr[r] + lane; q < d.offr_ptr[r + 1]; q += 32) acc += A[d.offr_src[q]] * d.work[d.offr_col[q]];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) d.work[r] -= acc;
}

// rows with a handful of off-diagonal entries (every ring node couples to the supply column only): one thread per row
__global__ void __launch_bounds__(256) lu_pull_tiny_kernel(LuView d, const double *__restrict__ A, int first, int count) {
  xb::pdl_wait();
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= count) return;
  const int r = d.pull_tiny_rows[first + t];
  double acc = 0.0;
  for (int q = d.offr_ptr[r]; q < d.offr_ptr[r + 1]; ++q) acc += A[d.offr_src[q]] * d.work[d.offr_col[q]];
  d.work[r] -= acc;
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

__global__ void __launch_bounds__(256) lu_pull_chunk_kernel(LuView d, const double *__restrict__ A, int first_chunk) {
  xb::pdl_wait();
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

__global__ void __launch_bounds__(256) lu_pull_finish_kernel(LuView d, int first_slot) {
  xb::pdl_wait();
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
__global__ void __launch_bounds__(32 * kWarpsPerCta) lu_solve_level_kernel(LuView d, const double *__restrict__ A,
                                                                          int first, int count,
                                                                          double *__restrict__ xout) {
  xb::pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.x * kWarpsPerCta + warp;
  if (idx >= count) return;
  const int b = d.level_blocks[first + idx];
  if (d.block_big[b]) return;                  // large blocks: row-level schedule; staged small blocks: lu_solve_staged_kernel
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
    const double yk = lu_div(y[k], d.Ux[ue]);
    __syncwarp();
    if (lane == 0) { y[k] = yk; xout[d.col_perm[k]] = yk; }
    for (int q = d.Up[k] + lane; q < ue; q += 32) y[d.Ui[q]] -= d.Ux[q] * yk;
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// Staged small blocks (block_big == 2): the whole factor of the block -- column pointers, row indices, values
// and one dense column -- sits in a per-warp slice of shared memory, so the column chain of the left-looking
// update runs on shared-memory latency instead of a dozen dependent global loads per column.
// Slice layout: double x[nb], Lx[nl], Ux[nu]; u16 Lp[nb+1], Up[nb+1], Li[nl], Ui[nu]  (indices local to the block:
// a staged block has at most 512 rows and its factor fewer than 1536 entries, so 16 bits are enough and more
// slices fit an SM).  The CTA shape (warps per CTA) is chosen at upload so that the resident warps per SM are
// maximal for this slice size (staged_warps).
// ---------------------------------------------------------------------------------------------
typedef unsigned short sidx_t;
struct StagedView { double *x, *Lx, *Ux; sidx_t *Lp, *Up, *Li, *Ui; int nb, nl, nu, k0, l0, u0; };
__device__ __forceinline__ StagedView staged_load(const LuView &d, int b, unsigned char *slice, int lane, bool values) {
  StagedView v;
  v.k0 = d.block_ptr[b]; v.nb = d.block_ptr[b + 1] - v.k0;
  v.l0 = d.Lp[v.k0]; v.nl = d.Lp[v.k0 + v.nb] - v.l0;
  v.u0 = d.Up[v.k0]; v.nu = d.Up[v.k0 + v.nb] - v.u0;
  v.x = (double *)slice; v.Lx = v.x + v.nb; v.Ux = v.Lx + v.nl;
  v.Lp = (sidx_t *)(v.Ux + v.nu); v.Up = v.Lp + v.nb + 1; v.Li = v.Up + v.nb + 1; v.Ui = v.Li + v.nl;
  for (int i = lane; i <= v.nb; i += 32) { v.Lp[i] = (sidx_t)(d.Lp[v.k0 + i] - v.l0); v.Up[i] = (sidx_t)(d.Up[v.k0 + i] - v.u0); }
  for (int i = lane; i < v.nl; i += 32) { v.Li[i] = (sidx_t)(d.Li[v.l0 + i] - v.k0); if (values) v.Lx[i] = d.Lx[v.l0 + i]; }
  for (int i = lane; i < v.nu; i += 32) { v.Ui[i] = (sidx_t)(d.Ui[v.u0 + i] - v.k0); if (values) v.Ux[i] = d.Ux[v.u0 + i]; }
  return v;
}

__global__ void __launch_bounds__(1024) lu_refactor_staged_kernel(LuView d, const double *__restrict__ A) {
  xb::pdl_wait();
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + warp;
  if (b >= d.nblocks || d.block_big[b] != 2) return;
  StagedView v = staged_load(d, b, smem + (size_t)warp * d.staged_bytes, lane, false);
  for (int i = lane; i < v.nl; i += 32) v.Lx[i] = 0.0;
  for (int i = lane; i < v.nu; i += 32) v.Ux[i] = 0.0;
  __syncwarp();
  // A -> factor slots (acol_dst: >= 0 index into Ux, < 0 ~index into Lx), all columns of the block at once
  const int a0 = d.acol_ptr[v.k0], a1 = d.acol_ptr[v.k0 + v.nb];
  for (int q = a0 + lane; q < a1; q += 32) {
    const int dst = d.acol_dst[q];
    const double val = A[d.acol_src[q]];
    if (dst >= 0) v.Ux[dst - v.u0] = val; else v.Lx[~dst - v.l0] = val;
  }
  __syncwarp();
  bool bad = false, weak = false;
  for (int k = 0; k < v.nb; ++k) {
    const int ub = v.Up[k], ue = v.Up[k + 1] - 1, lb = v.Lp[k], le = v.Lp[k + 1];
    // dense column from the slots
    for (int q = ub + lane; q < ue; q += 32) v.x[v.Ui[q]] = v.Ux[q];
    for (int q = lb + lane; q < le; q += 32) v.x[v.Li[q]] = v.Lx[q];
    if (lane == 0) v.x[k] = v.Ux[ue];
    __syncwarp();
    for (int q = ub; q < ue; ++q) {
      const int i = v.Ui[q];
      const double u = v.x[i];
      if (lane == 0) v.Ux[q] = u;
      for (int t = v.Lp[i] + lane; t < v.Lp[i + 1]; t += 32) v.x[v.Li[t]] -= v.Lx[t] * u;
      __syncwarp();
    }
    const double pivot = v.x[k];
    if (bad_pivot(pivot)) bad = true;
    if (lane == 0) v.Ux[ue] = pivot;
    for (int q = lb + lane; q < le; q += 32) {
      const double c = v.x[v.Li[q]];
      if (d.pivot_check && fabs(pivot) < d.pivot_tol * fabs(c)) weak = true;
      v.Lx[q] = lu_div(c, pivot);
    }
    __syncwarp();
  }
  if (bad && lane == 0) atomicOr(d.status, 1);
  if (weak) atomicOr(d.status, 4);
  for (int i = lane; i < v.nl; i += 32) d.Lx[v.l0 + i] = v.Lx[i];
  for (int i = lane; i < v.nu; i += 32) d.Ux[v.u0 + i] = v.Ux[i];
}

__global__ void __launch_bounds__(1024) lu_solve_staged_kernel(LuView d, int first, int count, double *__restrict__ xout) {
  xb::pdl_wait();
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.x * (blockDim.x >> 5) + warp;
  if (idx >= count) return;
  const int b = d.level_blocks[first + idx];
  if (d.block_big[b] != 2) return;
  StagedView v = staged_load(d, b, smem + (size_t)warp * d.staged_bytes, lane, true);
  for (int i = lane; i < v.nb; i += 32) v.x[i] = d.work[v.k0 + i];
  __syncwarp();
  for (int k = 0; k < v.nb; ++k) {                       // unit-lower forward substitution
    const double yk = v.x[k];
    for (int q = v.Lp[k] + lane; q < v.Lp[k + 1]; q += 32) v.x[v.Li[q]] -= v.Lx[q] * yk;
    __syncwarp();
  }
  for (int k = v.nb - 1; k >= 0; --k) {                  // backward substitution, pivot stored last in the column
    const int ue = v.Up[k + 1] - 1;
    const double yk = lu_div(v.x[k], v.Ux[ue]);
    __syncwarp();
    if (lane == 0) v.x[k] = yk;
    for (int q = v.Up[k] + lane; q < ue; q += 32) v.x[v.Ui[q]] -= v.Ux[q] * yk;
    __syncwarp();
  }
  for (int i = lane; i < v.nb; i += 32) { const double yi = v.x[i]; d.work[v.k0 + i] = yi; xout[d.col_perm[v.k0 + i]] = yi; }
}

// ---------------------------------------------------------------------------------------------
// Batched groups (block_big == 3, lu.h): a TILE of kBundle lanes owns one block, a one-warp CTA holds LB <= 32 / kBundle
// blocks.  The CTA keeps the factor values of its blocks in shared memory, slot-major ([slot][block], conflict-free);
// per step every lane executes ONE operation of the current bundle on its tile's column: loads, arithmetic, store,
// warp barrier (the next bundle may read what a neighbouring lane just wrote).  The program (kBundle 8-byte words
// {dst | type << 14, a, b, 0} per bundle, the same for every block of the group) sits in shared memory too.
// Per warp this is ~25 instructions per bundle instead of ~125 when one lane ran all kBundle operations.
// ---------------------------------------------------------------------------------------------
// 16-byte asynchronous copies global -> shared by the whole warp (n16 = number of 16-byte units); nothing is waited for here
__device__ __forceinline__ void stage_async16(const void *__restrict__ src, void *dst, int n16) {
  const unsigned sdst = (unsigned)__cvta_generic_to_shared(dst);
  const char *g = reinterpret_cast<const char *>(src);
  for (int i = threadIdx.x; i < n16; i += 32)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst + (unsigned)i * 16u), "l"(g + (size_t)i * 16) : "memory");
}
__device__ __forceinline__ void async_commit_wait() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
}
__host__ __device__ constexpr int round16(int bytes) { return (bytes + 15) & ~15; }

__global__ void __launch_bounds__(32) lu_refactor_batched_kernel(LuBatchDev g, const double *__restrict__ A, int LB, double tol,
                                                                 int check, int *status) {
  xb::pdl_wait();
  extern __shared__ __align__(16) unsigned char bsm_raw[];
  // shared memory: values [nu + nl][LB] doubles | program | this CTA's A indices [na][LB] ints | slot of each A entry [na]
  const int lane = threadIdx.x;
  const int ns = g.nu + g.nl;
  double *bsm = reinterpret_cast<double *>(bsm_raw);
  const int off_prog = round16(ns * LB * 8), off_idx = off_prog + round16(g.rf_bundles * kBundle * 8);
  const int off_dst = off_idx + round16(g.na * LB * 4);
  uint2 *prog_s = reinterpret_cast<uint2 *>(bsm_raw + off_prog);
  int *idx_s = reinterpret_cast<int *>(bsm_raw + off_idx);
  int *dst_s = reinterpret_cast<int *>(bsm_raw + off_dst);        // [na] factor slot of each A entry (padded to 16 bytes)
  const int j0 = blockIdx.x * LB;
  // Staging, two memory round trips in all: (1) program and index list of this CTA (contiguous, 16-byte asynchronous
  // copies) while the slots are zeroed; (2) A -> factor slots, one asynchronous 8-byte copy per entry straight into its
  // slot -- every gather of the CTA in flight at once, no registers tied up.
  const int idx_units = round16(g.na * LB * 4) / 16;
  stage_async16(g.rf_prog, prog_s, g.rf_bundles * kBundle * 8 / 16);
  stage_async16(g.a_src_cta + (size_t)blockIdx.x * idx_units * 4, idx_s, idx_units);
  stage_async16(g.a_dst, dst_s, round16(g.na * 4) / 16);
  for (int t = lane; t < ns * LB; t += 32) bsm[t] = 0.0;
  async_commit_wait();
  {
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(bsm);
    const int total = g.na * LB, de = 32 / LB, dl = 32 - de * LB;      // t += 32  <=>  e += de, lj += dl (with carry)
    int e = lane / LB, lj = lane - e * LB;
    for (int t = lane; t < total; t += 32) {
      const int dst = dst_s[e] * LB + lj;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sbase + (unsigned)dst * 8u), "l"(A + idx_s[t]) : "memory");
      e += de; lj += dl;
      if (lj >= LB) { lj -= LB; ++e; }
    }
    async_commit_wait();
  }
  const int blk = lane / kBundle, sub = lane - blk * kBundle;      // tile = block, lane of the tile = operation of the bundle
  const bool active = blk < LB;
  double *v = bsm + (active ? blk : 0);
  int flag = 0;
  uint2 nxt = prog_s[sub];
  for (int bi = 0; bi < g.rf_bundles; ++bi) {
    const uint2 cur = nxt;
    if (bi + 1 < g.rf_bundles) nxt = prog_s[(bi + 1) * kBundle + sub];
    const int ty = (cur.x >> 14) & 3, ds = cur.x & 0x3fff, a = cur.x >> 16, b = cur.y & 0xffff;
    if (active && ty != kOpNop) {
      const double x = v[a * LB];
      if (ty == kOpFnma) {
        const double d = v[ds * LB], y = v[b * LB];
        v[ds * LB] = d - x * y;
      } else if (ty == kOpDiv) {
        const double d = v[ds * LB];
        if (bad_pivot(x)) flag |= 1;
        if (check && fabs(x) < tol * fabs(d)) flag |= 4;      // the fixed pivot no longer passes KLU's threshold test
        v[ds * LB] = lu_div(d, x);
      } else {
        if (bad_pivot(x)) flag |= 1;
      }
    }
    __syncwarp();
  }
  // write back, the whole warp again (slot-major, block fastest: coalesced)
  {
    const int total = ns * LB, de = 32 / LB, dl = 32 - de * LB;
    int sl = lane / LB, lj = lane - sl * LB;
    for (int t = lane; t < total; t += 32) {
      if (j0 + lj < g.nblk) g.LUx[(size_t)sl * g.nblk + j0 + lj] = bsm[t];
      sl += de; lj += dl;
      if (lj >= LB) { lj -= LB; ++sl; }
    }
  }
  if (flag && active && j0 + blk < g.nblk) atomicOr(status, flag);
}

__global__ void __launch_bounds__(32) lu_solve_batched_kernel(LuBatchDev g, LuView d, int LB, double *__restrict__ xout) {
  xb::pdl_wait();
  extern __shared__ __align__(16) double bsm[];   // factor [nu + nl][LB], y [nb][LB], the program (later: column permutation)
  const int lane = threadIdx.x;
  const int ns = g.nu + g.nl;
  uint2 *prog_s = reinterpret_cast<uint2 *>(reinterpret_cast<unsigned char *>(bsm) + round16((ns + g.nb) * LB * 8));
  stage_async16(g.sv_prog, prog_s, g.sv_bundles * kBundle * 8 / 16);
  // factor values and right-hand sides: asynchronous 8-byte copies straight into shared memory, issued by the whole
  // warp (item t = slot t / LB of block t % LB -- coalesced), all in flight at once
  const int j0 = blockIdx.x * LB;
  const int de = 32 / LB, dl = 32 - de * LB;
  {
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(bsm);
    int sl = lane / LB, lj = lane - sl * LB;
    for (int t = lane; t < ns * LB; t += 32) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sbase + (unsigned)t * 8u),
                   "l"(g.LUx + (size_t)sl * g.nblk + min(j0 + lj, g.nblk - 1)) : "memory");
      sl += de; lj += dl;
      if (lj >= LB) { lj -= LB; ++sl; }
    }
  }
  const int blk = lane / kBundle, sub = lane - blk * kBundle;
  const bool active = blk < LB;
  const int kl = __ldg(g.k0 + min(j0 + (active ? blk : 0), g.nblk - 1));      // first position of this tile's block
  if (active) {      // right-hand side of the tile's block: its kBundle lanes share the rows
    const unsigned ybase = (unsigned)__cvta_generic_to_shared(bsm + (size_t)ns * LB + blk);
    for (int i = sub; i < g.nb; i += kBundle)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ybase + (unsigned)(i * LB) * 8u), "l"(d.work + kl + i) : "memory");
  }
  async_commit_wait();
  const double *v = bsm + (active ? blk : 0);
  double *y = bsm + (size_t)ns * LB + (active ? blk : 0);
  uint2 nxt = prog_s[sub];
  for (int bi = 0; bi < g.sv_bundles; ++bi) {
    const uint2 cur = nxt;
    if (bi + 1 < g.sv_bundles) nxt = prog_s[(bi + 1) * kBundle + sub];
    const int ty = (cur.x >> 14) & 3, ds = cur.x & 0x3fff, a = cur.x >> 16, b = cur.y & 0xffff;
    if (active && ty != kOpNop) {
      const double x = v[a * LB], dd = y[ds * LB];
      if (ty == kOpFnma) y[ds * LB] = dd - x * y[b * LB];
      else if (ty == kOpDiv) y[ds * LB] = lu_div(dd, x);
    }
    __syncwarp();
  }
  // solution back to the work vector and, through the column permutation, to the caller's ordering; the permutation
  // entries of the CTA's blocks are fetched (asynchronously, all at once) into the program area, which is dead now --
  // the program needs at least as much room as they do whenever a block has an entry per row
  int *cperm_s = reinterpret_cast<int *>(prog_s);
  const bool staged_perm = g.nb * LB * 4 <= g.sv_bundles * kBundle * 8;
  if (staged_perm) {
    if (active) {
      const unsigned cbase = (unsigned)__cvta_generic_to_shared(cperm_s + blk);
      for (int i = sub; i < g.nb; i += kBundle)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(cbase + (unsigned)(i * LB) * 4u), "l"(d.col_perm + kl + i) : "memory");
    }
    async_commit_wait();
  }
  if (active && j0 + blk < g.nblk)
    for (int i = sub; i < g.nb; i += kBundle) {
      const double yi = y[i * LB];
      d.work[kl + i] = yi; xout[staged_perm ? cperm_s[i * LB + blk] : __ldg(d.col_perm + kl + i)] = yi;
    }
}

// factor values of a batched group -> the ordinary Lx / Ux arrays (export, diagnostics)
__global__ void __launch_bounds__(256) lu_batch_export_kernel(LuBatchDev g, LuView d) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= g.nblk) return;
  const int u0 = g.Up0[j], l0 = g.Lp0[j];
  for (int s = 0; s < g.nu; ++s) d.Ux[u0 + s] = g.LUx[(size_t)s * g.nblk + j];
  for (int s = 0; s < g.nl; ++s) d.Lx[l0 + s] = g.LUx[(size_t)(g.nu + s) * g.nblk + j];
}

// ---------------------------------------------------------------------------------------------
// Large diagonal blocks.  Refactor: one warp per column, columns of one dependency level per launch;
// the column is built in place in the factor arrays (row -> slot by binary search in the fixed pattern).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int lower_bound_dev(const int *a, int lo, int hi, int key) {
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
  return lo;
}

__global__ void __launch_bounds__(256) lu_big_cols_kernel(LuView d, const double *__restrict__ A, int first, int count) {
  xb::pdl_wait();
  const int w = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= count) return;
  const int k = d.rf_cols[first + w];
  const int ub = d.Up[k], ue = d.Up[k + 1] - 1, lb = d.Lp[k], le = d.Lp[k + 1];
  for (int q = ub + lane; q <= ue; q += 32) d.Ux[q] = 0.0;
  for (int q = lb + lane; q < le; q += 32) d.Lx[q] = 0.0;
  __syncwarp();
  for (int q = d.acol_ptr[k] + lane; q < d.acol_ptr[k + 1]; q += 32) {
    const int dst = d.acol_dst[q];
    const double v = A[d.acol_src[q]];
    if (dst >= 0) d.Ux[dst] = v; else d.Lx[~dst] = v;
  }
  __syncwarp();
  for (int q = ub; q < ue; ++q) {
    const int i = d.Ui[q];
    const double u = d.Ux[q];
    for (int t = d.Lp[i] + lane; t < d.Lp[i + 1]; t += 32) {
      const int r = d.Li[t];
      const double val = d.Lx[t] * u;
      if (r < k) d.Ux[lower_bound_dev(d.Ui, q + 1, ue, r)] -= val;
      else if (r == k) d.Ux[ue] -= val;
      else d.Lx[lower_bound_dev(d.Li, lb, le, r)] -= val;
    }
    __syncwarp();
  }
  const double pivot = d.Ux[ue];
  if (lane == 0 && bad_pivot(pivot)) atomicOr(d.status, 1);
  bool weak = false;
  for (int q = lb + lane; q < le; q += 32) {
    const double c = d.Lx[q];
    if (d.pivot_check && fabs(pivot) < d.pivot_tol * fabs(c)) weak = true;
    d.Lx[q] = lu_div(c, pivot);
  }
  if (weak) atomicOr(d.status, 4);
}

// Dense columns (supply rails): x = L^-1 A(:,k) over the whole block by the forward row stages, then gathered.
__global__ void __launch_bounds__(256) lu_dense_scatter_kernel(LuView d, const double *__restrict__ A, int k) {
  xb::pdl_wait();
  const int q = d.acol_ptr[k] + blockIdx.x * 256 + threadIdx.x;
  if (q < d.acol_ptr[k + 1]) d.work2[d.acol_row[q]] = A[d.acol_src[q]];
}
__global__ void __launch_bounds__(256) lu_dense_gather_kernel(LuView d, int k) {
  xb::pdl_wait();
  const int ub = d.Up[k], ue = d.Up[k + 1] - 1, lb = d.Lp[k], le = d.Lp[k + 1];
  const double pivot = d.work2[k];
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t == 0) { d.Ux[ue] = pivot; if (bad_pivot(pivot)) atomicOr(d.status, 1); }
  if (t < ue - ub) d.Ux[ub + t] = d.work2[d.Ui[ub + t]];
  if (t < le - lb) {
    const double c = d.work2[d.Li[lb + t]];
    if (d.pivot_check && fabs(pivot) < d.pivot_tol * fabs(c)) atomicOr(d.status, 4);
    d.Lx[lb + t] = lu_div(c, pivot);
  }
}

// One stage of a row-form triangular sweep: vec[r] -= sum_{col < col_limit} L(r, col) vec[col].
// WARP: one warp per row (lane-strided partial sums, shuffle tree); otherwise one CTA per row (fixed tree).
template <bool WARP>
__global__ void __launch_bounds__(256) lu_fwd_rows_kernel(LuView d, double *vec, const int *__restrict__ rows, int first, int count,
                                                          int col_limit) {
  xb::pdl_wait();
  __shared__ double sh[256];
  const int lane = threadIdx.x & 31;
  const int idx = WARP ? ((blockIdx.x * 256 + threadIdx.x) >> 5) : blockIdx.x;
  if (idx >= count) return;
  const int r = rows[first + idx];
  double acc = 0.0;
  const int step = WARP ? 32 : 256, off = WARP ? lane : threadIdx.x;
  for (int q = d.Lr_ptr[r] + off; q < d.Lr_ptr[r + 1]; q += step) {
    const int c = d.Lr_col[q];
    if (c < col_limit) acc += d.Lx[d.Lr_src[q]] * vec[c];
  }
  if (WARP) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) vec[r] -= acc;
  } else {
    const double t = block_tree_sum(acc, sh);
    if (threadIdx.x == 0) vec[r] -= t;
  }
}
// backward stage: y[r] = (y[r] - sum U(r, col) y[col]) / U(r, r), solution scattered to the caller's ordering
template <bool WARP>
__global__ void __launch_bounds__(256) lu_bwd_rows_kernel(LuView d, const int *__restrict__ rows, int first, int count,
                                                          double *__restrict__ xout) {
  xb::pdl_wait();
  __shared__ double sh[256];
  const int lane = threadIdx.x & 31;
  const int idx = WARP ? ((blockIdx.x * 256 + threadIdx.x) >> 5) : blockIdx.x;
  if (idx >= count) return;
  const int r = rows[first + idx];
  double *y = d.work;
  double acc = 0.0;
  const int step = WARP ? 32 : 256, off = WARP ? lane : threadIdx.x;
  for (int q = d.Ur_ptr[r] + off; q < d.Ur_ptr[r + 1]; q += step) acc += d.Ux[d.Ur_src[q]] * y[d.Ur_col[q]];
  double total;
  if (WARP) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    total = acc;
  } else {
    total = block_tree_sum(acc, sh);
  }
  if ((WARP ? lane : (int)threadIdx.x) == 0) {
    const double yr = lu_div(y[r] - total, d.Ux[d.Up[r + 1] - 1]);
    y[r] = yr;
    xout[d.col_perm[r]] = yr;
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
  cudaFree(d.pull_tiny_rows); cudaFree(d.work2); cudaFree(d.block_big); cudaFree(d.acol_dst); cudaFree(d.rf_cols); cudaFree(d.Lr_ptr); cudaFree(d.Lr_col);
  cudaFree(d.Lr_src); cudaFree(d.Ur_ptr); cudaFree(d.Ur_col); cudaFree(d.Ur_src); cudaFree(d.fs_short_rows); cudaFree(d.fs_long_rows);
  cudaFree(d.bs_short_rows); cudaFree(d.bs_long_rows);
  cudaFree(d.row_scale); cudaFree(d.As); cudaFree(d.nz_rowpos);
  for (LuBatchDev &g : d.batch) {
    cudaFree(g.k0); cudaFree(g.a_dst); cudaFree(g.a_src_cta); cudaFree(g.rf_prog); cudaFree(g.sv_prog); cudaFree(g.LUx);
    cudaFree(g.Up0); cudaFree(g.Lp0);
  }
  const double tol = d.pivot_tol; const int chk = d.pivot_check;
  d = LuDev();
  d.pivot_tol = tol; d.pivot_check = chk;      // options survive a new analysis
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
  UP(pull_tiny_rows) UP(block_big) UP(acol_dst) UP(rf_cols) UP(Lr_ptr) UP(Lr_col) UP(Lr_src) UP(Ur_ptr) UP(Ur_col) UP(Ur_src)
  UP(fs_short_rows) UP(fs_long_rows) UP(bs_short_rows) UP(bs_long_rows)
#undef UP
  d.nnz_a = p.nnz_a;
  if (!p.row_scale.empty()) {
    if ((e = up(&d.row_scale, p.row_scale)) != cudaSuccess) return e;
    if ((e = up(&d.nz_rowpos, p.nz_rowpos)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&d.As, (size_t)(p.nnz_a > 0 ? p.nnz_a : 1) * sizeof(double))) != cudaSuccess) return e;
  }
  d.pull_tiny_ptr = p.pull_tiny_ptr; d.staged_bytes = p.staged_bytes;
  d.rf_level_ptr = p.rf_level_ptr; d.rf_dense_ptr = p.rf_dense_ptr; d.rf_dense_cols = p.rf_dense_cols;
  d.fs_short_ptr = p.fs_short_ptr; d.fs_long_ptr = p.fs_long_ptr; d.bs_short_ptr = p.bs_short_ptr; d.bs_long_ptr = p.bs_long_ptr;
  d.big_blocks = p.big_blocks; d.big_fs_begin = p.big_fs_begin; d.big_fs_end = p.big_fs_end;
  d.big_bs_begin = p.big_bs_begin; d.big_bs_end = p.big_bs_end; d.block_ptr_h = p.block_ptr;
  {
    std::vector<int> level_of_block(d.nblocks, 0);
    for (int l = 0; l < d.nlevels; ++l) for (int q = p.level_ptr[l]; q < p.level_ptr[l + 1]; ++q) level_of_block[p.level_blocks[q]] = l;
    d.block_level_of_big.clear();
    for (int b : p.big_blocks) d.block_level_of_big.push_back(level_of_block[b]);
    d.dense_col_block.clear();
    for (int k : p.rf_dense_cols) {
      int bi = 0;
      while (!(p.block_ptr[p.big_blocks[bi]] <= k && k < p.block_ptr[p.big_blocks[bi] + 1])) ++bi;
      d.dense_col_block.push_back(bi);
    }
  }
  if (d.staged_bytes > 0) {
    // CTA shape: as many resident warps per SM as the shared memory allows (227 KB per SM, 1 KB reserved per CTA,
    // at most 32 CTAs and 64 warps per SM); fewer, fatter CTAs win when the per-CTA reservation matters
    int best_w = 1, best_res = 0;
    for (int w = 1; w <= 32; ++w) {
      const long long per_cta = (long long)w * d.staged_bytes;
      if (per_cta > 227 * 1024 - 1024) break;
      int ctas = (int)((227LL * 1024) / (per_cta + 1024));
      ctas = std::min(ctas, std::min(32, 64 / w));
      if (ctas * w > best_res) { best_res = ctas * w; best_w = w; }
    }
    d.staged_warps = best_w;
    const int bytes = d.staged_warps * d.staged_bytes;
    if ((e = cudaFuncSetAttribute(lu_refactor_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(lu_solve_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
  }
  // batched groups: device copies, CTA width (blocks per one-warp CTA): the narrowest of 8 / 16 / 32 whose grid still
  // fits the SMs in one wave (more CTAs per SM = more latency chains in flight per SM)
  {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    size_t max_rf = 0, max_sv = 0;
    for (const LuPlan::BatchGroup &pg : p.batch) {
      LuBatchDev g;
      g.nblk = (int)pg.blocks.size(); g.nb = pg.nb; g.nu = pg.nu; g.nl = pg.nl; g.na = pg.na; g.level = pg.level;
      g.rf_bundles = pg.rf_bundles; g.sv_bundles = pg.sv_bundles;
      std::vector<int> k0(g.nblk), up0(g.nblk), lp0(g.nblk);
      for (int j = 0; j < g.nblk; ++j) { k0[j] = p.block_ptr[pg.blocks[j]]; up0[j] = p.Up[k0[j]]; lp0[j] = p.Lp[k0[j]]; }
      if ((e = up(&g.k0, k0)) != cudaSuccess) return e;
      if ((e = up(&g.Up0, up0)) != cudaSuccess) return e;
      if ((e = up(&g.Lp0, lp0)) != cudaSuccess) return e;
      { std::vector<int> ad(pg.a_dst); ad.resize((ad.size() + 3) / 4 * 4, 0); if ((e = up(&g.a_dst, ad)) != cudaSuccess) return e; }      // padded: copied in 16-byte units
      if ((e = up(&g.rf_prog, pg.rf_prog)) != cudaSuccess) return e;
      if ((e = up(&g.sv_prog, pg.sv_prog)) != cudaSuccess) return e;
      {   // values of the first (pivoting, host) factorization, interleaved -- a solve may follow the analysis directly
        std::vector<double> lux((size_t)(g.nu + g.nl) * g.nblk, 0.0);
        if (p.Ux.size() == p.Ui.size() && p.Lx.size() == p.Li.size())
          for (int j = 0; j < g.nblk; ++j) {
            for (int sl = 0; sl < g.nu; ++sl) lux[(size_t)sl * g.nblk + j] = p.Ux[up0[j] + sl];
            for (int sl = 0; sl < g.nl; ++sl) lux[(size_t)(g.nu + sl) * g.nblk + j] = p.Lx[lp0[j] + sl];
          }
        if ((e = up(&g.LUx, lux)) != cudaSuccess) return e;
      }
      // blocks per one-warp CTA (a tile of kBundle lanes each): the work is a latency chain per warp, so the time is
      // (number of waves) x (chain time); take the width that needs the fewest waves, and among those the widest
      // (fewer CTAs to launch and to stage programs for: measured 53 us against 92 us at two waves)
      auto pick = [&](size_t bytes_per_block, size_t prog_bytes) {
        int best = 1; long long best_waves = -1;
        for (int lb = 1; lb <= 32 / kBundle; ++lb) {
          const size_t per_cta = bytes_per_block * lb + prog_bytes + 1024;
          if (per_cta > 227 * 1024) break;
          const long long resident = (long long)sms * std::min<long long>(32, (227 * 1024) / per_cta);
          const long long ctas = (g.nblk + lb - 1) / lb, waves = (ctas + resident - 1) / resident;
          if (best_waves < 0 || waves <= best_waves) { best_waves = waves; best = lb; }      // ties: the widest (fewest CTAs; measured)
        }
        return best;
      };
      g.lanes_rf = pick((size_t)(g.nu + g.nl) * 8 + (size_t)g.na * 4, (size_t)g.rf_bundles * kBundle * 8 + (size_t)g.na * 4 + 64);
      g.lanes_sv = pick((size_t)(g.nu + g.nl + g.nb) * 8, (size_t)g.sv_bundles * kBundle * 8 + 32);
      g.smem_rf = round16((g.nu + g.nl) * g.lanes_rf * 8) + round16(g.rf_bundles * kBundle * 8) + round16(g.na * g.lanes_rf * 4) + round16(g.na * 4);
      g.smem_sv = round16((g.nu + g.nl + g.nb) * g.lanes_sv * 8) + g.sv_bundles * kBundle * 8;
      max_rf = std::max(max_rf, (size_t)g.smem_rf);
      max_sv = std::max(max_sv, (size_t)g.smem_sv);
      {   // A indices regrouped per CTA: [cta][entry][block of the CTA] (one contiguous, 16-byte padded chunk per CTA)
        const int lb = g.lanes_rf, nctas = (g.nblk + lb - 1) / lb, chunk = round16(g.na * lb * 4) / 4;
        std::vector<int> cta((size_t)nctas * chunk, 0);
        for (int c = 0; c < nctas; ++c)
          for (int en = 0; en < g.na; ++en)
            for (int lj = 0; lj < lb; ++lj)
              cta[(size_t)c * chunk + en * lb + lj] = pg.a_src[(size_t)en * g.nblk + std::min(c * lb + lj, g.nblk - 1)];
        if ((e = up(&g.a_src_cta, cta)) != cudaSuccess) return e;
      }
      d.batch.push_back(g);
    }
    if (max_rf > 48 * 1024 && (e = cudaFuncSetAttribute(lu_refactor_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_rf)) != cudaSuccess) return e;
    if (max_sv > 48 * 1024 && (e = cudaFuncSetAttribute(lu_solve_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_sv)) != cudaSuccess) return e;
  }
  if ((e = cudaMalloc((void **)&d.work2, (size_t)(p.n > 0 ? p.n : 1) * sizeof(double))) != cudaSuccess) return e;
  d.pull_short_ptr = p.pull_short_ptr; d.pull_long_ptr = p.pull_long_ptr; d.pull_chunk_ptr = p.pull_chunk_ptr;
  if ((e = cudaMalloc((void **)&d.pull_partials, (p.pull_chunk_begin.size() + 1) * sizeof(double))) != cudaSuccess) return e;
  if ((e = cudaMalloc((void **)&d.work, (size_t)(p.n > 0 ? p.n : 1) * sizeof(double))) != cudaSuccess) return e;
  if ((e = cudaMalloc((void **)&d.status, sizeof(int))) != cudaSuccess) return e;
  return cudaMemset(d.status, 0, sizeof(int));
}

namespace {
// forward row stages [s0, s1) of one large block applied to `vec`
int run_fwd_stages(const LuDev &d, double *vec, int s0, int s1, int col_limit, cudaStream_t s) {
  int launches = 0;
  for (int st = s0; st < s1; ++st) {
    const int ns = d.fs_short_ptr[st + 1] - d.fs_short_ptr[st], nl = d.fs_long_ptr[st + 1] - d.fs_long_ptr[st];
    if (ns > 0) { xb::launch_pdl(lu_fwd_rows_kernel<true>, dim3((ns * 32 + 255) / 256), dim3(256), 0, s, d, vec, d.fs_short_rows, d.fs_short_ptr[st], ns, col_limit); ++launches; }
    if (nl > 0) { xb::launch_pdl(lu_fwd_rows_kernel<false>, dim3(nl), dim3(256), 0, s, d, vec, d.fs_long_rows, d.fs_long_ptr[st], nl, col_limit); ++launches; }
  }
  return launches;
}
}  // namespace

int launch_refactor(const LuDev &d, const double *A, cudaStream_t s) {
  cudaMemsetAsync(d.status, 0, sizeof(int), s);
  int extra = 0;
  if (d.row_scale) { xb::launch_pdl(lu_scale_values_kernel, dim3((d.nnz_a + 255) / 256), dim3(256), 0, s, d, A); A = d.As; extra = 1; }
  const int ctas = (d.nblocks + kWarpsPerCta - 1) / kWarpsPerCta;
  xb::launch_pdl(lu_refactor_kernel, dim3(ctas), dim3(32 * kWarpsPerCta), 0, s, d, A);
  int launches = 1 + extra;
  if (d.staged_bytes > 0) {
    const int w = d.staged_warps;
    xb::launch_pdl(lu_refactor_staged_kernel, dim3((d.nblocks + w - 1) / w), dim3(32 * w), (size_t)w * d.staged_bytes, s, d, A); ++launches;
  }
  for (const LuBatchDev &g : d.batch) {
    const int lb = g.lanes_rf;
    xb::launch_pdl(lu_refactor_batched_kernel, dim3((g.nblk + lb - 1) / lb), dim3(32), (size_t)g.smem_rf, s, g, A, lb,
                   d.pivot_tol, d.pivot_check, d.status);
    ++launches;
  }
  // large blocks: columns level by level; dense columns of a level after its normal columns
  const int nlev = (int)d.rf_level_ptr.size() - 1;
  for (int l = 0; l < nlev; ++l) {
    const int first = d.rf_level_ptr[l], count = d.rf_level_ptr[l + 1] - first;
    if (count > 0) { xb::launch_pdl(lu_big_cols_kernel, dim3((count * 32 + 255) / 256), dim3(256), 0, s, d, A, first, count); ++launches; }
    for (int q = d.rf_dense_ptr[l]; q < d.rf_dense_ptr[l + 1]; ++q) {
      const int k = d.rf_dense_cols[q], bi = d.dense_col_block[q], b = d.big_blocks[bi];
      const int k0 = d.block_ptr_h[b], k1 = d.block_ptr_h[b + 1];
      cudaMemsetAsync(d.work2 + k0, 0, (size_t)(k1 - k0) * sizeof(double), s);
      xb::launch_pdl(lu_dense_scatter_kernel, dim3((k1 - k0 + 255) / 256 + 1), dim3(256), 0, s, d, A, k);      // >= number of A entries in the column
      launches += 1 + run_fwd_stages(d, d.work2, d.big_fs_begin[bi], d.big_fs_end[bi], k, s);
      xb::launch_pdl(lu_dense_gather_kernel, dim3((k1 - k0 + 255) / 256 + 1), dim3(256), 0, s, d, k);
      ++launches;
    }
  }
  return launches;
}

int launch_batch_export(const LuDev &d, cudaStream_t s) {
  int launches = 0;
  for (const LuBatchDev &g : d.batch) { lu_batch_export_kernel<<<(g.nblk + 255) / 256, 256, 0, s>>>(g, (const LuView &)d); ++launches; }
  return launches;
}

int launch_solve(const LuDev &d, const double *A, const double *rhs, double *x, cudaStream_t s) {
  if (d.row_scale) A = d.As;      // off-diagonal entries: the scaled copy made by the last refactorization of these values
  xb::launch_pdl(lu_permute_rhs_kernel, dim3((d.n + 255) / 256), dim3(256), 0, s, d, rhs);
  int launches = 1;
  for (int l = 0; l < d.nlevels; ++l) {
    const int first = d.level_ptr[l], count = d.level_ptr[l + 1] - first;
    if (count <= 0) continue;
    const int nt = d.pull_tiny_ptr[l + 1] - d.pull_tiny_ptr[l];
    if (nt > 0) { xb::launch_pdl(lu_pull_tiny_kernel, dim3((nt + 255) / 256), dim3(256), 0, s, d, A, d.pull_tiny_ptr[l], nt); ++launches; }
    const int ns = d.pull_short_ptr[l + 1] - d.pull_short_ptr[l];
    if (ns > 0) { xb::launch_pdl(lu_pull_short_kernel, dim3((ns * 32 + 255) / 256), dim3(256), 0, s, d, A, d.pull_short_ptr[l], ns); ++launches; }
    const int nc = d.pull_chunk_ptr[l + 1] - d.pull_chunk_ptr[l], nl = d.pull_long_ptr[l + 1] - d.pull_long_ptr[l];
    if (nl > 0) {
      xb::launch_pdl(lu_pull_chunk_kernel, dim3(nc), dim3(256), 0, s, d, A, d.pull_chunk_ptr[l]);
      xb::launch_pdl(lu_pull_finish_kernel, dim3(nl), dim3(256), 0, s, d, d.pull_long_ptr[l]);
      launches += 2;
    }
    xb::launch_pdl(lu_solve_level_kernel, dim3((count + kWarpsPerCta - 1) / kWarpsPerCta), dim3(32 * kWarpsPerCta), 0, s, d, A, first, count, x);
    ++launches;
    if (d.staged_bytes > 0) {
      const int w = d.staged_warps;
      xb::launch_pdl(lu_solve_staged_kernel, dim3((count + w - 1) / w), dim3(32 * w), (size_t)w * d.staged_bytes, s, d, first, count, x);
      ++launches;
    }
    for (const LuBatchDev &g : d.batch) {
      if (g.level != l) continue;
      const int lb = g.lanes_sv;
      xb::launch_pdl(lu_solve_batched_kernel, dim3((g.nblk + lb - 1) / lb), dim3(32), (size_t)g.smem_sv, s, g,
                     (const LuView &)d, lb, x);
      ++launches;
    }
    for (size_t bi = 0; bi < d.big_blocks.size(); ++bi) {
      if (d.block_level_of_big[bi] != l) continue;
      launches += run_fwd_stages(d, d.work, d.big_fs_begin[bi], d.big_fs_end[bi], 0x7fffffff, s);
      for (int st = d.big_bs_begin[bi]; st < d.big_bs_end[bi]; ++st) {
        const int nsr = d.bs_short_ptr[st + 1] - d.bs_short_ptr[st], nlr = d.bs_long_ptr[st + 1] - d.bs_long_ptr[st];
        if (nsr > 0) { xb::launch_pdl(lu_bwd_rows_kernel<true>, dim3((nsr * 32 + 255) / 256), dim3(256), 0, s, d, d.bs_short_rows, d.bs_short_ptr[st], nsr, x); ++launches; }
        if (nlr > 0) { xb::launch_pdl(lu_bwd_rows_kernel<false>, dim3(nlr), dim3(256), 0, s, d, d.bs_long_rows, d.bs_long_ptr[st], nlr, x); ++launches; }
      }
    }
  }
  return launches;
}

}  // namespace lu
}  // namespace xb
