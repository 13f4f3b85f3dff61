// xyce_b200 -- bordered block-diagonal solve and the multi-GPU pieces of the Newton step (sm_100a + NCCL).
//
// Reference behaviour being replaced: Xyce's MPI "parallel load" lets every rank load its device
// instances into overlapped vectors / matrices and then exports the ghost rows with Add
// (N_LOA_CktLoader.C:600-601, :816-829; N_LAS_EpetraMultiVector.C:843-849, N_LAS_EpetraMatrix.C:202-208), imports the
// halo of the solution (:468-470), and the direct solve gathers the matrix on one rank.  Here:
//   * every rank (one process per GPU) keeps its partition's unknowns in the order [interior | border]; the border
//     unknowns -- the ones touched from several partitions: supply rails, source branches -- are replicated on every
//     rank in the same order;
//   * xg_dist_reduce_border_rows: after the local assembly the border rows of F, Q, dFdxdVp, dQdxdVp are summed over
//     the ranks: pack -> ncclAllReduce -> unpack on the context's stream, no host synchronisation (the "export with
//     Add" of the reference);
//   * xg_border_solve: the linear system has bordered block-diagonal form.  Each rank factors its interior block
//     A_ii with the KLU-pattern LU (BTF blocks, batched groups, ...), solves A_ii Y = [A_is | b_i], forms its part of
//     the Schur complement  S = A_ss - A_si Y_s,  g = b_s - A_si y;  ONE all-reduce sums the parts, every rank solves the
//     small dense system redundantly and back-substitutes its interior unknowns.  The BTF diagonal blocks are thereby
//     distributed over the GPUs; only the border system is replicated (its size is the number of shared unknowns).
//   * xg_dist_allgather: the scalars of the Newton convergence test / step control (norm parts, device-convergence
//     flag) travel in one small all-gather per evaluation and are combined identically on every rank
//     (DeviceMgr::allDevicesConverged's reduction, Core/N_DEV_DeviceMgr.C:5628, and the Epetra norms).
// The same border machinery serves a single GPU (world = 1, no communicator): dense nodes declared as border leave
// the BTF blocks, e.g. the supply node of an inverter array, whose removal turns one 100 001-row block into 50 000
// equal 2 x 2 blocks (one batched group).
// NCCL is loaded at run time (dlopen "libnccl.so.2": the copy a host application such as PyTorch has already mapped,
// else the system one); the library itself has no link-time dependency on it.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "ctx.h"
#include "pdl.cuh"
#include "vecops.cuh"

namespace {

#define XD_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return xg_fail(ctx, 100 + (int)e_, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

// ---- the handful of NCCL entry points used, resolved at run time ----
struct NcclId { char internal[128]; };
struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(NcclId *) = nullptr;
  int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok() const { return GetUniqueId && CommInitRank && CommDestroy && AllReduce && AllGather && GetErrorString; }
};
constexpr int kNcclFloat64 = 8, kNcclSum = 0;

NcclApi &nccl() {
  static NcclApi api;
  if (api.lib) return api;
  for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
    api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return api;
  api.GetUniqueId = (int (*)(NcclId *))dlsym(api.lib, "ncclGetUniqueId");
  api.CommInitRank = (int (*)(void **, int, NcclId, int))dlsym(api.lib, "ncclCommInitRank");
  api.CommDestroy = (int (*)(void *))dlsym(api.lib, "ncclCommDestroy");
  api.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(api.lib, "ncclAllReduce");
  api.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(api.lib, "ncclAllGather");
  api.GetErrorString = (const char *(*)(int))dlsym(api.lib, "ncclGetErrorString");
  return api;
}

template <class T> cudaError_t up(T **d, const std::vector<T> &v) {
  cudaFree(*d); *d = nullptr;
  cudaError_t e = cudaMalloc((void **)d, (v.empty() ? 1 : v.size()) * sizeof(T));
  if (e == cudaSuccess && !v.empty()) e = cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return e;
}

constexpr int kSiChunk = 1024;

// ---- kernels ----
__global__ void __launch_bounds__(256) pack_rows_k(int nvec, int ns, int ni, double *v0, double *v1, double *v2, double *v3, double *buf) {
  xb::pdl_wait();
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= nvec * ns) return;
  double *const v[4] = {v0, v1, v2, v3};
  const int k = t / ns, r = t - k * ns;
  buf[t] = v[k][ni + r];
}
__global__ void __launch_bounds__(256) unpack_rows_k(int nvec, int ns, int ni, double *v0, double *v1, double *v2, double *v3, const double *buf) {
  xb::pdl_wait();
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= nvec * ns) return;
  double *const v[4] = {v0, v1, v2, v3};
  const int k = t / ns, r = t - k * ns;
  v[k][ni + r] = buf[t];
}

// B[c][row] = A_is(row, c) for the border columns (B zeroed before), B[ns][.] = b_i
__global__ void __launch_bounds__(256) scatter_is_k(int n_is, const int *__restrict__ pos, const int *__restrict__ row,
                                                    const int *__restrict__ col, const double *__restrict__ A, int ni, double *B) {
  xb::pdl_wait();
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t < n_is) B[(size_t)col[t] * ni + row[t]] = A[pos[t]];
}

__device__ __forceinline__ double tree256(double v, double *sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

// partial[c][chunk] = sum over the chunk's entries (border row r, interior column j) of A(r, j) * Y[c][j], c = 0 .. ns
// (fixed-shape tree: deterministic)
__global__ void __launch_bounds__(256) si_chunk_k(int ncols, int n_chunks, const int *__restrict__ cb, const int *__restrict__ ce,
                                                  const int *__restrict__ pos, const int *__restrict__ col,
                                                  const double *__restrict__ A, const double *__restrict__ Y, int ni, double *partial) {
  xb::pdl_wait();
  __shared__ double sh[256];
  const int ch = blockIdx.x, b = cb[ch], e = ce[ch];
  for (int c = 0; c < ncols; ++c) {
    double acc = 0.0;
    for (int q = b + threadIdx.x; q < e; q += 256) acc += A[pos[q]] * Y[(size_t)c * ni + col[q]];
    const double s = tree256(acc, sh);
    if (threadIdx.x == 0) partial[(size_t)c * n_chunks + ch] = s;
  }
}

// red[r][c] = base(r, c) - sum of row r's chunk partials for column c, in chunk order
//   base: c < ns: this rank's A_ss(r, c) (0 if not in the pattern); c == ns: b_s(r) when this rank carries the border
//   right-hand side (take_b), else 0
__global__ void __launch_bounds__(64) si_finish_k(int ns, int n_chunks, const int *__restrict__ row_chunk_ptr,
                                                  const double *__restrict__ partial, const int *__restrict__ ss_pos,
                                                  const double *__restrict__ A, const double *__restrict__ rhs, int ni, int take_b,
                                                  double *red) {
  xb::pdl_wait();
  const int t = blockIdx.x * 64 + threadIdx.x;
  if (t >= ns * (ns + 1)) return;
  const int r = t / (ns + 1), c = t - r * (ns + 1);
  double base;
  if (c < ns) { const int p = ss_pos[r * ns + c]; base = p >= 0 ? A[p] : 0.0; }
  else base = take_b ? rhs[ni + r] : 0.0;
  double acc = 0.0;
  for (int ch = row_chunk_ptr[r]; ch < row_chunk_ptr[r + 1]; ++ch) acc += partial[(size_t)c * n_chunks + ch];
  red[t] = base - acc;
}

// Dense border system [S | g] (ns x (ns + 1), row major in global memory): Gauss-Jordan with partial pivoting by one
// CTA; on exit column ns holds the solution.  status bit 3 is set when the system is singular.
__global__ void __launch_bounds__(256) dense_solve_k(int ns, double *M, int *status) {
  xb::pdl_wait();
  extern __shared__ double sm[];          // the augmented matrix
  __shared__ int piv;
  const int ld = ns + 1;
  for (int t = threadIdx.x; t < ns * ld; t += 256) sm[t] = M[t];
  __syncthreads();
  for (int k = 0; k < ns; ++k) {
    if (threadIdx.x == 0) {
      int best = k; double bv = fabs(sm[k * ld + k]);
      for (int r = k + 1; r < ns; ++r) { const double v = fabs(sm[r * ld + k]); if (v > bv) { bv = v; best = r; } }
      piv = best;
      if (!(bv > 0.0) || !(bv <= 1.7976931348623157e308)) atomicOr(status, 8);
    }
    __syncthreads();
    const int p = piv;
    if (p != k) for (int c = threadIdx.x; c < ld; c += 256) { const double t = sm[k * ld + c]; sm[k * ld + c] = sm[p * ld + c]; sm[p * ld + c] = t; }
    __syncthreads();
    const double d = sm[k * ld + k];
    // eliminate column k from every other row (each thread owns whole elements; the pivot row is read-only here)
    for (int t = threadIdx.x; t < ns * ld; t += 256) {
      const int r = t / ld, c = t - r * ld;
      if (r == k || c <= k) continue;
      sm[t] -= (sm[r * ld + k] / d) * sm[k * ld + c];
    }
    __syncthreads();
    for (int r = threadIdx.x; r < ns; r += 256) if (r != k) sm[r * ld + k] = 0.0;
    __syncthreads();
  }
  for (int r = threadIdx.x; r < ns; r += 256) M[r * ld + ns] = sm[r * ld + ns] / sm[r * ld + r];
}

// x_i = y - sum_c Y[c] xs[c] (c ascending), x_s = xs
__global__ void __launch_bounds__(256) back_subst_k(int ni, int ns, const double *__restrict__ Y, const double *__restrict__ red, double *x) {
  xb::pdl_wait();
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int ld = ns + 1;
  if (i < ni) {
    double v = Y[(size_t)ns * ni + i];
    for (int c = 0; c < ns; ++c) v -= Y[(size_t)c * ni + i] * red[c * ld + ns];
    x[i] = v;
  } else if (i < ni + ns) {
    x[i] = red[(i - ni) * ld + ns];
  }
}

// ---- small collectives over NVLink peer memory (one kernel, no NCCL launch) -------------------------------------
// The collectives of this path move a handful of doubles (border rows, the border system, norm parts).  NCCL needs
// ~15-25 us for such a message; the exchange below needs one kernel of a few microseconds: every rank owns a mailbox
// in its own HBM (cudaMalloc + CUDA IPC, mapped by all peers through NVSwitch), writes its values straight into slot
// [rank] of every peer's mailbox (P2P stores), publishes a per-sender epoch flag (release, system scope), waits for
// the flags of all senders in its own mailbox (acquire) and combines the slots IN RANK ORDER -- every rank computes
// bitwise the same sum.  Two slot sets alternate with the epoch parity: a sender can only be one epoch ahead of the
// slowest receiver (it needs that receiver's flag of the previous epoch to finish its own), so the set it writes is
// never the one a peer still reads.  A spin that does not finish within ~2 s sets an error flag instead of hanging.
constexpr int kBoxDoubles = 64;                    // payload per sender and epoch
struct P2PView {
  int rank, world;
  double *box[16];                                 // mailbox of every rank as mapped in THIS process (box[rank] = own)
  // mailbox layout: double slots[2][world][kBoxDoubles]; unsigned long long flags[world]; int error
};
__device__ __forceinline__ double *box_slots(double *box, int world, int parity, int sender) { return box + ((size_t)parity * world + sender) * kBoxDoubles; }
__device__ __forceinline__ unsigned long long *box_flags(double *box, int world) { return reinterpret_cast<unsigned long long *>(box + (size_t)2 * world * kBoxDoubles); }

// mode 0: out[0 .. k) = sum over ranks (rank order); mode 1: out[r * k + i] = value i of rank r (all-gather).
// in / out may be vectors addressed through an index list: gather_idx / scatter_idx (null = contiguous).
__global__ void __launch_bounds__(64) p2p_exchange_k(P2PView v, unsigned long long epoch, int k, int mode, const double *in, double *out,
                                                    double *const *in_vecs, int nvec, int ns, int ni) {
  xb::pdl_wait();
  const int t = threadIdx.x, parity = (int)(epoch & 1);
  // payload: either the contiguous `in`, or border rows [ni, ni + ns) of up to 4 vectors (fused pack)
  double mine = 0.0;
  if (t < k) mine = in_vecs ? in_vecs[t / ns][ni + (t % ns)] : in[t];
  if (t < k) for (int p = 0; p < v.world; ++p) box_slots(v.box[p], v.world, parity, v.rank)[t] = mine;
  __threadfence_system();
  __syncthreads();
  if (t < v.world) {
    unsigned long long *f = box_flags(v.box[t], v.world) + v.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
  }
  if (t < v.world) {
    const unsigned long long *f = box_flags(v.box[v.rank], v.world) + t;
    unsigned long long seen = 0;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(f) : "memory");
      if (seen < epoch && clock64() - t0 > 4000000000LL) { *reinterpret_cast<int *>(box_flags(v.box[v.rank], v.world) + v.world) = 1; break; }
    } while (seen < epoch);
  }
  __syncthreads();
  double *own = v.box[v.rank];
  if (mode == 0) {
    if (t < k) {
      double acc = 0.0;
      for (int r = 0; r < v.world; ++r) acc += box_slots(own, v.world, parity, r)[t];
      if (in_vecs) in_vecs[t / ns][ni + (t % ns)] = acc; else out[t] = acc;
    }
  } else {
    for (int q = t; q < k * v.world; q += 64) out[q] = box_slots(own, v.world, parity, q / k)[q % k];
  }
}

int nccl_fail(xgpu_ctx *ctx, int rc, const char *what) {
  return xg_fail(ctx, 300 + rc, std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(rc) : "NCCL error"));
}

XgDist *ensure(xgpu_ctx *ctx) {
  if (!ctx->dist) ctx->dist = new XgDist;
  return ctx->dist;
}

}  // namespace

void xg_dist_free(xgpu_ctx *ctx) {
  XgDist *d = ctx->dist;
  if (!d) return;
  if (d->comm && nccl().CommDestroy) nccl().CommDestroy(d->comm);
  for (int r = 0; r < 16; ++r) if (d->p2p_box[r] && r != d->rank && d->p2p_attached) cudaIpcCloseMemHandle(d->p2p_box[r]);
  if (d->p2p_own) cudaFree(d->p2p_own);
  cudaFree(d->p2p_vecs);
  cudaFree(d->is_pos); cudaFree(d->is_row); cudaFree(d->is_col); cudaFree(d->si_pos); cudaFree(d->si_col);
  cudaFree(d->chunk_row); cudaFree(d->chunk_begin); cudaFree(d->chunk_end); cudaFree(d->row_chunk_ptr); cudaFree(d->ss_pos);
  cudaFree(d->B); cudaFree(d->partials); cudaFree(d->red); cudaFree(d->pack); cudaFreeHost(d->h_pack);
  delete d;
  ctx->dist = nullptr;
}

bool xg_dist_multi(const xgpu_ctx *ctx) { return ctx->dist && (ctx->dist->comm || ctx->dist->p2p_attached) && ctx->dist->world > 1; }

namespace {
P2PView p2p_view(const XgDist *d) {
  P2PView v{};
  v.rank = d->rank; v.world = d->world;
  for (int r = 0; r < d->world && r < 16; ++r) v.box[r] = d->p2p_box[r];
  return v;
}
bool p2p_ready(const XgDist *d) { return d && d->p2p_attached && d->world > 1 && d->world <= 16; }
}  // namespace

int xg_dist_reduce_border_rows(xgpu_ctx *ctx, double *const *vecs, int nvec) {
  XgDist *d = ctx->dist;
  if (!d || d->world <= 1 || d->ns == 0 || nvec <= 0) return 0;
  if (nvec > 4) return xg_fail(ctx, 1, "at most 4 vectors per border reduction");
  if (p2p_ready(d) && nvec * d->ns <= kBoxDoubles) {      // fused pack + exchange + unpack: one kernel
    bool same = true;      // the vectors rarely change between calls: upload the pointer list only when they do
    for (int k = 0; k < nvec; ++k) same = same && d->p2p_vecs_host[k] == vecs[k];
    if (!same) {
      for (int k = 0; k < 4; ++k) d->p2p_vecs_host[k] = k < nvec ? vecs[k] : nullptr;
      XD_CUDA(cudaMemcpyAsync(d->p2p_vecs, d->p2p_vecs_host, 4 * sizeof(double *), cudaMemcpyHostToDevice, ctx->stream));
    }
    xb::launch_pdl(p2p_exchange_k, dim3(1), dim3(64), 0, ctx->stream, p2p_view(d), ++d->p2p_epoch, nvec * d->ns, 0, (const double *)nullptr,
                   (double *)nullptr, (double *const *)d->p2p_vecs, nvec, d->ns, d->ni);
    ++ctx->launches;
    return 0;
  }
  if (!d->comm) return 0;
  double *v[4] = {nullptr, nullptr, nullptr, nullptr};
  for (int k = 0; k < nvec; ++k) v[k] = vecs[k];
  const int cnt = nvec * d->ns, blocks = (cnt + 255) / 256;
  xb::launch_pdl(pack_rows_k, dim3(blocks), dim3(256), 0, ctx->stream, nvec, d->ns, d->ni, v[0], v[1], v[2], v[3], d->pack);
  const int rc = nccl().AllReduce(d->pack, d->pack, (size_t)cnt, kNcclFloat64, kNcclSum, d->comm, ctx->stream);
  if (rc) return nccl_fail(ctx, rc, "ncclAllReduce (border rows)");
  xb::launch_pdl(unpack_rows_k, dim3(blocks), dim3(256), 0, ctx->stream, nvec, d->ns, d->ni, v[0], v[1], v[2], v[3], (const double *)d->pack);
  ctx->launches += 2;
  return 0;
}

int xg_dist_allgather(xgpu_ctx *ctx, const double *d_send, int k) {
  XgDist *d = ctx->dist;
  if (!d || k <= 0 || k > 8) return xg_fail(ctx, 1, "allgather: 1 .. 8 doubles per rank");
  double *recv = d->pack + 8 * d->ns + 8;          // behind the border-row pack area
  if (p2p_ready(d) && k <= kBoxDoubles) {
    xb::launch_pdl(p2p_exchange_k, dim3(1), dim3(64), 0, ctx->stream, p2p_view(d), ++d->p2p_epoch, k, 1, d_send, recv,
                   (double *const *)nullptr, 0, 1, 0);
    ++ctx->launches;
  } else if (d->comm && d->world > 1) {
    const int rc = nccl().AllGather(d_send, recv, (size_t)k, kNcclFloat64, d->comm, ctx->stream);
    if (rc) return nccl_fail(ctx, rc, "ncclAllGather");
  } else {
    XD_CUDA(cudaMemcpyAsync(recv, d_send, k * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  XD_CUDA(cudaMemcpyAsync(d->h_pack, recv, (size_t)k * d->world * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return 0;
}

int xg_border_solve(xgpu_ctx *ctx, const double *J, const double *rhs, double *x, int rhs_border_reduced, bool defer_status) {
  XgDist *d = ctx->dist;
  if (!d || !d->analyzed) return xg_fail(ctx, 113, "xgpu_border_analyze has not been called");
  const int ni = d->ni, ns = d->ns;
  cudaStream_t s = ctx->stream;
  // interior factorization (fixed pattern and pivot sequence; status checked like xgpu_lu_refactor)
  // A bad or sub-threshold pivot re-pivots HERE (host analysis of this rank's interior on the current values, whose
  // factor values are then already those of J) so that every rank still issues the same sequence of collectives.
  // defer_status (single rank only): launches only, the caller reads the status word with its own synchronisation
  // (xg_lu_status) and repeats the solve after a re-analysis if it has to.
  if (ni > 0) {
    int rc = (defer_status && !xg_dist_multi(ctx) && !ctx->lu_repivot) ? xg_lu_refactor_async(ctx, J) : xgpu_lu_refactor(ctx, J);
    if (rc == 2 || rc == 3) { rc = xgpu_border_analyze(ctx, J); ++d->reanalyses; }
    if (rc) return rc;
  }
  // right-hand sides [A_is | b_i] -> Y = A_ii^-1 (.)
  if (ni > 0) {
    if (ns > 0) XD_CUDA(cudaMemsetAsync(d->B, 0, (size_t)ns * ni * sizeof(double), s));
    if (d->n_is > 0) { xb::launch_pdl(scatter_is_k, dim3((d->n_is + 255) / 256), dim3(256), 0, s, d->n_is, (const int *)d->is_pos, (const int *)d->is_row, (const int *)d->is_col, J, ni, d->B); ++ctx->launches; }
    XD_CUDA(cudaMemcpyAsync(d->B + (size_t)ns * ni, rhs, (size_t)ni * sizeof(double), cudaMemcpyDeviceToDevice, s));
    for (int c = 0; c <= ns; ++c) {
      if (c < ns && !d->col_nonzero[c]) continue;      // a border column without interior entries: A_ii^-1 0 = 0 (e.g. a source branch)
      const int rc = xgpu_lu_solve(ctx, J, d->B + (size_t)c * ni, d->B + (size_t)c * ni);
      if (rc) return rc;
    }
  }
  if (ns == 0) {
    if (ni > 0) XD_CUDA(cudaMemcpyAsync(x, d->B, (size_t)ni * sizeof(double), cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  // this rank's part of the reduced system
  if (d->n_chunks > 0) {
    xb::launch_pdl(si_chunk_k, dim3(d->n_chunks), dim3(256), 0, s, ns + 1, d->n_chunks, (const int *)d->chunk_begin, (const int *)d->chunk_end,
                   (const int *)d->si_pos, (const int *)d->si_col, J, (const double *)d->B, ni, d->partials);
    ++ctx->launches;
  }
  const int take_b = (!rhs_border_reduced || d->rank == 0 || !xg_dist_multi(ctx)) ? 1 : 0;
  xb::launch_pdl(si_finish_k, dim3((ns * (ns + 1) + 63) / 64), dim3(64), 0, s, ns, d->n_chunks, (const int *)d->row_chunk_ptr,
                 (const double *)d->partials, (const int *)d->ss_pos, J, rhs, ni, take_b, d->red);
  ++ctx->launches;
  if (p2p_ready(d) && ns * (ns + 1) <= kBoxDoubles) {
    xb::launch_pdl(p2p_exchange_k, dim3(1), dim3(64), 0, s, p2p_view(d), ++d->p2p_epoch, ns * (ns + 1), 0, (const double *)d->red, d->red,
                   (double *const *)nullptr, 0, 1, 0);
    ++ctx->launches;
  } else if (d->comm && d->world > 1) {
    const int rc = nccl().AllReduce(d->red, d->red, (size_t)ns * (ns + 1), kNcclFloat64, kNcclSum, d->comm, s);
    if (rc) return nccl_fail(ctx, rc, "ncclAllReduce (border system)");
  }
  xb::launch_pdl(dense_solve_k, dim3(1), dim3(256), (size_t)ns * (ns + 1) * sizeof(double), s, ns, d->red, ctx->lu_dev.status);
  xb::launch_pdl(back_subst_k, dim3((ni + ns + 255) / 256), dim3(256), 0, s, ni, ns, (const double *)d->B, (const double *)d->red, x);
  ctx->launches += 2;
  XD_CUDA(cudaGetLastError());
  return 0;
}

extern "C" {

int xgpu_comm_unique_id(unsigned char *id128) {
  if (!id128) return 1;
  if (!nccl().ok()) return 2;
  NcclId id;
  const int rc = nccl().GetUniqueId(&id);
  if (rc) return 300 + rc;
  std::memcpy(id128, id.internal, 128);
  return 0;
}

int xgpu_comm_init(xgpu_ctx *ctx, const unsigned char *id128, int rank, int world) {
  if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return 1;
  if (!nccl().ok()) return xg_fail(ctx, 2, "libnccl.so.2 could not be loaded");
  XD_CUDA(cudaSetDevice(ctx->device));
  XgDist *d = ensure(ctx);
  if (d->comm) { nccl().CommDestroy(d->comm); d->comm = nullptr; }
  NcclId id;
  std::memcpy(id.internal, id128, 128);
  const int rc = nccl().CommInitRank(&d->comm, world, id, rank);
  if (rc) return nccl_fail(ctx, rc, "ncclCommInitRank");
  d->rank = rank; d->world = world;
  cudaFreeHost(d->h_pack); d->h_pack = nullptr;
  XD_CUDA(cudaMallocHost((void **)&d->h_pack, (size_t)(8 * world + 64) * sizeof(double)));
  if (d->pack) {      // border already declared: make room for the gathers of this world size
    cudaFree(d->pack); d->pack = nullptr;
    XD_CUDA(cudaMalloc((void **)&d->pack, (size_t)(8 * d->ns + 8 + 8 * world + 64) * sizeof(double)));
  }
  return 0;
}

// ---- peer-memory mailboxes (CUDA IPC): xgpu_p2p_handle on every rank, exchange the 64-byte handles through the host
// application, xgpu_p2p_attach with all of them (rank order).  Needs xgpu_comm_init first (rank / world).
int xgpu_p2p_handle(xgpu_ctx *ctx, unsigned char *handle64) {
  if (!ctx || !handle64) return 1;
  XgDist *d = ctx->dist;
  if (!d || d->world < 1) return xg_fail(ctx, 1, "xgpu_comm_init must precede xgpu_p2p_handle");
  if (d->world > 16) return xg_fail(ctx, 1, "peer mailboxes support at most 16 ranks");
  XD_CUDA(cudaSetDevice(ctx->device));
  if (!d->p2p_own) {
    const size_t bytes = (size_t)2 * d->world * kBoxDoubles * sizeof(double) + (size_t)(d->world + 2) * sizeof(unsigned long long);
    XD_CUDA(cudaMalloc((void **)&d->p2p_own, bytes));
    XD_CUDA(cudaMemset(d->p2p_own, 0, bytes));
    XD_CUDA(cudaMalloc((void **)&d->p2p_vecs, 4 * sizeof(double *)));
    XD_CUDA(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  XD_CUDA(cudaIpcGetMemHandle(&h, d->p2p_own));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
  std::memcpy(handle64, &h, 64);
  return 0;
}
int xgpu_p2p_attach(xgpu_ctx *ctx, const unsigned char *handles64_by_rank) {
  if (!ctx || !handles64_by_rank) return 1;
  XgDist *d = ctx->dist;
  if (!d || !d->p2p_own) return xg_fail(ctx, 1, "xgpu_p2p_handle must precede xgpu_p2p_attach");
  XD_CUDA(cudaSetDevice(ctx->device));
  for (int r = 0; r < d->world; ++r) {
    if (r == d->rank) { d->p2p_box[r] = d->p2p_own; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles64_by_rank + 64 * (size_t)r, 64);
    void *p = nullptr;
    XD_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    d->p2p_box[r] = (double *)p;
  }
  d->p2p_attached = true;
  d->p2p_epoch = 0;
  return 0;
}
// 1 when a mailbox spin timed out since the last call (a peer did not take part in a collective)
int xgpu_p2p_error(xgpu_ctx *ctx) {
  if (!ctx || !ctx->dist || !ctx->dist->p2p_own) return 0;
  int e = 0;
  const XgDist *d = ctx->dist;
  cudaMemcpy(&e, reinterpret_cast<const unsigned long long *>(d->p2p_own + (size_t)2 * d->world * kBoxDoubles) + d->world, sizeof(int), cudaMemcpyDeviceToHost);
  return e;
}

int xgpu_comm_info(const xgpu_ctx *ctx, int *rank, int *world) {
  if (!ctx || !rank || !world) return 1;
  *rank = ctx->dist ? ctx->dist->rank : 0;
  *world = ctx->dist ? ctx->dist->world : 1;
  return 0;
}

int xgpu_border_set(xgpu_ctx *ctx, int n_border) {
  if (!ctx || n_border < 0) return 1;
  if (!ctx->finalized) return xg_fail(ctx, 15, "xgpu_finalize has not been called");
  if (n_border > ctx->n) return xg_fail(ctx, 1, "more border unknowns than unknowns");
  if (n_border > 96) return xg_fail(ctx, 1, "the replicated border system is limited to 96 unknowns");
  XD_CUDA(cudaSetDevice(ctx->device));
  XgDist *d = ensure(ctx);
  const int n = ctx->n, ns = n_border, ni = n - ns;
  d->ni = ni; d->ns = ns; d->analyzed = false; d->n_global = ni + ns;
  std::vector<int> is_pos, is_row, is_col, si_pos, si_col, cb, ce, crow, rcp(1, 0), ss(std::max(ns * ns, 1), -1);
  d->sub_rowptr.assign(1, 0); d->sub_colind.clear(); d->sub_index.clear();
  for (int r = 0; r < n; ++r) {
    const int first_si = (int)si_pos.size();
    for (int p = ctx->rowptr[r]; p < ctx->rowptr[r + 1]; ++p) {
      const int c = ctx->colind[p];
      if (r < ni && c < ni) { d->sub_colind.push_back(c); d->sub_index.push_back(p); }
      else if (r < ni) { is_pos.push_back(p); is_row.push_back(r); is_col.push_back(c - ni); }
      else if (c < ni) { si_pos.push_back(p); si_col.push_back(c); }
      else ss[(r - ni) * ns + (c - ni)] = p;
    }
    if (r < ni) d->sub_rowptr.push_back((int32_t)d->sub_colind.size());
    else {
      for (int b = first_si; b < (int)si_pos.size(); b += kSiChunk) { cb.push_back(b); ce.push_back(std::min(b + kSiChunk, (int)si_pos.size())); crow.push_back(r - ni); }
      rcp.push_back((int)cb.size());
    }
  }
  while ((int)rcp.size() < ns + 1) rcp.push_back((int)cb.size());
  d->n_is = (int)is_pos.size(); d->n_si = (int)si_pos.size(); d->n_chunks = (int)cb.size();
  d->col_nonzero.assign(ns, 0);
  for (int c : is_col) d->col_nonzero[c] = 1;
  XD_CUDA(up(&d->is_pos, is_pos)); XD_CUDA(up(&d->is_row, is_row)); XD_CUDA(up(&d->is_col, is_col));
  XD_CUDA(up(&d->si_pos, si_pos)); XD_CUDA(up(&d->si_col, si_col));
  XD_CUDA(up(&d->chunk_begin, cb)); XD_CUDA(up(&d->chunk_end, ce)); XD_CUDA(up(&d->chunk_row, crow)); XD_CUDA(up(&d->row_chunk_ptr, rcp));
  XD_CUDA(up(&d->ss_pos, ss));
  cudaFree(d->B); cudaFree(d->partials); cudaFree(d->red); cudaFree(d->pack);
  d->B = d->partials = d->red = d->pack = nullptr;
  XD_CUDA(cudaMalloc((void **)&d->B, std::max<size_t>((size_t)(ns + 1) * ni, 1) * sizeof(double)));
  XD_CUDA(cudaMalloc((void **)&d->partials, std::max<size_t>((size_t)(ns + 1) * d->n_chunks, 1) * sizeof(double)));
  XD_CUDA(cudaMalloc((void **)&d->red, std::max<size_t>((size_t)ns * (ns + 1), 1) * sizeof(double)));
  XD_CUDA(cudaMalloc((void **)&d->pack, (size_t)(8 * ns + 8 + 8 * d->world + 64) * sizeof(double)));
  if (!d->h_pack) XD_CUDA(cudaMallocHost((void **)&d->h_pack, (size_t)(8 * d->world + 64) * sizeof(double)));
  if ((size_t)ns * (ns + 1) * sizeof(double) > 48 * 1024)
    XD_CUDA(cudaFuncSetAttribute(dense_solve_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)ns * (ns + 1) * sizeof(double))));
  ctx->lu_ready = false;      // a whole-matrix plan, if any, no longer matches
  // size of the global system (norm denominators of the step control): sum of the interiors + the border once
  if (d->comm && d->world > 1) {
    double h = (double)ni, *dv = d->pack;
    XD_CUDA(cudaMemcpyAsync(dv, &h, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const int rc = nccl().AllReduce(dv, dv, 1, kNcclFloat64, kNcclSum, d->comm, ctx->stream);
    if (rc) return nccl_fail(ctx, rc, "ncclAllReduce (sizes)");
    XD_CUDA(cudaMemcpyAsync(&h, dv, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    XD_CUDA(cudaStreamSynchronize(ctx->stream));
    d->n_global = (long long)(h + 0.5) + ns;
  }
  return 0;
}

int xgpu_border_info(const xgpu_ctx *ctx, int *n_interior, int *n_border, long long *n_global) {
  if (!ctx || !ctx->dist) return 1;
  if (n_interior) *n_interior = ctx->dist->ni;
  if (n_border) *n_border = ctx->dist->ns;
  if (n_global) *n_global = ctx->dist->n_global;
  return 0;
}

int xgpu_border_analyze(xgpu_ctx *ctx, const double *d_vals) {
  if (!ctx || !d_vals) return 100;
  XgDist *d = ctx->dist;
  if (!d) return xg_fail(ctx, 113, "xgpu_border_set has not been called");
  XD_CUDA(cudaSetDevice(ctx->device));
  if (d->ni == 0) { d->analyzed = true; return 0; }
  std::vector<double> vals((size_t)ctx->nnz);
  XD_CUDA(cudaMemcpyAsync(vals.data(), d_vals, vals.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  XD_CUDA(cudaStreamSynchronize(ctx->stream));
  xb::lu::set_batching(ctx->lu_batch != 0);
  const int rc = xb::lu::analyze_and_factor(d->ni, d->sub_rowptr.data(), d->sub_colind.data(), vals.data(), 0.001, ctx->lu_plan,
                                            d->sub_index.data());
  if (rc == 1) return xg_fail(ctx, 1, "the interior block is structurally singular");
  for (xgpu_ctx::LuGraph *g : {&ctx->g_refactor, &ctx->g_solve}) { if (g->exec) cudaGraphExecDestroy(g->exec); *g = xgpu_ctx::LuGraph(); }
  XD_CUDA(xb::lu::upload_plan(ctx->lu_plan, ctx->lu_dev));
  ctx->lu_ready = true;
  d->analyzed = true;
  if (rc == 2) return xg_fail(ctx, 2, "the interior block is numerically singular");
  return 0;
}

int xgpu_border_solve(xgpu_ctx *ctx, const double *d_vals, const double *d_rhs, double *d_x, int rhs_border_reduced) {
  if (!ctx || !d_vals || !d_rhs || !d_x) return 100;
  return xg_border_solve(ctx, d_vals, d_rhs, d_x, rhs_border_reduced, false);
}

int xgpu_shared_reduce(xgpu_ctx *ctx, double *d_f, double *d_q, double *d_dFdxdVp, double *d_dQdxdVp) {
  if (!ctx) return 1;
  double *v[4]; int n = 0;
  for (double *p : {d_f, d_q, d_dFdxdVp, d_dQdxdVp}) if (p) v[n++] = p;
  return xg_dist_reduce_border_rows(ctx, v, n);
}

}  // extern "C"
