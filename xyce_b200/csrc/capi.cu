// xyce_b200 -- implementation of the C ABI declared in include/xyce_b200.h.
// Host-side C++: owns device copies of the parameter records and gather maps, builds the
// stamp -> CSR maps once, and launches the sm_100a kernels.  There is deliberately no CPU
// fallback: every entry point fails with an error code if CUDA is unavailable.

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "ctx.h"
#include "pdl.cuh"
#include "vecops.cuh"

using namespace xb;
using namespace xb::b4;

namespace {

int fail(xgpu_ctx *c, int code, const std::string &msg) { return xg_fail(c, code, msg); }
#define XG_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      return fail(ctx, 100 + (int)e_, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

template <class T>
cudaError_t upload(T **dst, const T *src, size_t count) {
  cudaError_t e = cudaMalloc((void **)dst, std::max<size_t>(count, 1) * sizeof(T));
  if (e != cudaSuccess) return e;
  if (count) e = cudaMemcpy(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice);
  return e;
}

// first-part length of a run of `count` instances (pipelined host path)
inline int xg_pipe_cut(const xgpu_ctx *ctx, int count) {
  int cut = (int)(ctx->pipe_frac * count + 0.5);
  return cut < 0 ? 0 : (cut > count ? count : cut);
}

cudaError_t upload_map(const GatherMapHost &h, GatherMapDev &d) {
  d.ndst = (int)h.ptr.size() - 1;
  d.total = (int64_t)h.src.size();
  d.nlong = (int)h.long_dst.size();
  cudaError_t e;
  if ((e = upload(&d.ptr, h.ptr.data(), h.ptr.size())) != cudaSuccess) return e;
  if ((e = upload(&d.src, h.src.data(), h.src.size())) != cudaSuccess) return e;
  if ((e = upload(&d.long_dst, h.long_dst.data(), h.long_dst.size())) != cudaSuccess) return e;
  std::vector<int64_t> cb;
  std::vector<int32_t> cslot, lptr(1, 0);
  for (size_t j = 0; j < h.long_dst.size(); ++j) {
    const int dst = h.long_dst[j];
    for (int64_t b = h.ptr[dst]; b < h.ptr[dst + 1]; b += kChunk) { cb.push_back(b); cslot.push_back((int32_t)j); }
    lptr.push_back((int32_t)cb.size());
  }
  d.nchunks = (int)cb.size();
  if ((e = upload(&d.chunk_begin, cb.data(), cb.size())) != cudaSuccess) return e;
  if ((e = upload(&d.chunk_dst_slot, cslot.data(), cslot.size())) != cudaSuccess) return e;
  if ((e = upload(&d.long_chunk_ptr, lptr.data(), lptr.size())) != cudaSuccess) return e;
  // ELL width: 2 when (almost) every short destination has at most 2 sources, else 4
  {
    size_t over2 = 0;
    for (int dd = 0; dd < d.ndst; ++dd) { const int64_t c = h.ptr[dd + 1] - h.ptr[dd]; if (c > 2 && c <= kLongThreshold) ++over2; }
    d.ell_w = (over2 * 50 <= (size_t)d.ndst) ? 2 : kEllMax;
    std::vector<int32_t> ell((size_t)d.ell_w * std::max(d.ndst, 1), -1);
    for (int dd = 0; dd < d.ndst; ++dd) {
      const int64_t b = h.ptr[dd], c = h.ptr[dd + 1] - b;
      if (c > kLongThreshold) { ell[dd] = kEllLong; continue; }
      for (int k = 0; k < d.ell_w && k < c; ++k) ell[(size_t)k * d.ndst + dd] = h.src[b + k];
      if (c > d.ell_w) ell[(size_t)(d.ell_w - 1) * d.ndst + dd] = kEllTail;
    }
    if ((e = upload(&d.ell, ell.data(), ell.size())) != cudaSuccess) return e;
  }
  if ((e = cudaMalloc((void **)&d.partials, std::max<size_t>(4 * cb.size(), 1) * sizeof(double))) != cudaSuccess) return e;
  if ((e = cudaMalloc((void **)&d.done, std::max<size_t>(h.long_dst.size(), 1) * sizeof(int32_t))) != cudaSuccess) return e;
  return cudaMemset(d.done, 0, std::max<size_t>(h.long_dst.size(), 1) * sizeof(int32_t));
}

// uniform view of one device group for pattern / map construction
struct GroupView {
  int n, R, S;
  const int32_t *lids;            // node-major [nodes][n]
  std::vector<int> srow, scol;    // node index of each slot's row / column
  int64_t *vec_base, *mat_base;
};
std::vector<GroupView> group_views(xgpu_ctx *ctx) {
  std::vector<GroupView> v;
  for (auto &g : ctx->groups) {
    GroupView w;
    w.n = g.n; w.R = g.general ? kRowsGeneral : kRowsDefault; w.S = g.general ? kSlotsGeneral : kSlotsDefault;
    w.lids = g.lids.data();
    for (int s = 0; s < w.S; ++s) { w.srow.push_back(g.general ? kSlotRow[s] : s / 4); w.scol.push_back(g.general ? kSlotCol[s] : s % 4); }
    w.vec_base = (int64_t *)&g.dev.vec_base; w.mat_base = (int64_t *)&g.dev.mat_base;
    v.push_back(w);
  }
  for (auto &g : ctx->sgroups) {
    GroupView w;
    w.n = g.n; w.R = g.nodes; w.S = g.slots; w.lids = g.lids.data(); w.srow = g.slot_row; w.scol = g.slot_col;
    w.vec_base = (int64_t *)&g.dev.vec_base; w.mat_base = (int64_t *)&g.dev.mat_base;
    v.push_back(w);
  }
  return v;
}

void finish_map(GatherMapHost &m, const std::vector<int64_t> &count) {
  const size_t nd = count.size();
  m.ptr.assign(nd + 1, 0);
  for (size_t d = 0; d < nd; ++d) m.ptr[d + 1] = m.ptr[d] + count[d];
  m.src.assign((size_t)m.ptr[nd], 0);
  m.long_dst.clear();
  for (size_t d = 0; d < nd; ++d)
    if (count[d] > kLongThreshold) m.long_dst.push_back((int32_t)d);
}

}  // namespace

int xg_fail(xgpu_ctx *c, int code, const std::string &msg) {
  if (c) c->err = msg;
  return code;
}

int xg_finalize_linear(xgpu_ctx *ctx);

extern "C" {

int xgpu_create(int device, xgpu_ctx **out) {
  if (!out) return 1;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) return 2;   // no CUDA device: no fallback by design
  if (device < 0 || device >= ndev) return 3;
  xgpu_ctx *ctx = new xgpu_ctx;
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreate(&ctx->stream) != cudaSuccess) {
    delete ctx;
    return 4;
  }
  ctx->own_stream = true;
  *out = ctx;
  return 0;
}

void xgpu_destroy(xgpu_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto &g : ctx->groups) {
    cudaFree(g.d_inst_d); cudaFree(g.d_von); cudaFree(g.d_topo); cudaFree(g.d_model_idx);
    cudaFree(g.d_size_idx); cudaFree(g.d_lids); cudaFree(g.d_sto0); cudaFree(g.d_sta0); cudaFree(g.d_orig); cudaFree(g.d_branch0); cudaFree(g.d_lead);
  }
  for (auto &g : ctx->sgroups) { cudaFree(g.d_rec); cudaFree(g.d_flags); cudaFree(g.d_lids); cudaFree(g.d_sto0); cudaFree(g.d_sta0); cudaFree(g.d_orig); cudaFree(g.d_branch0); cudaFree(g.d_lead); }
  cudaFree(ctx->d_models); cudaFree(ctx->d_sizes); cudaFree(ctx->d_vec_planes); cudaFree(ctx->d_mat_planes);
  for (GatherMapDev *m : {&ctx->vec_map, &ctx->mat_map}) { cudaFree(m->ptr); cudaFree(m->src); cudaFree(m->long_dst); cudaFree(m->chunk_begin); cudaFree(m->chunk_dst_slot); cudaFree(m->long_chunk_ptr); cudaFree(m->partials); cudaFree(m->ell); cudaFree(m->done); }
  cudaFree(ctx->d_conv);
  for (double *b : ctx->buf) cudaFree(b);
  cudaFree(ctx->d_lead_host);
  for (xgpu_ctx::LuGraph *g : {&ctx->g_refactor, &ctx->g_solve}) if (g->exec) cudaGraphExecDestroy(g->exec);
  cudaFree(ctx->tran_pool); cudaFree(ctx->tran_ints); cudaFreeHost(ctx->tran_pinned); cudaFree(ctx->d_hist);
  if (ctx->pipe.s2) cudaStreamDestroy(ctx->pipe.s2);
  if (ctx->pipe.ev_a) cudaEventDestroy(ctx->pipe.ev_a);
  if (ctx->pipe.ev_x) cudaEventDestroy(ctx->pipe.ev_x);
  xb::lu::free_plan(ctx->lu_dev);
  xg_dist_free(ctx);
  for (XgLinearPart *L : {&ctx->linG, &ctx->linC}) { cudaFree(L->rows); cudaFree(L->ptr); cudaFree(L->col); cudaFree(L->pos); cudaFree(L->val); }
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char *xgpu_last_error(const xgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int xgpu_set_stream(xgpu_ctx *ctx, void *s) {
  if (!ctx) return 1;
  if (ctx->own_stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
  ctx->stream = (cudaStream_t)s;
  return 0;
}

int xgpu_set_option(xgpu_ctx *ctx, const char *name, int value) {
  if (!ctx || !name) return 1;
  const std::string n(name);
  if (n == "b4_arith" && value >= 0 && value <= 2) { ctx->b4_arith = value; return 0; }
  if (n == "b4_minblocks" && value >= 1 && value <= 6) { ctx->b4_minblocks = value; return 0; }
  if (n == "b4_threads" && (value == 0 || (value >= 64 && value <= 512))) { ctx->b4_threads = value; return 0; }
  if (n == "b4_uniform" && (value == 0 || value == 1)) { ctx->b4_uniform = value; return 0; }
  if (n == "b4_lockstep" && (value == 0 || value == 1)) { ctx->b4_lockstep = value; return 0; }
  if (n == "b4_spec" && (value == 0 || value == 1)) { ctx->b4_spec = value; return 0; }
  // "pipeline_host": 1 (default) = xgpu_load_host_jr evaluates in two parts and overlaps the PCIe transfer of the first
  // window with the evaluation of the second part when the circuit's numbering allows it; "pipe_percent" = share of every
  // run in the first part (before xgpu_finalize)
  if (n == "pipeline_host" && (value == 0 || value == 1)) { ctx->pipeline_host = value; return 0; }
  if (n == "pipe_r_mapped" && (value == 0 || value == 1)) { ctx->pipe_r_mapped = value; return 0; }
  if (n == "pipe_percent" && value >= 10 && value <= 90 && !ctx->finalized) { ctx->pipe_frac = value / 100.0; return 0; }
  if (n == "lu_graphs" && (value == 0 || value == 1)) { ctx->lu_graphs = value; return 0; }
  // "lu_pivot_check": 1 (default) = the refactorization tests every pivot of the fixed sequence against KLU's threshold
  // (|pivot| >= 0.001 max |column|) and reports code 3 when one fails; "lu_repivot": 1 = KLU_REPIVOT=1 semantics, every
  // xgpu_lu_refactor re-runs the pivoting host factorization (the reference's default, N_LAS_AmesosSolver.C:316-318)
  if (n == "lu_pivot_check" && (value == 0 || value == 1)) {
    ctx->lu_dev.pivot_check = value;
    // captured refactor graphs carry the old value in their kernel parameters: drop them
    for (xgpu_ctx::LuGraph *g : {&ctx->g_refactor, &ctx->g_solve}) { if (g->exec) cudaGraphExecDestroy(g->exec); *g = xgpu_ctx::LuGraph(); }
    return 0;
  }
  if (n == "lu_batch" && (value == 0 || value == 1)) { ctx->lu_batch = value; return 0; }      // takes effect at the next analysis / import
  if (n == "lu_repivot" && (value == 0 || value == 1)) { ctx->lu_repivot = value; return 0; }
  if (n == "zero_copy_out" && (value == 0 || value == 1)) { ctx->zero_copy_out = value; return 0; }
  return fail(ctx, 16, "unknown option or value out of range: " + n);
}

int xgpu_sync(xgpu_ctx *ctx) {
  if (!ctx) return 1;
  XG_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int xgpu_pattern_set(xgpu_ctx *ctx, int n, const int32_t *rowptr, const int32_t *colind) {
  if (!ctx || n < 0 || !rowptr) return 1;
  if (ctx->finalized) return fail(ctx, 5, "pattern_set after finalize");
  ctx->n = n;
  ctx->rowptr.assign(rowptr, rowptr + n + 1);
  ctx->nnz = rowptr[n];
  ctx->colind.assign(colind, colind + ctx->nnz);
  for (int r = 0; r < n; ++r)
    for (int k = rowptr[r] + 1; k < rowptr[r + 1]; ++k)
      if (colind[k] <= colind[k - 1]) return fail(ctx, 6, "CSR columns must be strictly increasing within a row");
  return 0;
}

int xgpu_pattern_build(xgpu_ctx *ctx, int n) {
  if (!ctx || n <= 0) return 1;
  if (ctx->finalized) return fail(ctx, 5, "pattern_build after finalize");
  // (row, col) pairs of every stamp entry; sort + unique = generateRowColData
  std::vector<uint64_t> pairs;
  for (auto &g : group_views(ctx)) {
    const int gn = g.n;
    pairs.reserve(pairs.size() + (size_t)g.S * gn);
    for (int s = 0; s < g.S; ++s) {
      const int32_t *lr = g.lids + (size_t)g.srow[s] * gn, *lc = g.lids + (size_t)g.scol[s] * gn;
      for (int i = 0; i < gn; ++i) {
        if (lr[i] < 0 || lc[i] < 0) continue;
        if (lr[i] >= n || lc[i] >= n) return fail(ctx, 10, "node LID outside the pattern");
        pairs.push_back(((uint64_t)(uint32_t)lr[i] << 32) | (uint32_t)lc[i]);
      }
    }
  }
  for (const XgLinearPart *L : {&ctx->linG, &ctx->linC})
    for (size_t k = 0; k < L->h_row.size(); ++k) {
      if (L->h_row[k] >= n || L->h_col[k] >= n) return fail(ctx, 10, "linear stamp index outside the pattern");
      pairs.push_back(((uint64_t)(uint32_t)L->h_row[k] << 32) | (uint32_t)L->h_col[k]);
    }
  std::sort(pairs.begin(), pairs.end());
  pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
  ctx->n = n;
  ctx->nnz = (int64_t)pairs.size();
  ctx->rowptr.assign(n + 1, 0);
  ctx->colind.resize(pairs.size());
  for (size_t k = 0; k < pairs.size(); ++k) {
    ++ctx->rowptr[(pairs[k] >> 32) + 1];
    ctx->colind[k] = (int32_t)(pairs[k] & 0xffffffffu);
  }
  for (int r = 0; r < n; ++r) ctx->rowptr[r + 1] += ctx->rowptr[r];
  return 0;
}

int xgpu_pattern_nnz(const xgpu_ctx *ctx) { return ctx ? (int)ctx->nnz : -1; }

int xgpu_pattern_get(const xgpu_ctx *ctx, int32_t *rowptr, int32_t *colind) {
  if (!ctx || !rowptr || !colind || ctx->rowptr.empty()) return 1;
  std::copy(ctx->rowptr.begin(), ctx->rowptr.end(), rowptr);
  std::copy(ctx->colind.begin(), ctx->colind.end(), colind);
  return 0;
}

int xgpu_sizes_set(xgpu_ctx *ctx, int n_state, int n_store) {
  if (!ctx || n_state < 0 || n_store < 0) return 1;
  ctx->n_state = n_state;
  ctx->n_store = n_store;
  return 0;
}

int xgpu_b4_field_count(int which) {
  switch (which) {
    case 0: return kNumModelD;
    case 1: return kNumModelI;
    case 2: return kNumSizeD;
    case 3: return kNumInstD;
    case 4: return kNumInstI;
  }
  return -1;
}

const char *xgpu_b4_field_names(int which) {
#define NM(n) #n "\n"
  switch (which) {
    case 0: return XB_B4_MODEL_D(NM);
    case 1: return XB_B4_MODEL_I(NM);
    case 2: return XB_B4_SIZE_D(NM);
    case 3: return XB_B4_INST_D(NM);
    case 4: return XB_B4_INST_I(NM);
  }
#undef NM
  return "";
}

int xgpu_b4_models_set(xgpu_ctx *ctx, int n_models, const double *model_d, const int32_t *model_i, int n_sizes,
                       const double *size_d) {
  if (!ctx || n_models <= 0 || n_sizes <= 0 || !model_d || !model_i || !size_d) return 1;
  XG_CUDA(cudaSetDevice(ctx->device));
  std::vector<B4Model> M(n_models);
  std::vector<B4Size> P(n_sizes);
  for (int m = 0; m < n_models; ++m) {
    int k = 0;
#define GET(name) M[m].name = model_d[(size_t)m * kNumModelD + (k++)];
    XB_B4_MODEL_D(GET)
#undef GET
    k = 0;
#define GET(name) M[m].name = model_i[(size_t)m * kNumModelI + (k++)];
    XB_B4_MODEL_I(GET)
#undef GET
  }
  for (int s = 0; s < n_sizes; ++s) {
    int k = 0;
#define GET(name) P[s].name = size_d[(size_t)s * kNumSizeD + (k++)];
    XB_B4_SIZE_D(GET)
#undef GET
  }
  // only the two record tables are replaced; device groups (BSIM4 and small-device) keep their buffers
  XG_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(ctx->d_models); cudaFree(ctx->d_sizes);
  ctx->d_models = nullptr; ctx->d_sizes = nullptr;
  XG_CUDA(upload(&ctx->d_models, M.data(), M.size()));
  XG_CUDA(upload(&ctx->d_sizes, P.data(), P.size()));
  ctx->n_models = n_models;
  ctx->n_sizes = n_sizes;
  ctx->h_models = M;
  ctx->h_sizes = P;
  for (auto &g : ctx->groups) { g.dev.models = ctx->d_models; g.dev.sizes = ctx->d_sizes; g.packs_valid = false; }
  return 0;
}

int xgpu_b4_group_add(xgpu_ctx *ctx, int n, const double *inst_d, const int32_t *inst_i, const int32_t *model_idx,
                      const int32_t *size_idx, const int32_t *lids12, const int32_t *sto_lid0, int sto_stride,
                      const int32_t *sta_lid0, int sta_stride) {
  if (!ctx || n <= 0 || !inst_d || !inst_i || !model_idx || !size_idx || !lids12 || !sto_lid0 || !sta_lid0) return -1;
  if (ctx->finalized) { fail(ctx, 5, "group_add after finalize"); return -5; }
  if (!ctx->d_models) { fail(ctx, 7, "xgpu_b4_models_set must precede xgpu_b4_group_add"); return -7; }
  if (cudaSetDevice(ctx->device) != cudaSuccess) return -4;
  XgHostGroup g;
  g.n = n;
  // records -> structure of arrays; topology word; choose the kernel variant
  std::vector<double> soa((size_t)kNumInstD * n);
  std::vector<int> topo(n);
  bool all_default = true;
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < kNumInstD; ++k) soa[(size_t)k * n + i] = inst_d[(size_t)i * kNumInstD + k];
    B4Inst I{};
    int k = 0;
#define GET(name) I.name = inst_i[(size_t)i * kNumInstI + (k++)];
    XB_B4_INST_I(GET)
#undef GET
    if (I.trnqsMod) { fail(ctx, 8, "trnqsMod = 1 is not supported (the reference marks it unimplemented)"); return -8; }
    topo[i] = pack_topo(I);
    if (topo[i] & kDefaultTopoMask) all_default = false;
    if (model_idx[i] < 0 || model_idx[i] >= ctx->n_models || size_idx[i] < 0 || size_idx[i] >= ctx->n_sizes) {
      fail(ctx, 9, "model/size index out of range");
      return -9;
    }
  }
  g.general = all_default ? 0 : 1;
  for (int i = 0; i < n; ++i) {
    if (i > 0 && model_idx[i] == model_idx[i - 1] && size_idx[i] == size_idx[i - 1]) { ++g.run_count.back(); continue; }
    g.run_model.push_back(model_idx[i]); g.run_size.push_back(size_idx[i]); g.run_start.push_back(i); g.run_count.push_back(1);
  }
  const int nn = g.general ? kNumNodes : 4;
  g.lids.assign((size_t)kNumNodes * n, -1);
  for (int i = 0; i < n; ++i)
    for (int t = 0; t < kNumNodes; ++t) {
      const int l = lids12[(size_t)i * kNumNodes + t];
      if (l >= ctx->n && ctx->n > 0 && !ctx->rowptr.empty()) { fail(ctx, 10, "node LID outside the pattern"); return -10; }
      g.lids[(size_t)t * n + i] = l;
    }
  std::vector<double> von(n, 0.0);
  std::vector<int> orig(n, 1);
  bool ok = upload(&g.d_inst_d, soa.data(), soa.size()) == cudaSuccess &&
            upload(&g.d_topo, topo.data(), topo.size()) == cudaSuccess &&
            upload(&g.d_model_idx, model_idx, (size_t)n) == cudaSuccess &&
            upload(&g.d_size_idx, size_idx, (size_t)n) == cudaSuccess &&
            upload(&g.d_lids, g.lids.data(), (size_t)nn * n) == cudaSuccess &&
            upload(&g.d_sto0, sto_lid0, (size_t)n) == cudaSuccess &&
            upload(&g.d_sta0, sta_lid0, (size_t)n) == cudaSuccess &&
            upload(&g.d_von, von.data(), (size_t)n) == cudaSuccess &&
            upload(&g.d_orig, orig.data(), (size_t)n) == cudaSuccess;
  if (!ok) { fail(ctx, 11, "device allocation failed in group_add"); return -11; }
  GroupDev &d = g.dev;
  d.n = n; d.general = g.general; d.models = ctx->d_models; d.sizes = ctx->d_sizes;
  d.inst_d = g.d_inst_d; d.topo = g.d_topo; d.model_idx = g.d_model_idx; d.size_idx = g.d_size_idx;
  d.lids = g.d_lids; d.sto_lid0 = g.d_sto0; d.sta_lid0 = g.d_sta0; d.sto_stride = sto_stride;
  d.sta_stride = sta_stride; d.von = g.d_von; d.orig_flag = g.d_orig;
  ctx->groups.push_back(std::move(g));
  return (int)ctx->groups.size() - 1;
}

int xgpu_simple_group_add(xgpu_ctx *ctx, int type, int n, const double *rec, const int32_t *flags, const int32_t *lids,
                          const int32_t *sto_lid0, int sto_stride, const int32_t *sta_lid0, int sta_stride) {
  if (!ctx || n <= 0 || !rec || !lids) return -1;
  if (ctx->finalized) { fail(ctx, 5, "group_add after finalize"); return -5; }
  const xb::simple::TypeInfo *ti = xb::simple::type_info(type);
  if (!ti) { fail(ctx, 17, "unknown device type"); return -17; }
  if (cudaSetDevice(ctx->device) != cudaSuccess) return -4;
  XgSimpleGroup g;
  g.type = type; g.n = n; g.nodes = ti->nodes; g.slots = ti->slots; g.nfields = ti->nfields; g.nstore = ti->nstore; g.nstate = ti->nstate;
  g.slot_row.assign(ti->slot_row, ti->slot_row + ti->slots);
  g.slot_col.assign(ti->slot_col, ti->slot_col + ti->slots);
  std::vector<double> soa((size_t)g.nfields * n);
  for (int i = 0; i < n; ++i) for (int k = 0; k < g.nfields; ++k) soa[(size_t)k * n + i] = rec[(size_t)i * g.nfields + k];
  g.lids.assign((size_t)g.nodes * n, -1);
  for (int i = 0; i < n; ++i) for (int t = 0; t < g.nodes; ++t) g.lids[(size_t)t * n + i] = lids[(size_t)i * g.nodes + t];
  std::vector<int> fl(n, 0), zero(n, 0), one(n, 1);
  if (flags) fl.assign(flags, flags + n);
  bool ok = upload(&g.d_rec, soa.data(), soa.size()) == cudaSuccess && upload(&g.d_flags, fl.data(), fl.size()) == cudaSuccess &&
            upload(&g.d_lids, g.lids.data(), g.lids.size()) == cudaSuccess &&
            upload(&g.d_sto0, sto_lid0 ? sto_lid0 : zero.data(), (size_t)n) == cudaSuccess &&
            upload(&g.d_sta0, sta_lid0 ? sta_lid0 : zero.data(), (size_t)n) == cudaSuccess &&
            upload(&g.d_orig, one.data(), (size_t)n) == cudaSuccess;
  if (!ok) { fail(ctx, 11, "device allocation failed in simple_group_add"); return -11; }
  xb::simple::GroupDev &d = g.dev;
  d.type = type; d.n = n; d.rec = g.d_rec; d.flags = g.d_flags; d.lids = g.d_lids; d.sto_lid0 = g.d_sto0; d.sta_lid0 = g.d_sta0;
  d.sto_stride = sto_stride; d.sta_stride = sta_stride; d.orig_flag = g.d_orig;
  if (type == xb::simple::kBjt) {      // excess phase (model PTF != 0): the evaluation reads the store vector of the step before
    const size_t k = (size_t)xb::simple::bjt_excess_phase_field();
    for (int i = 0; i < n; ++i) if (rec[(size_t)i * g.nfields + k] != 0.0) { ctx->needs_last_sto = true; break; }
  }
  ctx->sgroups.push_back(std::move(g));
  return (int)ctx->sgroups.size() - 1;
}

int xgpu_needs_last_store(const xgpu_ctx *ctx) { return ctx && ctx->needs_last_sto ? 1 : 0; }

int xgpu_last_store_set(xgpu_ctx *ctx, double *d_last_sto) {
  if (!ctx) return 1;
  ctx->d_last_sto = d_last_sto;
  return 0;
}

int xgpu_simple_store_count(int type) {
  const xb::simple::TypeInfo *ti = xb::simple::type_info(type);
  return ti ? ti->nstore : -1;
}

int xgpu_simple_field_count(int type) {
  const xb::simple::TypeInfo *ti = xb::simple::type_info(type);
  return ti ? ti->nfields : -1;
}

// pipelined host-buffer path: {usable for this circuit, rows in the first window, nonzeros in the first window}
int xgpu_pipe_info(const xgpu_ctx *ctx, long long *info3) {
  if (!ctx || !info3) return 1;
  info3[0] = ctx->pipe.ok ? 1 : 0; info3[1] = ctx->pipe.vec_split; info3[2] = ctx->pipe.mat_split;
  return 0;
}

// mode-specialised kernel object the group's last evaluation ran (id of bsim4_spec_tuples.def), -1 = generic build
int xgpu_b4_group_spec(const xgpu_ctx *ctx, int group) {
  if (!ctx || group < 0 || group >= (int)ctx->groups.size()) return -2;
  return ctx->groups[group].last_spec;
}

int xgpu_adms_gen_count(void) { return xb::simple::adms_gen_count(); }
int xgpu_adms_gen_info(int idx, const char **name, const char **fields, int32_t *info5, int32_t *slot_row, int32_t *slot_col) {
  const xb::simple::TypeInfo *ti = xb::simple::type_info(xb::simple::kAdmsGenBase + idx);
  if (idx < 0 || idx >= xb::simple::adms_gen_count() || !ti) return 1;
  if (name) *name = xb::simple::adms_gen_name(idx);
  if (fields) *fields = xb::simple::adms_gen_fields(idx);
  if (info5) { info5[0] = xb::simple::kAdmsGenBase + idx; info5[1] = ti->nodes; info5[2] = xb::simple::adms_gen_ext(idx); info5[3] = ti->slots; info5[4] = ti->nfields; }
  for (int s = 0; s < ti->slots; ++s) { if (slot_row) slot_row[s] = ti->slot_row[s]; if (slot_col) slot_col[s] = ti->slot_col[s]; }
  return 0;
}

int xgpu_finalize(xgpu_ctx *ctx) {
  if (!ctx) return 1;
  if (ctx->finalized) return 0;
  if (ctx->rowptr.empty()) return fail(ctx, 12, "xgpu_pattern_set must precede xgpu_finalize");
  XG_CUDA(cudaSetDevice(ctx->device));
  // plane layout
  int64_t vb = 0, mb = 0;
  std::vector<GroupView> views = group_views(ctx);
  for (auto &g : views) {
    *g.vec_base = vb; *g.mat_base = mb;
    vb += (int64_t)g.R * g.n; mb += (int64_t)g.S * g.n;
  }
  if (vb >= (1LL << 31) || mb >= (1LL << 31)) return fail(ctx, 13, "contribution plane exceeds 2^31 elements");
  ctx->vec_plane = vb; ctx->mat_plane = mb;
  // gather maps: two passes (count, fill) in group-major, instance-minor, slot order
  const int n = ctx->n;
  GatherMapHost vm, mm;
  std::vector<int64_t> vcount(n, 0), mcount((size_t)ctx->nnz, 0);
  auto csr_find = [&](int r, int c) -> int64_t {
    const int32_t *b = ctx->colind.data() + ctx->rowptr[r], *e = ctx->colind.data() + ctx->rowptr[r + 1];
    const int32_t *p = std::lower_bound(b, e, c);
    return (p != e && *p == c) ? (int64_t)(p - ctx->colind.data()) : -1;
  };
  for (int pass = 0; pass < 2; ++pass) {
    std::vector<int64_t> vfill, mfill;
    if (pass == 1) {
      finish_map(vm, vcount); finish_map(mm, mcount);
      vfill.assign(vm.ptr.begin(), vm.ptr.end() - 1);
      mfill.assign(mm.ptr.begin(), mm.ptr.end() - 1);
    }
    for (auto &g : views) {
      const int gn = g.n;
      for (int i = 0; i < gn; ++i) {
        for (int r = 0; r < g.R; ++r) {
          const int l = g.lids[(size_t)r * gn + i];
          if (l < 0) continue;
          if (l >= n) return fail(ctx, 10, "node LID outside the pattern");
          if (pass == 0) ++vcount[l];
          else vm.src[(size_t)vfill[l]++] = (int32_t)(*g.vec_base + (int64_t)r * gn + i);
        }
        for (int s = 0; s < g.S; ++s) {
          const int lr = g.lids[(size_t)g.srow[s] * gn + i], lc = g.lids[(size_t)g.scol[s] * gn + i];
          if (lr < 0 || lc < 0) continue;
          const int64_t k = csr_find(lr, lc);
          if (k < 0) return fail(ctx, 14, "a device stamp entry is missing from the CSR pattern");
          if (pass == 0) ++mcount[(size_t)k];
          else mm.src[(size_t)mfill[(size_t)k]++] = (int32_t)(*g.mat_base + (int64_t)s * gn + i);
        }
      }
    }
  }
  XG_CUDA(upload_map(vm, ctx->vec_map));
  XG_CUDA(upload_map(mm, ctx->mat_map));
  // Pipelined host path: with ONE 4-terminal BSIM4 group of uniform runs, find the prefix of rows / nonzeros whose
  // contributions all come from the first part of the runs (and that are not long destinations).  Circuits numbered
  // device by device (arrays, anything a netlist lists block by block) have a long such prefix; others simply do not
  // pipeline.
  ctx->pipe.ok = false;
  if (ctx->groups.size() == 1 && ctx->sgroups.empty() && !ctx->groups[0].general &&
      (int)ctx->groups[0].run_start.size() <= kMaxUniformRuns && ctx->groups[0].n > 0) {
    const XgHostGroup &g = ctx->groups[0];
    std::vector<char> part1((size_t)g.n, 0);
    for (size_t r = 0; r < g.run_start.size(); ++r) {
      const int cut = xg_pipe_cut(ctx, g.run_count[r]);
      for (int k = cut; k < g.run_count[r]; ++k) part1[(size_t)g.run_start[r] + k] = 1;
    }
    auto prefix = [&](const GatherMapHost &h) -> int64_t {
      const int64_t nd = (int64_t)h.ptr.size() - 1;
      for (int64_t d = 0; d < nd; ++d) {
        if (h.ptr[d + 1] - h.ptr[d] > kLongThreshold) return d;
        for (int64_t k = h.ptr[d]; k < h.ptr[d + 1]; ++k) if (part1[(size_t)(h.src[k] % g.n)]) return d;      // plane element (row, i) = base + row * n + i, base = 0
      }
      return nd;
    };
    ctx->pipe.vec_split = (int)prefix(vm);
    ctx->pipe.mat_split = prefix(mm);
    ctx->pipe.ok = ctx->pipe.vec_split >= n / 4 && ctx->pipe.mat_split >= ctx->nnz / 4;
    if (ctx->pipe.ok && !ctx->pipe.s2) {
      XG_CUDA(cudaStreamCreateWithFlags(&ctx->pipe.s2, cudaStreamNonBlocking));
      XG_CUDA(cudaEventCreateWithFlags(&ctx->pipe.ev_a, cudaEventDisableTiming));
      XG_CUDA(cudaEventCreateWithFlags(&ctx->pipe.ev_x, cudaEventDisableTiming));
    }
  }
  XG_CUDA(cudaMalloc((void **)&ctx->d_vec_planes, std::max<int64_t>(4 * vb, 1) * sizeof(double)));
  XG_CUDA(cudaMalloc((void **)&ctx->d_mat_planes, std::max<int64_t>(2 * mb, 1) * sizeof(double)));
  XG_CUDA(cudaMemset(ctx->d_vec_planes, 0, std::max<int64_t>(4 * vb, 1) * sizeof(double)));
  XG_CUDA(cudaMemset(ctx->d_mat_planes, 0, std::max<int64_t>(2 * mb, 1) * sizeof(double)));
  XG_CUDA(cudaMalloc((void **)&ctx->d_conv, sizeof(int)));
  // context-owned system buffers
  const size_t sizes[12] = {(size_t)n, (size_t)n, (size_t)n, (size_t)n, (size_t)n, (size_t)ctx->nnz, (size_t)ctx->nnz,
                            (size_t)ctx->n_store, (size_t)ctx->n_store, (size_t)ctx->n_state, (size_t)ctx->n_state,
                            ctx->needs_last_sto ? (size_t)ctx->n_store : 0};
  for (int b = 0; b < 12; ++b) {
    XG_CUDA(cudaMalloc((void **)&ctx->buf[b], std::max<size_t>(sizes[b], 1) * sizeof(double)));
    XG_CUDA(cudaMemset(ctx->buf[b], 0, std::max<size_t>(sizes[b], 1) * sizeof(double)));
  }
  if (ctx->needs_last_sto && !ctx->d_last_sto) ctx->d_last_sto = ctx->buf[11];
  { const int rc = xg_finalize_linear(ctx); if (rc) return rc; }
  ctx->finalized = true;
  return 0;
}

namespace {
// Lead currents of BSIM4 instances.  Groups with internal nodes: the evaluation kernel has written the lead block
// (emit_lead, bsim4_load.h) to the group's [8][n] buffer.  4-terminal groups: with no internal nodes the lead quantities of
// Master::loadDAEVectors (N_DEV_MOSFET_B4.C:10933-10987) ARE the instance's own F and Q contributions to its
// drain / gate / source / bulk rows -- leadF[id] = -(ceqjd - ceqbd - ceqdrn + Idtoteq) np is the D' row term (:10691),
// leadQ[is] = -(Qg + Qb + Qd) np the S' row term (:10893), and so on -- which the evaluation kernel has just
// written to the contribution planes: this kernel copies them to the branch-data LIDs (assign, not accumulate).
__global__ void b4_lead_kernel(int n, const double *__restrict__ planeF, const double *__restrict__ planeQ,
                               const int *__restrict__ lids, const int *__restrict__ branch0,
                               const double *__restrict__ sol, double *leadF, double *leadQ, double *junctionV) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = branch0[i];
  if (b < 0) return;
  // plane rows: 0 drain, 1 gate, 2 source, 3 bulk; branch order: id, ig, is, ib
  const int row_of[4] = {0, 1, 2, 3};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    leadF[b + k] = planeF[(size_t)row_of[k] * n + i];
    leadQ[b + k] = planeQ[(size_t)row_of[k] * n + i];
  }
  const int ld = lids[i], lg = lids[(size_t)n + i], ls = lids[2 * (size_t)n + i];
  const double vd = ld >= 0 ? sol[ld] : 0.0, vg = lg >= 0 ? sol[lg] : 0.0, vs = ls >= 0 ? sol[ls] : 0.0;
  junctionV[b + 0] = vd - vs;
  junctionV[b + 1] = vg - vs;
  junctionV[b + 2] = 0.0;
  junctionV[b + 3] = 0.0;
}
}  // namespace

int xgpu_b4_lead_set(xgpu_ctx *ctx, int group, const int32_t *branch_lid0) {
  if (!ctx || group < 0 || group >= (int)ctx->groups.size() || !branch_lid0) return 1;
  XgHostGroup &g = ctx->groups[group];
  cudaFree(g.d_branch0); g.d_branch0 = nullptr;
  XG_CUDA(upload(&g.d_branch0, branch_lid0, (size_t)g.n));
  if (g.general && !g.d_lead) {      // devices with internal nodes: the evaluation kernel computes the lead block itself
    XG_CUDA(cudaMalloc((void **)&g.d_lead, (size_t)8 * std::max(g.n, 1) * sizeof(double)));
    XG_CUDA(cudaMemset(g.d_lead, 0, (size_t)8 * std::max(g.n, 1) * sizeof(double)));
    g.dev.lead = g.d_lead;
  }
  return 0;
}

int xgpu_simple_lead_set(xgpu_ctx *ctx, int group, const int32_t *branch_lid0) {
  if (!ctx || group < 0 || group >= (int)ctx->sgroups.size() || !branch_lid0) return 1;
  XgSimpleGroup &g = ctx->sgroups[group];
  if (xb::simple::lead_count(g.type) == 0) return fail(ctx, 17, "this device type has no lead currents in the library");
  cudaFree(g.d_branch0); g.d_branch0 = nullptr;
  XG_CUDA(upload(&g.d_branch0, branch_lid0, (size_t)g.n));
  if (g.type == xb::simple::kBjt && !g.d_lead) {
    XG_CUDA(cudaMalloc((void **)&g.d_lead, (size_t)8 * std::max(g.n, 1) * sizeof(double)));
    XG_CUDA(cudaMemset(g.d_lead, 0, (size_t)8 * std::max(g.n, 1) * sizeof(double)));
    g.dev.lead = g.d_lead;
  }
  return 0;
}

int xgpu_lead_load(xgpu_ctx *ctx, const double *d_sol, double *d_leadF, double *d_leadQ, double *d_junctionV) {
  return xgpu_b4_lead_load(ctx, d_sol, d_leadF, d_leadQ, d_junctionV);
}

int xgpu_b4_lead_load(xgpu_ctx *ctx, const double *d_sol, double *d_leadF, double *d_leadQ, double *d_junctionV) {
  if (!ctx || !d_sol || !d_leadF || !d_leadQ || !d_junctionV) return 1;
  if (!ctx->finalized) return fail(ctx, 15, "xgpu_finalize has not been called");
  for (auto &g : ctx->groups) {
    if (!g.d_branch0 || g.n == 0) continue;
    const double *pF = ctx->d_vec_planes + g.dev.vec_base, *pQ = ctx->d_vec_planes + ctx->vec_plane + g.dev.vec_base;
    if (g.general) { pF = g.d_lead; pQ = g.d_lead + (size_t)4 * g.n; }
    b4_lead_kernel<<<(g.n + 255) / 256, 256, 0, ctx->stream>>>(g.n, pF, pQ, g.d_lids, g.d_branch0, d_sol, d_leadF, d_leadQ, d_junctionV);
    ++ctx->launches;
  }
  for (auto &g : ctx->sgroups) {
    if (!g.d_branch0 || g.n == 0) continue;
    xb::simple::launch_lead(g.dev, g.d_branch0, ctx->d_vec_planes + g.dev.vec_base, ctx->d_vec_planes + ctx->vec_plane + g.dev.vec_base,
                            d_sol, d_leadF, d_leadQ, d_junctionV, ctx->stream);
    ++ctx->launches;
  }
  XG_CUDA(cudaGetLastError());
  return 0;
}

// host-buffer form for the adaptors: the three lead vectors go up (entries the kernels do not write keep their values, as in
// the reference), the lead kernels run against the solution of the last xgpu_load_host call, the vectors come back
int xgpu_lead_load_host(xgpu_ctx *ctx, int n_branch, double *h_leadF, double *h_leadQ, double *h_junctionV) {
  if (!ctx || n_branch <= 0 || !h_leadF || !h_leadQ || !h_junctionV) return 1;
  if (!ctx->finalized) return fail(ctx, 15, "xgpu_finalize has not been called");
  if (ctx->lead_len < n_branch) {
    cudaFree(ctx->d_lead_host); ctx->d_lead_host = nullptr; ctx->lead_len = 0;
    XG_CUDA(cudaMalloc((void **)&ctx->d_lead_host, (size_t)3 * n_branch * sizeof(double)));
    ctx->lead_len = n_branch;
  }
  double *d[3] = {ctx->d_lead_host, ctx->d_lead_host + n_branch, ctx->d_lead_host + 2 * (size_t)n_branch};
  double *h[3] = {h_leadF, h_leadQ, h_junctionV};
  for (int k = 0; k < 3; ++k) XG_CUDA(cudaMemcpyAsync(d[k], h[k], (size_t)n_branch * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  const int rc = xgpu_b4_lead_load(ctx, ctx->buf[0], d[0], d[1], d[2]);
  if (rc) return rc;
  for (int k = 0; k < 3; ++k) XG_CUDA(cudaMemcpyAsync(h[k], d[k], (size_t)n_branch * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  XG_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int xgpu_b4_von_set(xgpu_ctx *ctx, int group, const double *von) {
  if (!ctx || group < 0 || group >= (int)ctx->groups.size() || !von) return 1;
  XG_CUDA(cudaMemcpyAsync(ctx->groups[group].d_von, von, ctx->groups[group].n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  XG_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int xgpu_b4_von_get(xgpu_ctx *ctx, int group, double *von) {
  if (!ctx || group < 0 || group >= (int)ctx->groups.size() || !von) return 1;
  XG_CUDA(cudaMemcpyAsync(von, ctx->groups[group].d_von, ctx->groups[group].n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  XG_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int xgpu_update_state(xgpu_ctx *ctx, const double *d_sol, double *d_next_sta, double *d_curr_sta, double *d_next_sto,
                      double *d_curr_sto, const xgpu_solver_state *ss) {
  if (!ctx || !ss || !d_sol) return 1;
  if (!ctx->finalized) return fail(ctx, 15, "xgpu_finalize has not been called");
  LoadArgs a{};
  SolverFlags &S = a.S;
  S.dcopFlag = ss->dcopFlag; S.tranopFlag = ss->tranopFlag; S.acopFlag = ss->acopFlag;
  S.transientFlag = ss->transientFlag; S.dcsweepFlag = ss->dcsweepFlag; S.initJctFlag = ss->initJctFlag;
  S.initFixFlag = ss->initFixFlag; S.initTranFlag = ss->initTranFlag; S.newtonIter = ss->newtonIter;
  S.locaEnabledFlag = ss->locaEnabledFlag; S.artParameterFlag = ss->artParameterFlag;
  S.voltageLimiterFlag = ss->voltageLimiterFlag; S.gmin = ss->gmin; S.gainScale = ss->gainScale;
  S.nltermScale = ss->nltermScale; S.vgstConst = ss->vgstConst; S.vdsScaleMin = ss->vdsScaleMin;
  S.sizeScale = ss->sizeScale; S.currTimeStep = ss->currTimeStep; S.lastTimeStep = ss->lastTimeStep;
  S.beginIntegrationFlag = ss->beginIntegrationFlag;
  a.sol = d_sol; a.next_sta = d_next_sta; a.curr_sta = d_curr_sta; a.next_sto = d_next_sto; a.curr_sto = d_curr_sto;
  a.last_sto = ctx->d_last_sto ? ctx->d_last_sto : d_curr_sto;
  for (int p = 0; p < 4; ++p) a.vec_planes[p] = ctx->d_vec_planes + (int64_t)p * ctx->vec_plane;
  for (int p = 0; p < 2; ++p) a.mat_planes[p] = ctx->d_mat_planes + (int64_t)p * ctx->mat_plane;
  for (auto &g : ctx->groups) {
    const bool lockstep = ctx->b4_lockstep && ctx->b4_arith == 2;
    const bool uniform = (ctx->b4_uniform || lockstep) && (int)g.run_start.size() <= kMaxUniformRuns;
    if (lockstep && !uniform) return fail(ctx, 18, "b4_lockstep needs the uniform-record kernel (at most 64 model/bin runs per group)");
    if (uniform && !g.packs_valid) {
      g.packs.clear();
      for (size_t r = 0; r < g.run_start.size(); ++r) {
        if (r % kRunsPerPack == 0) { g.packs.emplace_back(); g.packs.back().nruns = 0; }
        BinRun &br = g.packs.back().run[g.packs.back().nruns++];
        br.M = ctx->h_models[g.run_model[r]]; br.P = ctx->h_sizes[g.run_size[r]];
        br.start = g.run_start[r]; br.count = g.run_count[r];
      }
      // the same runs cut in two (pipelined host path): part 0 = the first pipe_frac of every run, part 1 = the rest
      for (int part = 0; part < 2; ++part) {
        g.packs_part[part] = g.packs;
        for (auto &pk : g.packs_part[part])
          for (int r = 0; r < pk.nruns; ++r) {
            const int cut = xg_pipe_cut(ctx, pk.run[r].count);
            if (part == 0) pk.run[r].count = cut; else { pk.run[r].start += cut; pk.run[r].count -= cut; }
          }
      }
      g.packs_valid = true;
      // the specialised object whose mode tuple every run's model card carries (bsim4_spec_tuples.def), or none
      g.spec_id = -1;
      for (int t = 0; t < kNumSpecTuples && g.spec_id < 0; ++t) {
        bool ok = true;
        for (size_t r = 0; r < g.run_start.size(); ++r) {
          const B4Model &M = ctx->h_models[g.run_model[r]];
          int k = 0;
#define CHK(name) if (kSpecModes[t][k] != -2 && M.name != kSpecModes[t][k]) ok = false; ++k;
          XB_B4_MODEL_I(CHK)
#undef CHK
          if (!(M.versionDouble >= 4.8)) ok = false;      // the specialised builds are the 4.8.2 evaluator
        }
        if (ok) g.spec_id = t;
      }
    }
    // b4_threads == 0: pick the block shape (measured on B200 at the C2 operating point,
    // profiles/r01_b4_kernel_variants_v3.json).  Once the kernel image is compact enough for the instruction
    // caches, 12 warps per SM (128 x 3, 168 registers) beat 8 warps without spills; the mode-specialised build
    // on large groups gains a little more from 16 warps (128 x 4).
    const bool spec = ctx->b4_spec && uniform && g.spec_id >= 0 && ctx->b4_arith == 2 && !lockstep && !g.general;
    int threads = ctx->b4_threads, minblocks = ctx->b4_minblocks;
    if (threads == 0) { threads = 128; minblocks = (spec && g.n > 150000) ? 4 : 3; }      // crossover between 100k and 200k (profiles/r01_b4_occupancy.json)
    if (ctx->eval_part >= 0 && !uniform) return fail(ctx, 20, "partial evaluation needs the uniform-record kernel");
    const std::vector<BinPack> &pk = (ctx->eval_part >= 0) ? g.packs_part[ctx->eval_part] : g.packs;
    const int nl = launch_b4_group(g.dev, a, ctx->b4_arith, lockstep, threads, minblocks,
                                   uniform ? pk.data() : nullptr, uniform ? (int)pk.size() : 0, ctx->stream,
                                   spec ? g.spec_id : -1);
    g.last_spec = spec ? g.spec_id : -1;
    if (nl < 0) return fail(ctx, 19, "unsupported BSIM4 launch shape (b4_threads, b4_minblocks)");
    ctx->launches += nl;
  }
  if (ctx->eval_part <= 0)      // the small-device groups belong to the first part
    for (auto &g : ctx->sgroups) { xb::simple::launch_group(g.dev, a, ctx->stream); ++ctx->launches; }
  XG_CUDA(cudaGetLastError());
  return 0;
}

int xgpu_load_vectors(xgpu_ctx *ctx, double *d_f, double *d_q, double *d_fl, double *d_ql, int accumulate) {
  if (!ctx || !d_f || !d_q || !d_fl || !d_ql) return 1;
  if (!ctx->finalized) return fail(ctx, 15, "xgpu_finalize has not been called");
  const double *in[4]; double *out[4] = {d_f, d_q, d_fl, d_ql};
  for (int p = 0; p < 4; ++p) in[p] = ctx->d_vec_planes + (int64_t)p * ctx->vec_plane;
  launch_gather(ctx->vec_map, 4, in, ctx->vec_plane, out, accumulate != 0, ctx->stream);
  ++ctx->launches;
  XG_CUDA(cudaGetLastError());
  return 0;
}

int xgpu_load_matrices(xgpu_ctx *ctx, double *d_dFdx, double *d_dQdx, int accumulate) {
  if (!ctx || !d_dFdx || !d_dQdx) return 1;
  if (!ctx->finalized) return fail(ctx, 15, "xgpu_finalize has not been called");
  const double *in[2]; double *out[2] = {d_dFdx, d_dQdx};
  for (int p = 0; p < 2; ++p) in[p] = ctx->d_mat_planes + (int64_t)p * ctx->mat_plane;
  launch_gather(ctx->mat_map, 2, in, ctx->mat_plane, out, accumulate != 0, ctx->stream);
  ++ctx->launches;
  XG_CUDA(cudaGetLastError());
  return 0;
}

int xgpu_load_dae(xgpu_ctx *ctx, const double *d_sol, double *d_next_sta, double *d_curr_sta, double *d_next_sto,
                  double *d_curr_sto, const xgpu_solver_state *ss, double *d_f, double *d_q, double *d_fl, double *d_ql,
                  double *d_dFdx, double *d_dQdx, int accumulate) {
  if (!ctx || !d_f || !d_q || !d_fl || !d_ql || !d_dFdx || !d_dQdx) return 1;
  const int rc = xgpu_update_state(ctx, d_sol, d_next_sta, d_curr_sta, d_next_sto, d_curr_sto, ss);
  if (rc) return rc;
  const double *vin[4], *min_[2];
  double *vout[4] = {d_f, d_q, d_fl, d_ql}, *mout[2] = {d_dFdx, d_dQdx};
  for (int p = 0; p < 4; ++p) vin[p] = ctx->d_vec_planes + (int64_t)p * ctx->vec_plane;
  for (int p = 0; p < 2; ++p) min_[p] = ctx->d_mat_planes + (int64_t)p * ctx->mat_plane;
  ctx->launches += launch_gather_fused(ctx->vec_map, vin, vout, ctx->mat_map, min_, mout, accumulate != 0, ctx->stream);
  XG_CUDA(cudaGetLastError());
  return 0;
}

int xgpu_jacobian_combine(xgpu_ctx *ctx, double qs, const double *d_dQdx, double fs, const double *d_dFdx, double *d_jac) {
  if (!ctx || !d_dQdx || !d_dFdx || !d_jac) return 1;
  launch_linear_combo(ctx->nnz, qs, d_dQdx, fs, d_dFdx, d_jac, ctx->stream);
  ++ctx->launches;
  XG_CUDA(cudaGetLastError());
  return 0;
}

namespace {
__global__ void and_flags_k(const int *flags, int n, int *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flags[i] == 0) *out = 0;      // every writer stores the same value
}
}  // namespace

int xgpu_all_converged(xgpu_ctx *ctx, int *converged) {
  if (!ctx || !converged) return 1;
  if (!ctx->finalized) return fail(ctx, 15, "xgpu_finalize has not been called");
  int all = 1;
  XG_CUDA(cudaMemcpyAsync(ctx->d_conv, &all, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  for (auto &g : ctx->groups) { and_flags_k<<<(g.n + 255) / 256, 256, 0, ctx->stream>>>(g.d_orig, g.n, ctx->d_conv); ++ctx->launches; }
  for (auto &g : ctx->sgroups) { and_flags_k<<<(g.n + 255) / 256, 256, 0, ctx->stream>>>(g.d_orig, g.n, ctx->d_conv); ++ctx->launches; }
  XG_CUDA(cudaMemcpyAsync(&all, ctx->d_conv, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  XG_CUDA(cudaStreamSynchronize(ctx->stream));
  *converged = all;
  return 0;
}

int xgpu_state_set(xgpu_ctx *ctx, int which, const double *h) {
  if (!ctx || which < 0 || which > 4 || !h || !ctx->finalized) return 1;
  const size_t cnt = which == 4 ? (ctx->needs_last_sto ? ctx->n_store : 0) : which < 2 ? ctx->n_store : ctx->n_state;
  if (cnt == 0) return 0;
  XG_CUDA(cudaMemcpyAsync(ctx->buf[7 + which], h, cnt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  XG_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int xgpu_state_get(xgpu_ctx *ctx, int which, double *h) {
  if (!ctx || which < 0 || which > 4 || !h || !ctx->finalized) return 1;
  const size_t cnt = which == 4 ? (ctx->needs_last_sto ? ctx->n_store : 0) : which < 2 ? ctx->n_store : ctx->n_state;
  if (cnt == 0) return 0;
  XG_CUDA(cudaMemcpyAsync(h, ctx->buf[7 + which], cnt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  XG_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

double *xgpu_device_buffer(xgpu_ctx *ctx, int which) {
  if (!ctx || which < 0 || which > 11 || !ctx->finalized) return nullptr;
  return ctx->buf[which];
}

int xgpu_load_host(xgpu_ctx *ctx, const double *h_sol, const xgpu_solver_state *ss, double *h_f, double *h_q,
                   double *h_fl, double *h_ql, double *h_dFdx, double *h_dQdx) {
  if (!ctx || !h_sol || !ss) return 1;
  if (!ctx->finalized) return fail(ctx, 15, "xgpu_finalize has not been called");
  double **b = ctx->buf;
  XG_CUDA(cudaMemcpyAsync(b[0], h_sol, ctx->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  // Pinned, mapped host buffers (cudaHostAlloc / cudaHostRegister: what a GPU-aware caller provides) are written by the
  // assembly kernel directly over PCIe -- the transfer overlaps the assembly and the six device-to-host copies with
  // their launch overheads disappear.  Pageable buffers take the copy path below.
  if (ctx->zero_copy_out && h_f && h_q && h_fl && h_ql && h_dFdx && h_dQdx) {
    double *hp[6] = {h_f, h_q, h_fl, h_ql, h_dFdx, h_dQdx}, *dp[6];
    bool mapped = true;
    for (int p = 0; p < 6 && mapped; ++p)
      if (cudaHostGetDevicePointer((void **)&dp[p], hp[p], 0) != cudaSuccess) { cudaGetLastError(); mapped = false; }
    if (mapped) {
      const int rc0 = xgpu_load_dae(ctx, b[0], b[9], b[10], b[7], b[8], ss, dp[0], dp[1], dp[2], dp[3], dp[4], dp[5], 0);
      if (rc0) return rc0;
      XG_CUDA(cudaStreamSynchronize(ctx->stream));
      return 0;
    }
  }
  const int rc = xgpu_load_dae(ctx, b[0], b[9], b[10], b[7], b[8], ss, b[1], b[2], b[3], b[4], b[5], b[6], 0);
  if (rc) return rc;
  double *hv[4] = {h_f, h_q, h_fl, h_ql};
  for (int p = 0; p < 4; ++p)
    if (hv[p]) XG_CUDA(cudaMemcpyAsync(hv[p], b[1 + p], ctx->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (h_dFdx) XG_CUDA(cudaMemcpyAsync(h_dFdx, b[5], ctx->nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (h_dQdx) XG_CUDA(cudaMemcpyAsync(h_dQdx, b[6], ctx->nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  XG_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C++" {
namespace {
// J = qs dQdx + fs dFdx over the nnz entries and r = -(qs Q + fs F) (+ qs dQdxdVp + fs dFdxdVp) over the n rows, one launch.
// The outputs may be device buffers or pinned, mapped host memory (zero-copy: the stores travel over PCIe while the
// kernel runs, no separate copy).
__global__ void __launch_bounds__(256) jr_kernel(long long nnz, int n, double qs, double fs, const double *__restrict__ dQdx,
                                                 const double *__restrict__ dFdx, const double *__restrict__ F,
                                                 const double *__restrict__ Q, const double *__restrict__ Fl,
                                                 const double *__restrict__ Ql, int limiter, double *__restrict__ J,
                                                 double *__restrict__ r) {
  xb::pdl_wait();
  const long long k = (long long)blockIdx.x * 256 + threadIdx.x;      // callers pass pointers offset to their window
  if (k < nnz) J[k] = qs * dQdx[k] + fs * dFdx[k];
  if (k < n) {
    double v = -(qs * Q[k] + fs * F[k]);
    if (limiter) v += qs * Ql[k] + fs * Fl[k];
    r[k] = v;
  }
}
}  // namespace
}  // extern "C++"

int xgpu_load_host_jr(xgpu_ctx *ctx, const double *h_sol, const xgpu_solver_state *ss, double qscalar, double fscalar,
                      double *h_rhs, double *h_jac) {
  if (!ctx || !h_sol || !ss || !h_rhs || !h_jac) return 1;
  if (!ctx->finalized) return fail(ctx, 15, "xgpu_finalize has not been called");
  double **b = ctx->buf;
  XG_CUDA(cudaMemcpyAsync(b[0], h_sol, ctx->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (ctx->pipe.ok && ctx->pipeline_host && !ctx->zero_copy_out && ctx->b4_uniform && !ctx->b4_lockstep && !xg_dist_multi(ctx)) {
    // Two parts: evaluate the first part of every run, assemble and combine the destinations only it feeds, start their
    // DMA on the second stream; meanwhile evaluate the second part, assemble and combine the rest, copy it.  Same kernels
    // on the same data in the same order per destination: bitwise the result of the one-pass path.
    const int rs = ctx->pipe.vec_split; const long long ks = ctx->pipe.mat_split;
    // option "pipe_r_mapped" (off): the residual part (0.8 MB) stored straight into the caller's mapped pinned buffer, so
    // that two of the four DMA transfers go away.  Measured on config 2 (scripts/pipe_r_time.py): 168 us vs 162 us with the
    // DMA copies, bitwise equal -- the stores over PCIe cost more than the fixed cost of two transfers; kept for other hosts
    double *r_host = nullptr;
    if (ctx->pipe_r_mapped && cudaHostGetDevicePointer((void **)&r_host, h_rhs, 0) != cudaSuccess) { cudaGetLastError(); r_host = nullptr; }
    const double *vin[4], *min_[2];
    double *vout[4] = {b[1], b[2], b[3], b[4]}, *mout[2] = {b[5], b[6]};
    for (int p = 0; p < 4; ++p) vin[p] = ctx->d_vec_planes + (int64_t)p * ctx->vec_plane;
    for (int p = 0; p < 2; ++p) min_[p] = ctx->d_mat_planes + (int64_t)p * ctx->mat_plane;
    for (int part = 0; part < 2; ++part) {
      ctx->eval_part = part;
      const int rc = xgpu_update_state(ctx, b[0], b[9], b[10], b[7], b[8], ss);
      ctx->eval_part = -1;
      if (rc) return rc;
      xb::GatherMapDev mv = ctx->vec_map, mm = ctx->mat_map;
      if (part == 0) { mv.d_lo = 0; mv.d_hi = rs; mv.with_chunks = 0; mm.d_lo = 0; mm.d_hi = (int)ks; mm.with_chunks = 0; }
      else { mv.d_lo = rs; mv.d_hi = -1; mm.d_lo = (int)ks; mm.d_hi = -1; }
      ctx->launches += launch_gather_fused(mv, vin, vout, mm, min_, mout, false, ctx->stream);
      const long long k0 = part == 0 ? 0 : ks, k1 = part == 0 ? ks : ctx->nnz;
      const int r0 = part == 0 ? 0 : rs, r1 = part == 0 ? rs : ctx->n;
      const long long m = std::max<long long>(k1 - k0, r1 - r0);
      if (m > 0) {
        xb::launch_pdl(jr_kernel, dim3((unsigned)((m + 255) / 256)), dim3(256), 0, ctx->stream, k1 - k0, r1 - r0, qscalar, fscalar,
                       (const double *)(b[6] + k0), (const double *)(b[5] + k0), (const double *)(b[1] + r0), (const double *)(b[2] + r0),
                       (const double *)(b[3] + r0), (const double *)(b[4] + r0), ss->voltageLimiterFlag, b[5] + k0,
                       r_host ? r_host + r0 : b[1] + r0);
        ++ctx->launches;
      }
      cudaStream_t cs = ctx->stream;
      if (part == 0) {      // the first window travels on the second stream while the second part is evaluated
        XG_CUDA(cudaEventRecord(ctx->pipe.ev_a, ctx->stream));
        XG_CUDA(cudaStreamWaitEvent(ctx->pipe.s2, ctx->pipe.ev_a, 0));
        cs = ctx->pipe.s2;
      }
      if (r1 > r0 && !r_host) XG_CUDA(cudaMemcpyAsync(h_rhs + r0, b[1] + r0, (size_t)(r1 - r0) * sizeof(double), cudaMemcpyDeviceToHost, cs));
      if (k1 > k0) XG_CUDA(cudaMemcpyAsync(h_jac + k0, b[5] + k0, (size_t)(k1 - k0) * sizeof(double), cudaMemcpyDeviceToHost, cs));
    }
    XG_CUDA(cudaStreamSynchronize(ctx->pipe.s2));
    XG_CUDA(cudaStreamSynchronize(ctx->stream));
    XG_CUDA(cudaGetLastError());
    return 0;
  }
  const int rc = xgpu_load_dae(ctx, b[0], b[9], b[10], b[7], b[8], ss, b[1], b[2], b[3], b[4], b[5], b[6], 0);
  if (rc) return rc;
  double *dj = nullptr, *dr = nullptr;
  bool mapped = ctx->zero_copy_out != 0;
  if (mapped && (cudaHostGetDevicePointer((void **)&dj, h_jac, 0) != cudaSuccess ||
                 cudaHostGetDevicePointer((void **)&dr, h_rhs, 0) != cudaSuccess)) { cudaGetLastError(); mapped = false; }
  if (!mapped) { dj = b[5]; dr = b[1]; }      // in place: J over dFdx, r over F (element-wise, same index)
  const long long m = std::max<long long>(ctx->nnz, ctx->n);
  xb::launch_pdl(jr_kernel, dim3((unsigned)((m + 255) / 256)), dim3(256), 0, ctx->stream, (long long)ctx->nnz, ctx->n, qscalar,
                 fscalar, (const double *)b[6], (const double *)b[5], (const double *)b[1], (const double *)b[2],
                 (const double *)b[3], (const double *)b[4], ss->voltageLimiterFlag, dj, dr);
  ++ctx->launches;
  if (!mapped) {
    XG_CUDA(cudaMemcpyAsync(h_rhs, dr, ctx->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    XG_CUDA(cudaMemcpyAsync(h_jac, dj, ctx->nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  XG_CUDA(cudaStreamSynchronize(ctx->stream));
  XG_CUDA(cudaGetLastError());
  return 0;
}

extern "C++" {
namespace {
void lu_drop_graphs(xgpu_ctx *ctx) {
  for (xgpu_ctx::LuGraph *g : {&ctx->g_refactor, &ctx->g_solve}) {
    if (g->exec) cudaGraphExecDestroy(g->exec);
    *g = xgpu_ctx::LuGraph();
  }
}
// Run a launch sequence directly, or -- when it is long (launch-latency bound) -- as a CUDA graph captured on the
// first call with these device pointers.  The kernels read their inputs through those pointers only, so replaying
// the graph on new VALUES in the same buffers is the same computation.
constexpr int kGraphMinLaunches = 12;
template <class F>
int lu_run(xgpu_ctx *ctx, xgpu_ctx::LuGraph &g, const void *k0, const void *k1, const void *k2, F &&launch) {
  if (g.exec && g.k0 == k0 && g.k1 == k1 && g.k2 == k2) {
    XG_CUDA(cudaGraphLaunch(g.exec, ctx->stream));
    ctx->launches += g.launches;
    return 0;
  }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(ctx->stream, &cs);
  // the legacy default stream cannot be captured
  const bool capturable = ctx->stream != nullptr && ctx->stream != cudaStreamLegacy && ctx->stream != cudaStreamPerThread;
  const bool try_graph = ctx->lu_graphs && capturable && cs == cudaStreamCaptureStatusNone && g.launches != -1;
  if (try_graph && g.launches >= kGraphMinLaunches) {        // second call of a long sequence: capture it
    if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      const int n = launch();
      cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
      if (e == cudaSuccess && graph) e = cudaGraphInstantiate(&g.exec, graph, 0);
      if (graph) cudaGraphDestroy(graph);
      if (e == cudaSuccess && g.exec) {
        g.k0 = k0; g.k1 = k1; g.k2 = k2; g.launches = n;
        XG_CUDA(cudaGraphLaunch(g.exec, ctx->stream));
        ctx->launches += n;
        return 0;
      }
      cudaGetLastError();
      g.exec = nullptr; g.launches = -1;                     // capture not possible here: stay on plain launches
    }
  }
  const int n = launch();
  if (g.launches != -1) g.launches = n;
  ctx->launches += n;
  XG_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace
}  // extern "C++"

}  // extern "C"
int xg_lu_refactor_async(xgpu_ctx *ctx, const double *d_vals) {
  return lu_run(ctx, ctx->g_refactor, d_vals, nullptr, nullptr,
                [&] { return xb::lu::launch_refactor(ctx->lu_dev, d_vals, ctx->stream); });
}
int xg_lu_status(xgpu_ctx *ctx, int *status) {
  XG_CUDA(cudaMemcpyAsync(status, ctx->lu_dev.status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  XG_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" {

int xgpu_lu_analyze(xgpu_ctx *ctx, const double *d_vals) {
  if (!ctx || !d_vals) return 100;
  if (ctx->rowptr.empty()) return fail(ctx, 112, "no CSR pattern");
  XG_CUDA(cudaSetDevice(ctx->device));
  std::vector<double> vals((size_t)ctx->nnz);
  XG_CUDA(cudaMemcpyAsync(vals.data(), d_vals, vals.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  XG_CUDA(cudaStreamSynchronize(ctx->stream));
  xb::lu::set_batching(ctx->lu_batch != 0);
  const int rc = xb::lu::analyze_and_factor(ctx->n, ctx->rowptr.data(), ctx->colind.data(), vals.data(), 0.001, ctx->lu_plan);
  if (rc == 1) return fail(ctx, 1, "matrix is structurally singular");
  lu_drop_graphs(ctx);
  XG_CUDA(xb::lu::upload_plan(ctx->lu_plan, ctx->lu_dev));
  ctx->lu_ready = true;
  if (rc == 2) return fail(ctx, 2, "matrix is numerically singular");
  return 0;
}

int xgpu_lu_import(xgpu_ctx *ctx, const int32_t *row_perm, const int32_t *col_perm, int nblocks, const int32_t *block_ptr,
                   const int32_t *Lp, const int32_t *Li, const int32_t *Up, const int32_t *Ui, const double *row_scale) {
  if (!ctx || !row_perm || !col_perm || !block_ptr || !Lp || !Li || !Up || !Ui) return 100;
  if (ctx->rowptr.empty()) return fail(ctx, 112, "no CSR pattern");
  XG_CUDA(cudaSetDevice(ctx->device));
  const char *why = "";
  xb::lu::set_batching(ctx->lu_batch != 0);
  const int rc = xb::lu::import_factorization(ctx->n, ctx->rowptr.data(), ctx->colind.data(), row_perm, col_perm, nblocks,
                                              block_ptr, Lp, Li, Up, Ui, row_scale, ctx->lu_plan, &why);
  if (rc) { ctx->lu_ready = false; return fail(ctx, 3, std::string("xgpu_lu_import: ") + why); }
  lu_drop_graphs(ctx);
  XG_CUDA(xb::lu::upload_plan(ctx->lu_plan, ctx->lu_dev));
  ctx->lu_ready = true;
  return 0;
}

int xgpu_lu_export_sizes(const xgpu_ctx *ctx, int32_t *sizes4) {
  if (!ctx || !sizes4) return 100;
  if (!ctx->lu_ready) return 113;
  const xb::lu::LuPlan &p = ctx->lu_plan;
  sizes4[0] = p.n; sizes4[1] = (int)p.block_ptr.size() - 1; sizes4[2] = (int)p.Li.size(); sizes4[3] = (int)p.Ui.size();
  return 0;
}

int xgpu_lu_export(xgpu_ctx *ctx, int32_t *row_perm, int32_t *col_perm, int32_t *block_ptr, int32_t *Lp, int32_t *Li,
                   double *Lx, int32_t *Up, int32_t *Ui, double *Ux) {
  if (!ctx) return 100;
  if (!ctx->lu_ready) return fail(ctx, 113, "xgpu_lu_analyze has not been called");
  const xb::lu::LuPlan &p = ctx->lu_plan;
  if (row_perm) std::copy(p.row_perm.begin(), p.row_perm.end(), row_perm);
  if (col_perm) std::copy(p.col_perm.begin(), p.col_perm.end(), col_perm);
  if (block_ptr) std::copy(p.block_ptr.begin(), p.block_ptr.end(), block_ptr);
  if (Lp) std::copy(p.Lp.begin(), p.Lp.end(), Lp);
  if (Li) std::copy(p.Li.begin(), p.Li.end(), Li);
  if (Up) std::copy(p.Up.begin(), p.Up.end(), Up);
  if (Ui) std::copy(p.Ui.begin(), p.Ui.end(), Ui);
  // numeric values of the latest factorization on the device (batched groups keep theirs interleaved: copy them over)
  if ((Lx || Ux) && !ctx->lu_dev.batch.empty()) ctx->launches += xb::lu::launch_batch_export(ctx->lu_dev, ctx->stream);
  if (Lx && !p.Li.empty()) XG_CUDA(cudaMemcpyAsync(Lx, ctx->lu_dev.Lx, p.Li.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (Ux && !p.Ui.empty()) XG_CUDA(cudaMemcpyAsync(Ux, ctx->lu_dev.Ux, p.Ui.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  XG_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int xgpu_lu_refactor(xgpu_ctx *ctx, const double *d_vals) {
  if (!ctx || !d_vals) return 100;
  if (!ctx->lu_ready) return fail(ctx, 113, "xgpu_lu_analyze has not been called");
  if (ctx->lu_repivot) return xgpu_lu_analyze(ctx, d_vals);      // KLU_REPIVOT=1: pivoting factorization every time
  {
    const int rc = xg_lu_refactor_async(ctx, d_vals);
    if (rc) return rc;
  }
  int status = 0;
  { const int rc = xg_lu_status(ctx, &status); if (rc) return rc; }
  if (status & 1) return fail(ctx, 2, "zero or non-finite pivot during refactorization");
  if (status & 4) return fail(ctx, 3, "a pivot of the fixed sequence fails the partial-pivoting threshold: re-analyse (re-pivot)");
  return 0;
}

int xgpu_lu_solve(xgpu_ctx *ctx, const double *d_vals, const double *d_rhs, double *d_x) {
  if (!ctx || !d_vals || !d_rhs || !d_x) return 100;
  if (!ctx->lu_ready) return fail(ctx, 113, "xgpu_lu_analyze has not been called");
  return lu_run(ctx, ctx->g_solve, d_vals, d_rhs, d_x,
                [&] { return xb::lu::launch_solve(ctx->lu_dev, d_vals, d_rhs, d_x, ctx->stream); });
}

// One Newton iteration for a caller that keeps its vectors on the host: x in, dx (and optionally the residual) out.
// Everything between the two copies stays on the device: evaluation, assembly, linear-device replay, J and r in one
// pass, refactorization on the plan of the previous call (first call, or a failed pivot check: host analysis), solves.
int xgpu_newton_step_host(xgpu_ctx *ctx, const double *h_sol, const xgpu_solver_state *ss, double qscalar, double fscalar,
                          const double *h_hist, double *h_dx, double *h_rhs) {
  if (!ctx || !h_sol || !ss || !h_dx) return 1;
  if (!ctx->finalized) return fail(ctx, 15, "xgpu_finalize has not been called");
  double **b = ctx->buf;
  XG_CUDA(cudaMemcpyAsync(b[0], h_sol, ctx->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (h_hist) {      // the caller's history / source terms of the residual
    if (!ctx->d_hist) XG_CUDA(cudaMalloc((void **)&ctx->d_hist, (size_t)ctx->n * sizeof(double)));
    XG_CUDA(cudaMemcpyAsync(ctx->d_hist, h_hist, ctx->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  int rc = xgpu_load_dae(ctx, b[0], b[9], b[10], b[7], b[8], ss, b[1], b[2], b[3], b[4], b[5], b[6], 0);
  if (rc) return rc;
  // linear devices: F += G x, Q += C x, dFdx += G, dQdx += C (FilteredMatrix replay, N_LOA_CktLoader.C:700-788)
  const XgLinearPart &G = ctx->linG, &C = ctx->linC;
  xb::vec::spmv_add(G.nrows, G.rows, G.ptr, G.col, G.val, b[0], b[1], ctx->stream);
  xb::vec::spmv_add(C.nrows, C.rows, C.ptr, C.col, C.val, b[0], b[2], ctx->stream);
  xb::vec::scatter_add(G.nnz, G.pos, G.val, b[5], ctx->stream);
  xb::vec::scatter_add(C.nnz, C.pos, C.val, b[6], ctx->stream);
  ctx->launches += (G.nrows > 0) + (C.nrows > 0) + (G.nnz > 0) + (C.nnz > 0);
  const bool bordered = ctx->dist && (ctx->dist->ns > 0 || xg_dist_multi(ctx));
  if (bordered && xg_dist_multi(ctx)) {      // border rows hold per-rank partial sums (N_LOA_CktLoader.C:816-829)
    double *vv[4] = {b[1], b[2], b[3], b[4]};
    rc = xg_dist_reduce_border_rows(ctx, vv, 4);
    if (rc) return rc;
  }
  const long long m = std::max<long long>(ctx->nnz, ctx->n);
  xb::launch_pdl(jr_kernel, dim3((unsigned)((m + 255) / 256)), dim3(256), 0, ctx->stream, (long long)ctx->nnz, ctx->n, qscalar,
                 fscalar, (const double *)b[6], (const double *)b[5], (const double *)b[1], (const double *)b[2],
                 (const double *)b[3], (const double *)b[4], ss->voltageLimiterFlag, b[5], b[1]);
  ++ctx->launches;
  if (bordered) {      // interior LU per rank + the replicated border (Schur) system, dist.cu
    if (h_hist) { xb::vec::axpby(b[1], 1.0, b[1], -1.0, ctx->d_hist, ctx->n, ctx->stream); ++ctx->launches; }
    if (h_rhs) XG_CUDA(cudaMemcpyAsync(h_rhs, b[1], ctx->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (!ctx->dist->analyzed) { rc = xgpu_border_analyze(ctx, b[5]); if (rc != 0 && rc != 2) return rc; }
    const bool defer = !xg_dist_multi(ctx) && !ctx->lu_repivot;
    rc = xg_border_solve(ctx, b[5], b[1], b[2], xg_dist_multi(ctx) ? 1 : 0, defer);
    if (rc) return rc;
    XG_CUDA(cudaMemcpyAsync(h_dx, b[2], ctx->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    int status = 0;
    if (defer) { rc = xg_lu_status(ctx, &status); if (rc) return rc; }      // the one synchronisation of the call
    else XG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (status & 5) {      // bad or sub-threshold pivot of the fixed sequence: re-pivot the interior, solve again
      rc = xgpu_border_analyze(ctx, b[5]); ++ctx->dist->reanalyses;
      if (rc != 0 && rc != 2) return rc;
      rc = xg_border_solve(ctx, b[5], b[1], b[2], 0, false);
      if (rc) return rc;
      XG_CUDA(cudaMemcpyAsync(h_dx, b[2], ctx->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      XG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    XG_CUDA(cudaGetLastError());
    return 0;
  }
  if (h_hist) { xb::vec::axpby(b[1], 1.0, b[1], -1.0, ctx->d_hist, ctx->n, ctx->stream); ++ctx->launches; }   // r -= hist
  if (h_rhs) XG_CUDA(cudaMemcpyAsync(h_rhs, b[1], ctx->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  bool analysed = false;
  if (!ctx->lu_ready || ctx->lu_repivot) { rc = xgpu_lu_analyze(ctx, b[5]); analysed = true; }
  else rc = xg_lu_refactor_async(ctx, b[5]);
  if (rc) return rc;
  rc = xgpu_lu_solve(ctx, b[5], b[1], b[2]);
  if (rc) return rc;
  XG_CUDA(cudaMemcpyAsync(h_dx, b[2], ctx->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  int status = 0;
  if (!analysed) { rc = xg_lu_status(ctx, &status); if (rc) return rc; }      // the one synchronisation of the call
  else XG_CUDA(cudaStreamSynchronize(ctx->stream));
  if (status & 5) {                                  // bad or sub-threshold pivot of the fixed sequence: re-pivot
    rc = xgpu_lu_analyze(ctx, b[5]);
    if (rc) return rc;
    rc = xgpu_lu_solve(ctx, b[5], b[1], b[2]);
    if (rc) return rc;
    XG_CUDA(cudaMemcpyAsync(h_dx, b[2], ctx->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    XG_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  XG_CUDA(cudaGetLastError());
  return 0;
}

int xgpu_lu_host_factor_solve(int n, const int32_t *rowptr, const int32_t *colind, const double *vals,
                              const double *rhs, double *x, double *info) {
  if (n <= 0 || !rowptr || !colind || !vals || !rhs || !x) return 100;
  xb::lu::LuPlan p;
  const int rc = xb::lu::analyze_and_factor(n, rowptr, colind, vals, 0.001, p);
  if (rc == 1) return 1;
  xb::lu::solve_host(p, rhs, x);
  if (info) {
    int largest = 0;
    for (size_t b = 0; b + 1 < p.block_ptr.size(); ++b) largest = std::max(largest, p.block_ptr[b + 1] - p.block_ptr[b]);
    info[0] = p.n; info[1] = (double)p.block_ptr.size() - 1; info[2] = largest; info[3] = (double)p.Li.size();
    info[4] = (double)p.Ui.size(); info[5] = (double)p.off_row.size(); info[6] = (double)p.level_ptr.size() - 1;
    info[7] = p.refactor_flops;
  }
  return rc;
}

int xgpu_lu_host_batch_selfcheck(int n, const int32_t *rowptr, const int32_t *colind, const double *vals0, const double *vals1,
                                 double *out4) {
  if (n <= 0 || !rowptr || !colind || !vals0 || !vals1 || !out4) return 100;
  xb::lu::LuPlan p;
  const int rc = xb::lu::analyze_and_factor(n, rowptr, colind, vals0, 0.001, p);
  if (rc == 1) return 1;
  xb::lu::batch_selfcheck_host(p, vals1, out4);
  return rc;
}

int xgpu_lu_info(const xgpu_ctx *ctx, double *info) {
  if (!ctx || !info || !ctx->lu_ready) return 100;
  const xb::lu::LuPlan &p = ctx->lu_plan;
  int largest = 0;
  for (size_t b = 0; b + 1 < p.block_ptr.size(); ++b) largest = std::max(largest, p.block_ptr[b + 1] - p.block_ptr[b]);
  info[0] = p.n; info[1] = (double)p.block_ptr.size() - 1; info[2] = largest; info[3] = (double)p.Li.size();
  info[4] = (double)p.Ui.size(); info[5] = (double)p.off_row.size(); info[6] = (double)p.level_ptr.size() - 1;
  info[7] = p.refactor_flops;
  return 0;
}

int xgpu_measure_fp64_peak(xgpu_ctx *ctx, double *tflops) {
  if (!ctx || !tflops) return 1;
  XG_CUDA(cudaSetDevice(ctx->device));
  double t = 0.0;
  cudaError_t e = xb::measure_fp64_peak(ctx->stream, &t);
  if (e != cudaSuccess) return fail(ctx, 100 + (int)e, cudaGetErrorString(e));
  *tflops = t;
  return 0;
}

long long xgpu_launch_count(const xgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
