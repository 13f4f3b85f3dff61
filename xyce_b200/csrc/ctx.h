// xyce_b200 -- internal definition of the context behind the C ABI (shared by capi.cu and sim_gpu.cu).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/xyce_b200.h"
#include "assembly.cuh"
#include "b4_kernels.cuh"
#include "lu.h"
#include "simple_kernels.cuh"

struct XgHostGroup {
  int n = 0;
  int general = 0;
  std::vector<int32_t> lids;       // [12][n] transposed (node-major)
  xb::b4::GroupDev dev{};
  // owned device memory
  double *d_inst_d = nullptr, *d_von = nullptr;
  int *d_topo = nullptr, *d_model_idx = nullptr, *d_size_idx = nullptr, *d_lids = nullptr;
  int *d_sto0 = nullptr, *d_sta0 = nullptr, *d_orig = nullptr;
  int *d_branch0 = nullptr;        // [n] first branch-data LID of every instance (lead currents), null = not requested
  double *d_lead = nullptr;        // [8][n] lead block of a general-topology group (written by the evaluation kernel)
  // runs of equal (model, bin) along the instance order; packs = their records, rebuilt when the
  // model table changes (empty = more than kMaxUniformRuns runs -> per-thread-record kernel)
  std::vector<int32_t> run_model, run_size, run_start, run_count;
  std::vector<xb::b4::BinPack> packs;
  std::vector<xb::b4::BinPack> packs_part[2];      // the same runs cut at pipe_frac of their length (pipelined host path)
  bool packs_valid = false;
  int spec_id = -1;                // the mode-specialised kernel object that fits every run's model card (bsim4_spec_tuples.def), -1 = none
  int last_spec = -1;              // object used by the last evaluation (diagnostics: xgpu_b4_group_spec)
};

// groups of the small compact models (diode, MOSFET level 1, BJT, ADMS-shaped rlc): flat per-instance records
struct XgSimpleGroup {
  int type = 0;                    // xb::simple::Type
  int n = 0, nodes = 0, slots = 0, nfields = 0, nstore = 0, nstate = 0;
  std::vector<int32_t> lids;       // [nodes][n]
  std::vector<int> slot_row, slot_col;
  xb::simple::GroupDev dev{};
  double *d_rec = nullptr;
  int *d_flags = nullptr, *d_lids = nullptr, *d_sto0 = nullptr, *d_sta0 = nullptr, *d_orig = nullptr;
  int *d_branch0 = nullptr;        // [n] first branch-data LID (lead currents), null = not requested
  double *d_lead = nullptr;        // BJT: [8][n] lead block written by the evaluation kernel
};

// linear-device part (R, C, V, I): constant stamps replayed every load, like the reference's
// FilteredMatrix objects (N_LOA_CktLoader.C:504-578, :700-788)
struct XgLinearPart {
  int nrows = 0, nnz = 0;
  int *rows = nullptr, *ptr = nullptr, *col = nullptr, *pos = nullptr;   // CSR over non-empty rows; pos = index in the system CSR
  double *val = nullptr;
  std::vector<int32_t> h_row, h_col;     // merged COO (host), kept for pattern building
  std::vector<double> h_val;
};

struct XgSource { int row; double scale; int type; double p[7]; };

// Bordered block-diagonal form of the local system and the communicator of a multi-GPU run (dist.cu).
// Local unknown order: [interior | border]; border = unknowns shared between partitions (supply rails, source
// branches) and/or dense nodes the caller wants out of the BTF blocks.  With world > 1 every rank holds a replica of
// the border unknowns in the same order.
struct XgDist {
  int rank = 0, world = 1;
  void *comm = nullptr;              // ncclComm_t
  int ni = 0, ns = 0;                // interior / border unknowns of this rank (ns equal on all ranks)
  long long n_global = 0;            // sum of ni over the ranks + ns
  // A_is entries (interior row, border column): CSR position, row, border column index
  int n_is = 0; int *is_pos = nullptr, *is_row = nullptr, *is_col = nullptr;
  // A_si entries by border row (interior columns), cut into chunks of kSiChunk for the deterministic row reductions
  int n_si = 0, n_chunks = 0; int *si_pos = nullptr, *si_col = nullptr, *chunk_row = nullptr, *chunk_begin = nullptr, *chunk_end = nullptr;
  int *row_chunk_ptr = nullptr;      // [ns + 1] chunks of each border row
  int *ss_pos = nullptr;             // [ns * ns] CSR position of border entry (r, c) or -1
  double *B = nullptr;               // [ns + 1][ni] right-hand sides A_is | b_i, overwritten by A_ii^-1 (.)
  double *partials = nullptr;        // [ns + 1][n_chunks]
  double *red = nullptr;             // [ns][ns + 1] reduced (Schur) system, then its solution in column ns
  double *pack = nullptr;            // small device buffer of the collectives (>= 8 ns + 64 doubles)
  double *h_pack = nullptr;          // pinned mirror (>= 8 * world + 64 doubles)
  std::vector<int32_t> sub_rowptr, sub_colind, sub_index;      // interior x interior sub-pattern and its positions in the CSR values
  bool analyzed = false;
  // peer-memory mailboxes for the small collectives (dist.cu): own allocation, every rank's mailbox as mapped here
  double *p2p_own = nullptr, *p2p_box[16] = {nullptr};
  double **p2p_vecs = nullptr;        // device copy of up to 4 vector pointers (fused pack / unpack of the border rows)
  double *p2p_vecs_host[4] = {nullptr, nullptr, nullptr, nullptr};      // what p2p_vecs currently holds
  bool p2p_attached = false;
  unsigned long long p2p_epoch = 0;
  std::vector<char> col_nonzero;     // [ns] border column has entries in interior rows (its A_ii^-1 solve is needed)
  int reanalyses = 0;                // host re-pivots triggered inside xg_border_solve (bad or sub-threshold pivot)
};

struct xgpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  long long launches = 0;
  // BSIM4 kernel variant (xgpu_set_option): arithmetic 0 strict / 1 fma / 2 fma + reciprocal division,
  // block shape, uniform-record kernel on/off, lock-step barriers 0/1
  int b4_arith = 2, b4_threads = 0 /* auto */, b4_minblocks = 2, b4_uniform = 1, b4_lockstep = 0, b4_spec = 1;

  int n = 0;
  int64_t nnz = 0;
  std::vector<int32_t> rowptr, colind;
  int n_state = 0, n_store = 0;

  xb::b4::B4Model *d_models = nullptr;
  xb::b4::B4Size *d_sizes = nullptr;
  std::vector<xb::b4::B4Model> h_models;
  std::vector<xb::b4::B4Size> h_sizes;
  int n_models = 0, n_sizes = 0;
  std::vector<XgHostGroup> groups;
  std::vector<XgSimpleGroup> sgroups;       // evaluated after the BSIM4 groups, in insertion order
  bool finalized = false;

  // contribution planes
  int64_t vec_plane = 0, mat_plane = 0;
  double *d_vec_planes = nullptr;   // 4 * vec_plane
  double *d_mat_planes = nullptr;   // 2 * mat_plane
  xb::GatherMapDev vec_map, mat_map;
  int *d_conv = nullptr;

  // context-owned system buffers (host-convenience path)
  double *buf[12] = {nullptr};        // xgpu_device_buffer order; 11 = last store
  double *d_last_sto = nullptr;       // xgpu_last_store_set / the transient driver: store vector of the step before
  bool needs_last_sto = false;
  int pipe_r_mapped = 0;              // pipelined host path: store the residual part straight into mapped pinned host memory
  double *d_lead_host = nullptr; int lead_len = 0;      // xgpu_lead_load_host: leadF | leadQ | junctionV staging        // a BJT group with excess phase (PTF != 0) is present

  // linear devices and independent sources
  XgLinearPart linG, linC;
  std::vector<XgSource> sources;
  std::vector<double> pwl;           // (time, value) pairs of all PWL sources (XgSource::p = {td, offset, count, repeat, repeattime})
  double *d_bsrc = nullptr;          // staging for source values

  // sparse LU
  xb::lu::LuPlan lu_plan;
  xb::lu::LuDev lu_dev;
  bool lu_ready = false;
  // CUDA graphs of the refactor / solve launch sequences (large diagonal blocks: one launch per level, hundreds of
  // launches per call).  Captured on first use for a given set of device pointers, replayed afterwards;
  // dropped whenever the plan changes (xgpu_lu_analyze).
  struct LuGraph { cudaGraphExec_t exec = nullptr; const void *k0 = nullptr, *k1 = nullptr, *k2 = nullptr; int launches = 0; };
  LuGraph g_refactor, g_solve;
  int lu_graphs = 1;          // option "lu_graphs": 0 = plain stream launches
  int lu_batch = 1;           // option "lu_batch": 0 = no batched groups (every block on the warp-per-block kernels)
  int lu_repivot = 0;         // option "lu_repivot": 1 = every refactorization re-pivots on the host (KLU_REPIVOT=1)
  int zero_copy_out = 0;      // option "zero_copy_out" = 1: xgpu_load_host lets the assembly kernel write pinned, mapped host outputs directly
                              // (measured +2.6 % on the C2 host-buffer path; off by default: DMA copies behave predictably with many ranks)
  // work space of xgpu_tran_run, kept between runs (driver-level allocation and pinned-memory calls cost
  // milliseconds to tenths of a second each): device pool in doubles, small int arena, pinned readback words
  double *tran_pool = nullptr; size_t tran_pool_len = 0;
  int *tran_ints = nullptr; size_t tran_ints_len = 0;
  double *tran_pinned = nullptr;
  double *d_hist = nullptr;           // xgpu_newton_step_host: the caller's history term on the device
  // Pipelined host-buffer path (xgpu_load_host_jr): the instances are evaluated in two parts (every (model, bin) run cut
  // at pipe_frac of its length); the destinations that only the first part feeds -- a prefix of the rows / nonzeros for
  // circuits numbered device by device -- are assembled and shipped over PCIe while the second part is evaluated.
  struct Pipe { bool ok = false; int vec_split = 0; long long mat_split = 0; cudaStream_t s2 = nullptr; cudaEvent_t ev_a = nullptr, ev_x = nullptr; } pipe;
  int pipeline_host = 1;              // option "pipeline_host": 0 = never pipeline
  double pipe_frac = 0.5;             // option "pipe_percent": share of every run in the first part
  int eval_part = -1;                 // -1 = all instances, 0 / 1 = that part only (set around xgpu_update_state by the pipelined path)
  XgDist *dist = nullptr;     // bordered solve / multi-GPU state (xgpu_border_set, xgpu_comm_init); null = plain single-GPU path
};

// internal entry points shared between capi.cu, sim_gpu.cu and dist.cu
int xg_lu_refactor_async(xgpu_ctx *ctx, const double *d_vals);          // launches only; status stays on the device
int xg_lu_status(xgpu_ctx *ctx, int *status);                           // reads the status word (synchronises)
void xg_dist_free(xgpu_ctx *ctx);
bool xg_dist_multi(const xgpu_ctx *ctx);                                // a communicator with more than one rank is attached
int xg_dist_reduce_border_rows(xgpu_ctx *ctx, double *const *vecs, int nvec);
// every rank contributes k doubles (device); the world * k gathered values arrive in ctx->dist->h_pack (host, after the
// stream synchronisation the caller does anyway)
int xg_dist_allgather(xgpu_ctx *ctx, const double *d_send, int k);
int xg_border_solve(xgpu_ctx *ctx, const double *d_vals, const double *d_rhs, double *d_x, int rhs_border_reduced, bool defer_status = false);

int xg_fail(xgpu_ctx *c, int code, const std::string &msg);
