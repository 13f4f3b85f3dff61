// xyce_b200 -- sparse LU: shared data structures of the host symbolic phase (lu_host.cpp) and the
// GPU refactor / solve kernels (lu.cu).
#pragma once
#include <cstdint>
#include <vector>

namespace xb {
namespace lu {

// Permuted system  B = P A Q  (position t holds row row_perm[t] and column col_perm[t] of A) is upper
// block triangular with diagonal blocks [block_ptr[b], block_ptr[b+1]).  Each diagonal block is
// factored L U with L unit lower (diagonal not stored) and U upper, both CSC over positions; a U
// column stores its off-diagonal entries in ascending row order followed by the pivot as its LAST
// entry.  Entries of A outside the diagonal blocks stay unfactored ("off", by column).
struct LuPlan {
  int n = 0;
  int nnz_a = 0;
  bool structurally_singular = false;
  bool singular = false;
  std::vector<int> row_perm, col_perm, block_ptr;
  std::vector<int> Lp, Li, Up, Ui;
  std::vector<double> Lx, Ux;               // values of the first (pivoting) factorization
  // numeric scatter: column t of B inside its block = { (acol_row[q], A.values[acol_src[q]]) }
  std::vector<int> acol_ptr, acol_row, acol_src;
  // off-diagonal entries of column t: rows off_row[q] (in earlier blocks), values A.values[off_src[q]]
  std::vector<int> off_ptr, off_row, off_src;
  std::vector<double> off_val_host;
  // the same entries by ROW position (pull form used by the GPU solve): row r needs
  // sum_q A.values[offr_src[q]] * y[offr_col[q]] over q in [offr_ptr[r], offr_ptr[r+1])
  std::vector<int> offr_ptr, offr_col, offr_src;
  // block levels for the solve: blocks in level l depend only on blocks in levels < l
  std::vector<int> level_ptr, level_blocks;
  // rows (positions) of each level that own off-diagonal entries, split by list length:
  // short rows are pulled by one warp each, long rows (supply rails) by chunked block reductions
  std::vector<int> pull_tiny_ptr, pull_tiny_rows;        // per level: rows with at most 8 entries (one thread each)
  std::vector<int> pull_short_ptr, pull_short_rows;      // per level
  std::vector<int> pull_long_ptr, pull_long_rows;        // per level
  std::vector<int> pull_chunk_ptr, pull_chunk_row_slot, pull_chunk_begin;   // per level: chunks of the long rows
  std::vector<int> pull_long_chunk_ptr;                  // per long row slot: its chunk range
  double refactor_flops = 0.0;
  // optional row scaling of an imported factorization (KLU scale > 0): the factored matrix is diag(1 / row_scale) P A Q,
  // row_scale by pivot position; nz_rowpos[k] = pivot position of the row of A's k-th CSR entry.  Empty = none.
  std::vector<double> row_scale;
  std::vector<int> nz_rowpos;

  // ---- large diagonal blocks (more than kBigBlock rows): column-level / row-level schedules ----
  // A single warp per block (the small-block kernels) would walk such a block column by column; instead
  //   refactor: columns are grouped by dependency level (column k needs every column i with U(i,k) != 0);
  //             one warp factors one column in place (values addressed through binary search in the fixed
  //             pattern), all columns of a level in one launch.  Columns whose U part is longer than
  //             kDenseCol (supply rails) are computed as a forward substitution over the whole block instead.
  //   solve:    L and U of the block in row form; rows grouped by level, one warp (or one CTA for rows longer
  //             than kLongRow) per row.
  std::vector<int> block_big;                   // [nblocks] 0 = medium (warp per block, global indices), 1 = large, 2 = staged small block
  int staged_bytes = 0;                         // shared-memory slice per warp of the staged kernels
  std::vector<int> acol_dst;                    // per acol entry of a large-block column: >= 0 index into Ux, < 0 ~index into Lx
  std::vector<int> rf_level_ptr, rf_cols;       // normal columns of all large blocks by level
  std::vector<int> rf_dense_ptr, rf_dense_cols; // dense columns by level (processed one at a time after the level's normal columns)
  std::vector<int> Lr_ptr, Lr_col, Lr_src;      // [n+1] row form of L (rows of large blocks only; others empty), src indexes Lx
  std::vector<int> Ur_ptr, Ur_col, Ur_src;      // row form of strict U, src indexes Ux
  // stages: every stage is a set of independent rows; short rows (one warp each) and long rows (one CTA each)
  std::vector<int> fs_short_ptr, fs_short_rows, fs_long_ptr, fs_long_rows;   // forward (L) stages, all large blocks back to back
  std::vector<int> bs_short_ptr, bs_short_rows, bs_long_ptr, bs_long_rows;   // backward (U) stages
  std::vector<int> big_blocks;                  // ids of the large blocks
  std::vector<int> big_fs_begin, big_fs_end, big_bs_begin, big_bs_end;       // per large block: its stage ranges

  // ---- batched blocks (block_big == 3): diagonal blocks that share ONE symbolic pattern (same size, same L / U
  // pattern, same positions of the A entries -- every ring of a ring-oscillator array, every cell of a cell array) ----
  // Such a group is factored and solved with one THREAD per block: the factor values of the group are stored
  // interleaved, slot-major with the block index fastest (coalesced), the pattern is compiled once on the host into
  // a straight-line "elimination program" over factor slots that is the same for every block.  Independent operations
  // are packed into bundles of kBundle; a TILE of kBundle lanes owns one block and executes one bundle per step, one
  // operation per lane (all loads of the bundle, then the arithmetic, then the stores, then a warp barrier), so the
  // shared-memory latency of a column chain is paid once per bundle and a warp (32 / kBundle blocks) runs in lock
  // step with uniform program words.
  struct BatchGroup {
    int nb = 0, nu = 0, nl = 0, na = 0, level = 0;      // block size, U / L entries, A entries per block, solve level
    std::vector<int> blocks;                            // block ids
    std::vector<int> a_dst;                             // [na] factor slot of the e-th A entry (U slots [0, nu), L slots [nu, nu + nl))
    std::vector<int> a_src;                             // [na][nblk] CSR value index, block fastest
    std::vector<unsigned short> rf_prog, sv_prog;       // bundles of kBundle ops, op = {dst | type << 14, a, b, 0} (8 bytes)
    int rf_bundles = 0, sv_bundles = 0;
  };
  std::vector<BatchGroup> batch;
};
constexpr int kBigBlock = 512, kDenseCol = 4096, kLongRow = 2048, kStagedBytes = 12288;
constexpr int kBundle = 4;                 // operations per bundle of a batched-group program
constexpr int kBatchMinBlocks = 16;        // fewer equal blocks than this stay on the warp-per-block kernels
constexpr int kBatchMaxSlots = 3000;       // nu + nl + nb of a batched pattern (8 blocks per CTA must fit shared memory)
// op types of the batched programs (high 2 bits of the first word)
//   refactor: 0  v[dst] -= v[a] * v[b]      1  v[dst] = v[dst] / v[a] (+ pivot threshold test)    2  pivot check of v[a]    3 no-op
//   solve   : 0  y[dst] -= v[a] * y[b]      1  y[dst] = y[dst] / v[a]                                                      3 no-op
enum BatchOp { kOpFnma = 0, kOpDiv = 1, kOpChk = 2, kOpNop = 3 };

// Symbolic analysis + first numeric factorization with threshold partial pivoting (KLU defaults:
// pivot_tol = 0.001, diagonal preferred).  Returns 0 ok, 1 structurally singular, 2 numerically singular.
// val_index (optional): the pattern given here is a SUB-pattern of a larger CSR matrix and val_index[k] is the position
// of its k-th entry in that matrix's value array; `vals` is then the larger array, and the plan's scatter maps address
// it directly (bordered solve: the interior block is factored straight out of the full Jacobian values).
int analyze_and_factor(int n, const int *rowptr, const int *colind, const double *vals, double pivot_tol,
                       LuPlan &plan, const int *val_index = nullptr);
// batched groups on / off for the plans built from now on (option "lu_batch"; default on)
void set_batching(bool on);
void solve_host(const LuPlan &plan, const double *b, double *x);
void batch_selfcheck_host(const LuPlan &plan, const double *vals, double *out4);
// Plan from an external factorization's permutations, block boundaries and L / U patterns (lu_host.cpp);
// numeric values come from the first refactorization on the GPU.  0 ok, 3 malformed input.
int import_factorization(int n, const int *rowptr, const int *colind, const int *row_perm, const int *col_perm,
                         int nblocks, const int *block_ptr, const int *Lp, const int *Li, const int *Up, const int *Ui,
                         const double *row_scale, LuPlan &plan, const char **why);

// Device-resident copy of a plan plus work space; created by upload_plan, used by the kernels.
// LuView is the part the kernels see (plain pointers and sizes: it is the kernel parameter, passed by value);
// LuDev adds the host-side launch schedule.
struct LuView {
  int n = 0, nblocks = 0, nlevels = 0, staged_bytes = 0, staged_warps = 4;
  int *row_perm = nullptr, *col_perm = nullptr, *block_ptr = nullptr;
  int *Lp = nullptr, *Li = nullptr, *Up = nullptr, *Ui = nullptr;
  double *Lx = nullptr, *Ux = nullptr;
  int *acol_ptr = nullptr, *acol_row = nullptr, *acol_src = nullptr;
  int *offr_ptr = nullptr, *offr_col = nullptr, *offr_src = nullptr;
  int *level_blocks = nullptr;
  int *pull_tiny_rows = nullptr;
  int *pull_short_rows = nullptr, *pull_long_rows = nullptr, *pull_chunk_row_slot = nullptr, *pull_chunk_begin = nullptr;
  int *pull_long_chunk_ptr = nullptr;
  double *pull_partials = nullptr;
  double *work = nullptr;         // [n] dense column / solution work vector
  double *work2 = nullptr;        // [n] dense-column work vector of the large-block refactor
  int *status = nullptr;          // device flags: bit 0 = zero or non-finite pivot, bit 2 = a pivot failed the threshold test
  // Pivot monitor of the refactorization (fixed pivot sequence): KLU's partial pivoting accepts a pivot only when
  // |pivot| >= pivot_tol * max |candidate| of its column; a refactorization on an old pivot sequence (klu_refactor,
  // KLU_REPIVOT=0) never re-tests that.  With pivot_check the kernels do, on the candidates they divide anyway, and
  // report it (bit 2) so that the caller re-analyses = re-pivots exactly when the reference's default
  // (KLU_REPIVOT=1: pivoting factorization every time, N_LAS_AmesosSolver.C:316-318) would have chosen differently.
  double pivot_tol = 0.001;
  int pivot_check = 1;
  double *row_scale = nullptr, *As = nullptr;   // row scaling (imported plans): factors by position, scaled copy of A's values
  int *nz_rowpos = nullptr; int nnz_a = 0;
  // large blocks (see LuPlan)
  int *block_big = nullptr, *acol_dst = nullptr, *rf_cols = nullptr;
  int *Lr_ptr = nullptr, *Lr_col = nullptr, *Lr_src = nullptr, *Ur_ptr = nullptr, *Ur_col = nullptr, *Ur_src = nullptr;
  int *fs_short_rows = nullptr, *fs_long_rows = nullptr, *bs_short_rows = nullptr, *bs_long_rows = nullptr;
};
struct LuBatchDev {
  int nblk = 0, nb = 0, nu = 0, nl = 0, na = 0, level = 0, rf_bundles = 0, sv_bundles = 0;
  int lanes_rf = 8, lanes_sv = 8;        // blocks per one-warp CTA (each owned by a tile of kBundle lanes; at most 32 / kBundle) chosen at upload
  int *k0 = nullptr;                     // [nblk] first position of each block
  int *a_dst = nullptr;                  // [na] factor slot of each A entry
  int *a_src_cta = nullptr;              // CSR value indices regrouped per CTA of the refactor kernel: [cta][entry][block in CTA]
  int smem_rf = 0, smem_sv = 0;          // dynamic shared memory of the two kernels
  unsigned short *rf_prog = nullptr, *sv_prog = nullptr;
  double *LUx = nullptr;                 // [nu + nl][nblk] factor values, block fastest
  int *Up0 = nullptr, *Lp0 = nullptr;    // [nblk] first U / L entry of each block in the ordinary factor arrays (export)
};
struct LuDev : LuView {
  std::vector<LuBatchDev> batch;
  std::vector<int> level_ptr;     // host copy: one launch per level
  std::vector<int> pull_tiny_ptr;
  std::vector<int> pull_short_ptr, pull_long_ptr, pull_chunk_ptr;   // host copies
  std::vector<int> rf_level_ptr, rf_dense_ptr, rf_dense_cols, fs_short_ptr, fs_long_ptr, bs_short_ptr, bs_long_ptr;
  std::vector<int> big_blocks, big_fs_begin, big_fs_end, big_bs_begin, big_bs_end, block_ptr_h, block_level_of_big;
  std::vector<int> dense_col_block;   // block id of every dense column (parallel to rf_dense_cols)
};

}  // namespace lu
}  // namespace xb

#ifdef __CUDACC__
#include <cuda_runtime.h>
namespace xb {
namespace lu {
cudaError_t upload_plan(const LuPlan &p, LuDev &d);
void free_plan(LuDev &d);
int launch_refactor(const LuDev &d, const double *d_A, cudaStream_t s);                       // returns #launches
// copies the factor values of the batched groups into the ordinary Lx / Ux arrays (export / diagnostics only)
int launch_batch_export(const LuDev &d, cudaStream_t s);
int launch_solve(const LuDev &d, const double *d_A, const double *d_rhs, double *d_x, cudaStream_t s);
}  // namespace lu
}  // namespace xb
#endif
