// xyce_b200 -- device-side data layout and launch interface of the BSIM4 group kernels.
#pragma once
#include <cuda_runtime.h>
#include "bsim4_instance.h"
#include "bsim4_spec_tuples.def"

namespace xb {
namespace b4 {

// Packed per-instance topology word (replaces the 7 int members of B4Inst in HBM).
XB_HD int pack_topo(const B4Inst &I) {
  return (I.rgateMod & 3) | ((I.rbodyMod & 3) << 2) | ((I.trnqsMod & 1) << 4) | ((I.acnqsMod & 1) << 5) |
         ((I.OFF & 1) << 6) | ((I.drainMOSFET_B4Exists & 1) << 7) | ((I.sourceMOSFET_B4Exists & 1) << 8);
}
XB_HD void unpack_topo(int w, B4Inst &I) {
  I.rgateMod = w & 3; I.rbodyMod = (w >> 2) & 3; I.trnqsMod = (w >> 4) & 1; I.acnqsMod = (w >> 5) & 1;
  I.OFF = (w >> 6) & 1; I.drainMOSFET_B4Exists = (w >> 7) & 1; I.sourceMOSFET_B4Exists = (w >> 8) & 1;
}
constexpr int kDefaultTopoMask = 0x1bf;   // every bit except OFF must be zero for the 4-terminal fast path

// Contribution layout of one instance group inside the assembly planes.
//   default topology: 4 rows, 16 slots (slot = 4*row + col over {D,G,S,B})
//   general topology: 11 rows, 62 slots (B4Slot order)
constexpr int kRowsDefault = 4, kSlotsDefault = 16;
constexpr int kRowsGeneral = kNumRows, kSlotsGeneral = kNumSlots;

// One group of BSIM4 instances evaluated by one launch.  All pointers are device pointers.
struct GroupDev {
  int n;                    // instances
  int general;              // 0: default-topology fast path, 1: general stamp
  const B4Model *models;    // [n_models]
  const B4Size *sizes;      // [n_sizes]
  const double *inst_d;     // [kNumInstD][n]  structure of arrays
  const int *topo;          // [n] packed topology word
  const int *model_idx;     // [n]
  const int *size_idx;      // [n]
  const int *lids;          // [4][n] (default) or [12][n] (general); -1 = ground
  const int *sto_lid0;      // [n] LID of store slot 0
  const int *sta_lid0;      // [n] LID of state slot 0
  int sto_stride;           // LID distance between consecutive store slots of one instance
  int sta_stride;
  double *von;              // [n] carried limiter threshold (Instance::von)
  int *orig_flag;           // [n] Instance::isConverged(): 1 unless pnjlim limited a junction voltage in this evaluation
  // contribution planes (see assembly.cuh): element (row r, instance i) of this group lives at
  // vec_base + r*n + i inside each of the 4 vector planes; slot s at mat_base + s*n + i.
  long long vec_base, mat_base;
  // general-topology groups with lead currents requested: [8][n] = leadF(id, ig, is, ib), leadQ(id, ig, is, ib); else null
  double *lead;
};

struct LoadArgs {
  SolverFlags S;
  const double *sol;        // next solution (length >= n_unknowns)
  double *next_sto, *curr_sto;
  double *last_sto;         // store vector of the step before (BJT excess phase only; = curr_sto when the caller keeps no history)
  double *next_sta, *curr_sta;
  double *vec_planes[4];    // F, Q, dFdxdVp, dQdxdVp contribution planes
  double *mat_planes[2];    // dFdx, dQdx contribution planes
};

// Runs of instances that share one (model card, bin) pair, passed by value in the kernel parameter block.
struct BinRun {
  B4Model M;
  B4Size P;
  int start, count;         // instance range [start, start + count) inside the group
};
constexpr int kRunsPerPack = 8;
struct BinPack {
  BinRun run[kRunsPerPack];
  int nruns;
};
constexpr int kMaxUniformRuns = 64;   // groups with more distinct runs use the per-thread-record kernel

// (threads per block, resident blocks per SM) shapes compiled for the default-topology kernel;
// registers per thread = 65536 / (threads * blocks), capped at 255.
#if defined(XB_B4_EXTRA)       // experiment builds only (scripts/build_variant.sh)
#define XB_B4_EXTRA_SHAPES(X) X(128, 5) X(128, 6) X(352, 1)
#else
#define XB_B4_EXTRA_SHAPES(X)
#endif
#if defined(XB_B4_FEW_SHAPES)      // additional mode-specialised objects: only the two shapes the launcher picks by itself
#define XB_B4_LAUNCH_SHAPES(X) X(128, 3) X(128, 4)
#else
#define XB_B4_LAUNCH_SHAPES(X) \
  X(64, 4) X(64, 6) X(96, 4) X(128, 2) X(128, 3) X(128, 4) X(256, 1) X(384, 1) X(512, 1) XB_B4_EXTRA_SHAPES(X)
#endif
// (20-24 warps per SM -- 128 x 5, 128 x 6, 64 x 11 at 80-96 registers -- were measured and are slower at every
// group size: the local-memory spills cost more than the occupancy gains; profiles/r01_b4_occupancy.json)

// arith: 0 exact (no FMA contraction, IEEE division), 1 fma, 2 fast (fma + reciprocal division);
// lockstep: block-wide barriers between evaluation sections keep the warps of a block inside the same
// instruction-cache window (arith 2 only); packs != nullptr selects the uniform-record kernel.
// Returns the number of kernel launches, -1 for an unsupported shape.
int launch_b4_group_a0(const GroupDev &g, const LoadArgs &a, int threads, int minblocks, const BinPack *packs, int npacks, cudaStream_t stream);
int launch_b4_group_a1(const GroupDev &g, const LoadArgs &a, int threads, int minblocks, const BinPack *packs, int npacks, cudaStream_t stream);
int launch_b4_group_a2(const GroupDev &g, const LoadArgs &a, int threads, int minblocks, const BinPack *packs, int npacks, cudaStream_t stream);
int launch_b4_group_a2s(const GroupDev &g, const LoadArgs &a, int threads, int minblocks, const BinPack *packs, int npacks, cudaStream_t stream);
// mode-specialised builds (scripts/gen_spec.py, one object per tuple of bsim4_spec_tuples.def): only for groups whose
// model cards all carry that tuple
int launch_b4_group_a2x(const GroupDev &g, const LoadArgs &a, int threads, int minblocks, const BinPack *packs, int npacks, cudaStream_t stream);
int launch_b4_group_a2x1(const GroupDev &g, const LoadArgs &a, int threads, int minblocks, const BinPack *packs, int npacks, cudaStream_t stream);
int launch_b4_group_a2x2(const GroupDev &g, const LoadArgs &a, int threads, int minblocks, const BinPack *packs, int npacks, cudaStream_t stream);
// mode sets of the specialised objects in XB_B4_MODEL_I order; -2 = not specialised (dtype)
#define XB_SPEC_ROW(id, ...) {__VA_ARGS__},
constexpr int kSpecModes[kNumSpecTuples][17] = {XB_B4_SPEC_TUPLES(XB_SPEC_ROW)};
#undef XB_SPEC_ROW
// spec_id: -1 = generic build, else the tuple's id
inline int launch_b4_group(const GroupDev &g, const LoadArgs &a, int arith, int lockstep, int threads, int minblocks,
                           const BinPack *packs, int npacks, cudaStream_t stream, int spec_id = -1) {
  if (arith == 2 && spec_id >= 0 && !lockstep && !g.general && packs) {
    if (spec_id == 0) return launch_b4_group_a2x(g, a, threads, minblocks, packs, npacks, stream);
    if (spec_id == 1) return launch_b4_group_a2x1(g, a, threads, minblocks, packs, npacks, stream);
    if (spec_id == 2) return launch_b4_group_a2x2(g, a, threads, minblocks, packs, npacks, stream);
  }
  if (arith == 2) return lockstep ? launch_b4_group_a2s(g, a, threads, minblocks, packs, npacks, stream)
                                  : launch_b4_group_a2(g, a, threads, minblocks, packs, npacks, stream);
  if (arith == 1) return launch_b4_group_a1(g, a, threads, minblocks, packs, npacks, stream);
  return launch_b4_group_a0(g, a, threads, minblocks, packs, npacks, stream);
}

}  // namespace b4
}  // namespace xb
