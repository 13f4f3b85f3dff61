/* xyce_b200 -- C ABI of the B200-native Newton-step engine.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Xyce itself has
 * no C ABI on this path; each entry point below replaces the reference *C++ virtual* it cites
 * and is what adaptor subclasses (GpuMaster<Traits> : DeviceMaster<Traits>, GpuSolver :
 * Linear::Solver -- see INTEGRATION.md) call.  All functions return 0 on success and a
 * non-zero code otherwise (xgpu_last_error gives the text), never throw, and must be called
 * from the single host thread that owns the context (the reference's loaders are
 * single-threaded per rank).  "d_" arguments are DEVICE pointers.
 */
#ifndef XYCE_B200_H
#define XYCE_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct xgpu_ctx xgpu_ctx;

/* Per-launch constant block: the members of Device::SolverState
 * (src/DeviceModelPKG/Core/N_DEV_SolverState.h:114-217) and Device::DeviceOptions
 * (Core/N_DEV_DeviceOptions.C:79-150) that model evaluation reads.  Filled by the adaptor
 * from getSolverState()/getDeviceOptions() before every xgpu_update_state call
 * (replaces DeviceMgr::setupSolverInfo, Core/N_DEV_DeviceMgr.C:3872-3873). */
typedef struct xgpu_solver_state {
  int dcopFlag, tranopFlag, acopFlag, transientFlag, dcsweepFlag;
  int initJctFlag, initFixFlag, initTranFlag, newtonIter, locaEnabledFlag;
  int artParameterFlag, voltageLimiterFlag;
  double gmin, gainScale, nltermScale, vgstConst, vdsScaleMin, sizeScale, currTimeStep;
  double lastTimeStep;         /* SolverState::lastTimeStep_ (BJT excess phase, N_DEV_BJT.C:2768-2769) */
  int beginIntegrationFlag;    /* SolverState::beginIntegrationFlag_: first step out of a break point, t = 0 included */
} xgpu_solver_state;

/* ---- lifetime ---- */
int xgpu_create(int device, xgpu_ctx **out);
void xgpu_destroy(xgpu_ctx *ctx);
const char *xgpu_last_error(const xgpu_ctx *ctx);
/* Run all work of this context on an existing CUDA stream (cudaStream_t passed as void*). */
int xgpu_set_stream(xgpu_ctx *ctx, void *cuda_stream);
int xgpu_sync(xgpu_ctx *ctx);
/* Tuning knobs.  "b4_arith": 0 = strict arithmetic (no FMA contraction, IEEE division),
 * 1 = FMA contraction, 2 = FMA + branch-free reciprocal division (<= 2 ulp per divide; default --
 * validated against the reference at 1e-12 by tests/test_gpu_bsim4_parity.py like the others).
 * "b4_threads" / "b4_minblocks": block shape of the BSIM4 kernel (threads per block x resident blocks per
 * SM; registers per thread = 65536 / (threads * blocks), capped at 255); b4_threads = 0 (default) picks the
 * shape from the group size.  "b4_uniform": 1 (default) = model card and bin of a run of equal instances are
 * passed in the kernel parameter block, 0 = every thread loads its own records.  "b4_lockstep": 1 = block
 * barriers between evaluation sections (diagnostic variant, needs b4_uniform). */
int xgpu_set_option(xgpu_ctx *ctx, const char *name, int value);
/* Diagnostics: evaluate the fast-variant exp (which = 0), log (1) or a/b (2) on host arrays of length n. */
int xgpu_selftest_fastmath(xgpu_ctx *ctx, int which, int n, const double *h_a, const double *h_b, double *h_out);

/* ---- linear-system shape ----
 * CSR pattern shared by dFdx, dQdx and the Jacobian: the output of
 * SerialLSUtil::generateRowColData (src/TopoManagerPKG/N_TOP_SerialLSUtil.C:620-703), ground
 * rows/columns stripped, columns sorted within a row. */
int xgpu_pattern_set(xgpu_ctx *ctx, int n_unknowns, const int32_t *rowptr, const int32_t *colind);
/* Alternative to xgpu_pattern_set for callers without a Topology package: after all device
 * groups are added, derive the pattern from their Jacobian stamps exactly as
 * generateRowColData does (union of stamp entries, sorted unique columns per row, ground
 * stripped).  xgpu_pattern_get copies it out (rowptr: n+1, colind: nnz entries). */
int xgpu_pattern_build(xgpu_ctx *ctx, int n_unknowns);
int xgpu_pattern_nnz(const xgpu_ctx *ctx);
int xgpu_pattern_get(const xgpu_ctx *ctx, int32_t *rowptr, int32_t *colind);
/* Lengths of the state and store vectors (TimeIntg::DataStore, N_TIA_DataStore.h:145-300). */
int xgpu_sizes_set(xgpu_ctx *ctx, int n_state, int n_store);

/* ---- BSIM4 (level 14/54, v4.8.2) device group ----
 * Replaces Device::addModel / addInstance bookkeeping plus Instance::registerLIDs,
 * registerStateLIDs, registerStoreLIDs (N_DEV_MOSFET_B4.C:6179-6487): the adaptor passes the
 * constants Xyce's own processParams4p82_/updateTemperature4p82_ computed and the LIDs the
 * Topology package handed out.  Record field order: xgpu_b4_field_names(). */
/* which: 0 model doubles, 1 model ints, 2 size-bin doubles, 3 instance doubles, 4 instance ints */
int xgpu_b4_field_count(int which);
const char *xgpu_b4_field_names(int which);            /* newline-separated, in record order */
int xgpu_b4_models_set(xgpu_ctx *ctx, int n_models, const double *model_d, const int32_t *model_i,
                       int n_sizes, const double *size_d);
/* Adds n_inst instances (arrays are instance-major / "array of records").  lids12: the 12 node
 * LIDs {Drain,GateExt,Source,Body,DrainPrime,SourcePrime,GatePrime,GateMid,BodyPrime,SourceBody,
 * DrainBody,Charge} with collapsed nodes aliased exactly as registerLIDs does, -1 = ground.
 * Store slot s of instance i lives at sto_lid0[i] + s*sto_stride (same for state).
 * Returns the group id (>= 0) or a negative error. */
int xgpu_b4_group_add(xgpu_ctx *ctx, int n_inst, const double *inst_d, const int32_t *inst_i,
                      const int32_t *model_idx, const int32_t *size_idx, const int32_t *lids12,
                      const int32_t *sto_lid0, int sto_stride, const int32_t *sta_lid0, int sta_stride);
/* Diagnostics: which mode-specialised build of the BSIM4 kernel the group's last evaluation ran.  The uniform-record
 * kernel is compiled once per mode tuple (capMod, mobMod, igcMod, igbMod, rdsMod, dioMod, ...) listed in
 * xyce_b200/csrc/bsim4_spec_tuples.def; a group whose model cards all carry one of them (VERSION >= 4.8, default
 * topology, at most 64 model / bin runs) uses that object.  Returns the tuple id, -1 for the generic build. */
int xgpu_b4_group_spec(const xgpu_ctx *ctx, int group);
/* ---- small compact models: junction diode (type 1), MOSFET level 1 (2), Gummel-Poon BJT (3), ADMS-shaped
 * series RLC (4), ADMS-generated MVS 2.0.0 ETSOI transistor (5; N_DEV_ADMSmvs_2_0_0_etsoi.C) ----
 * One flat record per instance holding the model-card values and the temperature-adjusted instance
 * constants the reference computes in processParams/updateTemperature (field order:
 * xyce_b200/csrc/simple_fields.def), a flag word, the node LIDs in the device's own node order, and the
 * first store / state LID.  Replaces Device::addInstance + registerLIDs/StoreLIDs/StateLIDs of
 * N_DEV_Diode.C, N_DEV_MOSFET1.C, N_DEV_BJT.C and the ADMS-generated classes.  Returns the group id. */
int xgpu_simple_field_count(int type);
/* store-vector slots per instance that the type's kernel writes (diode 3: Vd, Qd, Cd; ...; translated ADMS models: their
 * output variables in registerStoreLIDs order, what Instance::updatePrimaryState copies to nextStoVector) */
int xgpu_simple_store_count(int type);
/* Models translated from admsXml output (the C++ that Xyce's `_nosac` templates emit, utils/ADMS/
 * xyceImplementationFile_nosac.xml; any of src/DeviceModelPKG/ADMS/N_DEV_ADMS*.C or a user plugin built with
 * buildxyceplugin) by xyce_b200/adms/translate.py and compiled into the library at build time.  They are small-device
 * types like the ones above: xgpu_simple_group_add with type = info5[0].
 *   info5 = {type id, unknowns per instance (nodes + branch currents, admsNodeID / admsBRA_ID order), external nodes,
 *            Jacobian stamp entries (constructor's jacobianElements order), record fields}
 *   fields = space-separated record layout: "M:x" = Model member x, "I:x" = Instance member x (after processParams /
 *            updateTemperature; admsTemperature and adms_vt_nom included when the analog block reads them)
 *   slot_row / slot_col (info5[3] entries each, may be NULL): unknown index of each stamp entry.
 * Replaces Instance::updateIntermediateVars + loadDAEFVector / loadDAEQVector / loadDAEdFdx / loadDAEdQdx of the
 * generated classes (e.g. N_DEV_ADMSmvs_2_0_0_etsoi.C:704-745, :819-1456, :1458-1532). */
int xgpu_adms_gen_count(void);
int xgpu_adms_gen_info(int idx, const char **name, const char **fields, int32_t *info5, int32_t *slot_row, int32_t *slot_col);
int xgpu_simple_group_add(xgpu_ctx *ctx, int type, int n_inst, const double *rec, const int32_t *flags,
                          const int32_t *lids, const int32_t *sto_lid0, int sto_stride,
                          const int32_t *sta_lid0, int sta_stride);
/* Builds the stamp -> CSR gather maps (replaces Instance::registerJacLIDs/setupPointers,
 * N_DEV_MOSFET_B4.C:6524-6846, and Indexor::matrixGlobalToLocal, N_TOP_Indexor.C:149-214). */
int xgpu_finalize(xgpu_ctx *ctx);
/* Store vector of the step before (DataStore::lastStoVectorRawPtr, device pointer) for the following xgpu_update_state
 * calls.  Only the Gummel-Poon BJT with excess phase (model PTF != 0) reads it: Weil's approximation keeps two history
 * values of the store entry CEXBC (Instance::oldDAEExcessPhaseCalculation1/2, N_DEV_BJT.C:2706-2799).  Without this
 * call (or with NULL) the current store stands in.  xgpu_tran_run keeps the history itself. */
int xgpu_last_store_set(xgpu_ctx *ctx, double *d_last_sto);
int xgpu_needs_last_store(const xgpu_ctx *ctx);      /* 1 when a group reads the last store (BJT with PTF != 0) */

/* Lead currents (Instance::loadLeadCurrent, set by .PRINT I(M1) / P(M1): Master::loadDAEVectors
 * N_DEV_MOSFET_B4.C:10933-10987; registerBranchDataLIDs :6482-6500).  branch_lid0[i] = li_branch_dev_id of instance i
 * (id, ig, is, ib are consecutive LIDs of the lead-current vectors; -1 = instance without lead currents).
 * xgpu_b4_lead_load, called after xgpu_update_state, ASSIGNS leadF, leadQ and junctionV at those LIDs (device
 * vectors of the DataStore: nextLeadCurrFCompRawPtr, nextLeadCurrQCompRawPtr, nextJunctionVCompRawPtr).
 * 4-terminal groups take the values from the contribution planes (they are the instance's own row terms); groups with
 * internal nodes get them from the evaluation kernel. */
int xgpu_b4_lead_set(xgpu_ctx *ctx, int group, const int32_t *branch_lid0);
int xgpu_b4_lead_load(xgpu_ctx *ctx, const double *d_sol, double *d_leadF, double *d_leadQ, double *d_junctionV);
/* The same for diode (1 branch-data entry), MOSFET level 1 (id ig is ib) and BJT (ib ie ic is) groups of
 * xgpu_simple_group_add (Master::loadDAEVectors, N_DEV_Diode.C:1889-1897, N_DEV_MOSFET1.C:4544-4572, N_DEV_BJT.C:4358-4383;
 * branch order of registerBranchDataLIDs).  xgpu_lead_load = xgpu_b4_lead_load: one call serves every group that has
 * branch LIDs set. */
int xgpu_simple_lead_set(xgpu_ctx *ctx, int group, const int32_t *branch_lid0);
int xgpu_lead_load(xgpu_ctx *ctx, const double *d_sol, double *d_leadF, double *d_leadQ, double *d_junctionV);
/* Host-buffer form (what Master::loadDAEVectors gets: leadF, leadQ, junctionV of length n_branch), evaluated at the
 * solution of the last xgpu_load_host / xgpu_load_host_jr call; entries no instance writes keep their values. */
int xgpu_lead_load_host(xgpu_ctx *ctx, int n_branch, double *h_leadF, double *h_leadQ, double *h_junctionV);

/* Carried per-instance limiter threshold (Instance::von); instance order = insertion order. */
int xgpu_b4_von_set(xgpu_ctx *ctx, int group, const double *von);
int xgpu_b4_von_get(xgpu_ctx *ctx, int group, double *von);

/* ---- hot path (device pointers) ----
 * xgpu_update_state   <- Device::updateState            (Core/N_DEV_Device.h:312-332;
 *                        DeviceMgr::updateState, Core/N_DEV_DeviceMgr.C:3857-3936)
 *   evaluates every instance of every group at d_sol, writes next store/state (and current state
 *   on the first transient Newton step) and the per-instance contributions.
 * xgpu_load_vectors   <- Device::loadDAEVectors         (N_DEV_Device.h:378-404; DeviceMgr :4156-4284)
 * xgpu_load_matrices  <- Device::loadDAEMatrices        (N_DEV_Device.h:427-450; DeviceMgr :3980-4115)
 *   accumulate != 0 keeps the "+=" contract (caller zeroed / other devices already loaded);
 *   accumulate == 0 overwrites, which saves the caller's zero-fill pass.
 * xgpu_all_converged  <- DeviceMgr::allDevicesConverged (Core/N_DEV_DeviceMgr.C:5603-5640) */
int xgpu_update_state(xgpu_ctx *ctx, const double *d_sol, double *d_next_sta, double *d_curr_sta,
                      double *d_next_sto, double *d_curr_sto, const xgpu_solver_state *ss);
int xgpu_load_vectors(xgpu_ctx *ctx, double *d_f, double *d_q, double *d_dFdxdVp, double *d_dQdxdVp,
                      int accumulate);
int xgpu_load_matrices(xgpu_ctx *ctx, double *d_dFdx, double *d_dQdx, int accumulate);
/* The three calls above in one: updateState, then loadDAEVectors and loadDAEMatrices assembled by the same three
 * launches (what NonlinearEquationLoader::loadRHS + loadJacobian ask for when both are wanted at one iterate,
 * src/LoaderServicesPKG/N_LOA_NonlinearEquationLoader.C:374-558).  Identical sums in identical order. */
int xgpu_load_dae(xgpu_ctx *ctx, const double *d_sol, double *d_next_sta, double *d_curr_sta, double *d_next_sto,
                  double *d_curr_sto, const xgpu_solver_state *ss, double *d_f, double *d_q, double *d_dFdxdVp,
                  double *d_dQdxdVp, double *d_dFdx, double *d_dQdx, int accumulate);
int xgpu_all_converged(xgpu_ctx *ctx, int *converged);
/* J = qscalar*dQdx + fscalar*dFdx  <- Matrix::linearCombo as used by OneStep::obtainJacobian
 * (N_LAS_EpetraMatrix.C:629-648, N_TIA_OneStep.C:490-495). */
int xgpu_jacobian_combine(xgpu_ctx *ctx, double qscalar, const double *d_dQdx, double fscalar,
                          const double *d_dFdx, double *d_jac);

/* ---- sparse direct solver (KLU-pattern LU) ----
 * Replaces Linear::AmesosSolver::doSolve with Amesos_Klu (N_LAS_AmesosSolver.C:216-470):
 *   xgpu_lu_analyze   <- SymbolicFactorization (:335) + the first pivoting NumericFactorization (:363):
 *                        BTF, per-block fill-reducing ordering, Gilbert-Peierls LU with threshold partial
 *                        pivoting (tol 0.001, diagonal preferred) on the HOST; fixes pattern + pivot order
 *   xgpu_lu_refactor  <- NumericFactorization with "Refactorize" (KLU_REPIVOT=0, :316-318) on the GPU
 *   xgpu_lu_solve     <- Solve (:396) on the GPU
 * d_vals: the nnz CSR values of the matrix with the pattern given to xgpu_pattern_set/_build.
 * Return codes: 0 ok, 1 structurally singular, 2 numerically singular (zero/non-finite pivot; the
 * reference then zeroes the update and warns, :372-388), other = CUDA / usage error.
 * info[8] (xgpu_lu_info): n, number of BTF blocks, largest block, nnz(L), nnz(U) incl. diagonal,
 * off-diagonal-block entries, solve levels, refactor flops. */
int xgpu_lu_analyze(xgpu_ctx *ctx, const double *d_vals);
int xgpu_lu_refactor(xgpu_ctx *ctx, const double *d_vals);
/* Symbolic result of an EXTERNAL factorization instead of xgpu_lu_analyze: what a KLU-enabled Xyce build gets from
 * klu_analyze + klu_factor through klu_extract (P, Q, block boundaries R, patterns of L and U; Amesos_Klu holds
 * them after SymbolicFactorization / NumericFactorization, N_LAS_AmesosSolver.C:335, :363) -- so that the GPU
 * refactorization and solves run on KLU's own ordering and pivot sequence.  Position t of the permuted matrix holds
 * row row_perm[t] and column col_perm[t] of A; diagonal blocks [block_ptr[b], block_ptr[b+1]); L, U in CSC over
 * positions (an explicit unit diagonal in L and any pivot position inside a U column are accepted).  row_scale
 * (NULL = none): the external solver factored diag(1 / row_scale) P A Q, factors by pivot position (klu_extract's Rs,
 * KLU scale = 1 or 2); the refactorization then divides A's values entry by entry, xgpu_lu_solve divides the
 * right-hand side.  Follow with xgpu_lu_refactor (numeric values) and xgpu_lu_solve.  3 = malformed input. */
int xgpu_lu_import(xgpu_ctx *ctx, const int32_t *row_perm, const int32_t *col_perm, int n_blocks,
                   const int32_t *block_ptr, const int32_t *Lp, const int32_t *Li, const int32_t *Up, const int32_t *Ui,
                   const double *row_scale);
/* The current plan in the same conventions (klu_extract analogue): sizes4 = {n, n_blocks, nnz(L), nnz(U)};
 * any output pointer may be NULL; Lx / Ux = values of the latest factorization on the device. */
int xgpu_lu_export_sizes(const xgpu_ctx *ctx, int32_t *sizes4);
int xgpu_lu_export(xgpu_ctx *ctx, int32_t *row_perm, int32_t *col_perm, int32_t *block_ptr, int32_t *Lp, int32_t *Li,
                   double *Lx, int32_t *Up, int32_t *Ui, double *Ux);
int xgpu_lu_solve(xgpu_ctx *ctx, const double *d_vals, const double *d_rhs, double *d_x);
int xgpu_lu_info(const xgpu_ctx *ctx, double *info8);
/* Host-only symbolic analysis + pivoting factorization + solve (no GPU, no context): the code path
 * xgpu_lu_analyze runs on the host, exposed so that CPU-only CI can test it. */
int xgpu_lu_host_factor_solve(int n, const int32_t *rowptr, const int32_t *colind, const double *vals,
                              const double *rhs, double *x, double *info8);
/* Diagonal blocks with one common symbolic pattern (every ring of a ring array, every cell of a cell array) are
 * refactored and solved as a BATCH: one thread per block, factor values interleaved across blocks, the pattern compiled
 * on the host into a straight-line program shared by all blocks (lu.h).  Host-only self check of those programs:
 * analyze on vals0, run the programs on vals1, compare with a plain left-looking refactorization / substitution.
 * out4 = {groups, batched blocks, largest relative deviation of the factors, of the block solves}. */
int xgpu_lu_host_batch_selfcheck(int n, const int32_t *rowptr, const int32_t *colind, const double *vals0,
                                 const double *vals1, double *out4);

/* ---- bordered block-diagonal solve and multi-GPU (one process per GPU, NCCL over NVLink) ----
 * Replaces Xyce's MPI "parallel load" + gathered direct solve: reverse export of ghost rows with Add
 * (N_LOA_CktLoader.C:600-601, :816-829; N_LAS_EpetraMultiVector.C:843-849, N_LAS_EpetraMatrix.C:202-208), halo import of
 * the solution (:468-470), DeviceMgr::allDevicesConverged's reduction (Core/N_DEV_DeviceMgr.C:5628).
 * Each rank holds its partition with the unknowns ordered [interior | border]; the border unknowns (shared between
 * partitions: supply rails, source branches; or dense nodes a single GPU wants out of its BTF blocks) are the LAST
 * n_border unknowns, replicated on every rank in the same order.  Devices attached only to border unknowns belong to
 * one rank; sources on border rows are replicated (the B vector is not reduced).
 *   xgpu_comm_unique_id / xgpu_comm_init: NCCL communicator of the context (the id is created on one rank and
 *     distributed by the host application: MPI_Bcast in Xyce, the store of torch.distributed in bench.py).  libnccl is
 *     loaded at run time; without it xgpu_comm_init fails and everything else works.
 *   xgpu_border_set: declares the border (after xgpu_finalize; at most 96 unknowns).  world = 1 is allowed.
 *   xgpu_shared_reduce: sums the border rows of the given vectors (NULL = skip) over the ranks, in place, on the
 *     context's stream: pack -> ncclAllReduce -> unpack, no host synchronisation.  No-op without a communicator.
 *   xgpu_border_analyze / xgpu_border_solve: the interior block A_ii is analysed / refactored with the KLU-pattern LU
 *     straight out of the full CSR values; Y = A_ii^-1 [A_is | b_i]; the rank's part of the Schur system
 *     [A_ss - A_si Y_s | b_s - A_si y] is all-reduced (ONE collective), solved redundantly by every rank, and the
 *     interior unknowns are back-substituted.  d_vals holds this rank's partial sums in the border block;
 *     rhs_border_reduced != 0: the border rows of d_rhs are already summed (replicated), else they are partial sums.
 *     Return codes as xgpu_lu_refactor.  xgpu_tran_run uses these automatically once a border is declared, and with a
 *     communicator combines the norms / convergence flags of the ranks in one small all-gather per evaluation. */
int xgpu_comm_unique_id(unsigned char *id128);
int xgpu_comm_init(xgpu_ctx *ctx, const unsigned char *id128, int rank, int world);
int xgpu_comm_info(const xgpu_ctx *ctx, int *rank, int *world);
/* Optional: peer-memory mailboxes for the small collectives (border rows, border system, norm parts -- a handful of
 * doubles each).  With them every such collective is ONE kernel that stores straight into the peers' HBM over NVLink /
 * NVSwitch and combines in rank order (a few microseconds, bitwise identical on all ranks) instead of an NCCL call
 * (15-25 us for such a message).  Call xgpu_p2p_handle on every rank (after xgpu_comm_init), exchange the 64-byte CUDA IPC
 * handles through the host application, then xgpu_p2p_attach with all of them in rank order (single node, <= 16 ranks).
 * xgpu_p2p_error returns 1 if a mailbox wait timed out (a peer did not take part in a collective). */
int xgpu_p2p_handle(xgpu_ctx *ctx, unsigned char *handle64);
int xgpu_p2p_attach(xgpu_ctx *ctx, const unsigned char *handles64_by_rank);
int xgpu_p2p_error(xgpu_ctx *ctx);
int xgpu_border_set(xgpu_ctx *ctx, int n_border);
int xgpu_border_info(const xgpu_ctx *ctx, int *n_interior, int *n_border, long long *n_global);
int xgpu_shared_reduce(xgpu_ctx *ctx, double *d_f, double *d_q, double *d_dFdxdVp, double *d_dQdxdVp);
int xgpu_border_analyze(xgpu_ctx *ctx, const double *d_vals);
int xgpu_border_solve(xgpu_ctx *ctx, const double *d_vals, const double *d_rhs, double *d_x, int rhs_border_reduced);

/* ---- linear devices, sources, Newton + transient loop (callers of the hot path) ----
 * xgpu_linear_set: constant conductance (G) and capacitance (C) stamps of the linear devices as COO
 *   triplets (duplicates are summed, entries with a -1 index = ground are dropped).  Replayed on every
 *   load exactly like the reference's FilteredMatrix pair (N_LOA_CktLoader.C:504-578, :700-788;
 *   N_LAS_FilteredMatrix.C:473-548, :632-667).  Call before xgpu_pattern_build / xgpu_finalize.
 * xgpu_sources_set: independent sources; source k adds scale[k]*s_k(t) to B[row[k]]
 *   (Vsrc: row = branch equation, scale +1, N_DEV_Vsrc.C:834-; Isrc: two entries).  type 0 = DC
 *   (params7[0]), 1 = PULSE(v1 v2 td tr tf pw per).  Evaluated on the host once per load
 *   (DeviceMgr::updateSources, Core/N_DEV_DeviceMgr.C:4900-4917).
 * xgpu_tran_run: .TRAN ... NOOP from the initial solution h_x0 -- DampedNewton (N_NLS_DampedNewton.C:362-515,
 *   :1191-1397) inside the variable-step trapezoid of OneStep (N_TIA_OneStep.C) with LTE step control.
 *   Outputs: accepted time points and probe waveforms, one record {t, h, newton iterations, order, status}
 *   per step attempt, and counters stats16 = {accepted, rejected, newton iterations, Jacobian loads,
 *   residual loads, linear solves, LU analyses, LU refactors, time points, attempts, driver rc,
 *   Newton iterations of the DC operating point, its convergence status, seconds spent in set-up (allocation,
 *   upload), seconds in the time loop, longest single wait for the per-iteration norm readback}. */
typedef struct xgpu_tran_params {
  double tstop, tstep, delmax;
  int maxNewtonStep;                 /* 0 = reference default (20) */
  double deltaXTol, absTol, relTol, RHSTol;   /* 0 = reference defaults 0.33, 1e-6, 1e-2, 1e-2 */
  double relErrorTol, absErrorTol;   /* 0 = 1e-3, 1e-6 */
  int maxOrder;                      /* 0 = 2 */
  int maxSteps;
  int method;                        /* .OPTIONS TIMEINT METHOD: 0 / 7 = trapezoid (OneStep, N_TIA_OneStep.C), 8 = Gear (N_TIA_Gear12.C) */
  int dcop;                          /* 0 = start from x0 (.TRAN ... NOOP / UIC); 1 = DC operating point from x0 first
                                        (NoTimeIntegration + DampedNewton DC_OP defaults, N_TIA_NoTimeIntegration.C:161-173, :291-298) */
} xgpu_tran_params;
int xgpu_linear_set(xgpu_ctx *ctx, int nG, const int32_t *g_row, const int32_t *g_col, const double *g_val,
                    int nC, const int32_t *c_row, const int32_t *c_col, const double *c_val);
int xgpu_sources_set(xgpu_ctx *ctx, int n_sources, const int32_t *row, const double *scale, const int32_t *type,
                     const double *params7);
/* The same by device: kind 0 Resistor {R}, 1 Capacitor {C}, 2 Inductor {L} (branch unknown: F[p] += i,
 * F[n] -= i, F[b] -= vp - vn, Q[b] += L i; N_DEV_Inductor.C:880-910, :960-985), 3 Vsrc (branch unknown; N_DEV_Vsrc.C:1323-,
 * :1454-1457), 4 ISRC (B[p] -= i(t), B[n] += i(t); N_DEV_ISRC.C:1100-1130).  nodes3[i] = {pos, neg, branch} unknown
 * indices (-1 = ground / no branch); value[i] for kinds 0-2; src_type[i] / src_params7[i] (the types of
 * xgpu_sources_set) for kinds 3-4.  Stamps are added to the replayed G / C pair, sources are appended; may be mixed
 * with xgpu_linear_set / xgpu_sources_set, which REPLACE the stamps / the source list: call those first.  Before xgpu_pattern_build. */
int xgpu_linear_devices_add(xgpu_ctx *ctx, int kind, int n, const int32_t *nodes3, const double *value,
                            const int32_t *src_type, const double *src_params7);
/* Table of the piece-wise linear sources: n_points (time, value) pairs shared by all PWL sources; a source of type 5
 * has params7 = {TD, offset (in points), count, REPEAT (0 / 1), REPEATTIME} (PWLinData, Core/N_DEV_SourceData.C:1770-1886).
 * Source types of xgpu_sources_set: 0 DC {v}, 1 PULSE {v1 v2 td tr tf pw per}, 2 SIN {v0 va freq td theta phase},
 * 3 EXP {v1 v2 td1 tau1 td2 tau2} (:811), 4 SFFM {v0 va fc mdi fs} (:2908), 5 PWL.  PULSE and PWL sources announce break
 * points (PulseData::getBreakPoints :1442, PWLinData::getBreakPoints :2044): xgpu_tran_run steps exactly onto them and
 * restarts the integration there (order 1, fresh initial step) as Transient / StepErrorControl do
 * (N_ANP_Transient.C:1900-1948, N_TIA_StepErrorControl.C:370-420, :742-850); a PULSE source also caps the step at a
 * tenth of its period (PulseData::getMaxTimeStepSize :1518). */
int xgpu_sources_pwl_set(xgpu_ctx *ctx, int n_points, const double *tv_pairs);
/* Host-only views of the source routines (no GPU, no context), for tests: value at time t, break points announced at t
 * (returns their number, writes at most max_out), step cap at t. */
double xgpu_source_value(int type, const double *params7, const double *pwl_tv_pairs, double t, double bp_tol);
int xgpu_source_breakpoints(int type, const double *params7, const double *pwl_tv_pairs, double t, int max_out, double *out);
double xgpu_source_max_step(int type, const double *params7, double t);
int xgpu_tran_run(xgpu_ctx *ctx, const xgpu_tran_params *tp, const double *h_x0, int n_probes,
                  const int32_t *probes, int max_out, int *n_out, double *h_times, double *h_wave,
                  int max_steps_out, int *n_steps_out, double *h_step_info5, double *stats16);

/* ---- host-buffer convenience path (what a non-GPU-aware caller uses; copies inside) ----
 * One updateState + loadDAEVectors + loadDAEMatrices pass.  Vectors have n_unknowns entries,
 * matrices nnz entries; next/curr store and state live in the context between calls. */
int xgpu_load_host(xgpu_ctx *ctx, const double *h_sol, const xgpu_solver_state *ss, double *h_f, double *h_q,
                   double *h_dFdxdVp, double *h_dQdxdVp, double *h_dFdx, double *h_dQdx);
/* The same pass returning only what a host-side Newton solver consumes (4 MB instead of 9.6 MB of PCIe traffic on
 * BASELINE config 2): the combined Jacobian  J = qscalar dQdx + fscalar dFdx  (OneStep::obtainJacobian,
 * N_TIA_OneStep.C:490-495: qscalar = -alphas/h, fscalar = 1 or 1/2; Gear12: a0/h, 1) and the device part of the residual
 *   r = -(qscalar Q + fscalar F) + qscalar dQdxdVp + fscalar dFdxdVp        (limiter terms when ss->voltageLimiterFlag)
 * to which the caller adds the terms that live in its own history vectors (OneStep::obtainResidual :223-272:
 * + qscalar qHistory[0] + fscalar B (- 1/2 qHistory[2] at order 2)).  h_rhs: n entries, h_jac: nnz entries.  With option
 * "zero_copy_out" = 1 and pinned, mapped buffers the kernel stores straight into host memory. */
int xgpu_load_host_jr(xgpu_ctx *ctx, const double *h_sol, const xgpu_solver_state *ss, double qscalar, double fscalar,
                      double *h_rhs, double *h_jac);
/* One whole Newton iteration behind host buffers: what NonlinearSolver::rhs_() + jacobian_() + newton_() do between two
 * solution vectors (N_NLS_DampedNewton.C:362-515: Loader::loadRHS + loadJacobian, then Linear::Solver::solve on
 * J dx = -r).  In: x (n), the time integrator's scalars as in xgpu_load_host_jr, and optionally h_hist (n) = the part of
 * the residual that lives in the caller's history vectors (qscalar qHistory[0] - fscalar B ...; null = none).
 * Out: dx (n) and optionally the right-hand side r = -(qscalar Q + fscalar F + hist) + limiter terms (n; null = not
 * wanted).  PCIe traffic is 8n bytes each way (0.8 MB on BASELINE config 2 instead of the 4 MB of J + r); evaluation,
 * assembly, linear-device replay, refactorization on the previous call's pivot sequence and the triangular solves run
 * back to back on the device with ONE host synchronisation at the end.  The first call (and any call whose pivots fail
 * the threshold check, or every call under "lu_repivot") analyses on the host like xgpu_lu_analyze. */
int xgpu_newton_step_host(xgpu_ctx *ctx, const double *h_sol, const xgpu_solver_state *ss, double qscalar, double fscalar,
                          const double *h_hist, double *h_dx, double *h_rhs);
/* xgpu_load_host_jr pipelines when the circuit allows it (option "pipeline_host", default 1): every (model, bin) run of
 * the BSIM4 group is evaluated in two parts ("pipe_percent" of its length first, default 50, set before xgpu_finalize);
 * the rows / nonzeros that only the first part feeds -- a prefix for circuits numbered device by device, found in
 * xgpu_finalize -- are assembled, combined and copied over PCIe on a second stream while the second part is evaluated.
 * Same sums in the same order: bitwise the one-pass result.  info3 = {usable for this circuit, rows in the first window,
 * nonzeros in the first window}. */
int xgpu_pipe_info(const xgpu_ctx *ctx, long long *info3);
/* which: 0 next store, 1 curr store, 2 next state, 3 curr state, 4 last store (no-op unless xgpu_needs_last_store) */
int xgpu_state_set(xgpu_ctx *ctx, int which, const double *h_vals);
int xgpu_state_get(xgpu_ctx *ctx, int which, double *h_vals);
/* Device pointers of the context-owned buffers used by xgpu_load_host:
 * 0 sol, 1 f, 2 q, 3 dFdxdVp, 4 dQdxdVp, 5 dFdx, 6 dQdx, 7 next sto, 8 curr sto, 9 next sta, 10 curr sta,
 * 11 last sto (one element unless a BJT group with excess phase is present) */
double *xgpu_device_buffer(xgpu_ctx *ctx, int which);

/* Measured FP64 FMA throughput of this GPU (dependent-chain DFMA microbenchmark, TFLOP/s with
 * FMA = 2 flops): the denominator for FP64-pipe roofline fractions. */
int xgpu_measure_fp64_peak(xgpu_ctx *ctx, double *tflops);

/* Kernel launches issued by this context since creation (bench bookkeeping). */
long long xgpu_launch_count(const xgpu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* XYCE_B200_H */
