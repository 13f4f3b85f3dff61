// xyce_b200 -- BLAS-1 style kernels of the Newton / time-integration layer on device vectors
// (what Epetra does for the reference: N_LAS_MultiVector.h:86-181; norms N_LAS_EpetraMultiVector.C:599-680).
// All reductions are two-stage with fixed shapes (bitwise reproducible).  HBM bound: 8-24 B per element.
#pragma once
#include <cuda_runtime.h>
namespace xb {
namespace vec {
void fill(double *d, double v, int n, cudaStream_t s);
void axpby(double *dst, double a, const double *x, double b, const double *y, int n, cudaStream_t s);   // dst = a x + b y
void sol_weights(double *dst, double rel, double abs, const double *a, const double *b, int n, cudaStream_t s);
void abs_weights(double *dst, double rel, double abs, const double *a, int n, cudaStream_t s);
enum Reduce { kSumSq = 0, kMaxAbs = 1, kWMaxAbs = 2, kWSumSq = 3 };
// returns the reduced scalar on the host (synchronises the stream); scratch: >= 1024 doubles of device memory
double reduce(Reduce mode, const double *x, const double *w, int n, double *scratch, cudaStream_t s);
// same reduction, result left in scratch[0] on the device (no copy, no synchronisation): multi-GPU norms combine the
// ranks' parts first
void reduce_dev(Reduce mode, const double *x, const double *w, int n, double *scratch, cudaStream_t s);
// Newton residual + everything the convergence test reads, two launches and no host round trip of their own:
//   form 0: rhs = -[(q - qh0) inv_h + fs f + (-fs) b (+ 0.5 qh2)] (+ qlim_coef qlim + fs flim)    (operation order of the
//   axpby sequence of OneStep::obtainResidual, so results are bit-identical to it)
//   form 1: rhs = -[(a0 q + a1 qh0 (+ a2 qh1)) inv_h + f - b] (+ qlim_coef qlim + flim)           (Gear12::obtainResidual)
//   form 2: rhs = -(f - b) (+ flim)                                                                (NoTimeIntegration)
//   out4[0] = sum rhs^2, out4[1] = max |rhs|, out4[2] = max |dx / w|, out4[3] = 1 if every flag in the listed int
//   arrays is non-zero else 0.  scratch: >= 3 * 1024 + 8 doubles.
struct ResidualArgs {
  double *rhs; const double *q, *qh0, *qh1, *f, *b, *qh2, *qlim, *flim, *dx, *w;
  double inv_h, fs, qlim_coef, a0, a1, a2; int form, order2, limiter, n;      // form: 0 OneStep, 1 Gear12, 2 DC (tran_driver.h)
  const int *flags[8]; int flag_n[8]; int nflag_arrays;
};
void residual_norms(const ResidualArgs &a, double *scratch, double *out4, cudaStream_t s);

// CSR y += A x restricted to the stored rows (linear-device replay, FilteredMatrix::axpy N_LAS_FilteredMatrix.C:473-548)
void spmv_add(int nrows, const int *rows, const int *ptr, const int *col, const double *val, const double *x,
              double *y, cudaStream_t s);
// vals[pos[k]] += v[k]  (FilteredMatrix::addToMatrix, N_LAS_FilteredMatrix.C:632-667); pos entries are unique
void scatter_add(int n, const int *pos, const double *v, double *vals, cudaStream_t s);
}  // namespace vec
}  // namespace xb
