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
// CSR y += A x restricted to the stored rows (linear-device replay, FilteredMatrix::axpy N_LAS_FilteredMatrix.C:473-548)
void spmv_add(int nrows, const int *rows, const int *ptr, const int *col, const double *val, const double *x,
              double *y, cudaStream_t s);
// vals[pos[k]] += v[k]  (FilteredMatrix::addToMatrix, N_LAS_FilteredMatrix.C:632-667); pos entries are unique
void scatter_add(int n, const int *pos, const double *v, double *vals, cudaStream_t s);
}  // namespace vec
}  // namespace xb
