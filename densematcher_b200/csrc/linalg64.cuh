// Small batched float64 dense linear algebra for the spectral ICP step (one CTA per mesh pair):
// SPD inverse through a Cholesky factorisation, and the orthogonal polar factor U I V^T of a k2 x k1
// matrix through a one-sided (Hestenes) Jacobi SVD.
#pragma once
#include "dm_internal.cuh"

namespace dm {

// scratch (in doubles) the two routines need per batch entry when the matrices do not fit in shared memory
size_t spd_inverse_scratch_doubles(int n);
size_t polar_scratch_doubles(int rows, int cols);

// Ginv[b] = G[b]^-1 for n_batch symmetric positive definite n x n matrices (row-major, dense, contiguous).
// status[0] is set to 1 if a pivot is not positive.  scratch: n_batch * spd_inverse_scratch_doubles(n).
int spd_inverse_launch(const double* G, double* Ginv, int n, int n_batch, double* scratch, int* status,
                       cudaStream_t st);

// C[b] = U I V^T where X[b] = U S V^T (rows x cols, row-major, contiguous): the nearest (partial) isometry.
// scratch: n_batch * polar_scratch_doubles(rows, cols) (Jacobi);  ns_scratch: polar_ns_scratch_doubles(...) doubles
// for the Newton-Schulz fast path (nullptr = Jacobi only).  X and C must not alias when ns_scratch is given.
size_t polar_ns_scratch_doubles(int rows, int cols, int n_batch);
int polar_factor_launch(const double* X, double* C, int rows, int cols, int n_batch, double* scratch, double* ns_scratch,
                        cudaStream_t st);

}  // namespace dm
