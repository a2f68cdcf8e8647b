// Batched / ragged float64 GEMM on CUDA cores for the small dense contractions of the functional-map
// path (k x k, N x k and k x N x d products; everything that must be float64-exact-grade).
//   C[b] (M x N) = sum_k  a(b, m, k) * b(b, n, k)
// Operands are described at run time (ragged row offsets, transposition, row gather, per-row scale).
#pragma once
#include "dm_internal.cuh"

namespace dm {

struct GemmOperand {
  const double* d = nullptr;  // float64 data, or
  const float* f = nullptr;   // float32 data
  int64_t ld = 0;
  const int64_t* off = nullptr;  // ragged: batch b owns matrix rows off[b]..off[b+1]
  const int64_t* in = nullptr;   // ragged, optional: the DATA of batch b starts at matrix row in[b] instead (mesh bank);
                                 // sizes, kscale and the output still follow `off`
  int64_t batch_stride = 0;      // else: elements between consecutive batches
  int rows = 0;                  // else: matrix rows per batch
  int col0 = 0;                  // first matrix column used
  // trans == 0: element(i, k) = Mat[row0 + i][col0 + k]      (i = output index, k = contraction index)
  // trans == 1: element(i, k) = Mat[row0 + k][col0 + i]
  int trans = 0;
  const void* gather = nullptr;  // trans == 1 only: matrix row k -> rows_of(gather_rows)[gather[goff + k]]
  int gather_i64 = 0;
  const int64_t* gather_off = nullptr;  // offsets of the gather index array (per batch), the gathered matrix uses `off`
  const int* gather_cnt = nullptr;      // optional: batch b gathers gather_cnt[b] rows (segments of fixed capacity) instead of
                                        // gather_off[b + 1] - gather_off[b]
  const double* kscale = nullptr;       // trans == 1 only: multiply by kscale[kbase + k] (kbase = gather_off or off)
};

struct GemmProblem {
  GemmOperand A, B;
  int M = 0, N = 0, K = 0;  // fixed sizes; a ragged dimension is taken from the operand offsets instead
  int maxM = 0, maxN = 0, maxK = 0;
  int n_batch = 0;
  const int* skip = nullptr;  // optional per-batch flag (device): batches with skip[b] != 0 are left untouched
  double* C = nullptr;
  int64_t ldc = 0;
  int64_t c_batch_stride = 0;      // used when the output rows are not ragged
  const int64_t* c_off = nullptr;  // ragged output rows (follows A's row offsets)
  const double* c_colscale = nullptr;  // optional: C[m][n] *= c_colscale[colscale_base + n]
  const int64_t* c_colscale_off = nullptr;
  double alpha = 1.0;
  // optional addend of the "both transposed" shape (gemm64_tt_kernel): C[b][m][n] = alpha acc + c_add[b][m][n]
  const double* c_add = nullptr;
  int64_t c_add_ld = 0, c_add_batch_stride = 0;
  int ksplit = 1;                 // number of K chunks; chunk s writes to C + s * split_stride
  int kchunk = 0;
  int64_t split_stride = 0;
};

int gemm64_launch(const GemmProblem& P, cudaStream_t st);
// out[i] = sum_s part[s * stride + i], i < n  (deterministic split-K reduction)
int sum_partials_launch(const double* part, int n_split, int64_t stride, int64_t n, double* out, cudaStream_t st);

}  // namespace dm
