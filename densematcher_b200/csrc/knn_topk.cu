// k nearest neighbours (k > 1) of every query row, float64, single pair: the `k` argument of knn_query
// (densematcher/pyFM/spectral/nn_utils.py:4-38; the reference's caller with k > 1 is projection_utils.py:178).
// Not on the throughput path: scores S = Y X^T come from the float64 DMMA GEMM in chunks of query rows, one warp per
// query keeps the k smallest |x_j|^2 - 2 S_ij of its lane-strided candidates in registers, the lanes' lists are merged
// by k rounds of a warp arg-min (lowest index on ties), and the k distances are re-evaluated directly as |y - x|
// (no cancellation) before the final ordering by (distance, index), which is what the kd-tree returns.
#include "dm_internal.cuh"
#include "gemm64.cuh"

namespace dm {
namespace {
constexpr int kMaxK = 16;
constexpr int kChunkRows = 2048;

__global__ void __launch_bounds__(256)
    knn_topk_kernel(const double* __restrict__ S, int64_t ldS, const double* __restrict__ Y, int64_t ldY,
                    const double* __restrict__ X, int64_t ldX, const double* __restrict__ xsq, int row0, int rows, int ndb,
                    int d, int k, int64_t* __restrict__ idx_out, double* __restrict__ dist_out) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const double* s = S + int64_t(r) * ldS;
  double bv[kMaxK];
  int bi[kMaxK];
#pragma unroll
  for (int t = 0; t < kMaxK; ++t) bv[t] = INFINITY, bi[t] = 0x7fffffff;
  for (int j = lane; j < ndb; j += 32) {  // ascending j: strict '<' keeps the lowest index among equals
    const double v = xsq[j] - 2.0 * s[j];
    if (v < bv[kMaxK - 1]) {
      bv[kMaxK - 1] = v, bi[kMaxK - 1] = j;
#pragma unroll
      for (int t = kMaxK - 1; t > 0; --t) {
        if (bv[t] < bv[t - 1]) {
          const double tv = bv[t]; bv[t] = bv[t - 1]; bv[t - 1] = tv;
          const int ti = bi[t]; bi[t] = bi[t - 1]; bi[t - 1] = ti;
        }
      }
    }
  }
  // k rounds: the lane with the smallest head (value, then index) wins and pops it
  int sel = 0x7fffffff;
  for (int round = 0; round < k; ++round) {
    double hv = bv[0];
    int hi = bi[0];
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, hv, sh);
      const int oi = __shfl_xor_sync(0xffffffffu, hi, sh);
      if (ov < hv || (ov == hv && oi < hi)) hv = ov, hi = oi;
    }
    if (bi[0] == hi && hi != 0x7fffffff) {  // pop
#pragma unroll
      for (int t = 0; t < kMaxK - 1; ++t) bv[t] = bv[t + 1], bi[t] = bi[t + 1];
      bv[kMaxK - 1] = INFINITY, bi[kMaxK - 1] = 0x7fffffff;
    }
    if (lane == round) sel = hi;
  }
  // exact distances of the k selected candidates (lane t < k owns candidate t), then order by (distance, index)
  const double* y = Y + int64_t(row0 + r) * ldY;
  double dist = INFINITY;
  for (int t = 0; t < k; ++t) {
    const int j = __shfl_sync(0xffffffffu, sel, t);
    double acc = 0.0;
    if (j != 0x7fffffff) {
      const double* x = X + int64_t(j) * ldX;
      for (int c = lane; c < d; c += 32) {
        const double df = y[c] - x[c];
        acc = fma(df, df, acc);
      }
    }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
    if (lane == t) dist = j != 0x7fffffff ? sqrt(acc) : INFINITY;
  }
  // rank of candidate `lane` among the k (k <= 16: all-pairs comparison through shuffles)
  int rank = 0;
  for (int t = 0; t < k; ++t) {
    const double od = __shfl_sync(0xffffffffu, dist, t);
    const int oj = __shfl_sync(0xffffffffu, sel, t);
    if (lane < k && t != lane && (od < dist || (od == dist && oj < sel))) ++rank;
  }
  if (lane < k) {
    idx_out[int64_t(row0 + r) * k + rank] = sel == 0x7fffffff ? 0 : sel;
    dist_out[int64_t(row0 + r) * k + rank] = dist;
  }
}

__global__ void __launch_bounds__(256) row_sqnorm_kernel(const double* __restrict__ X, int64_t ld, int n, int d, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const double* x = X + int64_t(r) * ld;
  double s = 0.0;
  for (int c = lane; c < d; c += 32) s = fma(x[c], x[c], s);
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
  if (lane == 0) out[r] = s;
}
}  // namespace
}  // namespace dm

using namespace dm;

extern "C" {

size_t dm_knn_workspace_bytes(int nq, int ndb, int d, int k) {
  (void)d, (void)k;
  if (nq < 0 || ndb < 0) return 0;
  Carver c(nullptr);
  c.take<double>(size_t(ndb));
  c.take<double>(size_t(nq < kChunkRows ? nq : kChunkRows) * size_t(ndb));
  return c.bytes();
}

int dm_knn_f64(const double* Y, int64_t ldY, int nq, const double* X, int64_t ldX, int ndb, int d, int k, int64_t* idx,
               double* dist, void* workspace, size_t workspace_bytes, dm_stream_t stream) {
  if (nq < 0 || ndb < 0 || d <= 0 || k < 1) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (k > kMaxK) DM_FAIL(DM_ERR_UNSUPPORTED, "k = %d: at most %d neighbours", k, kMaxK);
  if (k > ndb) DM_FAIL(DM_ERR_BADARG, "Expected n_neighbors <= n_samples_fit, but n_neighbors = %d, n_samples_fit = %d", k, ndb);
  if (nq == 0) return DM_OK;
  if (!Y || !X || !idx || !dist) DM_FAIL(DM_ERR_BADARG, "null argument");
  if (ldY < d || ldX < d) DM_FAIL(DM_ERR_BADARG, "leading dimension smaller than d");
  if (!workspace || dm_knn_workspace_bytes(nq, ndb, d, k) > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver c(workspace);
  double* xsq = c.take<double>(size_t(ndb));
  double* S = c.take<double>(size_t(nq < kChunkRows ? nq : kChunkRows) * size_t(ndb));
  row_sqnorm_kernel<<<unsigned((ndb + 7) / 8), 256, 0, st>>>(X, ldX, ndb, d, xsq);
  DM_LAUNCH_OK("row_sqnorm_kernel");
  for (int r0 = 0; r0 < nq; r0 += kChunkRows) {
    const int rows = nq - r0 < kChunkRows ? nq - r0 : kChunkRows;
    GemmProblem G;
    G.A.d = Y + int64_t(r0) * ldY, G.A.ld = ldY, G.A.rows = rows, G.A.trans = 0;
    G.B.d = X, G.B.ld = ldX, G.B.rows = ndb, G.B.trans = 0;
    G.M = rows, G.N = ndb, G.K = d, G.maxM = rows, G.maxN = ndb, G.maxK = d, G.n_batch = 1;
    G.C = S, G.ldc = ndb;
    int rc;
    if ((rc = gemm64_launch(G, st))) return rc;
    knn_topk_kernel<<<unsigned((rows + 7) / 8), 256, 0, st>>>(S, ndb, Y, ldY, X, ldX, xsq, r0, rows, ndb, d, k, idx, dist);
    DM_LAUNCH_OK("knn_topk_kernel");
  }
  return DM_OK;
}

}  // extern "C"
