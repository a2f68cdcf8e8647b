// extern "C" surface of libdm_b200.so (declared in include/dm_b200.h).
#include "dm_internal.cuh"

namespace dm {
const char* last_error();

namespace {
__global__ void __launch_bounds__(256)
    match_dist_kernel(const float* __restrict__ Y, int64_t ldY, const float* __restrict__ X, int64_t ldX,
                      const void* __restrict__ idx, int64_t n, int d, double* __restrict__ dist, int i64) {
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int64_t j = load_index(idx, row, i64 != 0);
  const float* y = Y + row * ldY;
  const float* x = X + j * ldX;
  double s = 0.0;
  for (int k = lane; k < d; k += 32) {
    const double df = double(y[k]) - double(x[k]);
    s = fma(df, df, s);
  }
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
  if (lane == 0) dist[row] = sqrt(s);
}

__global__ void set_single_pair_offsets(int64_t* q_off, int64_t nq, int64_t* db_off, int64_t ndb) {
  q_off[0] = 0, q_off[1] = nq, db_off[0] = 0, db_off[1] = ndb;
}

__global__ void __launch_bounds__(256)
    cvt_f64_f32_kernel(const double* __restrict__ src, int64_t lds, int64_t rows, int d, float* __restrict__ dst,
                       int ldd) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * ldd) return;
  const int64_t r = i / ldd;
  const int k = int(i % ldd);
  dst[i] = k < d ? float(src[r * lds + k]) : 0.f;
}
}  // namespace

int cvt_f64_f32(const double* src, int64_t lds, int64_t rows, int d, float* dst, int ldd, cudaStream_t st) {
  if (rows <= 0) return DM_OK;
  const int64_t n = rows * ldd;
  cvt_f64_f32_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(src, lds, rows, d, dst, ldd);
  DM_LAUNCH_OK("cvt_f64_f32_kernel");
  return DM_OK;
}
}  // namespace dm

using namespace dm;

extern "C" {

const char* dm_last_error(void) { return dm::last_error(); }
int dm_version(void) { return DM_VERSION; }
const char* dm_build_info(void) {
  return "libdm_b200 " __DATE__ " sm_100a engines=tcgen05(split-bf16,default),ffma";
}

size_t dm_nn_workspace_bytes(int n_pairs, int64_t total_q, int64_t total_db, int max_q, int max_db, int d,
                             int n_row_epi, int n_col_epi, int flags) {
  if (n_pairs < 0 || total_q < 0 || total_db < 0 || max_q < 0 || max_db < 0 || n_row_epi < 0 || n_col_epi < 0 ||
      n_row_epi > kMaxEpi || n_col_epi > kMaxEpi)
    return 0;
  return nn_workspace_bytes(n_pairs, total_q, total_db, max_q, max_db, d, n_row_epi, n_col_epi, flags);
}

int dm_nn_argmax_f32(const float* Y, int64_t ldY, const int64_t* q_off, int64_t total_q, int max_q, const float* X,
                     int64_t ldX, const int64_t* db_off, int64_t total_db, int max_db, int n_pairs, int d,
                     const dm_nn_epi* row_epi_h, int n_row_epi, const dm_nn_epi* col_epi_h, int n_col_epi, int flags,
                     void* workspace, size_t workspace_bytes, dm_stream_t stream) {
  if (n_row_epi < 0 || n_row_epi > kMaxEpi || n_col_epi < 0 || n_col_epi > kMaxEpi)
    DM_FAIL(DM_ERR_BADARG, "at most %d row and %d column epilogues", kMaxEpi, kMaxEpi);
  if ((n_row_epi && !row_epi_h) || (n_col_epi && !col_epi_h)) DM_FAIL(DM_ERR_BADARG, "epilogue array is null");
  NNRequest R{};
  R.Y = Y, R.ldY = ldY, R.X = X, R.ldX = ldX;
  R.Y64 = nullptr, R.X64 = nullptr;
  R.q_off = q_off, R.db_off = db_off, R.total_q = total_q, R.total_db = total_db;
  R.max_q = max_q, R.max_db = max_db, R.n_pairs = n_pairs, R.d = d;
  R.n_row = n_row_epi, R.n_col = n_col_epi;
  for (int e = 0; e < n_row_epi; ++e) R.row[e] = row_epi_h[e];
  for (int e = 0; e < n_col_epi; ++e) R.col[e] = col_epi_h[e];
  R.flags = flags;
  return nn_run(R, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t dm_nn_f64_workspace_bytes(int n_pairs, int64_t total_q, int64_t total_db, int max_q, int max_db, int d,
                                 int n_row_epi, int n_col_epi, int flags) {
  const size_t inner = dm_nn_workspace_bytes(n_pairs, total_q, total_db, max_q, max_db, d, n_row_epi, n_col_epi, flags);
  const size_t ldd = size_t((d + 3) / 4 * 4);
  Carver c(nullptr);
  c.take<char>(inner);  // first, so that dm_nn_read_stats finds the counters at the start of the workspace
  if (!nn_use_tc(flags)) {  // the CUDA-core engine runs on fp32 copies; the tensor-core engine splits the originals
    c.take<float>(size_t(total_q) * ldd);
    c.take<float>(size_t(total_db) * ldd);
  }
  return c.bytes();
}

int dm_nn_argmax_f64(const double* Y, int64_t ldY, const int64_t* q_off, int64_t total_q, int max_q, const double* X,
                     int64_t ldX, const int64_t* db_off, int64_t total_db, int max_db, int n_pairs, int d,
                     const dm_nn_epi* row_epi_h, int n_row_epi, const dm_nn_epi* col_epi_h, int n_col_epi, int flags,
                     void* workspace, size_t workspace_bytes, dm_stream_t stream) {
  if (n_row_epi < 0 || n_row_epi > kMaxEpi || n_col_epi < 0 || n_col_epi > kMaxEpi)
    DM_FAIL(DM_ERR_BADARG, "at most %d row and %d column epilogues", kMaxEpi, kMaxEpi);
  if ((n_row_epi && !row_epi_h) || (n_col_epi && !col_epi_h)) DM_FAIL(DM_ERR_BADARG, "epilogue array is null");
  if (total_q < 0 || total_db < 0 || d <= 0 || ldY < d || ldX < d) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n_pairs == 0 || (total_q == 0 && total_db == 0)) return DM_OK;
  if (!workspace) DM_FAIL(DM_ERR_WORKSPACE, "workspace is null");
  if (reinterpret_cast<uintptr_t>(workspace) % 256) DM_FAIL(DM_ERR_ALIGN, "workspace must be 256-byte aligned");
  const size_t need = dm_nn_f64_workspace_bytes(n_pairs, total_q, total_db, max_q, max_db, d, n_row_epi, n_col_epi, flags);
  if (need > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ldd = (d + 3) / 4 * 4;
  Carver c(workspace);
  const size_t inner = dm_nn_workspace_bytes(n_pairs, total_q, total_db, max_q, max_db, d, n_row_epi, n_col_epi, flags);
  char* inner_ws = c.take<char>(inner);
  float *Yf = nullptr, *Xf = nullptr;
  int rc;
  if (!nn_use_tc(flags)) {
    Yf = c.take<float>(size_t(total_q) * ldd);
    Xf = c.take<float>(size_t(total_db) * ldd);
    if ((rc = cvt_f64_f32(Y, ldY, total_q, d, Yf, ldd, st))) return rc;
    if ((rc = cvt_f64_f32(X, ldX, total_db, d, Xf, ldd, st))) return rc;
  }
  NNRequest R{};
  R.Y = Yf, R.ldY = ldd, R.X = Xf, R.ldX = ldd;
  R.Y64 = Y, R.ldY64 = ldY, R.X64 = X, R.ldX64 = ldX;
  R.q_off = q_off, R.db_off = db_off, R.total_q = total_q, R.total_db = total_db;
  R.max_q = max_q, R.max_db = max_db, R.n_pairs = n_pairs, R.d = d;
  R.n_row = n_row_epi, R.n_col = n_col_epi;
  for (int e = 0; e < n_row_epi; ++e) R.row[e] = row_epi_h[e];
  for (int e = 0; e < n_col_epi; ++e) R.col[e] = col_epi_h[e];
  R.flags = flags;
  return nn_run(R, inner_ws, inner, st);
}

int dm_nn_read_stats(const void* workspace, int64_t* out_h, dm_stream_t stream) {
  if (!workspace || !out_h) DM_FAIL(DM_ERR_BADARG, "null argument");
  unsigned int c[4];
  DM_CUDA_OK(cudaMemcpyAsync(c, workspace, sizeof(c), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  DM_CUDA_OK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  out_h[0] = c[1];
  out_h[1] = c[2];
  out_h[2] = int64_t(c[0]) + c[3];
  out_h[3] = c[3];  // results that needed the full float64 scan (the others were decided between two candidates)
  return DM_OK;
}

size_t dm_nn_debug_workspace_bytes(int nq, int ndb, int d, int flags) {
  if (nq < 0 || ndb < 0 || d <= 0) return 0;
  Carver c(nullptr);
  if (nn_use_tc(flags)) {
    const size_t kp = size_t(nn_tc_kp(d));
    c.take<int64_t>(4);
    c.take<float>(size_t(nq));
    c.take<float>(size_t(ndb));
    c.take<uint16_t>(size_t(nq) * kp);
    c.take<uint16_t>(size_t(nq) * kp);
    c.take<uint16_t>(size_t(ndb) * kp);
    c.take<uint16_t>(size_t(ndb) * kp);
  }
  return c.bytes();
}

int dm_nn_debug_scores_f32(const float* Y, int64_t ldY, int nq, const float* X, int64_t ldX, int ndb, int d,
                           float* S_out, int64_t ldS, int flags, void* workspace, size_t workspace_bytes,
                           dm_stream_t stream) {
  if (!Y || !X || !S_out || nq < 0 || ndb < 0 || d <= 0 || ldY < d || ldX < d || ldS < ndb)
    DM_FAIL(DM_ERR_BADARG, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!nn_use_tc(flags)) return nn_ffma_debug_scores(Y, ldY, nq, X, ldX, ndb, d, S_out, ldS, st);
  if (nq == 0 || ndb == 0) return DM_OK;
  const size_t need = dm_nn_debug_workspace_bytes(nq, ndb, d, flags);
  if (!workspace || need > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu", need);
  if (reinterpret_cast<uintptr_t>(workspace) % 256) DM_FAIL(DM_ERR_ALIGN, "workspace must be 256-byte aligned");
  const int kp = nn_tc_kp(d);
  Carver c(workspace);
  int64_t* off = c.take<int64_t>(4);
  float* nqv = c.take<float>(size_t(nq));
  float* ndv = c.take<float>(size_t(ndb));
  uint16_t* yh = c.take<uint16_t>(size_t(nq) * kp);
  uint16_t* yl = c.take<uint16_t>(size_t(nq) * kp);
  uint16_t* xh = c.take<uint16_t>(size_t(ndb) * kp);
  uint16_t* xl = c.take<uint16_t>(size_t(ndb) * kp);
  set_single_pair_offsets<<<1, 1, 0, st>>>(off, nq, off + 2, ndb);
  DM_LAUNCH_OK("set_single_pair_offsets");
  int rc;
  if ((rc = nn_prep_side(Y, 0, ldY, off, 1, nq, d, nqv, nullptr, 0, yh, yl, nullptr, kp, st))) return rc;
  if ((rc = nn_prep_side(X, 0, ldX, off + 2, 1, ndb, d, ndv, nullptr, 0, xh, xl, nullptr, kp, st))) return rc;
  NNProblem P{};
  P.q_off = off, P.db_off = off + 2, P.total_q = nq, P.total_db = ndb, P.max_q = nq, P.max_db = ndb;
  P.q_in = P.q_off, P.db_in = P.db_off, P.rows_q = nq, P.rows_db = ndb;
  P.n_pairs = 1, P.d = d, P.kp = kp, P.rt_rows = kFfmaRowTile, P.max_rt = (nq + kFfmaRowTile - 1) / kFfmaRowTile;
  return nn_tc_launch(P, yh, yl, xh, xl, S_out, ldS, st);
}

int dm_match_dist_f32(const float* Y, int64_t ldY, const float* X, int64_t ldX, const void* idx, int64_t n, int d,
                      double* dist, int flags, dm_stream_t stream) {
  if (n < 0 || d <= 0) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n == 0) return DM_OK;
  if (!Y || !X || !idx || !dist) DM_FAIL(DM_ERR_BADARG, "null argument");
  match_dist_kernel<<<unsigned((n + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      Y, ldY, X, ldX, idx, n, d, dist, (flags & DM_I64_OUT) ? 1 : 0);
  DM_LAUNCH_OK("match_dist_kernel");
  return DM_OK;
}

}  // extern "C"
