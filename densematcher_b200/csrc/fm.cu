// Functional-map stages: spectral projection, FM -> p2p (four index outputs
// from one score pass), p2p -> FM and the ZoomOut ladder.  Everything that enters C is float64; the
// N x N x k score pass runs through the fused NN engine with the float64 near-tie re-evaluation.
#include "dm_internal.cuh"
#include "gemm64.cuh"
#include "linalg64.cuh"
#include "tc_ptx.cuh"

namespace dm {
namespace {

// bias[j] = -1/2 |row_j|^2 over the first d columns (float64), one warp per row
__global__ void __launch_bounds__(256)
    neg_half_sqnorm_kernel(const double* __restrict__ M, int64_t ld, int64_t rows, int d, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const double* r = M + row * ld;
  double s = 0.0;
  for (int k = lane; k < d; k += 32) s = fma(r[k], r[k], s);
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
  if (lane == 0) out[row] = -0.5 * s;
}

int neg_half_sqnorm(const double* M, int64_t ld, int64_t rows, int d, double* out, cudaStream_t st) {
  if (rows <= 0) return DM_OK;
  neg_half_sqnorm_kernel<<<unsigned((rows + 7) / 8), 256, 0, st>>>(M, ld, rows, d, out);
  DM_LAUNCH_OK("neg_half_sqnorm_kernel");
  return DM_OK;
}

int pad4(int x) { return (x + 3) / 4 * 4; }

// ---- p2p -> FM (internal, with split-K workspace)
static const int kP2PChunk = [] { const char* e = getenv("DM_P2P_CHUNK"); return e ? atoi(e) : 256; }();
size_t p2p_to_fm_ws(int n_pairs, int max_n2, int k1, int k2) {
  const int ks = (max_n2 + kP2PChunk - 1) / kP2PChunk;
  Carver c(nullptr);
  if (ks > 1) c.take<double>(size_t(ks) * n_pairs * k1 * k2);
  return c.bytes();
}
size_t zoomout_pf_ws(int n_pairs, int64_t total_n2, int max_n2, int k1m, int k2m, int flags) {
  const size_t a = p2p_to_fm_ws(n_pairs, max_n2, k1m, k2m);
  const size_t b = ((flags & DM_FAST_FM) && proj_tc_supported(k2m, k1m))
                       ? proj_tc_workspace_bytes(n_pairs, total_n2, max_n2, k2m, k1m) : 0;
  return a > b ? a : b;
}

int p2p_to_fm_run(const void* p2p, int p2p_i64, const double* Phi1, int64_t ld1, const int64_t* off1,
                  const double* Phi2, int64_t ld2, const int64_t* off2, int max_n2, const double* area2, int n_pairs,
                  int k1, int k2, double* C, void* ws, cudaStream_t st) {
  if (n_pairs <= 0 || k1 <= 0 || k2 <= 0) return DM_OK;
  const int ks = (max_n2 + kP2PChunk - 1) / kP2PChunk;
  GemmProblem G;
  G.A.d = Phi2, G.A.ld = ld2, G.A.off = off2, G.A.trans = 1, G.A.kscale = area2;
  G.B.d = Phi1, G.B.ld = ld1, G.B.off = off1, G.B.trans = 1, G.B.gather = p2p, G.B.gather_i64 = p2p_i64,
  G.B.gather_off = off2;
  G.M = k2, G.N = k1, G.maxM = k2, G.maxN = k1, G.maxK = max_n2, G.n_batch = n_pairs;
  G.ldc = k1, G.c_batch_stride = int64_t(k1) * k2;
  int rc;
  if (ks <= 1) {
    G.C = C;
    return gemm64_launch(G, st);
  }
  Carver c(ws);
  double* part = c.take<double>(size_t(ks) * n_pairs * k1 * k2);
  G.C = part, G.ksplit = ks, G.kchunk = kP2PChunk, G.split_stride = int64_t(n_pairs) * k1 * k2;
  if ((rc = gemm64_launch(G, st))) return rc;
  return sum_partials_launch(part, ks, G.split_stride, G.split_stride, C, st);
}

// ---- p2p_21 of a functional map, upstream pyFM semantics (the ZoomOut / ICP inner conversion):
//      knn(tree = Phi1[:, :k1] C^T, query = Phi2[:, :k2])
struct P2P21Scratch {
  double* emb1;  // [total_n1, lde]
  float* Xf;     // [total_n1, ldf]
  void* nn_ws;
  size_t nn_ws_bytes;
  int lde, ldf;
  void* fact = nullptr;  // scratch of the factored (tensor-core embedding) path, k <= 256
  int fact_k1m = 0, fact_k2m = 0;  // widths it was sized for
  int x_kp = -1;         // padded width for which the split of Phi1 in `fact` is valid
};
int p2p21_run(const double* C, int k1, int k2, const double* Phi1, int64_t ld1, const int64_t* off1, int64_t total_n1,
              int max_n1, const double* Phi2, int64_t ld2, const float* Phi2f, int ldPhi2f, const int64_t* off2,
              int64_t total_n2, int max_n2, int n_pairs, void* p2p_out, int flags, P2P21Scratch& S,
              cudaStream_t st, int* y_kp_state = nullptr) {
  int rc;
  // k <= 128: the database side Phi1 C^T is embedded on the tensor cores, float64 rows on demand (embed_tc.cu)
  if (S.fact && k1 <= S.fact_k1m && k2 <= S.fact_k2m && p2p21_factored_applicable(k1, k2, flags))
    return p2p21_factored_run(C, k1, k2, Phi1, ld1, off1, total_n1, max_n1, Phi2, ld2, off2, total_n2, max_n2, n_pairs,
                              p2p_out, flags, S.fact, S.fact_k1m, S.fact_k2m, S.emb1, S.lde, S.nn_ws, S.nn_ws_bytes, st, &S.x_kp,
                              y_kp_state);
  GemmProblem G;
  G.A.d = Phi1, G.A.ld = ld1, G.A.off = off1, G.A.trans = 0;
  G.B.d = C, G.B.ld = k1, G.B.batch_stride = int64_t(k1) * k2, G.B.rows = k2, G.B.trans = 0;
  G.N = k2, G.K = k1, G.maxM = max_n1, G.maxN = k2, G.maxK = k1, G.n_batch = n_pairs;
  G.C = S.emb1, G.ldc = S.lde, G.c_off = off1;
  if ((rc = gemm64_launch(G, st))) return rc;
  const int dfast = pad4(k2);
  const bool tc = nn_use_tc(flags);  // the tensor-core engine splits the float64 operands itself
  if (!tc && (rc = cvt_f64_f32(S.emb1, S.lde, total_n1, k2, S.Xf, S.ldf, st))) return rc;
  // Xf columns k2..dfast must be zero: cvt writes ldf columns per row, zero beyond k2
  NNRequest R{};
  R.Y = tc ? nullptr : Phi2f, R.ldY = ldPhi2f, R.X = tc ? nullptr : S.Xf, R.ldX = S.ldf;
  R.Y64 = Phi2, R.ldY64 = ld2, R.X64 = S.emb1, R.ldX64 = S.lde;
  R.q_off = off2, R.db_off = off1, R.total_q = total_n2, R.total_db = total_n1;
  R.max_q = max_n2, R.max_db = max_n1, R.n_pairs = n_pairs, R.d = k2, R.d_fast = dfast;
  R.n_row = 1, R.n_col = 0;
  R.row[0] = dm_nn_epi{DM_SCALE_NONE, DM_BIAS_NEG_HALF_SQNORM, nullptr, nullptr, p2p_out};
  R.flags = flags;
  if (tc && y_kp_state) {
    // the query matrix Phi2 is the same at every rung: split it once per padded width (64 / 128 / 192 / 256), with all
    // the columns of that width (those beyond k2 meet the zero padding of the database side)
    const int kp = nn_tc_kp(k2);
    R.y_prep_d = int(ld2 < kp ? ld2 : kp);
    R.skip_prep_y = (*y_kp_state == kp);
    *y_kp_state = kp;
  }
  return nn_run(R, S.nn_ws, S.nn_ws_bytes, st);
}

}  // namespace
}  // namespace dm

using namespace dm;

extern "C" {

// ------------------------------------------------------------------ projection
static size_t project_f64_ws(int n_meshes, int max_n, int k, int d) {
  const int ks = (max_n + 255) / 256;
  Carver c(nullptr);
  if (ks > 1) c.take<double>(size_t(ks) * n_meshes * k * d);
  return c.bytes();
}

size_t dm_project_workspace_bytes(int n_meshes, int64_t total_n, int max_n, int k, int d) {
  if (n_meshes < 0 || total_n < 0 || max_n < 0 || k <= 0 || d <= 0) return 0;
  const size_t a = project_f64_ws(n_meshes, max_n, k, d);
  const size_t b = proj_tc_supported(k, d) ? proj_tc_workspace_bytes(n_meshes, total_n, max_n, k, d) : 0;
  return a > b ? a : b;
}

int dm_project_ex(const double* Phi, int64_t ldPhi, const double* area, const float* F, int64_t ldF,
                  const int64_t* row_off, int64_t total_n, int max_n, int n_meshes, int k, int d, double* out,
                  int flags, void* workspace, size_t workspace_bytes, dm_stream_t stream) {
  if (n_meshes < 0 || k <= 0 || d <= 0 || total_n < 0 || max_n < 0) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n_meshes == 0) return DM_OK;
  if (!Phi || !area || !F || !row_off || !out) DM_FAIL(DM_ERR_BADARG, "null argument");
  if (ldPhi < k || ldF < d) DM_FAIL(DM_ERR_BADARG, "leading dimension too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!(flags & DM_F64_GEMM) && proj_tc_supported(k, d)) {
    const size_t need = proj_tc_workspace_bytes(n_meshes, total_n, max_n, k, d);
    if (!workspace || need > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu", need);
    if (reinterpret_cast<uintptr_t>(workspace) % 256) DM_FAIL(DM_ERR_ALIGN, "workspace must be 256-byte aligned");
    return proj_tc_run(Phi, ldPhi, area, F, nullptr, ldF, nullptr, nullptr, 0, nullptr, row_off, total_n, max_n,
                       n_meshes, k, d, out, workspace, workspace_bytes, st);
  }
  const size_t need = project_f64_ws(n_meshes, max_n, k, d);
  if (need > workspace_bytes || (need && !workspace)) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu", need);
  const int ks = (max_n + 255) / 256;
  GemmProblem G;
  G.A.d = Phi, G.A.ld = ldPhi, G.A.off = row_off, G.A.trans = 1, G.A.kscale = area;
  G.B.f = F, G.B.ld = ldF, G.B.off = row_off, G.B.trans = 1;
  G.M = k, G.N = d, G.maxM = k, G.maxN = d, G.maxK = max_n, G.n_batch = n_meshes;
  G.ldc = d, G.c_batch_stride = int64_t(k) * d;
  if (ks <= 1) {
    G.C = out;
    return gemm64_launch(G, st);
  }
  Carver c(workspace);
  double* part = c.take<double>(size_t(ks) * n_meshes * k * d);
  G.C = part, G.ksplit = ks, G.kchunk = 256, G.split_stride = int64_t(n_meshes) * k * d;
  int rc;
  if ((rc = gemm64_launch(G, st))) return rc;
  return sum_partials_launch(part, ks, G.split_stride, G.split_stride, out, st);
}

int dm_project(const double* Phi, int64_t ldPhi, const double* area, const float* F, int64_t ldF,
               const int64_t* row_off, int64_t total_n, int max_n, int n_meshes, int k, int d, double* out,
               void* workspace, size_t workspace_bytes, dm_stream_t stream) {
  return dm_project_ex(Phi, ldPhi, area, F, ldF, row_off, total_n, max_n, n_meshes, k, d, out, 0, workspace,
                       workspace_bytes, stream);
}

// ------------------------------------------------------------------ FM -> p2p
size_t dm_fm_to_p2p_workspace_bytes(int n_pairs, int64_t total_n1, int64_t total_n2, int max_n1, int max_n2, int k1,
                                    int k2, int flags) {
  Carver c(nullptr);
  c.take<double>(size_t(total_n2) * k1);        // emb2 = Phi2 C
  c.take<double>(size_t(total_n1) * k2);        // emb1 = Phi1 C^T
  c.take<double>(size_t(total_n1));             // -1/2 |emb1|^2
  c.take<float>(size_t(total_n2) * pad4(k1));   // fp32 emb2
  c.take<float>(size_t(total_n1) * pad4(k1));   // fp32 Phi1
  c.take<char>(nn_workspace_bytes(n_pairs, total_n2, total_n1, max_n2, max_n1, k1, 2, 2, flags));
  const size_t fact = f2p_factored_applicable(k1, k2, flags)
                          ? f2p_factored_workspace_bytes(n_pairs, total_n1, total_n2, max_n1, max_n2, k1, k2, flags) : 0;
  return c.bytes() > fact ? c.bytes() : fact;
}

int dm_fm_to_p2p(const double* C, int k1, int k2, const double* Phi1, int64_t ld1, const int64_t* off1,
                 int64_t total_n1, int max_n1, const double* Phi2, int64_t ld2, const int64_t* off2, int64_t total_n2,
                 int max_n2, const double* area1, int n_pairs, void* p2p_21, void* p2p_12, void* dense_21,
                 void* dense_12, int flags, void* workspace, size_t workspace_bytes, dm_stream_t stream) {
  if (n_pairs < 0 || k1 <= 0 || k2 <= 0 || total_n1 < 0 || total_n2 < 0) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n_pairs == 0) return DM_OK;
  if (!C || !Phi1 || !Phi2 || !off1 || !off2) DM_FAIL(DM_ERR_BADARG, "null argument");
  if (ld1 < k1 || ld2 < k2) DM_FAIL(DM_ERR_BADARG, "eigenbasis has fewer columns than the functional map");
  if (dense_21 && !area1) DM_FAIL(DM_ERR_BADARG, "dense_21 needs area1");
  if (!p2p_21 && !p2p_12 && !dense_21 && !dense_12) DM_FAIL(DM_ERR_BADARG, "no output requested");
  const size_t need = dm_fm_to_p2p_workspace_bytes(n_pairs, total_n1, total_n2, max_n1, max_n2, k1, k2, flags);
  if (!workspace || need > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu", need);
  if (reinterpret_cast<uintptr_t>(workspace) % 256) DM_FAIL(DM_ERR_ALIGN, "workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // default: embeddings on the tensor cores, float64 on demand for the re-evaluated results only (embed_tc.cu)
  if (f2p_factored_applicable(k1, k2, flags))
    return f2p_factored_run(C, k1, k2, Phi1, ld1, off1, total_n1, max_n1, Phi2, ld2, off2, total_n2, max_n2, area1, n_pairs,
                            p2p_21, p2p_12, dense_21, dense_12, flags, workspace, st);
  Carver c(workspace);
  double* emb2 = c.take<double>(size_t(total_n2) * k1);
  double* emb1 = c.take<double>(size_t(total_n1) * k2);
  double* bias1 = c.take<double>(size_t(total_n1));
  const int ldf = pad4(k1);
  float* Yf = c.take<float>(size_t(total_n2) * ldf);
  float* Xf = c.take<float>(size_t(total_n1) * ldf);
  const size_t nn_bytes = nn_workspace_bytes(n_pairs, total_n2, total_n1, max_n2, max_n1, k1, 2, 2, flags);
  char* nn_ws = c.take<char>(nn_bytes);
  int rc;
  const bool skip_prep = (flags & DM_SKIP_PREP) != 0;  // profiling: the embeddings of a previous identical call are reused
  // emb2 = Phi2[:, :k2] C   (convert.py:134)
  if (!skip_prep) {
    GemmProblem G;
    G.A.d = Phi2, G.A.ld = ld2, G.A.off = off2, G.A.trans = 0;
    G.B.d = C, G.B.ld = k1, G.B.batch_stride = int64_t(k1) * k2, G.B.rows = k2, G.B.trans = 1;
    G.N = k1, G.K = k2, G.maxM = max_n2, G.maxN = k1, G.maxK = k2, G.n_batch = n_pairs;
    G.C = emb2, G.ldc = k1, G.c_off = off2;
    if ((rc = gemm64_launch(G, st))) return rc;
  }
  if (p2p_21 && !skip_prep) {  // |emb1_j|^2 with emb1 = Phi1[:, :k1] C^T   (convert.py:138)
    GemmProblem G;
    G.A.d = Phi1, G.A.ld = ld1, G.A.off = off1, G.A.trans = 0;
    G.B.d = C, G.B.ld = k1, G.B.batch_stride = int64_t(k1) * k2, G.B.rows = k2, G.B.trans = 0;
    G.N = k2, G.K = k1, G.maxM = max_n1, G.maxN = k2, G.maxK = k1, G.n_batch = n_pairs;
    G.C = emb1, G.ldc = k2, G.c_off = off1;
    if ((rc = gemm64_launch(G, st))) return rc;
    if ((rc = neg_half_sqnorm(emb1, k2, total_n1, k2, bias1, st))) return rc;
  }
  const bool tc = nn_use_tc(flags);  // the tensor-core engine splits the float64 operands itself
  if (!tc) {
    if ((rc = cvt_f64_f32(emb2, k1, total_n2, k1, Yf, ldf, st))) return rc;
    if ((rc = cvt_f64_f32(Phi1, ld1, total_n1, k1, Xf, ldf, st))) return rc;
  }
  NNRequest R{};
  R.Y = tc ? nullptr : Yf, R.ldY = ldf, R.X = tc ? nullptr : Xf, R.ldX = ldf;
  R.Y64 = emb2, R.ldY64 = k1, R.X64 = Phi1, R.ldX64 = ld1;
  R.q_off = off2, R.db_off = off1, R.total_q = total_n2, R.total_db = total_n1;
  R.max_q = max_n2, R.max_db = max_n1, R.n_pairs = n_pairs, R.d = k1, R.d_fast = ldf;
  R.n_row = 0, R.n_col = 0;
  if (p2p_21) R.row[R.n_row++] = dm_nn_epi{DM_SCALE_NONE, DM_BIAS_ARRAY, nullptr, bias1, p2p_21};
  if (dense_21) R.row[R.n_row++] = dm_nn_epi{DM_SCALE_ARRAY, DM_BIAS_NONE, area1, nullptr, dense_21};
  if (p2p_12) R.col[R.n_col++] = dm_nn_epi{DM_SCALE_NONE, DM_BIAS_NEG_HALF_SQNORM, nullptr, nullptr, p2p_12};
  if (dense_12) R.col[R.n_col++] = dm_nn_epi{DM_SCALE_NONE, DM_BIAS_NONE, nullptr, nullptr, dense_12};
  R.flags = flags;
  return nn_run(R, nn_ws, nn_bytes, st);
}

size_t dm_mapped_indicator_workspace_bytes(int n1, int k2) {
  Carver c(nullptr);
  c.take<double>(size_t(n1 > 0 ? n1 : 0) * (k2 > 0 ? k2 : 0));
  return c.bytes();
}

int dm_mapped_indicator(const double* C, int k1, int k2, const double* Phi1, int64_t ld1, int n1, const double* Phi2,
                        int64_t ld2, int n2, const double* area1, double* MI, int64_t ldMI, void* workspace,
                        size_t workspace_bytes, dm_stream_t stream) {
  if (k1 <= 0 || k2 <= 0 || n1 < 0 || n2 < 0 || ldMI < n1) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n1 == 0 || n2 == 0) return DM_OK;
  if (!C || !Phi1 || !Phi2 || !area1 || !MI) DM_FAIL(DM_ERR_BADARG, "null argument");
  if (ld1 < k1 || ld2 < k2) DM_FAIL(DM_ERR_BADARG, "eigenbasis has fewer columns than the functional map");
  if (!workspace || dm_mapped_indicator_workspace_bytes(n1, k2) > workspace_bytes)
    DM_FAIL(DM_ERR_WORKSPACE, "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver c(workspace);
  double* W = c.take<double>(size_t(n1) * k2);  // W = Phi1 C^T  [n1, k2];  MI = Phi2 W^T diag(area1)
  int rc;
  GemmProblem G;
  G.A.d = Phi1, G.A.ld = ld1, G.A.rows = n1, G.A.trans = 0;
  G.B.d = C, G.B.ld = k1, G.B.rows = k2, G.B.trans = 0;
  G.M = n1, G.N = k2, G.K = k1, G.maxM = n1, G.maxN = k2, G.maxK = k1, G.n_batch = 1;
  G.C = W, G.ldc = k2;
  if ((rc = gemm64_launch(G, st))) return rc;
  GemmProblem H;
  H.A.d = Phi2, H.A.ld = ld2, H.A.rows = n2, H.A.trans = 0;
  H.B.d = W, H.B.ld = k2, H.B.rows = n1, H.B.trans = 0;
  H.M = n2, H.N = n1, H.K = k2, H.maxM = n2, H.maxN = n1, H.maxK = k2, H.n_batch = 1;
  H.C = MI, H.ldc = ldMI, H.c_colscale = area1;
  return gemm64_launch(H, st);
}

// the same for a ragged batch: MI rows of pair p are rows off2[p] .. off2[p + 1] of MI [total_n2, ldMI >= max_n1], columns
// 0 .. n1_p - 1 (two ragged GEMMs for the whole batch; per pair the arithmetic of dm_mapped_indicator)
size_t dm_mapped_indicators_workspace_bytes(int64_t total_n1, int k2) {
  Carver c(nullptr);
  c.take<double>(size_t(total_n1 > 0 ? total_n1 : 0) * (k2 > 0 ? k2 : 0));
  return c.bytes();
}

int dm_mapped_indicators(const double* C, int k1, int k2, const double* Phi1, int64_t ld1, const int64_t* off1,
                         int64_t total_n1, int max_n1, const double* Phi2, int64_t ld2, const int64_t* off2, int max_n2,
                         const double* area1, int n_pairs, double* MI, int64_t ldMI, void* workspace, size_t workspace_bytes,
                         dm_stream_t stream) {
  if (k1 <= 0 || k2 <= 0 || n_pairs < 0 || total_n1 < 0 || max_n1 < 0 || max_n2 < 0 || ldMI < max_n1)
    DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n_pairs == 0 || total_n1 == 0 || max_n2 == 0) return DM_OK;
  if (!C || !Phi1 || !Phi2 || !off1 || !off2 || !area1 || !MI) DM_FAIL(DM_ERR_BADARG, "null argument");
  if (ld1 < k1 || ld2 < k2) DM_FAIL(DM_ERR_BADARG, "eigenbasis has fewer columns than the functional map");
  if (!workspace || dm_mapped_indicators_workspace_bytes(total_n1, k2) > workspace_bytes)
    DM_FAIL(DM_ERR_WORKSPACE, "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver c(workspace);
  double* W = c.take<double>(size_t(total_n1) * k2);  // W = Phi1 C^T, rows packed like Phi1
  int rc;
  GemmProblem G;
  G.A.d = Phi1, G.A.ld = ld1, G.A.off = off1, G.A.trans = 0;
  G.B.d = C, G.B.ld = k1, G.B.batch_stride = int64_t(k1) * k2, G.B.rows = k2, G.B.trans = 0;
  G.N = k2, G.K = k1, G.maxM = max_n1, G.maxN = k2, G.maxK = k1, G.n_batch = n_pairs;
  G.C = W, G.ldc = k2, G.c_off = off1;
  if ((rc = gemm64_launch(G, st))) return rc;
  GemmProblem H;
  H.A.d = Phi2, H.A.ld = ld2, H.A.off = off2, H.A.trans = 0;
  H.B.d = W, H.B.ld = k2, H.B.off = off1, H.B.trans = 0;
  H.K = k2, H.maxM = max_n2, H.maxN = max_n1, H.maxK = k2, H.n_batch = n_pairs;
  H.C = MI, H.ldc = ldMI, H.c_off = off2, H.c_colscale = area1, H.c_colscale_off = off1;
  return gemm64_launch(H, st);
}

// ------------------------------------------------------------------ p2p -> FM
size_t dm_p2p_to_fm_workspace_bytes(int n_pairs, int max_n2, int k1, int k2) {
  return p2p_to_fm_ws(n_pairs, max_n2, k1, k2);
}

int dm_p2p_to_fm(const void* p2p_21, const double* Phi1, int64_t ld1, const int64_t* off1, const double* Phi2,
                 int64_t ld2, const int64_t* off2, int max_n2, const double* area2, int n_pairs, int k1, int k2,
                 double* C, int flags, void* workspace, size_t workspace_bytes, dm_stream_t stream) {
  if (n_pairs < 0 || k1 <= 0 || k2 <= 0 || max_n2 < 0) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n_pairs == 0) return DM_OK;
  if (!p2p_21 || !Phi1 || !Phi2 || !off1 || !off2 || !C) DM_FAIL(DM_ERR_BADARG, "null argument");
  if (ld1 < k1 || ld2 < k2) DM_FAIL(DM_ERR_BADARG, "eigenbasis has fewer columns than requested");
  const size_t need = p2p_to_fm_ws(n_pairs, max_n2, k1, k2);
  if (need > workspace_bytes || (need && !workspace)) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu", need);
  return p2p_to_fm_run(p2p_21, (flags & DM_I64_OUT) ? 1 : 0, Phi1, ld1, off1, Phi2, ld2, off2, max_n2, area2, n_pairs,
                       k1, k2, C, workspace, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------ ZoomOut
size_t dm_zoomout_workspace_bytes(int n_pairs, int64_t total_n1, int64_t total_n2, int max_n1, int max_n2, int k1_0,
                                  int k2_0, int nit, int step1, int step2, int flags) {
  const int k1m = k1_0 + nit * step1, k2m = k2_0 + nit * step2;
  Carver c(nullptr);
  c.take<double>(size_t(n_pairs) * k1m * k2m);  // C ping
  c.take<double>(size_t(n_pairs) * k1m * k2m);  // C pong
  c.take<double>(size_t(total_n1) * k2m);       // emb1
  c.take<float>(size_t(total_n1) * pad4(k2m));  // fp32 emb1
  c.take<float>(size_t(total_n2) * pad4(k2m));  // fp32 Phi2
  c.take<int32_t>(size_t(total_n2) * 2);        // p2p (int32 or int64)
  c.take<int32_t>(size_t(total_n2) * 2);        // p2p of the previous rung
  c.take<char>(p2p_to_fm_delta_ws(n_pairs, total_n2));
  c.take<double>(size_t(n_pairs) * k1m * k2m);  // M: the full-width map kept resident by the incremental rungs
  c.take<char>(zoomout_pf_ws(n_pairs, total_n2, max_n2, k1m, k2m, flags));
  c.take<char>(nn_workspace_bytes(n_pairs, total_n2, total_n1, max_n2, max_n1, k2m, 1, 0, flags));
  c.take<char>(p2p21_factored_scratch_bytes(n_pairs, total_n1, k1m, k2m));
  return c.bytes();
}

int dm_zoomout(const double* C0, int k1_0, int k2_0, int nit, int step1, int step2, const double* Phi1, int64_t ld1,
               const int64_t* off1, int64_t total_n1, int max_n1, const double* Phi2, int64_t ld2, const int64_t* off2,
               int64_t total_n2, int max_n2, const double* area2, int n_pairs, double* C_out, void* p2p_out, int flags,
               void* workspace, size_t workspace_bytes, dm_stream_t stream) {
  if (n_pairs < 0 || k1_0 <= 0 || k2_0 <= 0 || nit < 0 || step1 < 0 || step2 < 0) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n_pairs == 0) return DM_OK;
  if (!C0 || !Phi1 || !Phi2 || !off1 || !off2 || !area2 || !C_out) DM_FAIL(DM_ERR_BADARG, "null argument");
  const int k1m = k1_0 + nit * step1, k2m = k2_0 + nit * step2;
  if (ld1 < k1m) DM_FAIL(DM_ERR_BADARG, "Not enough eigenvectors on source : %d are needed when %lld are provided", k1m, (long long)ld1);
  if (ld2 < k2m) DM_FAIL(DM_ERR_BADARG, "Not enough eigenvectors on target : %d are needed when %lld are provided", k2m, (long long)ld2);
  const size_t need = dm_zoomout_workspace_bytes(n_pairs, total_n1, total_n2, max_n1, max_n2, k1_0, k2_0, nit, step1, step2, flags);
  if (!workspace || need > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu", need);
  if (reinterpret_cast<uintptr_t>(workspace) % 256) DM_FAIL(DM_ERR_ALIGN, "workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver c(workspace);
  double* Cbuf[2];
  Cbuf[0] = c.take<double>(size_t(n_pairs) * k1m * k2m);
  Cbuf[1] = c.take<double>(size_t(n_pairs) * k1m * k2m);
  P2P21Scratch S;
  S.lde = k2m, S.ldf = pad4(k2m);
  S.emb1 = c.take<double>(size_t(total_n1) * k2m);
  S.Xf = c.take<float>(size_t(total_n1) * S.ldf);
  float* Phi2f = c.take<float>(size_t(total_n2) * S.ldf);
  void* p2p_buf[2];
  p2p_buf[0] = c.take<int32_t>(size_t(total_n2) * 2);
  p2p_buf[1] = c.take<int32_t>(size_t(total_n2) * 2);
  void* delta_ws = c.take<char>(p2p_to_fm_delta_ws(n_pairs, total_n2));
  double* Mfull = c.take<double>(size_t(n_pairs) * k1m * k2m);
  const size_t pf_bytes = zoomout_pf_ws(n_pairs, total_n2, max_n2, k1m, k2m, flags);
  void* pf_ws = c.take<char>(pf_bytes);
  const bool fast_fm = (flags & DM_FAST_FM) && proj_tc_supported(k2m, k1m);
  S.nn_ws_bytes = nn_workspace_bytes(n_pairs, total_n2, total_n1, max_n2, max_n1, k2m, 1, 0, flags);
  S.nn_ws = c.take<char>(S.nn_ws_bytes);
  S.fact = c.take<char>(p2p21_factored_scratch_bytes(n_pairs, total_n1, k1m, k2m));
  S.fact_k1m = k1m, S.fact_k2m = k2m;
  const int i64 = (flags & DM_I64_OUT) ? 1 : 0;
  int rc;
  if (!nn_use_tc(flags) && (rc = cvt_f64_f32(Phi2, ld2, total_n2, k2m, Phi2f, S.ldf, st))) return rc;
  const double* Ccur = C0;
  int k1 = k1_0, k2 = k2_0;
  int y_kp = -1;  // padded width for which the split of Phi2 in the workspace is valid
  // The rungs keep M = Phi2[:, :k2m]^T A2 Phi1[p, :k1m] (the widths of the LAST rung) resident and correct it with the
  // vertices whose image changed since the previous rung (zoomout_delta.cu); a rung's map is M's leading block.  A full
  // product re-anchors M every kZoAnchor rungs, and the last rung is always a fresh product of its own.
  // (the gathered correction runs on the cp.async GEMM: rows of both bases on 16-byte boundaries)
  constexpr int kZoAnchor = 64;
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(Phi1) | reinterpret_cast<uintptr_t>(Phi2)) & 15) == 0 && !((ld1 | ld2) & 1);
  const bool delta_ok = !fast_fm && aligned16 && nit > 2 && p2p_to_fm_delta_applicable();
  for (int it = 0; it < nit; ++it) {
    void* p2p = p2p_buf[it & 1];
    // the fp32 copy of emb1 is re-made with the current width; stale columns beyond k2 are zeroed by cvt
    if ((rc = p2p21_run(Ccur, k1, k2, Phi1, ld1, off1, total_n1, max_n1, Phi2, ld2, Phi2f, S.ldf, off2, total_n2,
                        max_n2, n_pairs, p2p, flags, S, st, &y_kp)))
      return rc;
    double* Cnext = (it == nit - 1) ? C_out : Cbuf[it & 1];
    if (fast_fm) {
      // C = Phi2[:, :k2+s2]^T (a2 * Phi1[p, :k1+s1]) on the tensor cores: A operand = Phi2, B operand = gathered, scaled Phi1
      if ((rc = proj_tc_run(Phi2, ld2, nullptr, nullptr, Phi1, ld1, area2, p2p, i64, off1, off2, total_n2, max_n2, n_pairs,
                            k2 + step2, k1 + step1, Cnext, pf_ws, pf_bytes, st)))
        return rc;
    } else if (delta_ok && it != nit - 1) {
      if (it % kZoAnchor == 0) {
        if ((rc = p2p_to_fm_run(p2p, i64, Phi1, ld1, off1, Phi2, ld2, off2, max_n2, area2, n_pairs, k1m, k2m, Mfull, pf_ws, st)))
          return rc;
      } else if ((rc = p2p_to_fm_delta_run(p2p, p2p_buf[(it - 1) & 1], i64, Phi1, ld1, off1, Phi2, ld2, off2, total_n2, max_n2,
                                           area2, n_pairs, k1m, k2m, Mfull, delta_ws, st)))
        return rc;
      if ((rc = extract_block_run(Mfull, k1m, k2m, k1 + step1, k2 + step2, n_pairs, Cnext, st))) return rc;
    } else if ((rc = p2p_to_fm_run(p2p, i64, Phi1, ld1, off1, Phi2, ld2, off2, max_n2, area2, n_pairs, k1 + step1,
                                   k2 + step2, Cnext, pf_ws, st)))
      return rc;
    Ccur = Cnext;
    k1 += step1, k2 += step2;
  }
  if (nit == 0)
    DM_CUDA_OK(cudaMemcpyAsync(C_out, C0, sizeof(double) * size_t(n_pairs) * k1 * k2, cudaMemcpyDeviceToDevice, st));
  if (p2p_out) {
    if ((rc = p2p21_run(nit == 0 ? C0 : C_out, k1, k2, Phi1, ld1, off1, total_n1, max_n1, Phi2, ld2, Phi2f, S.ldf, off2,
                        total_n2, max_n2, n_pairs, p2p_out, flags, S, st, &y_kp)))
      return rc;
  }
  return DM_OK;
}

// ------------------------------------------------------------------ spectral ICP
namespace {
struct IcpLayout {
  double *G, *Ginv, *Phi2p, *X, *lin, *ns;
  int* status;
  P2P21Scratch S;
  float* Phi2f;
  void* p2p;
  void* pf_ws;
  size_t pf_bytes, bytes;
};
IcpLayout icp_carve(void* ws, int n_pairs, int64_t total_n1, int64_t total_n2, int max_n1, int max_n2, int k1, int k2,
                    int flags) {
  Carver c(ws);
  IcpLayout L;
  L.status = c.take<int>(64);  // first: dm_icp_read_status needs no sizes
  L.G = c.take<double>(size_t(n_pairs) * k2 * k2);
  L.Ginv = c.take<double>(size_t(n_pairs) * k2 * k2);
  L.Phi2p = c.take<double>(size_t(total_n2) * k2);
  L.X = c.take<double>(size_t(n_pairs) * k2 * k1);
  const size_t lin = spd_inverse_scratch_doubles(k2) > polar_scratch_doubles(k2, k1) ? spd_inverse_scratch_doubles(k2)
                                                                                     : polar_scratch_doubles(k2, k1);
  L.lin = c.take<double>(lin * n_pairs);
  L.ns = c.take<double>(polar_ns_scratch_doubles(k2, k1, n_pairs));
  L.S.lde = k2, L.S.ldf = pad4(k2);
  L.S.emb1 = c.take<double>(size_t(total_n1) * k2);
  const bool tc = nn_use_tc(flags);
  L.S.Xf = c.take<float>(tc ? 0 : size_t(total_n1) * L.S.ldf);
  L.Phi2f = c.take<float>(tc ? 0 : size_t(total_n2) * L.S.ldf);
  L.p2p = c.take<int64_t>(size_t(total_n2));
  const int k = k1 > k2 ? k1 : k2;
  L.pf_bytes = p2p_to_fm_ws(n_pairs, max_n2, k, k2);
  L.pf_ws = c.take<char>(L.pf_bytes);
  L.S.nn_ws_bytes = nn_workspace_bytes(n_pairs, total_n2, total_n1, max_n2, max_n1, k2, 1, 0, flags);
  L.S.nn_ws = c.take<char>(L.S.nn_ws_bytes);
  L.S.fact = c.take<char>(p2p21_factored_scratch_bytes(n_pairs, total_n1, k1, k2));
  L.S.fact_k1m = k1, L.S.fact_k2m = k2;
  L.bytes = c.bytes();
  return L;
}
}  // namespace

size_t dm_icp_workspace_bytes(int n_pairs, int64_t total_n1, int64_t total_n2, int max_n1, int max_n2, int k1, int k2,
                              int flags) {
  if (n_pairs < 0 || total_n1 < 0 || total_n2 < 0 || k1 <= 0 || k2 <= 0) return 0;
  return icp_carve(nullptr, n_pairs, total_n1, total_n2, max_n1, max_n2, k1, k2, flags).bytes;
}

int dm_icp(const double* C0, int k1, int k2, int nit, const double* Phi1, int64_t ld1, const int64_t* off1,
           int64_t total_n1, int max_n1, const double* Phi2, int64_t ld2, const int64_t* off2, int64_t total_n2,
           int max_n2, int n_pairs, double* C_out, void* p2p_out, int flags, void* workspace, size_t workspace_bytes,
           dm_stream_t stream) {
  if (n_pairs < 0 || k1 <= 0 || k2 <= 0 || nit < 0 || total_n1 < 0 || total_n2 < 0) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n_pairs == 0) return DM_OK;
  if (!C0 || !Phi1 || !Phi2 || !off1 || !off2 || !C_out) DM_FAIL(DM_ERR_BADARG, "null argument");
  if (ld1 < k1 || ld2 < k2) DM_FAIL(DM_ERR_BADARG, "eigenbasis has fewer columns than the functional map");
  if (!workspace) DM_FAIL(DM_ERR_WORKSPACE, "workspace is null");
  if (reinterpret_cast<uintptr_t>(workspace) % 256) DM_FAIL(DM_ERR_ALIGN, "workspace must be 256-byte aligned");
  IcpLayout L = icp_carve(workspace, n_pairs, total_n1, total_n2, max_n1, max_n2, k1, k2, flags);
  if (L.bytes > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu", L.bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int i64 = (flags & DM_I64_OUT) ? 1 : 0;
  int rc;
  DM_CUDA_OK(cudaMemsetAsync(L.status, 0, 64 * sizeof(int), st));
  if (nit > 0) {
    // lstsq(Phi2, Phi1[p]) = (Phi2 G^-1)^T Phi1[p] with G = Phi2^T Phi2, factorised once   (icp.py:38 -> convert.py:51)
    if ((rc = p2p_to_fm_run(nullptr, 0, Phi2, ld2, off2, Phi2, ld2, off2, max_n2, nullptr, n_pairs, k2, k2, L.G, L.pf_ws,
                            st)))
      return rc;
    if ((rc = spd_inverse_launch(L.G, L.Ginv, k2, n_pairs, L.lin, L.status, st))) return rc;
    GemmProblem G;
    G.A.d = Phi2, G.A.ld = ld2, G.A.off = off2, G.A.trans = 0;
    G.B.d = L.Ginv, G.B.ld = k2, G.B.batch_stride = int64_t(k2) * k2, G.B.rows = k2, G.B.trans = 0;
    G.N = k2, G.K = k2, G.maxM = max_n2, G.maxN = k2, G.maxK = k2, G.n_batch = n_pairs;
    G.C = L.Phi2p, G.ldc = k2, G.c_off = off2;
    if ((rc = gemm64_launch(G, st))) return rc;
  }
  if (!nn_use_tc(flags) && (rc = cvt_f64_f32(Phi2, ld2, total_n2, k2, L.Phi2f, L.S.ldf, st))) return rc;
  const double* Ccur = C0;
  int y_kp = -1;  // the split of the static query matrix Phi2 is made once
  for (int it = 0; it < nit; ++it) {
    // p = p2p_21(C)  (icp.py:37; the other two outputs of FM_to_p2p are discarded there)
    if ((rc = p2p21_run(Ccur, k1, k2, Phi1, ld1, off1, total_n1, max_n1, Phi2, ld2, L.Phi2f, L.S.ldf, off2, total_n2,
                        max_n2, n_pairs, L.p2p, flags, L.S, st, &y_kp)))
      return rc;
    if ((rc = p2p_to_fm_run(L.p2p, i64, Phi1, ld1, off1, L.Phi2p, k2, off2, max_n2, nullptr, n_pairs, k1, k2, L.X,
                            L.pf_ws, st)))
      return rc;
    // C <- U I V^T  (icp.py:39-40)
    if ((rc = polar_factor_launch(L.X, C_out, k2, k1, n_pairs, L.lin, L.ns, st))) return rc;
    Ccur = C_out;
  }
  if (nit == 0)
    DM_CUDA_OK(cudaMemcpyAsync(C_out, C0, sizeof(double) * size_t(n_pairs) * k1 * k2, cudaMemcpyDeviceToDevice, st));
  if (p2p_out) {
    if ((rc = p2p21_run(nit == 0 ? C0 : C_out, k1, k2, Phi1, ld1, off1, total_n1, max_n1, Phi2, ld2, L.Phi2f, L.S.ldf,
                        off2, total_n2, max_n2, n_pairs, p2p_out, flags, L.S, st, &y_kp)))
      return rc;
  }
  return DM_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ the whole per-pair path in one call
namespace dm {
namespace {
// c00[p] = sign(Phi1[first row of p][0] * Phi2[first row][0]) * sqrt(sum area2 / sum area1)   (functional.py:654-658)
// in1 / in2: first row of pair p in Phi / area (the batch offsets themselves, or the rows of its meshes in a bank)
__global__ void __launch_bounds__(256)
    c00_kernel(const double* __restrict__ Phi1, int64_t ld1, const int64_t* __restrict__ off1,
               const double* __restrict__ Phi2, int64_t ld2, const int64_t* __restrict__ off2,
               const double* __restrict__ area1, const double* __restrict__ area2, double* __restrict__ c00,
               const int64_t* __restrict__ in1, const int64_t* __restrict__ in2) {
  const int p = blockIdx.x, t = threadIdx.x;
  double s1 = 0.0, s2 = 0.0;
  const int64_t b1 = in1[p], e1 = b1 + (off1[p + 1] - off1[p]), b2 = in2[p], e2 = b2 + (off2[p + 1] - off2[p]);
  for (int64_t r = b1 + t; r < e1; r += 256) s1 += area1[r];
  for (int64_t r = b2 + t; r < e2; r += 256) s2 += area2[r];
  __shared__ double r1[8], r2[8];
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, sh);
    s2 += __shfl_xor_sync(0xffffffffu, s2, sh);
  }
  if ((t & 31) == 0) r1[t >> 5] = s1, r2[t >> 5] = s2;
  __syncthreads();
  if (t == 0) {
    for (int w = 1; w < 8; ++w) s1 += r1[w], s2 += r2[w];
    const double pr = Phi1[b1 * ld1] * Phi2[b2 * ld2];
    const double sg = pr > 0.0 ? 1.0 : (pr < 0.0 ? -1.0 : 0.0);
    c00[p] = sg * sqrt(s2 / s1);
  }
}

struct MatchLayout {
  void* nn_ws;
  size_t nn_bytes;
  double *A, *B, *c00;
  void* proj_ws;
  size_t proj_bytes;
  void* solve_ws;
  size_t solve_bytes;
  void* p2p_ws;
  size_t p2p_bytes, bytes;
  uint16_t *p1h, *p1m, *p1l, *p2h, *p2m, *p2l;  // three-way splits of Phi1 / Phi2 [total, pad64(k)] for FM -> p2p
  float *p1norm, *p2norm;
};
MatchLayout match_carve(void* ws, int n_pairs, int64_t total_n1, int64_t total_n2, int max_n1, int max_n2, int d, int k,
                        int flags) {
  Carver c(ws);
  MatchLayout L;
  // the solve workspace leads: its status words are then the first bytes of the whole workspace
  // (dm_match_pairs_read_status == dm_fmap_solve_read_status)
  L.solve_bytes = dm_fmap_solve_workspace_bytes(n_pairs, k, k, d);
  L.solve_ws = c.take<char>(L.solve_bytes);
  L.nn_bytes = nn_workspace_bytes(n_pairs, total_n2, total_n1, max_n2, max_n1, d, 1, 1, flags | kFlagSplit3);
  L.nn_ws = c.take<char>(L.nn_bytes);
  L.A = c.take<double>(size_t(n_pairs) * k * d);
  L.B = c.take<double>(size_t(n_pairs) * k * d);
  L.c00 = c.take<double>(size_t(n_pairs));
  const size_t p1 = proj_tc_workspace_bytes(n_pairs, total_n1, max_n1, k, d);
  const size_t p2 = proj_tc_workspace_bytes(n_pairs, total_n2, max_n2, k, d);
  L.proj_bytes = p1 > p2 ? p1 : p2;
  L.proj_ws = c.take<char>(L.proj_bytes);
  L.p2p_bytes = dm_fm_to_p2p_workspace_bytes(n_pairs, total_n1, total_n2, max_n1, max_n2, k, k, flags);
  L.p2p_ws = c.take<char>(L.p2p_bytes);
  const size_t kpK = f2p_factored_applicable(k, k, flags) ? size_t(nn_tc_kp(k)) : 0;
  const size_t r1 = kpK ? size_t(total_n1) : 0, r2 = kpK ? size_t(total_n2) : 0;
  L.p1h = c.take<uint16_t>(r1 * kpK), L.p1m = c.take<uint16_t>(r1 * kpK), L.p1l = c.take<uint16_t>(r1 * kpK);
  L.p2h = c.take<uint16_t>(r2 * kpK), L.p2m = c.take<uint16_t>(r2 * kpK), L.p2l = c.take<uint16_t>(r2 * kpK);
  L.p1norm = c.take<float>(r1), L.p2norm = c.take<float>(r2);
  L.bytes = c.bytes();
  return L;
}
}  // namespace
}  // namespace dm

extern "C" {

size_t dm_match_pairs_workspace_bytes(int n_pairs, int64_t total_n1, int64_t total_n2, int max_n1, int max_n2, int d,
                                      int k, int flags) {
  if (n_pairs < 0 || total_n1 < 0 || total_n2 < 0 || d <= 0 || k < 2) return 0;
  return match_carve(nullptr, n_pairs, total_n1, total_n2, max_n1, max_n2, d, k, flags).bytes;
}

int dm_match_pairs(const float* F1, int64_t ldF1, const float* F2, int64_t ldF2, const double* Phi1, int64_t ld1,
                   const double* Phi2, int64_t ld2, const double* area1, const double* area2, const double* evals1,
                   const double* evals2, const int64_t* off1, int64_t total_n1, int max_n1, const int64_t* off2,
                   int64_t total_n2, int max_n2, int n_pairs, int d, int k, double w_descr, double w_lap,
                   void* nn_p2p_21, void* nn_p2p_12, double* C, void* p2p_21, void* p2p_12, void* dense_21,
                   void* dense_12, int flags, void* workspace, size_t workspace_bytes, dm_stream_t stream) {
  if (n_pairs < 0 || d <= 0 || k < 2 || total_n1 < 0 || total_n2 < 0) DM_FAIL(DM_ERR_BADARG, "bad size (need k >= 2)");
  if (n_pairs == 0) return DM_OK;
  if (!F1 || !F2 || !Phi1 || !Phi2 || !area1 || !area2 || !evals1 || !evals2 || !off1 || !off2 || !C || !nn_p2p_21 ||
      !nn_p2p_12)
    DM_FAIL(DM_ERR_BADARG, "null argument");
  if (ldF1 < d || ldF2 < d || ld1 < k || ld2 < k) DM_FAIL(DM_ERR_BADARG, "leading dimension too small");
  if (total_n1 == 0 || total_n2 == 0) DM_FAIL(DM_ERR_BADARG, "empty meshes");
  if (!nn_use_tc(flags) || !proj_tc_supported(k, d)) DM_FAIL(DM_ERR_UNSUPPORTED, "dm_match_pairs needs the tensor-core engines (d <= 512)");
  if (!workspace) DM_FAIL(DM_ERR_WORKSPACE, "workspace is null");
  if (reinterpret_cast<uintptr_t>(workspace) % 256) DM_FAIL(DM_ERR_ALIGN, "workspace must be 256-byte aligned");
  MatchLayout L = match_carve(workspace, n_pairs, total_n1, total_n2, max_n1, max_n2, d, k, flags);
  if (L.bytes > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu", L.bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  // The feature search (tensor-bound score pass) and the functional-map chain up to C (HBM-bound splits / reductions,
  // latency-bound solve) are independent once the features are split: the chain runs on a side stream forked after the
  // operand preparation and joined before FM -> p2p.  Stream and events are per host thread and device (the call stays
  // stream-ordered for the caller and capturable: the fork / join are event edges).  DM_MATCH_SERIAL=1: one stream.
  struct Fork {
    cudaStream_t side = nullptr;
    cudaEvent_t prep = nullptr, done = nullptr;
  };
  static thread_local Fork forks[64];
  static const bool serial = [] { const char* e = getenv("DM_MATCH_SERIAL"); return e && e[0] == '1'; }();
  int devi = 0;
  DM_CUDA_OK(cudaGetDevice(&devi));
  Fork* fk = (!serial && devi >= 0 && devi < 64) ? &forks[devi] : nullptr;
  if (fk && !fk->side) {
    DM_CUDA_OK(cudaStreamCreateWithFlags(&fk->side, cudaStreamNonBlocking));
    DM_CUDA_OK(cudaEventCreateWithFlags(&fk->prep, cudaEventDisableTiming));
    DM_CUDA_OK(cudaEventCreateWithFlags(&fk->done, cudaEventDisableTiming));
  }
  cudaStream_t sf = fk ? fk->side : st;  // stream of the functional-map chain
  // 1. feature NN: queries = mesh 2, database = mesh 1 (rows -> p2p_21, columns -> p2p_12); keeps the three-way splits
  NNRequest R{};
  R.Y = F2, R.ldY = ldF2, R.X = F1, R.ldX = ldF1;
  R.q_off = off2, R.db_off = off1, R.total_q = total_n2, R.total_db = total_n1;
  R.max_q = max_n2, R.max_db = max_n1, R.n_pairs = n_pairs, R.d = d;
  R.n_row = 1, R.n_col = 1;
  R.row[0] = dm_nn_epi{DM_SCALE_NONE, DM_BIAS_NONE, nullptr, nullptr, nn_p2p_21};
  R.col[0] = dm_nn_epi{DM_SCALE_NONE, DM_BIAS_NONE, nullptr, nullptr, nn_p2p_12};
  R.flags = flags | kFlagSplit3;
  R.after_prep_event = fk ? fk->prep : nullptr;
  NNSplits sp{};
  if ((rc = nn_run(R, L.nn_ws, L.nn_bytes, st, &sp))) return rc;
  if (fk) DM_CUDA_OK(cudaStreamWaitEvent(sf, fk->prep, 0));
  // The operand preparation of FM -> p2p (three-way splits + row norms of both eigenbases) does not depend on C: enqueued
  // here, behind the feature search on the main stream, these HBM-bound kernels run beside the solve of the side stream
  // instead of after it.  FM -> p2p then reads them like a mesh bank whose rows are the batch's own (in = off).
  static const bool early_prep = [] { const char* e = getenv("DM_MATCH_LATE_PREP"); return !(e && e[0] == '1'); }();
  const bool want_maps = p2p_21 || p2p_12 || dense_21 || dense_12;
  const bool presplit = early_prep && want_maps && f2p_factored_applicable(k, k, flags);
  if (presplit) {
    const int kpK = nn_tc_kp(k);
    if ((rc = nn_prep_side(Phi1, 1, ld1, off1, n_pairs, total_n1, k, L.p1norm, nullptr, 0, L.p1h, L.p1m, L.p1l, kpK, st))) return rc;
    if ((rc = nn_prep_side(Phi2, 1, ld2, off2, n_pairs, total_n2, k, L.p2norm, nullptr, 0, L.p2h, L.p2m, L.p2l, kpK, st))) return rc;
  }
  // 2. projections, reusing the feature splits (mesh 1 = database side, mesh 2 = query side)
  const void* s1[3] = {sp.xh, sp.xl, sp.xl2};
  const void* s2[3] = {sp.yh, sp.yl, sp.yl2};
  if ((rc = proj_tc_run(Phi1, ld1, area1, F1, nullptr, ldF1, nullptr, nullptr, 0, nullptr, off1, total_n1, max_n1, n_pairs,
                        k, d, L.A, L.proj_ws, L.proj_bytes, sf, s1)))
    return rc;
  if ((rc = proj_tc_run(Phi2, ld2, area2, F2, nullptr, ldF2, nullptr, nullptr, 0, nullptr, off2, total_n2, max_n2, n_pairs,
                        k, d, L.B, L.proj_ws, L.proj_bytes, sf, s2)))
    return rc;
  // 3. pinned entry, closed-form C
  c00_kernel<<<unsigned(n_pairs), 256, 0, sf>>>(Phi1, ld1, off1, Phi2, ld2, off2, area1, area2, L.c00, off1, off2);
  DM_LAUNCH_OK("c00_kernel");
  if ((rc = dm_fmap_solve(L.A, L.B, evals1, evals2, L.c00, w_descr, w_lap, n_pairs, k, k, d, C, L.solve_ws, L.solve_bytes,
                          static_cast<dm_stream_t>(sf))))
    return rc;
  if (fk) {
    DM_CUDA_OK(cudaEventRecord(fk->done, sf));
    DM_CUDA_OK(cudaStreamWaitEvent(st, fk->done, 0));
  }
  // 4. the four index maps
  if (!want_maps) return DM_OK;
  if (presplit) {
    const NNBankSide b1{off1, total_n1, L.p1h, L.p1m, L.p1l, L.p1norm}, b2{off2, total_n2, L.p2h, L.p2m, L.p2l, L.p2norm};
    if (dense_21 && !area1) DM_FAIL(DM_ERR_BADARG, "dense_21 needs area1");
    return f2p_factored_run(C, k, k, Phi1, ld1, off1, total_n1, max_n1, Phi2, ld2, off2, total_n2, max_n2, area1, n_pairs, p2p_21,
                            p2p_12, dense_21, dense_12, flags, L.p2p_ws, st, &b1, &b2);
  }
  return dm_fm_to_p2p(C, k, k, Phi1, ld1, off1, total_n1, max_n1, Phi2, ld2, off2, total_n2, max_n2, area1, n_pairs, p2p_21,
                      p2p_12, dense_21, dense_12, flags, L.p2p_ws, L.p2p_bytes, stream);
}

int dm_match_pairs_read_status(const void* workspace, int* out_h, dm_stream_t stream) {
  return dm_fmap_solve_read_status(workspace, out_h, stream);
}

int dm_icp_read_status(const void* workspace, int* out_h, dm_stream_t stream) {
  if (!workspace || !out_h) DM_FAIL(DM_ERR_BADARG, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DM_CUDA_OK(cudaMemcpyAsync(out_h, workspace, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  DM_CUDA_OK(cudaStreamSynchronize(st));
  return DM_OK;
}

// ------------------------------------------------------------------ batched SPD solve (normal equations of the lstsq branch)
size_t dm_spd_solve_workspace_bytes(int n_batch, int n) {
  if (n_batch < 0 || n <= 0) return 0;
  Carver c(nullptr);
  c.take<int>(64);
  c.take<double>(size_t(n_batch) * n * n);
  c.take<double>(spd_inverse_scratch_doubles(n) * size_t(n_batch));
  return c.bytes();
}

int dm_spd_solve(const double* G, const double* B, int n, int m, int n_batch, double* X, void* workspace,
                 size_t workspace_bytes, dm_stream_t stream) {
  if (n_batch < 0 || n <= 0 || m <= 0) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n_batch == 0) return DM_OK;
  if (!G || !B || !X) DM_FAIL(DM_ERR_BADARG, "null argument");
  if (!workspace || dm_spd_solve_workspace_bytes(n_batch, n) > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) % 256) DM_FAIL(DM_ERR_ALIGN, "workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver c(workspace);
  int* status = c.take<int>(64);  // first: dm_icp_read_status reads it
  double* Ginv = c.take<double>(size_t(n_batch) * n * n);
  double* lin = c.take<double>(spd_inverse_scratch_doubles(n) * size_t(n_batch));
  DM_CUDA_OK(cudaMemsetAsync(status, 0, 64 * sizeof(int), st));
  int rc;
  if ((rc = spd_inverse_launch(G, Ginv, n, n_batch, lin, status, st))) return rc;
  GemmProblem P;
  P.A.d = Ginv, P.A.ld = n, P.A.batch_stride = int64_t(n) * n, P.A.rows = n, P.A.trans = 0;
  P.B.d = B, P.B.ld = m, P.B.batch_stride = int64_t(n) * m, P.B.rows = n, P.B.trans = 1;
  P.M = n, P.N = m, P.K = n, P.maxM = n, P.maxN = m, P.maxK = n, P.n_batch = n_batch;
  P.C = X, P.ldc = m, P.c_batch_stride = int64_t(n) * m;
  return gemm64_launch(P, st);
}

// ------------------------------------------------------------------ polar factor (the SVD step of ICP, exposed)
size_t dm_polar_factor_workspace_bytes(int n_batch, int rows, int cols) {
  if (n_batch < 0 || rows <= 0 || cols <= 0) return 0;
  Carver c(nullptr);
  c.take<double>(polar_scratch_doubles(rows, cols) * size_t(n_batch));
  c.take<double>(polar_ns_scratch_doubles(rows, cols, n_batch));
  return c.bytes();
}

int dm_polar_factor(const double* X, int rows, int cols, int n_batch, double* C, int flags, void* workspace,
                    size_t workspace_bytes, dm_stream_t stream) {
  if (n_batch < 0 || rows <= 0 || cols <= 0) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n_batch == 0) return DM_OK;
  if (!X || !C || X == C) DM_FAIL(DM_ERR_BADARG, "null or aliased argument");
  const size_t need = dm_polar_factor_workspace_bytes(n_batch, rows, cols);
  if (!workspace || need > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu", need);
  Carver c(workspace);
  double* jac = c.take<double>(polar_scratch_doubles(rows, cols) * size_t(n_batch));
  double* ns = c.take<double>(polar_ns_scratch_doubles(rows, cols, n_batch));
  return polar_factor_launch(X, C, rows, cols, n_batch, jac, (flags & DM_POLAR_JACOBI) ? nullptr : ns,
                             static_cast<cudaStream_t>(stream));
}

}  // extern "C"

extern "C" {
// c00[p] = sign(Phi1[first vertex of pair p][0] * Phi2[...][0]) * sqrt(area(mesh 2) / area(mesh 1)): the pinned entry
// x0[0, 0] of FunctionalMapping.fit (pyFM/functional.py:654-658), for a ragged batch.
int dm_fmap_c00(const double* Phi1, int64_t ld1, const int64_t* off1, const double* Phi2, int64_t ld2, const int64_t* off2,
                const double* area1, const double* area2, int n_pairs, double* c00, dm_stream_t stream) {
  if (n_pairs < 0) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n_pairs == 0) return DM_OK;
  if (!Phi1 || !Phi2 || !off1 || !off2 || !area1 || !area2 || !c00) DM_FAIL(DM_ERR_BADARG, "null argument");
  c00_kernel<<<unsigned(n_pairs), 256, 0, static_cast<cudaStream_t>(stream)>>>(Phi1, ld1, off1, Phi2, ld2, off2, area1, area2, c00,
                                                                               off1, off2);
  DM_LAUNCH_OK("c00_kernel");
  return DM_OK;
}
}  // extern "C"

// ------------------------------------------------------------------ mesh bank: per-mesh preparation, pairs as id lists
// Dataset-shaped work (BASELINE config 5: 599 meshes, every intra-category pair; the evaluation loop around
// compute_surface_map, functional_map.py:9-81) matches each mesh against ~50 others.  Everything of the per-pair path that
// depends on ONE mesh only is done once per mesh here -- the bf16 splits of its features (feature search operands), the
// three-way split of its eigenbasis (embedding operands of FM -> p2p), the row norms, the projection Phi^T A F -- and the
// per-pair call reads those through each pair's mesh rows: no per-pair gather of the meshes' matrices, no per-pair
// operand preparation, no per-pair projection.  Results are bit-identical to dm_match_pairs on the assembled batch.
namespace dm {
namespace {
struct BankLayout {
  uint16_t *Fh, *Fl;        // [total_n, pad64(d)] bf16 splits of the features
  float* Fnorm;             // [total_n]
  uint16_t *Ph, *Pm, *Pl;   // [total_n, pad64(k)] three-way split of Phi[:, :k]
  float* Pnorm;             // [total_n]
  double* A;                // [n_meshes, k, d] Phi^T diag(area) F
  size_t bytes;
};
BankLayout bank_carve(void* state, int n_meshes, int64_t total_n, int d, int k) {
  Carver c(state);
  BankLayout L;
  const size_t kpF = size_t(nn_tc_kp(d)), kpK = size_t(nn_tc_kp(k));
  L.Fh = c.take<uint16_t>(size_t(total_n) * kpF);
  L.Fl = c.take<uint16_t>(size_t(total_n) * kpF);
  L.Fnorm = c.take<float>(size_t(total_n));
  L.Ph = c.take<uint16_t>(size_t(total_n) * kpK);
  L.Pm = c.take<uint16_t>(size_t(total_n) * kpK);
  L.Pl = c.take<uint16_t>(size_t(total_n) * kpK);
  L.Pnorm = c.take<float>(size_t(total_n));
  L.A = c.take<double>(size_t(n_meshes) * k * d);
  L.bytes = c.bytes();
  return L;
}
constexpr int kBankProjChunk = 128;  // meshes per projection launch (bounds the split-K partials)

struct BankPrepLayout {
  uint16_t* Fl2;  // third split of the features: only the projection reads it
  void* proj_ws;
  size_t proj_bytes, bytes;
};
BankPrepLayout bank_prep_carve(void* ws, int n_meshes, int64_t total_n, int max_n, int d, int k) {
  Carver c(ws);
  BankPrepLayout L;
  L.Fl2 = c.take<uint16_t>(size_t(total_n) * nn_tc_kp(d));
  const int nb = n_meshes < kBankProjChunk ? n_meshes : kBankProjChunk;
  L.proj_bytes = proj_tc_workspace_bytes(nb, total_n, max_n, k, d, true);
  L.proj_ws = c.take<char>(L.proj_bytes);
  L.bytes = c.bytes();
  return L;
}

// in[p] = bank_off[ids[p]] for both sides; an id outside the prepared range of meshes, or a size mismatch between the batch
// offsets and the bank, raises the flag that dm_match_bank_pairs_read_status reports
__global__ void __launch_bounds__(256)
    bank_starts_kernel(const int64_t* __restrict__ bank_off, int mesh_lo, int mesh_hi, const int64_t* __restrict__ ids1,
                       const int64_t* __restrict__ ids2, const int64_t* __restrict__ off1, const int64_t* __restrict__ off2,
                       int n_pairs, int64_t* __restrict__ in1, int64_t* __restrict__ in2, int* __restrict__ bad) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const int64_t a = ids1[p], b = ids2[p];
  const bool ok = a >= mesh_lo && a < mesh_hi && b >= mesh_lo && b < mesh_hi;  // inside the prepared range of the bank
  const int64_t ac = ok ? a : mesh_lo, bc = ok ? b : mesh_lo;
  in1[p] = bank_off[ac], in2[p] = bank_off[bc];
  if (!ok || bank_off[ac + 1] - bank_off[ac] != off1[p + 1] - off1[p] || bank_off[bc + 1] - bank_off[bc] != off2[p + 1] - off2[p])
    atomicExch(bad, 1);
}

// dst[p][j] = src[ids[p] * src_stride + j], j < len (per-mesh blocks -> per-pair blocks: projections, eigenvalues)
__global__ void __launch_bounds__(256)
    gather_blocks_kernel(const double* __restrict__ src, int64_t src_stride, const int64_t* __restrict__ ids, int len,
                         double* __restrict__ dst) {
  const int p = blockIdx.x;
  const double* s = src + ids[p] * src_stride;
  double* d = dst + int64_t(p) * len;
  for (int j = blockIdx.y * blockDim.x + threadIdx.x; j < len; j += gridDim.y * blockDim.x) d[j] = s[j];
}

struct BankMatchLayout {
  void* solve_ws;
  size_t solve_bytes;
  int* bad;
  int64_t *in1, *in2;
  double *A, *B, *ev1, *ev2, *c00;
  void* nn_ws;
  size_t nn_bytes;
  void* p2p_ws;
  size_t p2p_bytes, bytes;
};
BankMatchLayout bank_match_carve(void* ws, int n_pairs, int64_t total_n1, int64_t total_n2, int max_n1, int max_n2, int d, int k,
                                 int flags) {
  Carver c(ws);
  BankMatchLayout L;
  L.solve_bytes = dm_fmap_solve_workspace_bytes(n_pairs, k, k, d);  // leads: its status words open the workspace
  L.solve_ws = c.take<char>(L.solve_bytes);
  L.bad = c.take<int>(64);
  L.in1 = c.take<int64_t>(size_t(n_pairs));
  L.in2 = c.take<int64_t>(size_t(n_pairs));
  L.A = c.take<double>(size_t(n_pairs) * k * d);
  L.B = c.take<double>(size_t(n_pairs) * k * d);
  L.ev1 = c.take<double>(size_t(n_pairs) * k);
  L.ev2 = c.take<double>(size_t(n_pairs) * k);
  L.c00 = c.take<double>(size_t(n_pairs));
  L.nn_bytes = nn_workspace_bytes(n_pairs, total_n2, total_n1, max_n2, max_n1, d, 1, 1, flags | kFlagBankQ | kFlagBankDb);
  L.nn_ws = c.take<char>(L.nn_bytes);
  L.p2p_bytes = f2p_factored_workspace_bytes(n_pairs, total_n1, total_n2, max_n1, max_n2, k, k, flags);
  L.p2p_ws = c.take<char>(L.p2p_bytes);
  L.bytes = c.bytes();
  return L;
}
}  // namespace
}  // namespace dm

extern "C" {

size_t dm_bank_state_bytes(int n_meshes, int64_t total_n, int d, int k) {
  if (n_meshes < 0 || total_n < 0 || d <= 0 || k < 2) return 0;
  return bank_carve(nullptr, n_meshes, total_n, d, k).bytes;
}

size_t dm_bank_prepare_workspace_bytes(int n_meshes, int64_t total_n, int max_n, int d, int k) {
  if (n_meshes < 0 || total_n < 0 || max_n < 0 || d <= 0 || k < 2) return 0;
  return bank_prep_carve(nullptr, n_meshes, total_n, max_n, d, k).bytes;
}

int dm_bank_prepare(const float* F, int64_t ldF, const double* Phi, int64_t ldPhi, const double* area, const int64_t* off,
                    int64_t total_n, int max_n, int n_meshes, int d, int k, int mesh_lo, int mesh_hi, int64_t row_lo,
                    int64_t row_hi, void* state, size_t state_bytes, void* workspace, size_t workspace_bytes,
                    dm_stream_t stream) {
  if (n_meshes < 0 || total_n < 0 || max_n < 0 || d <= 0 || k < 2) DM_FAIL(DM_ERR_BADARG, "bad size (need k >= 2)");
  if (mesh_lo < 0 || mesh_hi > n_meshes || mesh_lo > mesh_hi || row_lo < 0 || row_hi > total_n || row_lo > row_hi)
    DM_FAIL(DM_ERR_BADARG, "bad mesh / row range");
  if (n_meshes == 0 || total_n == 0 || mesh_lo == mesh_hi || row_lo == row_hi) return DM_OK;
  if (!F || !Phi || !area || !off || !state) DM_FAIL(DM_ERR_BADARG, "null argument");
  if (ldF < d || ldPhi < k) DM_FAIL(DM_ERR_BADARG, "leading dimension too small");
  if (!proj_tc_supported(k, d) || !f2p_factored_applicable(k, k, 0))
    DM_FAIL(DM_ERR_UNSUPPORTED, "the mesh bank needs the tensor-core engines (d <= 512, k <= 128)");
  if (reinterpret_cast<uintptr_t>(state) % 256 || (workspace && reinterpret_cast<uintptr_t>(workspace) % 256))
    DM_FAIL(DM_ERR_ALIGN, "state and workspace must be 256-byte aligned");
  BankLayout S = bank_carve(state, n_meshes, total_n, d, k);
  if (S.bytes > state_bytes) DM_FAIL(DM_ERR_WORKSPACE, "bank state too small: need %zu", S.bytes);
  BankPrepLayout W = bank_prep_carve(workspace, n_meshes, total_n, max_n, d, k);
  if (!workspace || W.bytes > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu", W.bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  // the same kernels, per row, as the per-pair preparation of dm_match_pairs / dm_fm_to_p2p: identical bits
  // (row-wise kernels: the range is addressed by shifting every array to its first row)
  const size_t kpF = size_t(nn_tc_kp(d)), kpK = size_t(nn_tc_kp(k));
  const int64_t nr = row_hi - row_lo;
  if ((rc = nn_prep_side(F + row_lo * ldF, 0, ldF, off, 0, nr, d, S.Fnorm + row_lo, nullptr, 0, S.Fh + row_lo * kpF,
                         S.Fl + row_lo * kpF, W.Fl2 + row_lo * kpF, int(kpF), st)))
    return rc;
  if ((rc = nn_prep_side(Phi + row_lo * ldPhi, 1, ldPhi, off, 0, nr, k, S.Pnorm + row_lo, nullptr, 0, S.Ph + row_lo * kpK,
                         S.Pm + row_lo * kpK, S.Pl + row_lo * kpK, int(kpK), st)))
    return rc;
  const void* fs[3] = {S.Fh, S.Fl, W.Fl2};
  for (int m0 = mesh_lo; m0 < mesh_hi; m0 += kBankProjChunk) {
    const int nb = mesh_hi - m0 < kBankProjChunk ? mesh_hi - m0 : kBankProjChunk;
    // (the projection of a mesh sums fixed 256-vertex slices: it does not depend on the batch the mesh is in)
    if ((rc = proj_tc_run(Phi, ldPhi, area, F, nullptr, ldF, nullptr, nullptr, 0, nullptr, off + m0, total_n, max_n, nb, k, d,
                          S.A + size_t(m0) * k * d, W.proj_ws, W.proj_bytes, st, fs)))
      return rc;
  }
  return DM_OK;
}

size_t dm_match_bank_pairs_workspace_bytes(int n_pairs, int64_t total_n1, int64_t total_n2, int max_n1, int max_n2, int d,
                                           int k, int flags) {
  if (n_pairs < 0 || total_n1 < 0 || total_n2 < 0 || d <= 0 || k < 2) return 0;
  return bank_match_carve(nullptr, n_pairs, total_n1, total_n2, max_n1, max_n2, d, k, flags).bytes;
}

int dm_match_bank_pairs(const void* state, size_t state_bytes, const float* F, int64_t ldF, const double* Phi, int64_t ldPhi,
                        const double* area, const double* evals, int64_t ld_evals, const int64_t* bank_off, int64_t total_n,
                        int n_meshes, int mesh_lo, int mesh_hi, const int64_t* ids1, const int64_t* ids2, const int64_t* off1, int64_t total_n1,
                        int max_n1, const int64_t* off2, int64_t total_n2, int max_n2, int n_pairs, int d, int k,
                        double w_descr, double w_lap, void* nn_p2p_21, void* nn_p2p_12, double* C, void* p2p_21, void* p2p_12,
                        void* dense_21, void* dense_12, int flags, void* workspace, size_t workspace_bytes,
                        dm_stream_t stream) {
  if (n_pairs < 0 || d <= 0 || k < 2 || total_n1 < 0 || total_n2 < 0 || n_meshes <= 0 || total_n <= 0)
    DM_FAIL(DM_ERR_BADARG, "bad size (need k >= 2 and a non-empty bank)");
  if (n_pairs == 0) return DM_OK;
  if (!state || !F || !Phi || !area || !evals || !bank_off || !ids1 || !ids2 || !off1 || !off2 || !C || !nn_p2p_21 || !nn_p2p_12)
    DM_FAIL(DM_ERR_BADARG, "null argument");
  if (ldF < d || ldPhi < k || ld_evals < k) DM_FAIL(DM_ERR_BADARG, "leading dimension too small");
  if (total_n1 == 0 || total_n2 == 0) DM_FAIL(DM_ERR_BADARG, "empty meshes");
  if (mesh_lo < 0 || mesh_hi > n_meshes || mesh_lo >= mesh_hi) DM_FAIL(DM_ERR_BADARG, "bad prepared range of meshes");
  if (!nn_use_tc(flags) || !proj_tc_supported(k, d) || !f2p_factored_applicable(k, k, flags))
    DM_FAIL(DM_ERR_UNSUPPORTED, "dm_match_bank_pairs needs the tensor-core engines (d <= 512, k <= 128)");
  if (!workspace) DM_FAIL(DM_ERR_WORKSPACE, "workspace is null");
  if (reinterpret_cast<uintptr_t>(workspace) % 256 || reinterpret_cast<uintptr_t>(state) % 256)
    DM_FAIL(DM_ERR_ALIGN, "state and workspace must be 256-byte aligned");
  BankLayout S = bank_carve(const_cast<void*>(state), n_meshes, total_n, d, k);
  if (S.bytes > state_bytes) DM_FAIL(DM_ERR_WORKSPACE, "bank state too small: need %zu (prepared with other sizes?)", S.bytes);
  BankMatchLayout L = bank_match_carve(workspace, n_pairs, total_n1, total_n2, max_n1, max_n2, d, k, flags);
  if (L.bytes > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu", L.bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  // the functional-map chain up to C (gathers of the per-mesh projections / eigenvalues, pinned entry, solve) only needs
  // the row starts: it runs on a side stream beside the feature score pass and is joined before FM -> p2p
  // (stream-ordered for the caller and capturable, like dm_match_pairs).  DM_MATCH_SERIAL=1: one stream.
  struct Fork {
    cudaStream_t side = nullptr;
    cudaEvent_t start = nullptr, done = nullptr;
  };
  static thread_local Fork forks[64];
  static const bool serial = [] { const char* e = getenv("DM_MATCH_SERIAL"); return e && e[0] == '1'; }();
  int devi = 0;
  DM_CUDA_OK(cudaGetDevice(&devi));
  Fork* fk = (!serial && devi >= 0 && devi < 64) ? &forks[devi] : nullptr;
  if (fk && !fk->side) {
    DM_CUDA_OK(cudaStreamCreateWithFlags(&fk->side, cudaStreamNonBlocking));
    DM_CUDA_OK(cudaEventCreateWithFlags(&fk->start, cudaEventDisableTiming));
    DM_CUDA_OK(cudaEventCreateWithFlags(&fk->done, cudaEventDisableTiming));
  }
  cudaStream_t sf = fk ? fk->side : st;
  DM_CUDA_OK(cudaMemsetAsync(L.bad, 0, 64 * sizeof(int), st));
  bank_starts_kernel<<<unsigned((n_pairs + 255) / 256), 256, 0, st>>>(bank_off, mesh_lo, mesh_hi, ids1, ids2, off1, off2, n_pairs,
                                                                     L.in1, L.in2, L.bad);
  DM_LAUNCH_OK("bank_starts_kernel");
  if (fk) {
    DM_CUDA_OK(cudaEventRecord(fk->start, st));
    DM_CUDA_OK(cudaStreamWaitEvent(sf, fk->start, 0));
  }
  // 1. chain: A / B / eigenvalues of each pair's meshes, pinned entry, closed-form C
  {
    const int len = k * d;
    gather_blocks_kernel<<<dim3(unsigned(n_pairs), unsigned((len + 1023) / 1024)), 256, 0, sf>>>(S.A, int64_t(len), ids1, len, L.A);
    gather_blocks_kernel<<<dim3(unsigned(n_pairs), unsigned((len + 1023) / 1024)), 256, 0, sf>>>(S.A, int64_t(len), ids2, len, L.B);
    gather_blocks_kernel<<<dim3(unsigned(n_pairs), 1), 256, 0, sf>>>(evals, ld_evals, ids1, k, L.ev1);
    gather_blocks_kernel<<<dim3(unsigned(n_pairs), 1), 256, 0, sf>>>(evals, ld_evals, ids2, k, L.ev2);
    DM_LAUNCH_OK("gather_blocks_kernel");
    c00_kernel<<<unsigned(n_pairs), 256, 0, sf>>>(Phi, ldPhi, off1, Phi, ldPhi, off2, area, area, L.c00, L.in1, L.in2);
    DM_LAUNCH_OK("c00_kernel");
    if ((rc = dm_fmap_solve(L.A, L.B, L.ev1, L.ev2, L.c00, w_descr, w_lap, n_pairs, k, k, d, C, L.solve_ws, L.solve_bytes,
                            static_cast<dm_stream_t>(sf))))
      return rc;
    if (fk) DM_CUDA_OK(cudaEventRecord(fk->done, sf));
  }
  // 2. feature search: queries = mesh 2, database = mesh 1, both read from the bank's splits
  const NNBankSide fq{L.in2, total_n, S.Fh, S.Fl, nullptr, S.Fnorm}, fdb{L.in1, total_n, S.Fh, S.Fl, nullptr, S.Fnorm};
  NNRequest R{};
  R.Y = F, R.ldY = ldF, R.X = F, R.ldX = ldF;
  R.q_off = off2, R.db_off = off1, R.total_q = total_n2, R.total_db = total_n1;
  R.max_q = max_n2, R.max_db = max_n1, R.n_pairs = n_pairs, R.d = d;
  R.n_row = 1, R.n_col = 1;
  R.row[0] = dm_nn_epi{DM_SCALE_NONE, DM_BIAS_NONE, nullptr, nullptr, nn_p2p_21};
  R.col[0] = dm_nn_epi{DM_SCALE_NONE, DM_BIAS_NONE, nullptr, nullptr, nn_p2p_12};
  R.flags = flags;
  R.bank_q = &fq, R.bank_db = &fdb;
  if ((rc = nn_run(R, L.nn_ws, L.nn_bytes, st))) return rc;
  if (fk) DM_CUDA_OK(cudaStreamWaitEvent(st, fk->done, 0));
  // 3. the four index maps
  if (!p2p_21 && !p2p_12 && !dense_21 && !dense_12) return DM_OK;
  const NNBankSide b1{L.in1, total_n, S.Ph, S.Pm, S.Pl, S.Pnorm}, b2{L.in2, total_n, S.Ph, S.Pm, S.Pl, S.Pnorm};
  return f2p_factored_run(C, k, k, Phi, ldPhi, off1, total_n1, max_n1, Phi, ldPhi, off2, total_n2, max_n2, area, n_pairs, p2p_21,
                          p2p_12, dense_21, dense_12, flags, L.p2p_ws, st, &b1, &b2);
}

// [0..3]: the solve stage's status words (dm_fmap_solve_read_status); [4] != 0: an id was outside the prepared range of meshes
// or the batch offsets do not match the sizes of the meshes in the bank (results are then meaningless)
int dm_match_bank_pairs_read_status(const void* workspace, int n_pairs, int d, int k, int* out_h /* [5] */, dm_stream_t stream) {
  if (!workspace || !out_h) DM_FAIL(DM_ERR_BADARG, "null argument");
  BankMatchLayout L = bank_match_carve(const_cast<void*>(workspace), n_pairs, 0, 0, 0, 0, d, k, 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DM_CUDA_OK(cudaMemcpyAsync(out_h, workspace, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  DM_CUDA_OK(cudaMemcpyAsync(out_h + 4, L.bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  DM_CUDA_OK(cudaStreamSynchronize(st));
  return DM_OK;
}

}  // extern "C"
