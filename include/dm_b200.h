/*
 * dm_b200.h -- C ABI of libdm_b200.so: the B200 (sm_100a) implementation of the
 * DenseMatcher correspondence hot path (feature NN / functional-map solve /
 * FM->p2p / ZoomOut / spectral ICP).
 *
 * The reference has no FFI layer: its boundary is Python call signatures over
 * numpy arrays.  Each entry point below therefore cites
 * the reference *Python function* (file:line, relative to the upstream repo root
 * JunzheJosephZhu/DenseMatcher @ 96da493) whose arithmetic it replaces; the
 * ctypes binding a maintainer would add on the reference side is shown in
 * INTEGRATION.md and lives in densematcher_b200/_lib.py.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _h (host);
 *   - matrices are row-major with a leading dimension in ELEMENTS;
 *   - a batch of mesh pairs is "ragged-packed": the rows of all pairs are
 *     stacked in one matrix and an int64 offsets array of n_pairs+1 entries
 *     (device) delimits pair p as rows off[p] .. off[p+1]-1;
 *   - all calls are stream-ordered and asynchronous w.r.t. the host, never
 *     synchronise, never allocate: the caller owns every buffer including the
 *     workspace, whose size the matching *_workspace_bytes call returns;
 *   - no global state except a thread-local error string; thread-safe;
 *   - return value: 0 = OK, <0 = one of DM_ERR_*; never throws.
 *   - index outputs are int32, or int64 when DM_I64_OUT is set (the reference
 *     returns int64, nn_utils.py:30 / numpy argmax).
 */
#ifndef DM_B200_H
#define DM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DM_VERSION 200

typedef void* dm_stream_t; /* a cudaStream_t */

enum {
  DM_OK = 0,
  DM_ERR_BADARG = -1,
  DM_ERR_ALIGN = -2,
  DM_ERR_WORKSPACE = -3,
  DM_ERR_CUDA = -4,
  DM_ERR_UNSUPPORTED = -5
};

/* flags */
enum {
  DM_I64_OUT = 1 << 0,      /* index outputs are int64 instead of int32 */
  DM_NO_RECHECK = 1 << 1,   /* skip the float64 near-tie re-evaluation (results are then only fp32-grade) */
  DM_ENGINE_FFMA = 1 << 2,  /* force the CUDA-core fp32 score kernel (default: tcgen05 split-bf16 tensor-core kernel) */
  DM_ENGINE_TC = 1 << 3,    /* force the tcgen05 split-bf16 tensor-core score kernel (the default) */
  DM_RECHECK_ALL = 1 << 4,  /* testing: send every row through the float64 path */
  DM_SKIP_PREP = 1 << 5,    /* profiling: reuse the operand preparation a previous identical call left in the workspace */
  DM_SKIP_FINISH = 1 << 6,  /* profiling: stop after the score kernel (no column finalisation, no re-evaluation) */
  DM_F64_GEMM = 1 << 7,     /* projection: float64 CUDA-core contraction instead of the tcgen05 split-bf16 engine */
  DM_FAST_FM = 1 << 8,      /* ZoomOut: form C = Phi2^T A2 Phi1[p] on the tcgen05 split-bf16 engine (fp32-grade, ~2e-6)
                               instead of float64; p2p near ties may then resolve differently from the float64 reference */
  DM_FAST_LOSS = 1 << 9     /* dm_dense_energy_ex: logarithm / division of the entropy term in float32 (the reference's own
                               precision for these terms); energies and gradients then agree with float64 to ~1e-7 */
};

/* how the per-element scale / bias of one argmax epilogue is obtained */
enum { DM_SCALE_NONE = 0, DM_SCALE_ARRAY = 1, DM_SCALE_INVNORM = 2 };
enum { DM_BIAS_NONE = 0, DM_BIAS_ARRAY = 1, DM_BIAS_NEG_HALF_SQNORM = 2 };

/*
 * One fused argmax epilogue over the score matrix S = Y X^T of a pair.
 * A ROW epilogue reduces over database rows j:  out[i] = argmax_j S_ij * scale[j] + bias[j]
 * A COL epilogue reduces over query rows i:     out[j] = argmax_i S_ij * scale[i] + bias[i]
 * scale/bias are indexed by the reduced-over side (packed like that side's matrix),
 * float64 because the near-tie re-evaluation is done in float64 like the reference.
 *   cosine NN on unit rows (model.py:169 + nn_utils.py:4-38):  NONE / NONE
 *   cosine NN on arbitrary rows:                                INVNORM / NONE
 *   Euclidean 1-NN == knn_query (nn_utils.py:4-38):             NONE / NEG_HALF_SQNORM
 *   dense mapped-indicator argmax (functional_map.py:49-50):    ARRAY (= vertex areas) / NONE
 * Ties resolve to the lowest index (numpy argmax).
 */
typedef struct dm_nn_epi {
  int32_t scale_mode;
  int32_t bias_mode;
  const double* scale; /* DM_SCALE_ARRAY only */
  const double* bias;  /* DM_BIAS_ARRAY only */
  void* out;           /* [rows of the kept side] int32 | int64 */
} dm_nn_epi;

const char* dm_last_error(void);
int dm_version(void);
/* "sm_100a;tc=1;..." -- what the library was compiled with */
const char* dm_build_info(void);

/* ------------------------------------------------------------------------------------------
 * Fused similarity + argmax (never materialises S).
 * Replaces: knn_query  densematcher/pyFM/spectral/nn_utils.py:4-38  (k = 1)
 *           dense argmax override  densematcher/functional_map.py:49-50, :76-77
 * Y: queries  [total_q,  d] (ld = ldY), pair p = rows q_off[p]..q_off[p+1]
 * X: database [total_db, d] (ld = ldX), pair p = rows db_off[p]..db_off[p+1]
 * max_q / max_db: upper bounds of the per-pair row counts (host knowledge; sizes the grid).
 * Up to 2 row and 2 column epilogues share one pass over S.
 * Scores are evaluated in fp32-grade arithmetic (three bf16 tcgen05 passes over a hi/lo split of the
 * operands, fp32 accumulation in tensor memory); results whose top-2 gap is below a rigorous
 * rounding-error bound are re-evaluated in float64 so that the index equals the float64
 * argmax of the reference.
 * ---------------------------------------------------------------------------------------- */
size_t dm_nn_workspace_bytes(int n_pairs, int64_t total_q, int64_t total_db, int max_q, int max_db, int d,
                             int n_row_epi, int n_col_epi, int flags);

int dm_nn_argmax_f32(const float* Y, int64_t ldY, const int64_t* q_off, int64_t total_q, int max_q,
                     const float* X, int64_t ldX, const int64_t* db_off, int64_t total_db, int max_db,
                     int n_pairs, int d,
                     const dm_nn_epi* row_epi_h, int n_row_epi,
                     const dm_nn_epi* col_epi_h, int n_col_epi,
                     int flags, void* workspace, size_t workspace_bytes, dm_stream_t stream);

/* Same for float64 operands (the spectral embeddings of FM_to_p2p are float64 in the reference,
 * convert.py:134-140): the score pass runs on an fp32 copy made in the workspace, the near-tie
 * re-evaluation on the float64 originals.  (The tensor-core engine splits the float64 values directly.) */
size_t dm_nn_f64_workspace_bytes(int n_pairs, int64_t total_q, int64_t total_db, int max_q, int max_db, int d,
                                 int n_row_epi, int n_col_epi, int flags);
int dm_nn_argmax_f64(const double* Y, int64_t ldY, const int64_t* q_off, int64_t total_q, int max_q,
                     const double* X, int64_t ldX, const int64_t* db_off, int64_t total_db, int max_db,
                     int n_pairs, int d,
                     const dm_nn_epi* row_epi_h, int n_row_epi,
                     const dm_nn_epi* col_epi_h, int n_col_epi,
                     int flags, void* workspace, size_t workspace_bytes, dm_stream_t stream);

/* counters of the last call that used `workspace`: out_h[0] = row results re-evaluated in float64,
 * out_h[1] = column results re-evaluated, out_h[2] = their sum, out_h[3] = how many of those needed a scan of
 * the whole candidate set (the rest were decided between the two leading candidates).
 * Synchronises the stream (diagnostics only). */
int dm_nn_read_stats(const void* workspace, int64_t* out_h, dm_stream_t stream);

/* Testing aid: materialise the fp32-grade score matrix of ONE pair as the selected engine computes it
 * (S_out [nq, ldS]); used to measure the engine's rounding error against float64. */
size_t dm_nn_debug_workspace_bytes(int nq, int ndb, int d, int flags);
int dm_nn_debug_scores_f32(const float* Y, int64_t ldY, int nq, const float* X, int64_t ldX, int ndb, int d,
                           float* S_out, int64_t ldS, int flags, void* workspace, size_t workspace_bytes,
                           dm_stream_t stream);

/* k nearest neighbours (k > 1) of every row of Y among the rows of X, Euclidean, float64, one pair: knn_query with k > 1
 * (nn_utils.py:4-38; caller projection_utils.py:178).  idx [nq, k] int64 and dist [nq, k] float64, each row ordered by
 * (distance, index) like the kd-tree's result.  k <= 16, k <= ndb.  API parity, not the throughput path. */
size_t dm_knn_workspace_bytes(int nq, int ndb, int d, int k);
int dm_knn_f64(const double* Y, int64_t ldY, int nq, const double* X, int64_t ldX, int ndb, int d, int k, int64_t* idx,
               double* dist, void* workspace, size_t workspace_bytes, dm_stream_t stream);

/* Euclidean distance of matched rows: dist[i] = | Y[i] - X[idx[i]] |_2 in float64
 * (the `dists` return of knn_query, nn_utils.py:30-38).  idx is int32|int64 per flags, global row ids into X. */
int dm_match_dist_f32(const float* Y, int64_t ldY, const float* X, int64_t ldX, const void* idx, int64_t n,
                      int d, double* dist, int flags, dm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Spectral projection  out[m] = Phi_m[:, :k]^T diag(area_m) F_m     (k x d per mesh, float64 out)
 * Replaces: optimize/base_functions.py:526-532, TriMesh.project mesh/trimesh.py:533-556.
 * Phi [total_n, ldPhi] float64, area [total_n] float64, F [total_n, ldF] float32, meshes ragged-packed
 * by row_off.  Default engine: tcgen05 tensor cores on a three-way bf16 split of both operands (six bf16
 * products per fp32-grade product, fp32 accumulation in tensor memory, float64 reduction of the split-K
 * partials): relative error ~1e-6 of |Phi|.|F|, i.e. the reference's own fp32 arithmetic
 * (base_functions.py:516-532 runs in float32).  flags & DM_F64_GEMM selects the float64 CUDA-core
 * contraction (error ~1e-15) and is also the fallback when d > 512.
 * ---------------------------------------------------------------------------------------- */
size_t dm_project_workspace_bytes(int n_meshes, int64_t total_n, int max_n, int k, int d);
int dm_project(const double* Phi, int64_t ldPhi, const double* area, const float* F, int64_t ldF,
               const int64_t* row_off, int64_t total_n, int max_n, int n_meshes, int k, int d,
               double* out /* [n_meshes, k, d] */, void* workspace, size_t workspace_bytes, dm_stream_t stream);
int dm_project_ex(const double* Phi, int64_t ldPhi, const double* area, const float* F, int64_t ldF,
                  const int64_t* row_off, int64_t total_n, int max_n, int n_meshes, int k, int d,
                  double* out /* [n_meshes, k, d] */, int flags, void* workspace, size_t workspace_bytes,
                  dm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Functional-map solve: exact minimiser of  w_descr/2 |C A - B|^2 + w_lap/2 sum C^2 Delta  with column 0
 * pinned to c00 * e_0 -- the energy FunctionalMapping.fit minimises with L-BFGS-B when only the
 * descriptor and Laplacian terms are active (pyFM/functional.py:352-487, :629-660;
 * optimize/base_functions.py:31-56, :79-102, :480-763).  Rows decouple into k2 SPD systems: float32 Cholesky
 * as a preconditioner + float64 iterative refinement against the float64 Gram matrix (converges to the float64
 * solution; systems that do not contract fall back to a float64 Cholesky).  A [n_pairs,k1,d], B [n_pairs,k2,d], evals1 [n_pairs,k1], evals2 [n_pairs,k2], c00 [n_pairs].
 * ---------------------------------------------------------------------------------------- */
size_t dm_fmap_solve_workspace_bytes(int n_pairs, int k1, int k2, int d);
int dm_fmap_solve(const double* A, const double* B, const double* evals1, const double* evals2,
                  const double* c00, double w_descr, double w_lap, int n_pairs, int k1, int k2, int d,
                  double* C /* [n_pairs, k2, k1] */, void* workspace, size_t workspace_bytes, dm_stream_t stream);
/* Outcome of the last dm_fmap_solve (or dm_match_pairs) that used `workspace`; synchronises the stream.
 *   out_h[0] != 0  a system was not positive definite even in float64 (rank-deficient descriptors with w_lap = 0,
 *                  w_descr = 0, ...): its row of C holds NaN.  The reference's L-BFGS returns a finite map there;
 *                  callers must treat this as an error (the Python layer raises).
 *   out_h[1]       systems the float32 factorisation + float64 refinement could not finish and the float64 kernel redid
 *   out_h[2]       reserved        out_h[3]  refinement steps taken in total */
int dm_fmap_solve_read_status(const void* workspace, int* out_h /* [4] */, dm_stream_t stream);

/* c00[p] = sign(Phi1[first vertex of pair p][0] * Phi2[first vertex][0]) * sqrt(area(mesh 2) / area(mesh 1)): the pinned
 * entry x0[0, 0] of FunctionalMapping.fit (pyFM/functional.py:654-658) that dm_fmap_solve takes, for a ragged batch. */
int dm_fmap_c00(const double* Phi1, int64_t ld1, const int64_t* off1, const double* Phi2, int64_t ld2,
                const int64_t* off2, const double* area1, const double* area2, int n_pairs, double* c00,
                dm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * FM -> point-to-point maps, all four index outputs from one pass over S = (Phi2 C) Phi1^T.
 * Replaces: FM_to_p2p  pyFM/spectral/convert.py:96-147 and the override functional_map.py:49-50.
 *   p2p_21     [total_n2]  argmin_j | Phi2[i] - (Phi1 C^T)[j] |     (convert.py:138-140)
 *   p2p_12     [total_n1]  argmin_i | Phi1[j] - (Phi2 C)[i] |       (convert.py:134-136)
 *   dense_21   [total_n2]  argmax_j (Phi2 C Phi1^T A1)_ij           (functional_map.py:49)
 *   dense_12   [total_n1]  argmax_i (Phi2 C Phi1^T A1)_ij           (functional_map.py:50)
 * any output pointer may be NULL.  C [n_pairs, k2, k1] float64; Phi1 [total_n1, ld1] uses columns :k1,
 * Phi2 [total_n2, ld2] uses columns :k2.
 * ---------------------------------------------------------------------------------------- */
size_t dm_fm_to_p2p_workspace_bytes(int n_pairs, int64_t total_n1, int64_t total_n2, int max_n1, int max_n2,
                                    int k1, int k2, int flags);
int dm_fm_to_p2p(const double* C, int k1, int k2,
                 const double* Phi1, int64_t ld1, const int64_t* off1, int64_t total_n1, int max_n1,
                 const double* Phi2, int64_t ld2, const int64_t* off2, int64_t total_n2, int max_n2,
                 const double* area1, int n_pairs,
                 void* p2p_21, void* p2p_12, void* dense_21, void* dense_12,
                 int flags, void* workspace, size_t workspace_bytes, dm_stream_t stream);

/* mapped_indicator = Phi2 C Phi1^T A1 of ONE pair, float64 [n2, n1] (convert.py:144); API parity only
 * (the index outputs above never materialise it). */
size_t dm_mapped_indicator_workspace_bytes(int n1, int k2);
int dm_mapped_indicator(const double* C, int k1, int k2, const double* Phi1, int64_t ld1, int n1,
                        const double* Phi2, int64_t ld2, int n2, const double* area1,
                        double* MI, int64_t ldMI, void* workspace, size_t workspace_bytes, dm_stream_t stream);
/* The same for a ragged batch in one call (the Hungarian slots of compute_surface_map for many pairs,
 * functional_map.py:57,66,78): the matrix of pair p occupies rows off2[p] .. off2[p + 1] of MI [total_n2, ldMI >= max_n1],
 * columns 0 .. n1_p - 1; per pair the arithmetic of dm_mapped_indicator. */
size_t dm_mapped_indicators_workspace_bytes(int64_t total_n1, int k2);
int dm_mapped_indicators(const double* C, int k1, int k2, const double* Phi1, int64_t ld1, const int64_t* off1,
                         int64_t total_n1, int max_n1, const double* Phi2, int64_t ld2, const int64_t* off2, int max_n2,
                         const double* area1, int n_pairs, double* MI, int64_t ldMI, void* workspace,
                         size_t workspace_bytes, dm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * p2p -> FM with the target mass:  C = Phi2[:, :k2]^T (area2 * Phi1[p2p_21, :k1])   float64
 * Replaces: p2p_to_FM  pyFM/spectral/convert.py:39-48 (A2 given; area2 == NULL means unit weights,
 * i.e. the Phi2^T Phi1[p] factor of the lstsq branch :51).  p2p_21 holds LOCAL vertex ids
 * (int32 | int64 per flags).  Deterministic split-K over the vertices.
 * ---------------------------------------------------------------------------------------- */
size_t dm_p2p_to_fm_workspace_bytes(int n_pairs, int max_n2, int k1, int k2);
int dm_p2p_to_fm(const void* p2p_21, const double* Phi1, int64_t ld1, const int64_t* off1,
                 const double* Phi2, int64_t ld2, const int64_t* off2, int max_n2,
                 const double* area2, int n_pairs, int k1, int k2,
                 double* C /* [n_pairs, k2, k1] */, int flags, void* workspace, size_t workspace_bytes,
                 dm_stream_t stream);

/* Batched symmetric positive definite solve  X[b] = G[b]^-1 B[b]  (G [n_batch, n, n], B / X [n_batch, n, m], float64,
 * contiguous): the normal equations (Phi2^T Phi2) C = Phi2^T Phi1[p] of the least-squares branch of p2p_to_FM
 * (convert.py:51, A2 = None; scipy.linalg.lstsq in the reference).  Cholesky inverse + one GEMM; dm_icp_read_status on
 * the same workspace reports a non-positive pivot. */
size_t dm_spd_solve_workspace_bytes(int n_batch, int n);
int dm_spd_solve(const double* G, const double* B, int n, int m, int n_batch, double* X, void* workspace,
                 size_t workspace_bytes, dm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * ZoomOut ladder (upstream pyFM semantics; pyFM/refine/zoomout.py:7-115, the shipped call at :40 is
 * broken): nit times  p = p2p_21(C);  C <- Phi2[:, :k2+s2]^T A2 Phi1[p, :k1+s1].
 * C0 [n_pairs, k2_0, k1_0] -> C_out [n_pairs, k2_0+nit*s2, k1_0+nit*s1]; p2p_out (optional) is the final
 * p2p_21 of the refined map.  Runs the whole ladder stream-ordered without host synchronisation.
 * ---------------------------------------------------------------------------------------- */
size_t dm_zoomout_workspace_bytes(int n_pairs, int64_t total_n1, int64_t total_n2, int max_n1, int max_n2,
                                  int k1_0, int k2_0, int nit, int step1, int step2, int flags);
int dm_zoomout(const double* C0, int k1_0, int k2_0, int nit, int step1, int step2,
               const double* Phi1, int64_t ld1, const int64_t* off1, int64_t total_n1, int max_n1,
               const double* Phi2, int64_t ld2, const int64_t* off2, int64_t total_n2, int max_n2,
               const double* area2, int n_pairs,
               double* C_out, void* p2p_out,
               int flags, void* workspace, size_t workspace_bytes, dm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Spectral ICP (pyFM/refine/icp.py:10-107, driven by FunctionalMapping.icp_refine functional.py:564-586):
 * nit times  p = p2p_21(C);  C = lstsq(Phi2[:, :k2], Phi1[p, :k1]);  C <- U I V^T of its SVD.
 * The least squares is solved through the (once-factorised) Gram matrix Phi2^T Phi2 and the orthogonal
 * polar factor through a one-sided Jacobi SVD, all float64.  p2p_out (optional) = p2p_21 of the result.
 * ---------------------------------------------------------------------------------------- */
size_t dm_icp_workspace_bytes(int n_pairs, int64_t total_n1, int64_t total_n2, int max_n1, int max_n2, int k1,
                              int k2, int flags);
int dm_icp(const double* C0, int k1, int k2, int nit,
           const double* Phi1, int64_t ld1, const int64_t* off1, int64_t total_n1, int max_n1,
           const double* Phi2, int64_t ld2, const int64_t* off2, int64_t total_n2, int max_n2,
           int n_pairs, double* C_out, void* p2p_out,
           int flags, void* workspace, size_t workspace_bytes, dm_stream_t stream);
/* out_h[0] != 0: the Gram matrix Phi2^T Phi2 of some pair was not positive definite (rank-deficient basis); the
 * refined maps of the call are then meaningless.  Synchronises the stream. */
int dm_icp_read_status(const void* workspace, int* out_h /* [4] */, dm_stream_t stream);

/* Batched dense float64 product  C[b] = A[b] B[b]^T  (A [n_batch, m, k], B [n_batch, n, k], C [n_batch, m, n], contiguous)
 * on the library's DMMA GEMM: the small k x k x d products around the fit (A A^T, B A^T, C (A A^T); torch.bmm in the
 * reference's energy_func_std, pyFM/optimize/base_functions.py:516-532). */
int dm_bmm_nt_f64(const double* A, const double* B, int n_batch, int m, int n, int k, double* C, dm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Dense-map energy terms of the fit and their gradient with respect to C, without materialising the n2 x n1 map
 * M = Phi2 C Phi1^T A1 (pyFM/optimize/base_functions.py: p2p :296-325, doubly_stochastic :327-361, entropy :363-372,
 * range01 :374-385, sumto1 :387-428; evaluated per L-BFGS callback by energy_func_std :480-639).
 *   energy [n_pairs, 5]      UNWEIGHTED terms in the order p2p, stochastic, ent, range01, sumto1 (0 where the weight is 0)
 *   grad   [n_pairs, k2, k1] d/dC of  sum_t w_t * term_t
 * float64 throughout; k1 <= 128.
 * ---------------------------------------------------------------------------------------- */
size_t dm_dense_energy_workspace_bytes(int n_pairs, int64_t total_n1, int64_t total_n2, int max_n1, int max_n2, int k1,
                                       int k2);
int dm_dense_energy(const double* C, int k1, int k2,
                    const double* Phi1, int64_t ld1, const int64_t* off1, int64_t total_n1, int max_n1,
                    const double* Phi2, int64_t ld2, const int64_t* off2, int64_t total_n2, int max_n2,
                    const double* area1, int n_pairs,
                    double w_p2p, double w_stochastic, double w_ent, double w_range01, double w_sumto1,
                    double* energy, double* grad, void* workspace, size_t workspace_bytes, dm_stream_t stream);
int dm_dense_energy_ex(const double* C, int k1, int k2,
                       const double* Phi1, int64_t ld1, const int64_t* off1, int64_t total_n1, int max_n1,
                       const double* Phi2, int64_t ld2, const int64_t* off2, int64_t total_n2, int max_n2,
                       const double* area1, int n_pairs,
                       double w_p2p, double w_stochastic, double w_ent, double w_range01, double w_sumto1,
                       double* energy, double* grad, int flags /* DM_FAST_LOSS */, void* workspace,
                       size_t workspace_bytes, dm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * The whole per-pair hot path of compute_surface_map (functional_map.py:44-50) for a ragged batch, in one call:
 *   feature NN both directions (knn_query on the unit features)          -> nn_p2p_21 [total_n2], nn_p2p_12 [total_n1]
 *   A = Phi1^T A1 F1, B = Phi2^T A2 F2 (tensor cores; the bf16 splits of F prepared for the NN pass are reused)
 *   C = closed-form minimiser, column 0 pinned to sign(Phi1[0,0] Phi2[0,0]) sqrt(area2/area1)  -> C [n_pairs, k, k]
 *   FM -> p2p: p2p_21, p2p_12 (kd-tree equivalent) and dense_21, dense_12 (argmax override); any may be NULL
 * F1 [total_n1, ldF1] / F2 float32; Phi [.., ld >= k] float64; evals1 / evals2 [n_pairs, k] contiguous float64.
 * ---------------------------------------------------------------------------------------- */
size_t dm_match_pairs_workspace_bytes(int n_pairs, int64_t total_n1, int64_t total_n2, int max_n1, int max_n2, int d,
                                      int k, int flags);
int dm_match_pairs(const float* F1, int64_t ldF1, const float* F2, int64_t ldF2,
                   const double* Phi1, int64_t ld1, const double* Phi2, int64_t ld2,
                   const double* area1, const double* area2, const double* evals1, const double* evals2,
                   const int64_t* off1, int64_t total_n1, int max_n1, const int64_t* off2, int64_t total_n2, int max_n2,
                   int n_pairs, int d, int k, double w_descr, double w_lap,
                   void* nn_p2p_21, void* nn_p2p_12, double* C, void* p2p_21, void* p2p_12, void* dense_21,
                   void* dense_12, int flags, void* workspace, size_t workspace_bytes, dm_stream_t stream);
/* status of the solve stage of the last dm_match_pairs on `workspace` (see dm_fmap_solve_read_status) */
int dm_match_pairs_read_status(const void* workspace, int* out_h /* [4] */, dm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Mesh bank: the same per-pair path for DATASET-shaped work -- every mesh takes part in many pairs (the evaluation
 * loop around compute_surface_map, densematcher/functional_map.py:9-81, over a test split; BASELINE config 5).
 * dm_bank_prepare does, ONCE PER MESH, everything of the path that depends on one mesh only: the bf16 splits and row
 * norms of its features (knn_query operands, pyFM/spectral/nn_utils.py:4-38), the three-way split of Phi[:, :k] (embedding
 * operands of FM_to_p2p, pyFM/spectral/convert.py:96-147) and the projection Phi^T A F (optimize/base_functions.py:526-532).
 * dm_match_bank_pairs then takes the pairs as two id lists and reads each pair's operands through its meshes' rows in the
 * bank: no per-pair copy of the meshes, no per-pair preparation or projection.  Results are bit-identical to
 * dm_match_pairs on the assembled batch.
 *   F [total_n, ldF] float32, Phi [total_n, ldPhi >= k] float64, area [total_n], evals [n_meshes, ld_evals >= k],
 *   bank_off [n_meshes + 1] (device): rows of mesh m = bank_off[m] .. bank_off[m + 1]
 *   ids1 / ids2 [n_pairs] (device int64): mesh 1 / mesh 2 of each pair; off1 / off2 [n_pairs + 1]: the packing of the
 *   OUTPUTS (cumulative sizes of the pairs' meshes, as in dm_match_pairs); `state` is opaque device memory of
 *   dm_bank_state_bytes, filled by dm_bank_prepare with the same (n_meshes, total_n, d, k).
 * ---------------------------------------------------------------------------------------- */
size_t dm_bank_state_bytes(int n_meshes, int64_t total_n, int d, int k);
size_t dm_bank_prepare_workspace_bytes(int n_meshes, int64_t total_n, int max_n, int d, int k);
/* prepares meshes mesh_lo .. mesh_hi - 1, whose rows are row_lo .. row_hi - 1 (= bank_off[mesh_lo] .. bank_off[mesh_hi],
 * which the host knows); the rest of `state` is left as it is, so a bank can be prepared piecewise -- a rank of a sharded
 * job prepares only the meshes its block of pairs touches.  (0, n_meshes, 0, total_n) prepares everything. */
int dm_bank_prepare(const float* F, int64_t ldF, const double* Phi, int64_t ldPhi, const double* area,
                    const int64_t* bank_off, int64_t total_n, int max_n, int n_meshes, int d, int k,
                    int mesh_lo, int mesh_hi, int64_t row_lo, int64_t row_hi,
                    void* state, size_t state_bytes, void* workspace, size_t workspace_bytes, dm_stream_t stream);
size_t dm_match_bank_pairs_workspace_bytes(int n_pairs, int64_t total_n1, int64_t total_n2, int max_n1, int max_n2,
                                           int d, int k, int flags);
int dm_match_bank_pairs(const void* state, size_t state_bytes, const float* F, int64_t ldF, const double* Phi,
                        int64_t ldPhi, const double* area, const double* evals, int64_t ld_evals,
                        const int64_t* bank_off, int64_t total_n, int n_meshes, int mesh_lo, int mesh_hi,
                        const int64_t* ids1, const int64_t* ids2, const int64_t* off1, int64_t total_n1, int max_n1, const int64_t* off2,
                        int64_t total_n2, int max_n2, int n_pairs, int d, int k, double w_descr, double w_lap,
                        void* nn_p2p_21, void* nn_p2p_12, double* C, void* p2p_21, void* p2p_12, void* dense_21,
                        void* dense_12, int flags, void* workspace, size_t workspace_bytes, dm_stream_t stream);
/* (mesh_lo, mesh_hi above: the range of meshes that dm_bank_prepare has prepared.)
 * out_h[0..3]: the solve stage's status words (dm_fmap_solve_read_status); out_h[4] != 0: a mesh id was outside the prepared
 * range or off1 / off2 do not match the sizes of the meshes in the bank (the results are then meaningless).  Synchronises. */
int dm_match_bank_pairs_read_status(const void* workspace, int n_pairs, int d, int k, int* out_h /* [5] */,
                                    dm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Nearest (partial) isometry  C[b] = U I V^T  of  X[b] = U S V^T  (the SVD step of icp.py:39-40), float64,
 * X / C [n_batch, rows, cols] contiguous, X != C.  Newton-Schulz iteration on batched GEMMs; matrices it cannot
 * orthonormalise to 1e-12 in its fixed step count (ill-conditioned or rank-deficient) are redone by a one-sided
 * Jacobi SVD.  flags & DM_POLAR_JACOBI forces the Jacobi path for all.
 * ---------------------------------------------------------------------------------------- */
enum { DM_POLAR_JACOBI = 1 << 9 };
size_t dm_polar_factor_workspace_bytes(int n_batch, int rows, int cols);
int dm_polar_factor(const double* X, int rows, int cols, int n_batch, double* C, int flags, void* workspace,
                    size_t workspace_bytes, dm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Rectangular linear sum assignment, float64, batched: the `linear_sum_assignment(MI * eta - 1000 (1 - eta),
 * maximize=True)` calls of compute_surface_map (densematcher/functional_map.py:57,66,78; the solver itself is
 * scipy.optimize's shortest-augmenting-path implementation, a third-party dependency of the reference).
 * Problem b is the dense row-major matrix cost + cost_off[b] with row_off[b+1]-row_off[b] rows and
 * col_off[b+1]-col_off[b] columns (leading dimension = its column count); cost_off / row_off / col_off are device
 * int64 arrays of n_batch(+1) entries.  tall_elems = total element count of the problems with more rows than
 * columns (they are solved on a transposed copy in the workspace; 0 if there are none).
 * Output: col_of_row[row_off[b] + i] = column assigned to row i, or -1 (only possible when rows > columns);
 * (rows with an assignment, their columns) is exactly scipy's (row_ind, col_ind) -- the same assignment, including
 * scipy's tie-breaking, not merely one of equal cost.  status[b] (device int32): 0 = solved, 1 = the matrix holds
 * NaN / -inf (+inf when maximising), 2 = infeasible; scipy raises ValueError for both.  One CTA per problem:
 * at most 8192 columns / rows, n_batch <= 65535.  flags: DM_I64_OUT.
 * ---------------------------------------------------------------------------------------- */
size_t dm_lap_workspace_bytes(int n_batch, int max_nr, int max_nc, int64_t tall_elems);
int dm_lap_solve(const double* cost, const int64_t* cost_off, const int64_t* row_off, const int64_t* col_off, int n_batch,
                 int max_nr, int max_nc, int64_t tall_elems, int maximize, void* col_of_row, int* status, int flags,
                 void* workspace, size_t workspace_bytes, dm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Barycentric precise map (Ezuz & Ben-Chen): project_pc_to_triangles with precompute_dmin=True
 * (densematcher/pyFM/spectral/projection_utils.py:16-115; callers convert.py:186-231, functional.py:221-251,
 * functional_map.py:62).  For every point of emb2 [total_n2, p] (pair b: rows off2[b] .. off2[b+1]-1) the triangle
 * of the p-dimensional mesh (emb1 [total_n1, p] rows off1[b].., faces [total_faces, 3] int32 vertex ids LOCAL to the
 * pair's mesh, rows face_off[b] .. face_off[b+1]-1) it projects onto:
 *   face_match[i]  face index local to the pair (int32, int64 with DM_I64_OUT)
 *   bary[i, 0..2]  barycentric coordinates of the projection on that face (float64)
 * Row i of the reference's sparse (n2, n1) map holds bary[i] in the columns faces[face_match[i]]
 * (barycentric_to_precise, projection_utils.py:380-417).
 * ---------------------------------------------------------------------------------------- */
size_t dm_precise_map_workspace_bytes(int n_pairs, int64_t total_n1, int max_n1, int64_t total_n2, int64_t total_faces);
int dm_precise_map(const double* emb1, int64_t ld1, const int64_t* off1, int64_t total_n1, int max_n1, const int32_t* faces,
                   const int64_t* face_off, int64_t total_faces, const double* emb2, int64_t ld2, const int64_t* off2,
                   int64_t total_n2, int max_n2, int n_pairs, int p, void* face_match, double* bary, int flags,
                   void* workspace, size_t workspace_bytes, dm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Eigenbasis provider: the k lowest eigenpairs of  W phi = lambda A phi  (W: cotangent stiffness, CSR, float64, symmetric
 * positive semi-definite; A = diag(mass) lumped vertex areas), evecs^T A evecs = I, eigenvalues ascending.
 * Replaces: laplacian_spectrum, pyFM/mesh/laplacian.py:143-182 (scipy eigsh(W, k, M=A, sigma=-0.01)) as called by
 * TriMesh.process (mesh/trimesh.py:498-531) and diffusion_net/geometry.py's operator cache.
 * Method: Chebyshev-filtered block subspace iteration on A^-1/2 W A^-1/2 (no factorisation); Gram products on the float64
 * tensor-core GEMM, dense Rayleigh-Ritz eigenproblems on the device (dm_sym_eig's kernel).  Stops when every wanted
 * residual |S x - theta x| <= tol * theta_k (tol <= 0: 1e-10) or after max_iter outer iterations; `degree` <= 0 selects
 * the default filter degree.  The call synchronises the stream once per outer iteration (8 bytes read back).
 * info_h (host, optional) = {iterations, converged, block width, dense-solver status}; residual_h (host, optional).
 * ---------------------------------------------------------------------------------------- */
size_t dm_lbo_eigs_workspace_bytes(int n, int64_t nnz, int k);
int dm_lbo_eigs(const int64_t* indptr, const int32_t* indices, const double* values, int64_t nnz, const double* mass,
                int n, int k, double tol, int max_iter, int degree, double* evals /* [k] */,
                double* evecs /* [n, ld_evecs] */, int64_t ld_evecs, int* info_h, double* residual_h, void* workspace,
                size_t workspace_bytes, dm_stream_t stream);

/* Batched dense symmetric eigen-decomposition, float64: A [n_batch, m, m] -> w [n_batch, m] ascending,
 * V [n_batch, m, m] with eigenvectors in the columns (numpy.linalg.eigh convention); m <= 512.  Householder
 * tridiagonalisation + implicit QL, one CTA per matrix. */
size_t dm_sym_eig_workspace_bytes(int n_batch, int m);
int dm_sym_eig(const double* A, int m, int n_batch, double* w, double* V, void* workspace, size_t workspace_bytes,
               dm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * DiffusionNet spectral transforms (diffusion_net/geometry.py:572-598, layers.py:56-67).
 * dm_from_basis: out[rows of mesh b, :c] = Phi_b[:, :k] coef[b]   (coef [n_meshes, k, c] float64).
 * dm_spectral_diffusion: out = Phi (exp(-evals t) * (Phi^T diag(mass) X))  -- to_basis on the tcgen05 projection
 * engine (dm_project_ex, same flags), the coefficient scaling, from_basis on the float64 tensor-core GEMM.
 * evals [n_meshes, k], time [c], X [total_n, ldX] float32, out [total_n, ld_out] float64.
 * ---------------------------------------------------------------------------------------- */
int dm_from_basis(const double* coef, const double* Phi, int64_t ldPhi, const int64_t* row_off, int max_n, int n_meshes,
                  int k, int c, double* out, int64_t ld_out, dm_stream_t stream);
size_t dm_spectral_diffusion_workspace_bytes(int n_meshes, int64_t total_n, int max_n, int k, int c);
int dm_spectral_diffusion(const double* Phi, int64_t ldPhi, const double* mass, const double* evals, const float* X,
                          int64_t ldX, const double* time, const int64_t* row_off, int64_t total_n, int max_n,
                          int n_meshes, int k, int c, double* out, int64_t ld_out, int flags, void* workspace,
                          size_t workspace_bytes, dm_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Farthest point sampling with Euclidean distances, batched over meshes: the vertex subsample ZoomOut takes when
 * `subsample` is an integer (pyFM/refine/zoomout.py:150-156, 199-206 -> TriMesh.extract_fps, mesh/trimesh.py:847-893 with
 * geodesic=False -> mesh/geometry.py:813-851).  verts [total_n, 3] float64 ragged-packed by row_off, first [n_meshes]
 * start vertex per mesh (the reference draws it at random), idx [n_meshes, size] int64 (mesh-local indices).
 * Same float64 rounding as numpy's norm and lowest-index argmax: identical samples for the same start vertex.
 * ---------------------------------------------------------------------------------------- */
size_t dm_fps_workspace_bytes(int64_t total_n);
int dm_fps(const double* verts, const int64_t* row_off, int64_t total_n, int n_meshes, const int64_t* first, int size,
           int64_t* idx, void* workspace, size_t workspace_bytes, dm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DM_B200_H */
