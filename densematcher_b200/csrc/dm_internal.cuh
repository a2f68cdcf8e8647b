// Internal declarations shared by the translation units of libdm_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <float.h>
#include "../../include/dm_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libdm_b200 is written for sm_100a (B200) only"
#endif

namespace dm {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
#define DM_FAIL(code, ...)      \
  do {                          \
    ::dm::set_error(__VA_ARGS__); \
    return (code);              \
  } while (0)
#define DM_CUDA_OK(expr)                                                                  \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) DM_FAIL(DM_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)
#define DM_LAUNCH_OK(what)                                                                   \
  do {                                                                                       \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess) DM_FAIL(DM_ERR_CUDA, "launch %s: %s", what, cudaGetErrorString(e__)); \
  } while (0)

// ---------------------------------------------------------------- workspace carving
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base(static_cast<char*>(b)) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
  size_t bytes() const { return (off + 255) & ~size_t(255); }
};

// ---------------------------------------------------------------- top-3 tracking
// Running (best, runner-up, third) of a candidate set: indices of the first two, value of the third.
// The third value bounds every candidate other than i1 / i2, which is what lets the float64
// re-evaluation look at two rows instead of the whole database (emit_result below).
struct Top3 {
  float m1;
  int i1;
  float m2;
  int i2;
  float m3;
};  // 20 B: a stride of 5 words is conflict-free in shared memory
constexpr int kNoIdx = 0x7fffffff;
__device__ __forceinline__ Top3 top3_init() {
  Top3 t;
  t.m1 = t.m2 = t.m3 = -INFINITY;
  t.i1 = t.i2 = kNoIdx;
  return t;
}
// candidates arrive in ascending index order inside one thread: strict '>' keeps the lowest index. Branch-free.
__device__ __forceinline__ void top3_push(Top3& t, float v, int idx) {
  const bool c1 = v > t.m1, c2 = v > t.m2;
  t.m3 = fmaxf(t.m3, fminf(t.m2, v));
  t.i2 = c1 ? t.i1 : (c2 ? idx : t.i2);
  t.m2 = fmaxf(t.m2, fminf(t.m1, v));
  t.i1 = c1 ? idx : t.i1;
  t.m1 = fmaxf(t.m1, v);
}
// union of two disjoint candidate sets; equal values -> lower index first
__device__ __forceinline__ void top3_merge(Top3& a, float bm1, int bi1, float bm2, int bi2, float bm3) {
  const bool bfirst = (bm1 > a.m1) || (bm1 == a.m1 && bi1 < a.i1);
  const float hm1 = bfirst ? bm1 : a.m1, hm2 = bfirst ? bm2 : a.m2, hm3 = bfirst ? bm3 : a.m3;
  const int hi1 = bfirst ? bi1 : a.i1, hi2 = bfirst ? bi2 : a.i2;
  const float lm1 = bfirst ? a.m1 : bm1, lm2 = bfirst ? a.m2 : bm2;
  const int li1 = bfirst ? a.i1 : bi1;
  const bool slo = (lm1 > hm2) || (lm1 == hm2 && li1 < hi2);
  a.m1 = hm1;
  a.i1 = hi1;
  a.m2 = slo ? lm1 : hm2;
  a.i2 = slo ? li1 : hi2;
  a.m3 = slo ? fmaxf(hm2, lm2) : fmaxf(hm3, lm1);
}
__device__ __forceinline__ void top3_merge(Top3& a, const Top3& b) { top3_merge(a, b.m1, b.i1, b.m2, b.i2, b.m3); }

// Cheaper tracking for short-lived partial scans (the 32-row column partials of the tensor-core kernel): best value +
// index and the runner-up VALUE only.  Converted to a Top3 whose runner-up index is unknown (kNoIdx) and whose third
// value is the conservative bound m2; top3_merge then yields exact (m1, i1, m2), a known i2 whenever the runner-up
// comes from another partial, and an upper bound m3 -- exactly what emit_result needs (an unknown i2 or a loose m3
// only turns a two-candidate re-evaluation into a full one).
struct Top2 {
  float m1;
  int i1;
  float m2;
};
__device__ __forceinline__ Top2 top2_init() {
  Top2 t;
  t.m1 = t.m2 = -INFINITY;
  t.i1 = kNoIdx;
  return t;
}
__device__ __forceinline__ void top2_push(Top2& t, float v, int idx) {  // ascending idx: strict '>' keeps the lowest
  const bool c1 = v > t.m1;
  t.m2 = fmaxf(t.m2, fminf(t.m1, v));
  t.i1 = c1 ? idx : t.i1;
  t.m1 = fmaxf(t.m1, v);
}
__device__ __forceinline__ Top3 top3_from(const Top2& t) {
  Top3 r;
  r.m1 = t.m1, r.i1 = t.i1, r.m2 = t.m2, r.i2 = kNoIdx, r.m3 = t.m2;
  return r;
}

__device__ __forceinline__ void store_index(void* out, int64_t pos, int v, bool i64) {
  if (i64)
    static_cast<int64_t*>(out)[pos] = v;
  else
    static_cast<int32_t*>(out)[pos] = v;
}
__device__ __forceinline__ int64_t load_index(const void* in, int64_t pos, bool i64) {
  return i64 ? static_cast<const int64_t*>(in)[pos] : static_cast<const int32_t*>(in)[pos];
}

// ---------------------------------------------------------------- vectorised row access for the split kernels
// A lane reads W consecutive elements of a matrix row with one 16-byte load (W = 4 floats / 2 doubles).
template <typename T>
struct RowVec;
template <>
struct RowVec<float> {
  static constexpr int W = 4;
  typedef float4 V;
  static __device__ __forceinline__ void unpack(const V& v, double (&o)[4]) { o[0] = v.x, o[1] = v.y, o[2] = v.z, o[3] = v.w; }
};
template <>
struct RowVec<double> {
  static constexpr int W = 2;
  typedef double2 V;
  static __device__ __forceinline__ void unpack(const V& v, double (&o)[2]) { o[0] = v.x, o[1] = v.y; }
};
// packs W bf16 values into one 8- or 4-byte store
template <int W>
__device__ __forceinline__ void store_bf16_vec(void* dst, const unsigned short (&b)[W]);
template <>
__device__ __forceinline__ void store_bf16_vec<4>(void* dst, const unsigned short (&b)[4]) {
  *reinterpret_cast<uint2*>(dst) = make_uint2(uint32_t(b[0]) | (uint32_t(b[1]) << 16), uint32_t(b[2]) | (uint32_t(b[3]) << 16));
}
template <>
__device__ __forceinline__ void store_bf16_vec<2>(void* dst, const unsigned short (&b)[2]) {
  *reinterpret_cast<uint32_t*>(dst) = uint32_t(b[0]) | (uint32_t(b[1]) << 16);
}

// ---------------------------------------------------------------- the NN problem, device view
constexpr int kMaxEpi = 2;

// One argmax epilogue after preparation: fp32 and fp64 copies of scale/bias over the reduced-over side,
// per-pair maxima for the error bound, and where the result goes.
struct EpiDev {
  const float* sf;   // scale, fp32            [rows of reduced-over side]
  const float* bf;   // bias, fp32
  const double* sd;  // scale, fp64 (recheck)
  const double* bd;  // bias, fp64
  const float* G;    // per pair: max_j |v_j| * |scale_j|
  const float* Bm;   // per pair: max_j |bias_j|
  void* out;
  int identity;      // scale == 1 and bias == 0 everywhere (plain dot-product argmax)
};

struct FlagEntry {  // one result that must be re-evaluated in float64
  int pair;
  int local;  // local index of the kept-side row inside the pair
  int epi;    // epilogue number; bit 8 set = column epilogue
  int mode;   // kFlagFull: scan every candidate; kFlagCand: only c1 / c2 can be the argmax
  int c1, c2; // local indices on the reduced-over side (kFlagCand)
  int pad0, pad1;
};
constexpr int kFlagFull = 0, kFlagCand = 1;

struct NNProblem {
  // fp32 operands for the fast score pass
  const float* Y;
  int64_t ldY;
  const float* X;
  int64_t ldX;
  // operands for the float64 re-evaluation (either the same fp32 data or float64 originals)
  const void* Y64;
  int64_t ldY64;
  int y64_is_double;
  const void* X64;
  int64_t ldX64;
  int x64_is_double;
  const int64_t* q_off;
  const int64_t* db_off;
  // first row of pair p inside the OPERAND arrays (the bf16 splits read through TMA and the originals read by the float64
  // re-evaluation).  Equal to q_off / db_off unless a side lives in a per-mesh bank (NNBankSide): then the operands of
  // pair p start at q_in[p] / db_in[p] of the bank while every per-row array of the batch (norms, scale / bias, results)
  // stays packed by q_off / db_off.  rows_q / rows_db: rows of the operand arrays (TMA extents).
  const int64_t* q_in;
  const int64_t* db_in;
  int64_t rows_q, rows_db;
  int64_t total_q, total_db;
  int max_q, max_db;
  int n_pairs, d;
  int d_fast;  // inner dimension of the fp32 operands (>= d; extra columns are zero on the database side)
  int kp;      // inner dimension of the bf16 split operands of the tensor-core engine (d rounded up to 64)
  int n_row, n_col;
  EpiDev row[kMaxEpi];
  EpiDev col[kMaxEpi];
  const float* norm_q;   // |y_i| (fp32, rounded up)
  const float* norm_db;  // |x_j|
  float eps;             // relative error bound of one score: |S~ - S| <= eps |y| |x|
  int probe_skip_epilogue;  // profiling (DM_NN_PROBE=1): tcgen05 kernel runs TMA + MMA only; results are garbage
  float col_trunc;       // extra relative error of the column scores (tcgen05 engine: 5 low mantissa bits carry the row)
  float row_trunc;       // the same for the row scores (dual-accumulator engine nn_tc2.cu: packed keys on both sides)
  int i64_out;
  int recheck_all;
  // scratch
  Top3* col_partial;  // [n_col][n_pairs * max_rt][max_db]
  int max_rt;         // row tiles per pair upper bound
  int rt_rows;        // rows per row tile (engine dependent)
  FlagEntry* flags;        // two-candidate entries grow from the front, full-scan entries from the back
  int64_t flag_cap;        // entries allocated
  unsigned int* counters;  // [0] = two-candidate entries, [1] rows flagged, [2] cols flagged, [3] full-scan entries
};

// Writes the fp32-grade argmax and, when the top-2 gap is inside the rounding-error bound of the score
// pass, queues the result for the float64 re-evaluation.  |s~ - s| <= eps |y||x||scale| + 4u(|s| + |bias|)
// for every candidate (u = 2^-24), hence the threshold below (with 2x slack on the u term).  When the
// third-best score is outside that window only the two leaders can be the float64 argmax.
__device__ __forceinline__ void emit_result(const NNProblem& P, const EpiDev& E, bool is_col, int epi, int pair,
                                            int64_t gpos, int local, float own_norm, const Top3& s) {
  const int idx = (s.i1 == kNoIdx) ? 0 : s.i1;
  store_index(E.out, gpos, idx, P.i64_out != 0);
  if (P.flags == nullptr) return;
  const float thr = 2.f * P.eps * own_norm * E.G[pair] + 9.6e-7f * (fabsf(s.m1) + E.Bm[pair]);
  const float tr = is_col ? P.col_trunc : P.row_trunc;  // per-value truncation error of packed-key partials
  const bool safe = (s.m1 - s.m2) > thr + tr * (fabsf(s.m1) + fabsf(s.m2));  // NaN -> not safe
  if (!safe || P.recheck_all) {
    const bool two = !P.recheck_all && (s.m1 - s.m3) > thr + tr * (fabsf(s.m1) + fabsf(s.m3)) && s.i2 != kNoIdx;
    const unsigned slot = atomicAdd(&P.counters[two ? 0 : 3], 1u);
    FlagEntry f;
    f.pair = pair;
    f.local = local;
    f.epi = epi | (is_col ? 256 : 0);
    f.mode = two ? kFlagCand : kFlagFull;
    f.c1 = s.i1;
    f.c2 = s.i2;
    f.pad0 = f.pad1 = 0;
    P.flags[two ? int64_t(slot) : P.flag_cap - 1 - int64_t(slot)] = f;
    atomicAdd(&P.counters[is_col ? 2 : 1], 1u);
  }
}

// engines (nn_ffma.cu, nn_tc.cu)
int nn_ffma_launch(const NNProblem& P, cudaStream_t st);
int nn_ffma_debug_scores(const float* Y, int64_t ldY, int nq, const float* X, int64_t ldX, int ndb, int d,
                         float* S, int64_t ldS, cudaStream_t st);
constexpr int kFfmaRowTile = 128;
// tcgen05 engine: bf16 split operands [rows, kp] prepared by nn_prep_side; dbgS != nullptr materialises the scores
int nn_tc_kp(int d);
int nn_tc_launch(const NNProblem& P, const void* Yh, const void* Yl, const void* Xh, const void* Xl, float* dbgS,
                 int64_t ldS, cudaStream_t st);
inline bool nn_use_tc(int flags) { return !(flags & DM_ENGINE_FFMA); }
// dual-accumulator tcgen05 engine for the four-output FM -> p2p pass with a short contraction (nn_tc2.cu)
bool nn_tc2_applicable(int n_row, int n_col, int kp);
int nn_tc2_launch(const NNProblem& P, const void* Yh, const void* Yl, const void* Xh, const void* Xl, cudaStream_t st);
// 2-D TMA descriptor over a row-major [rows, kp] bf16 matrix with boxes of [box_rows x 64], 128-byte swizzle (nn_tc.cu)
int tc_make_map_bf16(void* tensor_map /* CUtensorMap* */, const void* base, int64_t rows, int kp, int box_rows);

// One side of a search whose operands were prepared ONCE PER MESH (dm_bank_prepare) instead of once per pair: the bf16
// splits and row norms of the whole bank, and where each pair's mesh starts in it.  The float64 originals (NNRequest::X64 /
// Y64 or X / Y) are then bank arrays too; scale arrays of the epilogues over this side are indexed by bank rows.
struct NNBankSide {
  const int64_t* in;         // [n_pairs] first bank row of the pair's mesh (device)
  int64_t rows;              // rows of the bank arrays
  const void *hi, *lo, *lo2; // [rows, kp] bf16 (lo2: third split, may be null)
  const float* norm;         // [rows] row norms as nn_prep_side leaves them
};
// shared stages (nn_common.cu)
struct SideEpiSpec {  // how to derive one epilogue's scale/bias for the rows of one side
  int scale_mode, bias_mode;
  const double* scale;
  const double* bias;
  float* sf;
  float* bf;
  double* sd;
  double* bd;
  float* G;
  float* Bm;
};
// norms + derived scale/bias + per-pair maxima for one side; M is float or double
int nn_prep_side(const void* M, int is_double, int64_t ld, const int64_t* off, int n_pairs, int64_t total, int d,
                 float* norm_out, const SideEpiSpec* specs, int n_specs, void* hi, void* lo, void* lo2, int kp,
                 cudaStream_t st);
// the per-pair part of the preparation of a side served from a mesh bank: batch-packed copies of the bank's row norms,
// the epilogues' scale / bias arrays and the per-pair maxima (no operand is read or split)
int nn_bank_side_rows(const NNBankSide& B, const int64_t* off, int n_pairs, int max_n, float* norm_out,
                      const SideEpiSpec* specs, int n_specs, cudaStream_t st);
int nn_col_finalize(const NNProblem& P, cudaStream_t st);
int nn_recheck(const NNProblem& P, cudaStream_t st);

// where nn_run carves its scratch (public so that the functional-map stages can fill parts of it themselves)
struct NNLayout {
  unsigned int* counters;
  float *norm_q, *norm_db;
  struct Arr {
    float *sf, *bf;
    double *sd, *bd;
    float *G, *Bm;
  } row[kMaxEpi], col[kMaxEpi];
  Top3* col_partial;
  FlagEntry* flags;
  int64_t flag_cap;
  uint16_t *yh, *yl, *xh, *xl;  // bf16 split operands of the tensor-core engine [rows, kp]
  uint16_t *yl2, *xl2;          // third split (kFlagSplit3 only)
  size_t bytes;
};
struct NNRequest;
// Optional stage overrides of nn_run (the factored FM -> p2p path, embed_tc.cu): the query operand is produced by a
// tensor-core embedding kernel instead of nn_prep_side, and float64 values are filled in on demand before the
// re-evaluation.
struct NNHooks {
  void* ctx = nullptr;
  int (*prep_y)(void* ctx, const NNLayout& L, NNProblem& P, const NNRequest& R, cudaStream_t st) = nullptr;
  int (*prep_x)(void* ctx, const NNLayout& L, NNProblem& P, const NNRequest& R, cudaStream_t st) = nullptr;
  int (*after_prep)(void* ctx, const NNLayout& L, NNProblem& P, cudaStream_t st) = nullptr;
  int (*before_recheck)(void* ctx, const NNLayout& L, NNProblem& P, cudaStream_t st) = nullptr;
};

// full driver used by the extern "C" entry points and by the FM kernels
struct NNRequest {
  const float* Y;
  int64_t ldY;
  const float* X;
  int64_t ldX;
  const double* Y64;  // optional float64 originals (else nullptr -> use Y)
  int64_t ldY64;
  const double* X64;
  int64_t ldX64;
  const int64_t* q_off;
  const int64_t* db_off;
  int64_t total_q, total_db;
  int max_q, max_db, n_pairs, d;
  int d_fast;  // 0 -> d
  dm_nn_epi row[kMaxEpi];
  int n_row;
  dm_nn_epi col[kMaxEpi];
  int n_col;
  int flags;
  const NNHooks* hooks = nullptr;
  // ladders (ZoomOut / ICP) search the same static query matrix at every rung: its split may be prepared once with
  // y_prep_d >= d columns (columns beyond d meet zeros on the database side) and reused while skip_prep_y is set
  int y_prep_d = 0;
  int skip_prep_y = 0;
  // recorded on the stream once both operands are split (before the score pass): lets a caller fork work that only needs
  // the splits onto another stream (dm_match_pairs: the projections and the solve run beside the score pass)
  cudaEvent_t after_prep_event = nullptr;
  // sides served from a mesh bank (tensor-core engines only; epilogues over such a side: DM_SCALE_NONE / DM_SCALE_ARRAY
  // with DM_BIAS_NONE / DM_BIAS_ARRAY, the bias array being a placeholder that a hook fills)
  const NNBankSide* bank_q = nullptr;
  const NNBankSide* bank_db = nullptr;
};
size_t nn_workspace_bytes(int n_pairs, int64_t total_q, int64_t total_db, int max_q, int max_db, int d, int n_row,
                          int n_col, int flags);
// internal flag: also keep the third bf16 split (v = hi + lo + lo2) of both operands in the workspace, so that the
// projection engine can reuse the feature splits of the nearest-neighbour stage (dm_match_pairs)
constexpr int kFlagSplit3 = 1 << 30;
// internal flags of nn_workspace_bytes: the split operands of that side come from a mesh bank (none are carved)
constexpr int kFlagBankQ = 1 << 28, kFlagBankDb = 1 << 29;
struct NNSplits {  // where nn_run left the split operands ([rows, kp] bf16 each)
  const void *yh, *yl, *yl2, *xh, *xl, *xl2;
  int kp;
};
int nn_run(const NNRequest& R, void* ws, size_t ws_bytes, cudaStream_t st, NNSplits* splits = nullptr);

// tcgen05 projection engine (proj_tc.cu): out[b] = (a_scale A)^T (b_scale B[gather]) over the rows of batch b
bool proj_tc_supported(int k, int d);
size_t proj_tc_workspace_bytes(int n_batch, int64_t total_n, int max_n, int k, int d, bool b_presplit = false);
// b_presplit (optional): three [total_n, pad64(d)] bf16 arrays (hi, mid, lo) of the B operand prepared elsewhere
int proj_tc_run(const double* A, int64_t ldA, const double* a_scale, const float* Bf, const double* Bd, int64_t ldB,
                const double* b_scale, const void* b_gather, int b_gather_i64, const int64_t* b_gather_src_off,
                const int64_t* off, int64_t total_n, int max_n, int n_batch, int k, int d, double* out, void* ws,
                size_t ws_bytes, cudaStream_t st, const void* const* b_presplit = nullptr);

// FM -> p2p with tensor-core embeddings and on-demand float64 (embed_tc.cu); k1, k2 <= 128
bool f2p_factored_applicable(int k1, int k2, int flags);
size_t f2p_factored_workspace_bytes(int n_pairs, int64_t n1, int64_t n2, int max_n1, int max_n2, int k1, int k2, int flags);
int f2p_factored_run(const double* C, int k1, int k2, const double* Phi1, int64_t ld1, const int64_t* off1, int64_t total_n1,
                     int max_n1, const double* Phi2, int64_t ld2, const int64_t* off2, int64_t total_n2, int max_n2,
                     const double* area1, int n_pairs, void* p2p_21, void* p2p_12, void* dense_21, void* dense_12, int flags,
                     void* ws, cudaStream_t st, const NNBankSide* bank1 = nullptr, const NNBankSide* bank2 = nullptr);

// p2p_21 of a functional map (the ZoomOut / ICP inner conversion) with the database side Phi1 C^T embedded on the tensor
// cores and float64 rows filled on demand (embed_tc.cu); k1, k2 <= 128.  `scratch` holds the three-way split of Phi1
// (valid while *x_kp_state == pad64(k1)), the splits of C and the flags.
bool p2p21_factored_applicable(int k1, int k2, int flags);
size_t p2p21_factored_scratch_bytes(int n_pairs, int64_t total_n1, int k1m, int k2m);
int p2p21_factored_run(const double* C, int k1, int k2, const double* Phi1, int64_t ld1, const int64_t* off1, int64_t total_n1,
                       int max_n1, const double* Phi2, int64_t ld2, const int64_t* off2, int64_t total_n2, int max_n2,
                       int n_pairs, void* p2p_out, int flags, void* scratch, int scratch_k1m, int scratch_k2m, double* emb1,
                       int lde, void* nn_ws, size_t nn_ws_bytes, cudaStream_t st, int* x_kp_state, int* y_kp_state);

// incremental p2p -> FM of the ZoomOut ladder (zoomout_delta.cu): the full-width map M = Phi2^T A2 Phi1[p] is kept resident
// and corrected with the changed entries of the vertex map; every rung reads its leading block
bool p2p_to_fm_delta_applicable();
size_t p2p_to_fm_delta_ws(int n_pairs, int64_t total_n2);
int p2p_to_fm_delta_run(const void* p_new, const void* p_old, int i64, const double* Phi1, int64_t ld1, const int64_t* off1,
                        const double* Phi2, int64_t ld2, const int64_t* off2, int64_t total_n2, int max_n2, const double* area2,
                        int n_pairs, int K1, int K2, double* M, void* ws, cudaStream_t st);
int extract_block_run(const double* M, int K1, int K2, int k1, int k2, int n_pairs, double* C, cudaStream_t st);

int num_sms();
// true exactly once per (call site, device): function attributes such as the dynamic shared-memory limit are
// per-device state, and one process may drive several devices
struct OncePerDevice {
  bool done[64] = {};
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    const bool f = !done[dev];
    done[dev] = true;
    return f;
  }
};
int cvt_f64_f32(const double* src, int64_t lds, int64_t rows, int d, float* dst, int ldd, cudaStream_t st);

}  // namespace dm
