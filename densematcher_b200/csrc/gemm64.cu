// Batched / ragged float64 GEMM on the FP64 tensor cores (see gemm64.cuh for the operand model).
//
// 128 x 128 (or 128 x 64) output tile per 256-thread CTA; the eight warps form a 4 x 2 grid and each warp accumulates
// its 32 x 64 (32 x 32) sub-tile with mma.sync.m8n8k4.f64 (DMMA): 256 fused multiply-adds per instruction, so the
// kernel is bound by the FP64 pipe rather than by instruction issue or shared-memory bandwidth (the FFMA-style inner
// loop it replaces needed 64 DFMA + 8 LDS.128 per thread and k, and reached 25-40 % of the pipe).  K advances 8 at a
// time through a double-buffered shared-memory stage ([k][m] layout, row pitch 132 doubles); the next stage's global
// loads are issued into registers before the current stage is consumed and are not touched until after it.
#include "gemm64.cuh"

namespace dm {
namespace {

constexpr int TM = 128, TN = 128, TK = 8, NT = 256;
constexpr int LDT = TM + 4;  // padded row of a stage: row pitch = 4 (mod 16) doubles makes the 8x4 / 4x8 fragment loads
                             // of the FP64 MMA conflict-free (16 lanes -> 16 distinct 8-byte banks)

struct OpView {  // an operand resolved for one batch
  const double* d;
  const float* f;
  int64_t ld;
  const void* gather;
  int gather_i64;
  int64_t gbase;
  const double* kscale;  // already offset to this batch
};

__device__ __forceinline__ OpView resolve(const GemmOperand& o, int b) {
  OpView v;
  const int64_t row0 = o.off ? o.off[b] : 0;
  const int64_t base = (o.off ? 0 : int64_t(b) * o.batch_stride) + ((o.off && o.in) ? o.in[b] : row0) * o.ld + o.col0;
  v.d = o.d ? o.d + base : nullptr;
  v.f = o.f ? o.f + base : nullptr;
  v.ld = o.ld;
  v.gather = o.gather;
  v.gather_i64 = o.gather_i64;
  v.gbase = o.gather_off ? o.gather_off[b] : 0;
  const int64_t kbase = o.gather_off ? o.gather_off[b] : row0;
  v.kscale = o.kscale ? o.kscale + kbase : nullptr;
  return v;
}

__device__ __forceinline__ int ragged_k(const GemmOperand& o, int b) {
  if (o.trans != 1) return -1;
  if (o.gather_off) return o.gather_cnt ? o.gather_cnt[b] : int(o.gather_off[b + 1] - o.gather_off[b]);
  if (o.off) return int(o.off[b + 1] - o.off[b]);
  return -1;
}

// The four elements of a TM x TK operand tile one thread moves per stage.
//   TRANS = false: element(i, k) = Mat[i][k]   -> thread owns k = t % 8 of rows i = t / 8 + 32 e
//   TRANS = true : element(i, k) = Mat[k][i]   -> thread owns i = t % 128 of contraction rows k = t / 128 + 2 e
// Raw registers of one stage: nothing here consumes a loaded value, so the loads stay in flight while the previous
// stage is being multiplied (any arithmetic on them would stall the in-order warp on the memory latency).
template <int NE>
struct Staged {
  double d[NE];
  float f[NE];
  double s[NE];
  long long row[NE];  // gathered operands: source rows of the NEXT stage, loaded one stage ahead of the data (the row
                      // index -> row data dependency would otherwise sit inside one stage's latency budget)
};

// TW = tile width along the output index (128 or 64); a thread moves NE = TW * TK / NT elements per stage
template <bool TRANS, int TW>
struct Loader {
  static constexpr int NE = TW * TK / NT;
  const OpView& v;
  int i0, lim;  // first output index of the tile, number of valid output indices (M or N)
  int t;
  // TRANS operands of float64 whose rows start on 16-byte boundaries are moved two output indices at a time (one
  // 16-byte global load and one 16-byte shared store instead of two of each: the load/store unit is shared with the
  // DMMA operand loads): thread owns i = 2 (t % (TW / 2)), i + 1 of contraction rows k = t / (TW / 2) + (2 NT / TW) e
  bool vec2;
  __device__ __forceinline__ Loader(const OpView& view, int i0_, int lim_, int t_) : v(view), i0(i0_), lim(lim_), t(t_) {
    vec2 = v.d != nullptr && ((reinterpret_cast<uintptr_t>(v.d) | uintptr_t(v.ld * sizeof(double))) & 15) == 0 &&
           (i0 & 1) == 0;
  }

  // source rows of the stage starting at k0 (TRANS operands with a row gather only)
  __device__ __forceinline__ void prefetch_rows(int k0, int kend, Staged<NE>& r) const {
    if (!TRANS || !v.gather) return;
    if (vec2) {
#pragma unroll
      for (int e = 0; e < NE / 2; ++e) {
        const int k = k0 + (t / (TW / 2)) + (2 * NT / TW) * e;
        r.row[2 * e] = k < kend ? load_index(v.gather, v.gbase + k, v.gather_i64 != 0) : 0;
      }
    } else {
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        const int k = k0 + (t / TW) + (NT / TW) * e;
        r.row[e] = k < kend ? load_index(v.gather, v.gbase + k, v.gather_i64 != 0) : 0;
      }
    }
  }

  __device__ __forceinline__ void fetch(int k0, int kend, Staged<NE>& r) const {
    if (!TRANS && vec2) {  // thread owns k = 2 (t % 4), k + 1 of rows i = t / 4 + 64 e
#pragma unroll
      for (int e = 0; e < NE / 2; ++e) {
        const int k = k0 + 2 * (t & 3), i = i0 + (t >> 2) + 64 * e;
        r.d[2 * e] = r.d[2 * e + 1] = 0.0;
        if (i < lim && k < kend) {
          const double* src = v.d + int64_t(i) * v.ld + k;
          if (k + 1 < kend) {
            const double2 x = __ldg(reinterpret_cast<const double2*>(src));
            r.d[2 * e] = x.x, r.d[2 * e + 1] = x.y;
          } else {
            r.d[2 * e] = __ldg(src);
          }
        }
      }
      return;
    }
    if (TRANS && vec2) {
#pragma unroll
      for (int e = 0; e < NE / 2; ++e) {
        const int i = i0 + 2 * (t % (TW / 2));
        const int k = k0 + (t / (TW / 2)) + (2 * NT / TW) * e;
        r.d[2 * e] = r.d[2 * e + 1] = 0.0;
        r.s[2 * e] = 1.0;
        if (i < lim && k < kend) {
          const int64_t row = v.gather ? r.row[2 * e] : k;
          const double* src = v.d + row * v.ld + i;
          if (i + 1 < lim) {
            const double2 x = __ldg(reinterpret_cast<const double2*>(src));
            r.d[2 * e] = x.x, r.d[2 * e + 1] = x.y;
          } else {
            r.d[2 * e] = __ldg(src);
          }
          if (v.kscale) r.s[2 * e] = __ldg(v.kscale + k);
        }
      }
      return;
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      int i, k;
      if (!TRANS) {
        k = k0 + (t & 7);
        i = i0 + (t >> 3) + 32 * e;
      } else {
        i = i0 + (t % TW);
        k = k0 + (t / TW) + (NT / TW) * e;
      }
      r.d[e] = 0.0, r.f[e] = 0.f, r.s[e] = 1.0;
      if (i < lim && k < kend) {
        int64_t idx;
        if (!TRANS) {
          idx = int64_t(i) * v.ld + k;
        } else {
          const int64_t row = v.gather ? r.row[e] : k;
          idx = row * v.ld + i;
        }
        if (v.d)
          r.d[e] = __ldg(v.d + idx);
        else
          r.f[e] = __ldg(v.f + idx);
        if (TRANS && v.kscale) r.s[e] = __ldg(v.kscale + k);
      }
    }
  }
  __device__ __forceinline__ void stash(double* stage, const Staged<NE>& r) const {
    if (!TRANS && vec2) {
#pragma unroll
      for (int e = 0; e < NE / 2; ++e) {
        stage[(2 * (t & 3)) * LDT + (t >> 2) + 64 * e] = r.d[2 * e];
        stage[(2 * (t & 3) + 1) * LDT + (t >> 2) + 64 * e] = r.d[2 * e + 1];
      }
      return;
    }
    if (TRANS && vec2) {
      const bool scaled = v.kscale != nullptr;
#pragma unroll
      for (int e = 0; e < NE / 2; ++e) {
        double x0 = r.d[2 * e], x1 = r.d[2 * e + 1];
        if (scaled) x0 *= r.s[2 * e], x1 *= r.s[2 * e];
        *reinterpret_cast<double2*>(stage + ((t / (TW / 2)) + (2 * NT / TW) * e) * LDT + 2 * (t % (TW / 2))) = make_double2(x0, x1);
      }
      return;
    }
    const bool is_d = v.d != nullptr;
    const bool scaled = TRANS && v.kscale != nullptr;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      double x = is_d ? r.d[e] : double(r.f[e]);
      if (scaled) x *= r.s[e];
      if (!TRANS)
        stage[(t & 7) * LDT + (t >> 3) + 32 * e] = x;
      else
        stage[((t / TW) + (NT / TW) * e) * LDT + (t % TW)] = x;
    }
  }
};

// NB = accumulator columns per thread: 8 -> 128 x 128 tile, one CTA per SM (long contractions);
//                                      4 -> 128 x 64 tile, two CTAs per SM (short contractions are latency-bound:
//                                           the second CTA covers the other one's prologue and epilogue)
template <bool TA, bool TB, int NB>
__global__ void __launch_bounds__(NT, NB == 8 ? 1 : 2) gemm64_kernel(const GemmProblem P, int tiles_m, int tiles_n) {
  constexpr int TNW = 16 * NB;  // tile width in N
  int bid = blockIdx.x;
  const int tn = bid % tiles_n;
  bid /= tiles_n;
  const int tm = bid % tiles_m;
  bid /= tiles_m;
  const int ks = bid % P.ksplit;
  const int b = bid / P.ksplit;
  if (P.skip && P.skip[b]) return;

  const int M = (!TA && P.A.off) ? int(P.A.off[b + 1] - P.A.off[b]) : P.M;
  const int N = (!TB && P.B.off) ? int(P.B.off[b + 1] - P.B.off[b]) : P.N;
  int K = ragged_k(P.A, b);
  if (K < 0) K = ragged_k(P.B, b);
  if (K < 0) K = P.K;
  const int m0 = tm * TM, n0 = tn * TNW;
  if (m0 >= M || n0 >= N) return;
  int kbeg = 0, kend = K;
  if (P.ksplit > 1) {
    kbeg = min(K, ks * P.kchunk);
    kend = min(K, kbeg + P.kchunk);
  }

  __shared__ __align__(16) double As[2][TK * LDT];
  __shared__ __align__(16) double Bs[2][TK * LDT];
  const OpView A = resolve(P.A, b), B = resolve(P.B, b);
  const int t = threadIdx.x;
  const Loader<TA, TM> la(A, m0, M, t);
  const Loader<TB, TNW> lb(B, n0, N, t);
  // warp grid 4 (M) x 2 (N): warp tile 32 x WN, as 4 x NT8 MMA tiles of 8 x 8
  constexpr int WN = TNW / 2, NT8 = WN / 8;
  const int lane = t & 31, warp = t >> 5;
  const int wm = (warp & 3) * 32, wn = (warp >> 2) * WN;
  const int g = lane >> 2, t4 = lane & 3;  // fragment coordinates: A[row g][k t4], B[k t4][col g], C[row g][col 2 t4 + {0,1}]

  // 8 x 8 output blocks of this warp that lie inside the matrix (warp-uniform)
  const int na = max(0, min(4, (M - (m0 + wm) + 7) / 8)), nc = max(0, min(NT8, (N - (n0 + wn) + 7) / 8));

  double acc[4][NT8][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < NT8; ++c) acc[a][c][0] = acc[a][c][1] = 0.0;

  Staged<Loader<TA, TM>::NE> ra;
  Staged<Loader<TB, TNW>::NE> rb;
  if (kbeg < kend) {
    la.prefetch_rows(kbeg, kend, ra);
    lb.prefetch_rows(kbeg, kend, rb);
    la.fetch(kbeg, kend, ra);
    lb.fetch(kbeg, kend, rb);
    la.prefetch_rows(kbeg + TK, kend, ra);
    lb.prefetch_rows(kbeg + TK, kend, rb);
    la.stash(As[0], ra);
    lb.stash(Bs[0], rb);
  }
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += TK) {
    const bool more = k0 + TK < kend;
    if (more) {
      la.fetch(k0 + TK, kend, ra);
      lb.fetch(k0 + TK, kend, rb);
      la.prefetch_rows(k0 + 2 * TK, kend, ra);
      lb.prefetch_rows(k0 + 2 * TK, kend, rb);
    }
    const double* as = As[buf] + wm + g;
    const double* bs = Bs[buf] + wn + g;
    if (na == 4 && nc == NT8) {
#pragma unroll
      for (int k4 = 0; k4 < TK; k4 += 4) {
        double av[4], bv[NT8];
#pragma unroll
        for (int a = 0; a < 4; ++a) av[a] = as[(k4 + t4) * LDT + 8 * a];
#pragma unroll
        for (int c = 0; c < NT8; ++c) bv[c] = bs[(k4 + t4) * LDT + 8 * c];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < NT8; ++c)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(acc[a][c][0]), "+d"(acc[a][c][1])
                         : "d"(av[a]), "d"(bv[c]));
      }
    } else if (na > 0 && nc > 0) {
      // ragged edge of the output (the ZoomOut ladder's (k + 1) x (k + 1) maps: k + 1 is rarely a multiple of the tile):
      // only the 8 x 8 blocks that hold output issue MMAs (warp-uniform), so the FP64 tensor pipe is not spent on padding
#pragma unroll
      for (int k4 = 0; k4 < TK; k4 += 4) {
        double av[4], bv[NT8];
#pragma unroll
        for (int a = 0; a < 4; ++a) av[a] = as[(k4 + t4) * LDT + 8 * a];
#pragma unroll
        for (int c = 0; c < NT8; ++c) bv[c] = bs[(k4 + t4) * LDT + 8 * c];
#pragma unroll
        for (int a = 0; a < 4; ++a)
          if (a < na)
#pragma unroll
            for (int c = 0; c < NT8; ++c)
              if (c < nc)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                             : "+d"(acc[a][c][0]), "+d"(acc[a][c][1])
                             : "d"(av[a]), "d"(bv[c]));
      }
    }
    if (more) {
      la.stash(As[buf ^ 1], ra);
      lb.stash(Bs[buf ^ 1], rb);
    }
    __syncthreads();
    buf ^= 1;
  }

  double* C = P.C + int64_t(ks) * P.split_stride + (P.c_off ? P.c_off[b] * P.ldc : int64_t(b) * P.c_batch_stride);
  const double* cs = P.c_colscale ? P.c_colscale + (P.c_colscale_off ? P.c_colscale_off[b] : 0) : nullptr;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int m = m0 + wm + 8 * a + g;
    if (m >= M) continue;
#pragma unroll
    for (int c = 0; c < NT8; ++c) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = n0 + wn + 8 * c + 2 * t4 + h;
        if (n >= N) continue;
        double v = P.alpha * acc[a][c][h];
        if (cs) v *= cs[n];
        C[int64_t(m) * P.ldc + n] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// "Both transposed" shape: C[b] (M x N) = sum_k A[b][k][m] * B[b][gather(k)][n] * kscale[k] -- the p2p -> FM product
// Phi2^T A2 Phi1[p] of every ZoomOut / ICP rung and the Gram products (convert.py:39-48).  The contraction runs over
// the VERTICES, both operands are row-major [vertex, column], so a stage of TK vertices is a set of contiguous row
// segments: they go global -> shared with 16-byte cp.async through a KSTAGES-deep ring (no register staging, the gather
// indices of a stage are fetched KSTAGES stages ahead), which is what the register-staged generic kernel above lacked
// on this shape: it restarted a dependent (index -> row) load chain every 8 vertices and ran at 8 TFLOP/s.
// Same 128 x 64 tile, warp grid and edge-block skipping as gemm64_kernel<.., .., 4>; the vertex scale is applied to the
// A fragments as they are read.
constexpr int KSTAGES = 4;
constexpr int TT_TNW = 64;
// TTK vertices per stage: 8 gave 32 DMMAs per warp between two block barriers (barrier stalls were the top stall reason in
// the first capture, 42 % DMMA pipe); 16 halves the barrier count at 103 KB of shared memory (still two blocks per SM)
constexpr size_t tt_smem(int ttk) {
  return size_t(KSTAGES) * ttk * (LDT + (TT_TNW + 4)) * sizeof(double) + size_t(KSTAGES) * ttk * sizeof(double);
}

__device__ __forceinline__ void cp_async16_zfill(double* dst, const double* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(uint32_t(__cvta_generic_to_shared(dst))), "l"(src),
               "r"(src_bytes)
               : "memory");
}

template <int TTK>
__global__ void __launch_bounds__(NT, 2) gemm64_tt_kernel(const GemmProblem P, int tiles_m, int tiles_n) {
  constexpr int TNW = TT_TNW, LDB = TNW + 4;
  constexpr int AH = TTK / 4, BH = TTK / 8;  // 16-byte chunks per thread and stage: A (TTK x 64 chunks), B (TTK x 32)
  int bid = blockIdx.x;
  const int tn = bid % tiles_n;
  bid /= tiles_n;
  const int tm = bid % tiles_m;
  bid /= tiles_m;
  const int ks = bid % P.ksplit;
  const int b = bid / P.ksplit;
  if (P.skip && P.skip[b]) return;
  const int M = P.M, N = P.N;
  int K = ragged_k(P.A, b);
  if (K < 0) K = ragged_k(P.B, b);
  if (K < 0) K = P.K;
  const int m0 = tm * TM, n0 = tn * TNW;
  if (m0 >= M || n0 >= N) return;
  int kbeg = 0, kend = K;
  if (P.ksplit > 1) {
    kbeg = min(K, ks * P.kchunk);
    kend = min(K, kbeg + P.kchunk);
  }
  extern __shared__ __align__(16) double tsm[];
  double* As = tsm;                                  // [KSTAGES][TTK][LDT]
  double* Bs = As + KSTAGES * TTK * LDT;             // [KSTAGES][TTK][LDB]
  double* Ss = Bs + KSTAGES * TTK * LDB;             // [KSTAGES][TTK] vertex scale
  const OpView A = resolve(P.A, b), B = resolve(P.B, b);
  const int t = threadIdx.x;
  const int a_row = t >> 6, a_ch = t & 63;           // A rows a_row + 4 h
  const int b_row = t >> 5, b_ch = t & 31;           // B rows b_row + 8 h
  const int n_stage = (kend - kbeg + TTK - 1) / TTK;
  auto a_src_row = [&](int k) -> int64_t { return A.gather ? load_index(A.gather, A.gbase + k, A.gather_i64 != 0) : k; };
  auto b_src_row = [&](int k) -> int64_t { return B.gather ? load_index(B.gather, B.gbase + k, B.gather_i64 != 0) : k; };
  // gather rows of the stage that will be ISSUED next (fetched one issue ahead: KSTAGES stages before use)
  int64_t ra[AH], rb[BH];
  auto load_rows = [&](int stage) {
    const int k = kbeg + stage * TTK;
#pragma unroll
    for (int h = 0; h < AH; ++h) ra[h] = (k + a_row + 4 * h < kend) ? a_src_row(k + a_row + 4 * h) : 0;
#pragma unroll
    for (int h = 0; h < BH; ++h) rb[h] = (k + b_row + 8 * h < kend) ? b_src_row(k + b_row + 8 * h) : 0;
  };
  auto issue = [&](int stage) {
    const int k = kbeg + stage * TTK, buf = stage % KSTAGES;
    double* as = As + buf * TTK * LDT;
    double* bs = Bs + buf * TTK * LDB;
#pragma unroll
    for (int h = 0; h < AH; ++h) {
      const int kk = a_row + 4 * h, i = m0 + 2 * a_ch;
      const int valid = (k + kk < kend) ? max(0, min(2, M - i)) : 0;
      const double* src = A.d + ra[h] * A.ld + min(i, max(M - 1, 0));
      cp_async16_zfill(as + kk * LDT + 2 * a_ch, valid > 0 ? src : A.d, 8 * valid);
    }
#pragma unroll
    for (int h = 0; h < BH; ++h) {
      const int kk = b_row + 8 * h, i = n0 + 2 * b_ch;
      const int valid = (k + kk < kend) ? max(0, min(2, N - i)) : 0;
      const double* src = B.d + rb[h] * B.ld + min(i, max(N - 1, 0));
      cp_async16_zfill(bs + kk * LDB + 2 * b_ch, valid > 0 ? src : B.d, 8 * valid);
    }
    if (t < TTK) {
      const double* ksc = A.kscale ? A.kscale : B.kscale;
      Ss[buf * TTK + t] = (k + t < kend) ? (ksc ? ksc[k + t] : 1.0) : 0.0;
    }
  };
  // warp grid 4 (M) x 2 (N): warp tile 32 x 32
  constexpr int WN = TNW / 2, NT8 = WN / 8;
  const int lane = t & 31, warp = t >> 5;
  const int wm = (warp & 3) * 32, wn = (warp >> 2) * WN;
  const int g = lane >> 2, t4 = lane & 3;
  const int na = max(0, min(4, (M - (m0 + wm) + 7) / 8)), nc = max(0, min(NT8, (N - (n0 + wn) + 7) / 8));
  double acc[4][NT8][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < NT8; ++c) acc[a][c][0] = acc[a][c][1] = 0.0;

  load_rows(0);
  for (int s0 = 0; s0 < KSTAGES - 1; ++s0) {
    if (s0 < n_stage) {
      issue(s0);
      load_rows(s0 + 1);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int st = 0; st < n_stage; ++st) {
    asm volatile("cp.async.wait_group %0;" ::"n"(KSTAGES - 2) : "memory");
    __syncthreads();  // stage st has landed for every thread, and everyone is done with the buffer reused below
    const int nxt = st + KSTAGES - 1;
    if (nxt < n_stage) {
      issue(nxt);
      load_rows(nxt + 1);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int buf = st % KSTAGES;
    const double* as = As + buf * TTK * LDT + wm + g;
    const double* bs = Bs + buf * TTK * LDB + wn + g;
    const double* ss = Ss + buf * TTK;
    if (na > 0 && nc > 0) {
#pragma unroll
      for (int k4 = 0; k4 < TTK; k4 += 4) {
        const double sc = ss[k4 + t4];
        double av[4], bv[NT8];
#pragma unroll
        for (int a = 0; a < 4; ++a) av[a] = as[(k4 + t4) * LDT + 8 * a] * sc;
#pragma unroll
        for (int c = 0; c < NT8; ++c) bv[c] = bs[(k4 + t4) * LDB + 8 * c];
#pragma unroll
        for (int a = 0; a < 4; ++a)
          if (a < na)
#pragma unroll
            for (int c = 0; c < NT8; ++c)
              if (c < nc)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                             : "+d"(acc[a][c][0]), "+d"(acc[a][c][1])
                             : "d"(av[a]), "d"(bv[c]));
      }
    }
  }
  double* C = P.C + int64_t(ks) * P.split_stride + (P.c_off ? P.c_off[b] * P.ldc : int64_t(b) * P.c_batch_stride);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int m = m0 + wm + 8 * a + g;
    if (m >= M) continue;
#pragma unroll
    for (int c = 0; c < NT8; ++c) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = n0 + wn + 8 * c + 2 * t4 + h;
        if (n >= N) continue;
        double v = P.alpha * acc[a][c][h];
        if (P.c_add) v += P.c_add[int64_t(b) * P.c_add_batch_stride + int64_t(m) * P.c_add_ld + n];
        C[int64_t(m) * P.ldc + n] = v;
      }
    }
  }
}

// the cp.async path needs float64 operands whose rows and column windows start on 16-byte boundaries
bool tt_shape(const GemmProblem& P) {
  static const bool off = [] { const char* e = getenv("DM_GEMM_NO_TT"); return e && e[0] == '1'; }();
  if (off || P.A.trans != 1 || P.B.trans != 1 || !P.A.d || !P.B.d || P.c_colscale) return false;
  if (P.c_add && P.ksplit > 1) return false;
  if (P.A.kscale && P.B.kscale) return false;
  auto ok = [](const GemmOperand& o) {
    return ((reinterpret_cast<uintptr_t>(o.d) | uintptr_t(o.ld * sizeof(double)) | uintptr_t(o.col0 * sizeof(double))) & 15) == 0;
  };
  return ok(P.A) && ok(P.B);
}

// ---------------------------------------------------------------------------------------------------------------
// "Embedding" shape: out[b] (M x N) = A[b] (M x K, rows contiguous in k) . B[b], with a SMALL B (K, N <= 104: the
// functional map C or its transpose) and a tall A (a mesh's eigenbasis).  The generic kernel above restarts its
// k pipeline for every 128 x 64 output tile and runs this shape at ~13 TFLOP/s; here B stays resident in shared
// memory, a CTA walks over several 64-row tiles of A (cp.async double buffer: the next tile streams in while the
// current one is multiplied) and every warp owns 8 rows x all columns, so one A fragment feeds up to 13 DMMAs.
// Measured at N = 2000, K = N = 100, 128 pairs: 263 us = 19.4 TFLOP/s (the DMMA pipe alone peaks at 37.1,
// scripts/micro/dmma_peak.cu).  Tried and not kept: a 4 x 2 warp grid (B fragments shared by two row fragments: 282 us),
// a lane-major B layout read with LDS.128 (412 us: bank conflicts), eight tiles per CTA (393 us: wave quantisation).
constexpr int ET = 64;          // rows of A per tile
constexpr int EKP = 104;        // padded K and N
constexpr int ELD = EKP + 4;    // pitch == 12 (mod 16): conflict-free 8 x 4 / 4 x 8 fragment loads
constexpr int ETPC = 4;         // tiles per CTA
constexpr int ENC8 = EKP / 8;   // 8-column accumulator tiles per warp
constexpr size_t kEmbedSmem = size_t(EKP + 2 * ET) * ELD * sizeof(double);

__device__ __forceinline__ void cp_async16(double* dst, const double* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(uint32_t(__cvta_generic_to_shared(dst))), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(double* dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(uint32_t(__cvta_generic_to_shared(dst))), "l"(src) : "memory");
}

template <bool TB, int NC8>  // NC8: 8-column accumulator tiles per warp (N <= 8 NC8)
__global__ void __launch_bounds__(NT, 1) embed64_kernel(const GemmProblem P, int chunks) {
  extern __shared__ __align__(16) double esm[];
  double* Bs = esm;                 // [EKP][ELD]  B[k][n]
  double* As = esm + EKP * ELD;     // [2][ET][ELD]
  const int b = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  if (P.skip && P.skip[b]) return;
  const int M = P.A.off ? int(P.A.off[b + 1] - P.A.off[b]) : P.M;
  const int N = P.N, K = P.K;
  const int row_begin = chunk * ET * ETPC;
  if (row_begin >= M) return;
  const int n_tiles = min(ETPC, (M - row_begin + ET - 1) / ET);
  const OpView A = resolve(P.A, b), B = resolve(P.B, b);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int kp = (K + 3) & ~3;
  constexpr int nc8 = NC8;

  // four threads per row of the tile; 16-byte copies when every row of A starts on a 16-byte boundary (half as many
  // requests through the load/store unit, which the DMMA operand loads share: 310 -> 282 us), 8-byte copies otherwise
  const bool vec2 = ((reinterpret_cast<uintptr_t>(A.d) | uintptr_t(A.ld * sizeof(double))) & 15) == 0 && (K & 1) == 0;
  auto prefetch = [&](int tile, int buf) {
    const int r = t >> 2, row = row_begin + tile * ET + r;
    double* dst = As + buf * ET * ELD + r * ELD;
    const double* src = A.d + int64_t(row) * A.ld;
    const bool row_ok = row < M;
    if (vec2) {
      for (int k = 2 * (t & 3); k < kp; k += 8) {
        if (row_ok && k < K)
          cp_async16(dst + k, src + k);
        else
          dst[k] = dst[k + 1] = 0.0;
      }
    } else {
      for (int k = t & 3; k < kp; k += 4) {
        if (row_ok && k < K)
          cp_async8(dst + k, src + k);
        else
          dst[k] = 0.0;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch(0, 0);
  // B[k][n]: TB: element(n, k) = Mat[k][n]; else element(n, k) = Mat[n][k]   (zero padding up to kp x 8 nc8)
  for (int e = t; e < kp * 8 * nc8; e += NT) {
    int k, n;
    if (TB) {
      k = e / (8 * nc8), n = e - k * (8 * nc8);
    } else {
      n = e / kp, k = e - n * kp;
    }
    double v = 0.0;
    if (k < K && n < N) v = TB ? B.d[int64_t(k) * B.ld + n] : B.d[int64_t(n) * B.ld + k];
    Bs[k * ELD + n] = v;
  }
  double* C = P.C + (P.c_off ? P.c_off[b] * P.ldc : int64_t(b) * P.c_batch_stride);
  const double* cs = P.c_colscale ? P.c_colscale + (P.c_colscale_off ? P.c_colscale_off[b] : 0) : nullptr;
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int buf = tile & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // this tile has landed, and every warp is done with the other buffer (the previous tile)
    if (tile + 1 < n_tiles) prefetch(tile + 1, buf ^ 1);  // streams in during the multiplication below
    const double* as = As + buf * ET * ELD + (8 * warp + g) * ELD + t4;
    const double* bs = Bs + t4 * ELD + g;
    double acc[NC8][2];
#pragma unroll
    for (int c = 0; c < NC8; ++c) acc[c][0] = acc[c][1] = 0.0;
    // the fragments of step k4 + 4 are loaded into a second register set while the DMMAs of step k4 issue: a DMMA
    // never waits for a shared-memory load that was issued right before it
    double fa[2], fb[2][NC8];
    auto load_frags = [&](int k4, int set) {
      fa[set] = as[k4];
#pragma unroll
      for (int c = 0; c < NC8; ++c) fb[set][c] = bs[k4 * ELD + 8 * c];
    };
    auto mma_step = [&](int set) {
#pragma unroll
      for (int c = 0; c < NC8; ++c) {
        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
            : "+d"(acc[c][0]), "+d"(acc[c][1])
            : "d"(fa[set]), "d"(fb[set][c]));
      }
    };
    load_frags(0, 0);
    for (int k4 = 0; k4 < kp; k4 += 8) {
      if (k4 + 4 < kp) load_frags(k4 + 4, 1);
      mma_step(0);
      if (k4 + 4 < kp) {
        if (k4 + 8 < kp) load_frags(k4 + 8, 0);
        mma_step(1);
      }
    }
    const int m = row_begin + tile * ET + 8 * warp + g;
    if (m < M) {
      double* crow = C + int64_t(m) * P.ldc;
      const bool st2 = ((reinterpret_cast<uintptr_t>(C) | uintptr_t(P.ldc * sizeof(double))) & 15) == 0;
#pragma unroll
      for (int c = 0; c < NC8; ++c) {
        const int n = 8 * c + 2 * t4;
        double v0 = P.alpha * acc[c][0], v1 = P.alpha * acc[c][1];
        if (cs) {
          if (n < N) v0 *= cs[n];
          if (n + 1 < N) v1 *= cs[n + 1];
        }
        if (st2 && n + 1 < N) {
          *reinterpret_cast<double2*>(crow + n) = make_double2(v0, v1);
        } else {
          if (n < N) crow[n] = v0;
          if (n + 1 < N) crow[n + 1] = v1;
        }
      }
    }
  }
}

bool embed_shape(const GemmProblem& P) {
  return !P.A.trans && P.A.d && !P.A.gather && P.B.d && !P.B.off && !P.B.gather && !P.B.kscale && P.ksplit <= 1 && P.K > 0 &&
         P.K <= EKP && P.N > 0 && P.N <= EKP && P.maxM >= 4 * ET;
}

__global__ void __launch_bounds__(256)
    sum_partials_kernel(const double* __restrict__ part, int n_split, int64_t stride, int64_t n, double* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int k = 0; k < n_split; ++k) s += part[int64_t(k) * stride + i];
  out[i] = s;
}

}  // namespace

int gemm64_launch(const GemmProblem& P, cudaStream_t st) {
  if (P.n_batch <= 0 || P.maxM <= 0 || P.maxN <= 0) return DM_OK;
  if (P.c_add && !tt_shape(P)) DM_FAIL(DM_ERR_UNSUPPORTED, "gemm64: an addend needs the both-transposed cp.async shape");
  if (embed_shape(P)) {
    const int chunks = (P.maxM + ET * ETPC - 1) / (ET * ETPC);
    const int64_t nb = int64_t(P.n_batch) * chunks;
    if (nb > 0x7fffffffLL) DM_FAIL(DM_ERR_BADARG, "gemm64: grid too large");
#define DM_EMBED(TB_, NC8_)                                                                                            \
  do {                                                                                                                \
    static OncePerDevice once;                                                                                        \
    if (once.first())                                                                                                 \
      DM_CUDA_OK(cudaFuncSetAttribute(embed64_kernel<TB_, NC8_>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                      int(kEmbedSmem)));                                                              \
    embed64_kernel<TB_, NC8_><<<unsigned(nb), NT, kEmbedSmem, st>>>(P, chunks);                                        \
  } while (0)
    const int nc8 = (P.N + 7) / 8;
    if (P.B.trans) {
      if (nc8 <= 4) DM_EMBED(true, 4); else if (nc8 <= 7) DM_EMBED(true, 7); else if (nc8 <= 10) DM_EMBED(true, 10); else DM_EMBED(true, 13);
    } else {
      if (nc8 <= 4) DM_EMBED(false, 4); else if (nc8 <= 7) DM_EMBED(false, 7); else if (nc8 <= 10) DM_EMBED(false, 10); else DM_EMBED(false, 13);
    }
#undef DM_EMBED
    DM_LAUNCH_OK("embed64_kernel");
    return DM_OK;
  }
  if (tt_shape(P)) {
    const int tiles_m = (P.maxM + TM - 1) / TM, tiles_n = (P.maxN + TT_TNW - 1) / TT_TNW;
    GemmProblem Q = P;
    if (Q.ksplit < 1) Q.ksplit = 1;
    const int64_t nblk = int64_t(P.n_batch) * Q.ksplit * tiles_m * tiles_n;
    if (nblk > 0x7fffffffLL) DM_FAIL(DM_ERR_BADARG, "gemm64: grid too large");
    static const int ttk = [] { const char* e = getenv("DM_TT_TK"); return e ? atoi(e) : 16; }();
    static OncePerDevice once;
    if (once.first()) {
      DM_CUDA_OK(cudaFuncSetAttribute(gemm64_tt_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(tt_smem(8))));
      DM_CUDA_OK(cudaFuncSetAttribute(gemm64_tt_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(tt_smem(16))));
    }
    if (ttk == 8)
      gemm64_tt_kernel<8><<<unsigned(nblk), NT, tt_smem(8), st>>>(Q, tiles_m, tiles_n);
    else
      gemm64_tt_kernel<16><<<unsigned(nblk), NT, tt_smem(16), st>>>(Q, tiles_m, tiles_n);
    DM_LAUNCH_OK("gemm64_tt_kernel");
    return DM_OK;
  }
  // narrow tiles when the contraction is short (latency-bound) or when they waste less of the N range
  const int waste128 = (P.maxN + 127) / 128 * 128 - P.maxN, waste64 = (P.maxN + 63) / 64 * 64 - P.maxN;
  const int kper = P.ksplit > 1 ? P.kchunk : P.maxK;
  static const int force_wide = [] { const char* e = getenv("DM_GEMM_WIDE"); return e ? atoi(e) : 0; }();
  bool narrow = kper <= 512 || waste64 < waste128;
  if (force_wide == 1 && P.maxN > 64) narrow = false;
  if (force_wide == 2 && P.maxN > 96) narrow = waste64 + 32 < waste128;
  const int tnw = narrow ? 64 : 128;
  const int tiles_m = (P.maxM + TM - 1) / TM, tiles_n = (P.maxN + tnw - 1) / tnw;
  const int64_t nblk = int64_t(P.n_batch) * (P.ksplit > 0 ? P.ksplit : 1) * tiles_m * tiles_n;
  if (nblk > 0x7fffffffLL) DM_FAIL(DM_ERR_BADARG, "gemm64: grid too large");
  GemmProblem Q = P;
  if (Q.ksplit < 1) Q.ksplit = 1;
  const unsigned grid = unsigned(nblk);
#define DM_GEMM(TA_, TB_)                                                          \
  do {                                                                             \
    if (narrow)                                                                    \
      gemm64_kernel<TA_, TB_, 4><<<grid, NT, 0, st>>>(Q, tiles_m, tiles_n);        \
    else                                                                           \
      gemm64_kernel<TA_, TB_, 8><<<grid, NT, 0, st>>>(Q, tiles_m, tiles_n);        \
  } while (0)
  if (Q.A.trans && Q.B.trans)
    DM_GEMM(true, true);
  else if (Q.A.trans)
    DM_GEMM(true, false);
  else if (Q.B.trans)
    DM_GEMM(false, true);
  else
    DM_GEMM(false, false);
#undef DM_GEMM
  DM_LAUNCH_OK("gemm64_kernel");
  return DM_OK;
}

int sum_partials_launch(const double* part, int n_split, int64_t stride, int64_t n, double* out, cudaStream_t st) {
  if (n <= 0) return DM_OK;
  sum_partials_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(part, n_split, stride, n, out);
  DM_LAUNCH_OK("sum_partials_kernel");
  return DM_OK;
}

}  // namespace dm
