// Dense-map energy terms of the functional-map fit and their gradient.
//
// The reference evaluates, per L-BFGS callback, the dense n2 x n1 matrix  M = Phi2 C Phi1^T A1  with A1 densified
// to n1 x n1, an element-wise loss on it, and its autograd backward (densematcher/pyFM/optimize/base_functions.py:
// p2p :296-325, doubly_stochastic :327-361, entropy :363-372, range01 :374-385, sumto1 :387-428).  Here M is never
// stored: a CTA owns 64 rows of one pair, keeps emb2 = (Phi2 C)[rows] resident in shared memory, sweeps the columns in
// tiles of 64, forms the tile of M on the FP64 pipe, applies the loss and its derivative element-wise and immediately
// contracts the derivative tile with Phi1 (T = dE/dM A1 Phi1, 64 x k1 in registers) -- the same "tile, reduce, never
// materialise" shape as the nearest-neighbour kernel.  The gradient is then Phi2^T T (one batched GEMM).
// The row/column sums that sumto1 needs have closed forms in O(N k); doubly_stochastic needs one extra sweep for the
// sums of squares.  Everything is float64 (the reference runs this in float32; its L-BFGS result moves by ~1e-4 with
// that rounding).
#include "dm_internal.cuh"
#include "gemm64.cuh"

namespace dm {
namespace {

constexpr int ET = 64;        // tile edge
constexpr int kEThreads = 256;
constexpr int kMaxK = 128;    // eigenbasis size supported by the register tile of T

struct EnergyParams {
  const double* emb2;  // [total_n2, k1]
  const double* Phi1;  // [total_n1, ld1]
  int64_t ld1;
  const double* area1;
  const int64_t* off1;
  const int64_t* off2;
  int k1, max_rt;
  double w_p2p, w_st, w_ent, w_r01, w_sum;
  // sumto1: row / column sums and their means (per pair: [2] = rbar, cbar)
  const double* rs;
  const double* cs;
  const double* means;
  // doubly stochastic: row / column sums of squares
  double* rs2;
  double* cs2;
  double* T;         // [total_n2, k1]
  double* partial;   // [n_pairs * max_rt][3]  p2p, ent, range01
};

__host__ __device__ inline int energy_ldk(int k1) { return ((k1 + 7) / 8 * 8 + 15) / 16 * 16 + 4; }  // pitch = 4 (mod 16)
constexpr int kLdG = ET + 4;

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// MODE 0: sums of squares only (doubly_stochastic, first sweep); MODE 1: energies + T.
// Both contractions run on the float64 tensor cores (mma.sync m8n8k4): the tile S = E P^T (64 x 64 x k1) with eight warps
// in a 2 x 4 grid (warp tile 32 x 16), the loss and its derivative applied in the accumulator layout, and T += G P
// (64 x k1 x 64) with the derivative tile staged through shared memory.  Row pitches = 4 (mod 16) doubles make every
// fragment load conflict-free.  (The first version used 4 x 4 register tiles on the DFMA pipe: 2.8 TFLOP/s, shared-memory
// bound at 8 loads per 16 FMAs.)
// NCB: column blocks of T per warp (k1 <= 32 NCB): the small eigenbases of the fit (k = 15 ... 50) leave room for two CTAs
// per SM, which is what hides the float64 latency of the element-wise part (log, divisions)
// FAST: the logarithm and the division of the entropy term in float32 (the reference evaluates every term in float32,
// base_functions.py:363-372 on float32 tensors); everything else, and every accumulation, stays float64.  The float64
// log / division were ~200 of the 240 instructions per element of M.
// TERMS: which terms can be active (bit 0 p2p, 1 stochastic, 2 entropy, 3 range01, 4 sumto1).  kAllTerms tests the weights
// at run time; the notebook's default set (entropy + sumto1) has its own instantiation, without the per-element tests
// and the dead branches of the others (~150 -> ~50 instructions per element of M).
constexpr int kTermP2P = 1, kTermSt = 2, kTermEnt = 4, kTermR01 = 8, kTermSum = 16, kAllTerms = 31;
template <int MODE, int NCB, bool FAST, int TERMS>
__global__ void __launch_bounds__(kEThreads, NCB <= 2 ? 2 : 1) dense_energy_kernel(const EnergyParams P) {
  constexpr bool kSpec = TERMS != kAllTerms;  // specialised: every term of TERMS is active, the others are absent
  const bool on_p2p = (TERMS & kTermP2P) && (kSpec || P.w_p2p != 0.0), on_st = (TERMS & kTermSt) && (kSpec || P.w_st != 0.0);
  const bool on_ent = (TERMS & kTermEnt) && (kSpec || P.w_ent != 0.0), on_r01 = (TERMS & kTermR01) && (kSpec || P.w_r01 != 0.0);
  const bool on_sum = (TERMS & kTermSum) && (kSpec || P.w_sum != 0.0);
  extern __shared__ double sm[];
  const int p = blockIdx.x / P.max_rt, rt = blockIdx.x % P.max_rt;
  const int64_t r0 = P.off2[p], c0 = P.off1[p];
  const int n2 = int(P.off2[p + 1] - r0), n1 = int(P.off1[p + 1] - c0);
  const int row0 = rt * ET;
  if (row0 >= n2) return;
  const int k1 = P.k1, ldk = energy_ldk(k1), kp4 = (k1 + 3) & ~3, kp8 = (k1 + 7) & ~7;
  double* Es = sm;                 // [ET][ldk]  emb2 rows of this tile (zero beyond k1)
  double* Ps = Es + ET * ldk;      // [ET][ldk]  Phi1 rows of the current column tile
  double* Gs = Ps + ET * ldk;      // [ET][kLdG] dE/dM * a_j of the current tile
  __shared__ double red[kEThreads / 32][3];
  __shared__ double rowsq[4][ET];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int rh = warp & 1, cq = warp >> 1;  // row half (32 rows), column quarter (16 columns) of the score tile
  for (int e = t; e < ET * kp8; e += kEThreads) {
    const int i = e / kp8, k = e % kp8;
    Es[i * ldk + k] = (row0 + i < n2 && k < k1) ? P.emb2[(r0 + row0 + i) * k1 + k] : 0.0;
  }
  // T: rows rh*32 + 8 a + g, column blocks cq + 4 c (interleaved: a small k1 still spreads over the four warp columns)
  double T[4][NCB][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < NCB; ++c) T[a][c][0] = T[a][c][1] = 0.0;
  const int n_cb = kp8 / 8;  // column blocks of T
  double e_p2p = 0.0, e_ent = 0.0, e_r01 = 0.0, rs2[4] = {0.0, 0.0, 0.0, 0.0};
  const double rbar = P.means ? P.means[2 * p] : 0.0, cbar = P.means ? P.means[2 * p + 1] : 0.0;
  const double n2n1 = double(n2) / double(n1);
  double ri[4], r2i[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = row0 + rh * 32 + 8 * a + g;
    ri[a] = (on_sum && P.rs && i < n2) ? 2.0 * P.w_sum * (P.rs[r0 + i] - rbar) : 0.0;   // row part of the sumto1 gradient
    r2i[a] = (MODE == 1 && on_st && i < n2) ? P.rs2[r0 + i] - 1.0 : 0.0;
  }
  const int nct = (n1 + ET - 1) / ET;
  for (int ct = 0; ct < nct; ++ct) {
    const int col0 = ct * ET;
    __syncthreads();  // previous tile's Ps / Gs are free
    for (int e = t; e < ET * kp8; e += kEThreads) {
      const int j = e / kp8, k = e % kp8;
      Ps[j * ldk + k] = (col0 + j < n1 && k < k1) ? P.Phi1[(c0 + col0 + j) * P.ld1 + k] : 0.0;
    }
    __syncthreads();
    double S[4][2][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) S[a][b][0] = S[a][b][1] = 0.0;
    {
      const double* ea = Es + (rh * 32 + g) * ldk + t4;
      const double* pb = Ps + (cq * 16 + g) * ldk + t4;
      for (int k = 0; k < kp4; k += 4) {
        double av[4], bv[2];
#pragma unroll
        for (int a = 0; a < 4; ++a) av[a] = ea[8 * a * ldk + k];
#pragma unroll
        for (int b = 0; b < 2; ++b) bv[b] = pb[8 * b * ldk + k];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b) dmma884(S[a][b], av[a], bv[b]);
      }
    }
#pragma unroll
    for (int b = 0; b < 2; ++b) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int jl = cq * 16 + 8 * b + 2 * t4 + h, j = col0 + jl;
        const bool jv = j < n1;
        const double aj = jv ? P.area1[c0 + j] : 0.0;
        const double cj = (on_sum && P.cs && jv) ? 2.0 * P.w_sum * (P.cs[c0 + j] - cbar) : 0.0;  // column part, as above
        const double c2j = (MODE == 1 && on_st && jv) ? P.cs2[c0 + j] - n2n1 : 0.0;
        double colsq = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int il = rh * 32 + 8 * a + g;
          const bool v = jv && (row0 + il < n2);
          const double m = S[a][b][h] * aj;
          if (MODE == 0) {
            const double mm = v ? m * m : 0.0;
            rs2[a] += mm;
            colsq += mm;
          } else {
            double gd = 0.0;
            if (v) {
              if (on_p2p) {
                const double q = m * m - m;
                e_p2p += q * q;
                gd += P.w_p2p * 2.0 * q * (2.0 * m - 1.0);
              }
              if (on_ent) {
                const double mc = fmin(fmax(m, 0.0), 1.0);
                double lg, ratio;
                if (FAST) {
                  const float x = float(mc + 1e-10);
                  lg = double(logf(x));
                  ratio = double(__fdividef(float(mc), x));
                } else {
                  lg = log(mc + 1e-10);
                  ratio = mc / (mc + 1e-10);
                }
                e_ent -= mc * lg;
                if (m >= 0.0 && m <= 1.0) gd += P.w_ent * (-lg - ratio);
              }
              if (on_r01) {
                const double lo = fmax(-m, 0.0), hi = fmax(m - 1.0, 0.0);
                e_r01 += lo * lo + hi * hi;
                gd += P.w_r01 * (2.0 * hi - 2.0 * lo);
              }
              if (on_sum) gd += cj + ri[a];
              if (on_st) gd += P.w_st * 2.0 * m * (2.0 * c2j + 2.0 * r2i[a]);
            }
            Gs[il * kLdG + jl] = gd * aj;
          }
        }
        if (MODE == 0) {
          // column sum of squares over this warp's 32 rows: lanes that share (t4, h) differ in g
#pragma unroll
          for (int sh = 4; sh < 32; sh <<= 1) colsq += __shfl_xor_sync(0xffffffffu, colsq, sh);
          if (g == 0 && jv && colsq != 0.0) atomicAdd(P.cs2 + c0 + j, colsq);
        }
      }
    }
    if (MODE == 1) {
      __syncthreads();
      // T[i][c] += sum_j G[i][j] Phi1[j][c]
      const double* ga = Gs + (rh * 32 + g) * kLdG + t4;
      for (int j = 0; j < ET; j += 4) {
        double av[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) av[a] = ga[8 * a * kLdG + j];
#pragma unroll
        for (int c = 0; c < NCB; ++c) {
          const int cb = cq + 4 * c;
          if (cb < n_cb) {
            const double bv = Ps[(j + t4) * ldk + 8 * cb + g];
#pragma unroll
            for (int a = 0; a < 4; ++a) dmma884(T[a][c], av[a], bv);
          }
        }
      }
    }
  }
  if (MODE == 0) {
    // row sums of squares: the four lanes of a row (t4), then the four warps of a row half (cq)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      double v = rs2[a];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (t4 == 0) rowsq[cq][rh * 32 + 8 * a + g] = v;
    }
    __syncthreads();
    if (t < ET && row0 + t < n2) P.rs2[r0 + row0 + t] = rowsq[0][t] + rowsq[1][t] + rowsq[2][t] + rowsq[3][t];
    return;
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = row0 + rh * 32 + 8 * a + g;
    if (i >= n2) continue;
#pragma unroll
    for (int c = 0; c < NCB; ++c) {
      const int cb = cq + 4 * c;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col = 8 * cb + 2 * t4 + h;
        if (cb < n_cb && col < k1) P.T[(r0 + i) * k1 + col] = T[a][c][h];
      }
    }
  }
  double e3[3] = {e_p2p, e_ent, e_r01};
#pragma unroll
  for (int q = 0; q < 3; ++q) {
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) e3[q] += __shfl_xor_sync(0xffffffffu, e3[q], sh);
    if ((t & 31) == 0) red[t >> 5][q] = e3[q];
  }
  __syncthreads();
  if (t < 3) {
    double s = 0.0;
    for (int w = 0; w < kEThreads / 32; ++w) s += red[w][t];
    P.partial[(int64_t(p) * P.max_rt + rt) * 3 + t] = s;
  }
}

// sumto1 in closed form: rowsum_i = emb2_i . (Phi1^T a1),  colsum_j = a_j Phi1_j . (emb2^T 1); one CTA per pair.
// Warps walk the rows (lanes over the k1 <= 128 columns: coalesced row reads), partial sums merged through shared memory;
// the first version gave each of the k1 column sums to ONE thread striding down the matrix (0.49 ms per call: 20 % of a
// fit iteration).
__global__ void __launch_bounds__(256)
    sums_kernel(const double* __restrict__ emb2, const double* __restrict__ Phi1, int64_t ld1,
                const double* __restrict__ area1, const int64_t* __restrict__ off1, const int64_t* __restrict__ off2, int k1,
                double* __restrict__ rs, double* __restrict__ cs, double* __restrict__ means) {
  extern __shared__ double sm[];
  double* u = sm;        // [k1]  Phi1^T a1
  double* w = sm + k1;   // [k1]  emb2^T 1
  __shared__ double part[8][2][kMaxK];
  __shared__ double red[8][2];
  const int p = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int64_t r0 = off2[p], c0 = off1[p];
  const int n2 = int(off2[p + 1] - r0), n1 = int(off1[p + 1] - c0);
  double au[4] = {0.0, 0.0, 0.0, 0.0}, aw[4] = {0.0, 0.0, 0.0, 0.0};
  for (int j = warp; j < n1; j += 8) {
    const double a = area1[c0 + j];
    const double* row = Phi1 + (c0 + j) * ld1;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (lane + 32 * q < k1) au[q] = fma(a, row[lane + 32 * q], au[q]);
  }
  for (int i = warp; i < n2; i += 8) {
    const double* row = emb2 + (r0 + i) * k1;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (lane + 32 * q < k1) aw[q] += row[lane + 32 * q];
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (lane + 32 * q < k1) part[warp][0][lane + 32 * q] = au[q], part[warp][1][lane + 32 * q] = aw[q];
  __syncthreads();
  for (int k = t; k < k1; k += 256) {
    double su = 0.0, sw = 0.0;
    for (int q = 0; q < 8; ++q) su += part[q][0][k], sw += part[q][1][k];
    u[k] = su, w[k] = sw;
  }
  __syncthreads();
  double uu[4], ww[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uu[q] = lane + 32 * q < k1 ? u[lane + 32 * q] : 0.0;
    ww[q] = lane + 32 * q < k1 ? w[lane + 32 * q] : 0.0;
  }
  double sr = 0.0, sc = 0.0;
  for (int i = warp; i < n2; i += 8) {
    const double* row = emb2 + (r0 + i) * k1;
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (lane + 32 * q < k1) s = fma(row[lane + 32 * q], uu[q], s);
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
    if (lane == 0) rs[r0 + i] = s, sr += s;
  }
  for (int j = warp; j < n1; j += 8) {
    const double* row = Phi1 + (c0 + j) * ld1;
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (lane + 32 * q < k1) s = fma(row[lane + 32 * q], ww[q], s);
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
    s *= area1[c0 + j];
    if (lane == 0) cs[c0 + j] = s, sc += s;
  }
  if (lane == 0) red[warp][0] = sr, red[warp][1] = sc;
  __syncthreads();
  if (t == 0) {
    sr = sc = 0.0;
    for (int q = 0; q < 8; ++q) sr += red[q][0], sc += red[q][1];
    means[2 * p] = sr / double(n2);
    means[2 * p + 1] = sc / double(n1);
  }
}

// per pair: energy[p][0..4] = p2p, stochastic, ent, range01, sumto1 (unweighted)
__global__ void __launch_bounds__(256)
    energy_finalize_kernel(const double* __restrict__ partial, int max_rt, const int64_t* __restrict__ off1,
                           const int64_t* __restrict__ off2, const double* __restrict__ rs, const double* __restrict__ cs,
                           const double* __restrict__ means, const double* __restrict__ rs2, const double* __restrict__ cs2,
                           int want_sum, int want_st, double* __restrict__ energy) {
  const int p = blockIdx.x, t = threadIdx.x;
  const int64_t r0 = off2[p], c0 = off1[p];
  const int n2 = int(off2[p + 1] - r0), n1 = int(off1[p + 1] - c0);
  const int nrt = (n2 + ET - 1) / ET;
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int rt = t; rt < nrt; rt += 256) {
    const double* q = partial + (int64_t(p) * max_rt + rt) * 3;
    acc[0] += q[0], acc[2] += q[1], acc[3] += q[2];
  }
  if (want_sum) {
    const double rbar = means[2 * p], cbar = means[2 * p + 1];
    for (int i = t; i < n2; i += 256) acc[4] += (rs[r0 + i] - rbar) * (rs[r0 + i] - rbar);
    for (int j = t; j < n1; j += 256) acc[4] += (cs[c0 + j] - cbar) * (cs[c0 + j] - cbar);
  }
  if (want_st) {
    const double tgt = double(n2) / double(n1);
    for (int i = t; i < n2; i += 256) acc[1] += (rs2[r0 + i] - 1.0) * (rs2[r0 + i] - 1.0);
    for (int j = t; j < n1; j += 256) acc[1] += (cs2[c0 + j] - tgt) * (cs2[c0 + j] - tgt);
  }
  __shared__ double red[8][5];
#pragma unroll
  for (int q = 0; q < 5; ++q) {
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], sh);
    if ((t & 31) == 0) red[t >> 5][q] = acc[q];
  }
  __syncthreads();
  if (t < 5) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w][t];
    energy[p * 5 + t] = s;
  }
}

struct EnergyLayout {
  double *emb2, *T, *rs, *cs, *means, *rs2, *cs2, *partial, *gpart;
  int max_rt, ksplit;
  size_t bytes;
};
EnergyLayout energy_carve(void* ws, int n_pairs, int64_t total_n1, int64_t total_n2, int max_n2, int k1, int k2) {
  Carver c(ws);
  EnergyLayout L;
  L.max_rt = (max_n2 + ET - 1) / ET;
  L.ksplit = (max_n2 + 255) / 256;
  L.emb2 = c.take<double>(size_t(total_n2) * k1);
  L.T = c.take<double>(size_t(total_n2) * k1);
  L.rs = c.take<double>(size_t(total_n2));
  L.cs = c.take<double>(size_t(total_n1));
  L.means = c.take<double>(size_t(n_pairs) * 2);
  L.rs2 = c.take<double>(size_t(total_n2));
  L.cs2 = c.take<double>(size_t(total_n1));
  L.partial = c.take<double>(size_t(n_pairs) * L.max_rt * 3);
  L.gpart = c.take<double>(L.ksplit > 1 ? size_t(L.ksplit) * n_pairs * k1 * k2 : 0);
  L.bytes = c.bytes();
  return L;
}

}  // namespace
}  // namespace dm

using namespace dm;

extern "C" {

size_t dm_dense_energy_workspace_bytes(int n_pairs, int64_t total_n1, int64_t total_n2, int max_n1, int max_n2, int k1,
                                       int k2) {
  (void)max_n1;
  if (n_pairs < 0 || total_n1 < 0 || total_n2 < 0 || k1 <= 0 || k2 <= 0) return 0;
  return energy_carve(nullptr, n_pairs, total_n1, total_n2, max_n2, k1, k2).bytes;
}

int dm_dense_energy(const double* C, int k1, int k2, const double* Phi1, int64_t ld1, const int64_t* off1,
                    int64_t total_n1, int max_n1, const double* Phi2, int64_t ld2, const int64_t* off2, int64_t total_n2,
                    int max_n2, const double* area1, int n_pairs, double w_p2p, double w_stochastic, double w_ent,
                    double w_range01, double w_sumto1, double* energy, double* grad, void* workspace,
                    size_t workspace_bytes, dm_stream_t stream) {
  return dm_dense_energy_ex(C, k1, k2, Phi1, ld1, off1, total_n1, max_n1, Phi2, ld2, off2, total_n2, max_n2, area1, n_pairs,
                            w_p2p, w_stochastic, w_ent, w_range01, w_sumto1, energy, grad, 0, workspace, workspace_bytes, stream);
}

int dm_dense_energy_ex(const double* C, int k1, int k2, const double* Phi1, int64_t ld1, const int64_t* off1,
                       int64_t total_n1, int max_n1, const double* Phi2, int64_t ld2, const int64_t* off2, int64_t total_n2,
                       int max_n2, const double* area1, int n_pairs, double w_p2p, double w_stochastic, double w_ent,
                       double w_range01, double w_sumto1, double* energy, double* grad, int flags, void* workspace,
                       size_t workspace_bytes, dm_stream_t stream) {
  if (n_pairs < 0 || k1 <= 0 || k2 <= 0 || total_n1 < 0 || total_n2 < 0) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n_pairs == 0) return DM_OK;
  if (!C || !Phi1 || !Phi2 || !off1 || !off2 || !area1 || !energy || !grad) DM_FAIL(DM_ERR_BADARG, "null argument");
  if (ld1 < k1 || ld2 < k2) DM_FAIL(DM_ERR_BADARG, "eigenbasis has fewer columns than the functional map");
  if (k1 > kMaxK) DM_FAIL(DM_ERR_UNSUPPORTED, "dense-map energy terms support k1 <= %d", kMaxK);
  if (!workspace) DM_FAIL(DM_ERR_WORKSPACE, "workspace is null");
  if (reinterpret_cast<uintptr_t>(workspace) % 256) DM_FAIL(DM_ERR_ALIGN, "workspace must be 256-byte aligned");
  EnergyLayout L = energy_carve(workspace, n_pairs, total_n1, total_n2, max_n2, k1, k2);
  if (L.bytes > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu", L.bytes);
  (void)max_n1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  {  // emb2 = Phi2[:, :k2] C
    GemmProblem G;
    G.A.d = Phi2, G.A.ld = ld2, G.A.off = off2, G.A.trans = 0;
    G.B.d = C, G.B.ld = k1, G.B.batch_stride = int64_t(k1) * k2, G.B.rows = k2, G.B.trans = 1;
    G.N = k1, G.K = k2, G.maxM = max_n2, G.maxN = k1, G.maxK = k2, G.n_batch = n_pairs;
    G.C = L.emb2, G.ldc = k1, G.c_off = off2;
    if ((rc = gemm64_launch(G, st))) return rc;
  }
  EnergyParams P{};
  P.emb2 = L.emb2, P.Phi1 = Phi1, P.ld1 = ld1, P.area1 = area1, P.off1 = off1, P.off2 = off2, P.k1 = k1, P.max_rt = L.max_rt;
  P.w_p2p = w_p2p, P.w_st = w_stochastic, P.w_ent = w_ent, P.w_r01 = w_range01, P.w_sum = w_sumto1;
  P.rs2 = L.rs2, P.cs2 = L.cs2, P.T = L.T, P.partial = L.partial;
  if (w_sumto1 != 0.0) {
    sums_kernel<<<n_pairs, 256, 2 * k1 * sizeof(double), st>>>(L.emb2, Phi1, ld1, area1, off1, off2, k1, L.rs, L.cs, L.means);
    DM_LAUNCH_OK("sums_kernel");
    P.rs = L.rs, P.cs = L.cs, P.means = L.means;
  }
  const size_t shm = sizeof(double) * (2 * size_t(ET) * energy_ldk(k1) + size_t(ET) * kLdG);
  const unsigned grid = unsigned(n_pairs) * L.max_rt;
  if (w_stochastic != 0.0) DM_CUDA_OK(cudaMemsetAsync(L.cs2, 0, sizeof(double) * size_t(total_n1), st));
  const bool ent_sum = w_p2p == 0.0 && w_stochastic == 0.0 && w_range01 == 0.0 && w_ent != 0.0 && w_sumto1 != 0.0;
#define DM_ENERGY1(NCB_, FAST_, TERMS_)                                                                                       \
  do {                                                                                                                        \
    static OncePerDevice once1;                                                                                               \
    if (once1.first())                                                                                                        \
      DM_CUDA_OK(cudaFuncSetAttribute(dense_energy_kernel<1, NCB_, FAST_, TERMS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      200 * 1024));                                                                           \
    dense_energy_kernel<1, NCB_, FAST_, TERMS_><<<grid, kEThreads, shm, st>>>(P);                                             \
  } while (0)
#define DM_ENERGY(NCB_)                                                                                                       \
  do {                                                                                                                        \
    if (w_stochastic != 0.0) {                                                                                                \
      static OncePerDevice once0;                                                                                             \
      if (once0.first())                                                                                                      \
        DM_CUDA_OK(cudaFuncSetAttribute(dense_energy_kernel<0, NCB_, false, kAllTerms>,                                        \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));                            \
      dense_energy_kernel<0, NCB_, false, kAllTerms><<<grid, kEThreads, shm, st>>>(P);                                        \
    }                                                                                                                         \
    const bool fast = (flags & DM_FAST_LOSS) != 0;                                                                            \
    if (ent_sum && fast) DM_ENERGY1(NCB_, true, kTermEnt | kTermSum);                                                         \
    else if (ent_sum) DM_ENERGY1(NCB_, false, kTermEnt | kTermSum);                                                           \
    else if (fast) DM_ENERGY1(NCB_, true, kAllTerms);                                                                         \
    else DM_ENERGY1(NCB_, false, kAllTerms);                                                                                  \
  } while (0)
  if (k1 <= 32) DM_ENERGY(1); else if (k1 <= 64) DM_ENERGY(2); else DM_ENERGY(4);
#undef DM_ENERGY1
#undef DM_ENERGY
  DM_LAUNCH_OK("dense_energy_kernel");
  energy_finalize_kernel<<<n_pairs, 256, 0, st>>>(L.partial, L.max_rt, off1, off2, L.rs, L.cs, L.means, L.rs2, L.cs2,
                                                  w_sumto1 != 0.0, w_stochastic != 0.0, energy);
  DM_LAUNCH_OK("energy_finalize_kernel");
  // grad = Phi2[:, :k2]^T T   (k2 x k1), split over the vertices
  GemmProblem G;
  G.A.d = Phi2, G.A.ld = ld2, G.A.off = off2, G.A.trans = 1;
  G.B.d = L.T, G.B.ld = k1, G.B.off = off2, G.B.trans = 1;
  G.M = k2, G.N = k1, G.maxM = k2, G.maxN = k1, G.maxK = max_n2, G.n_batch = n_pairs;
  G.ldc = k1, G.c_batch_stride = int64_t(k1) * k2;
  if (L.ksplit <= 1) {
    G.C = grad;
    return gemm64_launch(G, st);
  }
  G.C = L.gpart, G.ksplit = L.ksplit, G.kchunk = 256, G.split_stride = int64_t(n_pairs) * k1 * k2;
  if ((rc = gemm64_launch(G, st))) return rc;
  return sum_partials_launch(L.gpart, L.ksplit, G.split_stride, G.split_stride, grad, st);
}

}  // extern "C"

extern "C" int dm_bmm_nt_f64(const double* A, const double* B, int n_batch, int m, int n, int k, double* C, dm_stream_t stream) {
  if (n_batch < 0 || m <= 0 || n <= 0 || k <= 0) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n_batch == 0) return DM_OK;
  if (!A || !B || !C) DM_FAIL(DM_ERR_BADARG, "null argument");
  GemmProblem G;
  G.A.d = A, G.A.ld = k, G.A.batch_stride = int64_t(m) * k, G.A.rows = m, G.A.trans = 0;
  G.B.d = B, G.B.ld = k, G.B.batch_stride = int64_t(n) * k, G.B.rows = n, G.B.trans = 0;
  G.M = m, G.N = n, G.K = k, G.maxM = m, G.maxN = n, G.maxK = k, G.n_batch = n_batch;
  G.C = C, G.ldc = n, G.c_batch_stride = int64_t(m) * n;
  return gemm64_launch(G, static_cast<cudaStream_t>(stream));
}
