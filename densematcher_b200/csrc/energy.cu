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

// MODE 0: sums of squares only (doubly_stochastic, first sweep); MODE 1: energies + T
template <int MODE>
__global__ void __launch_bounds__(kEThreads, 1) dense_energy_kernel(const EnergyParams P) {
  extern __shared__ double sm[];
  const int p = blockIdx.x / P.max_rt, rt = blockIdx.x % P.max_rt;
  const int64_t r0 = P.off2[p], c0 = P.off1[p];
  const int n2 = int(P.off2[p + 1] - r0), n1 = int(P.off1[p + 1] - c0);
  const int row0 = rt * ET;
  if (row0 >= n2) return;
  const int k1 = P.k1, ldk = k1 + 1;
  double* Es = sm;                 // [ET][ldk]  emb2 rows of this tile
  double* Ps = Es + ET * ldk;      // [ET][ldk]  Phi1 rows of the current column tile
  double* Gs = Ps + ET * ldk;      // [ET][ET + 1]  dE/dM * a_j of the current tile
  __shared__ double red[kEThreads / 32][3];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  for (int e = t; e < ET * k1; e += kEThreads) {
    const int i = e / k1, k = e % k1;
    Es[i * ldk + k] = (row0 + i < n2) ? P.emb2[(r0 + row0 + i) * k1 + k] : 0.0;
  }
  double T[4][kMaxK / 16];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < kMaxK / 16; ++c) T[a][c] = 0.0;
  double e_p2p = 0.0, e_ent = 0.0, e_r01 = 0.0, rs2[4] = {0.0, 0.0, 0.0, 0.0};
  const double rbar = P.means ? P.means[2 * p] : 0.0, cbar = P.means ? P.means[2 * p + 1] : 0.0;
  const double n2n1 = double(n2) / double(n1);
  double ri[4], r2i[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = row0 + 4 * ty + a;
    ri[a] = (P.rs && i < n2) ? P.rs[r0 + i] - rbar : 0.0;
    r2i[a] = (MODE == 1 && P.w_st != 0.0 && i < n2) ? P.rs2[r0 + i] - 1.0 : 0.0;
  }
  const int nct = (n1 + ET - 1) / ET;
  for (int ct = 0; ct < nct; ++ct) {
    const int col0 = ct * ET;
    __syncthreads();  // previous tile's Ps / Gs are free
    for (int e = t; e < ET * k1; e += kEThreads) {
      const int j = e / k1, k = e % k1;
      Ps[j * ldk + k] = (col0 + j < n1) ? P.Phi1[(c0 + col0 + j) * P.ld1 + k] : 0.0;
    }
    __syncthreads();
    double S[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) S[a][b] = 0.0;
    for (int k = 0; k < k1; ++k) {
      double ev[4], pv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) ev[a] = Es[(4 * ty + a) * ldk + k];
#pragma unroll
      for (int b = 0; b < 4; ++b) pv[b] = Ps[(4 * tx + b) * ldk + k];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) S[a][b] = fma(ev[a], pv[b], S[a][b]);
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = col0 + 4 * tx + b;
      const bool jv = j < n1;
      const double aj = jv ? P.area1[c0 + j] : 0.0;
      const double cj = (P.cs && jv) ? P.cs[c0 + j] - cbar : 0.0;
      const double c2j = (MODE == 1 && P.w_st != 0.0 && jv) ? P.cs2[c0 + j] - n2n1 : 0.0;
      double colsq = 0.0;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const bool v = jv && (row0 + 4 * ty + a < n2);
        const double m = S[a][b] * aj;
        if (MODE == 0) {
          const double mm = v ? m * m : 0.0;
          rs2[a] += mm;
          colsq += mm;
        } else {
          double g = 0.0;
          if (v) {
            if (P.w_p2p != 0.0) {
              const double q = m * m - m;
              e_p2p += q * q;
              g += P.w_p2p * 2.0 * q * (2.0 * m - 1.0);
            }
            if (P.w_ent != 0.0) {
              const double mc = fmin(fmax(m, 0.0), 1.0);
              const double lg = log(mc + 1e-10);
              e_ent -= mc * lg;
              if (m >= 0.0 && m <= 1.0) g += P.w_ent * (-lg - mc / (mc + 1e-10));
            }
            if (P.w_r01 != 0.0) {
              const double lo = fmax(-m, 0.0), hi = fmax(m - 1.0, 0.0);
              e_r01 += lo * lo + hi * hi;
              g += P.w_r01 * (2.0 * hi - 2.0 * lo);
            }
            if (P.w_sum != 0.0) g += P.w_sum * (2.0 * cj + 2.0 * ri[a]);
            if (P.w_st != 0.0) g += P.w_st * 2.0 * m * (2.0 * c2j + 2.0 * r2i[a]);
          }
          Gs[(4 * ty + a) * (ET + 1) + 4 * tx + b] = g * aj;
        }
      }
      if (MODE == 0 && jv && colsq != 0.0) atomicAdd(P.cs2 + c0 + j, colsq);
    }
    if (MODE == 1) {
      __syncthreads();
      // T[i][c] += sum_j G[i][j] Phi1[j][c]
      for (int j = 0; j < ET; ++j) {
        double gv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) gv[a] = Gs[(4 * ty + a) * (ET + 1) + j];
#pragma unroll
        for (int c = 0; c < kMaxK / 16; ++c) {
          if (tx + 16 * c < k1) {
            const double pv = Ps[j * ldk + tx + 16 * c];
#pragma unroll
            for (int a = 0; a < 4; ++a) T[a][c] = fma(gv[a], pv, T[a][c]);
          }
        }
      }
    }
  }
  if (MODE == 0) {
    // row sums of squares: reduce over the 16 threads (tx) that share a row group
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      double v = rs2[a];
#pragma unroll
      for (int sh = 8; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
      const int i = row0 + 4 * ty + a;
      if (tx == 0 && i < n2) P.rs2[r0 + i] = v;
    }
    return;
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = row0 + 4 * ty + a;
    if (i >= n2) continue;
#pragma unroll
    for (int c = 0; c < kMaxK / 16; ++c)
      if (tx + 16 * c < k1) P.T[(r0 + i) * k1 + tx + 16 * c] = T[a][c];
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

// sumto1 in closed form: rowsum_i = emb2_i . (Phi1^T a1),  colsum_j = a_j Phi1_j . (emb2^T 1); one CTA per pair
__global__ void __launch_bounds__(256)
    sums_kernel(const double* __restrict__ emb2, const double* __restrict__ Phi1, int64_t ld1,
                const double* __restrict__ area1, const int64_t* __restrict__ off1, const int64_t* __restrict__ off2, int k1,
                double* __restrict__ rs, double* __restrict__ cs, double* __restrict__ means) {
  extern __shared__ double sm[];
  double* u = sm;        // [k1]  Phi1^T a1
  double* w = sm + k1;   // [k1]  emb2^T 1
  __shared__ double red[8][2];
  const int p = blockIdx.x, t = threadIdx.x;
  const int64_t r0 = off2[p], c0 = off1[p];
  const int n2 = int(off2[p + 1] - r0), n1 = int(off1[p + 1] - c0);
  for (int k = t; k < k1; k += 256) {
    double su = 0.0, sw = 0.0;
    for (int j = 0; j < n1; ++j) su = fma(area1[c0 + j], Phi1[(c0 + j) * ld1 + k], su);
    for (int i = 0; i < n2; ++i) sw += emb2[(r0 + i) * k1 + k];
    u[k] = su, w[k] = sw;
  }
  __syncthreads();
  double sr = 0.0, sc = 0.0;
  for (int i = t; i < n2; i += 256) {
    double s = 0.0;
    for (int k = 0; k < k1; ++k) s = fma(emb2[(r0 + i) * k1 + k], u[k], s);
    rs[r0 + i] = s;
    sr += s;
  }
  for (int j = t; j < n1; j += 256) {
    double s = 0.0;
    for (int k = 0; k < k1; ++k) s = fma(Phi1[(c0 + j) * ld1 + k], w[k], s);
    s *= area1[c0 + j];
    cs[c0 + j] = s;
    sc += s;
  }
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) {
    sr += __shfl_xor_sync(0xffffffffu, sr, sh);
    sc += __shfl_xor_sync(0xffffffffu, sc, sh);
  }
  if ((t & 31) == 0) red[t >> 5][0] = sr, red[t >> 5][1] = sc;
  __syncthreads();
  if (t == 0) {
    for (int q = 1; q < 8; ++q) sr += red[q][0], sc += red[q][1];
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
  const size_t shm = sizeof(double) * (2 * size_t(ET) * (k1 + 1) + size_t(ET) * (ET + 1));
  static OncePerDevice attr_once;
  if (attr_once.first()) {
    DM_CUDA_OK(cudaFuncSetAttribute(dense_energy_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DM_CUDA_OK(cudaFuncSetAttribute(dense_energy_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  const unsigned grid = unsigned(n_pairs) * L.max_rt;
  if (w_stochastic != 0.0) {
    DM_CUDA_OK(cudaMemsetAsync(L.cs2, 0, sizeof(double) * size_t(total_n1), st));
    dense_energy_kernel<0><<<grid, kEThreads, shm, st>>>(P);
    DM_LAUNCH_OK("dense_energy_kernel<0>");
  }
  dense_energy_kernel<1><<<grid, kEThreads, shm, st>>>(P);
  DM_LAUNCH_OK("dense_energy_kernel<1>");
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
