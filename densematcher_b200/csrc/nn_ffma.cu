// CUDA-core fp32 engine of the fused similarity + argmax pass.
//
// S = Y X^T is never written: each CTA owns 128 query rows of one mesh pair, sweeps all database
// column tiles, and reduces every 128x128 fp32 tile in registers into running (max, argmax, 2nd max)
// per row (row epilogues) and per-row-tile partials per column (column epilogues).
// Replaces the kd-tree search of knn_query (densematcher/pyFM/spectral/nn_utils.py:28-30) and the
// dense argmax of functional_map.py:49-50.  Used for small / odd inner dimensions (the k-dimensional
// spectral embeddings of FM_to_p2p and ZoomOut) and as the on-device cross-check of the tcgen05 engine.
#include "dm_internal.cuh"

namespace dm {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;

__device__ __forceinline__ int sub_idx(int t4, int a) {  // a in 0..7 -> offset inside the 128-wide tile
  return (a < 4) ? (t4 * 4 + a) : (64 + t4 * 4 + (a - 4));
}

template <bool VEC>
__device__ __forceinline__ void gload(const float* __restrict__ row, bool valid, int d, int k, float4& r) {
  if (VEC) {
    if (valid && k < d)
      r = __ldg(reinterpret_cast<const float4*>(row + k));
    else
      r = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    r.x = (valid && k + 0 < d) ? __ldg(row + k + 0) : 0.f;
    r.y = (valid && k + 1 < d) ? __ldg(row + k + 1) : 0.f;
    r.z = (valid && k + 2 < d) ? __ldg(row + k + 2) : 0.f;
    r.w = (valid && k + 3 < d) ? __ldg(row + k + 3) : 0.f;
  }
}

__device__ __forceinline__ void sstore(float (*T)[BM], int kq, int lrow, const float4& r0, const float4& r1) {
  T[kq + 0][lrow] = r0.x;
  T[kq + 1][lrow] = r0.y;
  T[kq + 2][lrow] = r0.z;
  T[kq + 3][lrow] = r0.w;
  T[kq + 8][lrow] = r1.x;
  T[kq + 9][lrow] = r1.y;
  T[kq + 10][lrow] = r1.z;
  T[kq + 11][lrow] = r1.w;
}

// acc[a][b] = sum_k Yt[row(a)][k] * Xt[col(b)][k] for one 128x128 tile; K sequential in fp32 FMA.
template <bool VEC>
__device__ __forceinline__ void tile_scores(float (&acc)[8][8], float* smem, const float* __restrict__ Yrow,
                                            bool yvalid, const float* __restrict__ Xrow, bool xvalid, int d) {
  float(*As)[BK][BM] = reinterpret_cast<float(*)[BK][BM]>(smem);
  float(*Bs)[BK][BN] = reinterpret_cast<float(*)[BK][BN]>(smem + 2 * BK * BM);
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int lrow = t & 127, kq = (t >> 7) * 4;
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
  const int nk = (d + BK - 1) / BK;
  float4 ra0, ra1, rb0, rb1;
  gload<VEC>(Yrow, yvalid, d, kq, ra0);
  gload<VEC>(Yrow, yvalid, d, kq + 8, ra1);
  gload<VEC>(Xrow, xvalid, d, kq, rb0);
  gload<VEC>(Xrow, xvalid, d, kq + 8, rb1);
  sstore(As[0], kq, lrow, ra0, ra1);
  sstore(Bs[0], kq, lrow, rb0, rb1);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      const int k0 = (kt + 1) * BK + kq;
      gload<VEC>(Yrow, yvalid, d, k0, ra0);
      gload<VEC>(Yrow, yvalid, d, k0 + 8, ra1);
      gload<VEC>(Xrow, xvalid, d, k0, rb0);
      gload<VEC>(Xrow, xvalid, d, k0 + 8, rb1);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    if (kt + 1 < nk) {
      sstore(As[cur ^ 1], kq, lrow, ra0, ra1);
      sstore(Bs[cur ^ 1], kq, lrow, rb0, rb1);
    }
    __syncthreads();
  }
}

template <int NR, int NC, bool VEC>
__global__ void __launch_bounds__(NT, 2) nn_ffma_kernel(const NNProblem P) {
  const int p = blockIdx.x / P.max_rt, rt = blockIdx.x % P.max_rt;
  const int64_t q0 = P.q_off[p];
  const int nq = int(P.q_off[p + 1] - q0);
  const int row0 = rt * BM;
  if (row0 >= nq) return;
  const int64_t d0 = P.db_off[p];
  const int nd = int(P.db_off[p + 1] - d0);

  // 32 KB of tiles, re-used (and sized) for the column reduction: [16][BN] x 20 B = 40 KB
  __shared__ __align__(16) float smem[16 * BN * sizeof(Top3) / sizeof(float)];
  static_assert(sizeof(smem) >= 2 * 2 * BK * BM * sizeof(float), "tile buffers must fit");
  __shared__ Top3 rowstate[NR > 0 ? NR : 1][BM];
  Top3(*red)[BN] = reinterpret_cast<Top3(*)[BN]>(smem);

  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int lrow = t & 127;
  if (t < BM)
#pragma unroll
    for (int r = 0; r < NR; ++r) rowstate[r][t] = top3_init();

  const bool yvalid = row0 + lrow < nq;
  const float* Yrow = P.Y + (q0 + row0 + (yvalid ? lrow : 0)) * P.ldY;

  for (int col0 = 0; col0 < nd; col0 += BN) {
    const bool xvalid = col0 + lrow < nd;
    const float* Xrow = P.X + (d0 + col0 + (xvalid ? lrow : 0)) * P.ldX;
    float acc[8][8];
    tile_scores<VEC>(acc, smem, Yrow, yvalid, Xrow, xvalid, P.d_fast);

    // ---- row epilogues: reduce over the 8 columns held here, then over the 16 threads sharing a row
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      float cs[8], cb[8];
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const int j = col0 + sub_idx(tx, b);
        const bool v = j < nd;
        cs[b] = v ? __ldg(P.row[r].sf + d0 + j) : 0.f;
        cb[b] = v ? __ldg(P.row[r].bf + d0 + j) : -INFINITY;
      }
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        Top3 s = top3_init();
#pragma unroll
        for (int b = 0; b < 8; ++b) top3_push(s, fmaf(acc[a][b], cs[b], cb[b]), col0 + sub_idx(tx, b));
#pragma unroll
        for (int sh = 1; sh < 16; sh <<= 1) {
          const float om1 = __shfl_xor_sync(0xffffffffu, s.m1, sh);
          const int oi1 = __shfl_xor_sync(0xffffffffu, s.i1, sh);
          const float om2 = __shfl_xor_sync(0xffffffffu, s.m2, sh);
          const int oi2 = __shfl_xor_sync(0xffffffffu, s.i2, sh);
          const float om3 = __shfl_xor_sync(0xffffffffu, s.m3, sh);
          top3_merge(s, om1, oi1, om2, oi2, om3);
        }
        if (tx == 0) {
          Top3 cur = rowstate[r][sub_idx(ty, a)];
          top3_merge(cur, s);
          rowstate[r][sub_idx(ty, a)] = cur;
        }
      }
    }

    // ---- column epilogues: reduce over the 8 rows held here, then over the 16 threads sharing a column
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float rs[8], rb[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int i = row0 + sub_idx(ty, a);
        const bool v = i < nq;
        rs[a] = v ? __ldg(P.col[c].sf + q0 + i) : 0.f;
        rb[a] = v ? __ldg(P.col[c].bf + q0 + i) : -INFINITY;
      }
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        Top3 s = top3_init();
#pragma unroll
        for (int a = 0; a < 8; ++a) top3_push(s, fmaf(acc[a][b], rs[a], rb[a]), row0 + sub_idx(ty, a));
        red[ty][sub_idx(tx, b)] = s;
      }
      __syncthreads();
      if (t < BN) {
        Top3 m = red[0][t];
#pragma unroll
        for (int y = 1; y < 16; ++y) top3_merge(m, red[y][t]);
        const int j = col0 + t;
        if (j < nd)
          P.col_partial[((int64_t(c) * P.n_pairs + p) * P.max_rt + rt) * P.max_db + j] = m;
      }
      __syncthreads();
    }
  }

  if (tx == 0) {
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int li = sub_idx(ty, a);
        const int i = row0 + li;
        if (i < nq) emit_result(P, P.row[r], false, r, p, q0 + i, i, P.norm_q[q0 + i], rowstate[r][li]);
      }
  }
}

template <bool VEC>
__global__ void __launch_bounds__(NT, 2)
    nn_ffma_scores_kernel(const float* Y, int64_t ldY, int nq, const float* X, int64_t ldX, int nd, int d, float* S,
                          int64_t ldS) {
  __shared__ __align__(16) float smem[2 * 2 * BK * BM];
  const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4, lrow = t & 127;
  const bool yvalid = row0 + lrow < nq, xvalid = col0 + lrow < nd;
  const float* Yrow = Y + int64_t(row0 + (yvalid ? lrow : 0)) * ldY;
  const float* Xrow = X + int64_t(col0 + (xvalid ? lrow : 0)) * ldX;
  float acc[8][8];
  tile_scores<VEC>(acc, smem, Yrow, yvalid, Xrow, xvalid, d);
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int i = row0 + sub_idx(ty, a), j = col0 + sub_idx(tx, b);
      if (i < nq && j < nd) S[int64_t(i) * ldS + j] = acc[a][b];
    }
}

bool vec_ok(const float* A, int64_t ld, int d) {
  return (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (ld % 4 == 0) && (d % 4 == 0);
}

template <int NR, int NC>
void launch(const NNProblem& P, bool vec, dim3 grid, cudaStream_t st) {
  if (vec)
    nn_ffma_kernel<NR, NC, true><<<grid, NT, 0, st>>>(P);
  else
    nn_ffma_kernel<NR, NC, false><<<grid, NT, 0, st>>>(P);
}

}  // namespace

int nn_ffma_launch(const NNProblem& P, cudaStream_t st) {
  if (P.n_pairs <= 0 || P.total_q <= 0) return DM_OK;
  const bool vec = vec_ok(P.Y, P.ldY, P.d_fast) && vec_ok(P.X, P.ldX, P.d_fast);
  const int64_t nblk = int64_t(P.n_pairs) * P.max_rt;
  if (nblk > 0x7fffffffLL) DM_FAIL(DM_ERR_BADARG, "too many row tiles (%lld)", (long long)nblk);
  dim3 grid((unsigned)nblk);
  const int key = P.n_row * 10 + P.n_col;
  switch (key) {
    case 0 * 10 + 1: launch<0, 1>(P, vec, grid, st); break;
    case 0 * 10 + 2: launch<0, 2>(P, vec, grid, st); break;
    case 1 * 10 + 0: launch<1, 0>(P, vec, grid, st); break;
    case 1 * 10 + 1: launch<1, 1>(P, vec, grid, st); break;
    case 1 * 10 + 2: launch<1, 2>(P, vec, grid, st); break;
    case 2 * 10 + 0: launch<2, 0>(P, vec, grid, st); break;
    case 2 * 10 + 1: launch<2, 1>(P, vec, grid, st); break;
    case 2 * 10 + 2: launch<2, 2>(P, vec, grid, st); break;
    default: DM_FAIL(DM_ERR_BADARG, "unsupported epilogue combination %d row / %d col", P.n_row, P.n_col);
  }
  DM_LAUNCH_OK("nn_ffma_kernel");
  return DM_OK;
}

int nn_ffma_debug_scores(const float* Y, int64_t ldY, int nq, const float* X, int64_t ldX, int ndb, int d, float* S,
                         int64_t ldS, cudaStream_t st) {
  if (nq <= 0 || ndb <= 0) return DM_OK;
  dim3 grid((ndb + BN - 1) / BN, (nq + BM - 1) / BM);
  if (vec_ok(Y, ldY, d) && vec_ok(X, ldX, d))
    nn_ffma_scores_kernel<true><<<grid, NT, 0, st>>>(Y, ldY, nq, X, ldX, ndb, d, S, ldS);
  else
    nn_ffma_scores_kernel<false><<<grid, NT, 0, st>>>(Y, ldY, nq, X, ldX, ndb, d, S, ldS);
  DM_LAUNCH_OK("nn_ffma_scores_kernel");
  return DM_OK;
}

}  // namespace dm
