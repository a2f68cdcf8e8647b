// Stages shared by both score engines: operand preparation (norms, scale/bias, error-bound maxima),
// column finalisation, the float64 near-tie re-evaluation, and the driver that strings them together.
#include <stdarg.h>
#include <stdio.h>
#include <cuda_bf16.h>
#include <cooperative_groups.h>
#include "dm_internal.cuh"

namespace dm {

// ---------------------------------------------------------------- error string
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        n <= 0)
      n = 148;
  }
  return n;
}

namespace {

__device__ __forceinline__ int find_pair(const int64_t* off, int n_pairs, int64_t row) {
  int lo = 0, hi = n_pairs;  // off[lo] <= row < off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= row)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

struct SpecArr {
  SideEpiSpec s[kMaxEpi];
};

// one warp per row: |row| in float64, then the epilogue scale/bias derived from it
template <typename T>
__global__ void __launch_bounds__(256)
    prep_side_kernel(const T* __restrict__ M, int64_t ld, const int64_t* __restrict__ off, int n_pairs, int64_t total,
                     int d, float* __restrict__ norm_out, SpecArr specs, int n_specs, __nv_bfloat16* __restrict__ hi,
                     __nv_bfloat16* __restrict__ lo, __nv_bfloat16* __restrict__ lo2, int kp, int need_sq64) {
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= total) return;
  const T* r = M + row * ld;
  double s = 0.0;
  float s32 = 0.f;
  typedef RowVec<T> RV;
  constexpr int W = RV::W, U = 4;
  const bool vec_ok = (ld % W) == 0 && (reinterpret_cast<uintptr_t>(M) & 15) == 0;
  if (hi && vec_ok) {
    // split representation for the tensor-core engine: v = hi + lo + O(2^-18 |v|), zero-padded to kp columns.
    // 16-byte loads, U of them in flight per lane before the first use; packed bf16 stores.
    __nv_bfloat16* h = hi + row * kp;
    __nv_bfloat16* l = lo + row * kp;
    for (int k0 = lane * W; k0 < kp; k0 += 32 * W * U) {
      typename RV::V buf[U];
      bool full[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = k0 + u * 32 * W;
        full[u] = k + W <= d;
        if (full[u]) buf[u] = __ldg(reinterpret_cast<const typename RV::V*>(r + k));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = k0 + u * 32 * W;
        if (k >= kp) continue;
        double v[W];
        float vf[W];
#pragma unroll
        for (int w = 0; w < W; ++w) {
          T e = T(0);
          if (full[u])
            e = reinterpret_cast<const T*>(&buf[u])[w];
          else if (k + w < d)
            e = r[k + w];
          vf[w] = float(e);
          v[w] = (sizeof(T) == 8 || need_sq64) ? double(e) : 0.0;
        }
        unsigned short bh[W], bl[W], bl2[W];
#pragma unroll
        for (int w = 0; w < W; ++w) {
          // fp32 split: hi = bf16(x), lo = bf16(x - hi) (the difference is exact in fp32).  Conversions between fp32
          // and fp64 are slow-pipe instructions, so the float64 sum of squares is formed only when an epilogue needs
          // it (cosine scale / Euclidean bias); otherwise an fp32 sum feeds the (rounded-up) error bound.
          const float x = sizeof(T) == 8 ? float(v[w]) : vf[w];
          if (sizeof(T) == 8 || need_sq64)
            s = fma(v[w], v[w], s);
          else
            s32 = fmaf(x, x, s32);
          const __nv_bfloat16 vh = __float2bfloat16_rn(x);
          const float r1 = x - __bfloat162float(vh);
          const __nv_bfloat16 vl = __float2bfloat16_rn(r1);
          bh[w] = __bfloat16_as_ushort(vh), bl[w] = __bfloat16_as_ushort(vl);
          bl2[w] = __bfloat16_as_ushort(__float2bfloat16_rn(r1 - __bfloat162float(vl)));
        }
        store_bf16_vec<W>(h + k, bh);
        store_bf16_vec<W>(l + k, bl);
        if (lo2) store_bf16_vec<W>(lo2 + row * kp + k, bl2);
      }
    }
  } else if (hi) {
    __nv_bfloat16* h = hi + row * kp;
    __nv_bfloat16* l = lo + row * kp;
    for (int k = lane; k < kp; k += 32) {
      const double v = k < d ? double(r[k]) : 0.0;
      s = fma(v, v, s);
      const __nv_bfloat16 vh = __float2bfloat16_rn(float(v));
      h[k] = vh;
      const double r1 = v - double(__bfloat162float(vh));
      const __nv_bfloat16 vl = __float2bfloat16_rn(float(r1));
      l[k] = vl;
      if (lo2) lo2[row * kp + k] = __float2bfloat16_rn(float(r1 - double(__bfloat162float(vl))));
    }
  } else {
    for (int k = lane; k < d; k += 32) {
      const double v = double(r[k]);
      s = fma(v, v, s);
    }
  }
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, sh);
    s32 += __shfl_xor_sync(0xffffffffu, s32, sh);
  }
  if (lane != 0) return;
  // the fp32 sum (only used when no epilogue needs the float64 one) carries a relative error <= (d + 32) 2^-24
  if (s == 0.0 && s32 > 0.f) s = double(s32) * (1.0 + (double(d) + 32.0) * 5.9604644775390625e-08);
  const double nrm = sqrt(s);
  const float nf = __double2float_ru(nrm) * 1.0000005f;
  norm_out[row] = nf;
  for (int e = 0; e < n_specs; ++e) {
    const SideEpiSpec& S = specs.s[e];
    double sc = 1.0, bi = 0.0;
    if (S.scale_mode == DM_SCALE_ARRAY)
      sc = S.scale[row];
    else if (S.scale_mode == DM_SCALE_INVNORM)
      sc = 1.0 / fmax(nrm, 1e-30);
    if (S.bias_mode == DM_BIAS_ARRAY)
      bi = S.bias[row];
    else if (S.bias_mode == DM_BIAS_NEG_HALF_SQNORM)
      bi = -0.5 * s;
    S.sd[row] = sc;
    S.bd[row] = bi;
    const float scf = float(sc), bif = float(bi);
    S.sf[row] = scf;
    S.bf[row] = bif;
  }
}

// per pair and epilogue: G = max_j |row_j| |scale_j| and Bm = max_j |bias_j| (the error-bound maxima), one CTA each
__global__ void __launch_bounds__(256)
    pair_max_kernel(const int64_t* __restrict__ off, const float* __restrict__ norm, SpecArr specs, int n_pairs) {
  const int p = blockIdx.x, e = blockIdx.y;
  const SideEpiSpec& S = specs.s[e];
  float g = 0.f, b = 0.f;
  for (int64_t r = off[p] + threadIdx.x; r < off[p + 1]; r += blockDim.x) {
    g = fmaxf(g, norm[r] * fabsf(S.sf[r]) * 1.0000005f);
    b = fmaxf(b, fabsf(S.bf[r]));
  }
  __shared__ float sg[8], sb[8];
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) {
    g = fmaxf(g, __shfl_xor_sync(0xffffffffu, g, sh));
    b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, sh));
  }
  if ((threadIdx.x & 31) == 0) sg[threadIdx.x >> 5] = g, sb[threadIdx.x >> 5] = b;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) g = fmaxf(g, sg[w]), b = fmaxf(b, sb[w]);
    S.G[p] = g;
    S.Bm[p] = b;
  }
}

// A side served from a mesh bank: pair p's rows are bank rows in[p] .. in[p] + n.  Copies the bank's row norms into the
// batch packing and fills the epilogues' scale / bias arrays (scale arrays are indexed by bank rows; a DM_BIAS_ARRAY bias is
// a placeholder that a hook overwrites).  Same values as prep_side_kernel writes for these modes.
__global__ void __launch_bounds__(256)
    bank_side_rows_kernel(const int64_t* __restrict__ in, const int64_t* __restrict__ off, const float* __restrict__ norm_bank,
                          float* __restrict__ norm_out, SpecArr specs, int n_specs) {
  const int p = blockIdx.x;
  const int64_t s0 = in[p], r0 = off[p];
  const int n = int(off[p + 1] - r0);
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
    norm_out[r0 + i] = norm_bank[s0 + i];
    for (int e = 0; e < n_specs; ++e) {
      const SideEpiSpec& S = specs.s[e];
      const double sc = S.scale_mode == DM_SCALE_ARRAY ? S.scale[s0 + i] : 1.0;
      S.sd[r0 + i] = sc;
      S.bd[r0 + i] = 0.0;
      S.sf[r0 + i] = float(sc);
      S.bf[r0 + i] = 0.f;
    }
  }
}

__global__ void __launch_bounds__(256) col_finalize_kernel(const NNProblem P) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t per_epi = int64_t(P.n_pairs) * P.max_db;
  if (idx >= per_epi * P.n_col) return;
  const int c = int(idx / per_epi);
  const int64_t rem = idx % per_epi;
  const int p = int(rem / P.max_db), j = int(rem % P.max_db);
  const int64_t d0 = P.db_off[p];
  const int nd = int(P.db_off[p + 1] - d0);
  if (j >= nd) return;
  const int nq = int(P.q_off[p + 1] - P.q_off[p]);
  const int nrt = (nq + P.rt_rows - 1) / P.rt_rows;
  const Top3* part = P.col_partial + ((int64_t(c) * P.n_pairs + p) * P.max_rt) * P.max_db + j;
  Top3 m = top3_init();
  for (int rt = 0; rt < nrt; ++rt) top3_merge(m, part[int64_t(rt) * P.max_db]);
  emit_result(P, P.col[c], true, c, p, d0 + j, j, P.norm_db[d0 + j], m);
}

// ---------------------------------------------------------------- float64 re-evaluation
constexpr int RC_THREADS = 256;

// Full float64 re-evaluation of one flagged result: argmax_j dot(vec, mat[j]) * sc[j] + bi[j] over all n candidate
// rows, lowest index on ties.  A thread-block CLUSTER of kScanCluster CTAs owns one flagged result: CTA r scans the
// r-th slice of the candidate rows (one warp per row, four rows in flight per warp, lane-strided partial sums + xor
// tree -- the same summation order as the two-candidate kernel, so both paths agree bit for bit), the per-CTA
// winners are combined by rank 0 through distributed shared memory.
constexpr int kScanCluster = 8;

template <typename TV, typename TM>
__device__ void scan64_slice(double* vec, double* sbest, int* sidx, const TV* __restrict__ v, const TM* __restrict__ mat,
                             int64_t ldm, int j_beg, int j_end, int d, const double* __restrict__ sc,
                             const double* __restrict__ bi) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int k = t; k < d; k += RC_THREADS) vec[k] = double(v[k]);
  __syncthreads();
  double best = -INFINITY;
  int besti = 0x7fffffff;
  constexpr int NW = RC_THREADS / 32, U = 4;
  for (int j0 = j_beg + warp; j0 < j_end; j0 += NW * U) {  // a warp walks its rows in ascending order
    double s[U];
    const TM* __restrict__ r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      s[u] = 0.0;
      r[u] = mat + int64_t(min(j0 + u * NW, j_end - 1)) * ldm;
    }
#pragma unroll 4
    for (int k = lane; k < d; k += 32) {
      const double x = vec[k];
#pragma unroll
      for (int u = 0; u < U; ++u) s[u] = fma(x, double(r[u][k]), s[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) s[u] += __shfl_xor_sync(0xffffffffu, s[u], sh);
      const int j = j0 + u * NW;
      if (j < j_end) {
        const double val = __dadd_rn(__dmul_rn(s[u], sc[j]), bi[j]);  // same two roundings as numpy's S *= s; S += b
        if (val > best) {
          best = val;
          besti = j;
        }
      }
    }
  }
  if (lane == 0) {
    sbest[1 + warp] = best;
    sidx[1 + warp] = besti;
  }
  __syncthreads();
  if (t == 0) {
    for (int w = 1; w < NW; ++w)
      if (sbest[1 + w] > best || (sbest[1 + w] == best && sidx[1 + w] < besti)) {
        best = sbest[1 + w];
        besti = sidx[1 + w];
      }
    sbest[0] = best;  // slot 0 = this CTA's winner, read by rank 0 of the cluster
    sidx[0] = besti;
  }
}

template <typename TY, typename TX>
__global__ void __cluster_dims__(kScanCluster, 1, 1) __launch_bounds__(RC_THREADS) recheck_kernel(const NNProblem P) {
  namespace cg = cooperative_groups;
  extern __shared__ double vec[];
  __shared__ double sbest[1 + RC_THREADS / 32];
  __shared__ int sidx[1 + RC_THREADS / 32];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const unsigned cid = blockIdx.x / kScanCluster, ncl = gridDim.x / kScanCluster;
  const unsigned count = P.counters[3];
  const TY* Y = static_cast<const TY*>(P.Y64);
  const TX* X = static_cast<const TX*>(P.X64);
  for (unsigned f = cid; f < count; f += ncl) {  // uniform over the cluster
    const FlagEntry e = P.flags[P.flag_cap - 1 - int64_t(f)];
    const int p = e.pair, epi = e.epi & 255;
    const bool is_col = (e.epi & 256) != 0;
    const int64_t q0 = P.q_off[p], d0 = P.db_off[p];
    const int64_t qi0 = P.q_in[p], di0 = P.db_in[p];  // operand rows (mesh bank) vs batch-packed per-row arrays
    const int nq = int(P.q_off[p + 1] - q0), nd = int(P.db_off[p + 1] - d0);
    const int n = is_col ? nq : nd;
    const int per = (n + kScanCluster - 1) / kScanCluster;
    const int j_beg = min(n, int(rank) * per), j_end = min(n, j_beg + per);
    const EpiDev& E = is_col ? P.col[epi] : P.row[epi];
    if (!is_col)
      scan64_slice<TY, TX>(vec, sbest, sidx, Y + (qi0 + e.local) * P.ldY64, X + di0 * P.ldX64, P.ldX64, j_beg, j_end, P.d,
                           E.sd + d0, E.bd + d0);
    else
      scan64_slice<TX, TY>(vec, sbest, sidx, X + (di0 + e.local) * P.ldX64, Y + qi0 * P.ldY64, P.ldY64, j_beg, j_end, P.d,
                           E.sd + q0, E.bd + q0);
    cluster.sync();
    if (rank == 0 && threadIdx.x == 0) {
      double best = sbest[0];
      int besti = sidx[0];
      for (unsigned r = 1; r < kScanCluster; ++r) {  // ascending slices: strict '>' keeps the lowest index
        const double ov = *cluster.map_shared_rank(&sbest[0], r);
        const int oi = *cluster.map_shared_rank(&sidx[0], r);
        if (ov > best) {
          best = ov;
          besti = oi;
        }
      }
      store_index(E.out, (is_col ? d0 : q0) + e.local, besti == 0x7fffffff ? 0 : besti, P.i64_out != 0);
    }
    cluster.sync();  // the winners may be overwritten only after rank 0 has read them
  }
}

// Two-candidate re-evaluation: one warp per flagged result computes the two float64 scores with the same
// rounding sequence for both (so exact duplicates tie exactly and resolve to the lower index).
template <typename TV, typename TM>
__device__ __forceinline__ double warp_dot64(const TV* __restrict__ v, const TM* __restrict__ r, int d, int lane) {
  double s = 0.0;
  for (int k = lane; k < d; k += 32) s = fma(double(v[k]), double(r[k]), s);
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
  return s;
}

template <typename TY, typename TX>
__global__ void __launch_bounds__(256) recheck_cand_kernel(const NNProblem P) {
  const unsigned count = P.counters[0];
  const int lane = threadIdx.x & 31;
  const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarp = gridDim.x * (blockDim.x >> 5);
  const TY* Y = static_cast<const TY*>(P.Y64);
  const TX* X = static_cast<const TX*>(P.X64);
  for (unsigned f = warp; f < count; f += nwarp) {
    const FlagEntry e = P.flags[f];
    const int p = e.pair, epi = e.epi & 255;
    const bool is_col = (e.epi & 256) != 0;
    const int64_t q0 = P.q_off[p], d0 = P.db_off[p];
    const int64_t qi0 = P.q_in[p], di0 = P.db_in[p];
    const int lo = min(e.c1, e.c2), hi = max(e.c1, e.c2);
    double vlo, vhi;
    const EpiDev& E = is_col ? P.col[epi] : P.row[epi];
    if (!is_col) {
      const TY* y = Y + (qi0 + e.local) * P.ldY64;
      vlo = warp_dot64(y, X + (di0 + lo) * P.ldX64, P.d, lane);
      vhi = warp_dot64(y, X + (di0 + hi) * P.ldX64, P.d, lane);
      vlo = __dadd_rn(__dmul_rn(vlo, E.sd[d0 + lo]), E.bd[d0 + lo]);
      vhi = __dadd_rn(__dmul_rn(vhi, E.sd[d0 + hi]), E.bd[d0 + hi]);
    } else {
      const TX* x = X + (di0 + e.local) * P.ldX64;
      vlo = warp_dot64(x, Y + (qi0 + lo) * P.ldY64, P.d, lane);
      vhi = warp_dot64(x, Y + (qi0 + hi) * P.ldY64, P.d, lane);
      vlo = __dadd_rn(__dmul_rn(vlo, E.sd[q0 + lo]), E.bd[q0 + lo]);
      vhi = __dadd_rn(__dmul_rn(vhi, E.sd[q0 + hi]), E.bd[q0 + hi]);
    }
    if (lane == 0) store_index(E.out, (is_col ? d0 : q0) + e.local, (vhi > vlo) ? hi : lo, P.i64_out != 0);
  }
}

}  // namespace

int nn_prep_side(const void* M, int is_double, int64_t ld, const int64_t* off, int n_pairs, int64_t total, int d,
                 float* norm_out, const SideEpiSpec* specs, int n_specs, void* hi_v, void* lo_v, void* lo2_v, int kp,
                 cudaStream_t st) {
  __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(hi_v);
  __nv_bfloat16* lo = static_cast<__nv_bfloat16*>(lo_v);
  __nv_bfloat16* lo2 = static_cast<__nv_bfloat16*>(lo2_v);
  if (total <= 0) return DM_OK;
  SpecArr arr;
  for (int e = 0; e < kMaxEpi; ++e) arr.s[e] = specs && e < n_specs ? specs[e] : SideEpiSpec{};
  const int wpb = 8;
  const unsigned grid = unsigned((total + wpb - 1) / wpb);
  int need_sq64 = hi ? 0 : 1;  // the scalar (CUDA-core engine) path always sums in float64
  for (int e = 0; e < n_specs; ++e)
    if (specs[e].scale_mode == DM_SCALE_INVNORM || specs[e].bias_mode == DM_BIAS_NEG_HALF_SQNORM) need_sq64 = 1;
  if (is_double)
    prep_side_kernel<double><<<grid, wpb * 32, 0, st>>>(static_cast<const double*>(M), ld, off, n_pairs, total, d,
                                                        norm_out, arr, n_specs, hi, lo, lo2, kp, need_sq64);
  else
    prep_side_kernel<float><<<grid, wpb * 32, 0, st>>>(static_cast<const float*>(M), ld, off, n_pairs, total, d,
                                                       norm_out, arr, n_specs, hi, lo, lo2, kp, need_sq64);
  DM_LAUNCH_OK("prep_side_kernel");
  if (n_specs > 0 && n_pairs > 0) {
    pair_max_kernel<<<dim3(unsigned(n_pairs), unsigned(n_specs)), 256, 0, st>>>(off, norm_out, arr, n_pairs);
    DM_LAUNCH_OK("pair_max_kernel");
  }
  return DM_OK;
}

int nn_bank_side_rows(const NNBankSide& B, const int64_t* off, int n_pairs, int max_n, float* norm_out,
                      const SideEpiSpec* specs, int n_specs, cudaStream_t st) {
  if (n_pairs <= 0 || max_n <= 0) return DM_OK;
  SpecArr arr;
  for (int e = 0; e < kMaxEpi; ++e) arr.s[e] = specs && e < n_specs ? specs[e] : SideEpiSpec{};
  for (int e = 0; e < n_specs; ++e)
    if (specs[e].scale_mode == DM_SCALE_INVNORM || specs[e].bias_mode == DM_BIAS_NEG_HALF_SQNORM)
      DM_FAIL(DM_ERR_UNSUPPORTED, "mesh-bank side: epilogue %d derives its scale / bias from the rows", e);
  bank_side_rows_kernel<<<dim3(unsigned(n_pairs), unsigned((max_n + 255) / 256)), 256, 0, st>>>(B.in, off, B.norm, norm_out,
                                                                                                arr, n_specs);
  DM_LAUNCH_OK("bank_side_rows_kernel");
  if (n_specs > 0) {
    pair_max_kernel<<<dim3(unsigned(n_pairs), unsigned(n_specs)), 256, 0, st>>>(off, norm_out, arr, n_pairs);
    DM_LAUNCH_OK("pair_max_kernel");
  }
  return DM_OK;
}

int nn_col_finalize(const NNProblem& P, cudaStream_t st) {
  if (P.n_col == 0 || P.total_db <= 0) return DM_OK;
  const int64_t n = int64_t(P.n_pairs) * P.max_db * P.n_col;
  col_finalize_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(P);
  DM_LAUNCH_OK("col_finalize_kernel");
  return DM_OK;
}

int nn_recheck(const NNProblem& P, cudaStream_t st) {
  if (P.flags == nullptr) return DM_OK;
  const size_t shm = sizeof(double) * size_t(P.d);
  if (shm > 200 * 1024) DM_FAIL(DM_ERR_UNSUPPORTED, "inner dimension %d too large for the float64 re-evaluation", P.d);
  const int grid = num_sms() / kScanCluster * kScanCluster * 2;  // a multiple of the cluster size
#define DM_RC(TY, TX)                                                                                         \
  do {                                                                                                        \
    if (shm > 48 * 1024)                                                                                      \
      DM_CUDA_OK(cudaFuncSetAttribute(recheck_kernel<TY, TX>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(shm))); \
    recheck_kernel<TY, TX><<<grid, RC_THREADS, shm, st>>>(P);                                                 \
  } while (0)
  if (P.y64_is_double && P.x64_is_double)
    DM_RC(double, double);
  else if (P.y64_is_double)
    DM_RC(double, float);
  else if (P.x64_is_double)
    DM_RC(float, double);
  else
    DM_RC(float, float);
#undef DM_RC
  DM_LAUNCH_OK("recheck_kernel");
  const int cgrid = num_sms() * 4;
  if (P.y64_is_double && P.x64_is_double)
    recheck_cand_kernel<double, double><<<cgrid, 256, 0, st>>>(P);
  else if (P.y64_is_double)
    recheck_cand_kernel<double, float><<<cgrid, 256, 0, st>>>(P);
  else if (P.x64_is_double)
    recheck_cand_kernel<float, double><<<cgrid, 256, 0, st>>>(P);
  else
    recheck_cand_kernel<float, float><<<cgrid, 256, 0, st>>>(P);
  DM_LAUNCH_OK("recheck_cand_kernel");
  return DM_OK;
}

// ---------------------------------------------------------------- driver
namespace {
int pick_rt_rows(int /*d*/, int /*flags*/) { return kFfmaRowTile; }  // both engines tile 128 query rows per CTA

NNLayout carve(void* ws, int n_pairs, int64_t total_q, int64_t total_db, int max_q, int max_db, int d, int n_row,
               int n_col, int flags) {
  Carver c(ws);
  NNLayout L;
  L.counters = c.take<unsigned int>(64);
  L.norm_q = c.take<float>(total_q);
  L.norm_db = c.take<float>(total_db);
  for (int e = 0; e < n_row; ++e) {
    L.row[e].sf = c.take<float>(total_db);
    L.row[e].bf = c.take<float>(total_db);
    L.row[e].sd = c.take<double>(total_db);
    L.row[e].bd = c.take<double>(total_db);
    L.row[e].G = c.take<float>(n_pairs);
    L.row[e].Bm = c.take<float>(n_pairs);
  }
  for (int e = 0; e < n_col; ++e) {
    L.col[e].sf = c.take<float>(total_q);
    L.col[e].bf = c.take<float>(total_q);
    L.col[e].sd = c.take<double>(total_q);
    L.col[e].bd = c.take<double>(total_q);
    L.col[e].G = c.take<float>(n_pairs);
    L.col[e].Bm = c.take<float>(n_pairs);
  }
  const int rt_rows = pick_rt_rows(d, flags);
  const int max_rt = (max_q + rt_rows - 1) / rt_rows;
  L.col_partial = c.take<Top3>(size_t(n_col) * n_pairs * max_rt * max_db);
  L.flag_cap = (flags & DM_NO_RECHECK) ? 0 : int64_t(n_row) * total_q + int64_t(n_col) * total_db;
  L.flags = c.take<FlagEntry>(size_t(L.flag_cap));
  L.yh = L.yl = L.xh = L.xl = L.yl2 = L.xl2 = nullptr;
  if (nn_use_tc(flags)) {
    const size_t kp = size_t(nn_tc_kp(d));
    const size_t nyq = (flags & kFlagBankQ) ? 0 : size_t(total_q), nxd = (flags & kFlagBankDb) ? 0 : size_t(total_db);
    L.yh = c.take<uint16_t>(nyq * kp);
    L.yl = c.take<uint16_t>(nyq * kp);
    L.xh = c.take<uint16_t>(nxd * kp);
    L.xl = c.take<uint16_t>(nxd * kp);
    if (flags & kFlagSplit3) {
      L.yl2 = c.take<uint16_t>(nyq * kp);
      L.xl2 = c.take<uint16_t>(nxd * kp);
    }
  }
  L.bytes = c.bytes();
  return L;
}
}  // namespace

size_t nn_workspace_bytes(int n_pairs, int64_t total_q, int64_t total_db, int max_q, int max_db, int d, int n_row,
                          int n_col, int flags) {
  return carve(nullptr, n_pairs, total_q, total_db, max_q, max_db, d, n_row, n_col, flags).bytes;
}

int nn_run(const NNRequest& R, void* ws, size_t ws_bytes, cudaStream_t st, NNSplits* splits) {
  if (R.n_pairs < 0 || R.d <= 0 || R.total_q < 0 || R.total_db < 0 || R.max_q < 0 || R.max_db < 0)
    DM_FAIL(DM_ERR_BADARG, "negative size");
  if (R.n_row < 0 || R.n_row > kMaxEpi || R.n_col < 0 || R.n_col > kMaxEpi || R.n_row + R.n_col == 0)
    DM_FAIL(DM_ERR_BADARG, "need 1..%d row and/or column epilogues", kMaxEpi);
  if (R.n_pairs == 0 || (R.total_q == 0 && R.total_db == 0)) return DM_OK;
  const bool tc = nn_use_tc(R.flags);
  // the tensor-core engine reads the float64 originals when they are given; the CUDA-core engine needs fp32 operands
  const bool haveY = R.Y || (tc && R.Y64), haveX = R.X || (tc && R.X64);
  if ((R.total_q > 0 && !haveY) || (R.total_db > 0 && !haveX) || !R.q_off || !R.db_off) DM_FAIL(DM_ERR_BADARG, "null operand");
  if ((R.Y && R.ldY < R.d) || (R.X && R.ldX < R.d) || (R.Y64 && R.ldY64 < R.d) || (R.X64 && R.ldX64 < R.d) ||
      (!tc && R.d_fast > 0 && (R.d_fast < R.d || R.ldY < R.d_fast || R.ldX < R.d_fast)))
    DM_FAIL(DM_ERR_BADARG, "leading dimension smaller than d");
  if ((R.flags & DM_ENGINE_FFMA) && (R.flags & DM_ENGINE_TC)) DM_FAIL(DM_ERR_BADARG, "both engines forced");
  for (int e = 0; e < R.n_row + R.n_col; ++e) {
    const dm_nn_epi& E = e < R.n_row ? R.row[e] : R.col[e - R.n_row];
    if (!E.out && (e < R.n_row ? R.total_q : R.total_db) > 0) DM_FAIL(DM_ERR_BADARG, "epilogue %d has no output", e);
    if (E.scale_mode == DM_SCALE_ARRAY && !E.scale) DM_FAIL(DM_ERR_BADARG, "epilogue %d: scale array missing", e);
    if (E.bias_mode == DM_BIAS_ARRAY && !E.bias) DM_FAIL(DM_ERR_BADARG, "epilogue %d: bias array missing", e);
    if (E.scale_mode < 0 || E.scale_mode > 2 || E.bias_mode < 0 || E.bias_mode > 2)
      DM_FAIL(DM_ERR_BADARG, "epilogue %d: bad mode", e);
  }
  if (!ws) DM_FAIL(DM_ERR_WORKSPACE, "workspace is null");
  if (reinterpret_cast<uintptr_t>(ws) % 256) DM_FAIL(DM_ERR_ALIGN, "workspace must be 256-byte aligned");
  if ((R.bank_q || R.bank_db) && !tc) DM_FAIL(DM_ERR_UNSUPPORTED, "mesh-bank operands need the tensor-core engine");
  if ((R.bank_q && R.skip_prep_y) || ((R.bank_q || R.bank_db) && (R.flags & DM_SKIP_PREP)))
    DM_FAIL(DM_ERR_BADARG, "mesh-bank operands cannot be combined with a skipped preparation");
  NNLayout L = carve(ws, R.n_pairs, R.total_q, R.total_db, R.max_q, R.max_db, R.d, R.n_row, R.n_col,
                     R.flags | (R.bank_q ? kFlagBankQ : 0) | (R.bank_db ? kFlagBankDb : 0));
  if (L.bytes > ws_bytes)
    DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", L.bytes, ws_bytes);
  // operands of a bank side: the splits prepared once per mesh (the hooks and the engines below see them through L)
  if (R.bank_q) {
    L.yh = static_cast<uint16_t*>(const_cast<void*>(R.bank_q->hi)), L.yl = static_cast<uint16_t*>(const_cast<void*>(R.bank_q->lo));
    L.yl2 = static_cast<uint16_t*>(const_cast<void*>(R.bank_q->lo2));
  }
  if (R.bank_db) {
    L.xh = static_cast<uint16_t*>(const_cast<void*>(R.bank_db->hi)), L.xl = static_cast<uint16_t*>(const_cast<void*>(R.bank_db->lo));
    L.xl2 = static_cast<uint16_t*>(const_cast<void*>(R.bank_db->lo2));
  }

  DM_CUDA_OK(cudaMemsetAsync(L.counters, 0, 64 * sizeof(unsigned int), st));

  NNProblem P{};
  P.Y = R.Y, P.ldY = R.ldY, P.X = R.X, P.ldX = R.ldX;
  P.Y64 = R.Y64 ? static_cast<const void*>(R.Y64) : R.Y, P.ldY64 = R.Y64 ? R.ldY64 : R.ldY, P.y64_is_double = R.Y64 != nullptr;
  P.X64 = R.X64 ? static_cast<const void*>(R.X64) : R.X, P.ldX64 = R.X64 ? R.ldX64 : R.ldX, P.x64_is_double = R.X64 != nullptr;
  P.q_off = R.q_off, P.db_off = R.db_off;
  P.q_in = R.bank_q ? R.bank_q->in : R.q_off, P.db_in = R.bank_db ? R.bank_db->in : R.db_off;
  P.rows_q = R.bank_q ? R.bank_q->rows : R.total_q, P.rows_db = R.bank_db ? R.bank_db->rows : R.total_db;
  P.total_q = R.total_q, P.total_db = R.total_db, P.max_q = R.max_q, P.max_db = R.max_db;
  P.n_pairs = R.n_pairs, P.d = R.d, P.n_row = R.n_row, P.n_col = R.n_col;
  P.d_fast = R.d_fast > 0 ? R.d_fast : R.d;
  P.norm_q = L.norm_q, P.norm_db = L.norm_db;
  P.i64_out = (R.flags & DM_I64_OUT) ? 1 : 0;
  P.recheck_all = (R.flags & DM_RECHECK_ALL) ? 1 : 0;
  P.col_partial = L.col_partial;
  P.rt_rows = pick_rt_rows(R.d, R.flags);
  P.max_rt = (R.max_q + P.rt_rows - 1) / P.rt_rows;
  P.flags = (R.flags & DM_NO_RECHECK) ? nullptr : L.flags;
  P.flag_cap = L.flag_cap;
  P.counters = L.counters;
  P.kp = nn_tc_kp(R.d);
  P.col_trunc = 0.f;
  P.row_trunc = 0.f;
  {
    static const bool probe = [] { const char* e = getenv("DM_NN_PROBE"); return e && e[0] == '1'; }();
    P.probe_skip_epilogue = probe ? 1 : 0;
  }
  if (tc) {
    // split-bf16 truncation 3 * 2^-18 (+5%) plus one fp32 rounding (with 2x slack) per accumulated MMA
    P.eps = float(1.2e-5 + (3.0 * (P.kp / 16) + 2.0) * 2.384185791015625e-07);
    P.col_trunc = 3.9e-6f;  // 2^-18 (+2%): the column partials keep the row in the 5 low mantissa bits (nn_tc.cu)
  } else
    // fp32 FMA chain of length d (+ the fp32 rounding of float64 originals): gamma_d = d u / (1 - d u)
    P.eps = float((double(R.d) + 4.0) * 5.9604644775390625e-08 * 1.01);

  SideEpiSpec rs[kMaxEpi], cs[kMaxEpi];
  for (int e = 0; e < R.n_row; ++e) {
    rs[e] = SideEpiSpec{R.row[e].scale_mode, R.row[e].bias_mode, R.row[e].scale, R.row[e].bias, L.row[e].sf,
                        L.row[e].bf,         L.row[e].sd,        L.row[e].bd,    L.row[e].G,    L.row[e].Bm};
    P.row[e] = EpiDev{L.row[e].sf, L.row[e].bf, L.row[e].sd, L.row[e].bd, L.row[e].G, L.row[e].Bm, R.row[e].out,
                      R.row[e].scale_mode == DM_SCALE_NONE && R.row[e].bias_mode == DM_BIAS_NONE};
  }
  for (int e = 0; e < R.n_col; ++e) {
    cs[e] = SideEpiSpec{R.col[e].scale_mode, R.col[e].bias_mode, R.col[e].scale, R.col[e].bias, L.col[e].sf,
                        L.col[e].bf,         L.col[e].sd,        L.col[e].bd,    L.col[e].G,    L.col[e].Bm};
    P.col[e] = EpiDev{L.col[e].sf, L.col[e].bf, L.col[e].sd, L.col[e].bd, L.col[e].G, L.col[e].Bm, R.col[e].out,
                      R.col[e].scale_mode == DM_SCALE_NONE && R.col[e].bias_mode == DM_BIAS_NONE};
  }
  int rc;
  if (!(R.flags & DM_SKIP_PREP)) {
    // database side carries the row epilogues' scale/bias, query side the column epilogues'
    if (R.hooks && R.hooks->prep_x) {
      if ((rc = R.hooks->prep_x(R.hooks->ctx, L, P, R, st))) return rc;
    } else if (R.bank_db) {
      if ((rc = nn_bank_side_rows(*R.bank_db, R.db_off, R.n_pairs, R.max_db, L.norm_db, rs, R.n_row, st))) return rc;
    } else if ((rc = nn_prep_side(P.X64, P.x64_is_double, P.ldX64, R.db_off, R.n_pairs, R.total_db, R.d, L.norm_db, rs,
                                  R.n_row, L.xh, L.xl, L.xl2, P.kp, st)))
      return rc;
    if (R.hooks && R.hooks->prep_y) {
      if ((rc = R.hooks->prep_y(R.hooks->ctx, L, P, R, st))) return rc;
    } else if (R.bank_q) {
      if ((rc = nn_bank_side_rows(*R.bank_q, R.q_off, R.n_pairs, R.max_q, L.norm_q, cs, R.n_col, st))) return rc;
    } else if (!R.skip_prep_y &&
               (rc = nn_prep_side(P.Y64, P.y64_is_double, P.ldY64, R.q_off, R.n_pairs, R.total_q,
                                  R.y_prep_d > R.d ? R.y_prep_d : R.d, L.norm_q, cs, R.n_col, L.yh, L.yl, L.yl2, P.kp, st)))
      return rc;
    if (R.hooks && R.hooks->after_prep && (rc = R.hooks->after_prep(R.hooks->ctx, L, P, st))) return rc;
  }
  if (splits) *splits = NNSplits{L.yh, L.yl, L.yl2, L.xh, L.xl, L.xl2, P.kp};
  if (R.after_prep_event) DM_CUDA_OK(cudaEventRecord(R.after_prep_event, st));
  if (tc && nn_tc2_applicable(P.n_row, P.n_col, P.kp)) {
    P.row_trunc = P.col_trunc;  // packed keys on both sides
    if ((rc = nn_tc2_launch(P, L.yh, L.yl, L.xh, L.xl, st))) return rc;
  } else if (tc) {
    if (P.n_col == 0) P.row_trunc = P.col_trunc;  // row-only passes scan packed keys (nn_tc.cu)
    if ((rc = nn_tc_launch(P, L.yh, L.yl, L.xh, L.xl, nullptr, 0, st))) return rc;
  } else {
    if ((rc = nn_ffma_launch(P, st))) return rc;
  }
  if (R.flags & DM_SKIP_FINISH) return DM_OK;
  if ((rc = nn_col_finalize(P, st))) return rc;
  if (R.hooks && R.hooks->before_recheck && P.flags && (rc = R.hooks->before_recheck(R.hooks->ctx, L, P, st))) return rc;
  if ((rc = nn_recheck(P, st))) return rc;
  return DM_OK;
}

}  // namespace dm
