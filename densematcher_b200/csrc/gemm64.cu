#include "gemm64.cuh"

namespace dm {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct OpView {  // an operand resolved for one batch
  const double* d;
  const float* f;
  int64_t ld;
  int col0;
  int trans;
  const void* gather;
  int gather_i64;
  int64_t gbase;
  const double* kscale;  // already offset to this batch
};

__device__ __forceinline__ OpView resolve(const GemmOperand& o, int b) {
  OpView v;
  const int64_t row0 = o.off ? o.off[b] : 0;
  const int64_t base = (o.off ? 0 : int64_t(b) * o.batch_stride) + row0 * o.ld;
  v.d = o.d ? o.d + base : nullptr;
  v.f = o.f ? o.f + base : nullptr;
  v.ld = o.ld;
  v.col0 = o.col0;
  v.trans = o.trans;
  v.gather = o.gather;
  v.gather_i64 = o.gather_i64;
  v.gbase = o.gather_off ? o.gather_off[b] : 0;
  const int64_t kbase = o.gather_off ? o.gather_off[b] : row0;
  v.kscale = o.kscale ? o.kscale + kbase : nullptr;
  return v;
}

__device__ __forceinline__ double op_load(const OpView& v, int i, int k) {
  int64_t r, c;
  if (v.trans == 0) {
    r = i;
    c = v.col0 + k;
  } else {
    r = v.gather ? load_index(v.gather, v.gbase + k, v.gather_i64 != 0) : k;
    c = v.col0 + i;
  }
  double x = v.d ? v.d[r * v.ld + c] : double(v.f[r * v.ld + c]);
  if (v.trans == 1 && v.kscale) x *= v.kscale[k];
  return x;
}

__device__ __forceinline__ int ragged_k(const GemmOperand& o, int b) {
  if (o.trans != 1) return -1;
  if (o.gather_off) return int(o.gather_off[b + 1] - o.gather_off[b]);
  if (o.off) return int(o.off[b + 1] - o.off[b]);
  return -1;
}

__global__ void __launch_bounds__(256) gemm64_kernel(const GemmProblem P, int tiles_m, int tiles_n) {
  int bid = blockIdx.x;
  const int tn = bid % tiles_n;
  bid /= tiles_n;
  const int tm = bid % tiles_m;
  bid /= tiles_m;
  const int ks = bid % P.ksplit;
  const int b = bid / P.ksplit;

  const int M = (P.A.trans == 0 && P.A.off) ? int(P.A.off[b + 1] - P.A.off[b]) : P.M;
  const int N = (P.B.trans == 0 && P.B.off) ? int(P.B.off[b + 1] - P.B.off[b]) : P.N;
  int K = ragged_k(P.A, b);
  if (K < 0) K = ragged_k(P.B, b);
  if (K < 0) K = P.K;
  const int m0 = tm * TM, n0 = tn * TN;
  if (m0 >= M || n0 >= N) return;
  int kbeg = 0, kend = K;
  if (P.ksplit > 1) {
    kbeg = min(K, ks * P.kchunk);
    kend = min(K, kbeg + P.kchunk);
  }

  __shared__ double As[TK][TM + 1];
  __shared__ double Bs[TK][TN + 1];
  const OpView A = resolve(P.A, b), B = resolve(P.B, b);
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.0;

  for (int k0 = kbeg; k0 < kend; k0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = t + 256 * i;
      int mm, kk;
      if (A.trans == 0) {
        kk = e % TK;
        mm = e / TK;
      } else {
        mm = e % TM;
        kk = e / TM;
      }
      As[kk][mm] = (m0 + mm < M && k0 + kk < kend) ? op_load(A, m0 + mm, k0 + kk) : 0.0;
      int nn;
      if (B.trans == 0) {
        kk = e % TK;
        nn = e / TK;
      } else {
        nn = e % TN;
        kk = e / TN;
      }
      Bs[kk][nn] = (n0 + nn < N && k0 + kk < kend) ? op_load(B, n0 + nn, k0 + kk) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = As[k][ty * 4 + a];
#pragma unroll
      for (int c = 0; c < 4; ++c) bv[c] = Bs[k][tx * 4 + c];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = fma(av[a], bv[c], acc[a][c]);
    }
    __syncthreads();
  }

  double* C = P.C + int64_t(ks) * P.split_stride + (P.c_off ? P.c_off[b] * P.ldc : int64_t(b) * P.c_batch_stride);
  const double* cs = P.c_colscale ? P.c_colscale + (P.c_colscale_off ? P.c_colscale_off[b] : 0) : nullptr;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int m = m0 + ty * 4 + a;
    if (m >= M) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int n = n0 + tx * 4 + c;
      if (n >= N) continue;
      double v = P.alpha * acc[a][c];
      if (cs) v *= cs[n];
      C[int64_t(m) * P.ldc + n] = v;
    }
  }
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
  const int tiles_m = (P.maxM + TM - 1) / TM, tiles_n = (P.maxN + TN - 1) / TN;
  const int64_t nblk = int64_t(P.n_batch) * (P.ksplit > 0 ? P.ksplit : 1) * tiles_m * tiles_n;
  if (nblk > 0x7fffffffLL) DM_FAIL(DM_ERR_BADARG, "gemm64: grid too large");
  GemmProblem Q = P;
  if (Q.ksplit < 1) Q.ksplit = 1;
  gemm64_kernel<<<unsigned(nblk), 256, 0, st>>>(Q, tiles_m, tiles_n);
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
