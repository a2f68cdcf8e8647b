// Closed-form functional-map solve: rows of C decouple into k2 SPD systems of size n = k1 - 1
//     (w_d Abar Abar^T + w_l diag(Delta_i)) c = w_d Abar (B_i - c_i0 A_0)^T
// that share the Gram matrix and differ on the diagonal.  Replaces the L-BFGS-B loop of FunctionalMapping.fit
// (densematcher/pyFM/functional.py:352-487; energy optimize/base_functions.py:31-56, :79-102) when only the descriptor
// and Laplacian terms are active.
//
// Two kernels:
//   fmap_solve32_kernel  (default)  float32 Cholesky in shared memory + float64 iterative refinement.  The factor is
//       only a preconditioner: residuals r - M x are formed in float64 from the float64 Gram matrix, so the result
//       converges to the float64 solution (error ~ (cond n 2^-24)^2 after one step; the row systems have cond ~ 30).
//       Half the shared memory of a float64 factor => twice the resident systems per SM, and the FMA pipe at full rate.
//       Systems whose refinement does not contract (ill-conditioned, non-positive float32 pivot) are queued and
//   fmap_solve_kernel    float64 packed Cholesky (round 1), now also the fallback over that queue.
#include "dm_internal.cuh"
#include "gemm64.cuh"
#include "tc_ptx.cuh"

namespace dm {
namespace {

constexpr int kSolveNB = 4;  // columns per block step
constexpr int kSolveThreads = 128;

// 1 / sqrt(v) in float64 from the FP64 reciprocal-square-root estimate (MUFU.RSQ64H, ~2^-22) and two Newton steps
// (scripts/micro/fp64_latency.cu: 64 cycles of dependent latency against 173 for 1 / sqrt()).
__device__ __forceinline__ double rsqrt64(double v) {
  if (!(v > 1e-300 && v < 1e300)) return 1.0 / sqrt(v);
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
  const double h = 0.5 * v;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  return y;
}

// status words in the workspace (device ints): [0] a float64 pivot was not positive (the result holds NaN / garbage),
// [1] systems sent to the float64 fallback, [2] systems solved, [3] refinement steps taken in total
enum { kStBad = 0, kStFallback = 1, kStSystems = 2, kStRefine = 3 };

// ------------------------------------------------------------------------------------------------ float64 kernel
// w_d * (A A^T)[1:, 1:] of every pair as a packed lower triangle (row r starts at r (r + 1) / 2)
__global__ void __launch_bounds__(256)
    solve_pack_kernel(const double* __restrict__ AAt, double wd, int k1, int64_t stride, double* __restrict__ Lp,
                      const int* __restrict__ only_if) {
  if (only_if && *only_if == 0) return;  // fallback mode: nothing was queued
  const int n = k1 - 1, n_tri = n * (n + 1) / 2;
  const double* aat = AAt + int64_t(blockIdx.x) * k1 * k1;
  double* out = Lp + int64_t(blockIdx.x) * stride;
  for (int e = threadIdx.x; e < n_tri; e += blockDim.x) {
    int r = int((sqrtf(8.f * float(e) + 1.f) - 1.f) * 0.5f);
    while (r * (r + 1) / 2 > e) --r;
    while ((r + 1) * (r + 2) / 2 <= e) ++r;
    const int c = e - r * (r + 1) / 2;
    out[e] = wd * aat[int64_t(r + 1) * k1 + c + 1];
  }
  if (threadIdx.x == 0 && stride > n_tri) out[n_tri] = 0.0;
}

// One 128-thread CTA per system: left-looking Cholesky on a packed lower triangle in shared memory, four columns per
// step, right-hand side carried as an extra row, register-resident back substitution in warp 0.
// list == nullptr: system = blockIdx.x.  Otherwise the CTAs walk the queue list[0 .. *count - 1] (fallback mode).
template <int RPT>  // rows per thread: (n + 1) <= 128 * RPT
__global__ void __launch_bounds__(kSolveThreads, RPT == 1 ? 5 : 1)
    fmap_solve_kernel(const double* __restrict__ AAt, const double* __restrict__ BAt, const double* __restrict__ Lp,
                      int64_t lp_stride, const double* __restrict__ ev1, const double* __restrict__ ev2,
                      const double* __restrict__ c00, double wd, double wl, int k1, int k2, double* __restrict__ C,
                      int* __restrict__ status, const int* __restrict__ list, const int* __restrict__ count) {
  extern __shared__ __align__(16) double sm[];
  constexpr int RPL = 4 * RPT;  // unknowns per lane in the single-warp back substitution
  const int n = k1 - 1;
  const int t = threadIdx.x, lane = t & 31;
  const int n_tri = n * (n + 1) / 2;
  double* L = sm;                                        // row r starts at r (r + 1) / 2; row n = right-hand side
  double* Dbuf = sm + size_t(n + 1) * (n + 2) / 2 + 1;   // [2][4][4] accumulated diagonal blocks (double-buffered)
  double* Dorig = Dbuf + 2 * kSolveNB * kSolveNB;        // [2][4] the block's diagonal entries before elimination
  double* invd = Dorig + 2 * kSolveNB;                   // [n] reciprocals of the diagonal of L
  __shared__ double s_scale[kSolveThreads / 32];
  __shared__ __align__(8) unsigned long long s_bar;
  const uint32_t bar = tc::smem_u32(&s_bar);
  if (t == 0) {
    tc::mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int n_work = list ? *count : int(gridDim.x);
  uint32_t parity = 0;
  for (int f = blockIdx.x; f < n_work; f += gridDim.x, parity ^= 1) {
    const int sys = list ? list[f] : f;
    const int b = sys / k2, i = sys % k2;
    const double* aat = AAt + int64_t(b) * k1 * k1;
    const double* bat = BAt + int64_t(b) * k2 * k1;
    const double* l1 = ev1 + int64_t(b) * k1;
    const double* l2 = ev2 + int64_t(b) * k2;
    // the shared part of the matrix arrives by one bulk copy while the threads prepare the system-specific part
    if (t == 0) {
      const uint32_t bytes = uint32_t(lp_stride * sizeof(double));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic accesses of L precede the copy
      tc::mbar_expect_tx(bar, bytes);
      tc::tma_load_1d(tc::smem_u32(L), Lp + int64_t(b) * lp_stride, bytes, bar);
    }
    double scale = -INFINITY;
    for (int j = t; j < k1; j += kSolveThreads) scale = fmax(scale, l1[j]);
    for (int j = t; j < k2; j += kSolveThreads) scale = fmax(scale, l2[j]);
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) scale = fmax(scale, __shfl_xor_sync(0xffffffffu, scale, sh));
    if (lane == 0) s_scale[t >> 5] = scale;
    __syncthreads();
    scale = fmax(fmax(s_scale[0], s_scale[1]), fmax(s_scale[2], s_scale[3]));
    const double ci0 = (i == 0) ? c00[b] : 0.0;
    const double l2i = l2[i] / scale;
    double dg[RPT], rh[RPT];
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int c = t + kSolveThreads * q;
      dg[q] = rh[q] = 0.0;
      if (c < n) {
        const double df = l1[c + 1] / scale - l2i;
        dg[q] = wl * (df * df);
        rh[q] = wd * (bat[int64_t(i) * k1 + c + 1] - ci0 * aat[c + 1]);
      }
    }
    tc::mbar_wait(bar, parity);
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int c = t + kSolveThreads * q;
      if (c < n) {
        L[size_t(c) * (c + 1) / 2 + c] += dg[q];
        L[n_tri + c] = rh[q];
      }
    }
    __syncthreads();
    bool bad = false;
    int par = 0;
    for (int j0 = 0; j0 < n; j0 += kSolveNB, par ^= 1) {
      const int nb = min(kSolveNB, n - j0);
      // rows of this block step, thread-cyclic from j0: r = j0 + t + 128 q  (row n = right-hand side included)
      double acc[RPT][kSolveNB];
      const double* rowr[RPT];
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int r = j0 + t + kSolveThreads * q;
        const int rc = min(r, n);
        rowr[q] = L + size_t(rc) * (rc + 1) / 2;
#pragma unroll
        for (int c = 0; c < kSolveNB; ++c) acc[q][c] = (r <= n && c < nb && j0 + c <= min(r, n - 1)) ? rowr[q][j0 + c] : 0.0;
      }
      if (t < kSolveNB) {  // a_jj of the block (row j0 + t), before elimination
        const double own = t == 0 ? acc[0][0] : t == 1 ? acc[0][1] : t == 2 ? acc[0][2] : acc[0][3];
        Dorig[par * kSolveNB + t] = t < nb ? own : 1.0;
      }
      const double* rowc[kSolveNB];
#pragma unroll
      for (int c = 0; c < kSolveNB; ++c) {
        const int jc = min(j0 + c, n - 1);
        rowc[c] = L + size_t(jc) * (jc + 1) / 2;
      }
      if (j0 + (t & ~31) <= n) {  // warps whose rows all lie beyond the matrix skip the bulk (warp-uniform)
#pragma unroll 4
        for (int k = 0; k < j0; ++k) {
          double lc[kSolveNB];
#pragma unroll
          for (int c = 0; c < kSolveNB; ++c) lc[c] = rowc[c][k];
#pragma unroll
          for (int q = 0; q < RPT; ++q) {
            const double lr = rowr[q][k];
#pragma unroll
            for (int c = 0; c < kSolveNB; ++c) acc[q][c] = fma(-lr, lc[c], acc[q][c]);
          }
        }
      }
      double* D = Dbuf + par * kSolveNB * kSolveNB;
      if (t < kSolveNB) {
#pragma unroll
        for (int c = 0; c < kSolveNB; ++c) D[t * kSolveNB + c] = acc[0][c];
      }
      __syncthreads();
      double l[kSolveNB][kSolveNB], li[kSolveNB];
#pragma unroll
      for (int c = 0; c < kSolveNB; ++c) {
#pragma unroll
        for (int c2 = 0; c2 <= c; ++c2) {
          double v = D[c * kSolveNB + c2];
#pragma unroll
          for (int c3 = 0; c3 < c2; ++c3) v = fma(-l[c][c3], l[c2][c3], v);
          if (c2 == c) {
            // a pivot that cancelled to rounding level (n u a_jj) means the matrix is numerically singular
            if (c < nb && !(v > 1e-13 * fabs(Dorig[par * kSolveNB + c]))) bad = true;
            li[c] = c < nb ? rsqrt64(v) : 1.0;
            l[c][c] = v * li[c];
          } else {
            l[c][c2] = c < nb ? v * li[c2] : 0.0;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < kSolveNB; ++c)
        if (t == c && c < nb) invd[j0 + c] = li[c];
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int r = j0 + t + kSolveThreads * q;
        if (r <= n) {
          double* row = L + size_t(r) * (r + 1) / 2;
          double v[kSolveNB];
#pragma unroll
          for (int c = 0; c < kSolveNB; ++c) {
            double x = acc[q][c];
#pragma unroll
            for (int c2 = 0; c2 < c; ++c2) x = fma(-v[c2], l[c][c2], x);
            v[c] = (r == j0 + c) ? l[c][c] : x * li[c];
            if (c < nb && j0 + c <= min(r, n - 1)) row[j0 + c] = v[c];
          }
        }
      }
      __syncthreads();
    }
    if (t == 0 && bad) atomicExch(status + kStBad, 1);
    if (t < 32) {
      // back substitution L^T x = y with y = row n; unknown r lives in lane r % 32, slot r / 32
      double x[RPL];
      {
        const double* y = L + n_tri;
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
          const int r = lane + 32 * q;
          x[q] = r < n ? y[r] : 0.0;
        }
      }
#pragma unroll
      for (int qj = RPL - 1; qj >= 0; --qj) {
        if (32 * qj >= n) continue;
        for (int lj = min(31, n - 1 - 32 * qj); lj >= 0; --lj) {
          const int j = 32 * qj + lj;
          const double* row = L + size_t(j) * (j + 1) / 2;
          const double xj = __shfl_sync(0xffffffffu, x[qj], lj) * invd[j];
#pragma unroll
          for (int q = 0; q < RPL; ++q) {
            if (q > qj) continue;
            const int r = lane + 32 * q;
            if (r == j) x[q] = xj;
            else if (r < j) x[q] = fma(-row[r], xj, x[q]);
          }
        }
      }
      double* Ci = C + (int64_t(b) * k2 + i) * k1;
      if (lane == 0) Ci[0] = ci0;
#pragma unroll
      for (int q = 0; q < RPL; ++q) {
        const int r = lane + 32 * q;
        if (r < n) Ci[r + 1] = x[q];
      }
    }
    __syncthreads();  // shared memory is reused by the next queued system
  }
}

int64_t solve_lp_stride(int k1) {  // packed lower triangle of the (k1 - 1)^2 system, rounded up to an even count
  const int64_t n = k1 - 1, n_tri = n * (n + 1) / 2;
  return (n_tri + 1) & ~int64_t(1);
}

// ------------------------------------------------------------------------------------------------ float32 kernel
// Shared-memory layout of the float32 factor: BLOCK COLUMNS of four.  Block column J holds, for every row r >= 4 J
// (matrix rows 4 J .. n - 1, then row n = the right-hand side), the four entries (r, 4 J .. 4 J + 3) as one float4:
//     offset(J) = 4 J (n + 1) - 8 J (J - 1),   entry (r, c) at offset(c / 4) + 4 (r - 4 (c / 4)) + c % 4.
// Every access of the factorisation is then one aligned 16-byte load: a thread's own row (consecutive threads read
// consecutive float4 = conflict-free) and the four pivot rows (broadcast), each feeding four FMAs per loaded value.
__host__ __device__ inline int bc_offset(int J, int n) { return 4 * J * (n + 1) - 8 * J * (J - 1); }
__host__ __device__ inline int bc_blocks(int n) { return (n + 3) / 4; }
__host__ __device__ inline int bc_floats(int n) { return bc_offset(bc_blocks(n), n); }

// w_d (A A^T)[1:, 1:] of every pair in that layout (float32; the diagonal and the right-hand-side row are written per
// system by the solver; entries above the diagonal inside a diagonal block and padding columns are zero)
__global__ void __launch_bounds__(256)
    solve_pack32_kernel(const double* __restrict__ AAt, double wd, int k1, int64_t stride, float* __restrict__ Lp) {
  const int n = k1 - 1, NJ = bc_blocks(n);
  const double* aat = AAt + int64_t(blockIdx.x) * k1 * k1;
  float* out = Lp + int64_t(blockIdx.x) * stride;
  for (int J = 0; J < NJ; ++J) {
    const int base = bc_offset(J, n), rows = n + 1 - 4 * J;
    for (int e = threadIdx.x; e < rows * 4; e += blockDim.x) {
      const int r = 4 * J + (e >> 2), c = 4 * J + (e & 3);
      float v = 0.f;
      if (r < n && c <= r) v = float(wd * aat[int64_t(r + 1) * k1 + c + 1]);
      out[base + e] = v;
    }
  }
}

template <int T>
__device__ __forceinline__ void block_max2(float& a, float& b, float* red /* [2][T/32] */) {
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) {
    a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, sh));
    b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, sh));
  }
  __syncthreads();  // previous readers of red are done
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a, red[T / 32 + (threadIdx.x >> 5)] = b;
  __syncthreads();
  a = red[0], b = red[T / 32];
#pragma unroll
  for (int w = 1; w < T / 32; ++w) a = fmaxf(a, red[w]), b = fmaxf(b, red[T / 32 + w]);
}

// Blocked triangular solves with the finished factor, all threads of the CTA: thread t owns rows r = t + T q and keeps
// their running right-hand sides in registers; one barrier per block of four (the block's four values travel through
// a double-buffered shared patch, every thread solves the 4 x 4 triangle redundantly).
//   forward : L y = s        backward : L^T x = s        (in place in s; rows >= n hold 0)
template <int T, int RPT>
__device__ __forceinline__ void tri_forward(const float* __restrict__ L, const float* __restrict__ invd, float* blk,
                                            int n, float (&s)[RPT]) {
  const int t = threadIdx.x, NJ = bc_blocks(n);
  const int wmax = (t | 31) + T * (RPT - 1);  // largest row of this warp
  for (int J = 0, par = 0; J < NJ; ++J, par ^= 1) {
    const int j0 = 4 * J;
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int r = t + T * q;
      if (r >= j0 && r < j0 + 4) blk[par * 4 + (r - j0)] = s[q];
    }
    __syncthreads();
    if (wmax < j0) continue;  // warp-uniform: all rows of this warp are final
    const float* Lb = L + bc_offset(J, n);
    const float4 bv = *reinterpret_cast<const float4*>(blk + par * 4);
    const float4 li = *reinterpret_cast<const float4*>(invd + j0);
    const float4 r1 = *reinterpret_cast<const float4*>(Lb + 4), r2 = *reinterpret_cast<const float4*>(Lb + 8),
                 r3 = *reinterpret_cast<const float4*>(Lb + 12);
    const float y0 = bv.x * li.x;
    const float y1 = fmaf(-r1.x, y0, bv.y) * li.y;
    const float y2 = fmaf(-r2.y, y1, fmaf(-r2.x, y0, bv.z)) * li.z;
    const float y3 = fmaf(-r3.z, y2, fmaf(-r3.y, y1, fmaf(-r3.x, y0, bv.w))) * li.w;
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int r = t + T * q;
      if (r >= j0 + 4 && r < n) {
        const float4 lr = *reinterpret_cast<const float4*>(Lb + 4 * (r - j0));
        s[q] = fmaf(-lr.w, y3, fmaf(-lr.z, y2, fmaf(-lr.y, y1, fmaf(-lr.x, y0, s[q]))));
      } else if (r >= j0 && r < j0 + 4) {
        s[q] = (r == j0) ? y0 : (r == j0 + 1) ? y1 : (r == j0 + 2) ? y2 : y3;
      }
    }
  }
}

template <int T, int RPT>
__device__ __forceinline__ void tri_backward(const float* __restrict__ L, const float* __restrict__ invd, float* blk,
                                             int n, float (&s)[RPT]) {
  const int t = threadIdx.x, NJ = bc_blocks(n);
  const int wmin = t & ~31;  // smallest row of this warp
  for (int J = NJ - 1, par = 0; J >= 0; --J, par ^= 1) {
    const int j0 = 4 * J;
    if (t < 4 && j0 + t >= n) blk[par * 4 + t] = 0.f;  // padding columns of the last block
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int r = t + T * q;
      if (r >= j0 && r < j0 + 4 && r < n) blk[par * 4 + (r - j0)] = s[q];
    }
    __syncthreads();
    if (wmin >= j0 + 4) continue;
    const float* Lb = L + bc_offset(J, n);
    const float4 bv = *reinterpret_cast<const float4*>(blk + par * 4);
    const float4 li = *reinterpret_cast<const float4*>(invd + j0);
    const float4 r1 = *reinterpret_cast<const float4*>(Lb + 4), r2 = *reinterpret_cast<const float4*>(Lb + 8),
                 r3 = *reinterpret_cast<const float4*>(Lb + 12);
    const float x3 = bv.w * li.w;
    const float x2 = fmaf(-r3.z, x3, bv.z) * li.z;
    const float x1 = fmaf(-r3.y, x3, fmaf(-r2.y, x2, bv.y)) * li.y;
    const float x0 = fmaf(-r3.x, x3, fmaf(-r2.x, x2, fmaf(-r1.x, x1, bv.x))) * li.x;
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int r = t + T * q;
      if (r < j0) {
        // column r of rows j0 .. j0 + 3: block column r / 4, four rows one float4 apart
        const float* col = L + bc_offset(r >> 2, n) + 4 * (j0 - (r & ~3)) + (r & 3);
        s[q] = fmaf(-col[12], x3, fmaf(-col[8], x2, fmaf(-col[4], x1, fmaf(-col[0], x0, s[q]))));
      } else if (r < j0 + 4 && r < n) {
        s[q] = (r == j0) ? x0 : (r == j0 + 1) ? x1 : (r == j0 + 2) ? x2 : x3;
      }
    }
  }
}

constexpr int kMaxRefine = 4;

template <int T, int RPT, int MINB>  // (n + 1) <= T * RPT
__global__ void __launch_bounds__(T, MINB)
    fmap_solve32_kernel(const double* __restrict__ AAt, const double* __restrict__ BAt, const float* __restrict__ Lp,
                        int64_t lp_stride, const double* __restrict__ ev1, const double* __restrict__ ev2,
                        const double* __restrict__ c00, double wd, double wl, int k1, int k2, double* __restrict__ C,
                        int* __restrict__ status, int* __restrict__ list) {
  extern __shared__ __align__(16) float smf[];
  const int n = k1 - 1, NJ = bc_blocks(n);
  const int t = threadIdx.x, lane = t & 31;
  float* L = smf;                          // block-column factor, bc_floats(n)
  float* invd = L + bc_floats(n);          // [4 NJ] reciprocal diagonal (0 in padding columns)
  float* Dbuf = invd + 4 * NJ;             // [2][16] accumulated diagonal blocks
  float* blk = Dbuf + 32;                  // [2][4] triangular-solve exchange
  float* red = blk + 8;                    // [2][T / 32]
  double* x64 = reinterpret_cast<double*>(red + 2 * (T / 32) + ((2 * (T / 32)) & 1 ? 1 : 0));  // [n] current solution
  __shared__ double s_scale[T / 32];
  __shared__ __align__(8) unsigned long long s_bar;
  const int sys = blockIdx.x;
  const int b = sys / k2, i = sys % k2;
  const double* aat = AAt + int64_t(b) * k1 * k1;
  const double* bat = BAt + int64_t(b) * k2 * k1;
  const double* l1 = ev1 + int64_t(b) * k1;
  const double* l2 = ev2 + int64_t(b) * k2;
  const uint32_t bar = tc::smem_u32(&s_bar);
  if (t == 0) {
    tc::mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = uint32_t(bc_floats(n) * sizeof(float));
    tc::mbar_expect_tx(bar, bytes);
    tc::tma_load_1d(tc::smem_u32(L), Lp + int64_t(b) * lp_stride, bytes, bar);
  }
  double scale = -INFINITY;
  for (int j = t; j < k1; j += T) scale = fmax(scale, l1[j]);
  for (int j = t; j < k2; j += T) scale = fmax(scale, l2[j]);
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) scale = fmax(scale, __shfl_xor_sync(0xffffffffu, scale, sh));
  if (lane == 0) s_scale[t >> 5] = scale;
  __syncthreads();  // also publishes the barrier initialisation to the waiting threads
  scale = s_scale[0];
#pragma unroll
  for (int w = 1; w < T / 32; ++w) scale = fmax(scale, s_scale[w]);
  const double ci0 = (i == 0) ? c00[b] : 0.0;
  const double l2i = l2[i] / scale;
  // system-specific diagonal and right-hand side of the rows this thread owns in the solves: c = t + T q
  double dgg[RPT], rh[RPT];  // dgg = FULL diagonal entry w_d G_cc + w_l Delta_ic
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int c = t + T * q;
    dgg[q] = 1.0, rh[q] = 0.0;
    if (c < n) {
      const double df = l1[c + 1] / scale - l2i;
      dgg[q] = fma(wd, aat[int64_t(c + 1) * k1 + c + 1], wl * (df * df));
      rh[q] = wd * (bat[int64_t(i) * k1 + c + 1] - ci0 * aat[c + 1]);
    }
  }
  if (t < 4 * NJ - n) invd[n + t] = 0.f;  // padding columns of the last block
  tc::mbar_wait(bar, 0);
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int c = t + T * q;
    if (c < n) {
      const int base = bc_offset(c >> 2, n) + (c & 3);
      L[base + 4 * (c - (c & ~3))] = float(dgg[q]);
      L[base + 4 * (n - (c & ~3))] = float(rh[q]);
    }
  }
  __syncthreads();

  // ---- factorisation: left-looking, four columns per step, rows thread-cyclic from the step's first column
  bool bad = false;
  for (int J = 0, par = 0; J < NJ; ++J, par ^= 1) {
    const int j0 = 4 * J, nb = min(4, n - j0);
    const bool warp_active = j0 + (t & ~31) <= n;  // warp-uniform
    float4 acc[RPT];
    float* Lb = L + bc_offset(J, n);
    if (warp_active) {
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int r = j0 + t + T * q;
        acc[q] = r <= n ? *reinterpret_cast<const float4*>(Lb + 4 * (r - j0)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      for (int K = 0; K < J; ++K) {
        const float* Lk = L + bc_offset(K, n) + 4 * (j0 - 4 * K);  // row j0 of block column K
        const float4 c0 = *reinterpret_cast<const float4*>(Lk), c1 = *reinterpret_cast<const float4*>(Lk + 4),
                     c2 = *reinterpret_cast<const float4*>(Lk + 8), c3 = *reinterpret_cast<const float4*>(Lk + 12);
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
          const int r = j0 + t + T * q;
          if (r <= n) {
            const float4 o = *reinterpret_cast<const float4*>(Lk + 4 * (r - j0));
            acc[q].x = fmaf(-o.w, c0.w, fmaf(-o.z, c0.z, fmaf(-o.y, c0.y, fmaf(-o.x, c0.x, acc[q].x))));
            acc[q].y = fmaf(-o.w, c1.w, fmaf(-o.z, c1.z, fmaf(-o.y, c1.y, fmaf(-o.x, c1.x, acc[q].y))));
            acc[q].z = fmaf(-o.w, c2.w, fmaf(-o.z, c2.z, fmaf(-o.y, c2.y, fmaf(-o.x, c2.x, acc[q].z))));
            acc[q].w = fmaf(-o.w, c3.w, fmaf(-o.z, c3.z, fmaf(-o.y, c3.y, fmaf(-o.x, c3.x, acc[q].w))));
          }
        }
      }
      if (t < 4) *reinterpret_cast<float4*>(Dbuf + par * 16 + 4 * t) = acc[0];  // rows j0 .. j0 + 3
    }
    __syncthreads();
    if (warp_active) {
      // every thread factorises the 4 x 4 diagonal block (no serial owner)
      const float* D = Dbuf + par * 16;
      float l[4][4], li[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int c2 = 0; c2 <= c; ++c2) {
          float v = D[c * 4 + c2];
#pragma unroll
          for (int c3 = 0; c3 < c2; ++c3) v = fmaf(-l[c][c3], l[c2][c3], v);
          if (c2 == c) {
            const bool real = c < nb;
            if (real && !(v > 0.f)) bad = true;
            float y = rsqrtf(v);
            y = y * fmaf(-0.5f * v * y, y, 1.5f);  // one Newton step: the factor only has to be consistent
            li[c] = real ? y : 0.f;
            l[c][c] = real ? v * y : 1.f;
          } else {
            l[c][c2] = c < nb ? v * li[c2] : 0.f;
          }
        }
      }
      if (t < 4 && t < nb) invd[j0 + t] = t == 0 ? li[0] : t == 1 ? li[1] : t == 2 ? li[2] : li[3];
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int r = j0 + t + T * q;
        if (r <= n) {
          float4 v;
          v.x = acc[q].x * li[0];
          v.y = fmaf(-v.x, l[1][0], acc[q].y) * li[1];
          v.z = fmaf(-v.y, l[2][1], fmaf(-v.x, l[2][0], acc[q].z)) * li[2];
          v.w = fmaf(-v.z, l[3][2], fmaf(-v.y, l[3][1], fmaf(-v.x, l[3][0], acc[q].w))) * li[3];
          const int dr = r - j0;  // rows of the diagonal block keep their factor entries, zeros above the diagonal
          if (dr == 0) v = make_float4(l[0][0], 0.f, 0.f, 0.f);
          if (dr == 1) v = make_float4(l[1][0], l[1][1], 0.f, 0.f);
          if (dr == 2) v = make_float4(l[2][0], l[2][1], l[2][2], 0.f);
          if (dr == 3) v = make_float4(l[3][0], l[3][1], l[3][2], l[3][3]);
          if (r == n && dr < 4) {  // the right-hand-side row inside the last diagonal block: plain forward substitution
            v.x = acc[q].x * li[0];
            v.y = fmaf(-v.x, l[1][0], acc[q].y) * li[1];
            v.z = fmaf(-v.y, l[2][1], fmaf(-v.x, l[2][0], acc[q].z)) * li[2];
            v.w = fmaf(-v.z, l[3][2], fmaf(-v.y, l[3][1], fmaf(-v.x, l[3][0], acc[q].w))) * li[3];
          }
          *reinterpret_cast<float4*>(Lb + 4 * dr) = v;
        }
      }
    }
    __syncthreads();
  }
  const int any_bad = __syncthreads_or(bad ? 1 : 0);

  // ---- first solution: y = row n (forward substitution came with the factorisation), x = L^-T y
  float s[RPT];
  double xr[RPT];
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int c = t + T * q;
    s[q] = c < n ? L[bc_offset(c >> 2, n) + 4 * (n - (c & ~3)) + (c & 3)] : 0.f;
  }
  tri_backward<T, RPT>(L, invd, blk, n, s);
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int c = t + T * q;
    xr[q] = double(s[q]);
    if (c < n) x64[c] = xr[q];
  }
  __syncthreads();

  // ---- float64 iterative refinement: r = rhs - M x (M from the float64 Gram matrix), L L^T dx = r, x += dx
  bool converged = false;
  int steps = 0;
  float rho_prev = 1.f;
  if (!any_bad) {
    for (; steps < kMaxRefine && !converged; ++steps) {
      float xm = 0.f, dm = 0.f;
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int c = t + T * q;
        double r0 = 0.0, r1 = 0.0;
        if (c < n) {
          const double* g = aat + k1 + (c + 1);  // column c + 1 of rows 1 .. n (the matrix is symmetric: coalesced)
          int j = 0;
          for (; j + 1 < n; j += 2) {
            r0 = fma(g[int64_t(j) * k1], x64[j], r0);
            r1 = fma(g[int64_t(j + 1) * k1], x64[j + 1], r1);
          }
          if (j < n) r0 = fma(g[int64_t(j) * k1], x64[j], r0);
          // the diagonal of M is dgg, not w_d G_cc: add the Laplacian part separately
          const double lap = dgg[q] - wd * g[int64_t(c) * k1];
          r0 = rh[q] - fma(wd, r0 + r1, lap * xr[q]);
        }
        s[q] = float(r0);
      }
      __syncthreads();  // x64 is rewritten below only after every thread has read it
      tri_forward<T, RPT>(L, invd, blk, n, s);
      tri_backward<T, RPT>(L, invd, blk, n, s);
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int c = t + T * q;
        if (c < n) {
          xr[q] += double(s[q]);
          x64[c] = xr[q];
          xm = fmaxf(xm, fabsf(float(xr[q])));
          dm = fmaxf(dm, fabsf(s[q]));
        }
      }
      block_max2<T>(xm, dm, red);  // its barriers also publish x64
      // correction ratio rho; the error left after this step is about rho * (rho / rho_prev)
      const float rho = dm / fmaxf(xm, 1e-30f);
      const float left = steps == 0 ? rho * rho : rho * (rho / rho_prev);
      converged = (rho <= 1e-12f) || (rho < 0.25f * rho_prev && left <= 1e-10f) || (xm == 0.f && dm == 0.f);
      if (!(rho < 0.5f)) break;  // not contracting (or NaN): leave it to the float64 kernel
      rho_prev = rho;
    }
  }
  if (t == 0) {
    atomicAdd(status + kStRefine, steps);
    if (!converged) list[atomicAdd(status + kStFallback, 1)] = sys;
  }
  if (!converged) return;
  double* Ci = C + (int64_t(b) * k2 + i) * k1;
  if (t == 0) Ci[0] = ci0;
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int c = t + T * q;
    if (c < n) Ci[c + 1] = xr[q];
  }
}

// ------------------------------------------------------------------------------------------------ one warp per system
// The block kernel above spends most of its issue slots waiting: towards the end of the factorisation one warp of four
// still owns rows, the 4 x 4 diagonal block is factorised once per warp, every step costs two block barriers (ncu:
// 5.7 warps stalled on the barrier per issued instruction, 50 k warp instructions per system).  Here ONE WARP owns a
// system: lane l holds rows j0 + l + 32 q (q < RPT) of the current step, so a pivot row loaded from shared memory feeds
// RPT x 16 FMAs, the diagonal block is factorised once, and the only synchronisation is __syncwarp.  ~17 k warp
// instructions per system; 9-10 systems (warps) resident per SM, limited by the 22 KB factor each keeps in shared memory.
template <int RPT>
__device__ __forceinline__ void wtri_forward(const float* __restrict__ L, const float* __restrict__ invd, float* blk,
                                             int n, int lane, float (&s)[RPT]) {
  const int NJ = bc_blocks(n);
  const float* Lb = L;
  for (int J = 0; J < NJ; ++J) {
    const int j0 = 4 * J, qq = j0 >> 5, l0 = j0 & 31;
    if (lane < 4 && j0 + lane >= n) blk[lane] = 0.f;
#pragma unroll
    for (int q = 0; q < RPT; ++q)
      if (q == qq && lane >= l0 && lane < l0 + 4 && lane + 32 * q < n) blk[lane - l0] = s[q];
    __syncwarp();
    const float4 bv = *reinterpret_cast<const float4*>(blk);
    const float4 li = *reinterpret_cast<const float4*>(invd + j0);
    const float4 r1 = *reinterpret_cast<const float4*>(Lb + 4), r2 = *reinterpret_cast<const float4*>(Lb + 8),
                 r3 = *reinterpret_cast<const float4*>(Lb + 12);
    const float y0 = bv.x * li.x;
    const float y1 = fmaf(-r1.x, y0, bv.y) * li.y;
    const float y2 = fmaf(-r2.y, y1, fmaf(-r2.x, y0, bv.z)) * li.z;
    const float y3 = fmaf(-r3.z, y2, fmaf(-r3.y, y1, fmaf(-r3.x, y0, bv.w))) * li.w;
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      if (32 * q + 31 < j0) continue;  // warp-uniform: these rows are final
      const int r = lane + 32 * q;
      if (r >= j0 + 4 && r < n) {
        const float4 lr = *reinterpret_cast<const float4*>(Lb + 4 * (r - j0));
        s[q] = fmaf(-lr.w, y3, fmaf(-lr.z, y2, fmaf(-lr.y, y1, fmaf(-lr.x, y0, s[q]))));
      } else if (r >= j0 && r < j0 + 4) {
        s[q] = (r == j0) ? y0 : (r == j0 + 1) ? y1 : (r == j0 + 2) ? y2 : y3;
      }
    }
    __syncwarp();  // blk is rewritten by the next step
    Lb += 4 * (n + 1 - j0);
  }
}

template <int RPT>
__device__ __forceinline__ void wtri_backward(const float* __restrict__ L, const float* __restrict__ invd, float* blk,
                                              int n, int lane, float (&s)[RPT]) {
  const int NJ = bc_blocks(n);
  // column r of row j: block column r / 4 at L + colbase[q] + 4 j   (bc_offset(r / 4) - 4 (r & ~3) + (r & 3))
  int colbase[RPT];
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int r = min(lane + 32 * q, n - 1);
    colbase[q] = bc_offset(r >> 2, n) - 4 * (r & ~3) + (r & 3);
  }
  for (int J = NJ - 1; J >= 0; --J) {
    const int j0 = 4 * J, qq = j0 >> 5, l0 = j0 & 31;
    if (lane < 4 && j0 + lane >= n) blk[lane] = 0.f;  // padding columns of the last block
#pragma unroll
    for (int q = 0; q < RPT; ++q)
      if (q == qq && lane >= l0 && lane < l0 + 4 && lane + 32 * q < n) blk[lane - l0] = s[q];
    __syncwarp();
    const float* Lb = L + bc_offset(J, n);
    const float4 bv = *reinterpret_cast<const float4*>(blk);
    const float4 li = *reinterpret_cast<const float4*>(invd + j0);
    const float4 r1 = *reinterpret_cast<const float4*>(Lb + 4), r2 = *reinterpret_cast<const float4*>(Lb + 8),
                 r3 = *reinterpret_cast<const float4*>(Lb + 12);
    const float x3 = bv.w * li.w;
    const float x2 = fmaf(-r3.z, x3, bv.z) * li.z;
    const float x1 = fmaf(-r3.y, x3, fmaf(-r2.y, x2, bv.y)) * li.y;
    const float x0 = fmaf(-r3.x, x3, fmaf(-r2.x, x2, fmaf(-r1.x, x1, bv.x))) * li.x;
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      if (32 * q >= j0 + 4) continue;  // warp-uniform: rows of this slot are all above the block
      const int r = lane + 32 * q;
      if (r < j0) {
        const float* col = L + colbase[q] + 4 * j0;
        s[q] = fmaf(-col[12], x3, fmaf(-col[8], x2, fmaf(-col[4], x1, fmaf(-col[0], x0, s[q]))));
      } else if (r < j0 + 4 && r < n) {
        s[q] = (r == j0) ? x0 : (r == j0 + 1) ? x1 : (r == j0 + 2) ? x2 : x3;
      }
    }
    __syncwarp();
  }
}

// One block step of the warp factorisation with NQ live row slots (compile time: no predicates in the inner loop).
// The operands of block column K + 1 are loaded before the FMAs of block column K issue.
template <int NQ>
__device__ __forceinline__ bool wfactor_step(const float* __restrict__ L, float* __restrict__ Lb, float* __restrict__ invd,
                                             float* __restrict__ Dbuf, int n, int J, int lane) {
  const int j0 = 4 * J, nb = min(4, n - j0);
  float4 acc[NQ];
  int ro[NQ];  // float offset of the own row inside a block column, relative to row j0 (rows beyond n: clamped)
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    ro[q] = 4 * (min(j0 + lane + 32 * q, n) - j0);
    acc[q] = *reinterpret_cast<const float4*>(Lb + ro[q]);
  }
  const float* Lk = L + 4 * j0;  // row j0 of block column 0; block column K + 1 starts 4 (n + 1) - 16 (K + 1) floats later
  // two register sets (A: even K, B: odd K) alternate, so that the loads of block column K + 1 are in flight while the
  // FMAs of block column K issue and no register moves are needed
  float4 ca[4], oa[NQ], cb[4], ob[NQ];
#define DM_WF_LOAD(C, O)                                                  \
  do {                                                                    \
    C[0] = *reinterpret_cast<const float4*>(Lk);                          \
    C[1] = *reinterpret_cast<const float4*>(Lk + 4);                      \
    C[2] = *reinterpret_cast<const float4*>(Lk + 8);                      \
    C[3] = *reinterpret_cast<const float4*>(Lk + 12);                     \
    _Pragma("unroll") for (int q = 0; q < NQ; ++q) O[q] = *reinterpret_cast<const float4*>(Lk + ro[q]); \
  } while (0)
#define DM_WF_FMA(C, O)                                                                                               \
  do {                                                                                                                \
    _Pragma("unroll") for (int q = 0; q < NQ; ++q) {                                                                  \
      acc[q].x = fmaf(-O[q].w, C[0].w, fmaf(-O[q].z, C[0].z, fmaf(-O[q].y, C[0].y, fmaf(-O[q].x, C[0].x, acc[q].x)))); \
      acc[q].y = fmaf(-O[q].w, C[1].w, fmaf(-O[q].z, C[1].z, fmaf(-O[q].y, C[1].y, fmaf(-O[q].x, C[1].x, acc[q].y)))); \
      acc[q].z = fmaf(-O[q].w, C[2].w, fmaf(-O[q].z, C[2].z, fmaf(-O[q].y, C[2].y, fmaf(-O[q].x, C[2].x, acc[q].z)))); \
      acc[q].w = fmaf(-O[q].w, C[3].w, fmaf(-O[q].z, C[3].z, fmaf(-O[q].y, C[3].y, fmaf(-O[q].x, C[3].x, acc[q].w)))); \
    }                                                                                                                 \
  } while (0)
  if (J > 0) DM_WF_LOAD(ca, oa);
  int K = 0;
  for (; K + 2 <= J; K += 2) {
    Lk += 4 * (n + 1) - 16 * (K + 1);
    DM_WF_LOAD(cb, ob);
    DM_WF_FMA(ca, oa);
    Lk += 4 * (n + 1) - 16 * (K + 2);
    if (K + 2 < J) DM_WF_LOAD(ca, oa);
    DM_WF_FMA(cb, ob);
  }
  if (K < J) DM_WF_FMA(ca, oa);
#undef DM_WF_LOAD
#undef DM_WF_FMA
  if (lane < 4) *reinterpret_cast<float4*>(Dbuf + 4 * lane) = acc[0];  // rows j0 .. j0 + 3
  __syncwarp();
  bool bad = false;
  float l[4][4], li[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int c2 = 0; c2 <= c; ++c2) {
      float v = Dbuf[c * 4 + c2];
#pragma unroll
      for (int c3 = 0; c3 < c2; ++c3) v = fmaf(-l[c][c3], l[c2][c3], v);
      if (c2 == c) {
        const bool real = c < nb;
        if (real && !(v > 0.f)) bad = true;
        const float y = rsqrtf(v);  // 2 ulp: the factor is a preconditioner, it only has to be consistent
        li[c] = real ? y : 0.f;
        l[c][c] = real ? v * y : 1.f;
      } else {
        l[c][c2] = c < nb ? v * li[c2] : 0.f;
      }
    }
  }
  if (lane < nb) invd[j0 + lane] = lane == 0 ? li[0] : lane == 1 ? li[1] : lane == 2 ? li[2] : li[3];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int r = j0 + lane + 32 * q;
    float4 v;
    v.x = acc[q].x * li[0];
    v.y = fmaf(-v.x, l[1][0], acc[q].y) * li[1];
    v.z = fmaf(-v.y, l[2][1], fmaf(-v.x, l[2][0], acc[q].z)) * li[2];
    v.w = fmaf(-v.z, l[3][2], fmaf(-v.y, l[3][1], fmaf(-v.x, l[3][0], acc[q].w))) * li[3];
    if (q == 0 && lane < 4 && r < n) {  // rows of the diagonal block keep their factor entries
      v = lane == 0   ? make_float4(l[0][0], 0.f, 0.f, 0.f)
          : lane == 1 ? make_float4(l[1][0], l[1][1], 0.f, 0.f)
          : lane == 2 ? make_float4(l[2][0], l[2][1], l[2][2], 0.f)
                      : make_float4(l[3][0], l[3][1], l[3][2], l[3][3]);
    }
    if (r <= n) *reinterpret_cast<float4*>(Lb + ro[q]) = v;
  }
  __syncwarp();
  return bad;
}

template <int RPT>  // (n + 1) <= 32 * RPT
__global__ void __launch_bounds__(32, 10)
    fmap_solve32w_kernel(const double* __restrict__ AAt, const double* __restrict__ BAt, const float* __restrict__ Lp,
                         int64_t lp_stride, const double* __restrict__ ev1, const double* __restrict__ ev2,
                         const double* __restrict__ c00, double wd, double wl, int k1, int k2, double* __restrict__ C,
                         int* __restrict__ status, int* __restrict__ list) {
  extern __shared__ __align__(16) float smf[];
  const int n = k1 - 1, NJ = bc_blocks(n);
  const int lane = threadIdx.x;
  float* L = smf;                  // block-column factor, bc_floats(n)
  float* invd = L + bc_floats(n);  // [4 NJ]
  float* Dbuf = invd + 4 * NJ;     // [16]
  float* blk = Dbuf + 16;          // [4]
  double* x64 = reinterpret_cast<double*>(blk + 4);  // [n]   (float count so far is a multiple of 4)
  __shared__ __align__(8) unsigned long long s_bar;
  const int sys = blockIdx.x;
  const int b = sys / k2, i = sys % k2;
  const double* aat = AAt + int64_t(b) * k1 * k1;
  const double* bat = BAt + int64_t(b) * k2 * k1;
  const double* l1 = ev1 + int64_t(b) * k1;
  const double* l2 = ev2 + int64_t(b) * k2;
  const uint32_t bar = tc::smem_u32(&s_bar);
  if (lane == 0) {
    tc::mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = uint32_t(bc_floats(n) * sizeof(float));
    tc::mbar_expect_tx(bar, bytes);
    tc::tma_load_1d(tc::smem_u32(L), Lp + int64_t(b) * lp_stride, bytes, bar);
  }
  double scale = -INFINITY;
  for (int j = lane; j < k1; j += 32) scale = fmax(scale, l1[j]);
  for (int j = lane; j < k2; j += 32) scale = fmax(scale, l2[j]);
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) scale = fmax(scale, __shfl_xor_sync(0xffffffffu, scale, sh));
  const double ci0 = (i == 0) ? c00[b] : 0.0;
  const double l2i = l2[i] / scale;
  double dgg[RPT], rh[RPT];  // full diagonal entry w_d G_cc + w_l Delta_ic and right-hand side of rows c = lane + 32 q
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int c = lane + 32 * q;
    dgg[q] = 1.0, rh[q] = 0.0;
    if (c < n) {
      const double df = l1[c + 1] / scale - l2i;
      dgg[q] = fma(wd, aat[int64_t(c + 1) * k1 + c + 1], wl * (df * df));
      rh[q] = wd * (bat[int64_t(i) * k1 + c + 1] - ci0 * aat[c + 1]);
    }
  }
  if (lane < 4 * NJ - n) invd[n + lane] = 0.f;  // padding columns of the last block
  __syncwarp();                                 // the barrier initialisation is visible to every lane
  tc::mbar_wait(bar, 0);
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int c = lane + 32 * q;
    if (c < n) {
      const int base = bc_offset(c >> 2, n) + (c & 3);
      L[base + 4 * (c - (c & ~3))] = float(dgg[q]);
      L[base + 4 * (n - (c & ~3))] = float(rh[q]);
    }
  }
  __syncwarp();

  // ---- factorisation: left-looking, four columns per step; lane l owns rows j0 + l + 32 q
  bool bad = false;
  {
    float* Lb = L;  // block column J
    for (int J = 0; J < NJ; ++J) {
      const int j0 = 4 * J;
      const int nq = (n - j0) / 32 + 1;  // slots that still hold a row <= n (warp-uniform)
      bool b = false;
      if (RPT >= 4 && nq == 4) b = wfactor_step<4>(L, Lb, invd, Dbuf, n, J, lane);
      else if (RPT >= 3 && nq == 3) b = wfactor_step<3>(L, Lb, invd, Dbuf, n, J, lane);
      else if (RPT >= 2 && nq == 2) b = wfactor_step<2>(L, Lb, invd, Dbuf, n, J, lane);
      else b = wfactor_step<1>(L, Lb, invd, Dbuf, n, J, lane);
      bad |= b;
      Lb += 4 * (n + 1 - j0);
    }
  }
  const bool any_bad = __any_sync(0xffffffffu, bad);

  // ---- first solution: y = row n (forward substitution came with the factorisation), x = L^-T y
  float s[RPT];
  double xr[RPT];
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int c = lane + 32 * q;
    s[q] = c < n ? L[bc_offset(c >> 2, n) + 4 * (n - (c & ~3)) + (c & 3)] : 0.f;
  }
  __syncwarp();
  wtri_backward<RPT>(L, invd, blk, n, lane, s);
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int c = lane + 32 * q;
    xr[q] = double(s[q]);
    if (c < n) x64[c] = xr[q];
  }
  __syncwarp();

  // ---- float64 iterative refinement: r = rhs - M x (M from the float64 Gram matrix), L L^T dx = r, x += dx
  bool converged = false;
  int steps = 0;
  float rho_prev = 1.f;
  if (!any_bad) {
    for (; steps < kMaxRefine && !converged; ++steps) {
      // column c + 1 of rows 1 .. n of the float64 Gram matrix (symmetric: coalesced over the lanes); lanes beyond the
      // matrix read a clamped column and discard the result.  JU rows per slot are in flight at a time.
      double r0[RPT], r1[RPT];
      const double* g[RPT];
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        r0[q] = r1[q] = 0.0;
        g[q] = aat + k1 + (min(lane + 32 * q, n - 1) + 1);
      }
      constexpr int JU = RPT >= 4 ? 4 : 8;
      const double* gj = g[0];  // row cursor (slot 0); the other slots sit 32 q columns to the right
      int qoff[RPT];
#pragma unroll
      for (int q = 0; q < RPT; ++q) qoff[q] = int(g[q] - g[0]);
      int j = 0;
      for (; j + JU <= n; j += JU) {
        double gv[RPT][JU];
#pragma unroll
        for (int u = 0; u < JU; ++u)
#pragma unroll
          for (int q = 0; q < RPT; ++q) gv[q][u] = gj[u * k1 + qoff[q]];
        gj += JU * k1;
#pragma unroll
        for (int u = 0; u < JU; ++u) {
          const double xv = x64[j + u];
#pragma unroll
          for (int q = 0; q < RPT; ++q) {
            if (u & 1) r1[q] = fma(gv[q][u], xv, r1[q]);
            else r0[q] = fma(gv[q][u], xv, r0[q]);
          }
        }
      }
      for (; j < n; ++j, gj += k1) {
        const double xa = x64[j];
#pragma unroll
        for (int q = 0; q < RPT; ++q) r0[q] = fma(gj[qoff[q]], xa, r0[q]);
      }
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int c = lane + 32 * q;
        double res = 0.0;
        if (c < n) {
          const double lap = dgg[q] - wd * g[q][int64_t(c) * k1];  // the Laplacian part of the diagonal
          res = rh[q] - fma(wd, r0[q] + r1[q], lap * xr[q]);
        }
        s[q] = float(res);
      }
      __syncwarp();
      wtri_forward<RPT>(L, invd, blk, n, lane, s);
      wtri_backward<RPT>(L, invd, blk, n, lane, s);
      float xm = 0.f, dm = 0.f;
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int c = lane + 32 * q;
        if (c < n) {
          xr[q] += double(s[q]);
          x64[c] = xr[q];
          xm = fmaxf(xm, fabsf(float(xr[q])));
          dm = fmaxf(dm, fabsf(s[q]));
        }
      }
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) {
        xm = fmaxf(xm, __shfl_xor_sync(0xffffffffu, xm, sh));
        dm = fmaxf(dm, __shfl_xor_sync(0xffffffffu, dm, sh));
      }
      __syncwarp();  // x64 is complete before the next residual
      // correction ratio rho; the error left after this step is about rho * (rho / rho_prev)
      const float rho = dm / fmaxf(xm, 1e-30f);
      const float left = steps == 0 ? rho * rho : rho * (rho / rho_prev);
      converged = (rho <= 1e-12f) || (rho < 0.25f * rho_prev && left <= 1e-10f) || (xm == 0.f && dm == 0.f);
      if (!(rho < 0.5f)) break;  // not contracting (or NaN): leave it to the float64 kernel
      rho_prev = rho;
    }
  }
  if (lane == 0) {
    atomicAdd(status + kStRefine, steps);
    if (!converged) list[atomicAdd(status + kStFallback, 1)] = sys;
  }
  if (!converged) return;
  double* Ci = C + (int64_t(b) * k2 + i) * k1;
  if (lane == 0) Ci[0] = ci0;
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int c = lane + 32 * q;
    if (c < n) Ci[c + 1] = xr[q];
  }
}

size_t solve32w_shmem(int n) {
  const int NJ = bc_blocks(n);
  return (size_t(bc_floats(n)) + 4 * NJ + 16 + 4) * sizeof(float) + size_t(n) * sizeof(double);
}

size_t solve32_shmem(int n, int T) {
  const int NJ = bc_blocks(n);
  size_t fl = size_t(bc_floats(n)) + 4 * NJ + 32 + 8 + 2 * (T / 32);
  fl = (fl + 1) & ~size_t(1);
  return fl * sizeof(float) + size_t(n) * sizeof(double);
}

int solve_mode() {  // DM_SOLVE: "f64" forces the float64 kernel, "f32t64" the block-per-system float32 kernel (A/B runs)
  const char* e = getenv("DM_SOLVE");
  if (!e) return 0;
  if (e[0] == 'f' && e[1] == '6') return 1;
  if (e[0] == 'f' && e[1] == '3' && e[3] == 't') return 2;
  return 0;
}

struct SolveLayout {
  int* status;
  double *AAt, *BAt;
  int* list;
  double* Lp64;
  float* Lp32;
  int64_t stride64, stride32;
  size_t bytes;
};
SolveLayout solve_carve(void* ws, int n_pairs, int k1, int k2) {
  Carver c(ws);
  SolveLayout L;
  L.status = c.take<int>(64);  // first: dm_fmap_solve_read_status needs no sizes
  L.AAt = c.take<double>(size_t(n_pairs) * k1 * k1);
  L.BAt = c.take<double>(size_t(n_pairs) * k2 * k1);
  L.list = c.take<int>(size_t(n_pairs) * k2);
  L.stride64 = solve_lp_stride(k1);
  L.stride32 = (bc_floats(k1 - 1) + 3) & ~3;
  // the float64 pack is written only when the fallback may run; both are carved so that the size does not depend on the mode
  L.Lp64 = c.take<double>(size_t(n_pairs) * L.stride64);
  L.Lp32 = c.take<float>(size_t(n_pairs) * L.stride32);
  L.bytes = c.bytes();
  return L;
}

}  // namespace
}  // namespace dm

using namespace dm;

extern "C" {

size_t dm_fmap_solve_workspace_bytes(int n_pairs, int k1, int k2, int d) {
  (void)d;
  if (n_pairs < 0 || k1 < 2 || k2 < 1) return 0;
  return solve_carve(nullptr, n_pairs, k1, k2).bytes;
}

int dm_fmap_solve(const double* A, const double* B, const double* evals1, const double* evals2, const double* c00,
                  double w_descr, double w_lap, int n_pairs, int k1, int k2, int d, double* C, void* workspace,
                  size_t workspace_bytes, dm_stream_t stream) {
  if (n_pairs < 0 || k1 < 2 || k2 < 1 || d <= 0) DM_FAIL(DM_ERR_BADARG, "bad size (need k1 >= 2)");
  if (n_pairs == 0) return DM_OK;
  if (!A || !B || !evals1 || !evals2 || !c00 || !C) DM_FAIL(DM_ERR_BADARG, "null argument");
  const size_t need = dm_fmap_solve_workspace_bytes(n_pairs, k1, k2, d);
  if (!workspace || need > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu", need);
  if (reinterpret_cast<uintptr_t>(workspace) % 256) DM_FAIL(DM_ERR_ALIGN, "workspace must be 256-byte aligned");
  const int n = k1 - 1;
  const size_t per64 = sizeof(double) * (size_t(n + 1) * (n + 2) / 2);
  if (per64 > 220 * 1024 || n + 1 > 256)
    DM_FAIL(DM_ERR_UNSUPPORTED, "k1 = %d too large for the in-shared-memory Cholesky (max 236)", k1);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SolveLayout L = solve_carve(workspace, n_pairs, k1, k2);
  DM_CUDA_OK(cudaMemsetAsync(L.status, 0, 64 * sizeof(int), st));
  int rc;
  // profiling (DM_SOLVE_SKIP_PREP=1): reuse the Gram matrices and the packed factor input a previous identical call left
  // in the workspace, so that the factor / refine kernel can be timed alone with CUDA events
  const char* sp_env = getenv("DM_SOLVE_SKIP_PREP");
  const bool skip_prep = sp_env && sp_env[0] == '1';
  if (!skip_prep) {
    GemmProblem G;
    G.A.d = A, G.A.ld = d, G.A.batch_stride = int64_t(k1) * d, G.A.trans = 0;
    G.B = G.A;
    G.M = k1, G.N = k1, G.K = d, G.maxM = k1, G.maxN = k1, G.maxK = d, G.n_batch = n_pairs;
    G.C = L.AAt, G.ldc = k1, G.c_batch_stride = int64_t(k1) * k1;
    if ((rc = gemm64_launch(G, st))) return rc;
    G.A.d = B, G.A.batch_stride = int64_t(k2) * d;
    G.M = k2, G.maxM = k2, G.C = L.BAt, G.c_batch_stride = int64_t(k2) * k1;
    if ((rc = gemm64_launch(G, st))) return rc;
  }
  const int64_t n_sys = int64_t(n_pairs) * k2;
  if (n_sys > 0x7fffffffLL) DM_FAIL(DM_ERR_BADARG, "too many systems");
  const int mode = solve_mode();
  const size_t shm64 = per64 + (1 + 2 * kSolveNB * kSolveNB + 2 * kSolveNB + size_t(n)) * sizeof(double);
#define DM_SOLVE64(RPT, GRID, LIST, COUNT)                                                                          \
  do {                                                                                                              \
    static OncePerDevice once;                                                                                      \
    if (shm64 > 48 * 1024 && once.first())                                                                          \
      DM_CUDA_OK(cudaFuncSetAttribute(fmap_solve_kernel<RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256)); \
    fmap_solve_kernel<RPT><<<unsigned(GRID), kSolveThreads, shm64, st>>>(                                           \
        L.AAt, L.BAt, L.Lp64, L.stride64, evals1, evals2, c00, w_descr, w_lap, k1, k2, C, L.status, LIST, COUNT);   \
    DM_LAUNCH_OK("fmap_solve_kernel");                                                                              \
  } while (0)
  if (mode == 1) {  // float64 only
    solve_pack_kernel<<<unsigned(n_pairs), 256, 0, st>>>(L.AAt, w_descr, k1, L.stride64, L.Lp64, nullptr);
    DM_LAUNCH_OK("solve_pack_kernel");
    if (n + 1 <= kSolveThreads)
      DM_SOLVE64(1, n_sys, nullptr, nullptr);
    else
      DM_SOLVE64(2, n_sys, nullptr, nullptr);
    return DM_OK;
  }
  if (!skip_prep) {
    solve_pack32_kernel<<<unsigned(n_pairs), 256, 0, st>>>(L.AAt, w_descr, k1, L.stride32, L.Lp32);
    DM_LAUNCH_OK("solve_pack32_kernel");
  }
#define DM_SOLVE32(T, RPT, MINB)                                                                                    \
  do {                                                                                                              \
    const size_t shm = solve32_shmem(n, T);                                                                         \
    static OncePerDevice once;                                                                                      \
    if (once.first())                                                                                               \
      DM_CUDA_OK(cudaFuncSetAttribute(fmap_solve32_kernel<T, RPT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      200 * 1024));                                                                 \
    fmap_solve32_kernel<T, RPT, MINB><<<unsigned(n_sys), T, shm, st>>>(L.AAt, L.BAt, L.Lp32, L.stride32, evals1,    \
                                                                       evals2, c00, w_descr, w_lap, k1, k2, C,      \
                                                                       L.status, L.list);                           \
    DM_LAUNCH_OK("fmap_solve32_kernel");                                                                            \
  } while (0)
#define DM_SOLVE32W(RPT)                                                                                            \
  do {                                                                                                              \
    const size_t shm = solve32w_shmem(n);                                                                           \
    static OncePerDevice once;                                                                                      \
    if (once.first())                                                                                               \
      DM_CUDA_OK(cudaFuncSetAttribute(fmap_solve32w_kernel<RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                      100 * 1024));                                                                 \
    fmap_solve32w_kernel<RPT><<<unsigned(n_sys), 32, shm, st>>>(L.AAt, L.BAt, L.Lp32, L.stride32, evals1, evals2,   \
                                                                c00, w_descr, w_lap, k1, k2, C, L.status, L.list);  \
    DM_LAUNCH_OK("fmap_solve32w_kernel");                                                                           \
  } while (0)
  if (mode == 2) {  // the block-per-system float32 kernel (A/B runs)
    if (n + 1 <= 128)
      DM_SOLVE32(128, 1, 8);
    else
      DM_SOLVE32(128, 2, 2);
  } else if (n + 1 <= 32)
    DM_SOLVE32W(1);
  else if (n + 1 <= 64)
    DM_SOLVE32W(2);
  else if (n + 1 <= 128)
    DM_SOLVE32W(4);
  else
    DM_SOLVE32(128, 2, 2);
#undef DM_SOLVE32
#undef DM_SOLVE32W
  // whatever the float32 path could not finish, in float64 (the queue is normally empty: the CTAs return at once)
  solve_pack_kernel<<<unsigned(n_pairs), 256, 0, st>>>(L.AAt, w_descr, k1, L.stride64, L.Lp64, L.status + kStFallback);
  DM_LAUNCH_OK("solve_pack_kernel");
  const int fb_grid = int(n_sys < 4 * num_sms() ? n_sys : 4 * num_sms());
  if (n + 1 <= kSolveThreads)
    DM_SOLVE64(1, fb_grid, L.list, L.status + kStFallback);
  else
    DM_SOLVE64(2, fb_grid, L.list, L.status + kStFallback);
#undef DM_SOLVE64
  return DM_OK;
}

int dm_fmap_solve_read_status(const void* workspace, int* out_h, dm_stream_t stream) {
  if (!workspace || !out_h) DM_FAIL(DM_ERR_BADARG, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DM_CUDA_OK(cudaMemcpyAsync(out_h, workspace, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  DM_CUDA_OK(cudaStreamSynchronize(st));
  return DM_OK;
}

}  // extern "C"
