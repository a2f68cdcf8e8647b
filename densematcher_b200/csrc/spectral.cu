// Eigenbasis provider and spectral transforms on the device.
//
//  (1) dm_lbo_eigs: the k lowest eigenpairs of the generalised problem  W phi = lambda A phi  (W = cotangent stiffness in
//      CSR, A = lumped vertex areas) that TriMesh.process obtains from scipy's shift-invert ARPACK call
//      (densematcher/pyFM/mesh/laplacian.py:143-182, trimesh.py:498-531; diffusion_net/geometry.py:330-380 does the same
//      for the operator cache).  ARPACK's shift-invert needs a sparse factorisation, which is a poor fit for a GPU; the
//      device algorithm is a Chebyshev-filtered block subspace iteration on the symmetrised operator
//      S = A^-1/2 W A^-1/2 (same spectrum, eigenvectors u = A^1/2 phi):
//          X <- p_d(S) X          d fused "sparse row times dense block" passes (spmm_cheb_kernel), three-term recurrence
//          X <- X T               orthonormalisation: Cholesky-QR of the column-scaled Gram matrix (chol_kernel)
//          H = X^T S X,  H = V Theta V^T,  X <- X V      Rayleigh-Ritz
//      with the block products on the float64 tensor-core GEMM (gemm64.cu) and the dense m x m symmetric eigenproblems
//      (m = k + guard columns <= 512) solved by sym_eig_kernel: Householder tridiagonalisation, explicit Q, implicit QL --
//      one CTA, matrix in L2.  The filter damps [theta_max(block), gershgorin bound]; iteration stops when every wanted
//      residual |S x - theta x| <= tol * theta_k.
//  (2) dm_from_basis / dm_spectral_diffusion: DiffusionNet's  from_basis(exp(-lambda t) * to_basis(x))
//      (diffusion_net/layers.py:56-67, geometry.py:572-598) as projection (tcgen05, fm.cu) -> coefficient scaling ->
//      float64 tensor-core GEMM.
#include "dm_internal.cuh"
#include "gemm64.cuh"

namespace dm {
namespace {

constexpr int kEigMaxM = 512;     // largest dense symmetric eigenproblem (block width)
constexpr int kEigThreads = 1024;
constexpr int kMaxDegree = 64;
constexpr int kRing = 4;          // sweeps in flight between the QL recurrence and the rotation application

// ---------------------------------------------------------------------------------------------------------------
// dense symmetric eigensolver: A (m x m, row-major, pitch lda; destroyed) -> w ascending, V columns = eigenvectors
// (row-major, pitch ldv).  Z: m x lda scratch.  One CTA per matrix.
// ---------------------------------------------------------------------------------------------------------------
struct EigShared {
  double d[kEigMaxM], e[kEigMaxM], tau[kEigMaxM], v[kEigMaxM], w[kEigMaxM];
  double part[4][kEigMaxM];
  double red[32];
  double scal[4];
  int ctl[4];
  // ring of QL sweeps: rotation (c, s) per index, {mm, lo} per sweep, progress of the consumer warps
  double ring_c[kRing][kEigMaxM], ring_s[kRing][kEigMaxM];
  int ring_meta[kRing][2];
  volatile int cons[32];
};

__device__ __forceinline__ double block_sum(double x, double* red) {
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) x += __shfl_xor_sync(0xffffffffu, x, sh);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // red may still be read by the previous call
  if (lane == 0) red[warp] = x;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < (blockDim.x >> 5); ++i) s += red[i];
  return s;
}

__global__ void __launch_bounds__(kEigThreads, 1)
    sym_eig_kernel(double* __restrict__ A_all, int lda, int m, double* __restrict__ Z_all, double* __restrict__ w_all,
                   double* __restrict__ V_all, int ldv, int64_t a_stride, int64_t w_stride, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char eig_smem[];
  EigShared& S = *reinterpret_cast<EigShared*>(eig_smem);
  double* A = A_all + blockIdx.x * a_stride;
  double* Z = Z_all + blockIdx.x * a_stride;
  double* V = V_all + blockIdx.x * a_stride;
  double* wout = w_all + blockIdx.x * w_stride;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = kEigThreads / 32;
  const int tx = tid & 255, part = tid >> 8;  // 256 columns x 4 row parts for the matrix-vector products

  long long tq0 = clock64();
  // ---- phase 1: Householder tridiagonalisation A = Q T Q^T (full symmetric trailing block kept up to date)
  for (int j = 0; j + 1 < m; ++j) {
    const int L = m - j - 1, base = j + 1;
    double ss = 0.0;
    for (int r = tid + 1; r < L; r += kEigThreads) {
      const double x = A[int64_t(base + r) * lda + j];
      ss = fma(x, x, ss);
    }
    ss = block_sum(ss, S.red);  // |x[1:]|^2
    if (tid == 0) {
      const double x0 = A[int64_t(base) * lda + j];
      S.d[j] = A[int64_t(j) * lda + j];
      if (ss == 0.0) {
        S.tau[j] = 0.0, S.e[j] = x0, S.scal[0] = 0.0;
      } else {
        const double beta = -copysign(sqrt(fma(x0, x0, ss)), x0);
        S.tau[j] = (beta - x0) / beta;
        S.e[j] = beta;
        S.scal[0] = 1.0 / (x0 - beta);
      }
    }
    __syncthreads();
    const double tau = S.tau[j], scale = S.scal[0];
    if (tau == 0.0) continue;  // uniform
    for (int r = tid; r < L; r += kEigThreads) {
      const double vr = r == 0 ? 1.0 : A[int64_t(base + r) * lda + j] * scale;
      S.v[r] = vr;
      if (r) A[int64_t(base + r) * lda + j] = vr;  // kept for phase 2
    }
    __syncthreads();
    // p = tau A22 v  (A22 symmetric: p_i = sum_c A22[c][i] v[c], coalesced over i)
    {
      const int c0 = (L * part) / 4, c1 = (L * (part + 1)) / 4;
      for (int i = tx; i < L; i += 256) {
        double a0 = 0.0, a1 = 0.0;
        int c = c0;
        for (; c + 1 < c1; c += 2) {
          a0 = fma(A[int64_t(base + c) * lda + base + i], S.v[c], a0);
          a1 = fma(A[int64_t(base + c + 1) * lda + base + i], S.v[c + 1], a1);
        }
        if (c < c1) a0 = fma(A[int64_t(base + c) * lda + base + i], S.v[c], a0);
        S.part[part][i] = a0 + a1;
      }
    }
    __syncthreads();
    double pv = 0.0;
    for (int i = tid; i < L; i += kEigThreads) {
      const double p = tau * (S.part[0][i] + S.part[1][i] + S.part[2][i] + S.part[3][i]);
      S.w[i] = p;
      pv = fma(p, S.v[i], pv);
    }
    pv = block_sum(pv, S.red);
    const double kk = 0.5 * tau * pv;
    for (int i = tid; i < L; i += kEigThreads) S.w[i] -= kk * S.v[i];
    __syncthreads();
    // A22 -= v w^T + w v^T
    for (int r = warp; r < L; r += nwarp) {
      const double vr = S.v[r], wr = S.w[r];
      double* row = A + int64_t(base + r) * lda + base;
      for (int c = lane; c < L; c += 32) row[c] -= fma(vr, S.w[c], wr * S.v[c]);
    }
    __syncthreads();
  }
  if (tid == 0) {
    S.d[m - 1] = A[int64_t(m - 1) * lda + m - 1];
    S.e[m - 1] = 0.0;
    if (m >= 2) S.tau[m - 1] = 0.0;
  }
  long long tq1 = clock64();
  // ---- phase 2: Z = Q = H_0 H_1 ... H_{m-2}
  for (int idx = tid; idx < m * m; idx += kEigThreads) {
    const int r = idx / m, c = idx - r * m;
    Z[int64_t(r) * lda + c] = r == c ? 1.0 : 0.0;
  }
  __syncthreads();
  for (int j = m - 2; j >= 0; --j) {
    const double tau = S.tau[j];
    if (tau == 0.0) continue;
    const int L = m - j - 1, base = j + 1;
    for (int r = tid; r < L; r += kEigThreads) S.v[r] = r == 0 ? 1.0 : A[int64_t(base + r) * lda + j];
    __syncthreads();
    {
      const int r0 = (L * part) / 4, r1 = (L * (part + 1)) / 4;
      for (int c = tx; c < L; c += 256) {
        double a0 = 0.0, a1 = 0.0;
        int r = r0;
        for (; r + 1 < r1; r += 2) {
          a0 = fma(Z[int64_t(base + r) * lda + base + c], S.v[r], a0);
          a1 = fma(Z[int64_t(base + r + 1) * lda + base + c], S.v[r + 1], a1);
        }
        if (r < r1) a0 = fma(Z[int64_t(base + r) * lda + base + c], S.v[r], a0);
        S.part[part][c] = a0 + a1;
      }
    }
    __syncthreads();
    for (int c = tid; c < L; c += kEigThreads) S.w[c] = tau * (S.part[0][c] + S.part[1][c] + S.part[2][c] + S.part[3][c]);
    __syncthreads();
    for (int r = warp; r < L; r += nwarp) {
      const double vr = S.v[r];
      double* row = Z + int64_t(base + r) * lda + base;
      for (int c = lane; c < L; c += 32) row[c] = fma(-vr, S.w[c], row[c]);
    }
    __syncthreads();
  }
  long long tq2 = clock64();
  // ---- phase 3: A <- Z^T (row i of A = eigenvector direction i), implicit QL on (d, e) rotating rows of A
  for (int idx = tid; idx < m * m; idx += kEigThreads) {
    const int r = idx / m, c = idx - r * m;
    A[int64_t(c) * lda + r] = Z[int64_t(r) * lda + c];
  }
  __syncthreads();
  // The QL iteration is a scalar recurrence (one thread) that emits a sweep of plane rotations, and every rotation must be
  // applied to two rows of A (all columns).  Producer / consumer split: thread 0 runs the whole recurrence and publishes
  // sweeps into a ring of kRing slots in shared memory; warps 1.. apply them (thread k owns column k, a sequential chain
  // along the sweep) as they appear, so the recurrence never waits for the application (it did: 16.5 + 6.9 M cycles in turn).
  {
    volatile int* prod = &S.ctl[0];        // sweeps published
    volatile int* fin = &S.ctl[1];         // the recurrence has finished
    const int n_cw = (m + 31) / 32;        // consumer warps 1 .. n_cw own the m columns; the others wait at the barrier
    if (tid == 0) *prod = 0, *fin = 0;
    for (int w = tid; w < 32; w += kEigThreads) S.cons[w] = 0;
    __syncthreads();
    if (warp == 0) {
      if (lane == 0) {
        int n_pub = 0;
        for (int l = 0; l < m; ++l) {
          for (int iter = 0;; ++iter) {
            int mm = l;
            for (; mm < m - 1; ++mm) {
              const double dd = fabs(S.d[mm]) + fabs(S.d[mm + 1]);
              if (fabs(S.e[mm]) <= 2.220446049250313e-16 * dd) break;
            }
            if (mm == l) break;
            if (iter >= 80) {
              if (status) atomicExch(status, 2);  // no convergence: give up on this eigenvalue
              break;
            }
            // a free slot: every consumer warp is done with sweep n_pub - kRing
            if (n_pub >= kRing) {
              for (;;) {
                int lowest = 0x7fffffff;
                for (int w = 1; w <= n_cw; ++w) lowest = min(lowest, int(S.cons[w]));
                if (lowest > n_pub - kRing) break;
                __nanosleep(40);
              }
            }
            double* cs = S.ring_c[n_pub % kRing];
            double* sn = S.ring_s[n_pub % kRing];
            double g = (S.d[l + 1] - S.d[l]) / (2.0 * S.e[l]);
            double r = sqrt(fma(g, g, 1.0));
            g = S.d[mm] - S.d[l] + S.e[l] / (g + copysign(r, g));
            double s = 1.0, c = 1.0, pp = 0.0;
            int i = mm - 1;
            bool under = false;
            // e[i], d[i], d[i + 1] of the running step are carried in registers and the next pair is fetched at the top
            // of the step, before the dependent chain (the stores below would otherwise fence the shared-memory loads)
            double e_i = S.e[i], d_i = S.d[i], d_i1 = S.d[i + 1];
            for (; i >= l; --i) {
              const double e_n = i > l ? S.e[i - 1] : 0.0, d_n = i > l ? S.d[i - 1] : 0.0;
              const double f = s * e_i, bb = c * e_i;
              // r = hypot(f, g), s = f / r, c = g / r through one reciprocal square root (the float64 sqrt and the two
              // divisions of the textbook form were 2/3 of the whole solve: this chain runs on ONE thread)
              const double x = fma(f, f, g * g);
              if (x == 0.0) {
                S.e[i + 1] = 0.0;
                S.d[i + 1] = d_i1 - pp;
                S.e[mm] = 0.0;
                under = true;
                break;
              }
              const double rinv = rsqrt(x);
              S.e[i + 1] = x * rinv;
              s = f * rinv, c = g * rinv;
              g = d_i1 - pp;
              r = fma(d_i - g, s, 2.0 * c * bb);
              pp = s * r;
              S.d[i + 1] = g + pp;
              g = fma(c, r, -bb);
              cs[i] = c, sn[i] = s;
              d_i1 = d_i, d_i = d_n, e_i = e_n;
            }
            const int lo = i + 1;
            if (!under) {
              S.d[l] -= pp;
              S.e[l] = g;
              S.e[mm] = 0.0;
            }
            if (lo < mm) {
              S.ring_meta[n_pub % kRing][0] = mm, S.ring_meta[n_pub % kRing][1] = lo;
              __threadfence_block();
              *prod = ++n_pub;
            }
          }
        }
        __threadfence_block();
        *fin = 1;
      }
    } else if (warp <= n_cw) {
      int mine = 0;
      const int k = tid - 32;  // column of this consumer thread (m <= 512 < 992)
      for (;;) {
        if (*prod <= mine) {
          if (!*fin) {
            __nanosleep(200);  // (polling faster slows the recurrence: the pollers share its scheduler and shared memory)
            continue;
          }
          if (*prod <= mine) break;  // the recurrence has finished (its last publication precedes `fin`) and nothing is left
        }
        __threadfence_block();
        const int slot = mine % kRing;
        const int mm = S.ring_meta[slot][0], lo = S.ring_meta[slot][1];
        const double* cs = S.ring_c[slot];
        const double* sn = S.ring_s[slot];
        if (k < m) {
          double zi1 = A[int64_t(mm) * lda + k];
          int i = mm - 1;
          for (; i - 7 >= lo; i -= 8) {  // the eight loads first: they do not depend on the rotation chain
            double z[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) z[u] = A[int64_t(i - u) * lda + k];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const double c = cs[i - u], s = sn[i - u];
              A[int64_t(i - u + 1) * lda + k] = fma(s, z[u], c * zi1);
              zi1 = fma(c, z[u], -s * zi1);
            }
          }
          for (; i >= lo; --i) {
            const double zi = A[int64_t(i) * lda + k];
            const double c = cs[i], s = sn[i];
            A[int64_t(i + 1) * lda + k] = fma(s, zi, c * zi1);
            zi1 = fma(c, zi, -s * zi1);
          }
          A[int64_t(lo) * lda + k] = zi1;
        }
        ++mine;
        __syncwarp();
        if (lane == 0) S.cons[warp] = mine;
      }
    }
    __syncthreads();
  }
  if (tid == 0 && status && blockIdx.x == 0) {
    long long tq3 = clock64();
    status[8] = int((tq1 - tq0) >> 10), status[9] = int((tq2 - tq1) >> 10), status[10] = int((tq3 - tq2) >> 10);
  }
  // ---- ascending order
  for (int t = tid; t < m; t += kEigThreads) {
    const double dt = S.d[t];
    int rank = 0;
    for (int j = 0; j < m; ++j) rank += (S.d[j] < dt || (S.d[j] == dt && j < t)) ? 1 : 0;
    S.w[t] = double(rank);
    wout[rank] = dt;
  }
  __syncthreads();
  for (int t = warp; t < m; t += nwarp) {
    const int rank = int(S.w[t]);
    for (int r = lane; r < m; r += 32) V[int64_t(r) * ldv + rank] = A[int64_t(t) * lda + r];
  }
}

int sym_eig_launch(double* A, int lda, int m, double* Z, double* w, double* V, int ldv, int n_batch, int64_t a_stride,
                   int64_t w_stride, int* status, cudaStream_t st) {
  if (m < 1 || m > kEigMaxM) DM_FAIL(DM_ERR_UNSUPPORTED, "dense symmetric eigenproblem of size %d (limit %d)", m, kEigMaxM);
  static OncePerDevice once;
  if (once.first())
    DM_CUDA_OK(cudaFuncSetAttribute(sym_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(EigShared))));
  sym_eig_kernel<<<n_batch, kEigThreads, sizeof(EigShared), st>>>(A, lda, m, Z, w, V, ldv, a_stride, w_stride, status);
  DM_LAUNCH_OK("sym_eig_kernel");
  return DM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// operator S = A^-1/2 W A^-1/2 in CSR and its products with a dense block
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_nonneg_f64(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), static_cast<unsigned long long>(__double_as_longlong(v)));
}

// sval[e] = W[e] / sqrt(a_row a_col); bound[0] = max_i sum_j |S_ij| (Gershgorin)
__global__ void __launch_bounds__(256)
    csr_symmetrise_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                          const double* __restrict__ W, const double* __restrict__ mass, int n, double* __restrict__ sval,
                          double* __restrict__ dinv, double* __restrict__ bound) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const double di = rsqrt(mass[i]);
  if (lane == 0) dinv[i] = di;
  double s = 0.0;
  for (int64_t e = indptr[i] + lane; e < indptr[i + 1]; e += 32) {
    const double v = W[e] * di * rsqrt(mass[indices[e]]);
    sval[e] = v;
    s += fabs(v);
  }
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
  if (lane == 0) atomic_max_nonneg_f64(bound, s);
}

// out[i, :] = alpha (S Y)[i, :] + beta Y[i, :] + gamma Zold[i, :]   (out may alias Zold).  coef = {alpha, beta, gamma}
// on the device, so that a whole filter is enqueued without reading the interval back.  One warp per row.
__global__ void __launch_bounds__(256)
    spmm_cheb_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const double* __restrict__ sval,
                     int n, int m, int ld, const double* __restrict__ Y, const double* Zold, double* out,
                     const double* __restrict__ coef) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const double alpha = coef[0], beta = coef[1], gamma = coef[2];
  const int64_t e0 = indptr[i], e1 = indptr[i + 1];
  for (int c0 = 0; c0 < m; c0 += 128) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t e = e0; alpha != 0.0 && e < e1; ++e) {
      const double v = sval[e];
      const double* yr = Y + int64_t(indices[e]) * ld + c0;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (c0 + lane + 32 * t < m) acc[t] = fma(v, yr[lane + 32 * t], acc[t]);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int c = c0 + lane + 32 * t;
      if (c < m) {
        const int64_t o = int64_t(i) * ld + c;
        double r = alpha * acc[t];
        if (beta != 0.0) r = fma(beta, Y[o], r);
        if (gamma != 0.0) r = fma(gamma, Zold[o], r);
        out[o] = r;
      }
    }
  }
}

// deterministic start block: column 0 = A^1/2 1 (the null vector of S), the rest hashed uniform values in (-1, 1)
__global__ void __launch_bounds__(256) init_block_kernel(double* __restrict__ X, int n, int m, int ld, const double* __restrict__ mass) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= int64_t(n) * m) return;
  const int i = int(idx / m), c = int(idx % m);
  uint64_t h = uint64_t(idx) * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
  h ^= h >> 32, h *= 0xD6E8FEB86659FD93ull, h ^= h >> 32, h *= 0xD6E8FEB86659FD93ull, h ^= h >> 32;
  const double u = double(h >> 11) * (1.0 / 9007199254740992.0);
  X[int64_t(i) * ld + c] = c == 0 ? sqrt(mass[i]) : 2.0 * u - 1.0;
}

// scaled Chebyshev filter of degree deg on [a, b] with the lower end of the spectrum at 0 (Zhou & Saad's scaling keeps
// p(0) = 1): coefficient triples for the deg passes.  a = theta[m - 1] (largest Ritz value of the block), b = bound.
__global__ void cheb_coef_kernel(const double* __restrict__ theta, int m, const double* __restrict__ bound, int deg,
                                 double* __restrict__ coef) {
  if (threadIdx.x || blockIdx.x) return;
  const double b = bound[0] * 1.01;
  double a = theta[m - 1];
  if (!(a > 0.0)) a = 1e-3 * b;
  if (a > 0.999 * b) a = 0.999 * b;
  const double e = 0.5 * (b - a), c = 0.5 * (b + a);
  // p(0) / max |p| on [a, b] = T_d(c / e) = cosh(d acosh(c / e)): beyond ~1e6 the damped directions of the block drop
  // below the rounding level of the kept ones and the block loses rank (tiny meshes, where the block is most of the
  // space and a is close to b) -- the passes after d_eff are identity copies
  const double growth = acosh(c / e);
  int d_eff = growth > 0.0 ? int(14.5 / growth) : deg;
  d_eff = max(1, min(deg, d_eff));
  double sigma = e / (0.0 - c);
  const double tau = 2.0 / sigma;
  coef[0] = sigma / e, coef[1] = -c * sigma / e, coef[2] = 0.0;
  for (int i = 2; i <= deg; ++i) {
    if (i > d_eff) {
      coef[3 * (i - 1) + 0] = 0.0, coef[3 * (i - 1) + 1] = 1.0, coef[3 * (i - 1) + 2] = 0.0;
      continue;
    }
    const double sn = 1.0 / (tau - sigma);
    coef[3 * (i - 1) + 0] = 2.0 * sn / e;
    coef[3 * (i - 1) + 1] = -2.0 * c * sn / e;
    coef[3 * (i - 1) + 2] = -sigma * sn;
    sigma = sn;
  }
}

// column scaling of the Gram matrix: dsc[i] = 1 / sqrt(G_ii); G <- diag(dsc) G diag(dsc), symmetrised
__global__ void __launch_bounds__(256) svqb_scale_kernel(double* __restrict__ G, int m, int ld, double* __restrict__ dsc) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * m) return;
  const int r = idx / m, c = idx % m;
  if (c < r) return;
  const double gr = G[int64_t(r) * ld + r], gc = G[int64_t(c) * ld + c];
  const double dr = gr > 0.0 ? rsqrt(gr) : 0.0, dc = gc > 0.0 ? rsqrt(gc) : 0.0;
  const double v = 0.5 * (G[int64_t(r) * ld + c] + G[int64_t(c) * ld + r]) * dr * dc;
  // other threads read the ORIGINAL diagonal: only the off-diagonal entries are written here, the scaled diagonal
  // (exactly 1) by svqb_diag_kernel afterwards
  if (r != c) {
    G[int64_t(r) * ld + c] = v;
    G[int64_t(c) * ld + r] = v;
  } else {
    dsc[r] = dr;
  }
}
__global__ void __launch_bounds__(256) svqb_diag_kernel(double* __restrict__ G, int m, int ld, const double* __restrict__ dsc) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < m) G[int64_t(r) * ld + r] = dsc[r] > 0.0 ? 1.0 : 0.0;
}
// Cholesky-QR of the block: G' = diag(dsc) X^T X diag(dsc) (unit diagonal, svqb_scale_kernel) = L L^T in place (lower), then
// T = diag(dsc) L^-T so that (X T)^T (X T) = I.  One CTA: m <= 512 steps of "pivot, scale the column, rank-1 update of the
// trailing lower triangle".  A pivot below 1e-14 (the block lost rank) is clamped and reported in status[1].
__global__ void __launch_bounds__(1024, 1) chol_kernel(double* __restrict__ G, int m, int ld, int* __restrict__ status) {
  __shared__ double col[kEigMaxM];
  __shared__ double inv_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int j = 0; j < m; ++j) {
    if (tid == 0) {
      double p = G[int64_t(j) * ld + j];
      if (!(p > 1e-14)) {
        p = 1e-14;
        if (status) status[1] = 1;
      }
      const double piv = sqrt(p);
      G[int64_t(j) * ld + j] = piv;
      inv_s = 1.0 / piv;
    }
    __syncthreads();
    const double inv = inv_s;
    for (int r = j + 1 + tid; r < m; r += blockDim.x) {
      const double v = G[int64_t(r) * ld + j] * inv;
      G[int64_t(r) * ld + j] = v;
      col[r] = v;
    }
    __syncthreads();
    for (int r = j + 1 + warp; r < m; r += 32) {
      const double vr = col[r];
      double* row = G + int64_t(r) * ld;
      for (int c = j + 1 + lane; c <= r; c += 32) row[c] = fma(-vr, col[c], row[c]);
    }
    __syncthreads();
  }
}
// T[c][i] = dsc[c] (L^-1)[i][c]: one warp per column c of L^-1 (forward substitution, the running column kept in T's row c)
__global__ void __launch_bounds__(1024, 1)
    chol_transform_kernel(const double* __restrict__ L, const double* __restrict__ dsc, int m, int ld, double* __restrict__ T) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < m; c += 32) {
    double* z = T + int64_t(c) * ld;  // z[i] = (L^-1)[i][c], i >= c
    for (int i = lane; i < c; i += 32) z[i] = 0.0;
    if (lane == 0) z[c] = 1.0 / L[int64_t(c) * ld + c];
    __syncwarp();
    for (int r = c + 1; r < m; ++r) {
      const double* lr = L + int64_t(r) * ld;
      double acc = 0.0;
      for (int i = c + lane; i < r; i += 32) acc = fma(lr[i], z[i], acc);
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
      if (lane == 0) z[r] = -acc / lr[r];
      __syncwarp();
    }
    const double dc = dsc[c];
    for (int i = c + lane; i < m; i += 32) z[i] *= dc;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) symmetrise_kernel(double* __restrict__ H, int m, int ld) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * m) return;
  const int r = idx / m, c = idx % m;
  if (c <= r) return;
  const double v = 0.5 * (H[int64_t(r) * ld + c] + H[int64_t(c) * ld + r]);
  H[int64_t(r) * ld + c] = v;
  H[int64_t(c) * ld + r] = v;
}

// res[0] = max_{c < k} |SX[:, c] - theta_c X[:, c]| / max(theta_{k-1}, tiny)   (one CTA per column, atomic max)
__global__ void __launch_bounds__(256)
    residual_kernel(const double* __restrict__ X, const double* __restrict__ SX, const double* __restrict__ theta, int n, int ld,
                    int k, double* __restrict__ res) {
  __shared__ double red[8];
  const int c = blockIdx.x;
  const double th = theta[c];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double r = SX[int64_t(i) * ld + c] - th * X[int64_t(i) * ld + c];
    s = fma(r, r, s);
  }
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) s += red[w];
    atomic_max_nonneg_f64(res, sqrt(s) / fmax(fabs(theta[k - 1]), 1e-300));
  }
}

// evecs[i, c] = X[i, c] / sqrt(a_i), c < k;  evals[c] = theta[c]
__global__ void __launch_bounds__(256)
    eigs_output_kernel(const double* __restrict__ X, int ld, const double* __restrict__ theta, const double* __restrict__ dinv,
                       int n, int k, double* __restrict__ evecs, int64_t ld_out, double* __restrict__ evals) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx < k) evals[idx] = theta[idx];
  if (idx >= int64_t(n) * k) return;
  const int i = int(idx / k), c = int(idx % k);
  evecs[int64_t(i) * ld_out + c] = X[int64_t(i) * ld + c] * dinv[i];
}

// coef[mesh][k][c] *= exp(-evals[mesh][k] * t[c])
__global__ void __launch_bounds__(256)
    diffusion_scale_kernel(double* __restrict__ coef, const double* __restrict__ evals, const double* __restrict__ t, int n_meshes,
                           int k, int c) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= int64_t(n_meshes) * k * c) return;
  const int ch = int(idx % c);
  const int64_t mk = idx / c;
  coef[idx] *= exp(-evals[mk] * t[ch]);
}

int block_width(int n, int k) {
  int m = k + (k / 5 > 24 ? k / 5 : 24);
  m = (m + 7) & ~7;
  return m < n ? m : n;
}

struct EigsLayout {
  double *sval, *dinv, *X, *Xn, *SX, *SXn, *G, *H, *Zs, *V, *T, *theta, *lam, *dsc, *coef, *scal, *part;
  int* status;
  int m, ld, ks;
  size_t bytes;
};
constexpr int kGramChunk = 128;
EigsLayout eigs_carve(void* ws, int n, int64_t nnz, int k) {
  Carver c(ws);
  EigsLayout L{};
  L.m = block_width(n, k);
  L.ld = (L.m + 7) & ~7;
  L.ks = (n + kGramChunk - 1) / kGramChunk;
  L.status = c.take<int>(64);
  L.scal = c.take<double>(8);  // [0] gershgorin bound, [1] residual
  L.sval = c.take<double>(size_t(nnz));
  L.dinv = c.take<double>(size_t(n));
  const size_t blk = size_t(n) * L.ld, sq = size_t(L.ld) * L.ld;
  L.X = c.take<double>(blk), L.Xn = c.take<double>(blk), L.SX = c.take<double>(blk), L.SXn = c.take<double>(blk);
  L.G = c.take<double>(sq), L.H = c.take<double>(sq), L.Zs = c.take<double>(sq), L.V = c.take<double>(sq), L.T = c.take<double>(sq);
  L.theta = c.take<double>(size_t(L.ld)), L.lam = c.take<double>(size_t(L.ld)), L.dsc = c.take<double>(size_t(L.ld));
  L.coef = c.take<double>(3 * (kMaxDegree + 1));
  L.part = c.take<double>(size_t(L.ks) * sq);
  L.bytes = c.bytes();
  return L;
}

// out (m x m, pitch ld) = P^T Q for two n x m blocks (split over the vertices, deterministic reduction)
int gram(const double* P, const double* Q, int n, int m, int ld, const EigsLayout& L, double* out, cudaStream_t st) {
  GemmProblem G;
  G.A.d = P, G.A.ld = ld, G.A.rows = n, G.A.trans = 1;
  G.B.d = Q, G.B.ld = ld, G.B.rows = n, G.B.trans = 1;
  G.M = m, G.N = m, G.K = n, G.maxM = m, G.maxN = m, G.maxK = n, G.n_batch = 1;
  G.ldc = ld, G.c_batch_stride = int64_t(ld) * ld;
  int rc;
  if (L.ks <= 1) {
    G.C = out;
    return gemm64_launch(G, st);
  }
  G.C = L.part, G.ksplit = L.ks, G.kchunk = kGramChunk, G.split_stride = int64_t(ld) * ld;
  if ((rc = gemm64_launch(G, st))) return rc;
  return sum_partials_launch(L.part, L.ks, G.split_stride, int64_t(ld) * ld, out, st);
}

// out (n x m) = P (n x m) T (m x m)
int times_small(const double* P, const double* T, int n, int m, int ld, double* out, cudaStream_t st) {
  GemmProblem G;
  G.A.d = P, G.A.ld = ld, G.A.rows = n, G.A.trans = 0;
  G.B.d = T, G.B.ld = ld, G.B.rows = m, G.B.trans = 1;
  G.M = n, G.N = m, G.K = m, G.maxM = n, G.maxN = m, G.maxK = m, G.n_batch = 1;
  G.C = out, G.ldc = ld;
  return gemm64_launch(G, st);
}

}  // namespace
}  // namespace dm

using namespace dm;

extern "C" {

size_t dm_lbo_eigs_workspace_bytes(int n, int64_t nnz, int k) {
  if (n <= 0 || nnz < 0 || k <= 0 || k > n) return 0;
  return eigs_carve(nullptr, n, nnz, k).bytes;
}

int dm_lbo_eigs(const int64_t* indptr, const int32_t* indices, const double* values, int64_t nnz, const double* mass, int n,
                int k, double tol, int max_iter, int degree, double* evals, double* evecs, int64_t ld_evecs,
                int* info_h /* [4] host: iterations, converged, block width, status */, double* residual_h, void* workspace,
                size_t workspace_bytes, dm_stream_t stream) {
  if (n <= 0 || k <= 0 || k > n || nnz < 0 || max_iter < 1) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (!indptr || !indices || !values || !mass || !evals || !evecs) DM_FAIL(DM_ERR_BADARG, "null argument");
  if (ld_evecs < k) DM_FAIL(DM_ERR_BADARG, "leading dimension too small");
  if (block_width(n, k) > kEigMaxM) DM_FAIL(DM_ERR_UNSUPPORTED, "k = %d needs a block of %d columns (limit %d)", k, block_width(n, k), kEigMaxM);
  EigsLayout L = eigs_carve(workspace, n, nnz, k);
  if (!workspace || L.bytes > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu", L.bytes);
  if (reinterpret_cast<uintptr_t>(workspace) % 256) DM_FAIL(DM_ERR_ALIGN, "workspace must be 256-byte aligned");
  if (degree <= 0) degree = 24;
  if (degree > kMaxDegree) degree = kMaxDegree;
  if (!(tol > 0.0)) tol = 1e-10;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int m = L.m, ld = L.ld;
  const unsigned row_blocks = unsigned((n + 7) / 8), sq_blocks = unsigned((m * m + 255) / 256);
  int rc;
  DM_CUDA_OK(cudaMemsetAsync(L.status, 0, sizeof(int) * 64, st));
  DM_CUDA_OK(cudaMemsetAsync(L.scal, 0, sizeof(double) * 8, st));
  const size_t blk = size_t(n) * ld;
  for (double* b : {L.X, L.Xn, L.SX, L.SXn}) DM_CUDA_OK(cudaMemsetAsync(b, 0, sizeof(double) * blk, st));  // zero padding columns
  csr_symmetrise_kernel<<<row_blocks, 256, 0, st>>>(indptr, indices, values, mass, n, L.sval, L.dinv, L.scal);
  DM_LAUNCH_OK("csr_symmetrise_kernel");
  init_block_kernel<<<unsigned((int64_t(n) * m + 255) / 256), 256, 0, st>>>(L.X, n, m, ld, mass);
  DM_LAUNCH_OK("init_block_kernel");
  const double plain[3] = {1.0, 0.0, 0.0};
  DM_CUDA_OK(cudaMemcpyAsync(L.coef + 3 * kMaxDegree, plain, sizeof(plain), cudaMemcpyHostToDevice, st));

  auto orthonormalise = [&](double* X, double* out) -> int {  // out = X T with out^T out = I (Cholesky-QR, scaled columns)
    int r;
    if ((r = gram(X, X, n, m, ld, L, L.G, st))) return r;
    svqb_scale_kernel<<<sq_blocks, 256, 0, st>>>(L.G, m, ld, L.dsc);
    svqb_diag_kernel<<<unsigned((m + 255) / 256), 256, 0, st>>>(L.G, m, ld, L.dsc);
    chol_kernel<<<1, 1024, 0, st>>>(L.G, m, ld, L.status);
    chol_transform_kernel<<<1, 1024, 0, st>>>(L.G, L.dsc, m, ld, L.T);
    DM_LAUNCH_OK("chol_kernel");
    return times_small(X, L.T, n, m, ld, out, st);
  };

  double* X = L.X;    // current block
  double* Xn = L.Xn;  // scratch / rotated block
  if ((rc = orthonormalise(X, Xn))) return rc;
  { double* t = X; X = Xn; Xn = t; }
  int it = 0, converged = 0;
  double res = 0.0;
  for (;; ++it) {
    // Rayleigh-Ritz
    spmm_cheb_kernel<<<row_blocks, 256, 0, st>>>(indptr, indices, L.sval, n, m, ld, X, nullptr, L.SX, L.coef + 3 * kMaxDegree);
    DM_LAUNCH_OK("spmm_cheb_kernel");
    if ((rc = gram(X, L.SX, n, m, ld, L, L.H, st))) return rc;
    symmetrise_kernel<<<sq_blocks, 256, 0, st>>>(L.H, m, ld);
    DM_LAUNCH_OK("symmetrise_kernel");
    if ((rc = sym_eig_launch(L.H, ld, m, L.Zs, L.theta, L.V, ld, 1, 0, 0, L.status, st))) return rc;
    if ((rc = times_small(X, L.V, n, m, ld, Xn, st))) return rc;
    if ((rc = times_small(L.SX, L.V, n, m, ld, L.SXn, st))) return rc;
    DM_CUDA_OK(cudaMemsetAsync(L.scal + 1, 0, sizeof(double), st));
    residual_kernel<<<k, 256, 0, st>>>(Xn, L.SXn, L.theta, n, ld, k, L.scal + 1);
    DM_LAUNCH_OK("residual_kernel");
    DM_CUDA_OK(cudaMemcpyAsync(&res, L.scal + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
    DM_CUDA_OK(cudaStreamSynchronize(st));
    { double* t = X; X = Xn; Xn = t; }
    if (res <= tol) {
      converged = 1;
      break;
    }
    if (it + 1 >= max_iter) break;
    // filter: X <- p(S) X   (Y_1 = a1 S X + b1 X;  Y_i = a_i S Y_{i-1} + b_i Y_{i-1} + g_i Y_{i-2})
    cheb_coef_kernel<<<1, 32, 0, st>>>(L.theta, m, L.scal, degree, L.coef);
    DM_LAUNCH_OK("cheb_coef_kernel");
    double* y0 = X;
    double* y1 = Xn;
    spmm_cheb_kernel<<<row_blocks, 256, 0, st>>>(indptr, indices, L.sval, n, m, ld, y0, nullptr, y1, L.coef);
    for (int i = 2; i <= degree; ++i) {
      spmm_cheb_kernel<<<row_blocks, 256, 0, st>>>(indptr, indices, L.sval, n, m, ld, y1, y0, y0, L.coef + 3 * (i - 1));
      double* t = y0; y0 = y1; y1 = t;
    }
    DM_LAUNCH_OK("spmm_cheb_kernel");
    // y1 holds the filtered block; orthonormalise into the other buffer
    if ((rc = orthonormalise(y1, y0))) return rc;
    X = y0, Xn = y1;
  }
  eigs_output_kernel<<<unsigned((int64_t(n) * k + 255) / 256), 256, 0, st>>>(X, ld, L.theta, L.dinv, n, k, evecs, ld_evecs, evals);
  DM_LAUNCH_OK("eigs_output_kernel");
  int status[2] = {0, 0};  // [0] dense eigensolver (2 = QL did not converge), [1] Cholesky-QR met a rank-deficient block
  DM_CUDA_OK(cudaMemcpyAsync(status, L.status, sizeof(status), cudaMemcpyDeviceToHost, st));
  DM_CUDA_OK(cudaStreamSynchronize(st));
  if (info_h) info_h[0] = it + 1, info_h[1] = converged, info_h[2] = m, info_h[3] = status[0] | (status[1] << 4);
  if (residual_h) *residual_h = res;
  return DM_OK;
}

// batched dense symmetric eigen-decomposition (float64): A [n_batch, m, m] -> w [n_batch, m] ascending, V [n_batch, m, m]
// (columns = eigenvectors).  The small eigenproblems of the subspace iteration, exposed for tests.
size_t dm_sym_eig_workspace_bytes(int n_batch, int m) {
  if (n_batch <= 0 || m <= 0) return 0;
  Carver c(nullptr);
  c.take<int>(64);
  c.take<double>(size_t(n_batch) * m * m);
  c.take<double>(size_t(n_batch) * m * m);
  return c.bytes();
}

int dm_sym_eig(const double* A, int m, int n_batch, double* w, double* V, void* workspace, size_t workspace_bytes,
               dm_stream_t stream) {
  if (m <= 0 || n_batch < 0) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n_batch == 0) return DM_OK;
  if (!A || !w || !V) DM_FAIL(DM_ERR_BADARG, "null argument");
  if (!workspace || dm_sym_eig_workspace_bytes(n_batch, m) > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver c(workspace);
  int* status = c.take<int>(64);
  double* Ac = c.take<double>(size_t(n_batch) * m * m);
  double* Z = c.take<double>(size_t(n_batch) * m * m);
  DM_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int) * 64, st));
  DM_CUDA_OK(cudaMemcpyAsync(Ac, A, sizeof(double) * size_t(n_batch) * m * m, cudaMemcpyDeviceToDevice, st));
  return sym_eig_launch(Ac, m, m, Z, w, V, m, n_batch, int64_t(m) * m, m, status, st);
}

// out[rows of mesh b, :c] = Phi_b[:, :k] coef[b]   (coef [n_meshes, k, c] float64; geometry.py:586-598)
int dm_from_basis(const double* coef, const double* Phi, int64_t ldPhi, const int64_t* row_off, int max_n, int n_meshes, int k,
                  int c, double* out, int64_t ld_out, dm_stream_t stream) {
  if (n_meshes < 0 || k <= 0 || c <= 0 || max_n < 0) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n_meshes == 0 || max_n == 0) return DM_OK;
  if (!coef || !Phi || !row_off || !out) DM_FAIL(DM_ERR_BADARG, "null argument");
  if (ldPhi < k || ld_out < c) DM_FAIL(DM_ERR_BADARG, "leading dimension too small");
  GemmProblem G;
  G.A.d = Phi, G.A.ld = ldPhi, G.A.off = row_off, G.A.trans = 0;
  G.B.d = coef, G.B.ld = c, G.B.batch_stride = int64_t(k) * c, G.B.rows = k, G.B.trans = 1;
  G.N = c, G.K = k, G.maxM = max_n, G.maxN = c, G.maxK = k, G.n_batch = n_meshes;
  G.C = out, G.ldc = ld_out, G.c_off = row_off;
  return gemm64_launch(G, static_cast<cudaStream_t>(stream));
}

size_t dm_spectral_diffusion_workspace_bytes(int n_meshes, int64_t total_n, int max_n, int k, int c) {
  if (n_meshes < 0 || total_n < 0 || k <= 0 || c <= 0) return 0;
  Carver cv(nullptr);
  cv.take<double>(size_t(n_meshes) * k * c);
  cv.take<char>(dm_project_workspace_bytes(n_meshes, total_n, max_n, k, c));
  return cv.bytes();
}

// LearnedTimeDiffusion.forward, method 'spectral' (diffusion_net/layers.py:56-67):
//   out = Phi (exp(-evals t) * (Phi^T diag(mass) X))
int dm_spectral_diffusion(const double* Phi, int64_t ldPhi, const double* mass, const double* evals, const float* X, int64_t ldX,
                          const double* time, const int64_t* row_off, int64_t total_n, int max_n, int n_meshes, int k, int c,
                          double* out, int64_t ld_out, int flags, void* workspace, size_t workspace_bytes, dm_stream_t stream) {
  if (n_meshes < 0 || k <= 0 || c <= 0 || total_n < 0 || max_n < 0) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n_meshes == 0 || total_n == 0) return DM_OK;
  if (!Phi || !mass || !evals || !X || !time || !row_off || !out) DM_FAIL(DM_ERR_BADARG, "null argument");
  const size_t need = dm_spectral_diffusion_workspace_bytes(n_meshes, total_n, max_n, k, c);
  if (!workspace || need > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu", need);
  if (reinterpret_cast<uintptr_t>(workspace) % 256) DM_FAIL(DM_ERR_ALIGN, "workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver cv(workspace);
  double* coef = cv.take<double>(size_t(n_meshes) * k * c);
  const size_t pw = dm_project_workspace_bytes(n_meshes, total_n, max_n, k, c);
  void* proj_ws = cv.take<char>(pw);
  int rc;
  if ((rc = dm_project_ex(Phi, ldPhi, mass, X, ldX, row_off, total_n, max_n, n_meshes, k, c, coef, flags, proj_ws, pw, stream)))
    return rc;
  const int64_t tot = int64_t(n_meshes) * k * c;
  diffusion_scale_kernel<<<unsigned((tot + 255) / 256), 256, 0, st>>>(coef, evals, time, n_meshes, k, c);
  DM_LAUNCH_OK("diffusion_scale_kernel");
  return dm_from_basis(coef, Phi, ldPhi, row_off, max_n, n_meshes, k, c, out, ld_out, stream);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Farthest point sampling, Euclidean (TriMesh.extract_fps(geodesic=False), mesh/trimesh.py:870-876 ->
// geometry.farthest_point_sampling_call, mesh/geometry.py:813-851): idx[0] = first, then size - 1 times
// "argmax of the running distances, minimum with the distances to the new point".  One CTA per mesh; distances
// sqrt((dx^2 + dy^2) + dz^2) rounded exactly like numpy's norm (no fused multiply-add), argmax ties to the lowest index.
// ---------------------------------------------------------------------------------------------------------------
namespace dm {
namespace {
__global__ void __launch_bounds__(1024, 1)
    fps_kernel(const double* __restrict__ V, const int64_t* __restrict__ off, const int64_t* __restrict__ first, int size,
               int64_t* __restrict__ out, double* __restrict__ dist_all) {
  __shared__ double red_v[32];
  __shared__ int red_i[32];
  __shared__ int cur_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t r0 = off[b];
  const int n = int(off[b + 1] - r0);
  const double* Vb = V + 3 * r0;
  double* dist = dist_all + r0;
  int64_t* ob = out + int64_t(b) * size;
  int cur = int(first[b]);
  if (cur < 0 || cur >= n) cur = 0;
  for (int it = 0; it < size; ++it) {
    if (tid == 0) ob[it] = cur;
    if (it + 1 == size) break;
    const double cx = Vb[3 * cur], cy = Vb[3 * cur + 1], cz = Vb[3 * cur + 2];
    double best = -1.0;
    int bi = 0x7fffffff;
    for (int i = tid; i < n; i += blockDim.x) {
      const double dx = Vb[3 * i] - cx, dy = Vb[3 * i + 1] - cy, dz = Vb[3 * i + 2] - cz;
      const double d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
      const double m = it == 0 ? d : fmin(dist[i], d);
      dist[i] = m;
      if (m > best) best = m, bi = i;  // ascending i inside a thread: strict '>' keeps the lowest index
    }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, sh);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, sh);
      if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
    }
    if (lane == 0) red_v[warp] = best, red_i[warp] = bi;
    __syncthreads();
    if (warp == 0) {
      best = red_v[lane], bi = red_i[lane];
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, sh);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, sh);
        if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
      }
      if (lane == 0) cur_s = bi;
    }
    __syncthreads();
    cur = cur_s;
  }
}
}  // namespace
}  // namespace dm

extern "C" {

size_t dm_fps_workspace_bytes(int64_t total_n) {
  if (total_n < 0) return 0;
  Carver c(nullptr);
  c.take<double>(size_t(total_n));
  return c.bytes();
}

int dm_fps(const double* verts, const int64_t* row_off, int64_t total_n, int n_meshes, const int64_t* first, int size,
           int64_t* idx, void* workspace, size_t workspace_bytes, dm_stream_t stream) {
  if (n_meshes < 0 || size < 0 || total_n < 0) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n_meshes == 0 || size == 0) return DM_OK;
  if (!verts || !row_off || !first || !idx) DM_FAIL(DM_ERR_BADARG, "null argument");
  if (!workspace || dm_fps_workspace_bytes(total_n) > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small");
  Carver c(workspace);
  double* dist = c.take<double>(size_t(total_n));
  fps_kernel<<<n_meshes, 1024, 0, static_cast<cudaStream_t>(stream)>>>(verts, row_off, first, size, idx, dist);
  DM_LAUNCH_OK("fps_kernel");
  return DM_OK;
}

}  // extern "C"
