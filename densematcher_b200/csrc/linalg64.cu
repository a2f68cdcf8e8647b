// Batched float64 dense linear algebra of the spectral ICP step, one CTA per mesh pair.
//
//   spd_inverse : G^-1 of the Gram matrix Phi2^T Phi2 (Cholesky + two triangular solves per unit vector); it turns the
//                 least squares of icp.py:38 / convert.py:51, lstsq(Phi2, Phi1[p]), into one contraction
//                 (Phi2 G^-1)^T Phi1[p] per iteration (differs from LAPACK gelsd by ~3e-15).
//   polar_factor: U I V^T of the SVD (icp.py:39-40) through a one-sided Jacobi iteration with a round-robin
//                 ordering: every warp orthogonalises one column pair per step.
//
// The matrices live in shared memory when they fit (k up to ~110 for the polar factor) and in an L2-resident
// global scratch otherwise; the code is the same through generic pointers.
#include "linalg64.cuh"
#include "gemm64.cuh"

namespace dm {
namespace {

constexpr size_t kSmemBudget = 200 * 1024;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
  return v;
}

// ------------------------------------------------------------------------------------------ SPD inverse
__global__ void __launch_bounds__(256)
    spd_inverse_kernel(const double* __restrict__ G, double* __restrict__ Ginv, int n, double* scratch, size_t per,
                       int use_smem, int* status) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, t = threadIdx.x;
  const int ldl = n + 1;
  double* L = use_smem ? sm : scratch + size_t(b) * per;  // [n][ldl]
  double* Z = L + size_t(n) * ldl;                         // [n][n]
  const double* g = G + size_t(b) * n * n;
  __shared__ int s_bad;
  if (t == 0) s_bad = 0;
  for (int e = t; e < n * n; e += blockDim.x) L[(e / n) * ldl + (e % n)] = g[e];
  __syncthreads();
  for (int j = 0; j < n; ++j) {  // right-looking Cholesky, lower triangle
    if (t == 0) {
      const double dj = L[j * ldl + j];
      if (!(dj > 0.0)) s_bad = 1;
      L[j * ldl + j] = sqrt(dj);
    }
    __syncthreads();
    const double inv = 1.0 / L[j * ldl + j];
    for (int r = j + 1 + t; r < n; r += blockDim.x) L[r * ldl + j] *= inv;
    __syncthreads();
    const int m = n - j - 1;
    for (int r = j + 1 + (t >> 4); r < n; r += (blockDim.x >> 4)) {
      const double lr = L[r * ldl + j];
      for (int c = j + 1 + (t & 15); c <= r; c += 16) L[r * ldl + c] = fma(-lr, L[c * ldl + j], L[r * ldl + c]);
    }
    (void)m;
    __syncthreads();
  }
  // column c of the inverse: L y = e_c, L^T x = y   (thread c; reads of L are warp broadcasts)
  for (int c = t; c < n; c += blockDim.x) {
    for (int j = 0; j < c; ++j) Z[j * n + c] = 0.0;
    for (int j = c; j < n; ++j) {
      double s = (j == c) ? 1.0 : 0.0;
      for (int i = c; i < j; ++i) s = fma(-L[j * ldl + i], Z[i * n + c], s);
      Z[j * n + c] = s / L[j * ldl + j];
    }
    for (int j = n - 1; j >= 0; --j) {
      double s = Z[j * n + c];
      for (int i = j + 1; i < n; ++i) s = fma(-L[i * ldl + j], Z[i * n + c], s);
      Z[j * n + c] = s / L[j * ldl + j];
    }
  }
  __syncthreads();
  double* out = Ginv + size_t(b) * n * n;
  // symmetrise: the two triangles agree to rounding; average them so that Ginv is exactly symmetric
  for (int e = t; e < n * n; e += blockDim.x) {
    const int r = e / n, c = e % n;
    out[e] = 0.5 * (Z[r * n + c] + Z[c * n + r]);
  }
  if (t == 0 && s_bad) atomicExch(status, 1);
}

// ------------------------------------------------------------------------------------------ polar factor
__global__ void __launch_bounds__(1024)
    polar_kernel(const double* X, double* C, int rows, int cols, double* scratch, size_t per, int use_smem,
                 const int* __restrict__ need) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarps = blockDim.x >> 5;
  if (need && need[b] == 0) return;  // the Newton-Schulz iteration already delivered this one
  const bool tr = rows < cols;  // iterate on X^T so that the orthogonalised columns are the long ones
  const int m = tr ? cols : rows, n = tr ? rows : cols;
  double* W = use_smem ? sm : scratch + size_t(b) * per;  // [n][m] columns of the working matrix
  double* J = W + size_t(n) * m;                           // [n][n] columns of the accumulated rotations
  double* sig = J + size_t(n) * n;                         // [n]
  const double* x = X + size_t(b) * rows * cols;
  double* cout = C + size_t(b) * rows * cols;
  __shared__ int s_rot;
  for (int e = t; e < n * m; e += blockDim.x) {
    const int c = e / m, r = e % m;
    W[e] = tr ? x[size_t(c) * cols + r] : x[size_t(r) * cols + c];
  }
  for (int e = t; e < n * n; e += blockDim.x) J[e] = (e / n == e % n) ? 1.0 : 0.0;
  if (t == 0) s_rot = 0;
  __syncthreads();
  const int np = (n + 1) & ~1;
  const double tol = sqrt(double(m)) * 2.220446049250313e-16;
  for (int sweep = 0; sweep < 40 && n > 1; ++sweep) {
    for (int step = 0; step < np - 1; ++step) {
      for (int pr = warp; pr < np / 2; pr += nwarps) {
        int p, q;
        if (pr == 0) {
          p = np - 1;
          q = step;
        } else {
          p = (step + pr) % (np - 1);
          q = (step + np - 1 - pr) % (np - 1);
        }
        if (p >= n || q >= n) continue;  // the bye of an odd column count
        if (p > q) {
          const int s = p;
          p = q;
          q = s;
        }
        double* wp = W + size_t(p) * m;
        double* wq = W + size_t(q) * m;
        double al = 0.0, be = 0.0, ga = 0.0;
        for (int r = lane; r < m; r += 32) {
          const double a = wp[r], c = wq[r];
          al = fma(a, a, al);
          be = fma(c, c, be);
          ga = fma(a, c, ga);
        }
        al = warp_sum(al), be = warp_sum(be), ga = warp_sum(ga);
        if (ga != 0.0 && fabs(ga) > tol * sqrt(al * be)) {
          const double zeta = (be - al) / (2.0 * ga);
          const double tt = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
          for (int r = lane; r < m; r += 32) {
            const double a = wp[r], c = wq[r];
            wp[r] = cs * a - sn * c;
            wq[r] = sn * a + cs * c;
          }
          double* jp = J + size_t(p) * n;
          double* jq = J + size_t(q) * n;
          for (int r = lane; r < n; r += 32) {
            const double a = jp[r], c = jq[r];
            jp[r] = cs * a - sn * c;
            jq[r] = sn * a + cs * c;
          }
          if (lane == 0) s_rot = 1;
        }
      }
      __syncthreads();
    }
    const int rot = s_rot;
    __syncthreads();
    if (!rot) break;
    if (t == 0) s_rot = 0;
    __syncthreads();
  }
  for (int i = warp; i < n; i += nwarps) {
    double s = 0.0;
    for (int r = lane; r < m; r += 32) s = fma(W[size_t(i) * m + r], W[size_t(i) * m + r], s);
    s = warp_sum(s);
    if (lane == 0) sig[i] = s > 0.0 ? 1.0 / sqrt(s) : 0.0;
  }
  __syncthreads();
  // P = sum_i (w_i / sigma_i) j_i^T  (m x n);  C = P or P^T
  for (int e = t; e < m * n; e += blockDim.x) {
    const int r = e % m, c = e / m;
    double s = 0.0;
    for (int i = 0; i < n; ++i) s = fma(W[size_t(i) * m + r] * sig[i], J[size_t(i) * n + c], s);
    if (tr)
      cout[size_t(c) * cols + r] = s;
    else
      cout[size_t(r) * cols + c] = s;
  }
}

// ------------------------------------------------------------------------------------------ Newton-Schulz polar
// Scaling from the first Gram matrix Z = X^T X (or X X^T): sigma_max^2 <= |Z|_inf, so A0 = X / sqrt(|Z|_inf) has all
// singular values in (0, 1] -- and for the near-isometries ICP produces the bound is tight (|Z|_inf ~ 1), which is
// what keeps the iteration count small.  Z is rescaled in place to the Gram matrix of A0.
__global__ void __launch_bounds__(256)
    ns_scale_kernel(const double* __restrict__ X, double* __restrict__ Z, double* __restrict__ A, int n, int rows, int cols) {
  const double* x = X + size_t(blockIdx.x) * rows * cols;
  double* a = A + size_t(blockIdx.x) * rows * cols;
  double* z = Z + size_t(blockIdx.x) * n * n;
  __shared__ double red[256];
  __shared__ double s_inv2;
  const int t = threadIdx.x;
  double m = 0.0;
  for (int r = t; r < n; r += 256) {
    double s = 0.0;
    for (int c = 0; c < n; ++c) s += fabs(z[size_t(r) * n + c]);
    m = fmax(m, s);
  }
  red[t] = m;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if (t < h) red[t] = fmax(red[t], red[t + h]);
    __syncthreads();
  }
  if (t == 0) s_inv2 = red[0] > 0.0 ? 1.0 / red[0] : 0.0;
  __syncthreads();
  const double inv2 = s_inv2, inv = sqrt(inv2);
  for (int e = t; e < rows * cols; e += 256) a[e] = x[e] * inv;
  for (int e = t; e < n * n; e += 256) z[e] *= inv2;
}

// One Newton-Schulz bookkeeping step per batch entry (one CTA each): if the Gram matrix Z of the current iterate is
// the identity to 1e-13 the iterate is final -- it is copied to the output and the entry is marked done, so that every
// later kernel skips it; otherwise Z is replaced by W = 1.5 I - 0.5 Z for the update A <- A W.
__global__ void __launch_bounds__(256)
    ns_step_kernel(double* __restrict__ Z, const double* __restrict__ A, double* __restrict__ C, int n, int rows, int cols,
                   int* __restrict__ done) {
  if (done[blockIdx.x]) return;
  double* z = Z + size_t(blockIdx.x) * n * n;
  __shared__ double red[256];
  const int t = threadIdx.x;
  double m = 0.0;
  for (int e = t; e < n * n; e += 256) {
    const double v = fabs(z[e] - ((e / n == e % n) ? 1.0 : 0.0));
    m = (v == v) ? fmax(m, v) : INFINITY;  // NaN -> never converged
  }
  red[t] = m;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if (t < h) red[t] = fmax(red[t], red[t + h]);
    __syncthreads();
  }
  if (red[0] < 1e-13) {
    const double* a = A + size_t(blockIdx.x) * rows * cols;
    double* c = C + size_t(blockIdx.x) * rows * cols;
    for (int e = t; e < rows * cols; e += 256) c[e] = a[e];
    if (t == 0) done[blockIdx.x] = 1;
  } else {
    for (int e = t; e < n * n; e += 256) z[e] = ((e / n == e % n) ? 1.5 : 0.0) - 0.5 * z[e];
  }
}

__global__ void ns_need_kernel(const int* __restrict__ done, int* __restrict__ need, int n_batch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_batch) need[i] = done[i] ? 0 : 1;
}

size_t spd_doubles(int n) { return size_t(n) * (n + 1) + size_t(n) * n; }
size_t polar_doubles(int rows, int cols) {
  const int m = rows > cols ? rows : cols, n = rows > cols ? cols : rows;
  return size_t(n) * m + size_t(n) * n + size_t(n);
}

}  // namespace

size_t spd_inverse_scratch_doubles(int n) { return spd_doubles(n) * 8 > kSmemBudget ? spd_doubles(n) : 0; }
size_t polar_scratch_doubles(int rows, int cols) {
  return polar_doubles(rows, cols) * 8 > kSmemBudget ? polar_doubles(rows, cols) : 0;
}

int spd_inverse_launch(const double* G, double* Ginv, int n, int n_batch, double* scratch, int* status, cudaStream_t st) {
  if (n_batch <= 0 || n <= 0) return DM_OK;
  const size_t per = spd_doubles(n), bytes = per * 8;
  const int use_smem = bytes <= kSmemBudget;
  if (!use_smem && !scratch) DM_FAIL(DM_ERR_WORKSPACE, "spd_inverse: scratch missing");
  if (use_smem && bytes > 48 * 1024)
    DM_CUDA_OK(cudaFuncSetAttribute(spd_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBudget)));
  spd_inverse_kernel<<<n_batch, 256, use_smem ? bytes : 0, st>>>(G, Ginv, n, scratch, per, use_smem, status);
  DM_LAUNCH_OK("spd_inverse_kernel");
  return DM_OK;
}

int polar_jacobi_launch(const double* X, double* C, int rows, int cols, int n_batch, double* scratch, const int* need,
                        cudaStream_t st) {
  if (n_batch <= 0 || rows <= 0 || cols <= 0) return DM_OK;
  const size_t per = polar_doubles(rows, cols), bytes = per * 8;
  const int use_smem = bytes <= kSmemBudget;
  if (!use_smem && !scratch) DM_FAIL(DM_ERR_WORKSPACE, "polar_factor: scratch missing");
  if (use_smem && bytes > 48 * 1024)
    DM_CUDA_OK(cudaFuncSetAttribute(polar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBudget)));
  const int n = rows < cols ? rows : cols;
  const int threads = n >= 48 ? 1024 : (n >= 16 ? 256 : 64);
  polar_kernel<<<n_batch, threads, use_smem ? bytes : 0, st>>>(X, C, rows, cols, scratch, per, use_smem, need);
  DM_LAUNCH_OK("polar_kernel");
  return DM_OK;
}

// Scratch of the Newton-Schulz stage in doubles: two iterates + the Gram matrix + the fallback flags
size_t polar_ns_scratch_doubles(int rows, int cols, int n_batch) {
  const int n = rows < cols ? rows : cols;
  return size_t(n_batch) * (2 * size_t(rows) * cols + size_t(n) * n) + size_t(n_batch) + 8;
}

// U I V^T of X = U S V^T.  Fast path: Newton-Schulz  A <- A (1.5 I - 0.5 A^T A)  on batched float64 GEMMs (quadratic
// convergence; entries leave the iteration as soon as they are orthonormal to 1e-13, up to kNsIters steps).  Every matrix whose
// iterate is not orthonormal to 1e-12 afterwards (ill-conditioned or rank-deficient input) is redone by the one-sided
// Jacobi SVD, which needs no conditioning assumption.
constexpr int kNsIters = 24;

int polar_factor_launch(const double* X, double* C, int rows, int cols, int n_batch, double* scratch, double* ns_scratch,
                        cudaStream_t st) {
  if (n_batch <= 0 || rows <= 0 || cols <= 0) return DM_OK;
  if (!ns_scratch) return polar_jacobi_launch(X, C, rows, cols, n_batch, scratch, nullptr, st);
  const bool wide = rows < cols;
  const int n = wide ? rows : cols;
  const size_t mat = size_t(rows) * cols;
  double* A0 = ns_scratch;
  double* A1 = A0 + size_t(n_batch) * mat;
  double* Z = A1 + size_t(n_batch) * mat;
  int* need = reinterpret_cast<int*>(Z + size_t(n_batch) * n * n);
  int rc;
  auto gram = [&](const double* A, const int* skip) {  // Z = A^T A (tall) or A A^T (wide), n x n
    GemmProblem G;
    G.A.d = A, G.A.ld = cols, G.A.batch_stride = int64_t(mat), G.A.rows = rows, G.A.trans = wide ? 0 : 1;
    G.B = G.A;
    G.M = n, G.N = n, G.K = wide ? cols : rows, G.maxM = n, G.maxN = n, G.maxK = G.K, G.n_batch = n_batch;
    G.C = Z, G.ldc = n, G.c_batch_stride = int64_t(n) * n, G.skip = skip;
    return gemm64_launch(G, st);
  };
  int* done = need + n_batch;
  DM_CUDA_OK(cudaMemsetAsync(done, 0, sizeof(int) * n_batch, st));
  if ((rc = gram(X, nullptr))) return rc;
  ns_scale_kernel<<<n_batch, 256, 0, st>>>(X, Z, A0, n, rows, cols);
  DM_LAUNCH_OK("ns_scale_kernel");
  double* cur = A0;
  double* nxt = A1;
  for (int it = 0; it < kNsIters; ++it) {
    ns_step_kernel<<<n_batch, 256, 0, st>>>(Z, cur, C, n, rows, cols, done);
    DM_LAUNCH_OK("ns_step_kernel");
    GemmProblem G;
    if (!wide) {  // A <- A W            (rows x n) (n x n), W symmetric
      G.A.d = cur, G.A.ld = cols, G.A.batch_stride = int64_t(mat), G.A.rows = rows, G.A.trans = 0;
      G.B.d = Z, G.B.ld = n, G.B.batch_stride = int64_t(n) * n, G.B.rows = n, G.B.trans = 0;
      G.M = rows, G.N = cols, G.K = n;
    } else {      // A <- W A            (n x n) (n x cols)
      G.A.d = Z, G.A.ld = n, G.A.batch_stride = int64_t(n) * n, G.A.rows = n, G.A.trans = 0;
      G.B.d = cur, G.B.ld = cols, G.B.batch_stride = int64_t(mat), G.B.rows = rows, G.B.trans = 1;
      G.M = rows, G.N = cols, G.K = n;
    }
    G.maxM = G.M, G.maxN = G.N, G.maxK = G.K, G.n_batch = n_batch;
    G.C = nxt, G.ldc = cols, G.c_batch_stride = int64_t(mat), G.skip = done;
    if ((rc = gemm64_launch(G, st))) return rc;
    double* tmp = cur;
    cur = nxt;
    nxt = tmp;
    if ((rc = gram(cur, done))) return rc;
  }
  ns_step_kernel<<<n_batch, 256, 0, st>>>(Z, cur, C, n, rows, cols, done);  // catches convergence in the last step
  DM_LAUNCH_OK("ns_step_kernel");
  ns_need_kernel<<<(n_batch + 255) / 256, 256, 0, st>>>(done, need, n_batch);
  DM_LAUNCH_OK("ns_need_kernel");
  return polar_jacobi_launch(X, C, rows, cols, n_batch, scratch, need, st);
}

}  // namespace dm
