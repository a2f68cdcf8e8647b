// dm-nvcc-flags: -fmad=false
// Barycentric "precise map": every vertex of mesh 2 (a point of the p-dimensional spectral embedding) is projected
// onto the nearest triangle of mesh 1's embedding (Ezuz & Ben-Chen, "Deblurring and denoising of maps between
// shapes"; reference: densematcher/pyFM/spectral/projection_utils.py:16-115, called from convert.py:186-231 and
// functional_map.py:62).
//
//   l_max[f]      longest edge of face f                                   (projection_utils.py:118-146)
//   Delta_min[i]  distance from point i to the nearest vertex              (:149-186)
//   delta_min     distance to the nearest of a face's three vertices, through |x|^2 - 2 x.y + |y|^2   (:294-326)
//   candidates    faces with delta_min - l_max < Delta_min;  each is projected with Eberly's seven-region
//                 point-triangle routine and the first minimum of the distances wins                 (:329-377)
//
// One float64 GEMM gives all vertex-point inner products; then ONE CTA PER POINT keeps its row of squared vertex
// distances in shared memory, scans the faces (a thread per face), and every candidate found is projected by its
// warp (lanes over the embedding dimension, shuffle-reduced inner products, warp-uniform case analysis).
// This file is compiled without FMA contraction: the case analysis compares sums of products against each other
// exactly as the reference's numpy expressions do.
#include "dm_internal.cuh"
#include "gemm64.cuh"

namespace dm {
namespace {

constexpr int kPmThreads = 256;
constexpr int kPmWarps = kPmThreads / 32;
constexpr unsigned kFullMask = 0xffffffffu;

struct Proj {
  double sq, s, t;
};

// Closest point of the triangle B + s E0 + t E1 from the inner products a = E0.E0, b = E0.E1, c = E1.E1,
// d = E0.(B-P), e = E1.(B-P), f = |B-P|^2.  Branch for branch the vectorised reference routine
// (projection_utils.py:483-757), INCLUDING the two region-4 branches where it forms the squared distance with the
// un-normalised s / t of the region test (:543-544, :563-564); a point with a single candidate triangle goes through
// the scalar routine (:820-976) instead, whose barycentric coordinates are the same and whose distance is not used.
__device__ Proj tri_closest_point(double a, double b, double c, double d, double e, double f) {
  const double det = a * c - b * b;
  double s = b * e - c * d;
  double t = b * d - a * e;
  Proj r;
  auto full = [&](double s_, double t_) { return s_ * (a * s_ + b * t_ + 2.0 * d) + t_ * (b * s_ + c * t_ + 2.0 * e) + f; };
  if (s + t <= det) {
    if (s < 0) {
      if (t < 0) {  // region 4
        if (d < 0) {
          if (-d >= a) return Proj{a + 2.0 * d + f, 1.0, 0.0};
          return Proj{d * s + f, -d / a, 0.0};
        }
        if (e >= 0) return Proj{f, 0.0, 0.0};
        if (-e >= c) return Proj{c + 2.0 * e + f, 0.0, 1.0};
        return Proj{e * t + f, 0.0, -e / c};
      }
      // region 3
      if (e >= 0) return Proj{f, 0.0, 0.0};
      if (-e >= c) return Proj{c + 2.0 * e + f, 0.0, 1.0};
      r.t = -e / c;
      return Proj{e * r.t + f, 0.0, r.t};
    }
    if (t < 0) {  // region 5
      if (d >= 0) return Proj{f, 0.0, 0.0};
      if (-d >= a) return Proj{a + 2.0 * d + f, 1.0, 0.0};
      r.s = -d / a;
      return Proj{d * r.s + f, r.s, 0.0};
    }
    const double inv = 1.0 / det;  // region 0
    s = s * inv;
    t = t * inv;
    return Proj{full(s, t), s, t};
  }
  if (s < 0) {  // region 2
    const double tmp0 = b + d, tmp1 = c + e;
    if (tmp1 > tmp0) {
      const double numer = tmp1 - tmp0, denom = a - 2.0 * b + c;
      if (numer >= denom) return Proj{a + 2.0 * d + f, 1.0, 0.0};
      s = numer / denom;
      t = 1 - s;
      return Proj{full(s, t), s, t};
    }
    if (tmp1 <= 0) return Proj{c + 2.0 * e + f, 0.0, 1.0};
    if (e >= 0) return Proj{f, 0.0, 0.0};
    r.t = -e / c;
    return Proj{e * r.t + f, 0.0, r.t};
  }
  if (t < 0) {  // region 6
    const double tmp0 = b + e, tmp1 = a + d;
    if (tmp1 > tmp0) {
      const double numer = tmp1 - tmp0, denom = a - 2.0 * b + c;
      if (numer >= denom) return Proj{c + 2.0 * e + f, 0.0, 1.0};
      t = numer / denom;
      s = 1 - t;
      return Proj{full(s, t), s, t};
    }
    if (tmp1 <= 0) return Proj{a + 2.0 * d + f, 1.0, 0.0};
    if (d >= 0) return Proj{f, 0.0, 0.0};
    r.s = -d / a;
    return Proj{d * r.s + f, r.s, 0.0};
  }
  const double numer = c + e - b - d;  // region 1
  if (numer <= 0) return Proj{c + 2.0 * e + f, 0.0, 1.0};
  const double denom = a - 2.0 * b + c;
  if (numer >= denom) return Proj{a + 2.0 * d + f, 1.0, 0.0};
  s = numer / denom;
  t = 1 - s;
  return Proj{full(s, t), s, t};
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

// sq[r] = (||x_r||_2)^2 the way the reference forms it: np.linalg.norm(...)**2 (:317-320)
__global__ void pm_sqnorm_kernel(const double* X, int64_t ld, int64_t n, int p, double* sq) {
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  double acc = 0.0;
  for (int k = lane; k < p; k += 32) {
    const double x = X[r * ld + k];
    acc += x * x;
  }
  acc = warp_sum(acc);
  const double nrm = sqrt(acc);
  if (lane == 0) sq[r] = nrm * nrm;
}

// longest edge of every face in the embedding
__global__ void pm_lmax_kernel(const double* X, int64_t ld, const int64_t* off1, const int32_t* faces, const int64_t* foff,
                               int n_pairs, int p, double* lmax) {
  const int64_t fg = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (fg >= foff[n_pairs]) return;
  int lo = 0, hi = n_pairs - 1;  // pair of this face
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (foff[mid] <= fg) lo = mid; else hi = mid - 1;
  }
  const double* base = X + off1[lo] * ld;
  const double* x0 = base + int64_t(faces[3 * fg + 0]) * ld;
  const double* x1 = base + int64_t(faces[3 * fg + 1]) * ld;
  const double* x2 = base + int64_t(faces[3 * fg + 2]) * ld;
  double s01 = 0.0, s12 = 0.0, s20 = 0.0;
  for (int k = lane; k < p; k += 32) {
    const double a = x0[k], b = x1[k], c = x2[k];
    s01 += (b - a) * (b - a);
    s12 += (c - b) * (c - b);
    s20 += (a - c) * (a - c);
  }
  s01 = warp_sum(s01), s12 = warp_sum(s12), s20 = warp_sum(s20);
  if (lane == 0) lmax[fg] = fmax(fmax(sqrt(s01), sqrt(s12)), sqrt(s20));
}

struct PmArgs {
  const double* emb1;
  int64_t ld1;
  const int64_t* off1;
  const int32_t* faces;
  const int64_t* foff;
  const double* emb2;
  int64_t ld2;
  const int64_t* off2;
  int n_pairs, p;
  const double* D;  // [total_n2, ldD] inner products point x vertex
  int64_t ldD;
  const double *sq1, *sq2, *lmax;
  void* face_match;
  int out_i64;
  double* bary;
};

struct Best {
  double dist;
  int face;
  double s, t;
};
// numpy's argmin: the first minimum, and a NaN (degenerate triangle) counts as smaller than any number
__device__ __forceinline__ bool better(double dist, int face, const Best& b) {
  const bool na = dist != dist, nb = b.dist != b.dist;
  if (na != nb) return na;
  if (na) return face < b.face;
  return dist < b.dist || (dist == b.dist && face < b.face);
}

__global__ void __launch_bounds__(kPmThreads) precise_map_kernel(PmArgs A) {
  extern __shared__ __align__(16) uint8_t pm_smem[];
  double* d2 = reinterpret_cast<double*>(pm_smem);  // [n1] clamped squared distance to every vertex
  __shared__ double red_v[kPmWarps];
  __shared__ int red_i[kPmWarps];
  __shared__ Best red_b[kPmWarps];
  __shared__ double s_Deltamin;

  const int64_t pt = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int lo = 0, hi = A.n_pairs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (A.off2[mid] <= pt) lo = mid; else hi = mid - 1;
  }
  const int pair = lo;
  const int64_t v0 = A.off1[pair];
  const int n1 = int(A.off1[pair + 1] - v0);
  const int64_t f0 = A.foff[pair];
  const int nf = int(A.foff[pair + 1] - f0);
  const double* X = A.emb1 + v0 * A.ld1;
  const double* y = A.emb2 + pt * A.ld2;
  const double* Drow = A.D + pt * A.ldD;
  const double sqy = A.sq2[pt];

  // squared distances (:225-236: X @ Y.T, *= -2, += |x|^2, += |y|^2, clamp at 0) and the nearest vertex
  double bv = INFINITY;
  int bi = 0x7fffffff;
  for (int v = tid; v < n1; v += kPmThreads) {
    double x = Drow[v] * -2.0;
    x += A.sq1[v0 + v];
    x += sqy;
    x = fmax(x, 0.0);
    d2[v] = x;
    if (x < bv) bv = x, bi = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(kFullMask, bv, o);
    const int oi = __shfl_xor_sync(kFullMask, bi, o);
    if (ov < bv || (ov == bv && oi < bi)) bv = ov, bi = oi;
  }
  if (lane == 0) red_v[warp] = bv, red_i[warp] = bi;
  __syncthreads();
  if (warp == 0) {
    bv = lane < kPmWarps ? red_v[lane] : INFINITY;
    bi = lane < kPmWarps ? red_i[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(kFullMask, bv, o);
      const int oi = __shfl_xor_sync(kFullMask, bi, o);
      if (ov < bv || (ov == bv && oi < bi)) bv = ov, bi = oi;
    }
    // Delta_min: the distance itself, from coordinate differences (what the kd-tree query returns, :181-184)
    double acc = 0.0;
    if (n1 > 0) {
      const double* xn = X + int64_t(bi) * A.ld1;
      for (int k = lane; k < A.p; k += 32) {
        const double df = y[k] - xn[k];
        acc += df * df;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) s_Deltamin = sqrt(acc);
  }
  __syncthreads();
  const double Deltamin = s_Deltamin;

  Best best{INFINITY, 0x7fffffff, 0.0, 0.0};
  for (int fb = warp * 32; fb < nf; fb += kPmThreads) {
    const int fi = fb + lane;
    bool cand = false;
    int i0 = 0, i1 = 0, i2 = 0;
    if (fi < nf) {
      i0 = A.faces[3 * (f0 + fi) + 0], i1 = A.faces[3 * (f0 + fi) + 1], i2 = A.faces[3 * (f0 + fi) + 2];
      const double dmin = sqrt(fmin(fmin(d2[i0], d2[i1]), d2[i2]));
      cand = dmin - A.lmax[f0 + fi] < Deltamin;
    }
    unsigned m = __ballot_sync(kFullMask, cand);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const int j0 = __shfl_sync(kFullMask, i0, src), j1 = __shfl_sync(kFullMask, i1, src), j2 = __shfl_sync(kFullMask, i2, src);
      const double* x0 = X + int64_t(j0) * A.ld1;
      const double* x1 = X + int64_t(j1) * A.ld1;
      const double* x2 = X + int64_t(j2) * A.ld1;
      double a = 0, b = 0, c = 0, d = 0, e = 0, f = 0;
      for (int k = lane; k < A.p; k += 32) {
        const double B = x0[k], E0 = x1[k] - B, E1 = x2[k] - B, Dv = B - y[k];
        a += E0 * E0, b += E0 * E1, c += E1 * E1, d += E0 * Dv, e += E1 * Dv, f += Dv * Dv;
      }
      a = warp_sum(a), b = warp_sum(b), c = warp_sum(c), d = warp_sum(d), e = warp_sum(e), f = warp_sum(f);
      const Proj pr = tri_closest_point(a, b, c, d, e, f);
      const double dist = sqrt(fmax(pr.sq, 0.0));  // :754-755
      const int face = fb + src;
      if (better(dist, face, best)) best = Best{dist, face, pr.s, pr.t};
    }
  }
  // all lanes of a warp hold the same `best`; combine the warps.  (With a single candidate in total its distance
  // plays no role, so the scalar routine the reference uses for that case need not be distinguished.)
  if (lane == 0) red_b[warp] = best;
  __syncthreads();
  if (tid == 0) {
    Best bb = red_b[0];
    for (int w = 1; w < kPmWarps; ++w)
      if (better(red_b[w].dist, red_b[w].face, bb)) bb = red_b[w];
    const int face = bb.face == 0x7fffffff ? 0 : bb.face;  // no candidate at all: vertex 0 of face 0
    if (A.out_i64)
      static_cast<int64_t*>(A.face_match)[pt] = face;
    else
      static_cast<int32_t*>(A.face_match)[pt] = face;
    A.bary[3 * pt + 0] = 1 - bb.s - bb.t;  // :759
    A.bary[3 * pt + 1] = bb.s;
    A.bary[3 * pt + 2] = bb.t;
  }
}

}  // namespace
}  // namespace dm

using namespace dm;

extern "C" {

size_t dm_precise_map_workspace_bytes(int n_pairs, int64_t total_n1, int max_n1, int64_t total_n2, int64_t total_faces) {
  (void)n_pairs;
  Carver c(nullptr);
  c.take<double>(size_t(total_n2 > 0 ? total_n2 : 0) * size_t(max_n1 > 0 ? max_n1 : 0));
  c.take<double>(size_t(total_n1 > 0 ? total_n1 : 0));
  c.take<double>(size_t(total_n2 > 0 ? total_n2 : 0));
  c.take<double>(size_t(total_faces > 0 ? total_faces : 0));
  return c.bytes();
}

int dm_precise_map(const double* emb1, int64_t ld1, const int64_t* off1, int64_t total_n1, int max_n1, const int32_t* faces,
                   const int64_t* face_off, int64_t total_faces, const double* emb2, int64_t ld2, const int64_t* off2,
                   int64_t total_n2, int max_n2, int n_pairs, int p, void* face_match, double* bary, int flags,
                   void* workspace, size_t workspace_bytes, dm_stream_t stream) {
  if (n_pairs < 0 || p <= 0 || total_n1 < 0 || total_n2 < 0 || total_faces < 0 || max_n1 < 0 || max_n2 < 0)
    DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n_pairs == 0 || total_n2 == 0) return DM_OK;
  if (!emb1 || !emb2 || !off1 || !off2 || !faces || !face_off || !face_match || !bary) DM_FAIL(DM_ERR_BADARG, "null argument");
  if (ld1 < p || ld2 < p) DM_FAIL(DM_ERR_BADARG, "leading dimension smaller than the embedding dimension");
  if (total_n1 == 0 || total_faces == 0) DM_FAIL(DM_ERR_BADARG, "mesh 1 has no vertices or no faces");
  const size_t smem = size_t(max_n1) * sizeof(double);
  if (smem > 200 * 1024) DM_FAIL(DM_ERR_UNSUPPORTED, "precise map: at most %d vertices per mesh", 200 * 1024 / 8);
  if (total_n2 > 0x7fffffffLL) DM_FAIL(DM_ERR_UNSUPPORTED, "too many points for one call");
  const size_t need = dm_precise_map_workspace_bytes(n_pairs, total_n1, max_n1, total_n2, total_faces);
  if (!workspace || need > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu", need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver c(workspace);
  double* D = c.take<double>(size_t(total_n2) * max_n1);
  double* sq1 = c.take<double>(size_t(total_n1));
  double* sq2 = c.take<double>(size_t(total_n2));
  double* lmax = c.take<double>(size_t(total_faces));
  int rc;
  GemmProblem G;
  G.A.d = emb2, G.A.ld = ld2, G.A.off = off2, G.A.trans = 0;
  G.B.d = emb1, G.B.ld = ld1, G.B.off = off1, G.B.trans = 0;
  G.K = p, G.maxM = max_n2, G.maxN = max_n1, G.maxK = p, G.n_batch = n_pairs;
  G.C = D, G.ldc = max_n1, G.c_off = off2;
  if ((rc = gemm64_launch(G, st))) return rc;
  pm_sqnorm_kernel<<<unsigned((total_n1 * 32 + 255) / 256), 256, 0, st>>>(emb1, ld1, total_n1, p, sq1);
  pm_sqnorm_kernel<<<unsigned((total_n2 * 32 + 255) / 256), 256, 0, st>>>(emb2, ld2, total_n2, p, sq2);
  pm_lmax_kernel<<<unsigned((total_faces * 32 + 255) / 256), 256, 0, st>>>(emb1, ld1, off1, faces, face_off, n_pairs, p, lmax);
  DM_LAUNCH_OK("precise map preparation");
  PmArgs A;
  A.emb1 = emb1, A.ld1 = ld1, A.off1 = off1, A.faces = faces, A.foff = face_off;
  A.emb2 = emb2, A.ld2 = ld2, A.off2 = off2, A.n_pairs = n_pairs, A.p = p;
  A.D = D, A.ldD = max_n1, A.sq1 = sq1, A.sq2 = sq2, A.lmax = lmax;
  A.face_match = face_match, A.out_i64 = (flags & DM_I64_OUT) ? 1 : 0, A.bary = bary;
  static OncePerDevice once;
  if (once.first())
    DM_CUDA_OK(cudaFuncSetAttribute(precise_map_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  precise_map_kernel<<<unsigned(total_n2), kPmThreads, smem, st>>>(A);
  DM_LAUNCH_OK("precise_map_kernel");
  return DM_OK;
}

}  // extern "C"
