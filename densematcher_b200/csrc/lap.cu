// Rectangular linear sum assignment in float64, one CTA per problem (mesh pair), batched over the grid.
//
// The reference calls scipy.optimize.linear_sum_assignment on the dense N2 x N1 map of every pair
// (densematcher/functional_map.py:57,66,78).  scipy's solver is the shortest-augmenting-path algorithm of
// Crouse ("On implementing 2D rectangular assignment algorithms", 2016): rows are added one at a time, each by a
// Dijkstra search over the columns with dual variables u, v.  The search is inherently sequential (the next row
// depends on the argmin of the current scan), so the parallelism here is
//   * inside a step: the scan over the remaining columns (512 threads, columns owned by fixed threads, their
//     v / shortest-path / position state in registers) and a two-level shuffle reduction with ONE barrier;
//   * across problems: one CTA per pair, 148 pairs in flight.
// Every floating-point expression is evaluated in scipy's order and ties are broken by scipy's rule (first minimum
// in the order of its `remaining` list, except that an unassigned column replaces an equal candidate), so the
// assignment is identical to scipy's, not merely of equal cost -- also on integer / constant matrices.
#include "dm_internal.cuh"

namespace dm {
namespace {

constexpr int kLapThreads = 512;
constexpr int kLapWarps = kLapThreads / 32;
constexpr int kLapMaxCpt = 16;  // columns per thread -> at most 8192 columns

struct LapArgs {
  const double* cost;
  const int64_t* cost_off;
  const int64_t* row_off;
  const int64_t* col_off;
  const int64_t* t_off;  // offset of the transposed copy of a tall problem inside tbuf
  double* tbuf;
  int maximize;
  void* out;
  int out_i64;
  int* status;
  int max_small, max_big;
};

__device__ __forceinline__ void store_idx(void* out, int i64, int64_t at, int v) {
  if (i64)
    static_cast<int64_t*>(out)[at] = v;
  else
    static_cast<int32_t*>(out)[at] = v;
}

// status[b] = 1 when the matrix holds a NaN or an entry that is -inf after the optional negation (scipy:
// "matrix contains invalid numeric entries"); also resets status[b] (first block of the problem, launched after
// lap_reset_kernel).
__global__ void lap_reset_kernel(LapArgs A, int n_batch, int64_t* t_off) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int64_t acc = 0;
    for (int b = 0; b < n_batch; ++b) {
      const int64_t nr = A.row_off[b + 1] - A.row_off[b], nc = A.col_off[b + 1] - A.col_off[b];
      A.status[b] = 0;
      t_off[b] = acc;
      if (nc < nr) acc += nr * nc;
    }
  }
}

__global__ void lap_validate_kernel(LapArgs A) {
  const int b = blockIdx.y;
  const int64_t nr = A.row_off[b + 1] - A.row_off[b], nc = A.col_off[b + 1] - A.col_off[b];
  const int64_t n = nr * nc;
  const double* c = A.cost + A.cost_off[b];
  int bad = 0;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += int64_t(gridDim.x) * blockDim.x) {
    double x = c[e];
    if (A.maximize) x = -x;
    if (x != x || x == -INFINITY) bad = 1;
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(A.status + b, 1);
}

// tall problems (more rows than columns) are solved on the transposed matrix, as scipy does
__global__ void lap_transpose_kernel(LapArgs A) {
  __shared__ double tile[32][33];
  const int b = blockIdx.z;
  const int nr = int(A.row_off[b + 1] - A.row_off[b]), nc = int(A.col_off[b + 1] - A.col_off[b]);
  if (nc >= nr) return;
  const double* src = A.cost + A.cost_off[b];
  double* dst = A.tbuf + A.t_off[b];
  for (int r0 = blockIdx.y * 32; r0 < nr; r0 += gridDim.y * 32)
    for (int c0 = blockIdx.x * 32; c0 < nc; c0 += gridDim.x * 32) {
      for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int r = r0 + y, c = c0 + threadIdx.x;
        if (r < nr && c < nc) tile[y][threadIdx.x] = src[int64_t(r) * nc + c];
      }
      __syncthreads();
      for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int c = c0 + y, r = r0 + threadIdx.x;
        if (r < nr && c < nc) dst[int64_t(c) * nr + r] = tile[threadIdx.x][y];
      }
      __syncthreads();
    }
}

constexpr unsigned kFull = 0xffffffffu;

// (value, key) lexicographic minimum over a warp; returns the winning lane.  The float64 value is mapped to an
// order-preserving 64-bit unsigned key and reduced as two 32-bit halves with redux.sync (three REDUX instead of five
// rounds of 64-bit shuffles + compares: this sits on the critical path of every Dijkstra step, twice).
__device__ __forceinline__ int warp_argmin(double val, unsigned key, double& vmin, unsigned& kmin) {
  const long long b = __double_as_longlong(__dadd_rn(val, 0.0));  // -0.0 -> +0.0: they compare equal as numbers
  const unsigned long long o = static_cast<unsigned long long>(b) ^ (static_cast<unsigned long long>(b >> 63) | 0x8000000000000000ull);
  const unsigned hi = unsigned(o >> 32), lo = unsigned(o);
  const unsigned mh = __reduce_min_sync(kFull, hi);
  const unsigned ml = __reduce_min_sync(kFull, hi == mh ? lo : 0xffffffffu);
  const bool is_min = hi == mh && lo == ml;
  kmin = __reduce_min_sync(kFull, is_min ? key : 0xffffffffu);
  const int src = __ffs(__ballot_sync(kFull, is_min && key == kmin)) - 1;
  vmin = __shfl_sync(kFull, val, src);
  return src;
}

// MINB = 2 caps the kernel at 64 registers so that two problems share an SM: 20 % slower for one problem, 20 % more
// throughput once there are more problems than SMs (124 ms instead of 156 ms for 296 problems of 2562 x 2562)
template <int CPT, int MINB>
__global__ void __launch_bounds__(kLapThreads, MINB) lap_kernel(LapArgs A) {
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t r0 = A.row_off[b];
  const int onr = int(A.row_off[b + 1] - r0), onc = int(A.col_off[b + 1] - A.col_off[b]);
  for (int i = tid; i < onr; i += kLapThreads) store_idx(A.out, A.out_i64, r0 + i, -1);
  if (onr == 0 || onc == 0 || A.status[b] != 0) return;
  const bool tall = onc < onr;
  const int nr = tall ? onc : onr, nc = tall ? onr : onc;
  const double* cost = tall ? A.tbuf + A.t_off[b] : A.cost + A.cost_off[b];
  const bool neg = A.maximize != 0;

  extern __shared__ __align__(16) uint8_t lap_smem[];
  double* u = reinterpret_cast<double*>(lap_smem);                 // [max_small]
  double* red_val = u + A.max_small;                               // [2][kLapWarps]
  unsigned* red_key = reinterpret_cast<unsigned*>(red_val + 2 * kLapWarps);  // [2][kLapWarps]
  int* red_j = reinterpret_cast<int*>(red_key + 2 * kLapWarps);
  int* red_row = red_j + 2 * kLapWarps;
  int* col4row = red_row + 2 * kLapWarps;                          // [max_small]
  int* path = col4row + A.max_small;                               // [max_big]
  int* row4col = path + A.max_big;                                 // [max_big]

  for (int i = tid; i < nr; i += kLapThreads) {
    u[i] = 0.0;
    col4row[i] = -1;
  }
  for (int j = tid; j < nc; j += kLapThreads) {
    path[j] = -1;
    row4col[j] = -1;
  }
  double v[CPT], spc[CPT];
#pragma unroll
  for (int m = 0; m < CPT; ++m) v[m] = 0.0;
  __syncthreads();

  int par = 0;
  constexpr unsigned kGone = 0xffffffffu;  // key of a column that is not in the `remaining` list (scanned, or j >= nc)
  for (int cur = 0; cur < nr; ++cur) {
    // key[m]: the column's rank in the tie-breaking order, kept up to date instead of being rebuilt in every step.
    // scipy scans `remaining` by ascending position and lets an unassigned column replace an equal candidate, so among
    // equal values the LAST unassigned position wins, else the FIRST assigned one:
    //   unassigned column at position q -> nc - 1 - q,   assigned column -> 0x80000000 | q      (minimum wins)
    // `remaining` starts as nc-1, nc-2, ..., 0, i.e. column j at position nc - 1 - j.
    unsigned key[CPT];
    unsigned scanned = 0;  // bit m: own column m was scanned in this search
#pragma unroll
    for (int m = 0; m < CPT; ++m) {
      const int j = tid + m * kLapThreads;
      spc[m] = INFINITY;
      key[m] = j < nc ? (row4col[j] < 0 ? unsigned(j) : (0x80000000u | unsigned(nc - 1 - j))) : kGone;
    }
    int nrem = nc, i = cur, sink = -1;
    double minVal = 0.0;
    while (true) {
      const double ui = u[i];
      const double* crow = cost + int64_t(i) * nc;
      double c[CPT];
#pragma unroll
      for (int m = 0; m < CPT; ++m) c[m] = key[m] != kGone ? __ldg(crow + tid + m * kLapThreads) : 0.0;
      double bval = INFINITY;
      unsigned bkey = kGone;
      int bm = 0;
#pragma unroll
      for (int m = 0; m < CPT; ++m)
        if (key[m] != kGone) {
          const double cc = neg ? -c[m] : c[m];
          const double r = __dsub_rn(__dsub_rn(__dadd_rn(minVal, cc), ui), v[m]);
          if (r < spc[m]) {
            spc[m] = r;
            path[tid + m * kLapThreads] = i;
          }
          if (spc[m] < bval || (spc[m] == bval && key[m] < bkey)) {
            bval = spc[m];
            bkey = key[m];
            bm = m;
          }
        }
      double wv;
      unsigned wk;
      int src = warp_argmin(bval, bkey, wv, wk);
      const int wj = __shfl_sync(kFull, tid + bm * kLapThreads, src);
      if (lane == 0) {
        red_val[par * kLapWarps + warp] = wv;
        red_key[par * kLapWarps + warp] = wk;
        red_j[par * kLapWarps + warp] = wj;
      }
      __syncthreads();
      const bool has = lane < kLapWarps;
      const double ev = has ? red_val[par * kLapWarps + lane] : INFINITY;
      const unsigned ek = has ? red_key[par * kLapWarps + lane] : kGone;
      const int ej = has ? red_j[par * kLapWarps + lane] : 0;
      double lowest;
      unsigned kmin;
      src = warp_argmin(ev, ek, lowest, kmin);
      const int jstar = __shfl_sync(kFull, ej, src);
      par ^= 1;
      if (lowest == INFINITY) {  // infeasible cost matrix
        if (tid == 0) A.status[b] = 2;
        return;
      }
      minVal = lowest;
      const int rowstar = row4col[jstar];
      const unsigned index = (kmin & 0x80000000u) ? (kmin & 0x7fffffffu) : unsigned(nc - 1) - kmin;
      // remaining[index] = remaining[--num_remaining]: the column at the last position moves to `index`; the chosen
      // column leaves the list
      const unsigned last_u = unsigned(nc - nrem), last_a = 0x80000000u | unsigned(nrem - 1);
      const unsigned moved_u = unsigned(nc - 1) - index, moved_a = 0x80000000u | index;
      const int mstar = (jstar - tid) / kLapThreads;  // own slot of the chosen column, if it is one of this thread's
      const bool mine = jstar - tid == mstar * kLapThreads;
#pragma unroll
      for (int m = 0; m < CPT; ++m) {
        unsigned k = key[m];
        k = k == last_u ? moved_u : k;
        k = k == last_a ? moved_a : k;
        key[m] = (mine && m == mstar) ? kGone : k;
      }
      if (mine) scanned |= 1u << mstar;
      --nrem;
      if (rowstar < 0) {
        sink = jstar;
        break;
      }
      i = rowstar;
    }
    // dual variables: rows reached in this search are exactly the rows of the scanned, assigned columns
#pragma unroll
    for (int m = 0; m < CPT; ++m)
      if (scanned & (1u << m)) {
        const double d = __dsub_rn(minVal, spc[m]);
        v[m] = __dsub_rn(v[m], d);
        const int rj = row4col[tid + m * kLapThreads];
        if (rj >= 0) u[rj] = __dadd_rn(u[rj], d);
      }
    if (tid == 0) u[cur] = __dadd_rn(u[cur], minVal);
    __syncthreads();
    if (tid == 0) {  // augment along the path
      int j = sink;
      while (true) {
        const int i2 = path[j];
        row4col[j] = i2;
        const int t = col4row[i2];
        col4row[i2] = j;
        j = t;
        if (i2 == cur) break;
      }
    }
    __syncthreads();
  }

  for (int i2 = tid; i2 < nr; i2 += kLapThreads) {
    if (tall)
      store_idx(A.out, A.out_i64, r0 + col4row[i2], i2);
    else
      store_idx(A.out, A.out_i64, r0 + i2, col4row[i2]);
  }
}

size_t lap_smem_bytes(int max_small, int max_big) {
  return size_t(max_small) * 12 + size_t(max_big) * 8 + 2 * kLapWarps * (8 + 4 + 4 + 4) + 64;
}

template <int CPT, int MINB>
int lap_launch_b(const LapArgs& A, int n_batch, size_t smem, cudaStream_t st) {
  static OncePerDevice once;
  if (once.first())
    DM_CUDA_OK(cudaFuncSetAttribute(lap_kernel<CPT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  lap_kernel<CPT, MINB><<<n_batch, kLapThreads, smem, st>>>(A);
  DM_LAUNCH_OK("lap_kernel");
  return DM_OK;
}

template <int CPT>
int lap_launch(const LapArgs& A, int n_batch, size_t smem, cudaStream_t st) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (CPT <= 8 && n_batch > sms && 2 * (smem + 1024) <= 227 * 1024) return lap_launch_b<CPT, (CPT <= 8 ? 2 : 1)>(A, n_batch, smem, st);
  return lap_launch_b<CPT, 1>(A, n_batch, smem, st);
}

}  // namespace
}  // namespace dm

using namespace dm;

extern "C" {

size_t dm_lap_workspace_bytes(int n_batch, int max_nr, int max_nc, int64_t tall_elems) {
  (void)max_nr;
  (void)max_nc;
  Carver c(nullptr);
  c.take<int64_t>(size_t(n_batch > 0 ? n_batch : 0) + 1);
  c.take<double>(size_t(tall_elems > 0 ? tall_elems : 0));
  return c.bytes();
}

int dm_lap_solve(const double* cost, const int64_t* cost_off, const int64_t* row_off, const int64_t* col_off, int n_batch,
                 int max_nr, int max_nc, int64_t tall_elems, int maximize, void* col_of_row, int* status, int flags,
                 void* workspace, size_t workspace_bytes, dm_stream_t stream) {
  if (n_batch < 0 || max_nr < 0 || max_nc < 0 || tall_elems < 0) DM_FAIL(DM_ERR_BADARG, "bad size");
  if (n_batch == 0) return DM_OK;
  if (!cost_off || !row_off || !col_off || !status) DM_FAIL(DM_ERR_BADARG, "null argument");
  if (max_nr > 0 && !col_of_row) DM_FAIL(DM_ERR_BADARG, "null argument");
  const int max_small = max_nr < max_nc ? max_nr : max_nc, max_big = max_nr < max_nc ? max_nc : max_nr;
  if (max_big > kLapThreads * kLapMaxCpt)
    DM_FAIL(DM_ERR_UNSUPPORTED, "assignment problems are limited to %d columns", kLapThreads * kLapMaxCpt);
  const size_t smem = lap_smem_bytes(max_small, max_big);
  if (smem > 227 * 1024) DM_FAIL(DM_ERR_UNSUPPORTED, "assignment problem does not fit in shared memory");
  const size_t need = dm_lap_workspace_bytes(n_batch, max_nr, max_nc, tall_elems);
  if (!workspace || need > workspace_bytes) DM_FAIL(DM_ERR_WORKSPACE, "workspace too small: need %zu", need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver c(workspace);
  int64_t* t_off = c.take<int64_t>(size_t(n_batch) + 1);
  double* tbuf = c.take<double>(size_t(tall_elems));
  LapArgs A;
  A.cost = cost;
  A.cost_off = cost_off;
  A.row_off = row_off;
  A.col_off = col_off;
  A.t_off = t_off;
  A.tbuf = tbuf;
  A.maximize = maximize;
  A.out = col_of_row;
  A.out_i64 = (flags & DM_I64_OUT) ? 1 : 0;
  A.status = status;
  A.max_small = max_small;
  A.max_big = max_big;
  lap_reset_kernel<<<1, 32, 0, st>>>(A, n_batch, t_off);
  DM_LAUNCH_OK("lap_reset_kernel");
  if (max_nr > 0 && max_nc > 0) {
    const int64_t max_elems = int64_t(max_nr) * max_nc;
    const int vblocks = int((max_elems + 256 * 16 - 1) / (256 * 16) < 592 ? (max_elems + 256 * 16 - 1) / (256 * 16) : 592);
    lap_validate_kernel<<<dim3(vblocks > 0 ? vblocks : 1, n_batch), 256, 0, st>>>(A);
    DM_LAUNCH_OK("lap_validate_kernel");
    if (tall_elems > 0) {
      const int gx = (max_nc + 31) / 32 < 64 ? (max_nc + 31) / 32 : 64, gy = (max_nr + 31) / 32 < 64 ? (max_nr + 31) / 32 : 64;
      lap_transpose_kernel<<<dim3(gx, gy, n_batch), dim3(32, 8), 0, st>>>(A);
      DM_LAUNCH_OK("lap_transpose_kernel");
    }
  }
  const int cpt = (max_big + kLapThreads - 1) / kLapThreads;  // columns per thread
  switch (cpt) {
    case 0: case 1: return lap_launch<1>(A, n_batch, smem, st);
    case 2: return lap_launch<2>(A, n_batch, smem, st);
    case 3: return lap_launch<3>(A, n_batch, smem, st);
    case 4: return lap_launch<4>(A, n_batch, smem, st);
    case 5: return lap_launch<5>(A, n_batch, smem, st);
    case 6: return lap_launch<6>(A, n_batch, smem, st);
    case 7: case 8: return lap_launch<8>(A, n_batch, smem, st);
    case 9: case 10: return lap_launch<10>(A, n_batch, smem, st);
    case 11: case 12: return lap_launch<12>(A, n_batch, smem, st);
    default: return lap_launch<16>(A, n_batch, smem, st);
  }
}

}  // extern "C"
