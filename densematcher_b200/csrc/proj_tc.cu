// tcgen05 engine of the spectral projection  out[b] = A_b^T B_b   (A_b: n_b x k scaled eigenvectors, B_b: n_b x d
// features; the contraction runs over the VERTICES of mesh b).
//
// Replaces the dense contraction Phi^T A F of densematcher/pyFM/optimize/base_functions.py:526-532 (and
// TriMesh.project, mesh/trimesh.py:533-556).  This is the one genuinely GEMM-shaped, HBM-bound stage of the path
// (arithmetic intensity ~40 flop/B), so it runs on the tensor cores:
//   * both operands stay in their natural row-major [vertex, column] layout, i.e. "MN-major" for the UMMA (the
//     contraction index is the slow one); TMA fetches [64 columns x 16 vertices] boxes with the 128-byte swizzle and
//     the shared-memory descriptors describe MN-major SWIZZLE_128B atoms (leading-dimension offset = distance between
//     64-column groups, stride offset = distance between 8-vertex groups);
//   * fp32-grade products from a three-way bf16 split v = h + m + l (8 + 8 + 8 mantissa bits) and the six products
//     hh, hm, mh, mm, hl, lh, accumulated in fp32 in tensor memory: |error| ~ 2^-23 per product;
//   * a CTA owns a 128 x (up to 512) output tile of one mesh for a slice of its vertices (split-K); the fp32 partial
//     tiles are summed in float64 by proj_reduce_kernel (deterministic).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#include "dm_internal.cuh"
#include "tc_ptx.cuh"

namespace dm {
namespace {
using namespace tc;

constexpr int PM = 128;   // UMMA M: eigen-index rows of one output tile
constexpr int PKC = 16;   // vertices per pipeline stage (= UMMA K for bf16)
constexpr int kProjThreads = 192;
constexpr int kSliceChunks = 16;  // vertex chunks (of PKC) per CTA: 256 vertices (short fp32 accumulation chains: the TMEM accumulate truncates)
constexpr uint32_t GROUP_BYTES = PKC * 128;  // one 64-column group of one stage: 16 vertices x 128 B
constexpr uint32_t A_BYTES = 2 * GROUP_BYTES;  // 128 columns

struct ProjMaps {
  CUtensorMap a[3], b[3];
};

struct ProjParams {
  const int64_t* b_off;  // packed row offsets of the B operand (and the true row counts)
  int64_t a_stride_rows; // rows reserved per batch in the padded A operand (multiple of 16)
  int n_batch, m_tiles, ksplit, nb_groups, stages;
  float* partial;        // [ksplit][n_batch][m_tiles * 128][nb_groups * 64]
};

// MN-major operand, SWIZZLE_128B: 64-element (128 B) rows, 8-row atoms of 1024 B; LBO = bytes between 64-element
// groups along M/N, SBO = bytes between 8-row groups along K.  Descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((saddr >> 4) & 0x3FFF);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}
// kind::f16, A = B = bf16, both MN-major (bits 15 / 16), D = fp32, M = 128
__device__ __forceinline__ uint32_t idesc_mn(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (uint32_t(n >> 3) << 17) | (uint32_t(PM >> 4) << 24);
}

__global__ void __launch_bounds__(kProjThreads, 1) proj_tc_kernel(const __grid_constant__ ProjMaps maps, const ProjParams P) {
  int bid = blockIdx.x;
  const int ks = bid % P.ksplit;
  bid /= P.ksplit;
  const int mt = bid % P.m_tiles;
  const int b = bid / P.m_tiles;
  const int64_t r0 = P.b_off[b];
  const int nrows = int(P.b_off[b + 1] - r0);
  const int n_chunks = (nrows + PKC - 1) / PKC;
  // fixed slices of kSliceChunks x 16 vertices: the summation order of a mesh does not depend on the batch it is in
  const int c_beg = min(n_chunks, ks * kSliceChunks), c_end = min(n_chunks, c_beg + kSliceChunks);
  const int n_my = c_end - c_beg;
  const int NN = P.nb_groups * 64;
  float* out = P.partial + ((int64_t(ks) * P.n_batch + b) * P.m_tiles + mt) * PM * int64_t(NN);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (n_my <= 0) {  // nothing to contract in this slice: the partial tile is zero
    for (int e = threadIdx.x; e < PM * NN / 4; e += kProjThreads) reinterpret_cast<float4*>(out)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }

  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t b_bytes = P.nb_groups * GROUP_BYTES;
  const uint32_t stage_bytes = 3 * A_BYTES + 3 * b_bytes;
  const uint32_t off_bar = P.stages * stage_bytes;
  const uint32_t bar_full = sbase + off_bar;            // [stages]
  const uint32_t bar_empty = bar_full + 8 * P.stages;   // [stages]
  const uint32_t bar_done = bar_empty + 8 * P.stages;   // accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + off_bar + 8 * (2 * P.stages + 1));

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < 3; ++i) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[i]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b[i]) : "memory");
      }
      int stage = 0;
      uint32_t phase = 0;
      const int64_t a_row0 = int64_t(b) * P.a_stride_rows;
      for (int c = c_beg; c < c_end; ++c) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        const uint32_t sb = sbase + stage * stage_bytes, fb = bar_full + 8 * stage;
        mbar_expect_tx(fb, stage_bytes);
        const int arow = int(a_row0 + int64_t(c) * PKC), brow = int(r0 + int64_t(c) * PKC);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          tma_load_3d(sb + i * A_BYTES, &maps.a[i], 0, arow, 2 * mt, fb);
          tma_load_3d(sb + 3 * A_BYTES + i * b_bytes, &maps.b[i], 0, brow, 0, fb);
        }
        if (++stage == P.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = c_beg; c < c_end; ++c) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        const uint32_t sa = sbase + stage * stage_bytes, sb = sa + 3 * A_BYTES;
        const uint64_t ah = umma_desc_mn_sw128(sa, GROUP_BYTES), am = umma_desc_mn_sw128(sa + A_BYTES, GROUP_BYTES),
                       al = umma_desc_mn_sw128(sa + 2 * A_BYTES, GROUP_BYTES);
        for (int g0 = 0; g0 < P.nb_groups; g0 += 4) {  // N chunks of at most 256 columns
          const int gn = min(4, P.nb_groups - g0);
          const uint32_t idesc = idesc_mn(gn * 64);
          const uint32_t tacc = tmem_base + g0 * 64;
          const uint64_t bh = umma_desc_mn_sw128(sb + g0 * GROUP_BYTES, GROUP_BYTES),
                         bm = umma_desc_mn_sw128(sb + b_bytes + g0 * GROUP_BYTES, GROUP_BYTES),
                         bl = umma_desc_mn_sw128(sb + 2 * b_bytes + g0 * GROUP_BYTES, GROUP_BYTES);
          // smallest terms first
          tc_mma_bf16(tacc, al, bh, idesc, c != c_beg);
          tc_mma_bf16(tacc, ah, bl, idesc, 1);
          tc_mma_bf16(tacc, am, bm, idesc, 1);
          tc_mma_bf16(tacc, am, bh, idesc, 1);
          tc_mma_bf16(tacc, ah, bm, idesc, 1);
          tc_mma_bf16(tacc, ah, bh, idesc, 1);
        }
        tc_commit(bar_empty + 8 * stage);
        if (++stage == P.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      tc_commit(bar_done);
    }
  } else {
    const int q = warp & 3;  // TMEM lanes 32 q .. 32 q + 31
    mbar_wait(bar_done, 0);
    tc_fence_after();
    float* orow = out + int64_t(32 * q + lane) * NN;
    const uint32_t taddr = tmem_base + (uint32_t(32 * q) << 16);
    for (int ch = 0; ch < NN / 32; ++ch) {
      float v[32];
      tmem_ld32(taddr + ch * 32, v);
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4)
        reinterpret_cast<float4*>(orow + ch * 32)[c4] = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// out[b][r][c] (float64) = sum_s partial[s][b][r][c],  r < k, c < d   (slices added in ascending order, in float64)
// VEC: four consecutive columns per thread (16-byte loads, all slices of a group in flight); needs d % 4 == 0
template <bool VEC>
__global__ void __launch_bounds__(256)
    proj_reduce_kernel(const float* __restrict__ partial, int ksplit, int n_batch, int rows_pad, int NN, int k, int d,
                       double* __restrict__ out) {
  constexpr int W = VEC ? 4 : 1;
  const int dq = d / W;
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = int64_t(n_batch) * k * dq;
  if (idx >= total) return;
  const int c = int(idx % dq) * W;
  const int r = int((idx / dq) % k);
  const int b = int(idx / (int64_t(dq) * k));
  const int64_t stride = int64_t(n_batch) * rows_pad * NN;
  const float* p = partial + (int64_t(b) * rows_pad + r) * NN + c;
  double* o = out + (int64_t(b) * k + r) * d + c;
  if (VEC) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int q = 0;
    for (; q + 4 <= ksplit; q += 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(p + (q + u) * stride));
#pragma unroll
      for (int u = 0; u < 4; ++u) s0 += double(v[u].x), s1 += double(v[u].y), s2 += double(v[u].z), s3 += double(v[u].w);
    }
    for (; q < ksplit; ++q) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p + q * stride));
      s0 += double(v.x), s1 += double(v.y), s2 += double(v.z), s3 += double(v.w);
    }
    *reinterpret_cast<double2*>(o) = make_double2(s0, s1);
    *reinterpret_cast<double2*>(o + 2) = make_double2(s2, s3);
  } else {
    double s = 0.0;
    for (int q = 0; q < ksplit; ++q) s += double(p[q * stride]);
    *o = s;
  }
}

// Three-way bf16 split of a (scaled, optionally gathered) matrix:  v = scale[row] * src[g(row)][col] = h + m + l.
// One warp per OUTPUT row.  Packed output (out_stride_rows == 0): output rows follow `off`.  Padded output: batch b
// owns rows b * out_stride_rows .. and the rows past its true count are zero.
template <typename T>
__global__ void __launch_bounds__(256)
    split3_kernel(const T* __restrict__ src, int64_t ld, const int64_t* __restrict__ off, int n_batch, int cols,
                  const double* __restrict__ rowscale, const void* __restrict__ gather, int gather_i64,
                  const int64_t* __restrict__ gather_src_off, __nv_bfloat16* __restrict__ h, __nv_bfloat16* __restrict__ m,
                  __nv_bfloat16* __restrict__ l, int kp, int64_t out_stride_rows, int64_t out_rows) {
  const int lane = threadIdx.x & 31;
  const int64_t orow = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (orow >= out_rows) return;
  int b;
  int64_t prow;  // packed row (index into rowscale / gather / src when not gathered)
  bool valid = true;
  if (out_stride_rows > 0) {
    b = int(orow / out_stride_rows);
    const int64_t r = orow % out_stride_rows;
    prow = off[b] + r;
    valid = prow < off[b + 1];
  } else {
    prow = orow;
    b = 0;
    if (gather) {  // batch of a packed row: binary search in the offsets
      int lo = 0, hi = n_batch;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= prow) lo = mid; else hi = mid;
      }
      b = lo;
    }
  }
  __nv_bfloat16* ph = h + orow * kp;
  __nv_bfloat16* pm = m + orow * kp;
  __nv_bfloat16* pl = l + orow * kp;
  const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
  if (!valid) {
    for (int c = lane; c < kp; c += 32) ph[c] = z, pm[c] = z, pl[c] = z;
    return;
  }
  int64_t srow = prow;
  if (gather) srow = (gather_src_off ? gather_src_off[b] : 0) + load_index(gather, prow, gather_i64 != 0);
  const double sc = rowscale ? rowscale[prow] : 1.0;
  const T* s = src + srow * ld;
  typedef RowVec<T> RV;
  constexpr int W = RV::W, U = 4;
  if ((ld % W) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    // 16-byte loads, U in flight per lane; packed bf16 stores
    for (int k0 = lane * W; k0 < kp; k0 += 32 * W * U) {
      typename RV::V buf[U];
      bool full[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = k0 + u * 32 * W;
        full[u] = k + W <= cols;
        if (full[u]) buf[u] = __ldg(reinterpret_cast<const typename RV::V*>(s + k));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = k0 + u * 32 * W;
        if (k >= kp) continue;
        T v[W];
#pragma unroll
        for (int w = 0; w < W; ++w) {
          v[w] = T(0);
          if (full[u])
            v[w] = reinterpret_cast<const T*>(&buf[u])[w];
          else if (k + w < cols)
            v[w] = s[k + w];
        }
        unsigned short bh[W], bm[W], bl[W];
#pragma unroll
        for (int w = 0; w < W; ++w) {
          // one rounding to fp32 (2^-24, the precision of the whole engine), then an exact three-way split in fp32:
          // conversions between fp32 and fp64 are slow-pipe instructions and are kept to one per element
          const float x = (rowscale || sizeof(T) == 8) ? float(sc * v[w]) : float(v[w]);
          const __nv_bfloat16 vh = __float2bfloat16_rn(x);
          const float r1 = x - __bfloat162float(vh);
          const __nv_bfloat16 vm = __float2bfloat16_rn(r1);
          const float r2 = r1 - __bfloat162float(vm);
          bh[w] = __bfloat16_as_ushort(vh), bm[w] = __bfloat16_as_ushort(vm);
          bl[w] = __bfloat16_as_ushort(__float2bfloat16_rn(r2));
        }
        store_bf16_vec<W>(ph + k, bh);
        store_bf16_vec<W>(pm + k, bm);
        store_bf16_vec<W>(pl + k, bl);
      }
    }
    return;
  }
  for (int c = lane; c < kp; c += 32) {
    const double v = c < cols ? sc * double(s[c]) : 0.0;
    const __nv_bfloat16 vh = __float2bfloat16_rn(float(v));
    const double r1 = v - double(__bfloat162float(vh));
    const __nv_bfloat16 vm = __float2bfloat16_rn(float(r1));
    const double r2 = r1 - double(__bfloat162float(vm));
    ph[c] = vh, pm[c] = vm, pl[c] = __float2bfloat16_rn(float(r2));
  }
}

int make_map3(CUtensorMap* m, const void* base, int64_t rows, int kp, int box_groups) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) DM_FAIL(DM_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  // dims: 64 columns of a group | rows (vertices) | column groups
  const cuuint64_t gdim[3] = {64, cuuint64_t(rows > 0 ? rows : 1), cuuint64_t(kp / 64)};
  const cuuint64_t gstr[2] = {cuuint64_t(kp) * 2, 128};
  const cuuint32_t box[3] = {64, cuuint32_t(PKC), cuuint32_t(box_groups)};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) DM_FAIL(DM_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", int(r));
  return DM_OK;
}

int pad_to(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

// ---------------------------------------------------------------- host interface (used by dm_project / ZoomOut)
bool proj_tc_supported(int k, int d) { return k >= 1 && d >= 1 && pad_to(d, 64) <= 512; }

int proj_tc_ksplit(int /*n_batch*/, int /*m_tiles*/, int max_n) {
  const int chunks = (max_n + PKC - 1) / PKC;
  const int ks = (chunks + kSliceChunks - 1) / kSliceChunks;
  return ks < 1 ? 1 : ks;
}

size_t proj_tc_workspace_bytes(int n_batch, int64_t total_n, int max_n, int k, int d, bool b_presplit) {
  const int kpA = pad_to(k, 128), kpB = pad_to(d, 64);
  const int m_tiles = kpA / 128;
  const int64_t a_rows = int64_t(n_batch) * pad_to(max_n, PKC);
  Carver c(nullptr);
  for (int i = 0; i < 3; ++i) c.take<uint16_t>(size_t(a_rows) * kpA);
  for (int i = 0; i < 3; ++i) c.take<uint16_t>(b_presplit ? 0 : size_t(total_n) * kpB);
  c.take<float>(size_t(proj_tc_ksplit(n_batch, m_tiles, max_n)) * n_batch * kpA * kpB);
  return c.bytes();
}

// out[b] (k x d, float64) = sum over the rows r of batch b of  (a_scale[r] A[ga(r)][:k])^T (b_scale[r] B[gb(r)][:d])
// A is float64, B float32 or float64 (exactly one of Bf / Bd non-null); gathers are optional local row indices.
int proj_tc_run(const double* A, int64_t ldA, const double* a_scale, const float* Bf, const double* Bd, int64_t ldB,
                const double* b_scale, const void* b_gather, int b_gather_i64, const int64_t* b_gather_src_off,
                const int64_t* off, int64_t total_n, int max_n, int n_batch, int k, int d, double* out, void* ws,
                size_t ws_bytes, cudaStream_t st, const void* const* b_presplit) {
  if (n_batch <= 0) return DM_OK;
  if (total_n > 0x7fffffffLL) DM_FAIL(DM_ERR_BADARG, "too many rows for TMA coordinates");
  const int kpA = pad_to(k, 128), kpB = pad_to(d, 64);
  const int m_tiles = kpA / 128, nb_groups = kpB / 64;
  const int a_stride = pad_to(max_n, PKC);
  const int64_t a_rows = int64_t(n_batch) * a_stride;
  if (a_rows > 0x7fffffffLL) DM_FAIL(DM_ERR_BADARG, "too many padded rows for TMA coordinates");
  const int ksplit = proj_tc_ksplit(n_batch, m_tiles, max_n);
  Carver c(ws);
  __nv_bfloat16* a3[3];
  __nv_bfloat16* b3[3];
  for (int i = 0; i < 3; ++i) a3[i] = reinterpret_cast<__nv_bfloat16*>(c.take<uint16_t>(size_t(a_rows) * kpA));
  for (int i = 0; i < 3; ++i)
    b3[i] = reinterpret_cast<__nv_bfloat16*>(c.take<uint16_t>(b_presplit ? 0 : size_t(total_n) * kpB));
  float* partial = c.take<float>(size_t(ksplit) * n_batch * kpA * kpB);
  if (c.bytes() > ws_bytes) DM_FAIL(DM_ERR_WORKSPACE, "projection workspace too small: need %zu", c.bytes());

  const int wpb = 8;
  split3_kernel<double><<<unsigned((a_rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(
      A, ldA, off, n_batch, k, a_scale, nullptr, 0, nullptr, a3[0], a3[1], a3[2], kpA, a_stride, a_rows);
  DM_LAUNCH_OK("split3_kernel(A)");
  if (b_presplit) {
    for (int i = 0; i < 3; ++i) b3[i] = static_cast<__nv_bfloat16*>(const_cast<void*>(b_presplit[i]));
  } else if (total_n > 0) {
    if (Bf)
      split3_kernel<float><<<unsigned((total_n + wpb - 1) / wpb), wpb * 32, 0, st>>>(
          Bf, ldB, off, n_batch, d, b_scale, b_gather, b_gather_i64, b_gather_src_off, b3[0], b3[1], b3[2], kpB, 0, total_n);
    else
      split3_kernel<double><<<unsigned((total_n + wpb - 1) / wpb), wpb * 32, 0, st>>>(
          Bd, ldB, off, n_batch, d, b_scale, b_gather, b_gather_i64, b_gather_src_off, b3[0], b3[1], b3[2], kpB, 0, total_n);
    DM_LAUNCH_OK("split3_kernel(B)");
  }
  ProjMaps maps;
  int rc;
  for (int i = 0; i < 3; ++i) {
    if ((rc = make_map3(&maps.a[i], a3[i], a_rows, kpA, 2))) return rc;
    if ((rc = make_map3(&maps.b[i], b3[i], total_n, kpB, nb_groups))) return rc;
  }
  ProjParams P;
  P.b_off = off, P.a_stride_rows = a_stride, P.n_batch = n_batch, P.m_tiles = m_tiles, P.ksplit = ksplit;
  P.nb_groups = nb_groups, P.partial = partial;
  const uint32_t stage_bytes = 3 * A_BYTES + 3 * nb_groups * GROUP_BYTES;
  int stages = int((200 * 1024) / stage_bytes);
  P.stages = stages > 6 ? 6 : (stages < 2 ? 2 : stages);
  const size_t shm = size_t(P.stages) * stage_bytes + 8 * (2 * P.stages + 1) + 16 + 1024;
  static OncePerDevice attr_once;
  if (attr_once.first()) {
    DM_CUDA_OK(cudaFuncSetAttribute(proj_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  const int64_t nblk = int64_t(n_batch) * m_tiles * ksplit;
  if (nblk > 0x7fffffffLL) DM_FAIL(DM_ERR_BADARG, "projection grid too large");
  proj_tc_kernel<<<unsigned(nblk), kProjThreads, shm, st>>>(maps, P);
  DM_LAUNCH_OK("proj_tc_kernel");
  // (kpB and the strides are multiples of 64 floats: the float4 loads are aligned whenever d % 4 == 0)
  const bool vec = d % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const int64_t n_out = int64_t(n_batch) * k * (vec ? d / 4 : d);
  if (vec)
    proj_reduce_kernel<true><<<unsigned((n_out + 255) / 256), 256, 0, st>>>(partial, ksplit, n_batch, kpA, kpB, k, d, out);
  else
    proj_reduce_kernel<false><<<unsigned((n_out + 255) / 256), 256, 0, st>>>(partial, ksplit, n_batch, kpA, kpB, k, d, out);
  DM_LAUNCH_OK("proj_reduce_kernel");
  return DM_OK;
}

}  // namespace dm
