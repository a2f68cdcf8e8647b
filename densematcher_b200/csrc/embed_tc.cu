// Spectral embeddings on the tensor cores + on-demand float64, for the FM -> p2p pass.
//
// FM_to_p2p (densematcher/pyFM/spectral/convert.py:134-140) forms emb2 = Phi2 C and emb1 = Phi1 C^T in float64 before its
// two kd-tree searches.  The score pass that replaces the searches is fp32-grade anyway, and float64 is only needed for
// the ~1 % of results whose top-2 gap is inside the rounding bound -- so the two N x k x k float64 GEMMs (0.53 ms per
// 128 pairs, 30 % of the stage) are replaced by
//   (1) embed_tc_kernel: rows_i = Phi_i B on tcgen05 from three-way bf16 splits of both operands (six products, fp32
//       accumulation in tensor memory; the leading product hh and the five corrections accumulate in SEPARATE
//       accumulators so that the accumulation rounding of the long sum does not scale with the large term), whose
//       epilogue writes what the score pass needs directly: the bf16 hi/lo split of the embedding, row norms, the
//       Euclidean bias -1/2 |row|^2, the per-pair maxima of the error bound.  The embedding error
//       |d row_i| <= eps_e |Phi_i| |B|_F enters the re-evaluation threshold through an inflated row norm / bias bound,
//       so the exactness argument of the near-tie re-evaluation is unchanged;
//   (2) factored_fill_kernel: for every queued re-evaluation the float64 rows Phi_i B (and biases) it will read are
//       computed on demand by one warp each; results that need a scan of ALL candidates mark their pair, and only
//       marked pairs run the float64 GEMM (skip flags).
#include <cuda_bf16.h>
#include "dm_internal.cuh"
#include "gemm64.cuh"
#include "tc_ptx.cuh"

namespace dm {
namespace {

using namespace tc;

constexpr int EB_ROWS = 128;  // rows per CTA (UMMA M) and rows reserved per pair in the B operand (UMMA N)
constexpr int EB_BK = 64;
constexpr int EB_UK = 16;
constexpr uint32_t EB_TILE = EB_ROWS * EB_BK * 2;  // one [128 x 64] bf16 box
constexpr int EB_THREADS = 192;

__host__ __device__ constexpr int eb_a_stages(int kc) { return kc == 1 ? 2 : 1; }  // A buffers that fit beside the resident B
__host__ __device__ constexpr uint32_t eb_smem_bytes(int kc) { return uint32_t(kc) * 3 * (eb_a_stages(kc) + 1) * EB_TILE + 128 + 1024; }

__device__ __forceinline__ uint64_t eb_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= uint64_t((saddr >> 4) & 0x3FFF);
  d |= uint64_t(1) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}
constexpr uint32_t kEbIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(EB_ROWS >> 3) << 17) | (uint32_t(EB_ROWS >> 4) << 24);

// ---------------------------------------------------------------- B operand: three-way bf16 split of C or C^T per pair
// side 0: Bt[o][k] = C[k][o]   (rows o < k1, contraction k < k2: emb2 = Phi2 C)
// side 1: Bt[o][k] = C[o][k]   (rows o < k2, contraction k < k1: emb1 = Phi1 C^T)
// output rows are padded to rb (128 or 256) per pair and the contraction to kp (zeros); fro_part: chunk sums of |C|_F^2.
constexpr int kCsplitChunks = 8;  // CTAs per (pair, side): the kernel is a latency-bound walk over <= 256 x 256 entries
__global__ void __launch_bounds__(256)
    csplit_kernel(const double* __restrict__ C, int k1, int k2, int kp0, int kp1, __nv_bfloat16* __restrict__ h0,
                  __nv_bfloat16* __restrict__ m0, __nv_bfloat16* __restrict__ l0, __nv_bfloat16* __restrict__ h1,
                  __nv_bfloat16* __restrict__ m1, __nv_bfloat16* __restrict__ l1, double* __restrict__ fro_part,
                  int first_side, double* __restrict__ Ct, int rb) {
  const int p = blockIdx.x, side = blockIdx.y + first_side, chunk = blockIdx.z;
  if (Ct) {  // float64 transpose [k1, k2] for the on-demand products Phi1_j C^T
    for (int e = chunk * blockDim.x + threadIdx.x; e < k1 * k2; e += blockDim.x * kCsplitChunks)
      Ct[int64_t(p) * k1 * k2 + e] = C[int64_t(p) * k1 * k2 + int64_t(e % k2) * k1 + e / k2];
  }
  const double* Cp = C + int64_t(p) * k1 * k2;
  const int n_out = side == 0 ? k1 : k2, n_in = side == 0 ? k2 : k1, kp = side == 0 ? kp0 : kp1;
  __nv_bfloat16* h = (side == 0 ? h0 : h1) + int64_t(p) * rb * kp;
  __nv_bfloat16* m = (side == 0 ? m0 : m1) + int64_t(p) * rb * kp;
  __nv_bfloat16* l = (side == 0 ? l0 : l1) + int64_t(p) * rb * kp;
  double ss = 0.0;
  const int rows_per = rb / kCsplitChunks;
  for (int e = chunk * rows_per * kp + threadIdx.x; e < (chunk + 1) * rows_per * kp; e += blockDim.x) {
    const int o = e / kp, k = e % kp;
    double v = 0.0;
    if (o < n_out && k < n_in) v = side == 0 ? Cp[int64_t(k) * k1 + o] : Cp[int64_t(o) * k1 + k];
    ss = fma(v, v, ss);
    const float x = float(v);
    const __nv_bfloat16 vh = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(vh);
    const __nv_bfloat16 vm = __float2bfloat16_rn(r1);
    h[e] = vh, m[e] = vm, l[e] = __float2bfloat16_rn(r1 - __bfloat162float(vm));
  }
  if (side == first_side) {
    __shared__ double red[8];
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, sh);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) ss += red[w];
      fro_part[p * kCsplitChunks + chunk] = ss;
    }
  }
}
// |C_p|_F rounded up, from the chunk sums in fixed order (evaluated by the consumers: no kernel of its own)
__device__ __forceinline__ float cfro_of(const double* __restrict__ fro_part, int p) {
  double ss = 0.0;
#pragma unroll
  for (int c = 0; c < kCsplitChunks; ++c) ss += fro_part[p * kCsplitChunks + c];
  return __double2float_ru(sqrt(ss)) * 1.000001f;
}

// ---------------------------------------------------------------- the embedding kernel
struct EbMaps {
  CUtensorMap a[3], b[3];
};

struct EbEpi {          // one argmax epilogue whose scale / bias live on the embedded side
  int bias_sqnorm;      // bias_i = -1/2 |row_i|^2 (else 0)
  const double* scale;  // optional scale array (else 1)
  float *sf, *bf, *G, *Bm;
  double *sd, *bd;      // float64 scale / bias read by the re-evaluation (nullptr: left alone); the float64 bias of a
                        // bias_sqnorm epilogue is filled on demand, only the scale is final here
};

struct EbParams {
  const int64_t* off;       // rows of pair p: off[p] .. off[p + 1]
  const int64_t* a_in;      // first row of pair p in the A operand (its split and a_norm): off, or the mesh-bank rows
  int max_rt, k_out, kp_out, n_epi;
  __nv_bfloat16 *hi, *lo;   // [total, kp_out] split of the embedding (nullptr: norms / biases only)
  float* norm;              // inflated row norm (nullptr: not wanted)
  const float* a_norm;      // |Phi_i| (rounded up)
  const double* fro_part;   // [n_pairs, kCsplitChunks] sums of squares of C: |C|_F = cfro_of(fro_part, p)
  float eps_e;              // |d row_i| <= eps_e |Phi_i| |C|_F
  float inv_eps;            // 1 / eps of the score pass
  EbEpi epi[kMaxEpi];
};

__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
  atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// A CTA owns EB_TPC consecutive 128-row tiles of one pair: the B operand (the split of C, up to 96 KB) is fetched once and
// stays resident, the A tiles (48 KB per 64-wide K chunk triple) and the two accumulators are double-buffered, so the TMA
// loads of tile t + 1 and the epilogue of tile t - 1 overlap the MMAs of tile t (the first version ran load -> MMA ->
// epilogue back to back in a one-CTA-per-SM kernel: 143 us per embedding).
constexpr int EB_TPC = 4;

template <int KC>
__global__ void __launch_bounds__(EB_THREADS, 1) embed_tc_kernel(const __grid_constant__ EbMaps maps, const EbParams P) {
  const int ctas_per_pair = (P.max_rt + EB_TPC - 1) / EB_TPC;
  const int p = blockIdx.x / ctas_per_pair, rt0 = (blockIdx.x % ctas_per_pair) * EB_TPC;
  const int64_t r0 = P.off[p];
  const int n = int(P.off[p + 1] - r0);
  if (rt0 * EB_ROWS >= n) return;
  const int n_tiles = min(EB_TPC, (n - rt0 * EB_ROWS + EB_ROWS - 1) / EB_ROWS);
  const float cfro_p = cfro_of(P.fro_part, p);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  constexpr uint32_t A_BYTES = KC * 3 * EB_TILE;
  constexpr int AST = eb_a_stages(KC);
  constexpr uint32_t OFF_B = AST * A_BYTES, OFF_BAR = OFF_B + KC * 3 * EB_TILE;
  const uint32_t bar_b = sbase + OFF_BAR;       // B resident
  const uint32_t bar_afull = bar_b + 8;         // [2]
  const uint32_t bar_aempty = bar_afull + 16;   // [2]
  const uint32_t bar_tfull = bar_aempty + 16;   // [2]
  const uint32_t bar_tempty = bar_tfull + 16;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + OFF_BAR + 80);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_b, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_afull + 8 * i, 1);
      mbar_init(bar_aempty + 8 * i, 1);
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 4);
    }
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
      mbar_expect_tx(bar_b, KC * 3 * EB_TILE);
      const int brow = p * EB_ROWS;
#pragma unroll
      for (int kc = 0; kc < KC; ++kc)
#pragma unroll
        for (int i = 0; i < 3; ++i) tma_load_2d(sbase + OFF_B + (kc * 3 + i) * EB_TILE, &maps.b[i], kc * EB_BK, brow, bar_b);
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % AST;
        mbar_wait_backoff(bar_aempty + 8 * s, ((t / AST) & 1) ^ 1);
        mbar_expect_tx(bar_afull + 8 * s, A_BYTES);
        const int arow = int(P.a_in[p] + int64_t(rt0 + t) * EB_ROWS);
#pragma unroll
        for (int kc = 0; kc < KC; ++kc)
#pragma unroll
          for (int i = 0; i < 3; ++i)
            tma_load_2d(sbase + s * A_BYTES + (kc * 3 + i) * EB_TILE, &maps.a[i], kc * EB_BK, arow, bar_afull + 8 * s);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait_backoff(bar_b, 0);
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t & 1, sa = t % AST;
        mbar_wait_backoff(bar_tempty + 8 * s, ((t >> 1) & 1) ^ 1);
        mbar_wait_backoff(bar_afull + 8 * sa, (t / AST) & 1);
        tc_fence_after();
        const uint32_t t_hh = tmem_base + s * 2 * EB_ROWS, t_cor = t_hh + EB_ROWS;
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) {
          uint64_t da[3], db[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            da[i] = eb_desc_sw128(sbase + sa * A_BYTES + (kc * 3 + i) * EB_TILE);
            db[i] = eb_desc_sw128(sbase + OFF_B + (kc * 3 + i) * EB_TILE);
          }
#pragma unroll
          for (int k = 0; k < EB_BK / EB_UK; ++k) {
            const uint64_t ko = uint64_t((k * EB_UK * 2) >> 4);
            const uint32_t first = (kc | k) != 0;
            tc_mma_bf16(t_hh, da[0] + ko, db[0] + ko, kEbIdesc, first);   // h h
            tc_mma_bf16(t_cor, da[0] + ko, db[1] + ko, kEbIdesc, first);  // h m
            tc_mma_bf16(t_cor, da[1] + ko, db[0] + ko, kEbIdesc, 1);      // m h
            tc_mma_bf16(t_cor, da[1] + ko, db[1] + ko, kEbIdesc, 1);      // m m
            tc_mma_bf16(t_cor, da[0] + ko, db[2] + ko, kEbIdesc, 1);      // h l
            tc_mma_bf16(t_cor, da[2] + ko, db[0] + ko, kEbIdesc, 1);      // l h
          }
        }
        tc_commit(bar_aempty + 8 * sa);
        tc_commit(bar_tfull + 8 * s);
      }
    }
  } else {
    const int q = warp & 3;
    for (int t = 0; t < n_tiles; ++t) {
      const int s = t & 1;
      const int i = (rt0 + t) * EB_ROWS + 32 * q + lane;  // row inside the pair
      const bool ok = i < n;
      const int64_t gi = r0 + i;
      mbar_wait(bar_tfull + 8 * s, (t >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + s * 2 * EB_ROWS + (uint32_t(32 * q) << 16);
      float ss = 0.f;
      const int n_ch = P.kp_out / 32;
      for (int ch = 0; ch < n_ch; ++ch) {
        float a[32], c[32];
        if (ch * 32 < P.k_out) {
          tmem_ld32(taddr + ch * 32, a);
          tmem_ld32(taddr + EB_ROWS + ch * 32, c);
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) a[e] = c[e] = 0.f;
        }
        uint32_t ph[16], pl[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float y0 = a[e] + c[e], y1 = a[e + 1] + c[e + 1];
          ss = fmaf(y0, y0, ss);
          ss = fmaf(y1, y1, ss);
          const __nv_bfloat16 h0 = __float2bfloat16_rn(y0), h1 = __float2bfloat16_rn(y1);
          const __nv_bfloat16 l0 = __float2bfloat16_rn(y0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(y1 - __bfloat162float(h1));
          ph[e >> 1] = uint32_t(__bfloat16_as_ushort(h0)) | (uint32_t(__bfloat16_as_ushort(h1)) << 16);
          pl[e >> 1] = uint32_t(__bfloat16_as_ushort(l0)) | (uint32_t(__bfloat16_as_ushort(l1)) << 16);
        }
        if (ok && P.hi) {
          uint4* dh = reinterpret_cast<uint4*>(P.hi + gi * P.kp_out + ch * 32);
          uint4* dl = reinterpret_cast<uint4*>(P.lo + gi * P.kp_out + ch * 32);
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            dh[v4] = make_uint4(ph[4 * v4], ph[4 * v4 + 1], ph[4 * v4 + 2], ph[4 * v4 + 3]);
            dl[v4] = make_uint4(pl[4 * v4], pl[4 * v4 + 1], pl[4 * v4 + 2], pl[4 * v4 + 3]);
          }
        }
      }
      // the accumulators of this tile are free again
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * s);
      // |row| rounded up (fp32 sum of squares: relative error <= (k_out + 2) 2^-24), embedding error e_i, inflated norm
      const float nrm = ok ? __fsqrt_ru(ss * (1.f + float(P.k_out + 4) * 6.0e-8f)) * 1.000001f : 0.f;
      const float e_i = ok ? P.eps_e * P.a_norm[P.a_in[p] + i] * cfro_p : 0.f;
      const float own = (nrm + e_i) + e_i * P.inv_eps;  // eps * own >= eps |row| + e_i
      if (ok && P.norm) P.norm[gi] = own;
      for (int e = 0; e < P.n_epi; ++e) {
        const EbEpi& E = P.epi[e];
        float g = 0.f, bm = 0.f;
        if (ok) {
          const float sc = E.scale ? float(E.scale[gi]) : 1.f;
          const float bi = E.bias_sqnorm ? -0.5f * ss : 0.f;
          E.sf[gi] = sc;
          E.bf[gi] = bi;
          if (E.sd) E.sd[gi] = E.scale ? E.scale[gi] : 1.0;
          if (E.bd) E.bd[gi] = 0.0;
          g = own * fabsf(sc) * 1.000001f;
          // the float64 bias differs from bi by <= |row| e_i + e_i^2 / 2 + the fp32 rounding of the sum of squares;
          // it enters the threshold through Bm (emit_result: 9.6e-7 (|s| + Bm)), hence the division
          const float db = E.bias_sqnorm ? (nrm * e_i + 0.5f * e_i * e_i + fabsf(bi) * float(P.k_out + 4) * 1.2e-7f) : 0.f;
          bm = fabsf(bi) + db * (2.f / 9.6e-7f);
        }
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) {
          g = fmaxf(g, __shfl_xor_sync(0xffffffffu, g, sh));
          bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, sh));
        }
        if (lane == 0) {
          atomic_max_nonneg(E.G + p, g);
          atomic_max_nonneg(E.Bm + p, bm);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------- wide embeddings: 128 < k <= 256 (upper ZoomOut rungs)
// The split of C no longer fits beside the A tiles (3 x 256 rows x 256 columns of bf16 = 384 KB), so BOTH operands are
// streamed per 64-wide K chunk through a two-stage ring (48 KB of A + 48 KB of B per stage; B comes from L2, every CTA of a
// pair re-reads it), and the output columns are produced in halves of 128 -- a work unit is (row tile, half), its two
// accumulators (hh / corrections) double-buffered in tensor memory exactly as above.  The epilogue thread carries the
// running sum of squares of its row across the halves and finalises norm / bias / maxima after the last one.
constexpr uint32_t EBS_STAGE = 6 * EB_TILE;  // A triple + B triple of one K chunk: 96 KB
constexpr uint32_t kEbsSmem = 2 * EBS_STAGE + 128 + 1024;

__global__ void __launch_bounds__(EB_THREADS, 1)
    embed_tc_stream_kernel(const __grid_constant__ EbMaps maps, const EbParams P, const int n_kc, const int n_half) {
  const int ctas_per_pair = (P.max_rt + EB_TPC - 1) / EB_TPC;
  const int p = blockIdx.x / ctas_per_pair, rt0 = (blockIdx.x % ctas_per_pair) * EB_TPC;
  const int64_t r0 = P.off[p];
  const int n = int(P.off[p + 1] - r0);
  if (rt0 * EB_ROWS >= n) return;
  const int n_tiles = min(EB_TPC, (n - rt0 * EB_ROWS + EB_ROWS - 1) / EB_ROWS);
  const float cfro_p = cfro_of(P.fro_part, p);
  const int n_units = n_tiles * n_half;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  constexpr uint32_t OFF_BAR = 2 * EBS_STAGE;
  const uint32_t bar_full = sbase + OFF_BAR;    // [2]
  const uint32_t bar_empty = bar_full + 16;     // [2]
  const uint32_t bar_tfull = bar_empty + 16;    // [2]
  const uint32_t bar_tempty = bar_tfull + 16;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + OFF_BAR + 80);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 4);
    }
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
      for (int u = 0; u < n_units; ++u) {
        const int t = u / n_half, h = u % n_half;
        const int arow = int(P.a_in[p] + int64_t(rt0 + t) * EB_ROWS), brow = (p * n_half + h) * EB_ROWS;
        for (int kc = 0; kc < n_kc; ++kc) {
          mbar_wait_backoff(bar_empty + 8 * stage, phase ^ 1);
          mbar_expect_tx(bar_full + 8 * stage, EBS_STAGE);
          const uint32_t sb = sbase + stage * EBS_STAGE;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            tma_load_2d(sb + i * EB_TILE, &maps.a[i], kc * EB_BK, arow, bar_full + 8 * stage);
            tma_load_2d(sb + (3 + i) * EB_TILE, &maps.b[i], kc * EB_BK, brow, bar_full + 8 * stage);
          }
          if (++stage == 2) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = 0; u < n_units; ++u) {
        const int s = u & 1;
        mbar_wait_backoff(bar_tempty + 8 * s, ((u >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t t_hh = tmem_base + s * 2 * EB_ROWS, t_cor = t_hh + EB_ROWS;
        for (int kc = 0; kc < n_kc; ++kc) {
          mbar_wait_backoff(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sb = sbase + stage * EBS_STAGE;
          uint64_t da[3], db[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            da[i] = eb_desc_sw128(sb + i * EB_TILE);
            db[i] = eb_desc_sw128(sb + (3 + i) * EB_TILE);
          }
#pragma unroll
          for (int k = 0; k < EB_BK / EB_UK; ++k) {
            const uint64_t ko = uint64_t((k * EB_UK * 2) >> 4);
            const uint32_t first = (kc | k) != 0;
            tc_mma_bf16(t_hh, da[0] + ko, db[0] + ko, kEbIdesc, first);   // h h
            tc_mma_bf16(t_cor, da[0] + ko, db[1] + ko, kEbIdesc, first);  // h m
            tc_mma_bf16(t_cor, da[1] + ko, db[0] + ko, kEbIdesc, 1);      // m h
            tc_mma_bf16(t_cor, da[1] + ko, db[1] + ko, kEbIdesc, 1);      // m m
            tc_mma_bf16(t_cor, da[0] + ko, db[2] + ko, kEbIdesc, 1);      // h l
            tc_mma_bf16(t_cor, da[2] + ko, db[0] + ko, kEbIdesc, 1);      // l h
          }
          tc_commit(bar_empty + 8 * stage);
          if (++stage == 2) stage = 0, phase ^= 1;
        }
        tc_commit(bar_tfull + 8 * s);
      }
    }
  } else {
    const int q = warp & 3;
    for (int t = 0; t < n_tiles; ++t) {
      const int i = (rt0 + t) * EB_ROWS + 32 * q + lane;  // row inside the pair
      const bool ok = i < n;
      const int64_t gi = r0 + i;
      float ss = 0.f;
      for (int h = 0; h < n_half; ++h) {
        const int u = t * n_half + h, s = u & 1;
        mbar_wait(bar_tfull + 8 * s, (u >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + s * 2 * EB_ROWS + (uint32_t(32 * q) << 16);
        const int col_base = h * EB_ROWS;
        const int n_ch = min(4, (P.kp_out - col_base) / 32);
        for (int ch = 0; ch < n_ch; ++ch) {
          float a[32], c[32];
          if (col_base + ch * 32 < P.k_out) {
            tmem_ld32(taddr + ch * 32, a);
            tmem_ld32(taddr + EB_ROWS + ch * 32, c);
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) a[e] = c[e] = 0.f;
          }
          uint32_t ph[16], pl[16];
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float y0 = a[e] + c[e], y1 = a[e + 1] + c[e + 1];
            ss = fmaf(y0, y0, ss);
            ss = fmaf(y1, y1, ss);
            const __nv_bfloat16 h0 = __float2bfloat16_rn(y0), h1 = __float2bfloat16_rn(y1);
            const __nv_bfloat16 l0 = __float2bfloat16_rn(y0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(y1 - __bfloat162float(h1));
            ph[e >> 1] = uint32_t(__bfloat16_as_ushort(h0)) | (uint32_t(__bfloat16_as_ushort(h1)) << 16);
            pl[e >> 1] = uint32_t(__bfloat16_as_ushort(l0)) | (uint32_t(__bfloat16_as_ushort(l1)) << 16);
          }
          if (ok && P.hi) {
            uint4* dh = reinterpret_cast<uint4*>(P.hi + gi * P.kp_out + col_base + ch * 32);
            uint4* dl = reinterpret_cast<uint4*>(P.lo + gi * P.kp_out + col_base + ch * 32);
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) {
              dh[v4] = make_uint4(ph[4 * v4], ph[4 * v4 + 1], ph[4 * v4 + 2], ph[4 * v4 + 3]);
              dl[v4] = make_uint4(pl[4 * v4], pl[4 * v4 + 1], pl[4 * v4 + 2], pl[4 * v4 + 3]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * s);
      }
      const float nrm = ok ? __fsqrt_ru(ss * (1.f + float(P.k_out + 4) * 6.0e-8f)) * 1.000001f : 0.f;
      const float e_i = ok ? P.eps_e * P.a_norm[P.a_in[p] + i] * cfro_p : 0.f;
      const float own = (nrm + e_i) + e_i * P.inv_eps;
      if (ok && P.norm) P.norm[gi] = own;
      for (int e = 0; e < P.n_epi; ++e) {
        const EbEpi& E = P.epi[e];
        float g = 0.f, bm = 0.f;
        if (ok) {
          const float sc = E.scale ? float(E.scale[gi]) : 1.f;
          const float bi = E.bias_sqnorm ? -0.5f * ss : 0.f;
          E.sf[gi] = sc;
          E.bf[gi] = bi;
          if (E.sd) E.sd[gi] = E.scale ? E.scale[gi] : 1.0;
          if (E.bd) E.bd[gi] = 0.0;
          g = own * fabsf(sc) * 1.000001f;
          const float db = E.bias_sqnorm ? (nrm * e_i + 0.5f * e_i * e_i + fabsf(bi) * float(P.k_out + 4) * 1.2e-7f) : 0.f;
          bm = fabsf(bi) + db * (2.f / 9.6e-7f);
        }
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) {
          g = fmaxf(g, __shfl_xor_sync(0xffffffffu, g, sh));
          bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, sh));
        }
        if (lane == 0) {
          atomic_max_nonneg(E.G + p, g);
          atomic_max_nonneg(E.Bm + p, bm);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------- on-demand float64 for the queued re-evaluations
struct FillParams {
  // Y side: query rows are Phi2[:, :k2] C                      (row i -> emb64[q0 + i], k1 values)
  const double* Phi2;
  int64_t ld2;
  // X-side bias of row epilogue `row_bias_epi`: -1/2 |C Phi1_j|^2   (-1: none)
  const double* Phi1;
  int64_t ld1;
  const double* C;  // [n_pairs, k2, k1]
  int k1, k2;
  double* emb64;    // [total_q, k1]
  int row_bias_epi, col_bias_epi;
  double *row_bd, *col_bd;  // float64 bias arrays of those epilogues
  int* skip_y;      // per pair: cleared when a result needs ALL float64 query rows of the pair
  int* skip_x;      // ... all float64 database-side biases
  const int64_t *in1, *in2;  // first row of pair p in Phi1 / Phi2 (the batch offsets, or mesh-bank rows)
};

// out[o] = sum_k phi[k] C[k][o], o < k1 (lanes over o: coalesced rows of C); returns sum_o out[o]^2 (all lanes).
// The row of Phi is held in registers (lane l: entries l, l + 32, ...) and broadcast by shuffles; four rows of C are in
// flight per step (the loop is otherwise one dependent L2 round trip per k: 15 us per product).  T slabs of 32: k <= 32 T.
template <int T>
__device__ __forceinline__ double warp_row_times_C(const double* __restrict__ phi, const double* __restrict__ Cp, int k1,
                                                   int k2, int lane, double* __restrict__ out) {
  double f[T];
#pragma unroll
  for (int t = 0; t < T; ++t) f[t] = (lane + 32 * t < k2) ? phi[lane + 32 * t] : 0.0;
  double acc[T];
#pragma unroll
  for (int a = 0; a < T; ++a) acc[a] = 0.0;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    if (32 * t >= k2) break;
    for (int kk = 0; kk < 32 && 32 * t + kk < k2; kk += 4) {
      double c[4][T];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = min(32 * t + kk + u, k2 - 1);
        const double* row = Cp + int64_t(k) * k1;
#pragma unroll
        for (int a = 0; a < T; ++a) c[u][a] = lane + 32 * a < k1 ? row[lane + 32 * a] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        double fk = __shfl_sync(0xffffffffu, f[t], kk + u);
        if (32 * t + kk + u >= k2) fk = 0.0;
#pragma unroll
        for (int a = 0; a < T; ++a) acc[a] = fma(fk, c[u][a], acc[a]);
      }
    }
  }
  double ss = 0.0;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int o = lane + 32 * t;
    if (o < k1) {
      if (out) out[o] = acc[t];
      ss = fma(acc[t], acc[t], ss);
    }
  }
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, sh);
  return ss;
}
// sum_o (sum_k C[o][k] phi[k])^2, o < k2 (lanes over k: coalesced rows of C), four rows in flight
__device__ __forceinline__ double warp_sqnorm_C_times_row(const double* __restrict__ phi, const double* __restrict__ Cp,
                                                          int k1, int k2, int lane) {
  double f[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) f[t] = (lane + 32 * t < k1) ? phi[lane + 32 * t] : 0.0;
  double ss = 0.0;
  for (int o = 0; o < k2; o += 4) {
    double s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double* row = Cp + int64_t(min(o + u, k2 - 1)) * k1;
      double a = 0.0;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (lane + 32 * t < k1) a = fma(row[lane + 32 * t], f[t], a);
      s[u] = a;
    }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1)
#pragma unroll
      for (int u = 0; u < 4; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], sh);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (o + u < k2) ss = fma(s[u], s[u], ss);
  }
  return ss;
}

__global__ void __launch_bounds__(256) factored_fill_kernel(const NNProblem P, const FillParams F) {
  const int lane = threadIdx.x & 31;
  const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarp = gridDim.x * (blockDim.x >> 5);
  const unsigned n_cand = P.counters[0], n_full = P.counters[3];
  for (unsigned f = warp; f < n_cand + n_full; f += nwarp) {
    const bool full = f >= n_cand;
    const FlagEntry e = full ? P.flags[P.flag_cap - 1 - int64_t(f - n_cand)] : P.flags[f];
    const int p = e.pair, epi = e.epi & 255;
    const bool is_col = (e.epi & 256) != 0;
    const int64_t q0 = P.q_off[p], d0 = P.db_off[p];
    const double* Cp = F.C + int64_t(p) * F.k1 * F.k2;
    if (!is_col) {
      // the result of query row `local`: its float64 embedding row, and the float64 biases of its candidates
      warp_row_times_C<4>(F.Phi2 + (F.in2[p] + e.local) * F.ld2, Cp, F.k1, F.k2, lane, F.emb64 + (q0 + e.local) * F.k1);
      if (epi == F.row_bias_epi) {
        if (full) {
          if (lane == 0) F.skip_x[p] = 0;
        } else {
          for (int c = 0; c < 2; ++c) {
            const int j = c == 0 ? e.c1 : e.c2;
            const double ss = warp_sqnorm_C_times_row(F.Phi1 + (F.in1[p] + j) * F.ld1, Cp, F.k1, F.k2, lane);
            if (lane == 0) F.row_bd[d0 + j] = -0.5 * ss;
          }
        }
      }
    } else {
      // the result of database row `local`: the float64 embedding rows (and biases) of its query-side candidates
      if (full) {
        if (lane == 0) F.skip_y[p] = 0;
      } else {
        for (int c = 0; c < 2; ++c) {
          const int i = c == 0 ? e.c1 : e.c2;
          const double ss = warp_row_times_C<4>(F.Phi2 + (F.in2[p] + i) * F.ld2, Cp, F.k1, F.k2, lane, F.emb64 + (q0 + i) * F.k1);
          if (lane == 0 && epi == F.col_bias_epi) F.col_bd[q0 + i] = -0.5 * ss;
        }
      }
    }
  }
}

// bias[row] = -1/2 |M[row, :d]|^2 for the rows of the pairs whose skip flag is clear (one warp per row)
__global__ void __launch_bounds__(256)
    bias_rows_flagged_kernel(const double* __restrict__ M, int64_t ld, const int64_t* __restrict__ off, int max_n, int d,
                             const int* __restrict__ skip, double* __restrict__ bias) {
  const int p = blockIdx.y;
  if (skip[p]) return;
  const int lane = threadIdx.x & 31;
  const int64_t r0 = off[p];
  const int n = int(off[p + 1] - r0);
  for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += gridDim.x * 8) {
    const double* r = M + (r0 + i) * ld;
    double s = 0.0;
    for (int k = lane; k < d; k += 32) s = fma(r[k], r[k], s);
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
    if (lane == 0) bias[r0 + i] = -0.5 * s;
  }
}

template <int KC>
int eb_launch(const EbMaps& maps, const EbParams& P, int n_pairs, cudaStream_t st) {
  static OncePerDevice once;
  if (once.first())
    DM_CUDA_OK(cudaFuncSetAttribute(embed_tc_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(eb_smem_bytes(KC))));
  const int64_t nblk = int64_t(n_pairs) * ((P.max_rt + EB_TPC - 1) / EB_TPC);
  if (nblk > 0x7fffffffLL) DM_FAIL(DM_ERR_BADARG, "too many row tiles");
  embed_tc_kernel<KC><<<unsigned(nblk), EB_THREADS, eb_smem_bytes(KC), st>>>(maps, P);
  DM_LAUNCH_OK("embed_tc_kernel");
  return DM_OK;
}

// kp_in <= 128 and k_out <= 128: resident-B kernel; wider: the streaming kernel (the B operand then holds
// 128 * ceil(k_out / 128) rows per pair)
int eb_b_rows(int k_out) { return k_out <= EB_ROWS ? EB_ROWS : 2 * EB_ROWS; }
int eb_run(const void* const a3[3], int64_t a_rows, const void* const b3[3], int kp_in, const EbParams& P, int n_pairs,
           cudaStream_t st) {
  EbMaps maps;
  int rc;
  const bool wide = kp_in > 2 * EB_BK || P.k_out > EB_ROWS;
  const int rb = wide ? eb_b_rows(P.k_out) : EB_ROWS;
  for (int i = 0; i < 3; ++i) {
    if ((rc = tc_make_map_bf16(&maps.a[i], a3[i], a_rows, kp_in, EB_ROWS))) return rc;
    if ((rc = tc_make_map_bf16(&maps.b[i], b3[i], int64_t(n_pairs) * rb, kp_in, EB_ROWS))) return rc;
  }
  if (!wide) return kp_in <= EB_BK ? eb_launch<1>(maps, P, n_pairs, st) : eb_launch<2>(maps, P, n_pairs, st);
  static OncePerDevice once;
  if (once.first())
    DM_CUDA_OK(cudaFuncSetAttribute(embed_tc_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kEbsSmem)));
  const int64_t nblk = int64_t(n_pairs) * ((P.max_rt + EB_TPC - 1) / EB_TPC);
  if (nblk > 0x7fffffffLL) DM_FAIL(DM_ERR_BADARG, "too many row tiles");
  embed_tc_stream_kernel<<<unsigned(nblk), EB_THREADS, kEbsSmem, st>>>(maps, P, kp_in / EB_BK, rb / EB_ROWS);
  DM_LAUNCH_OK("embed_tc_stream_kernel");
  return DM_OK;
}

// ---------------------------------------------------------------- the FM -> p2p hooks
struct F2PCtx {
  // inputs
  const double *C, *Phi1, *Phi2;
  int64_t ld1, ld2;
  const int64_t *off1, *off2;
  int64_t total_n1, total_n2;
  int max_n1, max_n2, n_pairs, k1, k2;
  int row_bias_epi, col_bias_epi;  // epilogue numbers of p2p_21 / p2p_12 (-1: not requested)
  // meshes served from a bank (three-way splits of Phi[:, :k] and row norms prepared once per mesh, dm_bank_prepare):
  // Phi1 / Phi2 are then the bank's matrix and in1 / in2 the first bank row of each pair's mesh; else in = off
  const NNBankSide *bank1, *bank2;
  const int64_t *in1, *in2;
  // scratch
  uint16_t *p2h, *p2m, *p2l;  // three-way split of Phi2 [total_n2, kp2]
  float* p2norm;
  uint16_t *c0h, *c0m, *c0l, *c1h, *c1m, *c1l;  // B operands [n_pairs * 128, kp]
  double* fro_part;  // [n_pairs, kCsplitChunks] partial sums of squares of C (|C|_F: cfro_of)
  float* g_scratch;
  double* emb2;   // [total_n2, k1] float64 query rows, filled on demand
  double* emb1;   // [total_n1, k2] float64, only for pairs that need every bias
  int *skip_y, *skip_x;
  int kp1, kp2;   // pad64(k1), pad64(k2)
};

float eb_eps(int kp_in) {
  // hh accumulates kp/16 MMAs into its own accumulator (2^-23 each with 2x slack for the truncating accumulate), the
  // corrections are 2^-8 smaller; + the final fp32 add and the dropped products (ml, lm, ll, remainders: 2^-25)
  return float((double(kp_in / 16) + 3.0) * 1.1920928955078125e-07 + 3.0e-8);
}

int f2p_prep_y(void* vctx, const NNLayout& L, NNProblem& P, const NNRequest& R, cudaStream_t st) {
  F2PCtx& X = *static_cast<F2PCtx*>(vctx);
  int rc;
  // three-way split + row norms of Phi2 (the A operand of the embedding), unless the bank holds them
  if (!X.bank2 && (rc = nn_prep_side(X.Phi2, 1, X.ld2, X.off2, X.n_pairs, X.total_n2, X.k2, X.p2norm, nullptr, 0, X.p2h, X.p2m,
                                     X.p2l, X.kp2, st)))
    return rc;
  csplit_kernel<<<dim3(unsigned(X.n_pairs), 2, kCsplitChunks), 256, 0, st>>>(
      X.C, X.k1, X.k2, X.kp2, X.kp1, reinterpret_cast<__nv_bfloat16*>(X.c0h), reinterpret_cast<__nv_bfloat16*>(X.c0m),
      reinterpret_cast<__nv_bfloat16*>(X.c0l), reinterpret_cast<__nv_bfloat16*>(X.c1h),
      reinterpret_cast<__nv_bfloat16*>(X.c1m), reinterpret_cast<__nv_bfloat16*>(X.c1l), X.fro_part, 0, nullptr, EB_ROWS);
  DM_LAUNCH_OK("csplit_kernel");
  EbParams E{};
  E.off = X.off2, E.a_in = X.in2, E.max_rt = (X.max_n2 + EB_ROWS - 1) / EB_ROWS, E.k_out = X.k1, E.kp_out = P.kp, E.n_epi = R.n_col;
  E.hi = reinterpret_cast<__nv_bfloat16*>(L.yh), E.lo = reinterpret_cast<__nv_bfloat16*>(L.yl);
  E.norm = L.norm_q, E.a_norm = X.bank2 ? X.bank2->norm : X.p2norm, E.fro_part = X.fro_part;
  E.eps_e = eb_eps(X.kp2), E.inv_eps = 1.f / P.eps;
  for (int e = 0; e < R.n_col; ++e) {
    if (R.col[e].scale_mode == DM_SCALE_INVNORM || R.col[e].bias_mode == DM_BIAS_ARRAY)
      DM_FAIL(DM_ERR_UNSUPPORTED, "factored query side: unsupported column epilogue");
    E.epi[e] = EbEpi{R.col[e].bias_mode == DM_BIAS_NEG_HALF_SQNORM, R.col[e].scale_mode == DM_SCALE_ARRAY ? R.col[e].scale : nullptr,
                     L.col[e].sf, L.col[e].bf, L.col[e].G, L.col[e].Bm, L.col[e].sd, L.col[e].bd};
    DM_CUDA_OK(cudaMemsetAsync(L.col[e].G, 0, sizeof(float) * X.n_pairs, st));
    DM_CUDA_OK(cudaMemsetAsync(L.col[e].Bm, 0, sizeof(float) * X.n_pairs, st));
  }
  const void* a3[3] = {X.p2h, X.p2m, X.p2l};
  const void* a3b[3] = {X.bank2 ? X.bank2->hi : nullptr, X.bank2 ? X.bank2->lo : nullptr, X.bank2 ? X.bank2->lo2 : nullptr};
  const void* b3[3] = {X.c0h, X.c0m, X.c0l};
  return eb_run(X.bank2 ? a3b : a3, X.bank2 ? X.bank2->rows : X.total_n2, b3, X.kp2, E, X.n_pairs, st);
}

// (runs after nn_prep_side of the database side = Phi1, which left its three-way split in L.xh / xl / xl2)
int f2p_after_prep(void* vctx, const NNLayout& L, NNProblem& P, cudaStream_t st) {
  F2PCtx& X = *static_cast<F2PCtx*>(vctx);
  if (X.row_bias_epi < 0) return DM_OK;
  const int e = X.row_bias_epi;
  EbParams E{};
  E.off = X.off1, E.a_in = X.in1, E.max_rt = (X.max_n1 + EB_ROWS - 1) / EB_ROWS, E.k_out = X.k2, E.kp_out = X.kp2, E.n_epi = 1;
  E.hi = E.lo = nullptr, E.norm = nullptr;
  E.a_norm = X.bank1 ? X.bank1->norm : L.norm_db, E.fro_part = X.fro_part;  // (L.norm_db is the batch-packed copy of the same values)
  E.eps_e = eb_eps(X.kp1), E.inv_eps = 1.f / P.eps;
  // only bf / Bm of the bias epilogue change: the scale stays 1 (the kernel rewrites sf = 1) and G must stay
  // max_j |Phi1_j| from nn_prep_side, so the kernel's G (the norm of emb1, not wanted) goes to a scratch array
  E.epi[0] = EbEpi{1, nullptr, L.row[e].sf, L.row[e].bf, X.g_scratch, L.row[e].Bm, nullptr, nullptr};
  DM_CUDA_OK(cudaMemsetAsync(L.row[e].Bm, 0, sizeof(float) * X.n_pairs, st));
  DM_CUDA_OK(cudaMemsetAsync(X.g_scratch, 0, sizeof(float) * X.n_pairs, st));
  const void* a3[3] = {L.xh, L.xl, L.xl2};  // (the bank's splits when Phi1 is served from a bank, see nn_run)
  const void* b3[3] = {X.c1h, X.c1m, X.c1l};
  (void)P;
  return eb_run(a3, X.bank1 ? X.bank1->rows : X.total_n1, b3, X.kp1, E, X.n_pairs, st);
}

int f2p_before_recheck(void* vctx, const NNLayout& L, NNProblem& P, cudaStream_t st) {
  F2PCtx& X = *static_cast<F2PCtx*>(vctx);
  // skip flags: 1 = nothing to do for the pair (the fill kernel clears them)
  DM_CUDA_OK(cudaMemsetAsync(X.skip_y, 1, sizeof(int) * X.n_pairs, st));
  DM_CUDA_OK(cudaMemsetAsync(X.skip_x, 1, sizeof(int) * X.n_pairs, st));
  FillParams F{};
  F.Phi2 = X.Phi2, F.ld2 = X.ld2, F.Phi1 = X.Phi1, F.ld1 = X.ld1, F.C = X.C, F.k1 = X.k1, F.k2 = X.k2;
  F.emb64 = X.emb2, F.row_bias_epi = X.row_bias_epi, F.col_bias_epi = X.col_bias_epi;
  F.skip_y = X.skip_y, F.skip_x = X.skip_x;
  F.in1 = X.in1, F.in2 = X.in2;
  F.row_bd = X.row_bias_epi >= 0 ? L.row[X.row_bias_epi].bd : nullptr;
  F.col_bd = X.col_bias_epi >= 0 ? L.col[X.col_bias_epi].bd : nullptr;
  factored_fill_kernel<<<num_sms() * 8, 256, 0, st>>>(P, F);
  DM_LAUNCH_OK("factored_fill_kernel");
  int rc;
  // pairs with a full scan on the query side: every float64 query row (+ its bias)
  {
    GemmProblem G;
    G.A.d = X.Phi2, G.A.ld = X.ld2, G.A.off = X.off2, G.A.in = X.bank2 ? X.in2 : nullptr, G.A.trans = 0;
    G.B.d = X.C, G.B.ld = X.k1, G.B.batch_stride = int64_t(X.k1) * X.k2, G.B.rows = X.k2, G.B.trans = 1;
    G.N = X.k1, G.K = X.k2, G.maxM = X.max_n2, G.maxN = X.k1, G.maxK = X.k2, G.n_batch = X.n_pairs;
    G.C = X.emb2, G.ldc = X.k1, G.c_off = X.off2, G.skip = X.skip_y;
    if ((rc = gemm64_launch(G, st))) return rc;
    if (X.col_bias_epi >= 0) {
      bias_rows_flagged_kernel<<<dim3(32, unsigned(X.n_pairs)), 256, 0, st>>>(X.emb2, X.k1, X.off2, X.max_n2, X.k1, X.skip_y,
                                                                             L.col[X.col_bias_epi].bd);
      DM_LAUNCH_OK("bias_rows_flagged_kernel");
    }
  }
  // pairs with a full scan of a result whose database-side bias is factored: every float64 bias -1/2 |C Phi1_j|^2
  if (X.row_bias_epi >= 0) {
    GemmProblem G;
    G.A.d = X.Phi1, G.A.ld = X.ld1, G.A.off = X.off1, G.A.in = X.bank1 ? X.in1 : nullptr, G.A.trans = 0;
    G.B.d = X.C, G.B.ld = X.k1, G.B.batch_stride = int64_t(X.k1) * X.k2, G.B.rows = X.k2, G.B.trans = 0;
    G.N = X.k2, G.K = X.k1, G.maxM = X.max_n1, G.maxN = X.k2, G.maxK = X.k1, G.n_batch = X.n_pairs;
    G.C = X.emb1, G.ldc = X.k2, G.c_off = X.off1, G.skip = X.skip_x;
    if ((rc = gemm64_launch(G, st))) return rc;
    bias_rows_flagged_kernel<<<dim3(32, unsigned(X.n_pairs)), 256, 0, st>>>(X.emb1, X.k2, X.off1, X.max_n1, X.k2, X.skip_x,
                                                                           L.row[X.row_bias_epi].bd);
    DM_LAUNCH_OK("bias_rows_flagged_kernel");
  }
  return DM_OK;
}

struct F2PLayout {
  F2PCtx c;
  void* nn_ws;
  size_t nn_bytes, bytes;
};
F2PLayout f2p_carve(void* ws, int n_pairs, int64_t n1, int64_t n2, int max_n1, int max_n2, int k1, int k2, int flags) {
  Carver c(ws);
  F2PLayout L{};
  const int kp1 = nn_tc_kp(k1), kp2 = nn_tc_kp(k2);
  L.c.kp1 = kp1, L.c.kp2 = kp2;
  L.c.p2h = c.take<uint16_t>(size_t(n2) * kp2);
  L.c.p2m = c.take<uint16_t>(size_t(n2) * kp2);
  L.c.p2l = c.take<uint16_t>(size_t(n2) * kp2);
  L.c.p2norm = c.take<float>(size_t(n2));
  const size_t cb0 = size_t(n_pairs) * EB_ROWS * kp2, cb1 = size_t(n_pairs) * EB_ROWS * kp1;
  L.c.c0h = c.take<uint16_t>(cb0), L.c.c0m = c.take<uint16_t>(cb0), L.c.c0l = c.take<uint16_t>(cb0);
  L.c.c1h = c.take<uint16_t>(cb1), L.c.c1m = c.take<uint16_t>(cb1), L.c.c1l = c.take<uint16_t>(cb1);
  L.c.fro_part = c.take<double>(size_t(n_pairs) * kCsplitChunks);
  L.c.g_scratch = c.take<float>(size_t(n_pairs));
  L.c.emb2 = c.take<double>(size_t(n2) * k1);
  L.c.emb1 = c.take<double>(size_t(n1) * k2);
  L.c.skip_y = c.take<int>(size_t(n_pairs));
  L.c.skip_x = c.take<int>(size_t(n_pairs));
  L.nn_bytes = nn_workspace_bytes(n_pairs, n2, n1, max_n2, max_n1, k1, 2, 2, flags | kFlagSplit3);
  L.nn_ws = c.take<char>(L.nn_bytes);
  L.bytes = c.bytes();
  return L;
}

// ---------------------------------------------------------------- the ladder conversion p2p_21(C) (ZoomOut / ICP)
// queries = Phi2[:, :k2] as they are, database = Phi1[:, :k1] C^T embedded on the tensor cores, one Euclidean row epilogue
struct P21Ctx {
  const double *C, *Phi1;
  int64_t ld1;
  const int64_t* off1;
  int64_t total_n1;
  int max_n1, n_pairs, k1, k2, kp1;
  uint16_t *p1h, *p1m, *p1l;  // three-way split of Phi1 [total_n1, kp1] (resident across rungs)
  float* p1norm;
  uint16_t *c1h, *c1m, *c1l;  // split of C (rows o < k2, contraction k < k1) [n_pairs * 128, kp1]
  double* fro_part;           // [n_pairs, kCsplitChunks]
  double* Ct;                 // [n_pairs, k1, k2] float64 transpose of C
  double* emb1;               // [total_n1, lde] float64 database rows, filled on demand
  int lde;
  int* skip_x;
};

template <int T>
__global__ void __launch_bounds__(256) ladder_fill_kernel(const NNProblem P, const double* __restrict__ Phi1, int64_t ld1,
                                                          const double* __restrict__ Ct, int k1, int k2,
                                                          double* __restrict__ emb1, int lde, double* __restrict__ bd,
                                                          int* __restrict__ skip_x) {
  const int lane = threadIdx.x & 31;
  const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarp = gridDim.x * (blockDim.x >> 5);
  const unsigned n_cand = P.counters[0], n_full = P.counters[3];
  for (unsigned f = warp; f < n_cand + n_full; f += nwarp) {
    const bool full = f >= n_cand;
    const FlagEntry e = full ? P.flags[P.flag_cap - 1 - int64_t(f - n_cand)] : P.flags[f];
    const int p = e.pair;
    if (full) {
      if (lane == 0) skip_x[p] = 0;  // every float64 database row of this pair is needed
      continue;
    }
    const int64_t d0 = P.db_off[p];
    for (int c = 0; c < 2; ++c) {
      const int j = c == 0 ? e.c1 : e.c2;
      // emb1_j[o] = sum_k Phi1[j][k] C[o][k] = sum_k Phi1[j][k] Ct[k][o]
      const double ss = warp_row_times_C<T>(Phi1 + (d0 + j) * ld1, Ct + int64_t(p) * k1 * k2, k2, k1, lane, emb1 + (d0 + j) * lde);
      if (lane == 0) bd[d0 + j] = -0.5 * ss;
    }
  }
}

int p21_prep_x(void* vctx, const NNLayout& L, NNProblem& P, const NNRequest& R, cudaStream_t st) {
  P21Ctx& X = *static_cast<P21Ctx*>(vctx);
  const bool wide = X.kp1 > 2 * EB_BK || X.k2 > EB_ROWS;
  csplit_kernel<<<dim3(unsigned(X.n_pairs), 1, kCsplitChunks), 256, 0, st>>>(
      X.C, X.k1, X.k2, 0, X.kp1, nullptr, nullptr, nullptr, reinterpret_cast<__nv_bfloat16*>(X.c1h),
      reinterpret_cast<__nv_bfloat16*>(X.c1m), reinterpret_cast<__nv_bfloat16*>(X.c1l), X.fro_part, 1, X.Ct,
      wide ? eb_b_rows(X.k2) : EB_ROWS);
  DM_LAUNCH_OK("csplit_kernel");
  EbParams E{};
  E.off = X.off1, E.a_in = X.off1, E.max_rt = (X.max_n1 + EB_ROWS - 1) / EB_ROWS, E.k_out = X.k2, E.kp_out = P.kp, E.n_epi = 1;
  E.hi = reinterpret_cast<__nv_bfloat16*>(L.xh), E.lo = reinterpret_cast<__nv_bfloat16*>(L.xl);
  E.norm = L.norm_db, E.a_norm = X.p1norm, E.fro_part = X.fro_part;
  E.eps_e = eb_eps(X.kp1), E.inv_eps = 1.f / P.eps;
  if (R.n_row != 1 || R.row[0].bias_mode != DM_BIAS_NEG_HALF_SQNORM || R.row[0].scale_mode != DM_SCALE_NONE)
    DM_FAIL(DM_ERR_UNSUPPORTED, "factored database side: one Euclidean row epilogue expected");
  E.epi[0] = EbEpi{1, nullptr, L.row[0].sf, L.row[0].bf, L.row[0].G, L.row[0].Bm, L.row[0].sd, L.row[0].bd};
  DM_CUDA_OK(cudaMemsetAsync(L.row[0].G, 0, sizeof(float) * X.n_pairs, st));
  DM_CUDA_OK(cudaMemsetAsync(L.row[0].Bm, 0, sizeof(float) * X.n_pairs, st));
  const void* a3[3] = {X.p1h, X.p1m, X.p1l};
  const void* b3[3] = {X.c1h, X.c1m, X.c1l};
  return eb_run(a3, X.total_n1, b3, X.kp1, E, X.n_pairs, st);
}

int p21_before_recheck(void* vctx, const NNLayout& L, NNProblem& P, cudaStream_t st) {
  P21Ctx& X = *static_cast<P21Ctx*>(vctx);
  DM_CUDA_OK(cudaMemsetAsync(X.skip_x, 1, sizeof(int) * X.n_pairs, st));
  if (X.k1 <= 128 && X.k2 <= 128)
    ladder_fill_kernel<4><<<num_sms() * 8, 256, 0, st>>>(P, X.Phi1, X.ld1, X.Ct, X.k1, X.k2, X.emb1, X.lde, L.row[0].bd, X.skip_x);
  else
    ladder_fill_kernel<8><<<num_sms() * 8, 256, 0, st>>>(P, X.Phi1, X.ld1, X.Ct, X.k1, X.k2, X.emb1, X.lde, L.row[0].bd, X.skip_x);
  DM_LAUNCH_OK("ladder_fill_kernel");
  GemmProblem G;
  G.A.d = X.Phi1, G.A.ld = X.ld1, G.A.off = X.off1, G.A.trans = 0;
  G.B.d = X.C, G.B.ld = X.k1, G.B.batch_stride = int64_t(X.k1) * X.k2, G.B.rows = X.k2, G.B.trans = 0;
  G.N = X.k2, G.K = X.k1, G.maxM = X.max_n1, G.maxN = X.k2, G.maxK = X.k1, G.n_batch = X.n_pairs;
  G.C = X.emb1, G.ldc = X.lde, G.c_off = X.off1, G.skip = X.skip_x;
  int rc;
  if ((rc = gemm64_launch(G, st))) return rc;
  bias_rows_flagged_kernel<<<dim3(32, unsigned(X.n_pairs)), 256, 0, st>>>(X.emb1, X.lde, X.off1, X.max_n1, X.k2, X.skip_x,
                                                                          L.row[0].bd);
  DM_LAUNCH_OK("bias_rows_flagged_kernel");
  return DM_OK;
}

struct P21Layout {
  P21Ctx c;
  size_t bytes;
};
constexpr int kP21MaxK = 4 * EB_BK;  // widest rung embedded on the tensor cores
P21Layout p21_carve(void* ws, int n_pairs, int64_t total_n1, int k1m, int k2m) {
  Carver c(ws);
  P21Layout L{};
  const int kp = nn_tc_kp(k1m < kP21MaxK ? k1m : kP21MaxK);
  const int rb = (k2m < kP21MaxK ? k2m : kP21MaxK) <= EB_ROWS && kp <= 2 * EB_BK ? EB_ROWS : 2 * EB_ROWS;
  L.c.p1h = c.take<uint16_t>(size_t(total_n1) * kp);
  L.c.p1m = c.take<uint16_t>(size_t(total_n1) * kp);
  L.c.p1l = c.take<uint16_t>(size_t(total_n1) * kp);
  L.c.p1norm = c.take<float>(size_t(total_n1));
  const size_t cb = size_t(n_pairs) * rb * kp;
  L.c.c1h = c.take<uint16_t>(cb), L.c.c1m = c.take<uint16_t>(cb), L.c.c1l = c.take<uint16_t>(cb);
  L.c.fro_part = c.take<double>(size_t(n_pairs) * kCsplitChunks);
  const int kt = k1m < kP21MaxK ? k1m : kP21MaxK, ku = k2m < kP21MaxK ? k2m : kP21MaxK;
  L.c.Ct = c.take<double>(size_t(n_pairs) * kt * ku);
  L.c.skip_x = c.take<int>(size_t(n_pairs));
  L.bytes = c.bytes();
  return L;
}

}  // namespace

bool f2p_factored_applicable(int k1, int k2, int flags) {
  static const bool off = [] { const char* e = getenv("DM_F2P_F64EMB"); return e && e[0] == '1'; }();
  return !off && nn_use_tc(flags) && k1 <= 2 * EB_BK && k2 <= 2 * EB_BK && !(flags & DM_RECHECK_ALL);
}

size_t f2p_factored_workspace_bytes(int n_pairs, int64_t n1, int64_t n2, int max_n1, int max_n2, int k1, int k2, int flags) {
  return f2p_carve(nullptr, n_pairs, n1, n2, max_n1, max_n2, k1, k2, flags).bytes;
}

int f2p_factored_run(const double* C, int k1, int k2, const double* Phi1, int64_t ld1, const int64_t* off1, int64_t total_n1,
                     int max_n1, const double* Phi2, int64_t ld2, const int64_t* off2, int64_t total_n2, int max_n2,
                     const double* area1, int n_pairs, void* p2p_21, void* p2p_12, void* dense_21, void* dense_12, int flags,
                     void* ws, cudaStream_t st, const NNBankSide* bank1, const NNBankSide* bank2) {
  F2PLayout L = f2p_carve(ws, n_pairs, total_n1, total_n2, max_n1, max_n2, k1, k2, flags);
  F2PCtx& X = L.c;
  X.C = C, X.Phi1 = Phi1, X.Phi2 = Phi2, X.ld1 = ld1, X.ld2 = ld2, X.off1 = off1, X.off2 = off2;
  X.bank1 = bank1, X.bank2 = bank2, X.in1 = bank1 ? bank1->in : off1, X.in2 = bank2 ? bank2->in : off2;
  X.total_n1 = total_n1, X.total_n2 = total_n2, X.max_n1 = max_n1, X.max_n2 = max_n2, X.n_pairs = n_pairs;
  X.k1 = k1, X.k2 = k2;
  NNHooks H;
  H.ctx = &X, H.prep_y = f2p_prep_y, H.after_prep = f2p_after_prep, H.before_recheck = f2p_before_recheck;
  NNRequest R{};
  R.Y = nullptr, R.X = nullptr;
  R.Y64 = X.emb2, R.ldY64 = k1, R.X64 = Phi1, R.ldX64 = ld1;
  R.q_off = off2, R.db_off = off1, R.total_q = total_n2, R.total_db = total_n1;
  R.max_q = max_n2, R.max_db = max_n1, R.n_pairs = n_pairs, R.d = k1, R.d_fast = 0;
  R.n_row = 0, R.n_col = 0;
  X.row_bias_epi = X.col_bias_epi = -1;
  if (p2p_21) {
    X.row_bias_epi = R.n_row;  // bias filled by f2p_after_prep (fp32) and factored_fill_kernel (float64, on demand)
    // declared as an array bias so that the epilogue is never treated as a plain dot product; the array handed to
    // nn_prep_side is only a valid placeholder (its values are overwritten / never read, see the hooks)
    R.row[R.n_row++] = dm_nn_epi{DM_SCALE_NONE, DM_BIAS_ARRAY, nullptr, X.emb1, p2p_21};
  }
  if (dense_21) R.row[R.n_row++] = dm_nn_epi{DM_SCALE_ARRAY, DM_BIAS_NONE, area1, nullptr, dense_21};
  if (p2p_12) {
    X.col_bias_epi = R.n_col;
    R.col[R.n_col++] = dm_nn_epi{DM_SCALE_NONE, DM_BIAS_NEG_HALF_SQNORM, nullptr, nullptr, p2p_12};
  }
  if (dense_12) R.col[R.n_col++] = dm_nn_epi{DM_SCALE_NONE, DM_BIAS_NONE, nullptr, nullptr, dense_12};
  R.flags = flags | kFlagSplit3;
  R.hooks = &H;
  R.bank_db = bank1;  // (area1, the scale of dense_21, is then indexed by bank rows)
  return nn_run(R, L.nn_ws, L.nn_bytes, st);
}

}  // namespace dm

namespace dm {

bool p2p21_factored_applicable(int k1, int k2, int flags) {
  static const bool off = [] { const char* e = getenv("DM_P21_F64EMB"); return e && e[0] == '1'; }();
  static const bool narrow = [] { const char* e = getenv("DM_P21_NARROW"); return e && e[0] == '1'; }();
  const int lim = narrow ? 2 * EB_BK : kP21MaxK;
  return !off && nn_use_tc(flags) && k1 <= lim && k2 <= lim && !(flags & (DM_RECHECK_ALL | DM_SKIP_PREP));
}

size_t p2p21_factored_scratch_bytes(int n_pairs, int64_t total_n1, int k1m, int k2m) {
  return p21_carve(nullptr, n_pairs, total_n1, k1m, k2m).bytes;
}

int p2p21_factored_run(const double* C, int k1, int k2, const double* Phi1, int64_t ld1, const int64_t* off1, int64_t total_n1,
                       int max_n1, const double* Phi2, int64_t ld2, const int64_t* off2, int64_t total_n2, int max_n2,
                       int n_pairs, void* p2p_out, int flags, void* scratch, int scratch_k1m, int scratch_k2m, double* emb1,
                       int lde, void* nn_ws, size_t nn_ws_bytes, cudaStream_t st, int* x_kp_state, int* y_kp_state) {
  // the carve uses the widest split the scratch was sized for; this rung uses the first kp1 columns of each row, so the
  // split of Phi1 is stored with the rung's own pitch kp1 and remade when the pitch changes (64 -> 128)
  // carved with the widths the scratch was sized for (p2p21_factored_scratch_bytes): the layout is the same at every rung
  if (k1 > scratch_k1m || k2 > scratch_k2m) DM_FAIL(DM_ERR_WORKSPACE, "factored conversion: scratch sized for narrower maps");
  P21Layout L = p21_carve(scratch, n_pairs, total_n1, scratch_k1m, scratch_k2m);
  P21Ctx& X = L.c;
  X.C = C, X.Phi1 = Phi1, X.ld1 = ld1, X.off1 = off1, X.total_n1 = total_n1, X.max_n1 = max_n1, X.n_pairs = n_pairs;
  X.k1 = k1, X.k2 = k2, X.kp1 = nn_tc_kp(k1), X.emb1 = emb1, X.lde = lde;
  int rc;
  if (!x_kp_state || *x_kp_state != X.kp1) {
    // all the columns of the padded width (those beyond k1 meet the zero padding of the split of C)
    const int d_all = int(ld1 < X.kp1 ? ld1 : X.kp1);
    if ((rc = nn_prep_side(Phi1, 1, ld1, off1, n_pairs, total_n1, d_all, X.p1norm, nullptr, 0, X.p1h, X.p1m, X.p1l, X.kp1, st)))
      return rc;
    if (x_kp_state) *x_kp_state = X.kp1;
  }
  NNHooks H;
  H.ctx = &X, H.prep_x = p21_prep_x, H.before_recheck = p21_before_recheck;
  NNRequest R{};
  R.Y64 = Phi2, R.ldY64 = ld2, R.X64 = emb1, R.ldX64 = lde;
  R.q_off = off2, R.db_off = off1, R.total_q = total_n2, R.total_db = total_n1;
  R.max_q = max_n2, R.max_db = max_n1, R.n_pairs = n_pairs, R.d = k2, R.d_fast = 0;
  R.n_row = 1, R.n_col = 0;
  R.row[0] = dm_nn_epi{DM_SCALE_NONE, DM_BIAS_NEG_HALF_SQNORM, nullptr, nullptr, p2p_out};
  R.flags = flags;
  R.hooks = &H;
  if (y_kp_state) {
    const int kp = nn_tc_kp(k2);
    R.y_prep_d = int(ld2 < kp ? ld2 : kp);
    R.skip_prep_y = (*y_kp_state == kp);
    *y_kp_state = kp;
  }
  return nn_run(R, nn_ws, nn_ws_bytes, st);
}

}  // namespace dm
