// Dual-accumulator tcgen05 engine for the FM -> p2p pass: FOUR argmax reductions (two over the database, two over the
// queries) of one score matrix S = Y X^T with a SHORT contraction (k <= 128: the spectral embeddings).
//
// What bounded the generic kernel (nn_tc.cu) on this shape (ncu, round 1/2): not the tensor pipe (15 %) but the
// epilogue -- the column reductions transpose every 32 x 32 block through shared memory, one warp group carried both
// of them and ran at 0.25 IPC while the two row groups spun on the accumulator barrier; and with K = 128 every
// 128 x 256 tile re-fetched its Y operand, so the L2 -> SM operand traffic (3.2 GB per 128 pairs) was as long as the math.
//
// Here each CTA keeps its 128 query rows of Y RESIDENT in shared memory for the whole sweep over the database, and the
// tensor core computes BOTH S (128 x 128: rows = queries) and S^T (rows = database rows) of every tile from the same
// staged operands -- a second, swapped tcgen05.mma costs no extra operand traffic and the tensor pipe has the headroom.
// Both accumulators live in tensor memory (2 x 128 columns each, double-buffered: all 512 columns), and ALL FOUR
// reductions become row scans: a thread owns one accumulator row (TMEM lane) and walks its 128 columns -- no
// transposition, no shared-memory patch, four equally loaded groups of four warps.
//   group 0 / 1: rows of S   -> the two "row" epilogues  (argmax over database rows j; state carried over all tiles)
//   group 2 / 3: rows of S^T -> the two "column" epilogues (argmax over this CTA's 128 query rows; one partial per
//                tile and row tile, merged by col_finalize_kernel like the generic engine's)
// A scan keeps the running top-2 of each 32-column chunk on PACKED keys (the 5 low mantissa bits carry the column, so
// best + runner-up with their indices cost 2.5 FMNMX per score) and merges it into the thread's (best, runner-up, third)
// state once per chunk.  The 2^-18 relative truncation is part of the re-evaluation threshold (NNProblem::row_trunc /
// col_trunc); exactness versus the float64 reference comes from the same near-tie re-evaluation as everywhere else.
//
// Replaces, for the four-output call: FM_to_p2p (densematcher/pyFM/spectral/convert.py:96-147) + the dense argmax
// override (densematcher/functional_map.py:49-50).
#include "dm_internal.cuh"
#include "tc_ptx.cuh"
#include "tc_scan.cuh"

namespace dm {
namespace {

using namespace tc;

constexpr int T2_ROWS = 128;   // query rows per CTA (UMMA M of S, UMMA N of S^T)
constexpr int T2_TN = 128;     // database rows per tile (UMMA N of S, UMMA M of S^T)
constexpr int T2_BK = 64;      // K elements per chunk (one 128-byte swizzle row of bf16)
constexpr int T2_UK = 16;      // UMMA K
constexpr int T2_NST = 2;      // X stages (each holds the whole contraction of one tile)
constexpr int T2_GROUPS = 4;
constexpr int T2_THREADS = 32 * (2 + 4 * T2_GROUPS);  // 576
constexpr uint32_t T2_TILE_BYTES = T2_ROWS * T2_BK * 2;  // one [128 x 64] bf16 box: 16 KB

__host__ __device__ constexpr uint32_t t2_operand_bytes(int kc) { return uint32_t(kc) * 2 * T2_TILE_BYTES; }  // hi + lo
__host__ __device__ constexpr uint32_t t2_off_x(int kc) { return t2_operand_bytes(kc); }
__host__ __device__ constexpr uint32_t t2_off_rowsb(int kc) { return t2_off_x(kc) + T2_NST * t2_operand_bytes(kc); }
constexpr uint32_t T2_ROWSB_BYTES = 2 /*parity*/ * 2 /*epi*/ * 2 /*scale, bias*/ * T2_TN * 4;
constexpr uint32_t T2_CSB_BYTES = 2 /*epi*/ * 2 * T2_ROWS * 4;
__host__ __device__ constexpr uint32_t t2_off_csb(int kc) { return t2_off_rowsb(kc) + T2_ROWSB_BYTES; }
__host__ __device__ constexpr uint32_t t2_off_bar(int kc) { return t2_off_csb(kc) + T2_CSB_BYTES; }
__host__ __device__ constexpr uint32_t t2_smem_bytes(int kc) { return t2_off_bar(kc) + 128 + 1024; }
static_assert(t2_smem_bytes(2) <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ uint64_t t2_desc_sw128(uint32_t saddr) {  // K-major, 128-byte swizzle, 8-row atoms 1024 B apart
  uint64_t d = 0;
  d |= uint64_t((saddr >> 4) & 0x3FFF);
  d |= uint64_t(1) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}
// kind::f16, A = B = bf16 (K-major), D = fp32, M = 128, N = 128
constexpr uint32_t kT2Idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(T2_TN >> 3) << 17) | (uint32_t(T2_ROWS >> 4) << 24);

__device__ __forceinline__ void t2_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct T2Maps {
  CUtensorMap yh, yl, xh, xl;
};

template <int KC>  // 64-wide K chunks: kp = 64 KC
__global__ void __launch_bounds__(T2_THREADS, 1)
    f2p_tc_kernel(const __grid_constant__ T2Maps maps, const NNProblem P, const uint32_t keymask) {
  constexpr uint32_t OPB = t2_operand_bytes(KC);
  const int p = blockIdx.x / P.max_rt, rt = blockIdx.x % P.max_rt;
  const int64_t q0 = P.q_off[p];
  const int nq = int(P.q_off[p + 1] - q0);
  const int row0 = rt * T2_ROWS;
  if (row0 >= nq) return;
  const int64_t d0 = P.db_off[p];
  const int nd = int(P.db_off[p + 1] - d0);
  const int n_ct = (nd + T2_TN - 1) / T2_TN;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar_base = sbase + t2_off_bar(KC);
  const uint32_t bar_y = bar_base;                    // Y resident
  const uint32_t bar_xfull = bar_base + 8;            // [NST]
  const uint32_t bar_xempty = bar_xfull + 8 * T2_NST; // [NST]
  const uint32_t bar_tfull = bar_xempty + 8 * T2_NST; // [2]
  const uint32_t bar_tempty = bar_tfull + 16;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + t2_off_bar(KC) + 8 * (1 + 2 * T2_NST + 4));
  float* rowsb = reinterpret_cast<float*>(sgen + t2_off_rowsb(KC));  // [parity][epi][scale | bias][TN]
  float* csb = reinterpret_cast<float*>(sgen + t2_off_csb(KC));      // [epi][scale | bias][ROWS]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_y, 1);
    for (int s = 0; s < T2_NST; ++s) {
      mbar_init(bar_xfull + 8 * s, 1);
      mbar_init(bar_xempty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 4 * T2_GROUPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.yh) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.yl) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.xh) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.xl) : "memory");
      const int yrow = int(P.q_in[p] + row0);  // operand rows (a side may live in a mesh bank), see NNProblem::q_in
      const int64_t xrow0 = P.db_in[p];
      mbar_expect_tx(bar_y, OPB);
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        tma_load_2d(sbase + (2 * kc + 0) * T2_TILE_BYTES, &maps.yh, kc * T2_BK, yrow, bar_y);
        tma_load_2d(sbase + (2 * kc + 1) * T2_TILE_BYTES, &maps.yl, kc * T2_BK, yrow, bar_y);
      }
      for (int ct = 0; ct < n_ct; ++ct) {
        const int stage = ct % T2_NST;
        const uint32_t phase = (ct / T2_NST) & 1;
        mbar_wait_backoff(bar_xempty + 8 * stage, phase ^ 1);
        const uint32_t sb = sbase + t2_off_x(KC) + stage * OPB, fb = bar_xfull + 8 * stage;
        const int xrow = int(xrow0 + int64_t(ct) * T2_TN);
        mbar_expect_tx(fb, OPB);
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) {
          tma_load_2d(sb + (2 * kc + 0) * T2_TILE_BYTES, &maps.xh, kc * T2_BK, xrow, fb);
          tma_load_2d(sb + (2 * kc + 1) * T2_TILE_BYTES, &maps.xl, kc * T2_BK, xrow, fb);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: S = Y X^T and S^T = X Y^T of every tile, three split products each
    if (lane == 0) {
      mbar_wait_backoff(bar_y, 0);
      tc_fence_after();
      for (int ct = 0; ct < n_ct; ++ct) {
        const int acc = ct & 1, stage = ct % T2_NST;
        mbar_wait_backoff(bar_tempty + 8 * acc, ((ct >> 1) & 1) ^ 1);
        mbar_wait_backoff(bar_xfull + 8 * stage, (ct / T2_NST) & 1);
        tc_fence_after();
        const uint32_t tS = tmem_base + acc * T2_TN, tT = tmem_base + 2 * T2_TN + acc * T2_ROWS;
        const uint32_t xb = sbase + t2_off_x(KC) + stage * OPB;
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) {
          const uint64_t dyh = t2_desc_sw128(sbase + (2 * kc + 0) * T2_TILE_BYTES);
          const uint64_t dyl = t2_desc_sw128(sbase + (2 * kc + 1) * T2_TILE_BYTES);
          const uint64_t dxh = t2_desc_sw128(xb + (2 * kc + 0) * T2_TILE_BYTES);
          const uint64_t dxl = t2_desc_sw128(xb + (2 * kc + 1) * T2_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < T2_BK / T2_UK; ++k) {
            const uint64_t ko = uint64_t((k * T2_UK * 2) >> 4);
            const uint32_t first = (kc | k) != 0;
            tc_mma_bf16(tS, dyh + ko, dxh + ko, kT2Idesc, first);
            tc_mma_bf16(tS, dyh + ko, dxl + ko, kT2Idesc, 1);
            tc_mma_bf16(tS, dyl + ko, dxh + ko, kT2Idesc, 1);
            tc_mma_bf16(tT, dxh + ko, dyh + ko, kT2Idesc, first);
            tc_mma_bf16(tT, dxl + ko, dyh + ko, kT2Idesc, 1);
            tc_mma_bf16(tT, dxh + ko, dyl + ko, kT2Idesc, 1);
          }
        }
        tc_commit(bar_xempty + 8 * stage);  // the X stage is reusable once these MMAs have read it
        tc_commit(bar_tfull + 8 * acc);     // both accumulators of the tile are complete
      }
    }
  } else {
    // ===================== four epilogue groups of four warps; warp w reads TMEM lanes 32 (w % 4) ..
    const int q = warp & 3;
    const int group = (warp - 2) >> 2;           // 0, 1: rows of S; 2, 3: rows of S^T
    const int gt = (threadIdx.x - 64) & 127;     // thread inside its group
    const int trow = 32 * q + lane;              // accumulator row (TMEM lane)
    const bool on_s = group < 2;
    const int e = group & 1;                     // epilogue number on its side
    const EpiDev& E = on_s ? P.row[e] : P.col[e];
    const bool ident = E.identity != 0;

    if (!on_s) {  // scale / bias over this CTA's query rows (the columns of S^T): once per CTA
      const int i = row0 + gt;
      const bool ok = i < nq;
      csb[(e * 2 + 0) * T2_ROWS + gt] = ok ? __ldg(E.sf + q0 + i) : 0.f;
      csb[(e * 2 + 1) * T2_ROWS + gt] = ok ? __ldg(E.bf + q0 + i) : kMaskedScore;
      t2_bar_sync(1 + group, 128);
    }
    const bool full_rows = row0 + T2_ROWS <= nq;
    Top3 st = top3_init();
    // the re-evaluation window of this thread's result without its score-dependent part (emit_result)
    const float gE = E.G[p], bE = 9.6e-7f * E.Bm[p];
    float thr_base = on_s ? 2.f * P.eps * ((row0 + trow < nq) ? P.norm_q[q0 + row0 + trow] : 0.f) * gE + bE : 0.f;

    for (int ct = 0; ct < n_ct; ++ct) {
      const int acc = ct & 1, col0 = ct * T2_TN;
      const bool full_cols = col0 + T2_TN <= nd;
      const float* sc;
      const float* bi;
      if (on_s) {
        // scale / bias over the tile's database rows (the columns of S), double-buffered by the tile parity
        float* rb = rowsb + ((acc * 2 + e) * 2) * T2_TN;
        const int j = col0 + gt;
        const bool ok = j < nd;
        rb[gt] = ok ? __ldg(E.sf + d0 + j) : 0.f;
        rb[T2_TN + gt] = ok ? __ldg(E.bf + d0 + j) : kMaskedScore;
        t2_bar_sync(1 + group, 128);
        sc = rb, bi = rb + T2_TN;
      } else {
        sc = csb + (e * 2 + 0) * T2_ROWS, bi = csb + (e * 2 + 1) * T2_ROWS;
        st = top3_init();
        const int j = col0 + trow;
        thr_base = 2.f * P.eps * (j < nd ? P.norm_db[d0 + j] : 0.f) * gE + bE;
      }
      const bool plain = ident && (on_s ? full_cols : full_rows);  // no scale / bias and nothing to mask

      mbar_wait(bar_tfull + 8 * acc, (ct >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (on_s ? acc * T2_TN : 2 * T2_TN + acc * T2_ROWS) + (uint32_t(32 * q) << 16);
      const int base0 = on_s ? col0 : row0;  // index of column 0 of the accumulator on the reduced-over side
      const int n_valid = on_s ? min(T2_TN, nd - col0) : min(T2_ROWS, nq - row0);
      const int n_ch = (n_valid + T2_CH - 1) / T2_CH;
      for (int ch = 0; ch < n_ch; ++ch) {
        float v[32];
        tmem_ld32(taddr + ch * T2_CH, v);
        float k1, k2;
        if (plain) {
          t2_chunk_top2<true>(v, nullptr, nullptr, keymask, k1, k2);
          t2_merge_chunk<true>(st, k1, k2, base0 + ch * T2_CH, v, nullptr, nullptr, keymask, thr_base);
        } else {
          t2_chunk_top2<false>(v, sc + ch * T2_CH, bi + ch * T2_CH, keymask, k1, k2);
          t2_merge_chunk<false>(st, k1, k2, base0 + ch * T2_CH, v, sc + ch * T2_CH, bi + ch * T2_CH, keymask, thr_base);
        }
      }
      tc_fence_before();
      if (!on_s) {
        // partial of (row tile rt, database row j) for col_finalize_kernel
        const int j = col0 + trow;
        if (j < nd) P.col_partial[((int64_t(e) * P.n_pairs + p) * P.max_rt + rt) * P.max_db + j] = st;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
    }
    if (on_s) {
      const int i = row0 + trow;
      if (i < nq) emit_result(P, E, false, e, p, q0 + i, i, P.norm_q[q0 + i], st);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int KC>
int t2_launch(const T2Maps& maps, const NNProblem& P, cudaStream_t st) {
  static OncePerDevice attr_once;
  if (attr_once.first())
    DM_CUDA_OK(cudaFuncSetAttribute(f2p_tc_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(t2_smem_bytes(KC))));
  const int64_t nblk = int64_t(P.n_pairs) * P.max_rt;
  if (nblk > 0x7fffffffLL) DM_FAIL(DM_ERR_BADARG, "too many row tiles (%lld)", (long long)nblk);
  f2p_tc_kernel<KC><<<unsigned(nblk), T2_THREADS, t2_smem_bytes(KC), st>>>(maps, P, ~uint32_t(T2_CH - 1));
  DM_LAUNCH_OK("f2p_tc_kernel");
  return DM_OK;
}

}  // namespace

// the dual-accumulator engine serves the four-output pass with a short contraction
bool nn_tc2_applicable(int n_row, int n_col, int kp) {
  static const bool off = [] { const char* e = getenv("DM_F2P_OLD"); return e && e[0] == '1'; }();
  return !off && n_row == 2 && n_col == 2 && kp <= 2 * T2_BK;
}

int nn_tc2_launch(const NNProblem& P, const void* Yh, const void* Yl, const void* Xh, const void* Xl, cudaStream_t st) {
  if (P.n_pairs <= 0 || P.total_q <= 0) return DM_OK;
  if (P.rows_q > 0x7fffffffLL || P.rows_db > 0x7fffffffLL) DM_FAIL(DM_ERR_BADARG, "too many rows for TMA coordinates");
  T2Maps maps;
  int rc;
  if ((rc = tc_make_map_bf16(&maps.yh, Yh, P.rows_q, P.kp, T2_ROWS))) return rc;
  if ((rc = tc_make_map_bf16(&maps.yl, Yl, P.rows_q, P.kp, T2_ROWS))) return rc;
  if ((rc = tc_make_map_bf16(&maps.xh, Xh, P.rows_db, P.kp, T2_TN))) return rc;
  if ((rc = tc_make_map_bf16(&maps.xl, Xl, P.rows_db, P.kp, T2_TN))) return rc;
  return P.kp <= T2_BK ? t2_launch<1>(maps, P, st) : t2_launch<2>(maps, P, st);
}

}  // namespace dm
