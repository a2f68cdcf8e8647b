// tcgen05 engine of the fused similarity + argmax pass (sm_100a).
//
// S = Y X^T is never written.  Each CTA owns 128 query rows of one mesh pair and sweeps the pair's
// database in tiles of 256 columns.  fp32-grade scores come from THREE bf16 tensor-core passes over a
// split representation v = hi + lo (hi = bf16(v), lo = bf16(v - hi)):
//     S ~= Yhi Xhi^T + Yhi Xlo^T + Ylo Xhi^T          (|error| <= 3 * 2^-18 |y||x| + accumulation)
// accumulated in fp32 in tensor memory.  Operand tiles are staged by TMA (128-byte swizzle) through a
// 2-stage mbarrier ring; one thread issues tcgen05.mma; the accumulator is double-buffered in TMEM
// (2 x 256 columns) so that the epilogue of tile t overlaps the MMAs of tile t+1.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue.  An epilogue thread owns one accumulator row (TMEM lane): the row epilogues are a
// purely sequential top-3 scan over the columns (no shuffles); the column epilogues transpose each
// 32x32 block through a padded shared-memory patch, scan 32 rows per lane, merge the four warps in
// shared memory and write one partial per (row tile, column) for nn_col_finalize.
//
// Replaces the kd-tree search of knn_query (densematcher/pyFM/spectral/nn_utils.py:28-30) and the dense
// argmax of functional_map.py:49-50; exactness versus the float64 reference is restored by the
// near-tie re-evaluation in nn_common.cu.
#include "dm_internal.cuh"
#include "tc_ptx.cuh"
#include "tc_scan.cuh"

namespace dm {
namespace tc {
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}
}  // namespace tc

namespace {

constexpr int TM_ROWS = 128;  // query rows per CTA (UMMA M)
constexpr int TN = 256;       // database columns per accumulator tile (UMMA N)
constexpr int TBK = 64;       // K elements per pipeline stage (one 128-byte swizzle row of bf16)
// Pipeline depth and X tile per CTA.  In PAIR mode two CTAs (a cluster of 2 = one TPC) compute a 256 x 256 tile with
// tcgen05.mma.cta_group::2: each CTA stages its own 128 rows of Y and only HALF of the X tile, so a stage is 64 KB
// instead of 96 KB -- the L2 -> SM traffic per flop (the measured limiter of the single-CTA kernel) drops by a third,
// the shared-memory operand reads per SM drop too, and a third stage fits.
__host__ __device__ constexpr int n_stages(bool pair) { return pair ? 3 : 2; }
constexpr int UMMA_K = 16;
constexpr int kGroupWarps = 4;              // one warp per TMEM lane quarter
// Two epilogue groups, or three when there are two row epilogues AND column epilogues (see the kernel).  Row epilogues
// only: FOUR groups share the chunks of every tile -- the sequential top-3 scan is latency-bound, not issue-bound (two
// warps per scheduler left the ZoomOut conversion at 520 us per 128 pairs where its MMAs need 160), so twice the warps
// is the cheapest way to hide it
__host__ __device__ constexpr bool three_groups(int nr, int nc, bool debug) { return nr == 2 && nc > 0 && !debug; }
__host__ __device__ constexpr int n_groups(int nr, int nc, bool debug) {
  return three_groups(nr, nc, debug) ? 3 : ((nc == 0 && nr > 0 && !debug) ? 4 : 2);
}
__host__ __device__ constexpr int epi_warps(int nr, int nc, bool debug) { return n_groups(nr, nc, debug) * kGroupWarps; }
__host__ __device__ constexpr int n_threads(int nr, int nc, bool debug) { return 32 * (2 + epi_warps(nr, nc, debug)); }
constexpr int CCH = 32;  // columns per epilogue chunk (one tcgen05.ld.32x32b.x32)

constexpr uint32_t SZ_Y = TM_ROWS * TBK * 2;  // 16 KB per half
__host__ __device__ constexpr uint32_t sz_x(bool pair) { return (pair ? TN / 2 : TN) * TBK * 2; }  // 32 KB / 16 KB per half
__host__ __device__ constexpr uint32_t stage_bytes(bool pair) { return 2 * SZ_Y + 2 * sz_x(pair); }
constexpr uint32_t OFF_PATCH = n_stages(false) * stage_bytes(false);
static_assert(n_stages(true) * stage_bytes(true) == OFF_PATCH, "both modes use the same staging area");
constexpr int PATCH_LD = 36;  // row pitch (words) of the transposition patch: 16-byte stores, conflict-free both ways
constexpr uint32_t PATCH_BYTES = 32 * PATCH_LD * 4;
constexpr uint32_t OFF_COLRED = OFF_PATCH + kGroupWarps * PATCH_BYTES;          // [2][kMaxEpi][4 warps][32] Top3
constexpr uint32_t COLRED_BYTES = 2 * kMaxEpi * kGroupWarps * CCH * sizeof(Top3);
constexpr uint32_t OFF_ROWSB = OFF_COLRED + COLRED_BYTES;                       // [kMaxEpi][2][TN] float
constexpr uint32_t ROWSB_BYTES = kMaxEpi * 2 * TN * 4;
constexpr uint32_t OFF_BAR = OFF_ROWSB + ROWSB_BYTES;                           // mbarriers + tmem pointer
constexpr uint32_t SMEM_BYTES = OFF_BAR + 128 + 1024;                           // + alignment slack
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

using namespace tc;
// named barriers of the epilogue: 1 = row group (or both groups when they share the row work), 2 = column group,
// 3 = second row group
__device__ __forceinline__ void bar_sync_n(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// Four candidates at a time into TWO independent running top-3 chains (a: first two, b: last two) so that consecutive
// updates do not serialise on one dependency chain.  Candidates below the running third-best cannot change a top-3,
// which is the common case after the first few hundred columns: the update is skipped when no lane of the warp needs
// it (warp-uniform branch on a vote, no divergence).
__device__ __forceinline__ void top3_offer4(Top3& a, Top3& b, float w0, float w1, float w2, float w3, int idx0) {
  const bool need = fmaxf(w0, w1) > a.m3 || fmaxf(w2, w3) > b.m3;
  if (__any_sync(0xffffffffu, need)) {
    top3_push(a, w0, idx0);
    top3_push(b, w2, idx0 + 2);
    top3_push(a, w1, idx0 + 1);
    top3_push(b, w3, idx0 + 3);
  }
}

// Two row epilogues at once: one vote for both, four independent chains in the update
__device__ __forceinline__ void top3_offer4x2(Top3& a0, Top3& b0, Top3& a1, Top3& b1, const float (&w)[4], const float (&u)[4],
                                              int idx0) {
  const bool need = fmaxf(w[0], w[1]) > a0.m3 || fmaxf(w[2], w[3]) > b0.m3 || fmaxf(u[0], u[1]) > a1.m3 ||
                    fmaxf(u[2], u[3]) > b1.m3;
  if (__any_sync(0xffffffffu, need)) {
    top3_push(a0, w[0], idx0);
    top3_push(a1, u[0], idx0);
    top3_push(b0, w[2], idx0 + 2);
    top3_push(b1, u[2], idx0 + 2);
    top3_push(a0, w[1], idx0 + 1);
    top3_push(a1, u[1], idx0 + 1);
    top3_push(b0, w[3], idx0 + 3);
    top3_push(b1, u[3], idx0 + 3);
  }
}

// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row atoms 1024 B apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= uint64_t((saddr >> 4) & 0x3FFF);
  d |= uint64_t(1) << 16;            // leading byte offset (unused for swizzled K-major), 16 B
  d |= uint64_t(1024 >> 4) << 32;    // stride byte offset between 8-row groups
  d |= uint64_t(1) << 46;            // descriptor version
  d |= uint64_t(2) << 61;            // SWIZZLE_128B
  return d;
}
// kind::f16, A = B = bf16 (K-major), D = fp32, M = 128, N = TN
__host__ __device__ constexpr uint32_t idesc(bool pair) {  // PAIR: M = 256 over the two CTAs
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(TN >> 3) << 17) | (uint32_t((pair ? 2 * TM_ROWS : TM_ROWS) >> 4) << 24);
}

struct TcMaps {
  CUtensorMap yh, yl, xh, xl;
};

struct DebugOut {
  float* S;
  int64_t ldS;
};

// ------------------------------------------------------------------ the kernel
template <int NR, int NC, bool DEBUG, bool PAIR>
__global__ void __launch_bounds__(n_threads(NR, NC, DEBUG), 1)
    nn_tc_kernel(const __grid_constant__ TcMaps maps, const NNProblem P, const DebugOut dbg, const uint32_t keymask) {
  constexpr int STAGES = n_stages(PAIR);
  constexpr uint32_t SZ_X = sz_x(PAIR), STAGE_BYTES = stage_bytes(PAIR);
  constexpr uint32_t kIdesc = idesc(PAIR);
  // PAIR: the grid has an even number of row tiles per pair of meshes; CTAs 2m / 2m+1 of a cluster own row tiles 2m / 2m+1
  const int mrt = PAIR ? (P.max_rt + 1) & ~1 : P.max_rt;
  const int p = blockIdx.x / mrt, rt = blockIdx.x % mrt;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  const int64_t q0 = P.q_off[p];
  const int nq = int(P.q_off[p + 1] - q0);
  const int row0 = rt * TM_ROWS;
  if ((PAIR ? (rt & ~1) * TM_ROWS : row0) >= nq) return;  // uniform over the cluster
  const bool active = row0 < nq;  // PAIR: a CTA past the last row still stages its half of X and takes part in the MMA
  const int64_t d0 = P.db_off[p];
  const int nd = int(P.db_off[p + 1] - d0);
  const int n_ct = (nd + TN - 1) / TN;
  const int n_kc = P.kp / TBK;
  const int n_k16 = max(1, min(P.kp / UMMA_K, (P.d + UMMA_K - 1) / UMMA_K));
  const int last_ksteps = n_k16 - (n_kc - 1) * (TBK / UMMA_K);  // K steps of the last chunk that meet data (1..4)

  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar_full = sbase + OFF_BAR;               // [STAGES]
  const uint32_t bar_empty = bar_full + 8 * STAGES;        // [STAGES]
  const uint32_t bar_tfull = bar_empty + 8 * STAGES;       // [2]
  const uint32_t bar_tempty = bar_tfull + 16;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + OFF_BAR + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, (PAIR ? 2 : 1) * epi_warps(NR, NC, DEBUG));  // PAIR: both CTAs' warps, on the leader
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {  // the same warp of both CTAs
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers are initialised and its tensor memory allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.yh) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.yl) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.xh) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.xl) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      // operand rows: the pair's meshes may live in a bank (q_in / db_in), results and per-row arrays are batch-packed
      const int yrow = int(P.q_in[p] + row0);
      const int64_t xrow0 = P.db_in[p];
      for (int ct = 0; ct < n_ct; ++ct) {
        const int xrow = int(xrow0 + int64_t(ct) * TN) + (PAIR ? int(rank) * (TN / 2) : 0);
        for (int kc = 0; kc < n_kc; ++kc) {
          mbar_wait_backoff(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sb = sbase + stage * STAGE_BYTES, fb = bar_full + 8 * stage;
          if (PAIR) {
            // both CTAs' bytes are counted on the LEADER's barrier (its MMA thread is the only consumer)
            const uint32_t fbl = mapa_shared(fb, 0);
            if (rank == 0) mbar_expect_tx(fb, 2 * STAGE_BYTES);
            tma_load_2d_pair(sb, &maps.yh, kc * TBK, yrow, fbl);
            tma_load_2d_pair(sb + SZ_Y, &maps.yl, kc * TBK, yrow, fbl);
            tma_load_2d_pair(sb + 2 * SZ_Y, &maps.xh, kc * TBK, xrow, fbl);
            tma_load_2d_pair(sb + 2 * SZ_Y + SZ_X, &maps.xl, kc * TBK, xrow, fbl);
          } else {
            mbar_expect_tx(fb, STAGE_BYTES);
            tma_load_2d(sb, &maps.yh, kc * TBK, yrow, fb);
            tma_load_2d(sb + SZ_Y, &maps.yl, kc * TBK, yrow, fb);
            tma_load_2d(sb + 2 * SZ_Y, &maps.xh, kc * TBK, xrow, fb);
            tma_load_2d(sb + 2 * SZ_Y + SZ_X, &maps.xl, kc * TBK, xrow, fb);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread; PAIR: of the leader CTA, for both SMs)
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ct = 0; ct < n_ct; ++ct) {
        const int acc = ct & 1;
        mbar_wait_backoff(bar_tempty + 8 * acc, ((ct >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + acc * TN;
        for (int kc = 0; kc < n_kc; ++kc) {
          mbar_wait_backoff(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sb = sbase + stage * STAGE_BYTES;
          const uint64_t dyh = umma_desc_sw128(sb), dyl = umma_desc_sw128(sb + SZ_Y);
          const uint64_t dxh = umma_desc_sw128(sb + 2 * SZ_Y), dxl = umma_desc_sw128(sb + 2 * SZ_Y + SZ_X);
          // the database side is zero beyond column d: the K steps of the last chunk that only meet padding are skipped
          // (k = 150 pads to 192 columns but needs 160; the upper ZoomOut rungs are bound by these MMAs)
          auto issue_kstep = [&](int k) {
            const uint64_t ko = uint64_t((k * UMMA_K * 2) >> 4);  // 32 bytes per K step inside the swizzle row
            if (PAIR) {
              tc_mma_bf16_pair(tacc, dyh + ko, dxh + ko, kIdesc, (kc | k) != 0);
              tc_mma_bf16_pair(tacc, dyh + ko, dxl + ko, kIdesc, 1);
              tc_mma_bf16_pair(tacc, dyl + ko, dxh + ko, kIdesc, 1);
            } else {
              tc_mma_bf16(tacc, dyh + ko, dxh + ko, kIdesc, (kc | k) != 0);
              tc_mma_bf16(tacc, dyh + ko, dxl + ko, kIdesc, 1);
              tc_mma_bf16(tacc, dyl + ko, dxh + ko, kIdesc, 1);
            }
          };
          if (kc + 1 < n_kc || last_ksteps == TBK / UMMA_K) {  // full chunk: branch-free issue (the issuing thread is the
#pragma unroll                                                 // critical path of MMA-bound launches)
            for (int k = 0; k < TBK / UMMA_K; ++k) issue_kstep(k);
          } else {
            for (int k = 0; k < last_ksteps; ++k) issue_kstep(k);
          }
          // smem slot reusable once these MMAs have read it (PAIR: in both CTAs)
          if (PAIR) tc_commit_pair(bar_empty + 8 * stage); else tc_commit(bar_empty + 8 * stage);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        // accumulator complete (PAIR: each CTA's epilogue reads its own 128 rows from its own tensor memory)
        if (PAIR) tc_commit_pair(bar_tfull + 8 * acc); else tc_commit(bar_tfull + 8 * acc);
      }
    }
  } else {
    // ===================== epilogue: groups of four warps; warp w reads TMEM lanes 32*(w%4) .. +31.
    // With row AND column epilogues, group 0 does the row epilogues and group 1 the column epilogues of every chunk
    // (both read the accumulator from TMEM); with TWO row epilogues a third group takes the second one (kG3).
    // With row epilogues only, two groups split the chunks of each tile and their per-row states are merged at the end.
    constexpr bool kG3 = three_groups(NR, NC, DEBUG);
    constexpr int NRS = kG3 ? 1 : NR;          // row epilogues handled by one thread
    const int q = warp & 3;
    const int group = (warp - 2) >> 2;         // 0: warps 2..5, 1: warps 6..9, 2: warps 10..13 (kG3)
    const int trow = 32 * q + lane;            // accumulator row of this thread
    const int gt = (threadIdx.x - 64) & 127;   // thread index inside its group
    constexpr bool kSplitRoles = (NR > 0 && NC > 0);
    constexpr bool kShareRows = (NC == 0);     // all the groups work on rows (also the DEBUG dump)
    // Row epilogues only: the scan runs on packed keys (tc_scan.cuh: 4 ALU-pipe instructions per score instead of the ~12
    // of the branchy top-3 chains below -- ncu on the ZoomOut conversion: ALU pipe 67 % busy, 16 instructions per score,
    // tensor pipe 20 %); the 2^-18 truncation of the scores is covered by NNProblem::row_trunc
    constexpr bool kPacked = (NC == 0 && NR > 0 && !DEBUG);
    const bool do_rows = kSplitRoles ? group != 1 : (NC == 0 ? true : false);
    const bool do_cols = NC > 0 && group == 1;
    const int r_base = kG3 && group == 2 ? 1 : 0;  // first row epilogue of this thread
    const int row_bar = kG3 && group == 2 ? 3 : 1;
    float* patch = reinterpret_cast<float*>(sgen + OFF_PATCH + q * PATCH_BYTES);
    Top3* colred = reinterpret_cast<Top3*>(sgen + OFF_COLRED);          // [2][kMaxEpi][4][CCH]
    float* rowsb = reinterpret_cast<float*>(sgen + OFF_ROWSB);          // [e][0=scale,1=bias][TN]
    constexpr int kGroups = n_groups(NR, NC, DEBUG);
    const int row_threads = kShareRows ? kGroups * kGroupWarps * 32 : kGroupWarps * 32;

    Top3 rowst[NRS > 0 ? NRS : 1], rowsu[NRS > 0 ? NRS : 1];  // two chains per row epilogue (merged at the end)
#pragma unroll
    for (int r = 0; r < NRS; ++r) rowst[r] = rowsu[r] = top3_init();
    // row epilogue descriptors of this thread (kG3: the group's own one)
    const float* rsf[NRS > 0 ? NRS : 1];
    const float* rbf[NRS > 0 ? NRS : 1];
    bool rident[NRS > 0 ? NRS : 1];
#pragma unroll
    for (int r = 0; r < NRS; ++r) {
      const bool second = kG3 ? r_base == 1 : r == 1;
      rsf[r] = second ? P.row[NR > 1 ? 1 : 0].sf : P.row[0].sf;
      rbf[r] = second ? P.row[NR > 1 ? 1 : 0].bf : P.row[0].bf;
      rident[r] = second ? P.row[NR > 1 ? 1 : 0].identity : P.row[0].identity;
    }

    // the re-evaluation window of this thread's results without its score-dependent part (emit_result), packed scans only
    float thr_base[NRS > 0 ? NRS : 1];
#pragma unroll
    for (int r = 0; r < NRS; ++r) {
      const EpiDev& E = P.row[(kG3 ? r_base : r) < NR ? (kG3 ? r_base : r) : 0];
      const float nq_i = (kPacked && row0 + trow < nq) ? P.norm_q[q0 + row0 + trow] : 0.f;
      thr_base[r] = kPacked ? 2.f * P.eps * nq_i * E.G[p] + 9.6e-7f * E.Bm[p] : 0.f;
    }

    // scale / bias of THIS thread's accumulator row for the column epilogues (applied before the transposition)
    float csc[NC > 0 ? NC : 1], cbi[NC > 0 ? NC : 1];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int i = row0 + trow;
      const bool ok = do_cols && i < nq;
      csc[c] = ok ? __ldg(P.col[c].sf + q0 + i) : 0.f;
      cbi[c] = ok ? __ldg(P.col[c].bf + q0 + i) : -3.0e38f;  // finite: the packed keys below must not become NaN
    }

    // Row-only passes with one epilogue: the per-column scale / bias of the WHOLE database side is staged once per CTA
    // (the transposition patches, column partials and per-tile arrays are unused here, minus the exchange area of the final
    // merge: 24 KB = 3072 columns) instead of once per tile behind two group barriers and a round trip to L2 -- with
    // short contractions those ~1.5 us per tile were a third of the CTA's time
    constexpr uint32_t kXchBytes = 3 * TM_ROWS * sizeof(Top3);  // (kGroups - 1) x NR = 1 states, see the merge below
    constexpr int kRowCacheCols = int((kGroupWarps * PATCH_BYTES + COLRED_BYTES + ROWSB_BYTES - kXchBytes) / 8) / TN * TN;
    const bool cache_rows = kPacked && NRS == 1 && n_ct * TN <= kRowCacheCols;
    float* rcache = reinterpret_cast<float*>(sgen + OFF_PATCH + kXchBytes);  // [scale: n_ct TN][bias: n_ct TN]
    if (cache_rows) {
      const int rt_idx = int(threadIdx.x) - 64, n_all = n_ct * TN;
      for (int j = rt_idx; j < n_all; j += row_threads) {
        const bool v = j < nd;
        rcache[j] = v ? __ldg(rsf[0] + d0 + j) : 0.f;
        rcache[n_all + j] = v ? __ldg(rbf[0] + d0 + j) : kMaskedScore;
      }
      bar_sync_n(row_bar, row_threads);
    }

    for (int ct = 0; ct < n_ct; ++ct) {
      const int acc = ct & 1;
      const int col0 = ct * TN;
      if (NR > 0 && do_rows && !cache_rows) {
        // per-column scale / bias of the row epilogues for this tile
        bar_sync_n(row_bar, row_threads);  // everyone is done with the previous tile's arrays
        const int rt_idx = kShareRows ? int(threadIdx.x) - 64 : gt;
#pragma unroll
        for (int r = 0; r < NRS; ++r)
          for (int jj = rt_idx; jj < TN; jj += row_threads) {
            const int j = col0 + jj;
            const bool v = j < nd;
            rowsb[((r_base + r) * 2 + 0) * TN + jj] = v ? __ldg(rsf[r] + d0 + j) : 0.f;
            rowsb[((r_base + r) * 2 + 1) * TN + jj] = v ? __ldg(rbf[r] + d0 + j) : (kPacked ? kMaskedScore : -INFINITY);
          }
        bar_sync_n(row_bar, row_threads);
      }

      mbar_wait(bar_tfull + 8 * acc, (ct >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * TN + (uint32_t(32 * q) << 16);
      const int n_ch = min(TN, nd - col0 + CCH - 1) / CCH;  // chunks that hold at least one valid column
      const int ch_beg = kShareRows ? group * (TN / CCH / kGroups) : 0;
      const int ch_end = kShareRows ? ch_beg + TN / CCH / kGroups : TN / CCH;
      for (int ch = ch_beg; ch < ch_end; ++ch) {
        if (ch >= n_ch) break;  // uniform over the group
        if ((!do_rows && !do_cols && !DEBUG) || P.probe_skip_epilogue) break;
        float v[32];
        tmem_ld32(taddr + ch * CCH, v);
        if (DEBUG) {
          const int i = row0 + trow;
          if (i < nq)
#pragma unroll
            for (int c = 0; c < CCH; ++c) {
              const int j = col0 + ch * CCH + c;
              if (j < nd) dbg.S[int64_t(i) * dbg.ldS + j] = v[c];
            }
        }
        if (NR > 0 && do_rows) {
          const bool full_tile = col0 + TN <= nd;  // no masked tail columns in this tile
          const int jb = col0 + ch * CCH;
          if (kPacked) {
#pragma unroll
            for (int r = 0; r < NRS; ++r) {
              float k1, k2;
              if (rident[r] && full_tile) {
                t2_chunk_top2<true>(v, nullptr, nullptr, keymask, k1, k2);
                t2_merge_chunk<true>(rowst[r], k1, k2, jb, v, nullptr, nullptr, keymask, thr_base[r]);
              } else {
                const float* sc = cache_rows ? rcache + col0 + ch * CCH : rowsb + ((r_base + r) * 2 + 0) * TN + ch * CCH;
                const float* bi = cache_rows ? rcache + n_ct * TN + col0 + ch * CCH : rowsb + ((r_base + r) * 2 + 1) * TN + ch * CCH;
                t2_chunk_top2<false>(v, sc, bi, keymask, k1, k2);
                t2_merge_chunk<false>(rowst[r], k1, k2, jb, v, sc, bi, keymask, thr_base[r]);
              }
            }
          } else if (NRS == 2 && !(full_tile && (rident[0] || rident[NRS - 1]))) {
            // both epilogues carry scale / bias: interleave them (one vote, four independent chains)
            const float4* s40 = reinterpret_cast<const float4*>(rowsb + 0 * TN + ch * CCH);
            const float4* b40 = reinterpret_cast<const float4*>(rowsb + 1 * TN + ch * CCH);
            const float4* s41 = reinterpret_cast<const float4*>(rowsb + 2 * TN + ch * CCH);
            const float4* b41 = reinterpret_cast<const float4*>(rowsb + 3 * TN + ch * CCH);
#pragma unroll
            for (int c4 = 0; c4 < CCH / 4; ++c4) {
              const float4 s0 = s40[c4], b0 = b40[c4], s1 = s41[c4], b1 = b41[c4];
              const float w[4] = {fmaf(v[4 * c4 + 0], s0.x, b0.x), fmaf(v[4 * c4 + 1], s0.y, b0.y),
                                  fmaf(v[4 * c4 + 2], s0.z, b0.z), fmaf(v[4 * c4 + 3], s0.w, b0.w)};
              const float u[4] = {fmaf(v[4 * c4 + 0], s1.x, b1.x), fmaf(v[4 * c4 + 1], s1.y, b1.y),
                                  fmaf(v[4 * c4 + 2], s1.z, b1.z), fmaf(v[4 * c4 + 3], s1.w, b1.w)};
              top3_offer4x2(rowst[0], rowsu[0], rowst[NRS - 1], rowsu[NRS - 1], w, u, jb + 4 * c4);
            }
          } else {
#pragma unroll
            for (int r = 0; r < NRS; ++r) {
              if (rident[r] && full_tile) {  // plain dot-product argmax: no scale / bias traffic
#pragma unroll
                for (int c4 = 0; c4 < CCH / 4; ++c4)
                  top3_offer4(rowst[r], rowsu[r], v[4 * c4 + 0], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3], jb + 4 * c4);
              } else {
                const float4* s4 = reinterpret_cast<const float4*>(rowsb + ((r_base + r) * 2 + 0) * TN + ch * CCH);
                const float4* b4 = reinterpret_cast<const float4*>(rowsb + ((r_base + r) * 2 + 1) * TN + ch * CCH);
#pragma unroll
                for (int c4 = 0; c4 < CCH / 4; ++c4) {
                  const float4 s = s4[c4], b = b4[c4];
                  top3_offer4(rowst[r], rowsu[r], fmaf(v[4 * c4 + 0], s.x, b.x), fmaf(v[4 * c4 + 1], s.y, b.y),
                              fmaf(v[4 * c4 + 2], s.z, b.z), fmaf(v[4 * c4 + 3], s.w, b.w), jb + 4 * c4);
                }
              }
            }
          }
        }
        if (NC > 0 && do_cols) {
          // per column epilogue: scale / bias by the own row, transpose the warp's 32x32 block through the patch
          // (lane becomes the column), scan the 32 rows in two independent ascending chains (rows 0..15, 16..31)
          Top3 cst[NC > 0 ? NC : 1];
          const int ib = row0 + 32 * q;
          float w[NC > 0 ? NC : 1][32];
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            if (c > 0) __syncwarp();  // the previous epilogue's reads of the patch are done
            if (P.col[c].identity && row0 + TM_ROWS <= nq) {  // plain dot products, no masked tail rows
#pragma unroll
              for (int c4 = 0; c4 < CCH / 4; ++c4)
                *reinterpret_cast<float4*>(patch + lane * PATCH_LD + 4 * c4) =
                    make_float4(v[4 * c4 + 0], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
            } else {
#pragma unroll
              for (int c4 = 0; c4 < CCH / 4; ++c4)
                *reinterpret_cast<float4*>(patch + lane * PATCH_LD + 4 * c4) =
                    make_float4(fmaf(v[4 * c4 + 0], csc[c], cbi[c]), fmaf(v[4 * c4 + 1], csc[c], cbi[c]),
                                fmaf(v[4 * c4 + 2], csc[c], cbi[c]), fmaf(v[4 * c4 + 3], csc[c], cbi[c]));
            }
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 32; ++r) w[c][r] = patch[r * PATCH_LD + lane];
          }
          // Running top-2 of the 32 rows on PACKED keys: the 5 low mantissa bits of the score are replaced by the row
          // (r = 0..31), so that best and runner-up with their rows cost three FMNMX per element instead of
          // compares and selects.  The relative error 2^-18 this adds to the column scores is part of the re-evaluation
          // threshold (NNProblem::col_trunc).  All epilogues' chains advance together: 2 NC independent chains.
          float k1a[NC > 0 ? NC : 1], k2a[NC > 0 ? NC : 1], k1b[NC > 0 ? NC : 1], k2b[NC > 0 ? NC : 1];
#pragma unroll
          for (int c = 0; c < NC; ++c) k1a[c] = k2a[c] = k1b[c] = k2b[c] = -INFINITY;
#pragma unroll
          for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              const float ka = __int_as_float((__float_as_int(w[c][r]) & ~31) | r);
              const float kb = __int_as_float((__float_as_int(w[c][16 + r]) & ~31) | (16 + r));
              k2a[c] = fmaxf(k2a[c], fminf(k1a[c], ka));
              k1a[c] = fmaxf(k1a[c], ka);
              k2b[c] = fmaxf(k2b[c], fminf(k1b[c], kb));
              k1b[c] = fmaxf(k1b[c], kb);
            }
          }
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const float k1 = fmaxf(k1a[c], k1b[c]);
            const float k2 = fmaxf(fminf(k1a[c], k1b[c]), fmaxf(k2a[c], k2b[c]));
            Top3 t;
            t.m1 = __int_as_float(__float_as_int(k1) & ~31), t.i1 = ib + (__float_as_int(k1) & 31);
            t.m2 = __int_as_float(__float_as_int(k2) & ~31), t.i2 = ib + (__float_as_int(k2) & 31);
            t.m3 = t.m2;  // conservative: the third-best of the 32 rows is not tracked
            cst[c] = t;
          }
          const int par = ch & 1;
#pragma unroll
          for (int c = 0; c < NC; ++c) colred[((par * kMaxEpi + c) * kGroupWarps + q) * CCH + lane] = cst[c];
          bar_sync_n(2, kGroupWarps * 32);  // also orders the patch reuse of the next chunk
          // warp c of the group merges the four row quarters of column epilogue c and writes the partial of this row
          // tile.  The merge order must be ascending in the row index: TMEM quarter q' holds rows 32 q' .. 32 q' + 31.
          // the merging warp rotates with the chunk so that the extra work is spread over the four warps
          const int slot = ((warp - 2) - ch) & 3;  // 0..3
          if (slot < NC) {
            const int j = col0 + ch * CCH + lane;
            Top3 m = colred[((par * kMaxEpi + slot) * kGroupWarps + 0) * CCH + lane];
#pragma unroll
            for (int w = 1; w < kGroupWarps; ++w) top3_merge(m, colred[((par * kMaxEpi + slot) * kGroupWarps + w) * CCH + lane]);
            if (j < nd && active) P.col_partial[((int64_t(slot) * P.n_pairs + p) * P.max_rt + rt) * P.max_db + j] = m;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && rank != 0)
          mbar_arrive_cluster(mapa_shared(bar_tempty + 8 * acc, 0));
        else
          mbar_arrive(bar_tempty + 8 * acc);
      }
    }

    if (NR > 0) {
#pragma unroll
      for (int r = 0; r < NRS; ++r) top3_merge(rowst[r], rowsu[r]);
      if (kShareRows) {
        // the other groups hand their per-row states to group 0 through shared memory (the patch area is free: NC == 0);
        // merged in ascending chunk order, i.e. ascending column index inside every tile
        static_assert((kGroups - 1) * (NR > 0 ? NR : 1) * TM_ROWS * sizeof(Top3) <= kGroupWarps * PATCH_BYTES, "exchange area");
        Top3* xch = reinterpret_cast<Top3*>(sgen + OFF_PATCH);  // [kGroups - 1][NR][128]
        if (group > 0) {
#pragma unroll
          for (int r = 0; r < NR; ++r) xch[((group - 1) * NR + r) * TM_ROWS + trow] = rowst[r];
        }
        bar_sync_n(1, row_threads);
        if (group == 0) {
#pragma unroll
          for (int g2 = 1; g2 < kGroups; ++g2)
#pragma unroll
            for (int r = 0; r < NR; ++r) top3_merge(rowst[r], xch[((g2 - 1) * NR + r) * TM_ROWS + trow]);
        }
      }
      const int i = row0 + trow;
      if (kG3) {
        if (i < nq && group == 0) emit_result(P, P.row[0], false, 0, p, q0 + i, i, P.norm_q[q0 + i], rowst[0]);
        if (i < nq && group == 2) emit_result(P, P.row[NR > 1 ? 1 : 0], false, 1, p, q0 + i, i, P.norm_q[q0 + i], rowst[0]);
      } else if (group == 0) {
#pragma unroll
        for (int r = 0; r < NR; ++r)
          if (i < nq) emit_result(P, P.row[r], false, r, p, q0 + i, i, P.norm_q[q0 + i], rowst[r]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the leader's MMAs wrote the peer's tensor memory; remote arrivals are all delivered
  if (warp == 1) {
    tc_fence_after();
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ------------------------------------------------------------------ host side
int make_map(CUtensorMap* m, const void* base, int64_t rows, int kp, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) DM_FAIL(DM_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[2] = {cuuint64_t(kp), cuuint64_t(rows > 0 ? rows : 1)};
  const cuuint64_t gstr[1] = {cuuint64_t(kp) * 2};
  const cuuint32_t box[2] = {cuuint32_t(TBK), cuuint32_t(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) DM_FAIL(DM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
  return DM_OK;
}

template <int NR, int NC, bool DEBUG, bool PAIR>
int launch_mode(const TcMaps& maps, const NNProblem& P, const DebugOut& dbg, cudaStream_t st) {
  static OncePerDevice attr_once;
  if (attr_once.first()) {
    DM_CUDA_OK(cudaFuncSetAttribute(nn_tc_kernel<NR, NC, DEBUG, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    int(SMEM_BYTES)));
  }
  const int mrt = PAIR ? (P.max_rt + 1) & ~1 : P.max_rt;
  const int64_t nblk = int64_t(P.n_pairs) * mrt;
  if (nblk > 0x7fffffffLL) DM_FAIL(DM_ERR_BADARG, "too many row tiles (%lld)", (long long)nblk);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)nblk);
  cfg.blockDim = dim3(n_threads(NR, NC, DEBUG));
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = PAIR ? 1 : 0;
  DM_CUDA_OK(cudaLaunchKernelEx(&cfg, nn_tc_kernel<NR, NC, DEBUG, PAIR>, maps, P, dbg, ~uint32_t(T2_CH - 1)));
  return DM_OK;
}

// DM_NN_SINGLE_CTA=1 in the environment selects the single-CTA (cta_group::1) kernel for A/B measurements;
// DM_NN_FORCE_PAIR=1 the CTA-pair kernel for every non-debug launch (tests of its ragged / odd-tile handling)
bool env_flag(const char* name) {
  const char* e = getenv(name);
  return e && e[0] == '1';
}
bool use_pair_mode() {
  static const bool single = env_flag("DM_NN_SINGLE_CTA");
  return !single;
}
bool force_pair_mode() {
  static const bool force = env_flag("DM_NN_FORCE_PAIR");
  return force;
}

// The pair kernel pays off when the MMA pipeline is the limiter (long contractions: the d = 384 feature search, the
// upper rungs of the ZoomOut ladder).  With a short contraction the epilogues bound the tile time, and coupling two
// CTAs' accumulator hand-offs only adds lock-step stalls (FM -> p2p at k = 100: 2.39 ms single vs 2.44 ms pair).
template <int NR, int NC, bool DEBUG>
int launch(const TcMaps& maps, const TcMaps& maps_pair, const NNProblem& P, const DebugOut& dbg, cudaStream_t st) {
  // ... and when no CTA of a pair would idle: all query sets of one size with an even number of row tiles (a ragged
  // batch pairs the last odd tile of a mesh with an empty one: 94.7 k vs 97.7 k pairs/s on the 1024-pair ragged config)
  const bool even_tiles = P.max_rt % 2 == 0 && P.total_q == int64_t(P.n_pairs) * P.max_q;
  if (!DEBUG && ((use_pair_mode() && P.kp >= 4 * TBK && even_tiles) || force_pair_mode()))
    return launch_mode<NR, NC, false, true>(maps_pair, P, dbg, st);
  return launch_mode<NR, NC, DEBUG, false>(maps, P, dbg, st);
}

}  // namespace

int nn_tc_kp(int d) { return (d + TBK - 1) / TBK * TBK; }

int tc_make_map_bf16(void* tensor_map, const void* base, int64_t rows, int kp, int box_rows) {
  return make_map(static_cast<CUtensorMap*>(tensor_map), base, rows, kp, box_rows);
}

int nn_tc_launch(const NNProblem& P, const void* Yh, const void* Yl, const void* Xh, const void* Xl, float* dbgS,
                 int64_t ldS, cudaStream_t st) {
  if (P.n_pairs <= 0 || P.total_q <= 0) return DM_OK;
  if (P.rows_q > 0x7fffffffLL || P.rows_db > 0x7fffffffLL) DM_FAIL(DM_ERR_BADARG, "too many rows for TMA coordinates");
  TcMaps maps;
  int rc;
  if ((rc = make_map(&maps.yh, Yh, P.rows_q, P.kp, TM_ROWS))) return rc;
  if ((rc = make_map(&maps.yl, Yl, P.rows_q, P.kp, TM_ROWS))) return rc;
  if ((rc = make_map(&maps.xh, Xh, P.rows_db, P.kp, TN))) return rc;
  if ((rc = make_map(&maps.xl, Xl, P.rows_db, P.kp, TN))) return rc;
  TcMaps maps_pair = maps;  // CTA-pair mode: each CTA stages half of the X tile
  if ((rc = make_map(&maps_pair.xh, Xh, P.rows_db, P.kp, TN / 2))) return rc;
  if ((rc = make_map(&maps_pair.xl, Xl, P.rows_db, P.kp, TN / 2))) return rc;
  DebugOut dbg{dbgS, ldS};
  if (dbgS) {
    if ((rc = launch<0, 0, true>(maps, maps_pair, P, dbg, st))) return rc;
  } else {
    const int key = P.n_row * 10 + P.n_col;
    switch (key) {
      case 0 * 10 + 1: rc = launch<0, 1, false>(maps, maps_pair, P, dbg, st); break;
      case 0 * 10 + 2: rc = launch<0, 2, false>(maps, maps_pair, P, dbg, st); break;
      case 1 * 10 + 0: rc = launch<1, 0, false>(maps, maps_pair, P, dbg, st); break;
      case 1 * 10 + 1: rc = launch<1, 1, false>(maps, maps_pair, P, dbg, st); break;
      case 1 * 10 + 2: rc = launch<1, 2, false>(maps, maps_pair, P, dbg, st); break;
      case 2 * 10 + 0: rc = launch<2, 0, false>(maps, maps_pair, P, dbg, st); break;
      case 2 * 10 + 1: rc = launch<2, 1, false>(maps, maps_pair, P, dbg, st); break;
      case 2 * 10 + 2: rc = launch<2, 2, false>(maps, maps_pair, P, dbg, st); break;
      default: DM_FAIL(DM_ERR_BADARG, "unsupported epilogue combination %d row / %d col", P.n_row, P.n_col);
    }
    if (rc) return rc;
  }
  DM_LAUNCH_OK("nn_tc_kernel");
  return DM_OK;
}

}  // namespace dm
