// Packed-key row scans shared by the tcgen05 engines (nn_tc.cu row-only passes, nn_tc2.cu): a thread owns one accumulator
// row, walks it in chunks of 32 columns (one tcgen05.ld.32x32b.x32) and keeps (best, runner-up, third) of everything seen.
#pragma once
#include "dm_internal.cuh"

namespace dm {
namespace {

constexpr int T2_CH = 32;      // columns per epilogue chunk
constexpr float kMaskedScore = -3.0e38f;  // finite: packed keys must not become NaN

// key = (score bits & mask) | column as ONE LOP3 (LUT 0xEA = (a & b) | c).  `mask` (= ~31) must sit in a register the
// assembler cannot fold -- it arrives as a kernel parameter -- otherwise the two immediates cost two ALU-pipe
// instructions, and the ALU pipe (min / max / logic) is what bounds this kernel (ncu: 81 % busy, FMA pipe 10 %).
__device__ __forceinline__ float t2_key(float w, uint32_t mask, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(__float_as_uint(w)), "r"(mask), "r"(c));
  return __uint_as_float(r);
}

// Top-2 of one 32-column chunk on packed keys.  w_c = v_c * s_c + b_c (IDENT: w_c = v_c); key = (w & ~31) | c.
template <bool IDENT>
__device__ __forceinline__ void t2_chunk_top2(const float (&v)[32], const float* __restrict__ sc,
                                              const float* __restrict__ bi, uint32_t mask, float& k1, float& k2) {
  k1 = k2 = -INFINITY;
#pragma unroll
  for (int c4 = 0; c4 < T2_CH / 4; ++c4) {
    float w0 = v[4 * c4 + 0], w1 = v[4 * c4 + 1], w2 = v[4 * c4 + 2], w3 = v[4 * c4 + 3];
    if (!IDENT) {
      const float4 s = *reinterpret_cast<const float4*>(sc + 4 * c4), b = *reinterpret_cast<const float4*>(bi + 4 * c4);
      w0 = fmaf(w0, s.x, b.x), w1 = fmaf(w1, s.y, b.y), w2 = fmaf(w2, s.z, b.z), w3 = fmaf(w3, s.w, b.w);
    }
    const float ka = t2_key(w0, mask, 4 * c4 + 0), kb = t2_key(w1, mask, 4 * c4 + 1);
    const float kc = t2_key(w2, mask, 4 * c4 + 2), kd = t2_key(w3, mask, 4 * c4 + 3);
    const float h1 = fmaxf(ka, kb), l1 = fminf(ka, kb), h2 = fmaxf(kc, kd), l2 = fminf(kc, kd);
    const float t1 = fmaxf(h1, h2);
    const float t2 = fmaxf(fmaxf(fminf(h1, h2), l1), l2);  // (the three-input FMNMX3 of sm_100 measured no faster here)
    k2 = fmaxf(fmaxf(fminf(k1, t1), k2), t2);
    k1 = fmaxf(k1, t1);
  }
}

// Third-largest key of a chunk, given its two leaders (only needed in the rare case below)
template <bool IDENT>
__device__ __forceinline__ float t2_chunk_third(const float (&v)[32], const float* __restrict__ sc, const float* __restrict__ bi,
                                             uint32_t mask, float k1, float k2) {
  float k3 = -INFINITY;
#pragma unroll
  for (int c = 0; c < T2_CH; ++c) {
    const float w = IDENT ? v[c] : fmaf(v[c], sc[c], bi[c]);
    const float k = t2_key(w, mask, c);
    k3 = fmaxf(k3, (k == k1 || k == k2) ? -INFINITY : k);
  }
  return k3;
}

// Merges the chunk's leaders into the running (best, runner-up, third).  Everything else of the chunk is <= its
// runner-up m2, which serves as the chunk's (conservative) third value.  That bound is loose in exactly one case: the
// chunk's two leaders are also the two running leaders -- then third == runner-up and a near tie could only be settled
// by a scan of ALL candidates.  In that case (rare after the first few chunks, warp-uniform test) the chunk's true third
// is computed from the scores still in registers.
template <bool IDENT>
__device__ __forceinline__ void t2_merge_chunk(Top3& st, float k1, float k2, int base, const float (&v)[32],
                                               const float* __restrict__ sc, const float* __restrict__ bi, uint32_t mask,
                                               float thr_base) {
  const float m1 = __int_as_float(__float_as_int(k1) & ~31), m2 = __int_as_float(__float_as_int(k2) & ~31);
  const int i1 = base + (__float_as_int(k1) & 31), i2 = base + (__float_as_int(k2) & 31);
  // ... and it only matters when the chunk's two leaders are close enough to be re-evaluated at all (the window of
  // emit_result, with slack): a fraction of a percent of the (row, chunk) pairs, so the common path is the plain merge
  const bool near = !((m1 - m2) > 1.25f * thr_base + 8.0e-6f * (fabsf(m1) + fabsf(m2)));
  if (!__any_sync(0xffffffffu, near)) {
    top3_merge(st, m1, i1, m2, i2, m2);
    return;
  }
  const Top3 prev = st;
  top3_merge(st, m1, i1, m2, i2, m2);
  if (near && ((st.i1 == i1 && st.i2 == i2) || (st.i1 == i2 && st.i2 == i1))) {
    const float k3 = t2_chunk_third<IDENT>(v, sc, bi, mask, k1, k2);
    st = prev;
    top3_merge(st, m1, i1, m2, i2, __int_as_float(__float_as_int(k3) & ~31));
  }
}


}  // namespace
}  // namespace dm
