// Incremental p2p -> FM for the ZoomOut ladder (densematcher/pyFM/refine/zoomout.py:38-42, 110-113).
//
// Every rung recomputes  C+ = Phi2[:, :k2+s2]^T A2 Phi1[p, :k1+s1]  from scratch in the reference: 2 N (k+s)^2 flops, the
// largest float64 item of the ladder (40 % of its time as a full tensor-core GEMM).  But from one rung to the next only a
// few per cent of the vertex map p changes (measured: 5 % on the icosphere(4) golden, 2 % on the synthetic bench pairs,
// scripts/zo_delta_probe.py), and every rung's map is the leading block of ONE matrix
//     M = Phi2[:, :K2]^T A2 Phi1[p, :K1]          (K1, K2 = the widths of the last rung)
// which changes by  sum_{n : p[n] != p_old[n]} Phi2[n]^T a2[n] (Phi1[p[n]] - Phi1[p_old[n]])  from rung to rung.  The ladder
// keeps M resident, corrects it with the changed vertices only (|changed| K^2 instead of N k^2 multiply-adds), and hands
// its leading (k2+s2) x (k1+s1) block to the next rung:
//   delta_fill_kernel                       the changed vertices of every pair, in vertex order (deterministic), as
//                                           interleaved (+a2, p) / (-a2, p_old) entries of a gather list (one fixed
//                                           segment per pair: no count / scan pass over the pairs)
//   gemm64_tt_kernel                        the correction as a gathered "both transposed" GEMM with M as its addend
//   extract_block_kernel                    the dense leading block for the next conversion
// The ladder re-anchors M with a full product every 64 rungs, and the last rung is always a fresh product, so the rounding
// of the running sum never exceeds that of a few dozen float64 additions per entry.
#include "dm_internal.cuh"
#include "gemm64.cuh"

namespace dm {
namespace {

// The changed vertices of pair b, in vertex order, as interleaved gather entries (vertex n, +a2, p_new) / (vertex n, -a2,
// p_old).  Pair b owns the fixed segment [2 off2[b], 2 off2[b + 1]) of the lists (capacity: every vertex changed), so one
// kernel does what needed a count, a scan over the pairs and a fill: seg[b] = start of the segment, cnt[b] = entries used.
__global__ void __launch_bounds__(256)
    delta_fill_kernel(const void* __restrict__ p_new, const void* __restrict__ p_old, int i64, const int64_t* __restrict__ off2,
                      const double* __restrict__ area2, int64_t* __restrict__ seg, int* __restrict__ cnt,
                      int32_t* __restrict__ ga, int32_t* __restrict__ gb, double* __restrict__ dscale) {
  __shared__ int wcount[8];
  __shared__ int running_s;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r0 = off2[b], o = 2 * r0;
  const int n = int(off2[b + 1] - r0);
  if (threadIdx.x == 0) running_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 256) {
    const int i = base + threadIdx.x;
    int64_t pn = 0, po = 0;
    bool ch = false;
    if (i < n) {
      pn = load_index(p_new, r0 + i, i64 != 0), po = load_index(p_old, r0 + i, i64 != 0);
      ch = pn != po;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, ch);
    if (lane == 0) wcount[warp] = __popc(bal);
    __syncthreads();
    int before = running_s;
    for (int w = 0; w < warp; ++w) before += wcount[w];
    if (ch) {
      const int64_t j = o + 2 * int64_t(before + __popc(bal & ((1u << lane) - 1u)));
      const double a = area2[r0 + i];
      ga[j] = i, gb[j] = int32_t(pn), dscale[j] = a;
      ga[j + 1] = i, gb[j + 1] = int32_t(po), dscale[j + 1] = -a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = running_s;
      for (int w = 0; w < 8; ++w) t += wcount[w];
      running_s = t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) seg[b] = o, cnt[b] = 2 * running_s;
}

// dense [n_pairs, k2, k1] copy of the leading block of M [n_pairs, K2, K1]
__global__ void __launch_bounds__(256)
    extract_block_kernel(const double* __restrict__ M, int K1, int K2, int k1, int k2, int n_pairs, double* __restrict__ C) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= int64_t(n_pairs) * k1 * k2) return;
  const int c = int(idx % k1);
  const int64_t t = idx / k1;
  const int r = int(t % k2), b = int(t / k2);
  C[idx] = M[(int64_t(b) * K2 + r) * K1 + c];
}

struct DeltaLayout {
  int* cnt;
  int64_t* doff;
  int32_t *ga, *gb;
  double* dscale;
  size_t bytes;
};
DeltaLayout delta_carve(void* ws, int n_pairs, int64_t total_n2) {
  Carver c(ws);
  DeltaLayout L{};
  L.cnt = c.take<int>(size_t(n_pairs));
  L.doff = c.take<int64_t>(size_t(n_pairs) + 1);
  L.ga = c.take<int32_t>(size_t(total_n2) * 2 + 2);
  L.gb = c.take<int32_t>(size_t(total_n2) * 2 + 2);
  L.dscale = c.take<double>(size_t(total_n2) * 2 + 2);
  L.bytes = c.bytes();
  return L;
}

}  // namespace

bool p2p_to_fm_delta_applicable() {
  static const bool off = [] { const char* e = getenv("DM_ZO_FULL_GEMM"); return e && e[0] == '1'; }();
  return !off;
}

size_t p2p_to_fm_delta_ws(int n_pairs, int64_t total_n2) { return delta_carve(nullptr, n_pairs, total_n2).bytes; }

// M [n_pairs, K2, K1] = Phi2[:, :K2]^T A2 Phi1[p_old, :K1]  ->  the same for p_new, in place
int p2p_to_fm_delta_run(const void* p_new, const void* p_old, int i64, const double* Phi1, int64_t ld1, const int64_t* off1,
                        const double* Phi2, int64_t ld2, const int64_t* off2, int64_t total_n2, int max_n2, const double* area2,
                        int n_pairs, int K1, int K2, double* M, void* ws, cudaStream_t st) {
  DeltaLayout L = delta_carve(ws, n_pairs, total_n2);
  delta_fill_kernel<<<n_pairs, 256, 0, st>>>(p_new, p_old, i64, off2, area2, L.doff, L.cnt, L.ga, L.gb, L.dscale);
  DM_LAUNCH_OK("delta_fill_kernel");
  GemmProblem G;
  G.A.d = Phi2, G.A.ld = ld2, G.A.off = off2, G.A.trans = 1, G.A.gather = L.ga, G.A.gather_off = L.doff, G.A.kscale = L.dscale;
  G.A.gather_cnt = L.cnt;
  G.B.d = Phi1, G.B.ld = ld1, G.B.off = off1, G.B.trans = 1, G.B.gather = L.gb, G.B.gather_off = L.doff, G.B.gather_cnt = L.cnt;
  G.M = K2, G.N = K1, G.maxM = K2, G.maxN = K1, G.maxK = 2 * max_n2, G.n_batch = n_pairs;
  G.C = M, G.ldc = K1, G.c_batch_stride = int64_t(K2) * K1;
  G.c_add = M, G.c_add_ld = K1, G.c_add_batch_stride = int64_t(K1) * K2;  // every entry is read and written by one thread
  return gemm64_launch(G, st);
}

int extract_block_run(const double* M, int K1, int K2, int k1, int k2, int n_pairs, double* C, cudaStream_t st) {
  const int64_t tot = int64_t(n_pairs) * k1 * k2;
  if (tot <= 0) return DM_OK;
  extract_block_kernel<<<unsigned((tot + 255) / 256), 256, 0, st>>>(M, K1, K2, k1, k2, n_pairs, C);
  DM_LAUNCH_OK("extract_block_kernel");
  return DM_OK;
}

}  // namespace dm
