// Micro-benchmark: peak rate of mma.sync.m8n8k4.f64 (DMMA) and of DFMA on this GPU, registers only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu && ./dmma_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void __launch_bounds__(256) dmma(double* out, int iters) {
  double c[NACC][2];
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the same with a distinct B fragment per accumulator (registers), and with the fragments re-read from shared memory
// before every use (the access pattern of a real GEMM inner loop: one LDS.64 per DMMA, issued a full step ahead)
template <int NACC, bool SMEM>
__global__ void __launch_bounds__(256) dmma_operands(double* out, int iters) {
  __shared__ double sb[NACC * 4 * 12 + 64];
  for (int i = threadIdx.x; i < NACC * 4 * 12 + 64; i += blockDim.x) sb[i] = 1.0 + i * 1e-9;
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  double c[NACC][2], b[2][NACC];
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0, b[0][i] = 1.0 + i * 1e-9 + lane, b[1][i] = 1.0 - i * 1e-9 + lane;
  double a = 1.0 + threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it += 2) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (SMEM) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) b[h ^ 1][i] = sb[t4 * 12 + g + 8 * i + ((it + h) & 7)];
      }
#pragma unroll
      for (int i = 0; i < NACC; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b[h][i]));
    }
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma(double* out, int iters) {
  double c[NACC];
  for (int i = 0; i < NACC; ++i) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  double* out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int ctas = 1; ctas <= 4; ctas *= 2) {
    dmma<8><<<148 * ctas, 256>>>(out, 100); cudaDeviceSynchronize();
    cudaEventRecord(e0); dmma<8><<<148 * ctas, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flop = 2.0 * 256 * 8 * double(iters) * 8 * 148 * ctas;   // 256 FMA per DMMA, 8 per iteration, 8 warps per CTA
    printf("DMMA  %d CTA/SM x 8 warps, 8 independent accumulators: %7.2f TFLOP/s\n", ctas, flop / ms / 1e9);
  }
  {
    float ms;
    double flop = 2.0 * 256 * 13 * double(iters) * 8 * 148;
    dmma_operands<13, false><<<148, 256>>>(out, 100); cudaDeviceSynchronize();
    cudaEventRecord(e0); dmma_operands<13, false><<<148, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DMMA  1 CTA/SM x 8 warps, 13 accumulators, distinct B registers:  %7.2f TFLOP/s\n", flop / ms / 1e9);
    dmma_operands<13, true><<<148, 256>>>(out, 100); cudaDeviceSynchronize();
    cudaEventRecord(e0); dmma_operands<13, true><<<148, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DMMA  1 CTA/SM x 8 warps, 13 accumulators, B from LDS.64 each use: %7.2f TFLOP/s\n", flop / ms / 1e9);
  }
  for (int ctas = 1; ctas <= 4; ctas *= 2) {
    dfma<8><<<148 * ctas, 256>>>(out, 100); cudaDeviceSynchronize();
    cudaEventRecord(e0); dfma<8><<<148 * ctas, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flop = 2.0 * 8 * double(iters) * 256 * 148 * ctas;
    printf("DFMA  %d CTA/SM x 256 threads, 8 independent chains:      %7.2f TFLOP/s\n", ctas, flop / ms / 1e9);
  }
  return 0;
}
