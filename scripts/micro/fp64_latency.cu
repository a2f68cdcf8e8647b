// Micro-benchmark: dependent-chain latencies of the FP64 operations on the critical path of the Cholesky kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double rsqrt64_cvt(double v) {  // fp32 estimate through conversions + two Newton steps
  double y = double(rsqrtf(float(v)));
  const double h = 0.5 * v;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  return y;
}
__device__ __forceinline__ double rsqrt64_mufu(double v) {  // MUFU.RSQ64H estimate + two Newton steps
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
  const double h = 0.5 * v;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  return y;
}

template <int MODE>
__global__ void chain(double* out, long long* cyc, double x0, int n) {
  double x = x0 + threadIdx.x * 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if (MODE == 0) x = fma(x, 1.0000001, 1e-9);
    if (MODE == 1) x = x * 1.0000001;
    if (MODE == 2) x = rsqrt64_cvt(x) + 1.0;
    if (MODE == 3) x = rsqrt64_mufu(x) + 1.0;
    if (MODE == 4) x = 1.0 / sqrt(x) + 1.0;
    if (MODE == 5) x = double(float(x)) + 1e-9;
    if (MODE == 6) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y + 1.0; }
    if (MODE == 7) { float f = __double2float_rn(x); f = fmaf(f, 1.0000001f, 1e-9f); x = f; }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMallocManaged(&cyc, 8);
  const int n = 4096;
  const char* names[] = {"DFMA", "DMUL", "rsqrt64 via f32 conversions (+DADD)", "rsqrt64 via MUFU.RSQ64H (+DADD)",
                         "1/sqrt (library) (+DADD)", "F2F f64->f32->f64 (+DADD)", "MUFU.RSQ64H alone (+DADD)", "F2F + FFMA + F2F"};
#define RUN(M)                                                                     \
  chain<M><<<1, 32>>>(out, cyc, 1.5, n); cudaDeviceSynchronize();                  \
  chain<M><<<1, 32>>>(out, cyc, 1.5, n); cudaDeviceSynchronize();                  \
  printf("%-42s %8.1f cycles per dependent step\n", names[M], double(*cyc) / n);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7)
  return 0;
}
