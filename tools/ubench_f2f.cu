// float <-> double conversions (F2F.F64.F32 / F2F.F32.F64) on B200 against the DFMA pipe: does a conversion occupy the fp64
// pipe, at what rate, and what does the integer-ALU widening (exponent re-bias with shifts) cost beside DFMA work?
// The chunked IIR scans convert every sample twice per filter group; this decides whether that is free.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_f2f tools/ubench_f2f.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
__device__ __forceinline__ double widen_alu(float f) {          // exact for normal floats and zero
  const unsigned b = __float_as_uint(f), a = b & 0x7fffffffu;
  unsigned hi = (a >> 3) + 0x38000000u;
  hi = a ? hi : 0u;
  return __hiloint2double((int)(hi | (b & 0x80000000u)), (int)(b << 29));
}
__global__ void k_dfma(double* out, double a, double b) {
  double v[8];
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fma(v[i], a, b);
  double s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 8 conversions float -> double per iteration, consumed by integer XORs (no fp64 arithmetic)
__global__ void k_f2d(double* out, float a) {
  float f[8]; long long acc = 0;
  for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 0.37f + i;
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) { const double d = (double)f[i]; acc ^= __double_as_longlong(d); f[i] = __int_as_float(__float_as_int(f[i]) + 1); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (double)acc;
}
__global__ void k_d2f(double* out, double a) {
  double d[8]; int acc = 0;
  for (int i = 0; i < 8; ++i) d[i] = threadIdx.x * 0.37 + i;
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float f = (float)d[i]; acc ^= __float_as_int(f); d[i] = __longlong_as_double(__double_as_longlong(d[i]) + 0x20000000ll); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (double)acc;
}
// the scan's pass 1: per sample one conversion and four DFMAs
template <int MODE>   // 0: F2F, 1: ALU widening, 2: no conversion (upper bound)
__global__ void k_pass1(double* out, double w0, double w1, double w2, double w3) {
  float f[8]; double p[4] = {0, 0, 0, 0};
  for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 0.37f + i;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double x;
      if (MODE == 0) x = (double)f[i];
      else if (MODE == 1) x = widen_alu(f[i]);
      else x = __hiloint2double(__float_as_int(f[i]), 0);
      p[0] = fma(w0, x, p[0]); p[1] = fma(w1, x, p[1]); p[2] = fma(w2, x, p[2]); p[3] = fma(w3, x, p[3]);
      f[i] = __int_as_float(__float_as_int(f[i]) + 1);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = p[0] + p[1] + p[2] + p[3];
}
template <typename F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int blocks = sms * 8, threads = 256;
  double* dout; cudaMalloc(&dout, blocks * threads * 8);
  const double warps = (double)blocks * threads / 32;
  auto report = [&](const char* name, float ms, double units, const char* what) {
    printf("%-44s %8.3f ms  %7.2f %s/clk/SM (at %d MHz nominal)\n", name, ms, units * 32 / (ms * 1e-3) / sms / (clk * 1e3), what, clk / 1000);
  };
  report("DFMA", timeit([&] { k_dfma<<<blocks, threads>>>(dout, 1.0000001, 1e-9); }), warps * 8 * ITER, "FMA");
  report("F2F.F64.F32 alone", timeit([&] { k_f2d<<<blocks, threads>>>(dout, 1.f); }), warps * 8 * ITER, "conversions");
  report("F2F.F32.F64 alone", timeit([&] { k_d2f<<<blocks, threads>>>(dout, 1.0); }), warps * 8 * ITER, "conversions");
  report("pass 1: F2F + 4 DFMA per sample", timeit([&] { k_pass1<0><<<blocks, threads>>>(dout, 1.1, 1.2, 1.3, 1.4); }), warps * 8 * ITER, "samples");
  report("pass 1: ALU widening + 4 DFMA per sample", timeit([&] { k_pass1<1><<<blocks, threads>>>(dout, 1.1, 1.2, 1.3, 1.4); }), warps * 8 * ITER, "samples");
  report("pass 1: 4 DFMA per sample, no conversion", timeit([&] { k_pass1<2><<<blocks, threads>>>(dout, 1.1, 1.2, 1.3, 1.4); }), warps * 8 * ITER, "samples");
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
