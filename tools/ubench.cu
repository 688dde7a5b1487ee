// Pipe-rate microbenchmarks that decide the biquad/resample inner-loop design on B200.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench tools/ubench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITER 4096
__global__ void k_dfma(double* out, double a, double b) {
  double v[8];
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fma(v[i], a, b);
  double s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, float a, float b) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], a, b);
  float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// f32 -> f64 convert + one DADD per element (the add keeps the convert alive)
__global__ void k_cvt_up(double* out, const float* in) {
  float f[8];
  for (int i = 0; i < 8; ++i) f[i] = in[threadIdx.x + i];
  double acc[8] = {0};
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] += (double)f[i]; f[i] = __int_as_float(__float_as_int(f[i]) ^ (it & 1)); }
  double s = 0; for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ double up_int(float f) {
  const unsigned b = __float_as_uint(f);
  const unsigned hi = (((b << 1) >> 4) + 0x38000000u) | (b & 0x80000000u);
  return __hiloint2double((int)hi, (int)(b << 29));
}
__global__ void k_cvt_up_int(double* out, const float* in) {
  float f[8];
  for (int i = 0; i < 8; ++i) f[i] = in[threadIdx.x + i];
  double acc[8] = {0};
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] += up_int(f[i]); f[i] = __int_as_float(__float_as_int(f[i]) ^ (it & 1)); }
  double s = 0; for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_cvt_down(float* out, double a) {
  double v[8];
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
  float acc[8] = {0};
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] += (float)v[i]; v[i] += a; }
  float s = 0; for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// DADD alone, to subtract from the cvt kernels
__global__ void k_dadd(double* out, double a) {
  double v[8];
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += a;
  double s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: per element 1 up-convert + 7 DFMA (biquad-like ratio), hardware cvt vs integer cvt
template <bool INT>
__global__ void k_mix(double* out, const float* in, double a, double b) {
  float f[4];
  for (int i = 0; i < 4; ++i) f[i] = in[threadIdx.x + i];
  double v[4] = {1, 2, 3, 4};
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double x = INT ? up_int(f[i]) : (double)f[i];
      f[i] = __int_as_float(__float_as_int(f[i]) ^ (it & 1));
      double y = fma(a, x, v[i]);
      double t = fma(b, x, v[(i + 1) & 3]);
      v[i] = fma(-a, y, t);
      double u = fma(b, y, v[(i + 2) & 3]);
      v[(i + 1) & 3] = fma(-b, y, a * x);
      v[(i + 2) & 3] = fma(u, a, b);
      v[(i + 3) & 3] = fma(v[(i + 3) & 3], a, y);
    }
  double s = 0; for (int i = 0; i < 4; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
  double* dout; float* fout; float* fin;
  cudaMalloc(&dout, blocks * threads * 8); cudaMalloc(&fout, blocks * threads * 4); cudaMalloc(&fin, 4096);
  cudaMemset(fin, 0x3f, 4096);
  const double n = (double)blocks * threads * 8 * ITER;
  auto report = [&](const char* name, float ms, double ops) {
    printf("%-28s %8.3f ms  %8.1f Gop/s  %6.1f op/clk/SM (at %d MHz nominal)\n", name, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  };
  report("DFMA", timeit([&] { k_dfma<<<blocks, threads>>>(dout, 1.0000001, 1e-9); }), n);
  report("FFMA", timeit([&] { k_ffma<<<blocks, threads>>>(fout, 1.0000001f, 1e-9f); }), n);
  report("DADD", timeit([&] { k_dadd<<<blocks, threads>>>(dout, 1e-9); }), n);
  report("F2F.F64.F32 + DADD", timeit([&] { k_cvt_up<<<blocks, threads>>>(dout, fin); }), n);
  report("int-cvt f32->f64 + DADD", timeit([&] { k_cvt_up_int<<<blocks, threads>>>(dout, fin); }), n);
  report("F2F.F32.F64 + FADD + DADD", timeit([&] { k_cvt_down<<<blocks, threads>>>(fout, 1e-9); }), n);
  report("mix hw-cvt (per 8 ops)", timeit([&] { k_mix<false><<<blocks, threads>>>(dout, fin, 0.9, 0.1); }), n / 2);
  report("mix int-cvt (per 8 ops)", timeit([&] { k_mix<true><<<blocks, threads>>>(dout, fin, 0.9, 0.1); }), n / 2);
  return 0;
}
