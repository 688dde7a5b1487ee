// fp64 tensor-core (DMMA) issue rate on B200 against the DFMA pipe: decides whether pass 1 of the chunked IIR scan
// (a [4 x L] x [L x chunks] product) is worth moving to mma.sync.m8n8k4.f64.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_dmma tools/ubench_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
__global__ void k_dmma884(double* out, double a, double b) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; }
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma16816(double* out, double a, double b) {
  double c[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x + i + j;
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%4,%4,%4,%4,%4,%4,%4}, {%5,%5,%5,%5}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b));
  double s = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
// DMMA and DFMA together: do they share a pipe?
__global__ void k_both(double* out, double a, double b) {
  double c[4][2], v[8];
  for (int i = 0; i < 4; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; }
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fma(v[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  for (int i = 0; i < 8; ++i) s += v[i];
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
  double* dout; cudaMalloc(&dout, blocks * threads * 8);
  const double warps = (double)blocks * threads / 32;
  auto report = [&](const char* name, float ms, double warp_instr, double fma_per_instr) {
    printf("%-34s %8.3f ms  %7.2f warp-instr/clk/SM  %7.1f FMA/clk/SM (at %d MHz nominal)\n", name, ms,
           warp_instr / (ms * 1e-3) / sms / (clk * 1e3), warp_instr * fma_per_instr / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  };
  report("DFMA", timeit([&] { k_dfma<<<blocks, threads>>>(dout, 1.0000001, 1e-9); }), warps * 8 * ITER, 32);
  report("DMMA m8n8k4", timeit([&] { k_dmma884<<<blocks, threads>>>(dout, 1.0000001, 1e-9); }), warps * 8 * ITER, 256);
  report("DMMA m16n8k16", timeit([&] { k_dmma16816<<<blocks, threads>>>(dout, 1.0000001, 1e-9); }), warps * 4 * ITER, 2048);
  report("4 DMMA m8n8k4 + 8 DFMA (per 12)", timeit([&] { k_both<<<blocks, threads>>>(dout, 1.0000001, 1e-9); }), warps * 12 * ITER, 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
