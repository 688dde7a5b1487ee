// The chunked IIR scan's inner loops in isolation on B200 (2 CTAs x 512 threads per SM, like aug_chain_kernel):
// DFMA dependent-issue latency, pass 1 (conversion + 4 DFMA per sample), pass 2 (conversion, 10 DFMA, conversion back,
// shared-memory store per sample) and the five-level warp scan.  Coefficients are kernel parameters (constant bank).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_iir tools/ubench_iir.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int L = 33, REP = 64;
struct K { double c[2][5]; double wt[L][4]; double mp[5][16]; };
__global__ void k_lat(double* out, double a, double b, int n) {
  double v = threadIdx.x;
  for (int i = 0; i < n; ++i) v = fma(v, a, b);
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
}
__global__ void __launch_bounds__(512, 2) k_pass1(double* out, const __grid_constant__ K k) {
  extern __shared__ float buf[];
  float* mine = buf + threadIdx.x * L;
  for (int j = 0; j < L; ++j) mine[j] = threadIdx.x * 0.01f + j;
  double p[4] = {0, 0, 0, 0};
  for (int r = 0; r < REP; ++r) {
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const double xv = (double)mine[j];
      p[0] = fma(k.wt[j][0], xv, p[0]); p[1] = fma(k.wt[j][1], xv, p[1]);
      p[2] = fma(k.wt[j][2], xv, p[2]); p[3] = fma(k.wt[j][3], xv, p[3]);
    }
    mine[r % L] = (float)p[0];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = p[0] + p[1] + p[2] + p[3];
}
// pass 1 with the chunk split in NSUB equal parts that share ONE weight table (each constant feeds NSUB DFMAs), joined
// with M^(L/NSUB):  p = M (M p_0 + p_1) + p_2
template <int NSUB>
__global__ void __launch_bounds__(512, 2) k_pass1_split(double* out, const __grid_constant__ K k) {
  extern __shared__ float buf[];
  constexpr int LS = L / NSUB;
  float* mine = buf + threadIdx.x * L;
  for (int j = 0; j < L; ++j) mine[j] = threadIdx.x * 0.01f + j;
  double acc = 0;
  for (int r = 0; r < REP; ++r) {
    double p[NSUB][4];
#pragma unroll
    for (int u = 0; u < NSUB; ++u) { p[u][0] = p[u][1] = p[u][2] = p[u][3] = 0; }
#pragma unroll
    for (int j = 0; j < LS; ++j) {
#pragma unroll
      for (int u = 0; u < NSUB; ++u) {
        const double xv = (double)mine[u * LS + j];
        p[u][0] = fma(k.wt[j][0], xv, p[u][0]); p[u][1] = fma(k.wt[j][1], xv, p[u][1]);
        p[u][2] = fma(k.wt[j][2], xv, p[u][2]); p[u][3] = fma(k.wt[j][3], xv, p[u][3]);
      }
    }
    double q[4] = {p[0][0], p[0][1], p[0][2], p[0][3]};
#pragma unroll
    for (int u = 1; u < NSUB; ++u) {
      double n[4] = {p[u][0], p[u][1], p[u][2], p[u][3]};
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int c = 0; c < (rr < 2 ? 2 : 4); ++c) n[rr] = fma(k.mp[0][rr * 4 + c], q[c], n[rr]);
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) q[rr] = n[rr];
    }
    acc += q[0] + q[1] + q[2] + q[3];
    mine[r % L] = (float)q[0];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// pass 1 with the weights in shared memory (two broadcast 128-bit loads per sample)
__global__ void __launch_bounds__(512, 2) k_pass1_smem(double* out, const __grid_constant__ K k) {
  extern __shared__ float buf[];
  __shared__ __align__(16) double wt[L][4];
  for (int i = threadIdx.x; i < L * 4; i += 512) (&wt[0][0])[i] = (&k.wt[0][0])[i];
  float* mine = buf + threadIdx.x * L;
  for (int j = 0; j < L; ++j) mine[j] = threadIdx.x * 0.01f + j;
  __syncthreads();
  double p[4] = {0, 0, 0, 0};
  for (int r = 0; r < REP; ++r) {
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const double xv = (double)mine[j];
      const double2 w01 = *reinterpret_cast<const double2*>(&wt[j][0]);
      const double2 w23 = *reinterpret_cast<const double2*>(&wt[j][2]);
      p[0] = fma(w01.x, xv, p[0]); p[1] = fma(w01.y, xv, p[1]);
      p[2] = fma(w23.x, xv, p[2]); p[3] = fma(w23.y, xv, p[3]);
    }
    mine[r % L] = (float)p[0];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = p[0] + p[1] + p[2] + p[3];
}
// pass 1 as the zero-state recurrence itself (10 DFMA per sample, ten constants, no output)
__global__ void __launch_bounds__(512, 2) k_pass1_rec(double* out, const __grid_constant__ K k) {
  extern __shared__ float buf[];
  float* mine = buf + threadIdx.x * L;
  for (int j = 0; j < L; ++j) mine[j] = threadIdx.x * 0.01f + j;
  double z[4] = {0, 0, 0, 0};
  for (int r = 0; r < REP; ++r) {
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const double xv = (double)mine[j];
      const double y0 = fma(k.c[0][0], xv, z[0]);
      z[0] = fma(-k.c[0][3], y0, fma(k.c[0][1], xv, z[1]));
      z[1] = fma(-k.c[0][4], y0, k.c[0][2] * xv);
      const double y1 = fma(k.c[1][0], y0, z[2]);
      z[2] = fma(-k.c[1][3], y1, fma(k.c[1][1], y0, z[3]));
      z[3] = fma(-k.c[1][4], y1, k.c[1][2] * y0);
    }
    mine[r % L] = (float)z[0];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = z[0] + z[1] + z[2] + z[3];
}
// pass 2 of group g fused with pass 1 of group g + 1 (weights of g + 1 from shared memory)
__global__ void __launch_bounds__(512, 2) k_pass2_fused(double* out, const __grid_constant__ K k) {
  extern __shared__ float buf[];
  __shared__ __align__(16) double wt[L][4];
  for (int i = threadIdx.x; i < L * 4; i += 512) (&wt[0][0])[i] = (&k.wt[0][0])[i];
  float* mine = buf + threadIdx.x * L;
  for (int j = 0; j < L; ++j) mine[j] = threadIdx.x * 0.01f + j;
  __syncthreads();
  double z[4] = {0, 0, 0, 0}, p[4] = {0, 0, 0, 0};
  for (int r = 0; r < REP; ++r) {
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const double xv = (double)mine[j];
      const double y0 = fma(k.c[0][0], xv, z[0]);
      z[0] = fma(-k.c[0][3], y0, fma(k.c[0][1], xv, z[1]));
      z[1] = fma(-k.c[0][4], y0, k.c[0][2] * xv);
      const double y1 = fma(k.c[1][0], y0, z[2]);
      z[2] = fma(-k.c[1][3], y1, fma(k.c[1][1], y0, z[3]));
      z[3] = fma(-k.c[1][4], y1, k.c[1][2] * y0);
      const double2 w01 = *reinterpret_cast<const double2*>(&wt[j][0]);
      const double2 w23 = *reinterpret_cast<const double2*>(&wt[j][2]);
      p[0] = fma(w01.x, y1, p[0]); p[1] = fma(w01.y, y1, p[1]);
      p[2] = fma(w23.x, y1, p[2]); p[3] = fma(w23.y, y1, p[3]);
      mine[j] = (float)y1;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = z[0] + z[1] + z[2] + z[3] + p[0] + p[1] + p[2] + p[3];
}
template <bool FROM_REGS>
__global__ void __launch_bounds__(512, 2) k_pass2(double* out, const __grid_constant__ K k) {
  extern __shared__ float buf[];
  float* mine = buf + threadIdx.x * L;
  float v[L];
  for (int j = 0; j < L; ++j) { mine[j] = threadIdx.x * 0.01f + j; v[j] = mine[j]; }
  double z[4] = {0, 0, 0, 0};
  for (int r = 0; r < REP; ++r) {
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const double xv = (double)(FROM_REGS ? v[j] : mine[j]);
      const double y0 = fma(k.c[0][0], xv, z[0]);
      z[0] = fma(-k.c[0][3], y0, fma(k.c[0][1], xv, z[1]));
      z[1] = fma(-k.c[0][4], y0, k.c[0][2] * xv);
      const double y1 = fma(k.c[1][0], y0, z[2]);
      z[2] = fma(-k.c[1][3], y1, fma(k.c[1][1], y0, z[3]));
      z[3] = fma(-k.c[1][4], y1, k.c[1][2] * y0);
      mine[j] = (float)y1;
    }
    if (FROM_REGS) v[r % L] += 1.f;
  }
  float s = 0; for (int j = 0; j < L; ++j) s += v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = z[0] + z[1] + z[2] + z[3] + s;
}
__global__ void __launch_bounds__(512, 2) k_scan(double* out, const __grid_constant__ K k) {
  const int lane = threadIdx.x & 31;
  double p[4] = {threadIdx.x * 1e-3, 1, 2, 3};
  for (int r = 0; r < REP; ++r) {
    double u[4];
#pragma unroll
    for (int d = 0; d < 5; ++d) {
#pragma unroll
      for (int s = 0; s < 4; ++s) u[s] = __shfl_up_sync(0xffffffffu, p[s], 1 << d);
      if (lane >= (1 << d)) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          double a = p[rr];
#pragma unroll
          for (int c = 0; c < (rr < 2 ? 2 : 4); ++c) a = fma(k.mp[d][rr * 4 + c], u[c], a);
          p[rr] = a;
        }
      }
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
  double* dout; cudaMalloc(&dout, sms * 2 * 512 * 8);
  K k;
  for (int i = 0; i < 10; ++i) (&k.c[0][0])[i] = 0.1 + 0.01 * i;
  for (int i = 0; i < L * 4; ++i) (&k.wt[0][0])[i] = 0.5 / (1 + i);
  for (int i = 0; i < 80; ++i) (&k.mp[0][0])[i] = 0.01 * (i % 7);
  const size_t smem = 512 * L * 4;
  cudaFuncSetAttribute(k_pass1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_pass2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_pass2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const double cyc = clk * 1e3;
  { const int n = 1 << 16; float ms = timeit([&] { k_lat<<<1, 32>>>(dout, 1.0000001, 1e-9, n); });
    printf("DFMA dependent chain, one warp: %.1f cycles per DFMA\n", ms * 1e-3 * cyc / n); }
  const double samples = (double)sms * 2 * 512 * L * REP;
  auto rep = [&](const char* name, float ms, double fma_per_sample) {
    const double spc = samples / (ms * 1e-3 * cyc) / sms;
    printf("%-46s %7.3f ms  %6.2f samples/clk/SM  (%5.1f DFMA/clk/SM)\n", name, ms, spc, spc * fma_per_sample);
  };
  rep("pass 1 (LDS, F2F, 4 DFMA)", timeit([&] { k_pass1<<<sms * 2, 512, smem>>>(dout, k); }), 4);
  cudaFuncSetAttribute(k_pass1_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_pass1_rec, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_pass2_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_pass1_split<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_pass1_split<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rep("pass 1 in 3 parts sharing one weight table", timeit([&] { k_pass1_split<3><<<sms * 2, 512, smem>>>(dout, k); }), 4);
  rep("pass 1 in 11 parts sharing one weight table", timeit([&] { k_pass1_split<11><<<sms * 2, 512, smem>>>(dout, k); }), 4);
  rep("pass 1, weights from shared memory", timeit([&] { k_pass1_smem<<<sms * 2, 512, smem>>>(dout, k); }), 4);
  rep("pass 1 as the zero-state recurrence (10 DFMA)", timeit([&] { k_pass1_rec<<<sms * 2, 512, smem>>>(dout, k); }), 10);
  rep("pass 2 + next group's pass 1 (14 DFMA)", timeit([&] { k_pass2_fused<<<sms * 2, 512, smem>>>(dout, k); }), 14);
  rep("pass 2 from registers (F2F, 10 DFMA, F2F, STS)", timeit([&] { k_pass2<true><<<sms * 2, 512, smem>>>(dout, k); }), 10);
  rep("pass 2 from shared (LDS, F2F, 10 DFMA, F2F, STS)", timeit([&] { k_pass2<false><<<sms * 2, 512, smem>>>(dout, k); }), 10);
  { float ms = timeit([&] { k_scan<<<sms * 2, 512>>>(dout, k); });
    printf("%-46s %7.3f ms  %6.0f cycles per 5-level scan (32 warps per SM in flight)\n", "warp scan (40 SHFL + 60 DFMA)", ms, ms * 1e-3 * cyc / REP); }
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
