// Shared device helpers for the PCG/ECG conditioning kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "these kernels are written for sm_100a (B200) only"
#endif

#include "../../include/mpcg_b200.h"   // return codes + the extern "C" prototypes

#define MPCG_LAUNCH_CHECK()                          \
  do {                                               \
    cudaError_t _e = cudaGetLastError();             \
    if (_e != cudaSuccess) return (int)_e;           \
  } while (0)

namespace mpcg {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float ld_stream(const float* p) {   // read-once data: keep it out of L1
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream4(float4* p, float4 v) {  // write-once data: evict first
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream2(float2* p, float2 v) {
  asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream(float* p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// ---- warp reductions ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
// (largest value, smallest index among equals) -- numpy/torch "first arg-max" rule.
__device__ __forceinline__ void warp_argmax_first(float& v, int& i) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    float ov = __shfl_xor_sync(kFull, v, o);
    int oi = __shfl_xor_sync(kFull, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}

// ---- block reductions (scratch: >= 32 elements of the given type, reused between calls) -------
// All threads must call; every thread receives the result.
template <int THREADS>
__device__ __forceinline__ float block_max(float v, float* scratch) {
  constexpr int NW = THREADS / kWarp;
  v = warp_max(v);
  __syncthreads();                       // scratch may still be read from a previous call
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = scratch[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) r = fmaxf(r, scratch[w]);
  return r;
}
template <int THREADS>
__device__ __forceinline__ float block_min(float v, float* scratch) {
  constexpr int NW = THREADS / kWarp;
  v = warp_min(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = scratch[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) r = fminf(r, scratch[w]);
  return r;
}
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  constexpr int NW = THREADS / kWarp;
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
#pragma unroll
  for (int w = 0; w < NW; ++w) r += scratch[w];     // fixed order: bit-reproducible run to run
  return r;
}
template <int THREADS>
__device__ __forceinline__ int block_max_i(int v, int* scratch) {
  constexpr int NW = THREADS / kWarp;
  v = warp_max_i(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  int r = scratch[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) r = max(r, scratch[w]);
  return r;
}
template <int THREADS>
__device__ __forceinline__ int block_min_i(int v, int* scratch) {
  constexpr int NW = THREADS / kWarp;
  v = warp_min_i(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  int r = scratch[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) r = min(r, scratch[w]);
  return r;
}
template <int THREADS>
__device__ __forceinline__ void block_argmax_first(float& v, int& i, float* fscratch, int* iscratch) {
  constexpr int NW = THREADS / kWarp;
  warp_argmax_first(v, i);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { fscratch[threadIdx.x >> 5] = v; iscratch[threadIdx.x >> 5] = i; }
  __syncthreads();
  v = fscratch[0]; i = iscratch[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) {
    float ov = fscratch[w]; int oi = iscratch[w];
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}

// ---- cooperative global <-> shared row copies ---------------------------------------------------
// The shared destination is "phase matched": element k of the row lives at sh[k], and the caller
// guarantees  ((uintptr_t)(g) & 15) == ((uintptr_t)(sh) & 15)  so 16-byte vectors line up on both
// sides.  Elements outside [0, n) are not touched.
// STREAM = true uses the non-coherent read-once path; pass false when the kernel also writes g.
template <int THREADS, bool STREAM = true>
__device__ __forceinline__ void copy_g2s(float* sh, const float* g, int n) {
  const int tid = threadIdx.x;
  int head = (int)(((16u - ((uintptr_t)g & 15u)) & 15u) >> 2);
  if (head > n) head = n;
  if (tid < head) sh[tid] = STREAM ? ld_stream(g + tid) : g[tid];
  const int nvec = (n - head) >> 2;
  const float4* gv = reinterpret_cast<const float4*>(g + head);
  float4* sv = reinterpret_cast<float4*>(sh + head);
  for (int i = tid; i < nvec; i += THREADS) sv[i] = STREAM ? ld_stream4(gv + i) : gv[i];
  const int done = head + (nvec << 2);
  if (tid < n - done) sh[done + tid] = STREAM ? ld_stream(g + done + tid) : g[done + tid];
}
template <int THREADS>
__device__ __forceinline__ void copy_s2g(float* g, const float* sh, int n) {
  const int tid = threadIdx.x;
  int head = (int)(((16u - ((uintptr_t)g & 15u)) & 15u) >> 2);
  if (head > n) head = n;
  if (tid < head) g[tid] = sh[tid];
  const int nvec = (n - head) >> 2;
  float4* gv = reinterpret_cast<float4*>(g + head);
  const float4* sv = reinterpret_cast<const float4*>(sh + head);
  for (int i = tid; i < nvec; i += THREADS) gv[i] = sv[i];
  const int done = head + (nvec << 2);
  if (tid < n - done) g[done + tid] = sh[done + tid];
}
// Offset (in floats, 0..3) that makes a 16-byte-aligned shared buffer phase-match pointer g.
__device__ __forceinline__ int phase_of(const float* g) { return (int)(((uintptr_t)g & 15u) >> 2); }

}  // namespace mpcg
