// extern "C" entry: mpcg_absmax_norm_f32 -- row-wise clip((x - mean) / max|x - mean|, -1, 1)
// (signalproc/torchproc.py:62-66, augment/torchaug.py:24-27, signalproc/normalize.py:20-30).
//
// One CTA per row, two sweeps: (sum, min, max) then the affine map.  max|x - mean| is attained at the
// row minimum or maximum, so  peak = max(max - mean, mean - min)  needs no third sweep.  The second sweep
// re-reads the row from L2 (a 30 s recording at 16 kHz is 1.9 MB; 126 MB of L2 holds a full wave of them).
// Mean and peak are fp64; samples are mapped in fp64 and rounded once.
#include "common.cuh"

namespace mpcg {

constexpr int kNmThreads = 512;

__device__ __forceinline__ float nan_to_num_f(float v) {
  if (v != v) return 0.f;
  if (v == INFINITY) return FLT_MAX;
  if (v == -INFINITY) return -FLT_MAX;
  return v;
}

template <bool FIX_NAN>
__device__ __forceinline__ void norm_stats(const float* __restrict__ xr, long long t, double& mean, double& inv_peak,
                                           int flags, double* dscr, float* fscr) {
  const int tid = threadIdx.x;
  double s = 0.0;
  float lo = INFINITY, hi = -INFINITY;
  const int head = (int)min((long long)(((16u - ((uintptr_t)xr & 15u)) & 15u) >> 2), t);
  if (tid < head) {
    float v = xr[tid];
    if (FIX_NAN) v = nan_to_num_f(v);
    s += (double)v; lo = fminf(lo, v); hi = fmaxf(hi, v);
  }
  const long long nvec = (t - head) >> 2;
  const float4* xv = reinterpret_cast<const float4*>(xr + head);
  for (long long i = tid; i < nvec; i += kNmThreads) {
    float4 q = xv[i];
    if (FIX_NAN) { q.x = nan_to_num_f(q.x); q.y = nan_to_num_f(q.y); q.z = nan_to_num_f(q.z); q.w = nan_to_num_f(q.w); }
    s += ((double)q.x + (double)q.y) + ((double)q.z + (double)q.w);
    lo = fminf(fminf(lo, q.x), fminf(q.y, fminf(q.z, q.w)));
    hi = fmaxf(fmaxf(hi, q.x), fmaxf(q.y, fmaxf(q.z, q.w)));
  }
  const long long done = head + (nvec << 2);
  if (tid < t - done) {
    float v = xr[done + tid];
    if (FIX_NAN) v = nan_to_num_f(v);
    s += (double)v; lo = fminf(lo, v); hi = fmaxf(hi, v);
  }
  s = block_sum<kNmThreads>(s, dscr);
  lo = block_min<kNmThreads>(lo, fscr);
  hi = block_max<kNmThreads>(hi, fscr);
  mean = s / (double)t;
  double peak = fmax((double)hi - mean, mean - (double)lo);
  if (flags & MPCG_NORM_PEAK_GT0) {
    inv_peak = (peak > 0.0) ? 1.0 / peak : 1.0;
  } else {
    inv_peak = 1.0 / fmax(peak, 1e-12);
  }
}

__device__ __forceinline__ float norm_map(float v, double mean, double inv_peak) {
  const double u = ((double)v - mean) * inv_peak;
  return (float)fmin(fmax(u, -1.0), 1.0);
}

template <bool FIX_NAN>
__global__ void __launch_bounds__(kNmThreads)
absmax_norm_rows_kernel(const float* __restrict__ x, float* __restrict__ y, long long t, int flags) {
  __shared__ double dscr[32];
  __shared__ float fscr[32];
  const long long row = blockIdx.x;
  const float* xr = x + row * t;
  float* yr = y + row * t;
  double mean, inv_peak;
  norm_stats<FIX_NAN>(xr, t, mean, inv_peak, flags, dscr, fscr);
  const int tid = threadIdx.x;
  const bool vec_ok = (((uintptr_t)xr ^ (uintptr_t)yr) & 15u) == 0;
  if (vec_ok) {
    const int head = (int)min((long long)(((16u - ((uintptr_t)xr & 15u)) & 15u) >> 2), t);
    if (tid < head) {
      float v = xr[tid];
      if (FIX_NAN) v = nan_to_num_f(v);
      yr[tid] = norm_map(v, mean, inv_peak);
    }
    const long long nvec = (t - head) >> 2;
    const float4* xv = reinterpret_cast<const float4*>(xr + head);
    float4* yv = reinterpret_cast<float4*>(yr + head);
    for (long long i = tid; i < nvec; i += kNmThreads) {
      float4 q = xv[i];
      if (FIX_NAN) { q.x = nan_to_num_f(q.x); q.y = nan_to_num_f(q.y); q.z = nan_to_num_f(q.z); q.w = nan_to_num_f(q.w); }
      q.x = norm_map(q.x, mean, inv_peak); q.y = norm_map(q.y, mean, inv_peak);
      q.z = norm_map(q.z, mean, inv_peak); q.w = norm_map(q.w, mean, inv_peak);
      st_stream4(yv + i, q);
    }
    const long long done = head + (nvec << 2);
    if (tid < t - done) {
      float v = xr[done + tid];
      if (FIX_NAN) v = nan_to_num_f(v);
      yr[done + tid] = norm_map(v, mean, inv_peak);
    }
  } else {
    for (long long i = tid; i < t; i += kNmThreads) {
      float v = xr[i];
      if (FIX_NAN) v = nan_to_num_f(v);
      yr[i] = norm_map(v, mean, inv_peak);
    }
  }
}

}  // namespace mpcg

extern "C" int mpcg_absmax_norm_f32(const float* x, float* y, int64_t rows, int64_t t, int flags, void* stream) {
  using namespace mpcg;
  if (rows < 0 || t < 0) return MPCG_EINVAL;
  if (rows == 0 || t == 0) return MPCG_OK;
  if (!x || !y) return MPCG_EINVAL;
  if (rows > 0x7fffffffLL) return MPCG_ERANGE;
  if (flags & MPCG_NORM_NAN_TO_NUM)
    absmax_norm_rows_kernel<true><<<(unsigned)rows, kNmThreads, 0, (cudaStream_t)stream>>>(x, y, (long long)t, flags);
  else
    absmax_norm_rows_kernel<false><<<(unsigned)rows, kNmThreads, 0, (cudaStream_t)stream>>>(x, y, (long long)t, flags);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}
