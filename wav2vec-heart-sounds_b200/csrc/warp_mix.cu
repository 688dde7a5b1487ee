// extern "C" entries: mpcg_time_warp_f32, mpcg_mix_noise_f32 -- the two NumPy-only augmentations whose reference
// arithmetic lives outside the repository (both "parity unpinned"):
//
//  * time stretch (augment/primitives.py:30-34) shells out to the rubberband phase vocoder, which cannot be
//    reproduced; this build DEFINES its time warp as a resampling gather  y[j] = x(j * rate)  with 4-point
//    Catmull-Rom interpolation and clamped edges, output length round(T / rate) (the length rubberband aims for).
//    For the PCG stretch range 1.004-1.006 the pitch shift this implies is 0.5 %.
//  * real-noise mixing (augment/noise_sources.py:33-64 + pipelines.py:59-60): crop a noise record, abs-max
//    normalise it, scale it, add it, abs-max normalise the sum.  The noise bank is device resident (already at the
//    signal rate); record loading / wfdb I/O stays on the host side of the boundary.
#include "common.cuh"

namespace mpcg {

__device__ __forceinline__ float catmull_rom(float p0, float p1, float p2, float p3, float u) {
  const float a = -0.5f * p0 + 1.5f * p1 - 1.5f * p2 + 0.5f * p3;
  const float b = p0 - 2.5f * p1 + 2.f * p2 - 0.5f * p3;
  const float c = -0.5f * p0 + 0.5f * p2;
  return ((a * u + b) * u + c) * u + p1;
}

// Each thread produces four consecutive outputs and writes them with one 128-bit store.
__global__ void __launch_bounds__(256)
time_warp_kernel(const float* __restrict__ x, float* __restrict__ y, long long t, long long n_out, double rate) {
  const long long row = blockIdx.y;
  const float* xr = x + row * t;
  float* yr = y + row * n_out;
  const long long j0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
  if (j0 >= n_out) return;
  float v[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double pos = (double)(j0 + q) * rate;                 // source position in samples
    long long i = (long long)floor(pos);
    const float u = (float)(pos - (double)i);
    auto at = [&](long long k) { k = k < 0 ? 0 : (k > t - 1 ? t - 1 : k); return __ldg(xr + k); };
    v[q] = catmull_rom(at(i - 1), at(i), at(i + 1), at(i + 2), u);
  }
  if (j0 + 3 < n_out && (((uintptr_t)(yr + j0)) & 15u) == 0) {
    st_stream4(reinterpret_cast<float4*>(yr + j0), make_float4(v[0], v[1], v[2], v[3]));
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (j0 + q < n_out) yr[j0 + q] = v[q];
  }
}

constexpr int kMxThreads = 512;
struct MxStat { double sum; float lo, hi; };
__device__ __forceinline__ void mx_finish(MxStat s, long long n, double& mean, double& inv, double* dscr, float* fscr) {
  const double tot = block_sum<kMxThreads>(s.sum, dscr);
  const float lo = block_min<kMxThreads>(s.lo, fscr);
  const float hi = block_max<kMxThreads>(s.hi, fscr);
  mean = tot / (double)n;
  const double peak = fmax((double)hi - mean, mean - (double)lo);
  inv = peak > 0.0 ? 1.0 / peak : 1.0;                          // NumPy abs_max_normalise: divide only if peak > 0
}
__device__ __forceinline__ float mx_norm(float v, double mean, double inv) {
  return fminf(fmaxf((float)(((double)v - mean) * inv), -1.f), 1.f);
}

__global__ void __launch_bounds__(kMxThreads)
mix_noise_kernel(const float* __restrict__ x, const float* __restrict__ bank, float* __restrict__ y, long long t,
                 long long bank_rows, long long bank_len, const long long* __restrict__ src_row,
                 const long long* __restrict__ src_start, const float* __restrict__ scale) {
  __shared__ double dscr[32];
  __shared__ float fscr[32];
  const long long row = blockIdx.x;
  const float* xr = x + row * t;
  // the tables come from the caller: keep the crop inside the bank whatever they hold
  const long long br = min(max(src_row[row], 0LL), bank_rows - 1), bs = min(max(src_start[row], 0LL), bank_len - t);
  const float* nr = bank + br * bank_len + bs;
  const float s = scale[row];
  const int tid = threadIdx.x;
  MxStat sn{0.0, INFINITY, -INFINITY};
  for (long long i = tid; i < t; i += kMxThreads) {
    const float v = nr[i];
    sn.sum += (double)v; sn.lo = fminf(sn.lo, v); sn.hi = fmaxf(sn.hi, v);
  }
  double mn, in_, ms, is_;
  mx_finish(sn, t, mn, in_, dscr, fscr);
  MxStat ss{0.0, INFINITY, -INFINITY};
  for (long long i = tid; i < t; i += kMxThreads) {
    const float v = xr[i] + s * mx_norm(nr[i], mn, in_);
    ss.sum += (double)v; ss.lo = fminf(ss.lo, v); ss.hi = fmaxf(ss.hi, v);
  }
  mx_finish(ss, t, ms, is_, dscr, fscr);
  for (long long i = tid; i < t; i += kMxThreads)
    y[row * t + i] = mx_norm(xr[i] + s * mx_norm(nr[i], mn, in_), ms, is_);
}

// out[r] = sum_c scale_c N(crop_c), optionally normalised again: the arithmetic of noise_sources.pcg_noise / ecg_noise.
__global__ void __launch_bounds__(kMxThreads)
noise_combine_kernel(const float* __restrict__ bank, float* __restrict__ out, long long t, long long bank_rows,
                     long long bank_len, int ncomp, const long long* __restrict__ src_row,
                     const long long* __restrict__ src_start, const float* __restrict__ scale, int normalise_sum) {
  __shared__ double dscr[32];
  __shared__ float fscr[32];
  const long long row = blockIdx.x;
  const int tid = threadIdx.x;
  const float* nr[4];
  float sc[4];
  double mn[4], in_[4];
  for (int c = 0; c < ncomp; ++c) {
    const long long br = min(max(src_row[row * ncomp + c], 0LL), bank_rows - 1);
    const long long bs = min(max(src_start[row * ncomp + c], 0LL), bank_len - t);
    nr[c] = bank + br * bank_len + bs;
    sc[c] = scale[row * ncomp + c];
    mn[c] = 0.0; in_[c] = 1.0;
    if (sc[c] != 0.f) {                                      // (block-uniform) statistics of the crop
      MxStat s{0.0, INFINITY, -INFINITY};
      for (long long i = tid; i < t; i += kMxThreads) {
        const float v = nr[c][i];
        s.sum += (double)v; s.lo = fminf(s.lo, v); s.hi = fmaxf(s.hi, v);
      }
      mx_finish(s, t, mn[c], in_[c], dscr, fscr);
    }
  }
  auto value = [&](long long i) {
    float v = 0.f;
    for (int c = 0; c < ncomp; ++c)
      if (sc[c] != 0.f) v += sc[c] * mx_norm(nr[c][i], mn[c], in_[c]);
    return v;
  };
  if (!normalise_sum) {
    for (long long i = tid; i < t; i += kMxThreads) out[row * t + i] = value(i);
    return;
  }
  MxStat ss{0.0, INFINITY, -INFINITY};
  for (long long i = tid; i < t; i += kMxThreads) {
    const float v = value(i);
    ss.sum += (double)v; ss.lo = fminf(ss.lo, v); ss.hi = fmaxf(ss.hi, v);
  }
  const float amax = fmaxf(fabsf(block_min<kMxThreads>(ss.lo, fscr)), fabsf(block_max<kMxThreads>(ss.hi, fscr)));
  double ms, is_;
  mx_finish(ss, t, ms, is_, dscr, fscr);
  for (long long i = tid; i < t; i += kMxThreads) {
    const float v = value(i);
    out[row * t + i] = amax > 0.f ? mx_norm(v, ms, is_) : v;   // "if np.max(np.abs(combined)) > 0" (noise_sources.py:48)
  }
}

}  // namespace mpcg

extern "C" int mpcg_noise_combine_f32(const float* bank, float* out, int64_t rows, int64_t t, int64_t bank_rows,
                                      int64_t bank_len, int ncomp, const int64_t* src_row, const int64_t* src_start,
                                      const float* scale, int normalise_sum, void* stream) {
  using namespace mpcg;
  if (rows < 0 || t < 0 || bank_rows < 1 || bank_len < t || ncomp < 1 || ncomp > 4) return MPCG_EINVAL;
  if (rows == 0 || t == 0) return MPCG_OK;
  if (!bank || !out || !src_row || !src_start || !scale) return MPCG_EINVAL;
  if (rows > 0x7fffffffLL) return MPCG_ERANGE;
  noise_combine_kernel<<<(unsigned)rows, kMxThreads, 0, (cudaStream_t)stream>>>(
      bank, out, (long long)t, (long long)bank_rows, (long long)bank_len, ncomp, (const long long*)src_row,
      (const long long*)src_start, scale, normalise_sum);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}

extern "C" int mpcg_time_warp_f32(const float* x, float* y, int64_t rows, int64_t t, int64_t n_out, double rate,
                                  void* stream) {
  using namespace mpcg;
  if (rows < 0 || t < 0 || n_out < 0 || !(rate > 0.0)) return MPCG_EINVAL;
  if (rows == 0 || n_out == 0) return MPCG_OK;
  if (t < 1 || !x || !y) return MPCG_EINVAL;
  for (int64_t r0 = 0; r0 < rows; r0 += 65535) {                      // (rows are a grid dimension: blocks of 65 535)
    const int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
    dim3 grid((unsigned)((n_out + 1023) / 1024), (unsigned)nr);
    time_warp_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x + r0 * t, y + r0 * n_out, (long long)t, (long long)n_out, rate);
  }
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}

extern "C" int mpcg_mix_noise_f32(const float* x, const float* bank, float* y, int64_t rows, int64_t t,
                                  int64_t bank_rows, int64_t bank_len, const int64_t* src_row,
                                  const int64_t* src_start, const float* scale, void* stream) {
  using namespace mpcg;
  if (rows < 0 || t < 0 || bank_rows < 1 || bank_len < t) return MPCG_EINVAL;
  if (rows == 0 || t == 0) return MPCG_OK;
  if (!x || !bank || !y || !src_row || !src_start || !scale) return MPCG_EINVAL;
  if (rows > 0x7fffffffLL) return MPCG_ERANGE;
  mix_noise_kernel<<<(unsigned)rows, kMxThreads, 0, (cudaStream_t)stream>>>(
      x, bank, y, (long long)t, (long long)bank_rows, (long long)bank_len, (const long long*)src_row, (const long long*)src_start,
      scale);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}
