// extern "C" entry: mpcg_fill_nans_f32 -- bridge NaN runs of every row by linear interpolation between the nearest
// valid samples, holding the first / last valid value at the edges (reference signalproc/normalize.py:11-17,
// interpolate_nans = np.interp over the valid indices in float64; called first thing by the NumPy chains,
// signalproc/preprocess.py:25,34, and again by normalize.abs_max_normalise).  A row without a valid sample is left as
// it is, as the reference leaves it.  In place; rows without NaN are only read.
#include "common.cuh"

namespace mpcg {

constexpr int kNfThreads = 256;

__global__ void __launch_bounds__(kNfThreads)
fill_nans_rows_kernel(float* __restrict__ x, long long pitch, int t_uniform, const int* __restrict__ row_len, int channels) {
  __shared__ int lastv[kNfThreads], firstv[kNfThreads];
  __shared__ int any_nan;
  const int tid = threadIdx.x;
  const long long row = blockIdx.x;
  const int t = row_len ? row_len[row / channels] : t_uniform;
  float* xr = x + row * pitch;
  if (tid == 0) any_nan = 0;
  __syncthreads();
  const int chunk = (t + kNfThreads - 1) / kNfThreads;
  const int lo = min(tid * chunk, t), hi = min(lo + chunk, t);
  int fv = 0x7fffffff, lv = -1, bad = 0;
  for (int i = lo; i < hi; ++i) {
    const float v = xr[i];
    if (v == v) { if (fv == 0x7fffffff) fv = i; lv = i; } else bad = 1;
  }
  if (bad) any_nan = 1;
  lastv[tid] = lv;
  firstv[tid] = fv;
  __syncthreads();
  if (!any_nan) return;
  // last valid index before my chunk / first valid index after it (inclusive scans, then shifted by one thread)
  for (int d = 1; d < kNfThreads; d <<= 1) {
    const int a = tid >= d ? lastv[tid - d] : -1;
    const int b = tid + d < kNfThreads ? firstv[tid + d] : 0x7fffffff;
    __syncthreads();
    lastv[tid] = max(lastv[tid], a);
    firstv[tid] = min(firstv[tid], b);
    __syncthreads();
  }
  const int prev0 = tid > 0 ? lastv[tid - 1] : -1;
  const int next0 = tid + 1 < kNfThreads ? firstv[tid + 1] : 0x7fffffff;
  if (lastv[kNfThreads - 1] < 0) return;                    // no valid sample at all
  int p = prev0;
  int i = lo;
  while (i < hi) {
    const float v = xr[i];
    if (v == v) { p = i; ++i; continue; }
    int j = i + 1;                                          // the NaN run [i, j) inside my chunk
    while (j < hi && !(xr[j] == xr[j])) ++j;
    const int n = j < hi ? j : next0;                       // next valid index (0x7fffffff: none)
    const double fp = p >= 0 ? (double)xr[p] : 0.0, fn = n != 0x7fffffff ? (double)xr[n] : 0.0;
    for (int k = i; k < j; ++k) {
      double r;
      if (p < 0) r = fn;                                    // left of the first valid sample
      else if (n == 0x7fffffff) r = fp;                     // right of the last one
      else r = (fn - fp) / (double)(n - p) * (double)(k - p) + fp;
      xr[k] = (float)r;
    }
    i = j;
  }
}

}  // namespace mpcg

extern "C" int mpcg_fill_nans_f32(float* x, int64_t recordings, int channels, int64_t t, const int32_t* row_len,
                                  void* stream) {
  using namespace mpcg;
  if (recordings < 0 || channels < 1 || t < 0) return MPCG_EINVAL;
  if (recordings == 0 || t == 0) return MPCG_OK;
  if (!x) return MPCG_EINVAL;
  if (t > 0x3fffffff || recordings * channels > 0x7fffffffLL) return MPCG_ERANGE;
  fill_nans_rows_kernel<<<(unsigned)(recordings * channels), kNfThreads, 0, (cudaStream_t)stream>>>(x, (long long)t, (int)t,
                                                                                                  row_len, channels);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}
