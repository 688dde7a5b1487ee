// extern "C" entry: mpcg_cycle_rebuild_f32 -- rebuild a training signal from rearranged cardiac cycles
// (reference datasets/heart_cycles.py:38-69: _crossfade + rebuild).
//
// The host decides the cycle order (heart_cycles.py:72-98 draws it from a Python RNG) and hands every row its cycles as
// (start, length) pairs into the source row; the kernel replays the reference's loop
//   out = cycle[0];  while len(out) < target: out = crossfade(out, cycle[i % K], n)
// with one CTA per row.  A join blends the last n samples of what has been built so far with the first n of the next
// cycle: linear ramp when either side is (nearly) flat (variance < 1e-5), otherwise the correlation-aware pair
//   skew = 9/16 sin(pi t / 2) + 1/16 sin(3 pi t / 2),  even = sqrt(max(0.5 / (1 + r) - (1 - r) / (1 + r) skew^2, 0)),
//   fade_in = clip(even + skew, 0, 1),  r = |corrcoef(tail, head)|,  t = linspace(-1, 1, n)
// Joins are sequential (a short cycle's blended samples are the next join's tail), rows are independent.  Statistics
// in fp64 with the two-pass variance NumPy uses.
#include "common.cuh"

namespace mpcg {

constexpr int kCyThreads = 256;

__global__ void __launch_bounds__(kCyThreads)
cycle_rebuild_kernel(const float* __restrict__ x, float* __restrict__ y, long long* __restrict__ out_len,
                     const int* __restrict__ starts, const int* __restrict__ lens, const int* __restrict__ counts,
                     long long t, long long cap, int kmax, long long target, int n) {
  __shared__ double dscr[32];
  const int tid = threadIdx.x;
  const long long row = blockIdx.x;
  const float* xr = x + row * t;
  float* yr = y + row * cap;
  const int k = counts[row];
  const int* st = starts + row * kmax;
  const int* ln = lens + row * kmax;
  if (k <= 0) {                                                    // no usable segmentation: the row passes through
    const long long m = t < cap ? t : cap;
    for (long long i = tid; i < m; i += kCyThreads) yr[i] = xr[i];
    if (tid == 0) out_len[row] = m;
    return;
  }
  long long cur = ln[0] < cap ? ln[0] : cap;
  for (long long i = tid; i < cur; i += kCyThreads) yr[i] = xr[st[0] + i];
  __syncthreads();
  long long step = 1;
  const long long max_steps = 10LL * k + 4;
  while (cur < target && step <= max_steps + 1) {
    const int c = (int)(step % k);
    const float* b = xr + st[c];
    const long long len = ln[c];
    long long skip = 0;
    if (n > 1 && cur >= n && len >= n) {
      const float* a = yr + (cur - n);
      // two-pass statistics of the n-sample tail and head
      double sa = 0.0, sb = 0.0;
      for (int j = tid; j < n; j += kCyThreads) { sa += (double)a[j]; sb += (double)b[j]; }
      sa = block_sum<kCyThreads>(sa, dscr);
      sb = block_sum<kCyThreads>(sb, dscr);
      const double ma = sa / n, mb = sb / n;
      double va = 0.0, vb = 0.0, cab = 0.0;
      for (int j = tid; j < n; j += kCyThreads) {
        const double da = (double)a[j] - ma, db = (double)b[j] - mb;
        va += da * da; vb += db * db; cab += da * db;
      }
      va = block_sum<kCyThreads>(va, dscr);
      vb = block_sum<kCyThreads>(vb, dscr);
      cab = block_sum<kCyThreads>(cab, dscr);
      const bool flat = va / n < 1e-5 || vb / n < 1e-5;
      double r = fabs(cab / sqrt(va * vb));
      if (!(r == r)) r = 0.0;
      r = fmin(r, 1.0);
      const double g0 = 0.5 / (1.0 + r), g1 = (1.0 - r) / (1.0 + r);
      const double dt = 2.0 / (double)(n - 1), du = 1.0 / (double)(n - 1);
      __syncthreads();                                              // every read of the tail precedes its overwrite
      for (int j = tid; j < n; j += kCyThreads) {
        double f;
        if (flat) {
          f = j == n - 1 ? 1.0 : (double)j * du;
        } else {
          const double tt = j == n - 1 ? 1.0 : -1.0 + (double)j * dt;
          const double skew = 0.5625 * sinpi(0.5 * tt) + 0.0625 * sinpi(1.5 * tt);
          const double even = sqrt(fmax(g0 - g1 * skew * skew, 0.0));
          f = fmin(fmax(even + skew, 0.0), 1.0);
        }
        yr[cur - n + j] = (float)((double)a[j] * (1.0 - f) + (double)b[j] * f);
      }
      skip = n;
    }
    const long long room = cap - cur;
    long long add = len - skip;
    if (add > room) add = room;
    for (long long i = tid; i < add; i += kCyThreads) yr[cur + i] = b[skip + i];
    cur += add > 0 ? add : 0;
    __syncthreads();
    ++step;
    if (room <= 0) break;
  }
  if (tid == 0) out_len[row] = cur;
}

}  // namespace mpcg

extern "C" int mpcg_cycle_rebuild_f32(const float* x, float* y, int64_t* out_len, const int32_t* starts, const int32_t* lens,
                                      const int32_t* counts, int64_t rows, int64_t t, int64_t cap, int kmax, int64_t target_len,
                                      int fade_n, void* stream) {
  using namespace mpcg;
  if (rows < 0 || t < 0 || cap < 0 || kmax < 0 || target_len < 0 || fade_n < 0) return MPCG_EINVAL;
  if (rows == 0) return MPCG_OK;
  if (!x || !y || !out_len || !counts || (kmax > 0 && (!starts || !lens))) return MPCG_EINVAL;
  if (rows > 0x7fffffffLL) return MPCG_ERANGE;
  cycle_rebuild_kernel<<<(unsigned)rows, kCyThreads, 0, (cudaStream_t)stream>>>(
      x, y, reinterpret_cast<long long*>(out_len), starts, lens, counts, (long long)t, (long long)cap, kmax, (long long)target_len,
      fade_n);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}
