// extern "C" entry: mpcg_row_normalise_f32 -- the other amplitude normalisers of signalproc/normalize.py on device rows
// (SURVEY 8f rank 3): min-max (normalize.py:33-44), z-score (:47-56) and k-peak (:59-78, the mean of the k largest /
// k smallest samples as the range).  One CTA per row; the row is swept from L2:
//   sweep A  sum (fp64), min, max                       [+ k-peak: histogram of the top 11 key bits]
//   sweep B  sum of squared deviations (z-score)        [+ k-peak: next 11 bits inside the two boundary bins, fp64 sum
//                                                          of everything strictly beyond them]
//   sweep C  k-peak only: last 10 bits inside the boundary bins -> the exact k-th largest / smallest key; the values
//            between it and the boundary are reconstructed from (key, count), so no further sum is needed
//   sweep D  y = a + (x - b) * s  (fp64 per sample, rounded once)
// Keys are the usual order-preserving integer image of a float, so the selection is exact whatever the data; ties at
// the k-th value are counted, not guessed.  Scope "global" reproduces what the reference's tensor functions do with a
// batched input (ONE min / max, or the mean over all rows' top-k values, for the whole tensor): sweep A-C write per-row
// statistics, a one-CTA kernel combines them and a third kernel applies the map.
#include "common.cuh"

namespace mpcg {

constexpr int kRnThreads = 512;
constexpr int kRnBins = 2048;
constexpr double kRnEps = 1e-8;                                    // normalize.py:8

struct RnShared {
  int hist_hi[kRnBins];
  int hist_lo[kRnBins];
  double dscr[32];
  float fscr[32];
  int sel[4];                                                      // bin, remaining (hi); bin, remaining (lo)
};

__device__ __forceinline__ unsigned rn_key(float v) {
  const unsigned b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float rn_val(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Warp 0: the bin holding the k-th element counted from the top (from_top) or from the bottom of `hist`, and how many
// elements of that bin are still needed.  Lane l owns nbins/32 consecutive bins in counting order.
__device__ __forceinline__ void rn_select(const int* hist, int nbins, long long k, bool from_top, int* out) {
  const int lane = threadIdx.x & 31;
  const int per = nbins >> 5;
  long long tot = 0;
  for (int j = 0; j < per; ++j) {
    const int pos = lane * per + j;
    tot += hist[from_top ? nbins - 1 - pos : pos];
  }
  long long inc = tot;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const long long o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  const long long before = inc - tot;
  if (before < k && inc >= k) {
    long long need = k - before;
    for (int j = 0; j < per; ++j) {
      const int pos = lane * per + j;
      const int bin = from_top ? nbins - 1 - pos : pos;
      const int c = hist[bin];
      if (need <= c) { out[0] = bin; out[1] = (int)need; break; }
      need -= c;
    }
  }
}

struct RnStats { double mn, mx, mean, sd, hi_ref, lo_ref; };

// Map coefficients of one row (or of the whole tensor): y = isconst ? a : a + (x - b) * s
__device__ __forceinline__ void rn_coeffs(int mode, int flags, double lo, double hi, const RnStats& st,
                                          double& a, double& b, double& s, bool& isconst) {
  isconst = false;
  if (mode == MPCG_RN_ZSCORE) { a = 0.0; b = st.mean; s = 1.0 / (st.sd + kRnEps); return; }
  const double r0 = mode == MPCG_RN_MINMAX ? st.mn : st.lo_ref;
  const double r1 = mode == MPCG_RN_MINMAX ? st.mx : st.hi_ref;
  const double span = r1 - r0;
  a = lo; b = r0;
  if (flags & MPCG_RN_EPS) { s = (hi - lo) / (span + kRnEps); return; }
  if (span <= 0.0) { isconst = true; a = (lo + hi) * 0.5; s = 0.0; return; }      // normalize.py:37-38, 70-71
  s = (hi - lo) / span;
}

// fn(x[i], c) for every i < t, the 16-byte aligned middle of the row as float4; c in 0..3 is the sample's slot in its
// vector (a compile-time constant at each call site: callers keep one accumulator per slot to break the fp64 add chain)
template <typename F>
__device__ __forceinline__ void rn_sweep(const float* __restrict__ xr, long long t, int tid, int stride, F fn) {
  const int head = (int)min((long long)(((16u - ((uintptr_t)xr & 15u)) & 15u) >> 2), t);
  if (tid < head) fn(xr[tid], 0);
  const long long nvec = (t - head) >> 2;
  const float4* xv = reinterpret_cast<const float4*>(xr + head);
  for (long long i = tid; i < nvec; i += stride) {
    const float4 q = xv[i];
    fn(q.x, 0); fn(q.y, 1); fn(q.z, 2); fn(q.w, 3);
  }
  const long long done = head + (nvec << 2);
  if (tid < t - done) fn(xr[done + tid], 0);
}

__device__ __forceinline__ void rn_apply(const float* __restrict__ xr, float* __restrict__ yr, long long t, double a, double b,
                                         double s, bool isconst, int tid, int stride) {
  if (isconst) {
    const float c = (float)a;
    for (long long i = tid; i < t; i += stride) yr[i] = c;
    return;
  }
  auto map = [&](float v) { return (float)(a + ((double)v - b) * s); };
  if ((((uintptr_t)xr ^ (uintptr_t)yr) & 15u) != 0) {
    for (long long i = tid; i < t; i += stride) yr[i] = map(xr[i]);
    return;
  }
  const int head = (int)min((long long)(((16u - ((uintptr_t)xr & 15u)) & 15u) >> 2), t);
  if (tid < head) yr[tid] = map(xr[tid]);
  const long long nvec = (t - head) >> 2;
  const float4* xv = reinterpret_cast<const float4*>(xr + head);
  float4* yv = reinterpret_cast<float4*>(yr + head);
  for (long long i = tid; i < nvec; i += stride) {
    float4 q = xv[i];
    q.x = map(q.x); q.y = map(q.y); q.z = map(q.z); q.w = map(q.w);
    st_stream4(yv + i, q);
  }
  const long long done = head + (nvec << 2);
  if (tid < t - done) yr[done + tid] = map(xr[done + tid]);
}

template <bool KPEAK>
__global__ void __launch_bounds__(kRnThreads, KPEAK ? 1 : 2)
row_normalise_kernel(const float* __restrict__ x, float* __restrict__ y, double* __restrict__ stats, long long t, int mode, int k,
                     double lo, double hi, int flags, int apply) {
  __shared__ RnShared sm;
  const int tid = threadIdx.x;
  const long long row = blockIdx.x;
  const float* xr = x + row * t;
  constexpr bool kpeak = KPEAK;
  if (kpeak) {
    for (int i = tid; i < kRnBins; i += kRnThreads) { sm.hist_hi[i] = 0; sm.hist_lo[i] = 0; }
    __syncthreads();
  }
  // ---- sweep A
  double s4[4] = {0.0, 0.0, 0.0, 0.0};
  float mn = INFINITY, mx = -INFINITY;
  rn_sweep(xr, t, tid, kRnThreads, [&](float v, int c) {
    s4[c] += (double)v; mn = fminf(mn, v); mx = fmaxf(mx, v);
    if (kpeak) atomicAdd(&sm.hist_hi[rn_key(v) >> 21], 1);
  });
  const double s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
  RnStats st;
  st.mean = block_sum<kRnThreads>(s, sm.dscr) / (double)t;
  st.mn = (double)block_min<kRnThreads>(mn, sm.fscr);
  st.mx = (double)block_max<kRnThreads>(mx, sm.fscr);
  st.sd = 0.0; st.hi_ref = st.mx; st.lo_ref = st.mn;
  if (mode == MPCG_RN_ZSCORE) {
    // ---- sweep B (z-score): population variance around the fp64 mean
    double q4[4] = {0.0, 0.0, 0.0, 0.0};
    rn_sweep(xr, t, tid, kRnThreads, [&](float v, int c) { const double d = (double)v - st.mean; q4[c] = fma(d, d, q4[c]); });
    st.sd = sqrt(block_sum<kRnThreads>((q4[0] + q4[1]) + (q4[2] + q4[3]), sm.dscr) / (double)t);
  } else if constexpr (KPEAK) {
    const long long kk = k < t ? k : t;
    __syncthreads();
    if (tid < 32) { rn_select(sm.hist_hi, kRnBins, kk, true, sm.sel); rn_select(sm.hist_hi, kRnBins, kk, false, sm.sel + 2); }
    __syncthreads();
    const unsigned b1_hi = sm.sel[0], b1_lo = sm.sel[2];
    const long long need1_hi = sm.sel[1], need1_lo = sm.sel[3];
    __syncthreads();
    for (int i = tid; i < kRnBins; i += kRnThreads) { sm.hist_hi[i] = 0; sm.hist_lo[i] = 0; }
    __syncthreads();
    // ---- sweep B (k-peak)
    double s_hi = 0.0, s_lo = 0.0;
    rn_sweep(xr, t, tid, kRnThreads, [&](float v, int) {
      const unsigned key = rn_key(v), p = key >> 21;
      if (p > b1_hi) s_hi += (double)v; else if (p == b1_hi) atomicAdd(&sm.hist_hi[(key >> 10) & 2047u], 1);
      if (p < b1_lo) s_lo += (double)v; else if (p == b1_lo) atomicAdd(&sm.hist_lo[(key >> 10) & 2047u], 1);
    });
    __syncthreads();
    if (tid < 32) { rn_select(sm.hist_hi, kRnBins, need1_hi, true, sm.sel); rn_select(sm.hist_lo, kRnBins, need1_lo, false, sm.sel + 2); }
    __syncthreads();
    const unsigned p_hi = (b1_hi << 11) | (unsigned)sm.sel[0], p_lo = (b1_lo << 11) | (unsigned)sm.sel[2];
    const long long need2_hi = sm.sel[1], need2_lo = sm.sel[3];
    __syncthreads();
    for (int i = tid; i < 1024; i += kRnThreads) { sm.hist_hi[i] = 0; sm.hist_lo[i] = 0; }
    __syncthreads();
    // ---- sweep C: inside the boundary bins of sweep A; beyond the 22-bit prefix -> sum, on it -> last 10 bits
    rn_sweep(xr, t, tid, kRnThreads, [&](float v, int) {
      const unsigned key = rn_key(v), p = key >> 10;
      if ((p >> 11) == b1_hi) { if (p > p_hi) s_hi += (double)v; else if (p == p_hi) atomicAdd(&sm.hist_hi[key & 1023u], 1); }
      if ((p >> 11) == b1_lo) { if (p < p_lo) s_lo += (double)v; else if (p == p_lo) atomicAdd(&sm.hist_lo[key & 1023u], 1); }
    });
    __syncthreads();
    if (tid < 32) { rn_select(sm.hist_hi, 1024, need2_hi, true, sm.sel); rn_select(sm.hist_lo, 1024, need2_lo, false, sm.sel + 2); }
    __syncthreads();
    // the members of the last bins follow from (key, count); the k-th key itself contributes `remaining` copies
    {
      const int c_hi = sm.sel[0], c_lo = sm.sel[2];
      for (int j = tid; j < 1024; j += kRnThreads) {
        if (j > c_hi) s_hi += (double)sm.hist_hi[j] * (double)rn_val((p_hi << 10) | (unsigned)j);
        if (j < c_lo) s_lo += (double)sm.hist_lo[j] * (double)rn_val((p_lo << 10) | (unsigned)j);
      }
      if (tid == 0) {
        s_hi += (double)sm.sel[1] * (double)rn_val((p_hi << 10) | (unsigned)c_hi);
        s_lo += (double)sm.sel[3] * (double)rn_val((p_lo << 10) | (unsigned)c_lo);
      }
    }
    st.hi_ref = block_sum<kRnThreads>(s_hi, sm.dscr) / (double)kk;
    st.lo_ref = block_sum<kRnThreads>(s_lo, sm.dscr) / (double)kk;
  }
  if (stats != nullptr && tid == 0) {
    double* o = stats + row * 8;
    o[0] = st.mn; o[1] = st.mx; o[2] = st.mean; o[3] = st.sd; o[4] = st.hi_ref; o[5] = st.lo_ref; o[6] = 0.0; o[7] = 0.0;
  }
  if (!apply) return;
  double a, b, sc;
  bool isconst;
  rn_coeffs(mode, flags, lo, hi, st, a, b, sc, isconst);
  rn_apply(xr, y + row * t, t, a, b, sc, isconst, tid, kRnThreads);
}

// Whole-tensor statistics from the per-row ones (the reference's tensor functions reduce over every element:
// normalize.py:41-44 `x.max() - x.min()`, :76-78 `topk(...).values.mean()`).  One CTA; writes the map into coef[4].
__global__ void __launch_bounds__(256)
row_normalise_combine_kernel(const double* __restrict__ stats, long long rows, int mode, double lo, double hi, int flags,
                             double* __restrict__ coef) {
  __shared__ double dscr[32];
  __shared__ double dmin[8], dmax[8];
  const int tid = threadIdx.x;
  double mn = INFINITY, mx = -INFINITY, sh = 0.0, sl = 0.0;
  for (long long r = tid; r < rows; r += 256) {
    const double* o = stats + r * 8;
    mn = fmin(mn, o[0]); mx = fmax(mx, o[1]); sh += o[4]; sl += o[5];
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, d));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  }
  if ((tid & 31) == 0) { dmin[tid >> 5] = mn; dmax[tid >> 5] = mx; }
  sh = block_sum<256>(sh, dscr);
  sl = block_sum<256>(sl, dscr);
  __syncthreads();
  if (tid == 0) {
    RnStats st;
    st.mn = dmin[0]; st.mx = dmax[0];
    for (int w = 1; w < 8; ++w) { st.mn = fmin(st.mn, dmin[w]); st.mx = fmax(st.mx, dmax[w]); }
    st.mean = 0.0; st.sd = 0.0;
    st.hi_ref = sh / (double)rows; st.lo_ref = sl / (double)rows;      // every row contributes k values
    double a, b, s;
    bool isconst;
    rn_coeffs(mode, flags, lo, hi, st, a, b, s, isconst);
    coef[0] = a; coef[1] = b; coef[2] = s; coef[3] = isconst ? 1.0 : 0.0;
  }
}

__global__ void __launch_bounds__(kRnThreads)
row_normalise_apply_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, const double* __restrict__ coef) {
  const double a = coef[0], b = coef[1], s = coef[2];
  const bool isconst = coef[3] != 0.0;
  rn_apply(x, y, n, a, b, s, isconst, (int)(blockIdx.x * kRnThreads + threadIdx.x), (int)(gridDim.x * kRnThreads));
}

}  // namespace mpcg

extern "C" int mpcg_row_normalise_f32(const float* x, float* y, double* stats, int64_t rows, int64_t t, int mode, int k,
                                      double lo, double hi, int flags, void* stream) {
  using namespace mpcg;
  if (rows < 0 || t < 0) return MPCG_EINVAL;
  if (mode != MPCG_RN_MINMAX && mode != MPCG_RN_ZSCORE && mode != MPCG_RN_KPEAK) return MPCG_EINVAL;
  if (mode == MPCG_RN_KPEAK && k < 1) return MPCG_EINVAL;
  const bool global = (flags & MPCG_RN_GLOBAL) != 0;
  if (global && mode == MPCG_RN_ZSCORE) return MPCG_EINVAL;
  if (rows == 0 || t == 0) return MPCG_OK;
  if (!x || !y || (global && !stats)) return MPCG_EINVAL;
  if (rows > 0x7fffffffLL) return MPCG_ERANGE;
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == MPCG_RN_KPEAK)
    row_normalise_kernel<true><<<(unsigned)rows, kRnThreads, 0, st>>>(x, y, stats, (long long)t, mode, k, lo, hi, flags, global ? 0 : 1);
  else
    row_normalise_kernel<false><<<(unsigned)rows, kRnThreads, 0, st>>>(x, y, stats, (long long)t, mode, k, lo, hi, flags, global ? 0 : 1);
  MPCG_LAUNCH_CHECK();
  if (global) {
    double* coef = stats + rows * 8;                                 // stats holds [rows + 1, 8] doubles in this scope
    row_normalise_combine_kernel<<<1, 256, 0, st>>>(stats, (long long)rows, mode, lo, hi, flags, coef);
    MPCG_LAUNCH_CHECK();
    const long long n = (long long)rows * t;
    const unsigned grid = (unsigned)min((long long)148 * 8, (n + kRnThreads - 1) / kRnThreads);
    row_normalise_apply_kernel<<<grid, kRnThreads, 0, st>>>(x, y, n, coef);
    MPCG_LAUNCH_CHECK();
  }
  return MPCG_OK;
}
