// extern "C" entry: mpcg_gen_condition_f32 -- the per-item conditioning of the generator datasets, batched
// (reference datasets/generative.py:77-115: abs_max_normalise -> _fade -> fit_length [-> add_chirp], and
// signalproc/preprocess.py:45-64 for fit_length / add_chirp; the mel of the conditioning waveform is mpcg_mel_*).
//   y[i]     = i < t ? N(x)[i] * fade(i) : 0            for i < crop      (fade: 128-sample linear ramps at both ends of
//                                                                          the length-t signal, skipped when t < 2*128)
//   chirp[i] = y[i] + cos(2 pi * 0.5 * beta * t_i^2) * max(0.5, max|y|),   beta = (fs/2) / t_last,  t_i = i / fs
// One CTA per row, three sweeps (statistics, output + its maximum, chirp), the row re-read from L2.
#include "aug.cuh"

namespace mpcg {

constexpr int kGnThreads = 512;

__global__ void __launch_bounds__(kGnThreads)
gen_condition_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ chirp,
                     const long long* __restrict__ row_len, long long stride, long long crop, int fade_n, double fs, int norm_flags) {
  __shared__ double dscr[32];
  __shared__ float fscr[32];
  const long long row = blockIdx.x;
  const float* xr = x + row * stride;
  const long long t = row_len ? min(max(row_len[row], 0LL), stride) : stride;   // valid samples of this row
  float* yr = y + row * crop;
  const int tid = threadIdx.x;
  // sweep 1: row statistics of the whole recording (normalisation precedes the crop)
  RowStats st;
  stats_init(st);
  for (long long i = tid; i < t; i += kGnThreads) stats_add(st, xr[i]);
  const double tot = block_sum<kGnThreads>(st.sum, dscr);
  const float lo = block_min<kGnThreads>(st.lo, fscr), hi = block_max<kGnThreads>(st.hi, fscr);
  double mean = t > 0 ? tot / (double)t : 0.0;
  const double peak = fmax((double)hi - mean, mean - (double)lo);
  double inv_peak;
  if (norm_flags & MPCG_NORM_PEAK_GT0) inv_peak = (peak > 0.0) ? 1.0 / peak : 1.0;
  else inv_peak = 1.0 / fmax(peak, 1e-12);
  bool clampit = (norm_flags & MPCG_NORM_PEAK_GT0) == 0;            // the tensor path clamps, the NumPy path does not
  if (norm_flags & MPCG_GEN_NO_NORM) { mean = 0.0; inv_peak = 1.0; clampit = false; }   // rows rebuilt from normalised cycles
  const bool fade = fade_n > 1 && t >= 2LL * fade_n;
  const double step = fade_n > 1 ? 1.0 / (double)(fade_n - 1) : 0.0;
  // sweep 2: normalise, fade, fit to `crop`; maximum magnitude of what is written
  float vmax = 0.f;
  for (long long i = tid; i < crop; i += kGnThreads) {
    float v = 0.f;
    if (i < t) {
      double u = ((double)xr[i] - mean) * inv_peak;
      if (clampit) u = fmin(fmax(u, -1.0), 1.0);
      if (fade) {
        if (i < fade_n) u *= (double)i * step;                      // np.linspace(0, 1, n)[i]
        else if (i >= t - fade_n) u *= 1.0 - (double)(i - (t - fade_n)) * step;
      }
      v = (float)u;
    }
    yr[i] = v;
    vmax = fmaxf(vmax, fabsf(v));
  }
  if (chirp == nullptr) return;
  vmax = block_max<kGnThreads>(vmax, fscr);
  __syncthreads();
  // sweep 3: reference + full-band linear chirp (scipy.signal.chirp: cos(2 pi (f0 t + beta t^2 / 2)), f0 = 0)
  const float gain = fmaxf(0.5f, vmax);
  const double t_last = crop > 1 ? (double)(crop - 1) / fs : 1.0;
  const double beta = (fs * 0.5) / t_last;
  float* cr = chirp + row * crop;
  for (long long i = tid; i < crop; i += kGnThreads) {
    const double ti = (double)i / fs;
    double turns = 0.5 * beta * ti * ti;                            // phase / 2 pi
    turns -= floor(turns);
    cr[i] = yr[i] + (float)cospi(2.0 * turns) * gain;
  }
}

}  // namespace mpcg

extern "C" int mpcg_gen_condition_rows_f32(const float* x, float* y, float* chirp, const int64_t* row_len, int64_t rows,
                                           int64_t t, int64_t crop, int fade_n, double fs, int norm_flags, void* stream) {
  using namespace mpcg;
  if (rows < 0 || t < 0 || crop < 0 || fade_n < 0 || !(fs > 0.0)) return MPCG_EINVAL;
  if (rows == 0 || crop == 0) return MPCG_OK;
  if (!y || (t > 0 && !x)) return MPCG_EINVAL;
  if (rows > 0x7fffffffLL) return MPCG_ERANGE;
  gen_condition_kernel<<<(unsigned)rows, kGnThreads, 0, (cudaStream_t)stream>>>(
      x, y, chirp, reinterpret_cast<const long long*>(row_len), (long long)t, (long long)crop, fade_n, fs, norm_flags);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}

extern "C" int mpcg_gen_condition_f32(const float* x, float* y, float* chirp, int64_t rows, int64_t t, int64_t crop,
                                      int fade_n, double fs, int norm_flags, void* stream) {
  return mpcg_gen_condition_rows_f32(x, y, chirp, nullptr, rows, t, crop, fade_n, fs, norm_flags, stream);
}
