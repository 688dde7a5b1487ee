// extern "C" entry: mpcg_sosfiltfilt_f32 -- zero-phase IIR filtering of rows.
// Reference: signalproc/filters.py:44-90 (butter_bandpass / butter_lowpass / butter_highpass / band_stop call
// scipy.signal.sosfiltfilt; notch / notch_chain call scipy.signal.filtfilt on one biquad, the same arithmetic).
// SciPy's recipe (site-packages/scipy/signal/_signaltools.py, sosfiltfilt): odd extension of the row by `edge` samples
// on both sides, forward pass started from the steady-state conditions zi * ext[0], backward pass over the reversed
// result started from zi * (its first sample), reverse again, drop the extension.
//
// One CTA per row, two streamed passes over tiles of 256 x 61 samples with the chunked linear-recurrence scan of
// biquad.cuh (fp64 state); the forward result (length t + 2 edge, float32) goes to a caller-supplied workspace and is
// read back in reverse order by the backward pass, which writes only the t interior samples.
#include "biquad.cuh"

namespace mpcg {

constexpr int kFfThreads = 256;
constexpr int kFfTile = kFfThreads * kBqL;
struct FfSmem {
  BqScratch<kFfThreads> sc;
  float tile[kFfTile + 8];
};
struct FfZi {
  double z[kBqMaxGroups][4];               // steady-state DF-II-T states of each two-section group for a unit step
};

// value of the odd extension at extended index e (0 <= e < t + 2 edge)
__device__ __forceinline__ float ff_ext(const float* __restrict__ x, long long t, long long edge, long long e) {
  if (e < edge) return 2.f * x[0] - x[edge - e];
  const long long i = e - edge;
  if (i < t) return x[i];
  return 2.f * x[t - 1] - x[t - 2 - (i - t)];
}

__global__ void __launch_bounds__(kFfThreads)
filtfilt_rows_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ work, long long t, long long edge,
                     const __grid_constant__ BqPlan plan, const __grid_constant__ FfZi zi, int epilogue) {
  extern __shared__ __align__(16) unsigned char ff_raw[];
  FfSmem& sm = *reinterpret_cast<FfSmem*>(ff_raw);
  const long long row = blockIdx.x;
  const float* xr = x + row * t;
  float* yr = y + row * t;
  const long long te = t + 2 * edge;
  float* wr = work + row * te;
  const int tid = threadIdx.x;
  for (int pass = 0; pass < 2; ++pass) {
    bq_init_scratch<kFfThreads>(sm.sc, plan);
    __syncthreads();
    // initial conditions: zi * (first sample this pass sees)
    const float first = pass == 0 ? ff_ext(xr, t, edge, 0) : wr[te - 1];
    if (tid < plan.ngroups * 4) sm.sc.carry[tid >> 2][tid & 3] = zi.z[tid >> 2][tid & 3] * (double)first;
    __syncthreads();
    for (long long t0 = 0; t0 < te; t0 += kFfTile) {
      const int n = (int)((te - t0) < (long long)kFfTile ? (te - t0) : (long long)kFfTile);
      float* sh = sm.tile;
      if (pass == 0) {
        for (int i = tid; i < n; i += kFfThreads) sh[i] = ff_ext(xr, t, edge, t0 + i);
      } else {
        for (int i = tid; i < n; i += kFfThreads) sh[i] = wr[te - 1 - (t0 + i)];       // reversed
      }
      for (int i = n + tid; i < kFfTile; i += kFfThreads) sh[i] = 0.f;
      __syncthreads();
      bq_filter_tile<kFfThreads>(sh, sm.sc, plan);
      __syncthreads();
      if (pass == 0) {
        for (int i = tid; i < n; i += kFfThreads) wr[t0 + i] = sh[i];
      } else {
        for (int i = tid; i < n; i += kFfThreads) {
          const long long e = te - 1 - (t0 + i);                                         // extended index of this output
          if (e >= edge && e < edge + t) yr[e - edge] = epilogue == MPCG_EPI_EXP ? (float)exp((double)sh[i]) : sh[i];
        }
      }
      __syncthreads();
    }
    __threadfence_block();
    __syncthreads();
  }
}

}  // namespace mpcg

extern "C" int mpcg_sosfiltfilt_epi_f32(const float* x, float* y, float* work, int64_t rows, int64_t t, const double* sos,
                                        int n_sections, const double* zi, int64_t edge, int epilogue, void* stream) {
  using namespace mpcg;
  if (rows < 0 || t < 0 || edge < 0) return MPCG_EINVAL;
  if (epilogue != MPCG_EPI_NONE && epilogue != MPCG_EPI_EXP) return MPCG_EINVAL;
  BqPlan plan;
  const int rc = bq_make_plan(sos, n_sections, &plan);
  if (rc != MPCG_OK) return rc;
  if (!zi) return MPCG_EINVAL;
  if (rows == 0 || t == 0) return MPCG_OK;
  if (t <= edge) return MPCG_EINVAL;                        // SciPy: "length of the input vector x must be greater than padlen"
  if (!x || !y || !work) return MPCG_EINVAL;
  if (rows > 0x7fffffffLL) return MPCG_ERANGE;
  FfZi z;
  for (int g = 0; g < kBqMaxGroups; ++g)
    for (int s = 0; s < 4; ++s) {
      const int sec = 2 * g + (s >> 1);
      z.z[g][s] = sec < n_sections ? zi[2 * sec + (s & 1)] : 0.0;
    }
  cudaError_t e = cudaFuncSetAttribute(filtfilt_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FfSmem));
  if (e != cudaSuccess) return (int)e;
  filtfilt_rows_kernel<<<(unsigned)rows, kFfThreads, sizeof(FfSmem), (cudaStream_t)stream>>>(x, y, work, (long long)t,
                                                                                            (long long)edge, plan, z, epilogue);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}

extern "C" int mpcg_sosfiltfilt_f32(const float* x, float* y, float* work, int64_t rows, int64_t t, const double* sos,
                                    int n_sections, const double* zi, int64_t edge, void* stream) {
  return mpcg_sosfiltfilt_epi_f32(x, y, work, rows, t, sos, n_sections, zi, edge, MPCG_EPI_NONE, stream);
}
