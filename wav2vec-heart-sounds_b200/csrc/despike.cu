// extern "C" entry: mpcg_despike_f32 -- one CTA per recording; frame maxima cached in shared memory so a
// pass rescans only the frame it edited (the reference recomputes all maxima and the median every pass).
#include "despike.cuh"

namespace mpcg {

constexpr int kDsThreads = 256;

__global__ void __launch_bounds__(kDsThreads)
despike_rows_kernel(float* __restrict__ x, long long t, int win, int nframes, double threshold, int max_iter,
                    int median_mode, int* __restrict__ edits, int* __restrict__ trace, int trace_cap) {
  extern __shared__ __align__(16) float ds_smem[];
  __shared__ float fscr[40];
  __shared__ int iscr[32];
  float* tops = ds_smem;                                   // [nframes]
  float* frame_buf = ds_smem + ((nframes + 3) & ~3) + 4;   // [win + 4], phase-matched below
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long row = blockIdx.x;
  float* xr = x + row * t;

  // frame maxima: one warp per frame, coalesced
  for (int f = warp; f < nframes; f += kDsThreads / 32) {
    const float* p = xr + (long long)f * win;
    float m = 0.f;
    for (int i = lane; i < win; i += 32) m = fmaxf(m, fabsf(p[i]));
    m = warp_max(m);
    if (lane == 0) tops[f] = m;
  }
  __syncthreads();

  int passes = 0;
  for (; passes < max_iter; ++passes) {
    const SpikeDecision dec = spike_decide<kDsThreads>(tops, nframes, threshold, median_mode, fscr, iscr);
    if (!dec.active) break;
    float* g = xr + (long long)dec.worst * win;
    float* fr = frame_buf + phase_of(g);                   // 16-byte phase match with the global frame
    copy_g2s<kDsThreads, false>(fr, g, win);
    __syncthreads();
    int peak, lo, hi;
    bool changed;
    float new_top;
    spike_flatten<kDsThreads>(fr, win, peak, lo, hi, changed, new_top, fscr, iscr);
    for (int i = lo + tid; i < hi; i += kDsThreads) g[i] = kSpikeFill;
    if (tid == 0) {
      tops[dec.worst] = new_top;
      if (trace && passes < trace_cap) {
        int* tr = trace + ((long long)row * trace_cap + passes) * 4;
        tr[0] = dec.worst; tr[1] = peak; tr[2] = lo; tr[3] = hi;
      }
    }
    __syncthreads();
    if (!changed) { ++passes; break; }                     // fixed point: later passes would repeat this one
  }
  if (edits && tid == 0) edits[row] = passes;
}

}  // namespace mpcg

extern "C" int mpcg_despike_f32(float* x, int64_t rows, int64_t t, int64_t win, double threshold, int max_iterations,
                                int median_mode, int32_t* edits, int32_t* trace, int trace_cap, void* stream) {
  using namespace mpcg;
  if (rows < 0 || t < 0 || max_iterations < 0 || trace_cap < 0) return MPCG_EINVAL;
  if (median_mode != MPCG_MEDIAN_LOWER && median_mode != MPCG_MEDIAN_MEAN) return MPCG_EINVAL;
  if (rows == 0) return MPCG_OK;
  if (win < 1 || t < win) {                                 // the reference returns the clone untouched
    if (edits) {
      cudaError_t e = cudaMemsetAsync(edits, 0, sizeof(int32_t) * rows, (cudaStream_t)stream);
      if (e != cudaSuccess) return (int)e;
    }
    return MPCG_OK;
  }
  if (!x) return MPCG_EINVAL;
  const int64_t nframes = t / win;
  if (nframes > kDespikeMaxFrames || win > 50000 || rows > 0x7fffffffLL) return MPCG_ERANGE;
  const size_t smem = (size_t)(((nframes + 3) & ~3) + 4 + win + 8) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(despike_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  despike_rows_kernel<<<(unsigned)rows, kDsThreads, smem, (cudaStream_t)stream>>>(
      x, (long long)t, (int)win, (int)nframes, threshold, max_iterations, median_mode, edits, trace, trace_cap);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}
