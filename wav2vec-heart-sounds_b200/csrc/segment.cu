// extern "C" entries: mpcg_segment_f32, mpcg_window_count -- overlapping-window gather
// (signalproc/torchproc.py:119-129: drop `start` samples, zero-pad to one window if short, unfold;
//  signalproc/segment.py:30-52 for the [N, win, C] loader layout).
//
// Pure data movement: every output element is written once with 128-bit stores; sources are read
// through L1/L2 (neighbouring windows overlap by 0.25 s, so the re-read hits cache).  The integer
// window arithmetic is the reference's and is checked bit-exactly by the tests.
#include "common.cuh"

namespace mpcg {

constexpr int kSgThreads = 256;

// planar: out[row, ch, k, j] = x[row, ch, start + k*hop + j]
__global__ void __launch_bounds__(kSgThreads)
segment_planar_kernel(const float* __restrict__ x, float* __restrict__ out, long long t, long long start,
                      long long win, long long hop, long long n) {
  const long long line = blockIdx.y;                       // (row*channels + ch)
  const long long k = blockIdx.z;
  const float* src = x + line * t;
  float* dst = out + (line * n + k) * win;
  const long long s0 = start + k * hop;
  const long long j0 = ((long long)blockIdx.x * kSgThreads + threadIdx.x) * 4;
  if (j0 >= win) return;
  float v[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const long long s = s0 + j0 + q;
    v[q] = (j0 + q < win && s < t) ? __ldg(src + s) : 0.f;
  }
  if (j0 + 3 < win && (((uintptr_t)(dst + j0)) & 15u) == 0) {
    st_stream4(reinterpret_cast<float4*>(dst + j0), make_float4(v[0], v[1], v[2], v[3]));
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (j0 + q < win) dst[j0 + q] = v[q];
  }
}

// channels last: out[row, k, j, ch] = x[row, ch, start + k*hop + j]; a 32-sample x C tile is transposed
// through shared memory so both sides stay coalesced.
constexpr int kSgTileJ = 256;
__global__ void __launch_bounds__(kSgThreads)
segment_interleave_kernel(const float* __restrict__ x, float* __restrict__ out, int channels, long long t,
                          long long start, long long win, long long hop, long long n) {
  extern __shared__ float sg_smem[];                       // [channels][kSgTileJ + 1]
  const long long row = blockIdx.y;
  const long long k = blockIdx.z;
  const long long j0 = (long long)blockIdx.x * kSgTileJ;
  const long long s0 = start + k * hop + j0;
  const int nj = (int)min((long long)kSgTileJ, win - j0);
  for (int idx = threadIdx.x; idx < channels * kSgTileJ; idx += kSgThreads) {
    const int c = idx / kSgTileJ, j = idx - c * kSgTileJ;
    const long long s = s0 + j;
    sg_smem[c * (kSgTileJ + 1) + j] = (j < nj && s < t) ? __ldg(x + (row * channels + c) * t + s) : 0.f;
  }
  __syncthreads();
  float* dst = out + ((row * n + k) * win + j0) * channels;
  for (int idx = threadIdx.x; idx < nj * channels; idx += kSgThreads) {
    const int j = idx / channels, c = idx - j * channels;
    dst[idx] = sg_smem[c * (kSgTileJ + 1) + j];
  }
}

}  // namespace mpcg

extern "C" int64_t mpcg_window_count(int64_t t, int64_t start, int64_t win, int64_t hop) {
  if (win < 1 || hop < 1 || t < 0 || start < 0) return MPCG_EINVAL;
  int64_t rem = t - start;
  if (rem < win) rem = win;                                // short (or empty) remainder is zero-padded to one window
  return (rem - win) / hop + 1;
}

extern "C" int mpcg_segment_f32(const float* x, float* out, int64_t rows, int64_t channels, int64_t t, int64_t start,
                                int64_t win, int64_t hop, int64_t n, int channels_last, void* stream) {
  using namespace mpcg;
  if (rows < 0 || channels < 1 || t < 0 || start < 0 || win < 1 || hop < 1 || n < 0) return MPCG_EINVAL;
  if (n != mpcg_window_count(t, start, win, hop)) return MPCG_EINVAL;
  if (rows == 0) return MPCG_OK;
  if (!out || (!x && t > 0)) return MPCG_EINVAL;
  if (n > 65535) return MPCG_ERANGE;
  if (!channels_last || channels == 1) {
    const int64_t lines = rows * channels;
    for (int64_t l0 = 0; l0 < lines; l0 += 65535) {
      const int64_t nl = lines - l0 < 65535 ? lines - l0 : 65535;
      dim3 grid((unsigned)((win + kSgThreads * 4 - 1) / (kSgThreads * 4)), (unsigned)nl, (unsigned)n);
      segment_planar_kernel<<<grid, kSgThreads, 0, (cudaStream_t)stream>>>(x + l0 * t, out + l0 * n * win, (long long)t,
                                                                          (long long)start, (long long)win,
                                                                          (long long)hop, (long long)n);
      MPCG_LAUNCH_CHECK();
    }
  } else {
    if (channels > 64) return MPCG_ERANGE;
    const size_t smem = (size_t)channels * (kSgTileJ + 1) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(segment_interleave_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
    for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
      const int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
      dim3 grid((unsigned)((win + kSgTileJ - 1) / kSgTileJ), (unsigned)nr, (unsigned)n);
      segment_interleave_kernel<<<grid, kSgThreads, smem, (cudaStream_t)stream>>>(
          x + r0 * channels * t, out + r0 * n * win * channels, (int)channels, (long long)t, (long long)start,
          (long long)win, (long long)hop, (long long)n);
      MPCG_LAUNCH_CHECK();
    }
  }
  return MPCG_OK;
}
