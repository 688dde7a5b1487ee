// extern "C" entries: mpcg_mel_f32, mpcg_logmap_f32 -- mel conditioning features
// (signalproc/spectrogram.py:13-45: torchaudio MelSpectrogram(power=1, normalized=True, center=True, reflect pad,
// periodic Hann, HTK filterbank) followed by  clamp((20*log10(clamp(mel, 1e-5)) - 20 + 100) / 100, 0, 1)).
//
// Framing + windowed DFT is a contraction  [frames x n_fft] . [n_fft x 2*bins]  against a basis that already
// carries the window and the 1/sqrt(sum w^2) normalisation, restricted to the DFT bins that have mel weight
// (31 of 513 at 16 kHz / n_fft 1024).  This first version runs it on the fp32 FMA pipe: one CTA stages the
// reflect-padded samples of FT frames in shared memory, warps own frame groups, lanes own bins, so sample
// loads are warp broadcasts and basis loads are coalesced and L1-resident across the CTA's warps.  Magnitude,
// mel projection and the log map are fused behind it; only the mel matrix goes back to HBM.
#include "common.cuh"

namespace mpcg {

constexpr int kMelThreads = 256;
constexpr int kMelWarps = kMelThreads / 32;

struct MelArgs {
  const float* x;         // [rows, t]
  float* out;             // [rows, n_mels, frames]
  const void* basis;      // [n_hi - n_lo][2][kpad]   (cos | -sin), windowed and normalised, zero-padded bins;
                          // double for the exact path, float for the fast one
  const float* fb;        // [nbins][n_mels]
  long long t;
  int n_fft, hop, n_lo, n_hi, nbins, kpad, n_mels, frames, log_map;
};

__device__ __forceinline__ float log_map(float mel) {
  const float db = 20.f * log10f(fmaxf(mel, 1e-5f)) - 20.f;
  return fminf(fmaxf((db + 100.f) / 100.f, 0.f), 1.f);
}

// ACC = double: basis and accumulation in fp64.  Leakage skirts of tonal inputs sit 5-6 decades below the peak
// and the dB map clamps at 1e-5, so float accumulation (error ~1e-7 of the frame's peak bin) is visible there;
// fp64 keeps the 1e-5 bound against the float64 oracle on every input.  ACC = float is the fast path.
// FPW = frames per warp (4, or 1 when 31 hops + a window of samples would not fit shared memory): the CTA covers
// kMelFT = 8 FPW frames.
template <typename ACC, int FPW>
__global__ void __launch_bounds__(kMelThreads)
mel_kernel(const MelArgs a) {
  constexpr int kMelFPW = FPW, kMelFT = kMelWarps * FPW;
  extern __shared__ __align__(16) float ml_smem[];
  const int span = (kMelFT - 1) * a.hop + (a.n_hi - a.n_lo);     // samples this CTA's frames touch
  float* xs = ml_smem;                                           // [span]
  float* mag = ml_smem + ((span + 3) & ~3);                      // [kMelFT][kpad + 1] (odd stride: no bank conflicts)
  const int ms = a.kpad + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long row = blockIdx.y;
  const int f0 = blockIdx.x * kMelFT;
  const float* xr = a.x + row * a.t;
  const int pad = a.n_fft / 2;
  // stage: sample index in the reflect-padded signal = f*hop + n ; un-padded = that - pad
  const long long base = (long long)f0 * a.hop + a.n_lo - pad;
  for (int m = tid; m < span; m += kMelThreads) {
    long long j = base + m;
    if (j < 0) j = -j;
    if (j >= a.t) j = 2 * (a.t - 1) - j;
    xs[m] = (j >= 0 && j < a.t) ? xr[j] : 0.f;
  }
  __syncthreads();
  const int nn = a.n_hi - a.n_lo;
  const float* xw = xs + warp * kMelFPW * a.hop;                 // first frame of this warp
  for (int kb = 0; kb < a.kpad; kb += 32) {
    ACC re[kMelFPW], im[kMelFPW];
#pragma unroll
    for (int f = 0; f < kMelFPW; ++f) { re[f] = (ACC)0; im[f] = (ACC)0; }
    const ACC* bc = reinterpret_cast<const ACC*>(a.basis) + kb + lane;
#pragma unroll 4
    for (int n = 0; n < nn; ++n) {
      const ACC c = __ldg(bc + (long long)n * 2 * a.kpad);
      const ACC s = __ldg(bc + (long long)n * 2 * a.kpad + a.kpad);
#pragma unroll
      for (int f = 0; f < kMelFPW; ++f) {
        const ACC v = (ACC)xw[f * a.hop + n];                    // warp-wide broadcast
        re[f] = fma(v, c, re[f]);
        im[f] = fma(v, s, im[f]);
      }
    }
#pragma unroll
    for (int f = 0; f < kMelFPW; ++f)
      mag[(warp * kMelFPW + f) * ms + kb + lane] = (float)sqrt(re[f] * re[f] + im[f] * im[f]);
  }
  __syncthreads();
  // mel projection: out[m, frame] = sum_k fb[k, m] * mag[frame, k]; lanes run over frames so stores are contiguous
  for (int o = tid; o < a.n_mels * kMelFT; o += kMelThreads) {
    const int m = o / kMelFT, f = o - m * kMelFT;
    if (f0 + f >= a.frames) continue;
    const float* mg = mag + f * ms;
    float acc = 0.f;
    for (int k = 0; k < a.nbins; ++k) acc = fmaf(__ldg(a.fb + (long long)k * a.n_mels + m), mg[k], acc);
    a.out[(row * a.n_mels + m) * a.frames + f0 + f] = a.log_map ? log_map(acc) : acc;
  }
}

__global__ void logmap_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = log_map(x[i]);
}

}  // namespace mpcg

extern "C" int mpcg_mel_f32(const float* x, float* out, int64_t rows, int64_t t, int n_fft, int hop, int n_lo, int n_hi,
                            int nbins, int kpad, const void* basis, int basis_f64, const float* fb, int n_mels,
                            int64_t frames, int log_map, void* stream) {
  using namespace mpcg;
  if (rows < 0 || t < 0 || n_fft < 2 || hop < 1 || n_lo < 0 || n_hi <= n_lo || n_hi > n_fft || nbins < 1 ||
      kpad < nbins || (kpad & 31) != 0 || n_mels < 1 || frames < 0)
    return MPCG_EINVAL;
  if (frames != 1 + t / hop) return MPCG_EINVAL;
  if (rows == 0 || frames == 0) return MPCG_OK;
  if (!x || !out || !basis || !fb) return MPCG_EINVAL;
  if (t <= n_fft / 2) return MPCG_EINVAL;                    // reflect padding needs pad < length (torch raises too)
  if (rows > 65535 || frames > 0x3fffffffLL) return MPCG_ERANGE;
  MelArgs a;
  a.x = x; a.out = out; a.basis = basis; a.fb = fb; a.t = t; a.n_fft = n_fft; a.hop = hop; a.n_lo = n_lo; a.n_hi = n_hi;
  a.nbins = nbins; a.kpad = kpad; a.n_mels = n_mels; a.frames = (int)frames; a.log_map = log_map;
  auto smem_of = [&](int ft) {
    const int span = (ft - 1) * hop + (n_hi - n_lo);
    return (size_t)(((span + 3) & ~3) + ft * (kpad + 1)) * sizeof(float);
  };
  const int fpw = smem_of(4 * kMelWarps) <= 220 * 1024 ? 4 : 1;     // long hops: fewer frames share a CTA's sample span
  const int ft = fpw * kMelWarps;
  const size_t smem = smem_of(ft);
  if (smem > 220 * 1024) return MPCG_ERANGE;
  dim3 grid((unsigned)((frames + ft - 1) / ft), (unsigned)rows);
#define MPCG_MEL_LAUNCH(ACC, FPW)                                                                                     \
  {                                                                                                                   \
    cudaError_t e = cudaFuncSetAttribute(mel_kernel<ACC, FPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return (int)e;                                                                              \
    mel_kernel<ACC, FPW><<<grid, kMelThreads, smem, (cudaStream_t)stream>>>(a);                                        \
  }
  if (basis_f64) {
    if (fpw == 4) MPCG_MEL_LAUNCH(double, 4) else MPCG_MEL_LAUNCH(double, 1)
  } else {
    if (fpw == 4) MPCG_MEL_LAUNCH(float, 4) else MPCG_MEL_LAUNCH(float, 1)
  }
#undef MPCG_MEL_LAUNCH
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}

extern "C" int mpcg_logmap_f32(const float* x, float* y, int64_t n, void* stream) {
  using namespace mpcg;
  if (n < 0) return MPCG_EINVAL;
  if (n == 0) return MPCG_OK;
  if (!x || !y) return MPCG_EINVAL;
  logmap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, (long long)n);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}
