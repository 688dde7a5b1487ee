// Rational polyphase resampler in dense frame form (replaces torchaudio.functional.resample's
// strided conv1d behind torchproc.resample, signalproc/torchproc.py:56-59, and SciPy's
// resample_poly/upfirdn behind signalproc/resample.py:11-22).
//
//   y[i*UP + p] = sum_{d<D} x[i*DOWN + off + d] * G[p][d]          (x = 0 outside the row)
//
// One thread owns FR consecutive frames i: it pulls the (FR-1)*DOWN + D input samples it needs from a
// shared-memory tile into registers once, then runs UP*FR*D FFMAs whose tap operand is an instruction
// immediate: the taps of the named rate pairs are baked in at build time (resample_taps_gen.cuh) and the
// launcher takes a baked instance only when the host-supplied taps match it bit for bit; any other
// ratio goes to the generic kernel, whose taps ride in the parameter bank.  Outputs go through a second
// shared tile so the global stores are fully coalesced.  Lane strides in both tiles are made odd by a
// one-word-per-stride skew, which keeps every shared access conflict-free.
#include "resample.cuh"

namespace mpcg {

constexpr int kRsThreads = 256;

template <int UP, int DOWN, int D, int FR>
struct RsShape {
  using T = RsTile<UP, DOWN, D, FR, 1, kRsThreads>;
  static constexpr int SOUT = FR * UP;                       // lane stride in the output tile
  static constexpr int OUT_WORDS = rs_skew(T::NOUT - 1, SOUT) + 1;
  static constexpr size_t SMEM = (size_t)(T::IN_WORDS + OUT_WORDS) * sizeof(float);
};

template <int UP, int DOWN, int D, int FR>
__global__ void __launch_bounds__(kRsThreads)
resample_frames_kernel(const float* __restrict__ x, float* __restrict__ y, long long t_in, long long t_out,
                       long long off) {
  using S = RsShape<UP, DOWN, D, FR>;
  using T = typename S::T;
  extern __shared__ __align__(16) float rs_smem[];
  float* xs = rs_smem;
  float* ys = rs_smem + T::IN_WORDS;
  const int tid = threadIdx.x;
  const long long row = blockIdx.y;
  const long long f0 = (long long)blockIdx.x * T::NF;         // first frame of this CTA
  const float* xr = x + row * t_in;
  float* yr = y + row * t_out;

  T::stage(xs, xr, f0 * DOWN + off, t_in);
  __syncthreads();
  auto sink = [&](int frame, int p, float v) { ys[rs_skew(frame * UP + p, S::SOUT)] = v; };
  T::compute(xs, sink);
  __syncthreads();

  // ---- coalesced store of the CTA's output span
  const long long o0 = f0 * UP;
  for (int o = tid; o < T::NOUT; o += kRsThreads) {
    const long long dst = o0 + o;
    if (dst < t_out) st_stream(yr + dst, ys[rs_skew(o, S::SOUT)]);
  }
}

// Generic fallback for ratios without a specialised instance: taps live in shared memory.
constexpr int kRsGenericMaxTaps = 7168;            // up * D floats, bounded by the 32 KB parameter space
struct RsGenericTaps {
  float g[kRsGenericMaxTaps];
};
constexpr int kRsGenericOut = 4096;                // outputs per CTA (upper bound)

__global__ void __launch_bounds__(kRsThreads)
resample_generic_kernel(const float* __restrict__ x, float* __restrict__ y, long long t_in, long long t_out,
                        long long off, int up, int down, int D, int frames_per_cta,
                        const __grid_constant__ RsGenericTaps taps) {
  extern __shared__ __align__(16) float rs_smem[];
  float* ts = rs_smem;                              // up * D taps
  float* xs = rs_smem + up * D;                     // (frames_per_cta - 1) * down + D inputs
  const int tid = threadIdx.x;
  const long long row = blockIdx.y;
  const long long f0 = (long long)blockIdx.x * frames_per_cta;
  const float* xr = x + row * t_in;
  float* yr = y + row * t_out;
  for (int i = tid; i < up * D; i += kRsThreads) ts[i] = taps.g[i];
  const int nin = (frames_per_cta - 1) * down + D;
  const long long in0 = f0 * down + off;
  for (int m = tid; m < nin; m += kRsThreads) {
    const long long src = in0 + m;
    xs[m] = (src >= 0 && src < t_in) ? ld_stream(xr + src) : 0.f;
  }
  __syncthreads();
  const int nout = frames_per_cta * up;
  const long long o0 = f0 * up;
  for (int o = tid; o < nout; o += kRsThreads) {
    const long long dst = o0 + o;
    if (dst >= t_out) break;
    const int f = o / up, p = o - f * up;
    const float* xi = xs + f * down;
    const float* tp = ts + p * D;
    float acc = 0.f;
    for (int d = 0; d < D; ++d) acc = fmaf(xi[d], tp[d], acc);
    st_stream(yr + dst, acc);
  }
}

template <int UP, int DOWN, int D, int FR>
static int launch_frames(const float* x, float* y, int64_t rows, int64_t t_in, int64_t t_out, int64_t off,
                         cudaStream_t stream) {
  using S = RsShape<UP, DOWN, D, FR>;
  auto kern = resample_frames_kernel<UP, DOWN, D, FR>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM);
  if (e != cudaSuccess) return (int)e;
  const int64_t frames = (t_out + UP - 1) / UP;
  const int64_t tiles = (frames + S::T::NF - 1) / S::T::NF;
  if (tiles > 0x7fffffffLL) return MPCG_ERANGE;
  for (int64_t r0 = 0; r0 < rows; r0 += 65535) {                      // (rows are a grid dimension: blocks of 65 535)
    const int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
    dim3 grid((unsigned)tiles, (unsigned)nr);
    kern<<<grid, kRsThreads, S::SMEM, stream>>>(x + r0 * t_in, y + r0 * t_out, (long long)t_in, (long long)t_out, (long long)off);
  }
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}

}  // namespace mpcg

extern "C" int mpcg_resample_f32(const float* x, float* y, int64_t rows, int64_t t_in, int64_t t_out,
                                 const float* taps, int up, int down, int taps_per_phase, int64_t offset,
                                 void* stream_) {
  using namespace mpcg;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rows < 0 || t_in < 0 || t_out < 0 || up < 1 || down < 1 || taps_per_phase < 1 || !taps) return MPCG_EINVAL;
  if (rows == 0 || t_out == 0) return MPCG_OK;
  if (!x || !y) return MPCG_EINVAL;
  // rows beyond the grid-y limit are processed in slabs
  const int64_t slab = 65535;
  for (int64_t r0 = 0; r0 < rows; r0 += slab) {
    const int64_t nr = rows - r0 < slab ? rows - r0 : slab;
    const float* xs = x + r0 * t_in;
    float* ys = y + r0 * t_out;
    int rc;
    const int D = taps_per_phase;
#define MPCG_RS_CASE(U, DN, DD, FR)                                                            \
  if (up == U && down == DN && D == DD && rs_taps_match<U, DN, DD>(taps, offset)) {             \
    rc = launch_frames<U, DN, DD, FR>(xs, ys, nr, t_in, t_out, offset, stream);                 \
    if (rc != MPCG_OK) return rc;                                                               \
    continue;                                                                                   \
  }
    // tensor-path (torchaudio sinc/Hann) shapes: 2k->16k, 2k->4125, 4k->4125
    MPCG_RS_CASE(8, 1, 15, 4)
    MPCG_RS_CASE(33, 16, 30, 1)
    MPCG_RS_CASE(33, 32, 46, 1)
    // NumPy-path (SciPy Kaiser resample_poly) shapes for the same three ratios
    MPCG_RS_CASE(8, 1, 22, 4)
    MPCG_RS_CASE(33, 16, 36, 1)
    MPCG_RS_CASE(33, 32, 52, 1)
#undef MPCG_RS_CASE
    // generic fallback
    if ((int64_t)up * D > kRsGenericMaxTaps) return MPCG_ERANGE;
    int fpc = kRsGenericOut / up;
    if (fpc < 1) fpc = 1;
    const int64_t nin = (int64_t)(fpc - 1) * down + D;
    const size_t smem = (size_t)((int64_t)up * D + nin) * sizeof(float);
    if (smem > 200 * 1024) return MPCG_ERANGE;
    RsGenericTaps tp;
    for (int i = 0; i < up * D; ++i) tp.g[i] = taps[i];
    cudaError_t e = cudaFuncSetAttribute(resample_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int64_t frames = (t_out + up - 1) / up;
    const int64_t tiles = (frames + fpc - 1) / fpc;
    if (tiles > 0x7fffffffLL) return MPCG_ERANGE;
    dim3 grid((unsigned)tiles, (unsigned)nr);
    resample_generic_kernel<<<grid, kRsThreads, smem, stream>>>(xs, ys, (long long)t_in, (long long)t_out,
                                                                 (long long)offset, up, down, D, fpc, tp);
    MPCG_LAUNCH_CHECK();
  }
  return MPCG_OK;
}
