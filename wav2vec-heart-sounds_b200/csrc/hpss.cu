// extern "C" entries of the HPSS augmentation (augment/primitives.py:88-123, whose arithmetic is librosa 0.11's
// stft / decompose.hpss / softmask / istft and scipy.ndimage.median_filter underneath):
//   mpcg_hpss_stft_f32        frames -> one-sided complex spectrum (periodic Hann, centred, zero padded)
//   mpcg_hpss_median_f32      running median of |S| along time (harmonic) or along frequency (percussive)
//   mpcg_hpss_istft_f32       soft masks -> H / P / R spectra -> inverse FFT -> windowed overlap-add
//   mpcg_hpss_finish_f32      divide by the window sum-of-squares, trim n_fft/2 on both sides
//   mpcg_hpss_mix_f32         the random re-weighting tail of hpss_recombine (two mixes, three normalisations)
//
// Layout: spectra are frame-major, S[row][frame][bin] (bin contiguous), so FFT output and the frequency median are
// unit-stride and the time median is coalesced across bins.  FFTs run in shared memory with radix-8 / radix-4 register butterflies, one frame per
// CTA; medians keep a sorted window per thread in shared memory and slide it (remove oldest, insert newest, one
// branch-free pass over the k slots),
// O(k) per output instead of a fresh selection.  Median results are bit-exact functions of the magnitudes:
// window [i - k/2, i - k/2 + k - 1], half-sample-symmetric reflection, rank k/2 (upper median for even k).
#include <stdlib.h>
#include "common.cuh"

namespace mpcg {

constexpr int kFftThreads = 256;

__device__ __forceinline__ unsigned bitrev(unsigned v, int bits) { return __brev(v) >> (32 - bits); }

// Shared-memory index padding (in float2 elements): one pad per 8 and per 64 elements keeps the stride-8 first pass and the
// bit-reversed scatter of the callers off each other's banks.  Every index into the work buffer goes through fpad().
__host__ __device__ __forceinline__ int fpad(int i) { return i + (i >> 3) + (i >> 6); }

__device__ __forceinline__ float2 fmulc(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <int E>
__device__ __forceinline__ float2 fmul_root8(float2 z) {            // z * exp(-2 pi i E / 8)
  constexpr float h = 0.70710678118654752440f;
  if (E == 0) return z;
  if (E == 1) return make_float2((z.x + z.y) * h, (z.y - z.x) * h);
  if (E == 2) return make_float2(z.y, -z.x);
  return make_float2((z.y - z.x) * h, -(z.x + z.y) * h);
}
template <int E>
__device__ __forceinline__ float2 fmul_root8c(float2 z) {           // z * exp(+2 pi i E / 8)
  constexpr float h = 0.70710678118654752440f;
  if (E == 0) return z;
  if (E == 1) return make_float2((z.x - z.y) * h, (z.x + z.y) * h);
  if (E == 2) return make_float2(-z.y, z.x);
  return make_float2(-(z.x + z.y) * h, (z.x - z.y) * h);
}

// butterflies of merged stage U on local indices >= M (decimation in time; see fft_pass)
template <int Q, bool INV, int U, int M>
struct FftStage {
  static __device__ __forceinline__ void run(float2 (&r)[1 << Q], const float2 (&base)[Q]) {
    if constexpr (M < (1 << Q)) {
      if constexpr ((M & (1 << U)) == 0) {
        constexpr int K = M & ((1 << U) - 1);
        constexpr int E = (K << (2 - U)) & 3;
        const float2 w = INV ? fmul_root8c<E>(base[U]) : fmul_root8<E>(base[U]);
        const float2 a = r[M], b = fmulc(r[M + (1 << U)], w);
        r[M] = make_float2(a.x + b.x, a.y + b.y);
        r[M + (1 << U)] = make_float2(a.x - b.x, a.y - b.y);
      }
      FftStage<Q, INV, U, M + 1>::run(r, base);
    }
  }
};

// Stages s .. s+Q-1 of the decimation-in-time transform in one pass: a radix-2^Q butterfly held in registers.  Stage
// s+u needs exp(-+2 pi i (p + k 2^s) / 2^(s+u+1)) = base_u * (8th root)^(k << (2-u)), base_u = omega_(s+u+1)^p =
// base_(u+1)^2: one twiddle load per group.
template <int Q, bool INV>
__device__ __forceinline__ void fft_pass(float2* a, int log2n, const float2* __restrict__ tw, int s) {
  constexpr int R = 1 << Q;
  const int hl = log2n - 1, lg = log2n - Q;
  for (int j = threadIdx.x; j < (1 << lg); j += kFftThreads) {
    const int p = j & ((1 << s) - 1);
    const int i0 = ((j >> s) << (s + Q)) + p;
    float2 r[R], base[Q];
#pragma unroll
    for (int m = 0; m < R; ++m) r[m] = a[fpad(i0 + (m << s))];
    base[Q - 1] = __ldg(tw + (p << (hl - (s + Q - 1))));
    if (INV) base[Q - 1].y = -base[Q - 1].y;
#pragma unroll
    for (int u = Q - 2; u >= 0; --u) base[u] = fmulc(base[u + 1], base[u + 1]);
    FftStage<Q, INV, 0, 0>::run(r, base);
    if constexpr (Q > 1) FftStage<Q, INV, 1, 0>::run(r, base);
    if constexpr (Q > 2) FftStage<Q, INV, 2, 0>::run(r, base);
#pragma unroll
    for (int m = 0; m < R; ++m) a[fpad(i0 + (m << s))] = r[m];
  }
  __syncthreads();
}

// In-place decimation-in-time FFT over the padded buffer a (input already bit-reversed).  tw[k] = exp(-2 pi i k / n).
// Radix-8 passes while more than four stages remain, then radix-4 / radix-2 (a 1024-point transform: 8, 8, 4, 4).
template <bool INV>
__device__ __forceinline__ void fft_shared_dir(float2* a, int log2n, const float2* __restrict__ tw) {
  int s = 0;
  while (log2n - s > 4 || log2n - s == 3) { fft_pass<3, INV>(a, log2n, tw, s); s += 3; }
  while (log2n - s >= 2) { fft_pass<2, INV>(a, log2n, tw, s); s += 2; }
  if (log2n - s == 1) fft_pass<1, INV>(a, log2n, tw, s);
}
// The same with the transform size and the stage known at compile time (n_fft 512 / 1024 / 2048, the reference's choices):
// a pass at stage S touches a[fpad(i0 + (m << S))], and because bits S .. S+Q-1 of i0 are zero the padded index is
// fpad(i0) + a CONSTANT per m -- one index computation per butterfly instead of 2^(Q+1) (the block transform is issue
// bound: 78 % of the slots busy in the inverse kernel).
template <int Q, bool INV, int LOG2N, int S>
__device__ __forceinline__ void fft_pass_s(float2* a, const float2* __restrict__ tw) {
  constexpr int R = 1 << Q;
  constexpr int hl = LOG2N - 1, lg = LOG2N - Q;
  for (int j = threadIdx.x; j < (1 << lg); j += kFftThreads) {
    const int p = j & ((1 << S) - 1);
    const int i0 = ((j >> S) << (S + Q)) + p;
    float2* a0 = a + fpad(i0);
    float2 r[R], base[Q];
#pragma unroll
    for (int m = 0; m < R; ++m) r[m] = a0[fpad(m << S)];
    base[Q - 1] = __ldg(tw + (p << (hl - (S + Q - 1))));
    if (INV) base[Q - 1].y = -base[Q - 1].y;
#pragma unroll
    for (int u = Q - 2; u >= 0; --u) base[u] = fmulc(base[u + 1], base[u + 1]);
    FftStage<Q, INV, 0, 0>::run(r, base);
    if constexpr (Q > 1) FftStage<Q, INV, 1, 0>::run(r, base);
    if constexpr (Q > 2) FftStage<Q, INV, 2, 0>::run(r, base);
#pragma unroll
    for (int m = 0; m < R; ++m) a0[fpad(m << S)] = r[m];
  }
  __syncthreads();
}
template <bool INV, int LOG2N, int S = 0>
__device__ __forceinline__ void fft_static(float2* a, const float2* __restrict__ tw) {
  if constexpr (LOG2N - S > 4 || LOG2N - S == 3) {
    fft_pass_s<3, INV, LOG2N, S>(a, tw);
    fft_static<INV, LOG2N, S + 3>(a, tw);
  } else if constexpr (LOG2N - S >= 2) {
    fft_pass_s<2, INV, LOG2N, S>(a, tw);
    fft_static<INV, LOG2N, S + 2>(a, tw);
  } else if constexpr (LOG2N - S == 1) {
    fft_pass_s<1, INV, LOG2N, S>(a, tw);
  }
}
template <bool INV>
__device__ __forceinline__ void fft_shared_any(float2* a, int log2n, const float2* __restrict__ tw) {
  switch (log2n) {
    case 9: fft_static<INV, 9>(a, tw); break;
    case 10: fft_static<INV, 10>(a, tw); break;
    case 11: fft_static<INV, 11>(a, tw); break;
    default: fft_shared_dir<INV>(a, log2n, tw);
  }
}
__device__ __forceinline__ void fft_shared(float2* a, int n, int log2n, bool inverse, const float2* __restrict__ tw) {
  (void)n;
  if (inverse) fft_shared_any<true>(a, log2n, tw); else fft_shared_any<false>(a, log2n, tw);
}

// ---------------------------------------------------------------------------------------------- STFT
// Two real frames per complex FFT: z = a + i b  ->  A[k] = (Z[k] + conj(Z[N-k])) / 2,  B[k] = (Z[k] - conj(Z[N-k])) / (2i).
__global__ void __launch_bounds__(kFftThreads)
hpss_stft_kernel(const float* __restrict__ x, float2* __restrict__ spec, long long t, int n_fft, int log2n, int hop,
                 int frames, const float* __restrict__ window, const float2* __restrict__ tw) {
  extern __shared__ float2 fft_buf[];
  const int frame = 2 * blockIdx.x;                               // this CTA: frames `frame` and `frame + 1`
  const bool two = frame + 1 < frames;
  const long long row = blockIdx.y;
  const float* xr = x + row * t;
  const long long base = (long long)frame * hop - n_fft / 2;
  for (int i = threadIdx.x; i < n_fft; i += kFftThreads) {
    const long long j = base + i, j2 = j + hop;
    const float w = __ldg(window + i);
    const float va = (j >= 0 && j < t) ? xr[j] * w : 0.f;
    const float vb = (two && j2 >= 0 && j2 < t) ? xr[j2] * w : 0.f;
    fft_buf[fpad(bitrev((unsigned)i, log2n))] = make_float2(va, vb);
  }
  __syncthreads();
  fft_shared(fft_buf, n_fft, log2n, false, tw);
  const int bins = n_fft / 2 + 1;
  float2* out_a = spec + ((long long)row * frames + frame) * bins;
  float2* out_b = out_a + bins;
  for (int k = threadIdx.x; k < bins; k += kFftThreads) {
    const float2 z = fft_buf[fpad(k)], zc = fft_buf[fpad((n_fft - k) & (n_fft - 1))];     // Z[k], Z[N-k] (N-0 -> 0)
    out_a[k] = make_float2(0.5f * (z.x + zc.x), 0.5f * (z.y - zc.y));
    if (two) out_b[k] = make_float2(0.5f * (z.y + zc.y), 0.5f * (zc.x - z.x));
  }
}

// ---------------------------------------------------------------------------------------------- medians
constexpr int kMedThreads = 128;
constexpr int kMedMaxK = 64;

__device__ __forceinline__ int reflect_idx(int i, int n) {       // d c b a | a b c d | d c b a
  if (n == 1) return 0;
  const int period = 2 * n;
  i %= period;
  if (i < 0) i += period;
  return i < n ? i : period - 1 - i;
}

// Each thread owns one line (a bin for the time median, a frame for the frequency median) and walks along it.
// The current window is kept sorted; ring[kk][thread] (shared) holds the same magnitudes in arrival order, so the sample
// that leaves is not fetched and rooted again.  The slide -- drop one copy of `gone`, insert `come` -- is ONE uniform pass
// over the slots with selects only (no data-dependent loop, the lanes of a warp never diverge):
//   b[p]   = a[p] < gone ? a[p] : a[p+1]                 (the window without the first copy of `gone`;  a[k] = +inf)
//   new[p] = max(b[p-1], min(b[p], come))                (insertion into a sorted list;  b[-1] = -inf)
// Windows of up to 32 samples live in REGISTERS (K slots, K a multiple of 4): the k real samples are framed by -inf
// sentinels below and +inf above so that the median always sits in slot K/2 -- no dynamic register index.  Larger
// windows run the same pass over a shared-memory column per thread.
struct MedLine {
  const float2* src;
  float* dst;
  int len;
  long long stride;
  __device__ __forceinline__ float mag(int i) const {
    const int j = (unsigned)i < (unsigned)len ? i : reflect_idx(i, len);
    const float2 v = src[(long long)j * stride];
    return sqrtf(v.x * v.x + v.y * v.y);
  }
};

// first window (indices -left .. -left + k - 1) insertion-sorted into sw[], arrival order into ring[]
__device__ __forceinline__ void med_first_window(const MedLine& ln, float* sw, float* ring, int k, int left) {
  for (int q = 0; q < k; ++q) {
    const float v = ln.mag(q - left);
    ring[q * kMedThreads] = v;
    int p = q;
    while (p > 0 && sw[(p - 1) * kMedThreads] > v) { sw[p * kMedThreads] = sw[(p - 1) * kMedThreads]; --p; }
    sw[p * kMedThreads] = v;
  }
}

__device__ __forceinline__ bool med_setup(MedLine& ln, const float2* spec, float* out, int frames, int bins, int along_time,
                                          long long rows) {
  // lines are numbered across the whole batch (a CTA's 128 threads may straddle two rows): no CTA is left with the one
  // or two lines that 513 bins leave over after four full CTAs -- a lone lane costs a whole warp's issue slots
  const long long per_row = along_time ? bins : frames;
  const long long gline = (long long)blockIdx.x * kMedThreads + threadIdx.x;
  if (gline >= rows * per_row) return false;
  const long long row = gline / per_row;
  const int line = (int)(gline - row * per_row);
  ln.len = along_time ? frames : bins;
  ln.stride = along_time ? bins : 1;
  const long long off = row * frames * bins + (along_time ? line : (long long)line * bins);
  ln.src = spec + off;
  ln.dst = out + off;
  return true;
}

template <int K>
__device__ __forceinline__ void med_slide(float (&a)[K], float gone, float come) {
  float below_b = -INFINITY;
#pragma unroll
  for (int p = 0; p < K; ++p) {
    const float b = a[p] < gone ? a[p] : (p + 1 < K ? a[p + 1] : INFINITY);
    a[p] = fmaxf(below_b, fminf(b, come));
    below_b = b;
  }
}

// Frequency direction (a thread owns a frame and walks along its bins): neighbouring threads are a whole frame apart in
// memory, so the magnitudes are staged through shared memory -- every warp loads 32 consecutive bins of one frame
// (256 contiguous bytes), the medians of a chunk go back the same way.  The walk consumes the reflected input stream
// position by position: stream position s is bin s - k/2, output i is complete once position i + k - 1 has arrived.
// CH = bins staged per round (32 or 16: a warp loads 32 / CH frames at a time).  Windows of up to 20 samples take 16: the
// tiles shrink to 8 KB each, 25 KB per CTA instead of 42, nine CTAs per SM instead of five (1.28 -> 1.12 ms at k = 17);
// longer windows keep 32 (their first-window scratch sets the size of the output tile anyway).
template <int K, int CH>
__global__ void __launch_bounds__(kMedThreads)
hpss_median_freq_kernel(const float2* __restrict__ spec, float* __restrict__ out, int frames, int bins, int k) {
  extern __shared__ float med_sorted[];                          // ring [k][T] | in [32][T+1] | out [32][T+1] (first-window sort inside out)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* ring = med_sorted + tid;
  float* tin = med_sorted + k * kMedThreads;
  float* tout = tin + CH * (kMedThreads + 1);
  float* sw = tout + tid;                                        // first-window scratch: dead before tout is written
  const long long row = blockIdx.y;
  const int f0 = blockIdx.x * kMedThreads;
  const float2* src = spec + ((long long)row * frames + f0) * bins;
  float* dst = out + ((long long)row * frames + f0) * bins;
  const int nf = min(kMedThreads, frames - f0);                  // frames of this CTA
  const int left = k / 2, below = K / 2 - left;
  // stage stream positions [s0, s0 + 32): magnitude of bin reflect(s - left) of every frame of the CTA
  constexpr int FPW = 32 / CH;                              // frames a warp stages per iteration
  const int cl = lane % CH, fsub = lane / CH;
  auto stage_in = [&](int s0) {
    const int j = s0 + cl - left;
    const int b = (unsigned)j < (unsigned)bins ? j : reflect_idx(j, bins);
    for (int f = warp * FPW + fsub; f < kMedThreads; f += kMedThreads / 32 * FPW) {
      float m = 0.f;
      if (f < nf) { const float2 v = src[(long long)f * bins + b]; m = sqrtf(v.x * v.x + v.y * v.y); }
      tin[cl * (kMedThreads + 1) + f] = m;
    }
  };
  // first window: stream positions 0 .. k-1, one or two staged chunks
  for (int q0 = 0; q0 < k; q0 += CH) {
    if (q0) __syncthreads();
    stage_in(q0);
    __syncthreads();
    for (int q = q0; q < k && q < q0 + CH; ++q) {
      const float v = tin[(q - q0) * (kMedThreads + 1) + tid];
      ring[q * kMedThreads] = v;
      int p = q;
      while (p > 0 && sw[(p - 1) * kMedThreads] > v) { sw[p * kMedThreads] = sw[(p - 1) * kMedThreads]; --p; }
      sw[p * kMedThreads] = v;
    }
  }
  float a[K];
#pragma unroll
  for (int p = 0; p < K; ++p) {
    const int q = p - below;
    a[p] = q < 0 ? -INFINITY : (q < k ? sw[q * kMedThreads] : INFINITY);
  }
  if (tid < nf) dst[(long long)tid * bins] = a[K / 2];             // output 0
  int slot = 0;
  const int total = bins + k - 1;                                 // stream length
  for (int s0 = k; s0 < total; s0 += CH) {
    __syncthreads();
    stage_in(s0);
    __syncthreads();
    const int n = min(CH, total - s0);
    for (int c = 0; c < n; ++c) {
      const float come = tin[c * (kMedThreads + 1) + tid];
      const float gone = ring[slot * kMedThreads];
      ring[slot * kMedThreads] = come;
      slot = slot + 1 == k ? 0 : slot + 1;
      med_slide<K>(a, gone, come);
      tout[c * (kMedThreads + 1) + tid] = a[K / 2];                // output s0 + c - k + 1
    }
    __syncthreads();
    const int i0 = s0 - k + 1;
    for (int f = warp * FPW + fsub; f < nf; f += kMedThreads / 32 * FPW)
      if (cl < n) dst[(long long)f * bins + i0 + cl] = tout[cl * (kMedThreads + 1) + f];
  }
}

template <int K>
__global__ void __launch_bounds__(kMedThreads)
hpss_median_reg_kernel(const float2* __restrict__ spec, float* __restrict__ out, int frames, int bins, int k, int along_time,
                       long long rows) {
  extern __shared__ float med_sorted[];                          // [k][kMedThreads] first window + [k][kMedThreads] ring
  MedLine ln;
  if (!med_setup(ln, spec, out, frames, bins, along_time, rows)) return;
  float* sw = med_sorted + threadIdx.x;
  float* ring = med_sorted + k * kMedThreads + threadIdx.x;
  const int left = k / 2, below = K / 2 - left;                  // -inf sentinels under the window
  med_first_window(ln, sw, ring, k, left);
  float a[K];
#pragma unroll
  for (int p = 0; p < K; ++p) {
    const int q = p - below;
    a[p] = q < 0 ? -INFINITY : (q < k ? sw[q * kMedThreads] : INFINITY);
  }
  ln.dst[0] = a[K / 2];
  int slot = 0;                                                  // ring position of the oldest sample
  // the entering samples are fetched three steps ahead: a step is ~130 instructions, a miss in L2 several hundred cycles
  float ahead0 = ln.mag(k - left), ahead1 = ln.mag(k - left + 1), ahead2 = ln.mag(k - left + 2);
  for (int i = 1; i < ln.len; ++i) {
    const float gone = ring[slot * kMedThreads];
    const float come = ahead0;                                   // index i - left + k - 1
    ahead0 = ahead1; ahead1 = ahead2;
    ahead2 = ln.mag(i - left + k + 2);
    ring[slot * kMedThreads] = come;
    slot = slot + 1 == k ? 0 : slot + 1;
    med_slide<K>(a, gone, come);
    ln.dst[(long long)i * ln.stride] = a[K / 2];
  }
}

__global__ void __launch_bounds__(kMedThreads)
hpss_median_kernel(const float2* __restrict__ spec, float* __restrict__ out, int frames, int bins, int k, int along_time,
                   long long rows) {
  extern __shared__ float med_sorted[];                          // [k][kMedThreads] sorted + [k][kMedThreads] ring
  MedLine ln;
  if (!med_setup(ln, spec, out, frames, bins, along_time, rows)) return;
  float* sw = med_sorted + threadIdx.x;
  float* ring = med_sorted + k * kMedThreads + threadIdx.x;
  const int left = k / 2;
  med_first_window(ln, sw, ring, k, left);
  ln.dst[0] = sw[left * kMedThreads];
  int slot = 0;
  float ahead = ln.mag(k - left);
  for (int i = 1; i < ln.len; ++i) {
    const float gone = ring[slot * kMedThreads];
    const float come = ahead;
    ahead = ln.mag(i - left + k);
    ring[slot * kMedThreads] = come;
    slot = slot + 1 == k ? 0 : slot + 1;
    float below_b = -INFINITY, cur = sw[0];
#pragma unroll 4
    for (int p = 0; p < k; ++p) {
      const float nxt = p + 1 < k ? sw[(p + 1) * kMedThreads] : INFINITY;
      const float b = cur < gone ? cur : nxt;
      sw[p * kMedThreads] = fmaxf(below_b, fminf(b, come));
      below_b = b;
      cur = nxt;
    }
    ln.dst[(long long)i * ln.stride] = sw[left * kMedThreads];
  }
}

// ---------------------------------------------------------------------------------------------- masks + ISTFT
__device__ __forceinline__ float softmask2(float x, float ref, bool split_zeros) {
  // librosa.util.softmask with power 2:  (x/z)^2 / ((x/z)^2 + (ref/z)^2),  z = max(x, ref).  One of the two ratios is
  // exactly 1 (z / z), so only the other one is divided: two divisions instead of three.
  // The two divisions are reciprocal + multiply (MUFU.RCP, within 2 ulp = 1.2e-7 of the mask; magnitudes stay far below
  // the 2^126 bound of that form): the inverse kernel is issue bound and four IEEE divisions per bin were 14 % of it.
  const float z = fmaxf(x, ref);
  if (z < FLT_MIN) return split_zeros ? 0.5f : 0.f;
  const float q = __fdividef(fminf(x, ref), z);
  const float q2 = q * q;
  return __fdividef(x >= ref ? 1.f : q2, q2 + 1.f);
}

// A CTA transforms kIstftGroup consecutive frames of one row and overlap-adds them in shared memory first: with n_fft / hop
// = 16 or 32 overlapping frames per sample, one global atomic per sample and FRAME was the bound of this kernel; now it is
// one per sample and GROUP ((group - 1) hop + n_fft atomics instead of group * n_fft).
constexpr int kIstftSlots = 3;                      // bins per thread fetched one frame ahead (n_fft <= 1024)
constexpr int kIstftGroup = 16;                     // at most; fewer when (group - 1) hop + n_fft would not fit
__global__ void __launch_bounds__(kFftThreads, 4)
hpss_istft_kernel(const float2* __restrict__ spec, const float* __restrict__ harm, const float* __restrict__ perc,
                  float* __restrict__ acc, int n_fft, int log2n, int hop, int frames, long long acc_len,
                  float margin_h, float margin_p, const float* __restrict__ window, const float2* __restrict__ tw, int group) {
  extern __shared__ float2 fft_buf[];            // [n_fft] work + 2*[span] overlap-add
  const int bins = n_fft / 2 + 1;
  const int span = (group - 1) * hop + n_fft;
  float* oh = reinterpret_cast<float*>(fft_buf + fpad(n_fft));
  float* op = oh + span;
  const int frame0 = blockIdx.x * group;
  const int nfr = min(group, frames - frame0);
  const long long row = blockIdx.y;
  const bool split = (margin_h == 1.f && margin_p == 1.f);
  const float scale = 1.f / (float)n_fft;
  for (int i = threadIdx.x; i < 2 * span; i += kFftThreads) oh[i] = 0.f;
  // A frame's spectrum and medians are fetched into REGISTERS one frame ahead (kIstftSlots bins per thread), so their
  // DRAM latency hides behind the previous frame's transform; frames with more bins per thread take the direct path.
  const bool pre = bins <= kIstftSlots * kFftThreads;
  float2 nsv[kIstftSlots];
  float nh[kIstftSlots], np_[kIstftSlots];
  auto fetch = [&](int f) {
    const long long off = ((long long)row * frames + frame0 + f) * bins;
#pragma unroll
    for (int u = 0; u < kIstftSlots; ++u) {
      const int k = threadIdx.x + u * kFftThreads;
      if (k < bins) { nsv[u] = spec[off + k]; nh[u] = harm[off + k]; np_[u] = perc[off + k]; }
    }
  };
  auto scatter = [&](int k, float2 sv, float h, float p) {
    const float wh = softmask2(h, p * margin_h, split), wp = softmask2(p, h * margin_p, split);
    // ONE inverse FFT per frame: H and P are spectra of real signals, so ifft(H + i P) = h + i p (h, p real).  The
    // residual needs no transform at all: istft is linear and istft(stft(x)) = x, hence r = x - h - p (finish kernel).
    float2 hv = make_float2(sv.x * wh, sv.y * wh), pv = make_float2(sv.x * wp, sv.y * wp);
    if (k == 0 || k == n_fft / 2) { hv.y = 0.f; pv.y = 0.f; }         // irfft ignores the imaginary part of DC / Nyquist
    fft_buf[fpad(bitrev((unsigned)k, log2n))] = make_float2(hv.x - pv.y, hv.y + pv.x);
    if (k > 0 && k < n_fft / 2) fft_buf[fpad(bitrev((unsigned)(n_fft - k), log2n))] = make_float2(hv.x + pv.y, pv.x - hv.y);
  };
  if (pre && nfr > 0) fetch(0);
  for (int f = 0; f < nfr; ++f) {
    __syncthreads();                             // the previous frame's overlap-add has read the work buffer
    if (pre) {
#pragma unroll
      for (int u = 0; u < kIstftSlots; ++u) {
        const int k = threadIdx.x + u * kFftThreads;
        if (k < bins) scatter(k, nsv[u], nh[u], np_[u]);
      }
      if (f + 1 < nfr) fetch(f + 1);
    } else {
      const long long off = ((long long)row * frames + frame0 + f) * bins;
      for (int k = threadIdx.x; k < bins; k += kFftThreads) scatter(k, spec[off + k], harm[off + k], perc[off + k]);
    }
    __syncthreads();
    fft_shared(fft_buf, n_fft, log2n, true, tw);
    // frames are added one after the other (the barrier above separates them), each thread on its own samples of the frame
    for (int i = threadIdx.x; i < n_fft; i += kFftThreads) {
      const float w = scale * __ldg(window + i);
      const float2 v = fft_buf[fpad(i)];
      oh[f * hop + i] += v.x * w;
      op[f * hop + i] += v.y * w;
    }
  }
  __syncthreads();
  const int used = (nfr - 1) * hop + n_fft;
  float* dst_h = acc + ((long long)row * 3 + 0) * acc_len + (long long)frame0 * hop;
  float* dst_p = acc + ((long long)row * 3 + 1) * acc_len + (long long)frame0 * hop;
  for (int i = threadIdx.x; i < used; i += kFftThreads) {
    atomicAdd(dst_h + i, oh[i]);
    atomicAdd(dst_p + i, op[i]);
  }
}

// y[row][0..1][i] = acc / wsum (istft's normalisation and centre trim); y[row][2][i] = x[row][i] - harmonic - percussive
__global__ void hpss_finish3_kernel(const float* __restrict__ acc, const float* __restrict__ wsum, const float* __restrict__ x,
                                    float* __restrict__ y, long long acc_len, long long n_out, long long t, int pad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = blockIdx.y;
  if (i >= n_out) return;
  const float w = __ldg(wsum + i + pad);
  float h = acc[(row * 3 + 0) * acc_len + i + pad], p = acc[(row * 3 + 1) * acc_len + i + pad];
  if (w > FLT_MIN) { h /= w; p /= w; }
  y[(row * 3 + 0) * n_out + i] = h;
  y[(row * 3 + 1) * n_out + i] = p;
  y[(row * 3 + 2) * n_out + i] = x[row * t + i] - (h + p);
}

__global__ void hpss_finish_kernel(const float* __restrict__ acc, const float* __restrict__ wsum, float* __restrict__ y,
                                   long long acc_len, long long n_out, int pad, long long lines) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long line = blockIdx.y;
  if (i >= n_out || line >= lines) return;
  const float w = __ldg(wsum + i + pad);
  const float v = acc[line * acc_len + i + pad];
  y[line * n_out + i] = (w > FLT_MIN) ? v / w : v;
}

// ---------------------------------------------------------------------------------------------- recombination
constexpr int kMixThreads = 512;
constexpr int kMixMaxParts = 8;
struct MixArgs {
  const float* parts;        // [nparts][rows][n]
  float* out;                // [rows][n]
  long long rows, n;
  int nparts;
  float w1[kMixMaxParts], w2[kMixMaxParts], wmix;
};
struct MixStat {
  double sum;
  float lo, hi;
};
__device__ __forceinline__ void mix_finish(MixStat s, long long n, double& mean, double& inv, double* dscr, float* fscr) {
  const double tot = block_sum<kMixThreads>(s.sum, dscr);
  const float lo = block_min<kMixThreads>(s.lo, fscr);
  const float hi = block_max<kMixThreads>(s.hi, fscr);
  mean = tot / (double)n;
  const double peak = fmax((double)hi - mean, mean - (double)lo);
  inv = peak > 0.0 ? 1.0 / peak : 1.0;                          // NumPy rule: divide only when the peak is positive
}
__device__ __forceinline__ float mix_norm(float v, double mean, double inv) {
  return fminf(fmaxf((float)(((double)v - mean) * inv), -1.f), 1.f);
}

__global__ void __launch_bounds__(kMixThreads)
hpss_mix_kernel(const MixArgs a) {
  __shared__ double dscr[32];
  __shared__ float fscr[32];
  const long long row = blockIdx.x;
  const int tid = threadIdx.x;
  const float* base = a.parts + row * a.n;
  const long long pstride = a.rows * a.n;
  double pm[kMixMaxParts], pi[kMixMaxParts];
  for (int p = 0; p < a.nparts; ++p) {                           // statistics of every part
    MixStat s{0.0, INFINITY, -INFINITY};
    for (long long i = tid; i < a.n; i += kMixThreads) {
      const float v = base[p * pstride + i];
      s.sum += (double)v; s.lo = fminf(s.lo, v); s.hi = fmaxf(s.hi, v);
    }
    mix_finish(s, a.n, pm[p], pi[p], dscr, fscr);
  }
  auto raw = [&](long long i, float& m1, float& m2) {
    m1 = 0.f; m2 = 0.f;
    for (int p = 0; p < a.nparts; ++p) {
      const float v = base[p * pstride + i];
      m1 = fmaf(a.w1[p], v, m1);
      m2 = fmaf(a.w2[p], mix_norm(v, pm[p], pi[p]), m2);
    }
  };
  MixStat s1{0.0, INFINITY, -INFINITY}, s2{0.0, INFINITY, -INFINITY};
  for (long long i = tid; i < a.n; i += kMixThreads) {
    float m1, m2;
    raw(i, m1, m2);
    s1.sum += (double)m1; s1.lo = fminf(s1.lo, m1); s1.hi = fmaxf(s1.hi, m1);
    s2.sum += (double)m2; s2.lo = fminf(s2.lo, m2); s2.hi = fmaxf(s2.hi, m2);
  }
  double mean1, inv1, mean2, inv2, mean3, inv3;
  mix_finish(s1, a.n, mean1, inv1, dscr, fscr);
  mix_finish(s2, a.n, mean2, inv2, dscr, fscr);
  MixStat s3{0.0, INFINITY, -INFINITY};
  for (long long i = tid; i < a.n; i += kMixThreads) {
    float m1, m2;
    raw(i, m1, m2);
    const float v = mix_norm(m1, mean1, inv1) + a.wmix * mix_norm(m2, mean2, inv2);
    s3.sum += (double)v; s3.lo = fminf(s3.lo, v); s3.hi = fmaxf(s3.hi, v);
  }
  mix_finish(s3, a.n, mean3, inv3, dscr, fscr);
  for (long long i = tid; i < a.n; i += kMixThreads) {
    float m1, m2;
    raw(i, m1, m2);
    const float v = mix_norm(m1, mean1, inv1) + a.wmix * mix_norm(m2, mean2, inv2);
    a.out[row * a.n + i] = mix_norm(v, mean3, inv3);
  }
}

// ---------------------------------------------------------------------------------------------- 1024-point transforms, one WARP per frame
// The block FFT above spends a 1024-point frame on 256 threads and four barrier-separated passes: at n_fft = 1024 it is
// latency bound (1.3 us per frame and SM).  Here a warp owns a frame: lane l holds x[l + 32 j], j = 0..31, in REGISTERS,
//   1. a 32-point FFT over j in registers (radix-2 DIF, twiddles are immediates; the result sits in bit-reversed
//      registers, which costs nothing: register indices are compile-time),
//   2. the twiddles W_1024^(l k1) from a 32 x 32 shared table and a transpose through the warp's private shared tile
//      (row stride 33: conflict-free both ways),
//   3. a second 32-point FFT in registers: lane k1 ends with X[k1 + 32 k2], k2 = 0..31,
// no CTA barrier anywhere.  The forward kernel packs two real frames into one transform and unpacks them with one
// shuffle pair per bin (Z[N - k] lives in lane 32 - k1, register 31 - k2): 6.7 -> 3.0 ms for the 1536 window transforms of
// the composed pipeline.  Other transform sizes keep the block FFT, and so does the inverse direction: the same scheme
// there (Hermitian halves of H + i P mirrored with shuffles, a round of four frames overlap-added cooperatively) came
// out at 16.8 - 19.6 ms against 16.3 ms -- its frame is bound by the three input streams and the mask divisions, not
// by the transform, and 106 registers leave 12 - 16 warps per SM to hide them.
constexpr int kWfWarps = 4;
constexpr int kWfThreads = 32 * kWfWarps;
constexpr int kWfN = 1024;
constexpr int kWfRow = 33;                          // float2 elements per row of a warp's transpose tile
constexpr int kWfPairsPerWarp = 8;                  // forward: frame pairs a warp walks through

__host__ __device__ constexpr int wf_br5(int k) {
  return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4);
}
// d * exp(-/+ 2 pi i m / 32), m a compile-time constant after unrolling
template <bool INV>
__device__ __forceinline__ float2 wf_mul_root32(float2 d, int m) {
  float c, s;
  switch (m) {
    case 0: return d;
    case 1: c = 0.98078528040323043f; s = 0.19509032201612825f; break;
    case 2: c = 0.92387953251128674f; s = 0.38268343236508978f; break;
    case 3: c = 0.83146961230254524f; s = 0.55557023301960218f; break;
    case 4: c = 0.70710678118654752f; s = 0.70710678118654752f; break;
    case 5: c = 0.55557023301960218f; s = 0.83146961230254524f; break;
    case 6: c = 0.38268343236508978f; s = 0.92387953251128674f; break;
    case 7: c = 0.19509032201612825f; s = 0.98078528040323043f; break;
    case 8: return INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
    case 9: c = -0.19509032201612825f; s = 0.98078528040323043f; break;
    case 10: c = -0.38268343236508978f; s = 0.92387953251128674f; break;
    case 11: c = -0.55557023301960218f; s = 0.83146961230254524f; break;
    case 12: c = -0.70710678118654752f; s = 0.70710678118654752f; break;
    case 13: c = -0.83146961230254524f; s = 0.55557023301960218f; break;
    case 14: c = -0.92387953251128674f; s = 0.38268343236508978f; break;
    default: c = -0.98078528040323043f; s = 0.19509032201612825f; break;
  }
  // forward: (c - i s); inverse: (c + i s)
  return INV ? make_float2(d.x * c - d.y * s, d.x * s + d.y * c) : make_float2(d.x * c + d.y * s, d.y * c - d.x * s);
}
// in-place 32-point DFT of a[0..31]; X[k] ends in a[wf_br5(k)]
template <bool INV>
__device__ __forceinline__ void wf_fft32(float2 (&a)[32]) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
#pragma unroll
    for (int b = 0; b < 32; b += 2 * h) {
#pragma unroll
      for (int i = 0; i < h; ++i) {
        const float2 u = a[b + i], v = a[b + i + h];
        a[b + i] = make_float2(u.x + v.x, u.y + v.y);
        a[b + i + h] = wf_mul_root32<INV>(make_float2(u.x - v.x, u.y - v.y), i * (16 / h));
      }
    }
  }
}
// The whole transform: in: a[j] = x[lane + 32 j]; out: X[lane + 32 k2] in a[wf_br5(k2)].  tws: [32][32] twiddles
// W^(l k1) at tws[k1 * 32 + l] (already conjugated for the inverse); tile: this warp's [32][kWfRow] scratch.
template <bool INV>
__device__ __forceinline__ void wf_fft1024(float2 (&a)[32], const float2* __restrict__ tws, float2* __restrict__ tile, int lane) {
  wf_fft32<INV>(a);
  __syncwarp();                                      // the tile's previous readers are done
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) tile[k1 * kWfRow + lane] = fmulc(a[wf_br5(k1)], tws[k1 * 32 + lane]);
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) a[n2] = tile[lane * kWfRow + n2];
  wf_fft32<INV>(a);
}
__device__ __forceinline__ void wf_build_twiddles(float2* tws, const float2* __restrict__ tw, bool inverse) {
  for (int e = threadIdx.x; e < 1024; e += kWfThreads) {
    const int m = ((e >> 5) * (e & 31)) & 1023;      // k1 * l
    float2 w = __ldg(tw + (m & 511));                // exp(-2 pi i m / 1024), m < 512
    if (m >= 512) w = make_float2(-w.x, -w.y);
    if (inverse) w.y = -w.y;
    tws[e] = w;
  }
}

__global__ void __launch_bounds__(kWfThreads)
hpss_stft_w1024_kernel(const float* __restrict__ x, float2* __restrict__ spec, long long t, int hop, int frames,
                       const float* __restrict__ window, const float2* __restrict__ tw) {
  __shared__ float2 tws[1024];
  __shared__ float2 tiles[kWfWarps][32 * kWfRow];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = blockIdx.y;
  const float* xr = x + row * t;
  constexpr int bins = kWfN / 2 + 1;
  wf_build_twiddles(tws, tw, false);
  __syncthreads();
  const int npairs = (frames + 1) / 2;
  const int p0 = (blockIdx.x * kWfWarps + warp) * kWfPairsPerWarp;
  for (int pp = p0; pp < p0 + kWfPairsPerWarp && pp < npairs; ++pp) {
    const int frame = 2 * pp;
    const bool two = frame + 1 < frames;
    const long long base = (long long)frame * hop - kWfN / 2 + lane;
    float2 a[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const long long i0 = base + 32 * j, i1 = i0 + hop;
      const float w = __ldg(window + lane + 32 * j);
      a[j] = make_float2((i0 >= 0 && i0 < t) ? __ldg(xr + i0) * w : 0.f, (two && i1 >= 0 && i1 < t) ? __ldg(xr + i1) * w : 0.f);
    }
    wf_fft1024<false>(a, tws, tiles[warp], lane);
    // two real frames from one transform: A[k] = (Z[k] + conj Z[N-k]) / 2, B[k] = (Z[k] - conj Z[N-k]) / (2i)
    float2* out_a = spec + ((long long)row * frames + frame) * bins;
    float2* out_b = out_a + bins;
    const int src = (32 - lane) & 31;
#pragma unroll
    for (int k2 = 0; k2 <= 16; ++k2) {
      const float2 z = a[wf_br5(k2)];
      const float2 other = a[wf_br5(31 - (k2 & 15))];                 // (k2 = 16: unused for lanes > 0)
      const float ox = __shfl_sync(kFull, other.x, src), oy = __shfl_sync(kFull, other.y, src);
      const float2 self = a[wf_br5((32 - k2) & 31)];
      const float2 zc = lane ? make_float2(ox, oy) : self;
      if (k2 < 16 || lane == 0) {
        const int k = lane + 32 * k2;
        out_a[k] = make_float2(0.5f * (z.x + zc.x), 0.5f * (z.y - zc.y));
        if (two) out_b[k] = make_float2(0.5f * (z.y + zc.y), 0.5f * (zc.x - z.x));
      }
    }
  }
}

static int log2_exact(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return (1 << l) == n ? l : -1;
}

}  // namespace mpcg

extern "C" int mpcg_hpss_stft_f32(const float* x, float* spec, int64_t rows, int64_t t, int n_fft, int hop,
                                  int64_t frames, const float* window, const float* twiddle, void* stream) {
  using namespace mpcg;
  const int l2 = log2_exact(n_fft);
  if (rows < 0 || t < 0 || l2 < 1 || n_fft > 8192 || hop < 1 || frames != 1 + t / hop) return MPCG_EINVAL;
  if (rows == 0) return MPCG_OK;
  if (!x || !spec || !window || !twiddle) return MPCG_EINVAL;
  if (rows > 65535 || frames > 0x7fffffffLL) return MPCG_ERANGE;
  if (n_fft == kWfN && !getenv("MPCG_HPSS_BLOCK_FFT")) {           // one warp per frame pair (no block barriers)
    const int64_t pairs = (frames + 1) / 2;
    dim3 grid((unsigned)((pairs + kWfWarps * kWfPairsPerWarp - 1) / (kWfWarps * kWfPairsPerWarp)), (unsigned)rows);
    hpss_stft_w1024_kernel<<<grid, kWfThreads, 0, (cudaStream_t)stream>>>(x, (float2*)spec, (long long)t, hop, (int)frames,
                                                                        window, (const float2*)twiddle);
    MPCG_LAUNCH_CHECK();
    return MPCG_OK;
  }
  const size_t smem = (size_t)fpad(n_fft) * sizeof(float2);
  cudaError_t e = cudaFuncSetAttribute(hpss_stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((unsigned)((frames + 1) / 2), (unsigned)rows);       // two frames per CTA (one packed complex FFT)
  hpss_stft_kernel<<<grid, kFftThreads, smem, (cudaStream_t)stream>>>(x, (float2*)spec, (long long)t, n_fft, l2, hop,
                                                                    (int)frames, window, (const float2*)twiddle);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}

extern "C" int mpcg_hpss_median_f32(const float* spec, float* out, int64_t rows, int64_t frames, int bins, int k,
                                    int along_time, void* stream) {
  using namespace mpcg;
  if (rows < 0 || frames < 1 || bins < 1 || k < 1) return MPCG_EINVAL;
  if (k > kMedMaxK) return MPCG_ERANGE;
  if (rows == 0) return MPCG_OK;
  if (!spec || !out) return MPCG_EINVAL;
  if (rows > 65535 || frames > 0x7fffffffLL) return MPCG_ERANGE;
  const int64_t nlines = along_time ? bins : frames;
  const size_t smem = 2 * (size_t)k * kMedThreads * sizeof(float);
  dim3 grid((unsigned)((nlines + kMedThreads - 1) / kMedThreads), (unsigned)rows);
  const long long flat_ctas = (rows * nlines + kMedThreads - 1) / kMedThreads;      // lines numbered across the batch
  if (flat_ctas > 0x7fffffffLL) return MPCG_ERANGE;
  dim3 flat((unsigned)flat_ctas);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(hpss_median_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const float2* sp = (const float2*)spec;
  const int ch = (k <= 20 || bins <= 320) ? 16 : 32;              // staging chunk of the frequency direction (measured: k 17 / 30, 257 - 1025 bins)
  const size_t out_tile = (size_t)ch * (kMedThreads + 1) > (size_t)k * kMedThreads ? (size_t)ch * (kMedThreads + 1)
                                                                                  : (size_t)k * kMedThreads;
  const size_t smem_f = ((size_t)k * kMedThreads + (size_t)ch * (kMedThreads + 1) + out_tile) * sizeof(float);
#define MED_REG(KK)                                                                                                  \
  if (along_time) {                                                                                                  \
    hpss_median_reg_kernel<KK><<<flat, kMedThreads, smem, st>>>(sp, out, (int)frames, bins, k, 1, (long long)rows);     \
  } else if (ch == 16) {                                                                                             \
    cudaError_t e = cudaFuncSetAttribute(hpss_median_freq_kernel<KK, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem_f);                                                               \
    if (e != cudaSuccess) return (int)e;                                                                             \
    hpss_median_freq_kernel<KK, 16><<<grid, kMedThreads, smem_f, st>>>(sp, out, (int)frames, bins, k);                 \
  } else {                                                                                                           \
    cudaError_t e = cudaFuncSetAttribute(hpss_median_freq_kernel<KK, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem_f);                                                               \
    if (e != cudaSuccess) return (int)e;                                                                             \
    hpss_median_freq_kernel<KK, 32><<<grid, kMedThreads, smem_f, st>>>(sp, out, (int)frames, bins, k);                 \
  }
  switch ((k + 3) / 4) {
    case 1: MED_REG(4); break;
    case 2: MED_REG(8); break;
    case 3: MED_REG(12); break;
    case 4: MED_REG(16); break;
    case 5: MED_REG(20); break;
    case 6: MED_REG(24); break;
    case 7: MED_REG(28); break;
    case 8: MED_REG(32); break;
    default: hpss_median_kernel<<<flat, kMedThreads, smem, st>>>(sp, out, (int)frames, bins, k, along_time, (long long)rows);
  }
#undef MED_REG
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}

extern "C" int mpcg_hpss_istft_f32(const float* spec, const float* harm, const float* perc, float* acc, int64_t rows,
                                   int n_fft, int hop, int64_t frames, float margin_h, float margin_p,
                                   const float* window, const float* twiddle, void* stream) {
  using namespace mpcg;
  const int l2 = log2_exact(n_fft);
  if (rows < 0 || l2 < 1 || n_fft > 8192 || hop < 1 || frames < 1) return MPCG_EINVAL;
  if (rows == 0) return MPCG_OK;
  if (!spec || !harm || !perc || !acc || !window || !twiddle) return MPCG_EINVAL;
  if (rows > 65535 || frames > 0x7fffffffLL) return MPCG_ERANGE;
  const int bins = n_fft / 2 + 1;
  const long long acc_len = (long long)n_fft + (long long)hop * (frames - 1);
  const size_t fixed = (size_t)fpad(n_fft) * sizeof(float2) + 2 * (size_t)n_fft * sizeof(float);
  int group = kIstftGroup;
  while (group > 1 && fixed + 2 * (size_t)(group - 1) * hop * sizeof(float) > 160 * 1024) --group;
  const size_t smem = fixed + 2 * (size_t)(group - 1) * hop * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(hpss_istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(acc, 0, sizeof(float) * (size_t)rows * 3 * acc_len, (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((unsigned)((frames + group - 1) / group), (unsigned)rows);
  hpss_istft_kernel<<<grid, kFftThreads, smem, (cudaStream_t)stream>>>((const float2*)spec, harm, perc, acc, n_fft, l2,
                                                                     hop, (int)frames, acc_len, margin_h, margin_p,
                                                                     window, (const float2*)twiddle, group);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}

extern "C" int mpcg_hpss_finish_f32(const float* acc, const float* wsum, float* y, int64_t lines, int n_fft, int hop,
                                    int64_t frames, void* stream) {
  using namespace mpcg;
  if (lines < 0 || n_fft < 2 || hop < 1 || frames < 1) return MPCG_EINVAL;
  const long long acc_len = (long long)n_fft + (long long)hop * (frames - 1);
  const long long n_out = (long long)hop * (frames - 1);
  if (lines == 0 || n_out == 0) return MPCG_OK;
  if (!acc || !wsum || !y) return MPCG_EINVAL;
  if (lines > 65535) return MPCG_ERANGE;
  dim3 grid((unsigned)((n_out + 255) / 256), (unsigned)lines);
  hpss_finish_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(acc, wsum, y, acc_len, n_out, n_fft / 2, (long long)lines);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}

extern "C" int mpcg_hpss_finish3_f32(const float* acc, const float* wsum, const float* x, float* y, int64_t rows, int64_t t,
                                     int n_fft, int hop, int64_t frames, void* stream) {
  using namespace mpcg;
  if (rows < 0 || n_fft < 2 || hop < 1 || frames < 1 || t < 0) return MPCG_EINVAL;
  const long long acc_len = (long long)n_fft + (long long)hop * (frames - 1);
  const long long n_out = (long long)hop * (frames - 1);
  if (rows == 0 || n_out == 0) return MPCG_OK;
  if (!acc || !wsum || !x || !y || n_out > t) return MPCG_EINVAL;
  if (rows > 65535) return MPCG_ERANGE;
  dim3 grid((unsigned)((n_out + 255) / 256), (unsigned)rows);
  hpss_finish3_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(acc, wsum, x, y, acc_len, n_out, (long long)t, n_fft / 2);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}

extern "C" int mpcg_hpss_mix_f32(const float* parts, float* out, int64_t rows, int64_t n, int nparts, const float* w1,
                                 const float* w2, float wmix, void* stream) {
  using namespace mpcg;
  if (rows < 0 || n < 0 || nparts < 1 || !w1 || !w2) return MPCG_EINVAL;
  if (nparts > kMixMaxParts) return MPCG_ERANGE;
  if (rows == 0 || n == 0) return MPCG_OK;
  if (!parts || !out) return MPCG_EINVAL;
  if (rows > 0x7fffffffLL) return MPCG_ERANGE;
  MixArgs a;
  a.parts = parts; a.out = out; a.rows = rows; a.n = n; a.nparts = nparts; a.wmix = wmix;
  for (int p = 0; p < kMixMaxParts; ++p) { a.w1[p] = p < nparts ? w1[p] : 0.f; a.w2[p] = p < nparts ? w2[p] : 0.f; }
  hpss_mix_kernel<<<(unsigned)rows, kMixThreads, 0, (cudaStream_t)stream>>>(a);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}
