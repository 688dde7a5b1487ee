// extern "C" entries of the augmentation chain (augment/torchaug.py:24-111):
//   mpcg_aug_stage_f32     one stage of augment_pcg_batch -- transform, per-row Bernoulli blend, re-normalise --
//                          or the bare transform of add_white_noise / sinusoidal_envelope / baseline_wander
//   mpcg_aug_warp_f32      amplitude_warp: per-row 65-tap smoothing FIR over a reflect-padded row
//   mpcg_aug_eq_mix_f32    the tail of parametric_eq: N(N(coloured)/50 + N(x)), optionally blended and re-normalised
//
// One CTA per row.  A stage that re-normalises sweeps its row twice (statistics, then the affine map); the second
// sweep re-reads the row from L2 (a 4 s window at 16 kHz is 256 KB, a full wave of CTAs keeps < 126 MB live), so HBM
// sees one read and one write per stage.  The transform is recomputed in the second sweep rather than stored, which
// is why the in-kernel noise is a counter-based Philox stream: (seed, row, sample) always yields the same value.
#include "aug.cuh"

namespace mpcg {

template <bool NORMALISE>
__global__ void __launch_bounds__(kAgThreads)
aug_stage_kernel(const float* __restrict__ x, float* __restrict__ y, long long t, StageArgs a) {
  __shared__ double dscr[32];
  __shared__ float fscr[32];
  const long long row = blockIdx.x;
  const float* xr = x + row * t;
  float* yr = y + row * t;
  const float* p = a.rowp ? a.rowp + row * 8 : nullptr;
  const float* nz = a.noise ? a.noise + row * t : nullptr;
  const bool on = a.op != MPCG_AUG_IDENTITY && (a.mask == nullptr || a.mask[row] != 0.f);
  const bool philox = on && a.op == MPCG_AUG_NOISE && nz == nullptr;
  const int tid = threadIdx.x;
  const long long nq = (t + 3) >> 2;                                 // groups of four samples
  double mean = 0.0, inv_peak = 1.0;
  const bool norm_row = NORMALISE && (!a.norm_masked_only || on);     // (block-uniform)
  if (norm_row) {
    RowStats st;
    stats_init(st);
    for (long long q = tid; q < nq; q += kAgThreads) {
      float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (philox) z4 = philox_normal4(a.seed, a.stream, row, q);
      const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long i = 4 * q + k;
        if (i < t) {
          const float v = xr[i];
          stats_add(st, on ? stage_value(a, p, nz, row, i, v, zz[k]) : v);
        }
      }
    }
    stats_finish(st, t, mean, inv_peak, dscr, fscr);
  }
  const bool vec = ((reinterpret_cast<uintptr_t>(xr) | reinterpret_cast<uintptr_t>(yr)) & 15u) == 0;
  for (long long q = tid; q < nq; q += kAgThreads) {
    float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (philox) z4 = philox_normal4(a.seed, a.stream, row, q);
    const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
    if (vec && 4 * q + 3 < t) {                                      // whole group: 128-bit load and store
      const float4 v4 = *reinterpret_cast<const float4*>(xr + 4 * q);
      const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
      float ww[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float w = on ? stage_value(a, p, nz, row, 4 * q + k, vv[k], zz[k]) : vv[k];
        ww[k] = norm_row ? norm_apply(w, mean, inv_peak) : w;
      }
      *reinterpret_cast<float4*>(yr + 4 * q) = make_float4(ww[0], ww[1], ww[2], ww[3]);
      continue;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = 4 * q + k;
      if (i < t) {
        const float v = xr[i];
        const float w = on ? stage_value(a, p, nz, row, i, v, zz[k]) : v;
        yr[i] = norm_row ? norm_apply(w, mean, inv_peak) : w;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- amplitude warp
constexpr int kWpTile = 4096;
constexpr int kWpMaxTaps = 257;
constexpr int kWpOut = 8;                            // consecutive outputs per thread (register tile)
__host__ __device__ constexpr int wp_skew(int m) { return m + (m >> 5); }   // lanes step 8 samples: one pad word per 32

// Each thread computes kWpOut consecutive outputs from a sliding register window: per tap one new sample and one tap
// are read from shared memory for kWpOut FFMAs (the first version read two words per FFMA and was LDS bound).
__global__ void __launch_bounds__(256)
aug_warp_kernel(const float* __restrict__ x, float* __restrict__ y, long long t, const float* __restrict__ curves,
                int ntaps) {
  extern __shared__ float wp_smem[];                // skewed [kWpTile + ntaps - 1 + kWpOut] samples, then [ntaps] taps
  float* xs = wp_smem;
  const int span = kWpTile + ntaps - 1;
  float* taps = wp_smem + wp_skew(span + kWpOut) + 1;
  const long long row = blockIdx.y;
  const long long i0 = (long long)blockIdx.x * kWpTile;
  const float* xr = x + row * t;
  const int half = ntaps / 2;
  for (int k = threadIdx.x; k < ntaps; k += 256) taps[k] = curves[row * ntaps + k];
  for (int m = threadIdx.x; m < span + kWpOut; m += 256) {
    long long j = i0 + m - half;                    // index into the un-padded row, reflect (no edge repeat)
    if (j < 0) j = -j;
    if (j >= t) j = 2 * (t - 1) - j;
    xs[wp_skew(m)] = (m < span && j >= 0 && j < t) ? xr[j] : 0.f;
  }
  __syncthreads();
  for (int o0 = threadIdx.x * kWpOut; o0 < kWpTile; o0 += 256 * kWpOut) {
    if (i0 + o0 >= t) break;
    float acc[kWpOut], win[kWpOut];
#pragma unroll
    for (int u = 0; u < kWpOut; ++u) { acc[u] = 0.f; win[u] = xs[wp_skew(o0 + u)]; }
    int k = 0;
    for (; k + kWpOut <= ntaps; k += kWpOut) {       // kWpOut taps per trip: the window rotates through its registers
#pragma unroll
      for (int s = 0; s < kWpOut; ++s) {
        const float tp = taps[k + s];
#pragma unroll
        for (int u = 0; u < kWpOut; ++u) acc[u] = fmaf(tp, win[(u + s) % kWpOut], acc[u]);
        win[s] = xs[wp_skew(o0 + k + s + kWpOut)];
      }
    }
    for (; k < ntaps; ++k) {                          // remaining taps
      const float tp = taps[k];
#pragma unroll
      for (int u = 0; u < kWpOut; ++u) acc[u] = fmaf(tp, xs[wp_skew(o0 + k + u)], acc[u]);
    }
#pragma unroll
    for (int u = 0; u < kWpOut; ++u)
      if (i0 + o0 + u < t) y[row * t + i0 + o0 + u] = acc[u];
  }
}

// ---------------------------------------------------------------------------------------------- EQ tail
// e = N( N(c)/50 + N(x) );  MIX_ONLY: y = e.  Otherwise y = N( m ? e : x ).
template <bool MIX_ONLY>
__global__ void __launch_bounds__(kAgThreads)
aug_eq_mix_kernel(const float* __restrict__ x, const float* __restrict__ c, float* __restrict__ y, long long t,
                  const float* __restrict__ mask) {
  __shared__ double dscr[32];
  __shared__ float fscr[32];
  const long long row = blockIdx.x;
  const float* xr = x + row * t;
  const float* cr = c + row * t;
  float* yr = y + row * t;
  const int tid = threadIdx.x;
  const bool on = MIX_ONLY || mask == nullptr || mask[row] != 0.f;
  double mx = 0.0, ix = 1.0, mc = 0.0, ic = 1.0, mv = 0.0, iv = 1.0, mw = 0.0, iw = 1.0;
  if (on) {
    RowStats sx, sc;
    stats_init(sx); stats_init(sc);
    for (long long i = tid; i < t; i += kAgThreads) { stats_add(sx, xr[i]); stats_add(sc, cr[i]); }
    stats_finish(sx, t, mx, ix, dscr, fscr);
    stats_finish(sc, t, mc, ic, dscr, fscr);
    RowStats sv;
    stats_init(sv);
    for (long long i = tid; i < t; i += kAgThreads)
      stats_add(sv, __fadd_rn(__fdiv_rn(norm_apply(cr[i], mc, ic), 50.f), norm_apply(xr[i], mx, ix)));
    stats_finish(sv, t, mv, iv, dscr, fscr);
  }
  if (MIX_ONLY) {
    for (long long i = tid; i < t; i += kAgThreads)
      yr[i] = norm_apply(__fadd_rn(__fdiv_rn(norm_apply(cr[i], mc, ic), 50.f), norm_apply(xr[i], mx, ix)), mv, iv);
    return;
  }
  RowStats sw;
  stats_init(sw);
  for (long long i = tid; i < t; i += kAgThreads) {
    const float w = on ? norm_apply(__fadd_rn(__fdiv_rn(norm_apply(cr[i], mc, ic), 50.f), norm_apply(xr[i], mx, ix)), mv, iv)
                       : xr[i];
    stats_add(sw, w);
  }
  stats_finish(sw, t, mw, iw, dscr, fscr);
  for (long long i = tid; i < t; i += kAgThreads) {
    const float w = on ? norm_apply(__fadd_rn(__fdiv_rn(norm_apply(cr[i], mc, ic), 50.f), norm_apply(xr[i], mx, ix)), mv, iv)
                       : xr[i];
    yr[i] = norm_apply(w, mw, iw);
  }
}

}  // namespace mpcg

extern "C" int mpcg_aug_stage_f32(const float* x, float* y, int64_t rows, int64_t t, int op, float fs,
                                  const float* rowp, const float* noise, const float* mask, int normalise,
                                  uint64_t seed, uint64_t stream_id, void* stream) {
  using namespace mpcg;
  if (rows < 0 || t < 0) return MPCG_EINVAL;
  if (op < MPCG_AUG_IDENTITY || op > MPCG_AUG_SELECT) return MPCG_EINVAL;
  if (rows == 0 || t == 0) return MPCG_OK;                    // nothing to do (empty tensors carry null pointers)
  if (op != MPCG_AUG_IDENTITY && op != MPCG_AUG_SELECT && !rowp) return MPCG_EINVAL;
  if (op == MPCG_AUG_SELECT && !noise) return MPCG_EINVAL;
  if ((op == MPCG_AUG_SINE_MUL || op == MPCG_AUG_SINE_ADD) && !(fs > 0.f)) return MPCG_EINVAL;
  if (rows == 0 || t == 0) return MPCG_OK;
  if (!x || !y) return MPCG_EINVAL;
  if (rows > 0x7fffffffLL) return MPCG_ERANGE;
  StageArgs a;
  a.op = op; a.fs = fs; a.rowp = rowp; a.noise = noise; a.mask = mask; a.seed = seed; a.stream = stream_id;
  a.norm_masked_only = (normalise == 2) ? 1 : 0;
  if (normalise)
    aug_stage_kernel<true><<<(unsigned)rows, kAgThreads, 0, (cudaStream_t)stream>>>(x, y, (long long)t, a);
  else
    aug_stage_kernel<false><<<(unsigned)rows, kAgThreads, 0, (cudaStream_t)stream>>>(x, y, (long long)t, a);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}

extern "C" int mpcg_aug_warp_f32(const float* x, float* y, int64_t rows, int64_t t, const float* curves, int ntaps,
                                 void* stream) {
  using namespace mpcg;
  if (rows < 0 || t < 0 || ntaps < 1 || (ntaps & 1) == 0) return MPCG_EINVAL;
  if (ntaps > kWpMaxTaps) return MPCG_ERANGE;
  if (rows == 0 || t == 0) return MPCG_OK;
  if (!x || !y || !curves) return MPCG_EINVAL;
  if (t <= ntaps / 2) return MPCG_EINVAL;                           // reflect padding needs pad < length
  const size_t smem = (size_t)(wp_skew(kWpTile + ntaps - 1 + kWpOut) + 1 + ntaps) * sizeof(float);
  for (int64_t r0 = 0; r0 < rows; r0 += 65535) {                      // (rows are a grid dimension: blocks of 65 535)
    const int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
    dim3 grid((unsigned)((t + kWpTile - 1) / kWpTile), (unsigned)nr);
    aug_warp_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x + r0 * t, y + r0 * t, (long long)t, curves + r0 * ntaps, ntaps);
  }
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}

extern "C" int mpcg_aug_eq_mix_f32(const float* x, const float* coloured, float* y, int64_t rows, int64_t t,
                                   const float* mask, int mix_only, void* stream) {
  using namespace mpcg;
  if (rows < 0 || t < 0) return MPCG_EINVAL;
  if (rows == 0 || t == 0) return MPCG_OK;
  if (!x || !coloured || !y) return MPCG_EINVAL;
  if (rows > 0x7fffffffLL) return MPCG_ERANGE;
  if (mix_only)
    aug_eq_mix_kernel<true><<<(unsigned)rows, kAgThreads, 0, (cudaStream_t)stream>>>(x, coloured, y, (long long)t, mask);
  else
    aug_eq_mix_kernel<false><<<(unsigned)rows, kAgThreads, 0, (cudaStream_t)stream>>>(x, coloured, y, (long long)t, mask);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}

// ---------------------------------------------------------------------------------------------- per-row draws in one launch
// The throughput mode of augment_pcg_batch (torchaug.py, fast_draws) needs, per row, the parameters of the three
// elementwise stages and four Bernoulli masks.  One thread per row draws them from Philox (keyed by the
// call's seed and the row) and writes the [3][rows][8] parameter table and the [4][rows] mask table the chain kernel
// reads -- instead of five small framework launches per call.
namespace mpcg {
struct DrawSpec {
  float scale[3][8], offset[3][8];           // table[i][row][j] = offset[i][j] + U * scale[i][j]   (scale 0: no draw)
  float prob[4];                             // mask[i][row] = U < prob[i]
};
__global__ void aug_draw_kernel(float* __restrict__ tab, float* __restrict__ masks, long long rows, const DrawSpec s,
                                unsigned long long seed, unsigned long long sid) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const uint2 key = make_uint2((unsigned)(seed ^ (sid * 0x9E3779B97F4A7C15ull)), (unsigned)((seed >> 32) ^ sid));
  float u[28];                                 // one uniform per table entry (those without a draw are not used) + 4 masks
#pragma unroll
  for (int c = 0; c < 7; ++c) {
    const uint4 r = philox4x32_10(make_uint4((unsigned)row, (unsigned)(row >> 32), (unsigned)c, 0x6d706367u), key);
    u[4 * c + 0] = (float)(r.x >> 8) * 5.9604644775390625e-08f;     // 24 bits -> [0, 1)
    u[4 * c + 1] = (float)(r.y >> 8) * 5.9604644775390625e-08f;
    u[4 * c + 2] = (float)(r.z >> 8) * 5.9604644775390625e-08f;
    u[4 * c + 3] = (float)(r.w >> 8) * 5.9604644775390625e-08f;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(u[8 * i + j], s.scale[i][j], s.offset[i][j]);
    float4* dst = reinterpret_cast<float4*>(tab + ((long long)i * rows + row) * 8);
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) masks[(long long)i * rows + row] = u[24 + i] < s.prob[i] ? 1.f : 0.f;
}
}  // namespace mpcg

extern "C" int mpcg_aug_draw_f32(float* tab, float* masks, int64_t rows, const float* scale, const float* offset,
                                 const float* prob, uint64_t seed, uint64_t sid, void* stream) {
  using namespace mpcg;
  if (rows < 0) return MPCG_EINVAL;
  if (rows == 0) return MPCG_OK;
  if (!tab || !masks || !scale || !offset || !prob || ((uintptr_t)tab & 15u)) return MPCG_EINVAL;
  DrawSpec s;
  for (int i = 0; i < 24; ++i) {
    (&s.scale[0][0])[i] = scale[i];
    (&s.offset[0][0])[i] = offset[i];
  }
  for (int i = 0; i < 4; ++i) s.prob[i] = prob[i];
  aug_draw_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(tab, masks, (long long)rows, s, seed, sid);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}
