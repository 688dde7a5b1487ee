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
#include "common.cuh"

namespace mpcg {

constexpr int kAgThreads = 512;

// ---------------------------------------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// Four standard normals for samples 4q .. 4q+3 of a row (Box-Muller on two uniform pairs).
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned long long stream, long long row,
                                                 long long q) {
  const uint4 r = philox4x32_10(make_uint4((unsigned)q, (unsigned)(q >> 32), (unsigned)row, (unsigned)(row >> 32)),
                                make_uint2((unsigned)(seed ^ (stream * 0x9E3779B97F4A7C15ull)),
                                           (unsigned)((seed >> 32) ^ stream)));
  const float u0 = ((float)r.x + 0.5f) * 2.3283064365386963e-10f, u1 = ((float)r.y + 0.5f) * 2.3283064365386963e-10f;
  const float u2 = ((float)r.z + 0.5f) * 2.3283064365386963e-10f, u3 = ((float)r.w + 0.5f) * 2.3283064365386963e-10f;
  const float ra = sqrtf(-2.f * __logf(u0)), rb = sqrtf(-2.f * __logf(u2));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u1, &s0, &c0);
  __sincosf(6.283185307179586f * u3, &s1, &c1);
  return make_float4(ra * c0, ra * s0, rb * c1, rb * s1);
}

// ---------------------------------------------------------------------------------------------- row statistics
struct RowStats {
  double sum;
  float lo, hi;
};
__device__ __forceinline__ void stats_init(RowStats& s) { s.sum = 0.0; s.lo = INFINITY; s.hi = -INFINITY; }
__device__ __forceinline__ void stats_add(RowStats& s, float v) {
  s.sum += (double)v;
  s.lo = fminf(s.lo, v);
  s.hi = fmaxf(s.hi, v);
}
// Block-wide combine; every thread receives (mean, 1/peak) of  clip((v - mean) / max(max|v - mean|, 1e-12)).
__device__ __forceinline__ void stats_finish(RowStats s, long long t, double& mean, double& inv_peak, double* dscr,
                                             float* fscr) {
  const double tot = block_sum<kAgThreads>(s.sum, dscr);
  const float lo = block_min<kAgThreads>(s.lo, fscr);
  const float hi = block_max<kAgThreads>(s.hi, fscr);
  mean = tot / (double)t;
  const double peak = fmax((double)hi - mean, mean - (double)lo);
  inv_peak = 1.0 / fmax(peak, 1e-12);
}
__device__ __forceinline__ float norm_apply(float v, double mean, double inv_peak) {
  const float u = (float)(((double)v - mean) * inv_peak);
  return fminf(fmaxf(u, -1.f), 1.f);
}

// ---------------------------------------------------------------------------------------------- stage transforms
struct StageArgs {
  int op;                    // MPCG_AUG_*
  float fs;
  const float* rowp;         // [rows, 8] per-row parameters (see include/mpcg_b200.h)
  const float* noise;        // [rows, t] injected standard normals, or NULL -> Philox
  const float* mask;         // [rows] 0/1, or NULL -> all rows transformed
  unsigned long long seed, stream;
};

// value of the transformed sample i of this row, given the input sample v
__device__ __forceinline__ float stage_value(const StageArgs& a, const float* p, const float* nz, long long row,
                                             long long i, float v, float zphilox) {
  switch (a.op) {
    case MPCG_AUG_NOISE: {
      const float z = nz ? nz[i] : zphilox;
      return __fadd_rn(v, __fmul_rn(p[0], z));                       // x + (scale*std) * noise
    }
    case MPCG_AUG_SINE_MUL:
    case MPCG_AUG_SINE_ADD: {
      // the reference builds t = arange(T) / fs and the phases in float32; follow its rounding sequence
      const float tt = __fdiv_rn((float)i, a.fs);
      const float two_pi = 6.283185307179586f;
      const float m0 = __fmul_rn(p[0], sinf(__fmul_rn(two_pi, __fadd_rn(__fmul_rn(p[1], tt), p[2]))));
      const float m1 = __fmul_rn(p[3], sinf(__fmul_rn(two_pi, __fadd_rn(__fmul_rn(p[4], tt), p[5]))));
      const float mod = __fadd_rn(__fadd_rn(0.f, m0), m1);
      return a.op == MPCG_AUG_SINE_MUL ? __fmul_rn(v, __fadd_rn(1.f, mod)) : __fadd_rn(v, mod);
    }
    case MPCG_AUG_SELECT:
      return nz[i];                                                  // blend of two tensors (torchaug._apply)
    default:
      return v;
  }
}

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
  if (NORMALISE) {
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
  for (long long q = tid; q < nq; q += kAgThreads) {
    float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (philox) z4 = philox_normal4(a.seed, a.stream, row, q);
    const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = 4 * q + k;
      if (i < t) {
        const float v = xr[i];
        const float w = on ? stage_value(a, p, nz, row, i, v, zz[k]) : v;
        yr[i] = NORMALISE ? norm_apply(w, mean, inv_peak) : w;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- amplitude warp
constexpr int kWpTile = 4096;
constexpr int kWpMaxTaps = 257;
__global__ void __launch_bounds__(256)
aug_warp_kernel(const float* __restrict__ x, float* __restrict__ y, long long t, const float* __restrict__ curves,
                int ntaps) {
  extern __shared__ float wp_smem[];                // [kWpTile + ntaps - 1] samples, then [ntaps] taps
  float* xs = wp_smem;
  float* taps = wp_smem + kWpTile + ntaps - 1;
  const long long row = blockIdx.y;
  const long long i0 = (long long)blockIdx.x * kWpTile;
  const float* xr = x + row * t;
  const int half = ntaps / 2;
  const int span = kWpTile + ntaps - 1;
  for (int k = threadIdx.x; k < ntaps; k += 256) taps[k] = curves[row * ntaps + k];
  for (int m = threadIdx.x; m < span; m += 256) {
    long long j = i0 + m - half;                    // index into the un-padded row, reflect (no edge repeat)
    if (j < 0) j = -j;
    if (j >= t) j = 2 * (t - 1) - j;
    xs[m] = (j >= 0 && j < t) ? xr[j] : 0.f;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < kWpTile; o += 256) {
    const long long i = i0 + o;
    if (i >= t) break;
    float acc = 0.f;
    for (int k = 0; k < ntaps; ++k) acc = fmaf(taps[k], xs[o + k], acc);
    y[row * t + i] = acc;
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
  if (op != MPCG_AUG_IDENTITY && op != MPCG_AUG_SELECT && !rowp) return MPCG_EINVAL;
  if (op == MPCG_AUG_SELECT && !noise) return MPCG_EINVAL;
  if ((op == MPCG_AUG_SINE_MUL || op == MPCG_AUG_SINE_ADD) && !(fs > 0.f)) return MPCG_EINVAL;
  if (rows == 0 || t == 0) return MPCG_OK;
  if (!x || !y) return MPCG_EINVAL;
  if (rows > 0x7fffffffLL) return MPCG_ERANGE;
  StageArgs a;
  a.op = op; a.fs = fs; a.rowp = rowp; a.noise = noise; a.mask = mask; a.seed = seed; a.stream = stream_id;
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
  if (rows > 65535) return MPCG_ERANGE;
  const size_t smem = (size_t)(kWpTile + 2 * ntaps) * sizeof(float);
  dim3 grid((unsigned)((t + kWpTile - 1) / kWpTile), (unsigned)rows);
  aug_warp_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, y, (long long)t, curves, ntaps);
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
