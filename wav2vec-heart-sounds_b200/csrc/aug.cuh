// Device helpers shared by the per-stage augmentation kernels (aug.cu) and the fused chain kernel (aug_chain.cu):
// counter-based Philox normals, row statistics, the stage transforms of augment/torchaug.py:39-66.
#pragma once
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
  int norm_masked_only;      // 1: only the rows whose mask is on are re-normalised (the NumPy primitives' rule)
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

}  // namespace mpcg
