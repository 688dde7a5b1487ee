// extern "C" entries: mpcg_beamform_fwd_f32, mpcg_beamform_bwd_f32 -- the time-varying sinc delay-and-sum gather of
// classify/beamformer.py:41-55 (TimeVaryingSincBeamformer._delay_channel + the sum of squares over microphones) and
// its backward pass (the module is learnable: the delays come from a small transformer).
//
//   kern[k] = sinc(tau_k - d) * hamming[k] / sum_k(...),  tau_k = k - K/2,  d = delays[b, m, t]   (per output sample)
//   y[b, m, t] = sum_k xp[b, m, t + k] * kern[k],         xp = x reflect-padded by K/2
//   out[b, t]  = sum_m y[b, m, t]^2
//
// tau_k is an integer, so sin(pi (tau_k - d)) = -(-1)^tau_k sin(pi d): one sine per output sample instead of K.
// Backward: dL/dy = 2 y g; the gradient of x is a GATHER over the (at most two, at the reflected edges) padded
// positions that read x[j], each collecting K taps of its neighbours' kernels; the gradient of d follows from
//   d kern_k / dd = (u_k' S - u_k S') / S^2,  u_k = sinc(tau_k - d) h_k,  u_k' = -sinc'(tau_k - d) h_k.
#include "common.cuh"

namespace mpcg {

constexpr int kBfMaxK = 129;
struct BfWindow {                   // the taper travels as a kernel parameter: no device state shared between calls
  float w[kBfMaxK];
};

struct BfTap {                      // u_k and u_k' for one (sample, tap)
  float u, du;
};
// z = tau - d;  sinc(z) = sin(pi z) / (pi z),  sinc'(z) = (cos(pi z) - sinc(z)) / z;  with integer tau:
// sin(pi z) = -sgn sin(pi d),  cos(pi z) = sgn cos(pi d),  sgn = (-1)^tau
__device__ __forceinline__ BfTap bf_tap(int tau, float d, float sd, float cd, float h) {
  const float z = (float)tau - d;
  const float sgn = (tau & 1) ? -1.f : 1.f;
  BfTap t;
  if (fabsf(z) < 1e-6f) {
    t.u = h;                        // sinc(0) = 1, sinc'(0) = 0
    t.du = 0.f;
  } else {
    const float inv = 1.f / z;
    const float s = -sgn * sd * 0.3183098861837907f * inv;
    t.u = s * h;
    t.du = -((sgn * cd - s) * inv) * h;
  }
  return t;
}
__device__ __forceinline__ long long bf_reflect(long long p, long long t) {      // padded position -> source index
  if (p < 0) p = -p;
  if (p >= t) p = 2 * (t - 1) - p;
  return p;
}

// One thread per (b, t): loops the microphones.  aux (optional, [B, M, T, 4]) keeps (coef = 2 y g' / S with g' = 1,
// i.e. 2 y / S;  dy/dd;  sin(pi d);  unused) for the backward pass.
__global__ void __launch_bounds__(256)
beamform_fwd_kernel(const float* __restrict__ x, const float* __restrict__ delays, float* __restrict__ out,
                    float* __restrict__ aux, long long batch, int mics, long long t, int K,
                    const __grid_constant__ BfWindow win) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * t) return;
  const long long b = i / t, tt = i - b * t;
  const int half = K / 2;
  float acc = 0.f;
  for (int m = 0; m < mics; ++m) {
    const float* xr = x + (b * mics + m) * t;
    const float d = delays[(b * mics + m) * t + tt];
    float sd, cd;
    sincospif(d, &sd, &cd);
    float S = 0.f, dS = 0.f, xu = 0.f, xdu = 0.f;
    for (int k = 0; k < K; ++k) {
      const BfTap tp = bf_tap(k - half, d, sd, cd, win.w[k]);
      const float xv = xr[bf_reflect(tt + k - half, t)];
      S += tp.u; dS += tp.du;
      xu = fmaf(xv, tp.u, xu);
      xdu = fmaf(xv, tp.du, xdu);
    }
    const float y = xu / S;
    acc = fmaf(y, y, acc);
    if (aux) {
      float4 a;
      a.x = 2.f * y / S;                                  // dL/dy / (g S): the weight each tap u_k carries back to x
      a.y = 2.f * y * (xdu / S - y * dS / S);             // d(y^2)/dd
      a.z = sd;
      a.w = cd;
      reinterpret_cast<float4*>(aux)[(b * mics + m) * t + tt] = a;
    }
  }
  out[i] = acc;
}

// One thread per (b, m, j): gradient of x[b, m, j] and of delays[b, m, j].
__global__ void __launch_bounds__(256)
beamform_bwd_kernel(const float* __restrict__ delays, const float* __restrict__ aux, const float* __restrict__ gout,
                    float* __restrict__ gx, float* __restrict__ gd, long long batch, int mics, long long t, int K,
                    const __grid_constant__ BfWindow win) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * mics * t) return;
  const long long bm = i / t, j = i - bm * t;
  const long long b = bm / mics;
  const int half = K / 2;
  const float4* ar = reinterpret_cast<const float4*>(aux) + bm * t;
  const float* dr = delays + bm * t;
  const float* gr = gout + b * t;
  if (gd) gd[i] = gr[j] * ar[j].y;
  if (!gx) return;
  // padded positions whose source is j: p = j, and the mirror images 2*0 - j (if within the left pad) and
  // 2 (t - 1) - j (if within the right pad); output sample s reads position p with tap k = p - s + half
  float acc = 0.f;
  long long ps[3];
  int np = 0;
  ps[np++] = j;
  if (j >= 1 && j <= half) ps[np++] = -j;
  if (j <= t - 2 && j >= t - 1 - half) ps[np++] = 2 * (t - 1) - j;
  for (int q = 0; q < np; ++q) {
    const long long p = ps[q];
    for (int k = 0; k < K; ++k) {
      const long long s = p - (k - half);
      if (s < 0 || s >= t) continue;
      const float4 a = ar[s];
      const BfTap tp = bf_tap(k - half, dr[s], a.z, a.w, win.w[k]);
      acc = fmaf(gr[s] * a.x, tp.u, acc);
    }
  }
  gx[i] = acc;
}

}  // namespace mpcg


extern "C" int mpcg_beamform_fwd_f32(const float* x, const float* delays, float* out, float* aux, int64_t batch, int mics,
                                     int64_t t, const float* window, int kernel_size, void* stream_) {
  using namespace mpcg;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (batch < 0 || mics < 1 || t < 0 || kernel_size < 1 || (kernel_size & 1) == 0) return MPCG_EINVAL;
  if (kernel_size > kBfMaxK) return MPCG_ERANGE;
  if (batch == 0 || t == 0) return MPCG_OK;
  if (!x || !delays || !out || !window) return MPCG_EINVAL;
  if (t <= kernel_size / 2) return MPCG_EINVAL;                       // reflect padding needs pad < length
  BfWindow win;
  for (int k = 0; k < kBfMaxK; ++k) win.w[k] = k < kernel_size ? window[k] : 0.f;
  const long long n = batch * t;
  beamform_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(x, delays, out, aux, batch, mics, t, kernel_size, win);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}

extern "C" int mpcg_beamform_bwd_f32(const float* delays, const float* aux, const float* grad_out, float* grad_x,
                                     float* grad_delays, int64_t batch, int mics, int64_t t, const float* window,
                                     int kernel_size, void* stream_) {
  using namespace mpcg;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (batch < 0 || mics < 1 || t < 0 || kernel_size < 1 || (kernel_size & 1) == 0) return MPCG_EINVAL;
  if (kernel_size > kBfMaxK) return MPCG_ERANGE;
  if (batch == 0 || t == 0) return MPCG_OK;
  if (!delays || !aux || !grad_out || !window) return MPCG_EINVAL;
  BfWindow win;
  for (int k = 0; k < kBfMaxK; ++k) win.w[k] = k < kernel_size ? window[k] : 0.f;
  const long long n = batch * mics * t;
  beamform_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(delays, aux, grad_out, grad_x, grad_delays, batch, mics,
                                                                         t, kernel_size, win);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}
