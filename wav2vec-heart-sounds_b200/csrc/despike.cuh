// Schmidt spike removal building blocks (reference: signalproc/torchproc.py:69-98 for the tensor
// path with torch.median's lower-middle rule; signalproc/despike.py:16-54 for the NumPy path with the
// mean-of-middles median and the extra "median == 0" stop).
//
// Everything integer here (worst frame, peak index, flattened span) follows the reference's
// first-maximum / strict-sign-flip rules exactly; the parity tests compare those integers bit for bit.
#pragma once
#include "common.cuh"

namespace mpcg {

constexpr float kSpikeFill = 1e-4f;               // the reference writes Python 1e-4 into the tensor
constexpr int kDespikeMaxFrames = 8192;

// Decision for one pass, computed redundantly by every thread of the CTA from the frame maxima.
struct SpikeDecision {
  bool active;       // some frame exceeds threshold * median
  int worst;         // first frame holding the largest maximum
};

// k-th smallest (0-based) of v[0..n) by rank counting; all threads call, all get the answer.
// Scratch: one float.  O(n^2 / THREADS): n is the number of 500 ms frames, 60 for a 30 s recording.
template <int THREADS>
__device__ __forceinline__ float block_kth(const float* v, int n, int k, float* out1) {
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += THREADS) {
    const float vi = v[i];
    int less = 0, leq = 0;
    for (int j = 0; j < n; ++j) {
      const float vj = v[j];
      less += (vj < vi);
      leq += (vj <= vi);
    }
    if (less <= k && k < leq) *out1 = vi;      // equal values may race; they write the same bits
  }
  __syncthreads();
  return *out1;
}

template <int THREADS>
__device__ __forceinline__ SpikeDecision spike_decide(const float* tops, int nframes, double threshold,
                                                      int median_mode, float* fscr, int* iscr) {
  SpikeDecision d;
  // worst frame: first arg-max
  float bv = -1.f;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < nframes; i += THREADS) {
    const float v = tops[i];
    if (v > bv) { bv = v; bi = i; }             // ascending i per thread keeps the first maximum
  }
  block_argmax_first<THREADS>(bv, bi, fscr, iscr);
  d.worst = bi;
  const float lo_mid = block_kth<THREADS>(tops, nframes, (nframes - 1) >> 1, fscr + 32);
  if (median_mode == MPCG_MEDIAN_LOWER) {
    // tensor path: fp32 tensor times Python scalar -> fp32 product, then fp32 compare
    const float cut = __fmul_rn((float)threshold, lo_mid);
    d.active = bv > cut;
  } else {
    const float hi_mid = block_kth<THREADS>(tops, nframes, nframes >> 1, fscr + 33);
    const double med = ((double)lo_mid + (double)hi_mid) * 0.5;
    d.active = (med != 0.0) && ((double)bv > threshold * med);
  }
  return d;
}

// Flatten the spike in one frame held in shared memory.  Returns (through refs) the integer decisions;
// `changed` tells whether any sample actually moved (an unchanged pass is a fixed point: every later
// pass of the reference would repeat it, so the caller may stop).  All threads call.
template <int THREADS>
__device__ __forceinline__ void spike_flatten(float* fr, int win, int& peak, int& lo, int& hi, bool& changed,
                                              float& new_top, float* fscr, int* iscr) {
  const int tid = threadIdx.x;
  float bv = -1.f;
  int bi = 0x7fffffff;
  for (int i = tid; i < win; i += THREADS) {
    const float a = fabsf(fr[i]);
    if (a > bv) { bv = a; bi = i; }
  }
  block_argmax_first<THREADS>(bv, bi, fscr, iscr);
  peak = bi;
  // strict sign flips between i and i+1 (zeros are not flips)
  int last_before = -1, first_after = 0x7fffffff;
  for (int i = tid; i < win - 1; i += THREADS) {
    const float a = fr[i], b = fr[i + 1];
    const bool flip = (a > 0.f && b < 0.f) || (a < 0.f && b > 0.f);
    if (flip) {
      if (i < peak) last_before = max(last_before, i);
      else first_after = min(first_after, i);
    }
  }
  last_before = block_max_i<THREADS>(last_before, iscr);
  first_after = block_min_i<THREADS>(first_after, iscr);
  lo = last_before + 1;                          // -1 + 1 = 0 when there is no earlier flip
  hi = (first_after == 0x7fffffff) ? win - 1 : first_after;
  int moved = 0;
  for (int i = lo + tid; i < hi; i += THREADS) {
    moved |= (fr[i] != kSpikeFill);
    fr[i] = kSpikeFill;
  }
  changed = block_max_i<THREADS>(moved, iscr) != 0;
  float m = 0.f;
  for (int i = tid; i < win; i += THREADS) m = fmaxf(m, fabsf(fr[i]));   // block_max's leading sync orders the writes
  new_top = block_max<THREADS>(m, fscr);
}

}  // namespace mpcg
