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

// Same decision computed by ONE WARP from shared memory, no block barrier: every warp of a CTA (and of
// every CTA of a cluster) runs it redundantly and arrives at identical bits.  O(n^2 / 32) per warp.
__device__ __forceinline__ SpikeDecision spike_decide_warp(const float* tops, int nframes, double threshold,
                                                           int median_mode) {
  const int lane = threadIdx.x & 31;
  SpikeDecision d;
  float bv = -1.f;
  int bi = 0x7fffffff;
  for (int i = lane; i < nframes; i += 32) {
    const float v = tops[i];
    if (v > bv) { bv = v; bi = i; }
  }
  warp_argmax_first(bv, bi);
  d.worst = bi;
  const int k_lo = (nframes - 1) >> 1, k_hi = nframes >> 1;
  float c_lo = -1.f, c_hi = -1.f;                    // frame maxima are >= 0, so -1 means "not mine"
  for (int i = lane; i < nframes; i += 32) {
    const float vi = tops[i];
    int less = 0, leq = 0;
    for (int j = 0; j < nframes; ++j) {
      const float vj = tops[j];
      less += (vj < vi);
      leq += (vj <= vi);
    }
    if (less <= k_lo && k_lo < leq) c_lo = vi;
    if (less <= k_hi && k_hi < leq) c_hi = vi;
  }
  const float lo_mid = warp_max(c_lo), hi_mid = warp_max(c_hi);
  if (median_mode == MPCG_MEDIAN_LOWER) {
    d.active = bv > __fmul_rn((float)threshold, lo_mid);
  } else {
    const double med = ((double)lo_mid + (double)hi_mid) * 0.5;
    d.active = (med != 0.0) && ((double)bv > threshold * med);
  }
  return d;
}

// <= 64 frames (a 30 s recording has 60): bitonic sort of 64 (value, ~index) keys held two per lane.
// key = float bits << 32 | (0xffffffff - index): frame maxima are >= 0 so the bit pattern orders like the
// value, and among equal values the smaller index sorts higher, which makes the largest key the FIRST arg-max.
__device__ __forceinline__ SpikeDecision spike_decide_sort64(const float* tops, int nframes, double threshold,
                                                             int median_mode) {
  const int lane = threadIdx.x & 31;
  unsigned long long k0, k1;                       // elements 2*lane and 2*lane+1 of the 64-element array
  {
    const int i0 = 2 * lane, i1 = 2 * lane + 1;
    // padding keys are the smallest possible (0) so real frames occupy the top `nframes` sorted slots
    k0 = (i0 < nframes) ? (((unsigned long long)__float_as_uint(tops[i0]) << 32) | (0xffffffffu - (unsigned)i0)) : 0ull;
    k1 = (i1 < nframes) ? (((unsigned long long)__float_as_uint(tops[i1]) << 32) | (0xffffffffu - (unsigned)i1)) : 0ull;
  }
  // bitonic network over 64 elements, element e = 2*lane + b
#pragma unroll
  for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride == 1) {                           // partner is the other element of the same lane
        const bool up = ((2 * lane) & size) == 0;  // ascending block?
        const unsigned long long lo = k0 < k1 ? k0 : k1, hi = k0 < k1 ? k1 : k0;
        k0 = up ? lo : hi;
        k1 = up ? hi : lo;
      } else {                                     // partner element lives in lane ^ (stride / 2), same b
        const int pl = stride >> 1;
        const unsigned long long o0 = __shfl_xor_sync(kFull, k0, pl), o1 = __shfl_xor_sync(kFull, k1, pl);
        const bool up = ((2 * lane) & size) == 0;
        const bool lower = (lane & pl) == 0;       // am I the lower index of the pair?
        const bool take_min = (up == lower);
        k0 = take_min ? (k0 < o0 ? k0 : o0) : (k0 < o0 ? o0 : k0);
        k1 = take_min ? (k1 < o1 ? k1 : o1) : (k1 < o1 ? o1 : k1);
      }
    }
  }
  // ascending: sorted position s holds element s; real frames sit at positions 64 - nframes .. 63
  auto at = [&](int pos) -> unsigned long long {
    const unsigned long long a = __shfl_sync(kFull, k0, pos >> 1), b = __shfl_sync(kFull, k1, pos >> 1);
    return (pos & 1) ? b : a;
  };
  const int base = 64 - nframes;
  const unsigned long long top = at(63);
  SpikeDecision d;
  d.worst = (int)(0xffffffffu - (unsigned)(top & 0xffffffffu));
  const float bv = __uint_as_float((unsigned)(top >> 32));
  const float lo_mid = __uint_as_float((unsigned)(at(base + ((nframes - 1) >> 1)) >> 32));
  if (median_mode == MPCG_MEDIAN_LOWER) {
    d.active = bv > __fmul_rn((float)threshold, lo_mid);
  } else {
    const float hi_mid = __uint_as_float((unsigned)(at(base + (nframes >> 1)) >> 32));
    const double med = ((double)lo_mid + (double)hi_mid) * 0.5;
    d.active = (med != 0.0) && ((double)bv > threshold * med);
  }
  return d;
}

// ---------------------------------------------------------------------------------------------------------
// Fast path for frames resident in shared memory, driven by ONE warp with no block barrier.
//   * per-frame maxima of 32-sample blocks (bmax) make "first arg-max" and "maximum after the edit" O(blocks);
//   * sign flips are searched outward from the peak in 32-sample ballots;
//   * the frame maxima of the whole recording are kept as a sorted array of 64-bit keys
//     (value bits << 32 | ~index), updated by one insertion per pass instead of a fresh sort.
// Every step reproduces the reference's rules: first maximum, strict +/- flips, [lo, hi) fill, median rank.
struct SpikeSorted {                                  // lives in shared memory, owned by warp 0
  unsigned long long key[64];
};
__device__ __forceinline__ unsigned long long spike_key(float top, int frame) {
  return ((unsigned long long)__float_as_uint(top) << 32) | (0xffffffffu - (unsigned)frame);
}
// Sort the (<= 64) frame maxima into sk (ascending, zero keys pad the bottom).
__device__ __forceinline__ void spike_sort_init(SpikeSorted& sk, const float* tops, int nframes) {
  const int lane = threadIdx.x & 31;
  unsigned long long k0 = (2 * lane < nframes) ? spike_key(tops[2 * lane], 2 * lane) : 0ull;
  unsigned long long k1 = (2 * lane + 1 < nframes) ? spike_key(tops[2 * lane + 1], 2 * lane + 1) : 0ull;
#pragma unroll
  for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const bool up = ((2 * lane) & size) == 0;
      if (stride == 1) {
        const unsigned long long lo = k0 < k1 ? k0 : k1, hi = k0 < k1 ? k1 : k0;
        k0 = up ? lo : hi;
        k1 = up ? hi : lo;
      } else {
        const int pl = stride >> 1;
        const unsigned long long o0 = __shfl_xor_sync(kFull, k0, pl), o1 = __shfl_xor_sync(kFull, k1, pl);
        const bool take_min = (up == ((lane & pl) == 0));
        k0 = take_min ? (k0 < o0 ? k0 : o0) : (k0 < o0 ? o0 : k0);
        k1 = take_min ? (k1 < o1 ? k1 : o1) : (k1 < o1 ? o1 : k1);
      }
    }
  }
  sk.key[2 * lane] = k0;
  sk.key[2 * lane + 1] = k1;
  __syncwarp();
}
// Frame `frame` dropped from old_top to new_top (new <= old): move its key down to its new place.
__device__ __forceinline__ void spike_sort_update(SpikeSorted& sk, int frame, float old_top, float new_top) {
  const int lane = threadIdx.x & 31;
  const unsigned long long ko = spike_key(old_top, frame), kn = spike_key(new_top, frame);
  if (ko == kn) return;
  const unsigned long long a = sk.key[2 * lane], b = sk.key[2 * lane + 1];
  const unsigned m0 = __ballot_sync(kFull, a == ko), m1 = __ballot_sync(kFull, b == ko);
  const int pos_old = m0 ? 2 * (__ffs(m0) - 1) : 2 * (__ffs(m1) - 1) + 1;
  const int below = __popc(__ballot_sync(kFull, a < kn)) + __popc(__ballot_sync(kFull, b < kn));   // insertion point
  unsigned long long na, nb;
  {
    const int i = 2 * lane;
    na = (i == below) ? kn : ((i > below && i <= pos_old) ? sk.key[i - 1] : a);
    const int j = i + 1;
    nb = (j == below) ? kn : ((j > below && j <= pos_old) ? a : b);
  }
  __syncwarp();
  sk.key[2 * lane] = na;
  sk.key[2 * lane + 1] = nb;
  __syncwarp();
}
__device__ __forceinline__ SpikeDecision spike_sort_decide(const SpikeSorted& sk, int nframes, double threshold,
                                                          int median_mode) {
  const int base = 64 - nframes;
  const unsigned long long top = sk.key[63];
  SpikeDecision d;
  d.worst = (int)(0xffffffffu - (unsigned)(top & 0xffffffffu));
  const float bv = __uint_as_float((unsigned)(top >> 32));
  const float lo_mid = __uint_as_float((unsigned)(sk.key[base + ((nframes - 1) >> 1)] >> 32));
  if (median_mode == MPCG_MEDIAN_LOWER) {
    d.active = bv > __fmul_rn((float)threshold, lo_mid);
  } else {
    const float hi_mid = __uint_as_float((unsigned)(sk.key[base + (nframes >> 1)] >> 32));
    const double med = ((double)lo_mid + (double)hi_mid) * 0.5;
    d.active = (med != 0.0) && ((double)bv > threshold * med);
  }
  return d;
}

// One flattening pass on a shared-memory frame by one warp, in two halves.  bm[0..nblk) are the maxima of the frame's
// 32-sample blocks and frame_top == max(bm).
//   spike_find_span: the reference's integer decisions (first arg-max, surrounding strict sign flips -> [lo, hi)).
//   spike_fill_span: writes the fill, refreshes the touched block maxima, returns whether any sample moved and the
//                    frame's new maximum.  `undo` (optional) first receives the hi - lo samples it overwrites.
__device__ __forceinline__ void spike_find_span(const float* fr, int win, const float* bm, int nblk, float frame_top,
                                                int& peak, int& lo, int& hi) {
  const int lane = threadIdx.x & 31;
  // first block holding the maximum, then the first sample inside it
  int blk = -1;
  for (int b0 = 0; b0 < nblk && blk < 0; b0 += 32) {
    const int b = b0 + lane;
    const unsigned m = __ballot_sync(kFull, b < nblk && bm[b] == frame_top);
    if (m) blk = b0 + __ffs(m) - 1;
  }
  {
    const int i = blk * 32 + lane;
    const unsigned m = __ballot_sync(kFull, i < win && fabsf(fr[i]) == frame_top);
    peak = blk * 32 + __ffs(m) - 1;
  }
  auto flip = [&](int i) {                               // strict sign change between samples i and i + 1
    if (i < 0 || i > win - 2) return false;
    const float a = fr[i], b = fr[i + 1];
    return (a > 0.f && b < 0.f) || (a < 0.f && b > 0.f);
  };
  int last_before = -1;
  for (int c = peak - 1; c >= 0; c -= 32) {              // candidates c-31 .. c, highest first
    const unsigned m = __ballot_sync(kFull, flip(c - 31 + lane));
    if (m) { last_before = c - 31 + (31 - __clz(m)); break; }
  }
  int first_after = 0x7fffffff;
  for (int c = peak; c <= win - 2; c += 32) {            // candidates c .. c+31, lowest first
    const unsigned m = __ballot_sync(kFull, flip(c + lane));
    if (m) { first_after = c + __ffs(m) - 1; break; }
  }
  lo = last_before + 1;
  hi = (first_after == 0x7fffffff) ? win - 1 : first_after;
}
__device__ __forceinline__ void spike_fill_span(float* fr, int win, float* bm, int nblk, int lo, int hi, bool& changed,
                                                float& new_top, float* undo = nullptr) {
  const int lane = threadIdx.x & 31;
  int moved = 0;
  for (int i = lo + lane; i < hi; i += 32) {
    const float v = fr[i];
    if (undo) undo[i - lo] = v;
    moved |= (v != kSpikeFill);
    fr[i] = kSpikeFill;
  }
  changed = __any_sync(kFull, moved != 0);
  __syncwarp();
  if (hi > lo) {                                         // refresh the maxima of the blocks the span touched
    for (int b = lo >> 5; b <= ((hi - 1) >> 5); ++b) {
      const int i = b * 32 + lane;
      const unsigned m = __reduce_max_sync(kFull, __float_as_uint(fmaxf(i < win ? fabsf(fr[i]) : 0.f, 0.f)));
      if (lane == 0) bm[b] = __uint_as_float(m);
    }
    __syncwarp();
  }
  float m = 0.f;
  for (int b = lane; b < nblk; b += 32) m = fmaxf(m, bm[b]);
  new_top = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(m)));
}
__device__ __forceinline__ void spike_pass_warp(float* fr, int win, float* bm, int nblk, float frame_top, int& peak,
                                                int& lo, int& hi, bool& changed, float& new_top) {
  spike_find_span(fr, win, bm, nblk, frame_top, peak, lo, hi);
  spike_fill_span(fr, win, bm, nblk, lo, hi, changed, new_top);
}

// Flatten the spike in one frame held in shared memory.  Returns (through refs) the integer decisions;
// `changed` tells whether any sample actually moved (an unchanged pass is a fixed point: every later
// pass of the reference would repeat it, so the caller may stop).  All threads call.
template <int THREADS>
__device__ __forceinline__ void spike_flatten(float* fr, int win, int& peak, int& lo, int& hi, bool& changed,
                                              float& new_top, float* fscr, int* iscr) {
  constexpr int NW = THREADS / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // (1) first arg-max of |x|
  float bv = -1.f;
  int bi = 0x7fffffff;
  for (int i = tid; i < win; i += THREADS) {
    const float a = fabsf(fr[i]);
    if (a > bv) { bv = a; bi = i; }
  }
  block_argmax_first<THREADS>(bv, bi, fscr, iscr);
  peak = bi;
  // (2) strict sign flips between i and i+1 (zeros are not flips): last one before the peak, first at/after it.
  //     One exchange for both: iscr[0..NW) holds the per-warp maxima, iscr[NW..2NW) the per-warp minima.
  int last_before = -1, first_after = 0x7fffffff;
  for (int i = tid; i < win - 1; i += THREADS) {
    const float a = fr[i], b = fr[i + 1];
    const bool flip = (a > 0.f && b < 0.f) || (a < 0.f && b > 0.f);
    if (flip) {
      if (i < peak) last_before = max(last_before, i);
      else first_after = min(first_after, i);
    }
  }
  last_before = warp_max_i(last_before);
  first_after = warp_min_i(first_after);
  __syncthreads();
  if (lane == 0) { iscr[warp] = last_before; iscr[NW + warp] = first_after; }
  __syncthreads();
  last_before = iscr[0]; first_after = iscr[NW];
#pragma unroll
  for (int w = 1; w < NW; ++w) { last_before = max(last_before, iscr[w]); first_after = min(first_after, iscr[NW + w]); }
  lo = last_before + 1;                          // -1 + 1 = 0 when there is no earlier flip
  hi = (first_after == 0x7fffffff) ? win - 1 : first_after;
  // (3) fill the span; the frame's new maximum is max(|fill|, max |x| outside the span); one exchange for
  //     (moved?, outside maximum)
  int moved = 0;
  float m = 0.f;
  for (int i = tid; i < win; i += THREADS) {
    const float v = fr[i];
    if (i >= lo && i < hi) {
      moved |= (v != kSpikeFill);
      fr[i] = kSpikeFill;
      m = fmaxf(m, kSpikeFill);
    } else {
      m = fmaxf(m, fabsf(v));
    }
  }
  moved = warp_max_i(moved);
  m = warp_max(m);
  __syncthreads();                               // every thread is done reading iscr from step (2)
  if (lane == 0) { iscr[warp] = moved; fscr[warp] = m; }
  __syncthreads();
  moved = iscr[0]; m = fscr[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) { moved |= iscr[w]; m = fmaxf(m, fscr[w]); }
  changed = moved != 0;
  new_top = m;
}

}  // namespace mpcg
