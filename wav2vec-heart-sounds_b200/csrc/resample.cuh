// Polyphase resampler tile primitives shared by the stand-alone resample kernel and the fused
// preprocess kernel.  Dense frame form (see resample.cu):
//   y[i*UP + p] = sum_{d<D} x[i*DOWN + off + d] * G[p][d]
// Taps are compile-time constants (resample_taps_gen.cuh), so the unrolled loops are FFMA-immediate.
#pragma once
#include "common.cuh"
#include "resample_taps_gen.cuh"

namespace mpcg {

__host__ __device__ constexpr int rs_skew(int idx, int stride) { return (stride & 1) ? idx : idx + idx / stride; }

// THREADS threads work on NF = (THREADS/PS)*FR frames at a time: thread (grp, fi) computes the phases of
// group grp (a contiguous 1/PS share of the UP phases; warp-uniform) for frames fi*FR .. fi*FR+FR-1.
template <int UP, int DOWN, int D, int FR, int PS, int THREADS>
struct RsTile {
  static_assert(THREADS % PS == 0 && (THREADS / PS) % 32 == 0, "phase groups must be whole warps");
  static constexpr int NFT = THREADS / PS;
  static constexpr int NF = NFT * FR;
  static constexpr int NIN = (NF - 1) * DOWN + D;
  static constexpr int NOUT = NF * UP;
  static constexpr int SIN = FR * DOWN;
  static constexpr int PER_THREAD_IN = (FR - 1) * DOWN + D;
  static constexpr int IN_WORDS = rs_skew(NIN - 1, SIN) + 1;
  using Taps = BakedTaps<UP, DOWN, D>;

  // xs[skew(m)] = x_row[in0 + m], zero outside [0, t_in)
  __device__ static __forceinline__ void stage(float* xs, const float* __restrict__ x_row, long long in0,
                                               long long t_in) {
    for (int m = threadIdx.x; m < NIN; m += THREADS) {
      const long long src = in0 + m;
      xs[rs_skew(m, SIN)] = (src >= 0 && src < t_in) ? ld_stream(x_row + src) : 0.f;
    }
  }

  // PB outputs are accumulated side by side so four independent FFMA chains are always in flight.
  static constexpr int PB = 4;
  template <int P0, int P1, class Sink>
  __device__ static __forceinline__ void phases(const float (&in)[PER_THREAD_IN], int frame0, Sink& sink) {
#pragma unroll
    for (int fr = 0; fr < FR; ++fr) {
#pragma unroll
      for (int p = P0; p < P1; p += PB) {
        float acc[PB];
#pragma unroll
        for (int k = 0; k < PB; ++k) acc[k] = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
#pragma unroll
          for (int k = 0; k < PB; ++k)
            if (p + k < P1) acc[k] = fmaf(in[fr * DOWN + d], Taps::tap(p + k < UP ? p + k : 0, d), acc[k]);
        }
#pragma unroll
        for (int k = 0; k < PB; ++k)
          if (p + k < P1) sink(frame0 + fr, p + k, acc[k]);
      }
    }
  }

  // Register-prefetched staging: issue the global loads for a tile early (fetch), park them in registers
  // while the previous tile is being computed, then drop them into shared memory (commit).
  static constexpr int NPRE = (NIN + THREADS - 1) / THREADS;
  __device__ static __forceinline__ void fetch(float (&pre)[NPRE], const float* __restrict__ x_row, long long in0,
                                               long long t_in) {
#pragma unroll
    for (int k = 0; k < NPRE; ++k) {
      const int m = threadIdx.x + k * THREADS;
      const long long src = in0 + m;
      pre[k] = (m < NIN && src >= 0 && src < t_in) ? ld_stream(x_row + src) : 0.f;
    }
  }
  __device__ static __forceinline__ void commit(float* xs, const float (&pre)[NPRE]) {
#pragma unroll
    for (int k = 0; k < NPRE; ++k) {
      const int m = threadIdx.x + k * THREADS;
      if (m < NIN) xs[rs_skew(m, SIN)] = pre[k];
    }
  }

  template <int G, class Sink>
  __device__ static __forceinline__ void dispatch(int grp, const float (&in)[PER_THREAD_IN], int frame0, Sink& sink) {
    if constexpr (G < PS) {
      if (grp == G) phases<(G * UP) / PS, ((G + 1) * UP) / PS>(in, frame0, sink);
      else dispatch<G + 1>(grp, in, frame0, sink);
    }
  }

  // sink(frame_in_tile, phase, value) is called once per output of this thread.
  template <class Sink>
  __device__ static __forceinline__ void compute(const float* xs, Sink& sink) {
    const int grp = threadIdx.x / NFT;
    const int fi = threadIdx.x - grp * NFT;
    float in[PER_THREAD_IN];
    const int base = fi * SIN;
#pragma unroll
    for (int d = 0; d < PER_THREAD_IN; ++d) in[d] = xs[rs_skew(base + d, SIN)];
    dispatch<0>(grp, in, fi * FR, sink);
  }
};

// Does the host-supplied tap matrix equal the baked instance bit for bit?
template <int UP, int DOWN, int D>
static inline bool rs_taps_match(const float* taps, long long off) {
  if (off != BakedTaps<UP, DOWN, D>::kOffset) return false;
  for (int p = 0; p < UP; ++p)
    for (int d = 0; d < D; ++d) {
      const float a = taps[p * D + d], b = BakedTaps<UP, DOWN, D>::tap(p, d);
      if (!(a == b)) return false;
    }
  return true;
}

}  // namespace mpcg
