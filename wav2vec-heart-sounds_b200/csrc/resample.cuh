// Polyphase resampler tile primitives shared by the stand-alone resample kernel and the fused
// preprocess kernel.  Dense frame form (see resample.cu):
//   y[i*UP + p] = sum_{d<D} x[i*DOWN + off + d] * G[p][d]
// Taps are compile-time constants (resample_taps_gen.cuh), so the unrolled loops are FFMA-immediate.
#pragma once
#include <type_traits>
#include "common.cuh"
#include "resample_taps_gen.cuh"

namespace mpcg {

// Compile-time counted loop: f(std::integral_constant<int, I>) for I in [I0, N).  Keeps tap indices constant
// expressions, so every tap is an instruction immediate and zero taps vanish at compile time.
template <int I, int N, class F>
__device__ __forceinline__ void rs_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    rs_for<I + 1, N>(f);
  }
}

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
    rs_for<0, FR>([&](auto frc) {
      constexpr int fr = decltype(frc)::value;
      rs_for<0, (P1 - P0 + PB - 1) / PB>([&](auto gc) {
        constexpr int p = P0 + decltype(gc)::value * PB;
        float acc[PB];
#pragma unroll
        for (int k = 0; k < PB; ++k) acc[k] = 0.f;
        rs_for<0, D>([&](auto dc) {
          constexpr int d = decltype(dc)::value;
          rs_for<0, PB>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            if constexpr (p + k < P1) {
              constexpr float t = Taps::g[p + k][d];
              if constexpr (t != 0.f) acc[k] = fmaf(in[fr * DOWN + d], t, acc[k]);
            }
          });
        });
        rs_for<0, PB>([&](auto kc) {
          constexpr int k = decltype(kc)::value;
          if constexpr (p + k < P1) sink(frame0 + fr, p + k, acc[k]);
        });
      });
    });
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

// ---------------------------------------------------------------------------------------------------------
// Team form used by the fused kernel: a TEAM of PS warps shares one staged block of 32 frames (coalesced global
// loads -> skewed shared memory); warp g of the team computes phase group g (a contiguous 1/PS share of the UP
// phases) for the 32 frames, one frame per lane.  Taps are instruction immediates and STRUCTURALLY ZERO TAPS ARE
// SKIPPED: the Hann-windowed sinc (and the Kaiser polyphase bank) reaches only 12-13 (20-21) of the D columns of
// the dense frame form per phase, the rest are +-0 in float32 (x * 0 contributes nothing for finite x).
// Only the input columns a phase group really touches are read from shared memory.
template <int UP, int DOWN, int D, int PS>
struct RsTeam {
  using Taps = BakedTaps<UP, DOWN, D>;
  static constexpr int FB = 32;                         // frames per block (one per lane)
  static constexpr int NIN = (FB - 1) * DOWN + D;       // staged inputs per block
  static constexpr int WORDS = rs_skew(NIN - 1, DOWN) + 1;
  static constexpr int TEAM = PS * 32;
  static constexpr int NPRE = (NIN + TEAM - 1) / TEAM;
  static constexpr int PB = 4;                          // outputs accumulated side by side

  __host__ __device__ static constexpr int p_begin(int g) { return (g * UP) / PS; }
  __host__ __device__ static constexpr int d_lo(int p0, int p1) {
    int m = D - 1;
    for (int p = p0; p < p1; ++p) m = Taps::first[p] < m ? Taps::first[p] : m;
    return m;
  }
  __host__ __device__ static constexpr int d_hi(int p0, int p1) {
    int m = 0;
    for (int p = p0; p < p1; ++p) m = Taps::last[p] > m ? Taps::last[p] : m;
    return m;
  }

  // inputs [in0, in0 + NIN) of the row -> registers (zero outside the row); tt = thread index inside the team
  __device__ static __forceinline__ void fetch(float (&pre)[NPRE], const float* __restrict__ x_row, long long in0,
                                               long long t_in, int tt) {
    if (in0 >= 0 && in0 + NIN <= t_in) {                // block-uniform: interior block, no bounds checks
#pragma unroll
      for (int k = 0; k < NPRE; ++k) {
        const int m = tt + k * TEAM;
        pre[k] = (m < NIN) ? ld_stream(x_row + in0 + m) : 0.f;
      }
    } else {
#pragma unroll
      for (int k = 0; k < NPRE; ++k) {
        const int m = tt + k * TEAM;
        const long long src = in0 + m;
        pre[k] = (m < NIN && src >= 0 && src < t_in) ? ld_stream(x_row + src) : 0.f;
      }
    }
  }
  __device__ static __forceinline__ void commit(float* xs, const float (&pre)[NPRE], int tt) {
#pragma unroll
    for (int k = 0; k < NPRE; ++k) {
      const int m = tt + k * TEAM;
      if (m < NIN) xs[rs_skew(m, DOWN)] = pre[k];
    }
  }

  // phases [P0, P1) of the frame staged at xs[skew(lane*DOWN + d)]; dst[p] receives phase p.
  template <int P0, int P1>
  __device__ static __forceinline__ void phases(const float* xs, int lane, float* dst) {
    constexpr int DLO = d_lo(P0, P1), DHI = d_hi(P0, P1);
    float in[DHI - DLO + 1];
    const int base = (DOWN & 1) ? lane * DOWN : lane * (DOWN + 1);   // == rs_skew(lane * DOWN, DOWN)
    rs_for<DLO, DHI + 1>([&](auto dc) {
      constexpr int d = decltype(dc)::value;
      in[d - DLO] = xs[base + rs_skew(d, DOWN)];
    });
    rs_for<0, (P1 - P0 + PB - 1) / PB>([&](auto gc) {
      constexpr int p = P0 + decltype(gc)::value * PB;
      float acc[PB];
#pragma unroll
      for (int k = 0; k < PB; ++k) acc[k] = 0.f;
      rs_for<DLO, DHI + 1>([&](auto dc) {
        constexpr int d = decltype(dc)::value;
        rs_for<0, PB>([&](auto kc) {
          constexpr int k = decltype(kc)::value;
          if constexpr (p + k < P1) {
            constexpr float t = Taps::g[p + k][d];
            if constexpr (t != 0.f) acc[k] = fmaf(in[d - DLO], t, acc[k]);
          }
        });
      });
      rs_for<0, PB>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        if constexpr (p + k < P1) dst[p + k] = acc[k];
      });
    });
  }
  template <int G>
  __device__ static __forceinline__ void dispatch(int grp, const float* xs, int lane, float* dst) {
    if constexpr (G < PS) {
      if (grp == G) phases<p_begin(G), p_begin(G + 1)>(xs, lane, dst);
      else dispatch<G + 1>(grp, xs, lane, dst);
    }
  }
};

// ---------------------------------------------------------------------------------------------------------
// Vector form of the team resampler (fused kernel, rows whose first sample is 16-byte aligned): the staged block is
// moved and read in 16-byte chunks.  A lane owns FR consecutive frames with FR * DOWN a multiple of 4, so every
// lane's first input sits at the same offset SH inside its chunk and all register indices are compile-time:
//   global -> registers (128-bit loads, prefetched one block ahead) -> shared (128-bit stores, one pad chunk per
//   eight so that the lanes' 128-bit reads, LS chunks apart, are free of bank conflicts) -> 128-bit reads of only
//   the chunks a phase group touches -> FFMA with immediate taps -> the UP outputs of each frame.
template <int UP, int DOWN, int D, int PS>
struct RsVec {
  using Taps = BakedTaps<UP, DOWN, D>;
  static constexpr int FR = (DOWN % 4 == 0) ? 1 : ((DOWN % 2 == 0) ? 2 : 4);   // frames per lane
  static constexpr int FB = 32 * FR;                    // frames per block
  static constexpr int LS = FR * DOWN / 4;              // chunks between neighbouring lanes
  static constexpr int SH = (int)(((Taps::kOffset % 4) + 4) % 4);
  static constexpr int NCH = ((FR - 1) * DOWN + D + SH + 3) / 4;       // chunks one lane may read
  static constexpr int NCH_BLK = 31 * LS + NCH;         // chunks per staged block
  __host__ __device__ static constexpr int skew(int c) { return (LS & 1) ? c : c + (c >> 3); }
  static constexpr int WORDS = 4 * (skew(NCH_BLK - 1) + 1);
  static constexpr int TEAM = PS * 32;
  static constexpr int NPRE = (NCH_BLK + TEAM - 1) / TEAM;
  static constexpr int PB = 4;                          // outputs accumulated side by side
  static constexpr int ALIGN_F = (DOWN % 4 == 0) ? 1 : (4 / (DOWN % 2 == 0 ? 2 : 1));   // frame alignment of a block start

  __host__ __device__ static constexpr int p_begin(int g) { return (g * UP) / PS; }
  __host__ __device__ static constexpr int d_lo(int p0, int p1) {
    int m = D - 1;
    for (int p = p0; p < p1; ++p) m = Taps::first[p] < m ? Taps::first[p] : m;
    return m;
  }
  __host__ __device__ static constexpr int d_hi(int p0, int p1) {
    int m = 0;
    for (int p = p0; p < p1; ++p) m = Taps::last[p] > m ? Taps::last[p] : m;
    return m;
  }

  // chunks of the block whose first frame is f0 (f0 * DOWN a multiple of 4) -> registers; zero outside the row.
  // 32-bit index arithmetic (rows are shorter than 2^30 samples); a block that lies inside the row takes no
  // per-chunk range checks.
  __device__ static __forceinline__ void fetch(float4 (&pre)[NPRE], const float* __restrict__ x_row, int f0, int t_in, int tt) {
    const int a0 = f0 * DOWN + (int)Taps::kOffset - SH;       // a multiple of 4
    const float* p = x_row + a0 + 4 * tt;
    if (a0 >= 0 && a0 + 4 * NCH_BLK <= t_in) {                // block-uniform: interior block
#pragma unroll
      for (int k = 0; k < NPRE; ++k) {
        if ((k + 1) * TEAM <= NCH_BLK || tt + k * TEAM < NCH_BLK)
          pre[k] = ld_stream4(reinterpret_cast<const float4*>(p + 4 * k * TEAM));
      }
    } else {
#pragma unroll
      for (int k = 0; k < NPRE; ++k) {
        const int c = tt + k * TEAM;
        const int idx = a0 + 4 * c;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < NCH_BLK && idx >= 0 && idx < t_in) {
          if (idx + 4 <= t_in) {
            v = ld_stream4(reinterpret_cast<const float4*>(x_row + idx));
          } else {                                            // the chunk holding the end of a row whose length is not a multiple of 4
            v.x = ld_stream(x_row + idx);
            if (idx + 1 < t_in) v.y = ld_stream(x_row + idx + 1);
            if (idx + 2 < t_in) v.z = ld_stream(x_row + idx + 2);
          }
        }
        pre[k] = v;
      }
    }
  }
  __device__ static __forceinline__ void commit(float* xs, const float4 (&pre)[NPRE], int tt) {
    float4* dst = reinterpret_cast<float4*>(xs);
#pragma unroll
    for (int k = 0; k < NPRE; ++k) {
      const int c = tt + k * TEAM;
      if ((k + 1) * TEAM <= NCH_BLK || c < NCH_BLK) dst[skew(c)] = pre[k];
    }
  }

  // phases [P0, P1) of the lane's FR frames; dst points at phase 0 of the lane's first frame.
  template <int P0, int P1>
  __device__ static __forceinline__ void phases(const float* xs, int lane, float* dst) {
    constexpr int DLO = d_lo(P0, P1), DHI = d_hi(P0, P1);
    constexpr int C0 = (SH + DLO) / 4, C1 = (SH + (FR - 1) * DOWN + DHI) / 4;      // chunks this phase group reads
    float in[4 * (C1 - C0 + 1)];
    const float4* xc = reinterpret_cast<const float4*>(xs);
    rs_for<C0, C1 + 1>([&](auto cc) {
      constexpr int c = decltype(cc)::value;
      const float4 v = xc[skew(lane * LS + c)];
      in[4 * (c - C0)] = v.x; in[4 * (c - C0) + 1] = v.y; in[4 * (c - C0) + 2] = v.z; in[4 * (c - C0) + 3] = v.w;
    });
    rs_for<0, FR>([&](auto frc) {
      constexpr int fr = decltype(frc)::value;
      rs_for<0, (P1 - P0 + PB - 1) / PB>([&](auto gc) {
        constexpr int p = P0 + decltype(gc)::value * PB;
        float acc[PB];
#pragma unroll
        for (int k = 0; k < PB; ++k) acc[k] = 0.f;
        rs_for<DLO, DHI + 1>([&](auto dc) {
          constexpr int d = decltype(dc)::value;
          rs_for<0, PB>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            if constexpr (p + k < P1) {
              constexpr float t = Taps::g[p + k][d];
              if constexpr (t != 0.f) acc[k] = fmaf(in[SH + fr * DOWN + d - 4 * C0], t, acc[k]);
            }
          });
        });
        rs_for<0, PB>([&](auto kc) {
          constexpr int k = decltype(kc)::value;
          if constexpr (p + k < P1) dst[fr * UP + p + k] = acc[k];
        });
      });
    });
  }
  template <int G>
  __device__ static __forceinline__ void dispatch(int grp, const float* xs, int lane, float* dst) {
    if constexpr (G < PS) {
      if (grp == G) phases<p_begin(G), p_begin(G + 1)>(xs, lane, dst);
      else dispatch<G + 1>(grp, xs, lane, dst);
    }
  }
};

// ---------------------------------------------------------------------------------------------------------
// Run form for integer up-sampling (DOWN == 1, UP a multiple of 4): a lane owns a RUN of FR = 5 consecutive frames,
// i.e. 5 * UP consecutive outputs, computed from FR - 1 + D staged inputs held in registers and stored with 128-bit
// stores.  The odd run length keeps the lanes' scalar input reads free of bank conflicts and spreads their 128-bit
// stores over the banks (lane stride 5 * UP * 4 bytes: two-way at worst).  One warp per block, warp-private staging.
template <int UP, int D>
struct RsRun {
  using Taps = BakedTaps<UP, 1, D>;
  static_assert(UP % 4 == 0, "the run form stores whole 16-byte groups of outputs");
  static constexpr int FR = 5;
  static constexpr int FB = 32 * FR;                    // frames per block
  static constexpr int NIN = FB - 1 + D;                // staged inputs per block
  static constexpr int WORDS = (NIN + 3) & ~3;
  static constexpr int NPRE = (NIN + 31) / 32;
  static constexpr int PER_LANE = FR - 1 + D;

  __device__ static __forceinline__ void fetch(float (&pre)[NPRE], const float* __restrict__ x_row, long long f0, int t_in,
                                               int lane) {
    const long long in0 = f0 + Taps::kOffset;
#pragma unroll
    for (int k = 0; k < NPRE; ++k) {
      const int m = lane + 32 * k;
      const long long src = in0 + m;
      pre[k] = (m < NIN && src >= 0 && src < t_in) ? ld_stream(x_row + src) : 0.f;
    }
  }
  __device__ static __forceinline__ void commit(float* xs, const float (&pre)[NPRE], int lane) {
#pragma unroll
    for (int k = 0; k < NPRE; ++k) {
      const int m = lane + 32 * k;
      if (m < NIN) xs[m] = pre[k];
    }
  }
  // dst: 16-byte aligned address of phase 0 of the lane's first frame
  __device__ static __forceinline__ void run(const float* xs, int lane, float* dst) {
    float in[PER_LANE];
    rs_for<0, PER_LANE>([&](auto jc) {
      constexpr int j = decltype(jc)::value;
      in[j] = xs[lane * FR + j];
    });
    rs_for<0, FR>([&](auto frc) {
      constexpr int fr = decltype(frc)::value;
      rs_for<0, UP / 4>([&](auto gc) {
        constexpr int p = decltype(gc)::value * 4;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        rs_for<0, D>([&](auto dc) {
          constexpr int d = decltype(dc)::value;
          rs_for<0, 4>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            constexpr float t = Taps::g[p + k][d];
            if constexpr (t != 0.f) acc[k] = fmaf(in[fr + d], t, acc[k]);
          });
        });
        *reinterpret_cast<float4*>(dst + fr * UP + p) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      });
    });
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
