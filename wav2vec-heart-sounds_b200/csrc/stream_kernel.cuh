// Kernel behind mpcg_preprocess_segment_f32 (fused.cu) -- the whole conditioning chain of one recording channel
//   resample -> Schmidt despike -> low-pass + high-pass -> abs-max normalise -> overlapping windows
// (torchproc.preprocess_pcg / preprocess_ecg + segment, signalproc/torchproc.py:101-129; NumPy twins
// signalproc/preprocess.py:24-37 + segment.py:40-52) in ONE launch.
//
// Row-streaming form.  A persistent CTA takes whole rows (one channel of one recording) from a ticket counter and
// streams each row through shared memory in tiles of kSkTile samples; nothing is exchanged between CTAs, so a row
// may have any length (480 000 samples at 16 kHz, or a different length per recording) and a CTA never waits for
// another one.  Per row:
//   A (despiked kinds only)  tile by tile: resample -> frame maxima; the resampled tile is parked in this CTA's row
//       buffer (global memory the size of one row, reused row after row, so it lives in L2);
//     the Schmidt passes then run in the reference's order on frames fetched on demand into a small shared-memory
//       cache, every flattened span written through to the row buffer;
//   B  tile by tile: resample (or reload the despiked tile) -> low-pass + high-pass as one 4-state chunked
//       linear-recurrence scan (fp64 state, carried from tile to tile in shared memory) -> row statistics; the
//       filtered samples go to their windows in `out` un-normalised;
//   C  the row's windows are rescaled in place (they were just written, so they are read back from L2) with
//       streaming stores.
// A row that fits one tile never leaves shared memory between A and C.
// The filter recipe (pass-1 weights, scan matrices, section coefficients) travels as a __grid_constant__ kernel
// parameter and is read as constant-bank instruction operands: no table loads in the inner loops and no device
// state shared between launches.
#pragma once
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include "resample.cuh"
#include "biquad.cuh"
#include "despike.cuh"

namespace mpcg {

constexpr int kSkThreads = 512;
constexpr int kSkWarps = kSkThreads / 32;
constexpr int kSkL = 36;                          // samples per filter chunk; 4 x odd: 128-bit shared accesses without bank conflicts
constexpr int kSkTile = kSkThreads * kSkL;        // samples per tile (18 432)
constexpr int kSkGuard = 40;                      // floats before and after the tile that edge frames of the resampler spill into
constexpr int kSkStageWords = 4608;               // resampler input staging: teams x buffers x block
constexpr int kSkMaxFrames = 1024;                // despike frames per row (a 30 s recording has 60)
constexpr int kSkSlots = 8;                       // despike frame cache (streamed rows)
constexpr int kSkBmWords = 1536;                  // 32-sample block maxima of the cached frames
constexpr int kSkWorkHeader = 64;                 // floats at the start of the workspace (ticket counter)
static_assert(kSkL % 4 == 0 && ((kSkL / 4) & 1) == 1, "chunk length must be 4 x odd");
static_assert(kSkWarps == 16, "the cross-warp scan is written for 16 warps");

struct SkKind {                                   // per channel kind (PCG / ECG): despike on/off + its filter
  int despike;
  int pad_;
  double c[2][5];                                 // two sections, b0 b1 b2 a1 a2
  double wt[kSkL][4];                             // A^(L-1-j) B
  double mp[10][16];                              // M^(2^d), d = 0..9, M = A^L
};

struct SkParams {
  const float* x;
  float* out;
  float* work_rows;                               // [ctas, work_stride] row buffers (despiked kinds)
  unsigned int* ticket;                           // row counter, zero at launch
  int* edits;
  int* trace;
  long long* dbg;                                 // optional [ctas, 16] cycles per phase (tools/ only)
  const int* row_t_in;                            // optional per-recording lengths before / after resampling (ragged batch)
  const int* row_t;
  const long long* row_out;                       // optional per-recording element offset of its block in `out`
  long long work_stride;
  long long x_stride;                             // elements between rows of x
  long long recordings;
  long long plane;                                // layout 2: elements per channel plane
  int trace_cap;
  int channels;
  int t_in, t;                                    // uniform row lengths (used where the tables are NULL)
  int off;                                        // resampler input offset
  int win_d;                                      // despike frame length
  double threshold;
  int max_iter, median_mode, norm_flags;
  int start, win, hop, n;                         // window geometry (n: windows of a uniform row)
  int layout;                                     // 0: out[rec, ch, k, j]  1: out[rec, k, j, ch]  2: out[ch, rec, k, j]
  unsigned char kind_of_channel[8];
  unsigned char chan_order[8];                    // channels in processing order: despiked kinds first (longest rows first)
  SkKind kinds[2];
};

struct SkShared {
  union {                                         // phases that follow one another share this space:
    float xs[kSkStageWords];                      //   resampler input staging
    struct {                                      //   despike
      float bm[kSkBmWords];
      SpikeSorted sorted;
    } d;
  };
  float tops[kSkMaxFrames];                       // frame maxima of the current row
  double mtab[16][32];                            // M^lane of the current kind, element-major
  double wagg[kSkWarps][4];                       // warp aggregates
  double wcar[kSkWarps][4];                       // state at the start of each warp's first chunk
  double carry[4];                                // state at the start of the next tile
  double rsum[kSkWarps];
  float rlo[kSkWarps], rhi[kSkWarps];
  float fscr[8];
  int slot_frame[kSkSlots];
  int req, victim;
  unsigned int ticket;
  int pad_;
};
static_assert(sizeof(SkShared) % 16 == 0, "the sample tile behind SkShared must stay 16-byte aligned");

__device__ __forceinline__ float4 ld_cg4(const float4* p) {      // re-read of data this CTA wrote: L2, not L1
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_cg(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void sk_team_sync(int team, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(count) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// Resample samples [t0, t0 + n) of the row into sig[0, n).  Teams of PS warps walk the tile in blocks of 32 frames:
// coalesced loads -> registers (prefetched one block ahead) -> skewed staging -> each warp of the team computes
// its phase group.  Edge frames spill into the guard floats around the tile.  Ends with a CTA barrier.
template <int UP, int DOWN, int D, int PS>
__device__ __forceinline__ void sk_resample_tile(float* sig, float* xs_all, const float* __restrict__ xr, int t_in, int off,
                                                 int t0, int n) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if constexpr (UP == DOWN) {
    for (int i = tid; i < n; i += kSkThreads) sig[i] = ld_stream(xr + t0 + i);
  } else {
    using T = RsTeam<UP, DOWN, D, PS>;
    constexpr int NTEAMS = kSkWarps / PS;
    constexpr int NBUF = (NTEAMS * 2 * T::WORDS <= kSkStageWords) ? 2 : 1;
    static_assert(kSkWarps % PS == 0 && (PS == 1 || NTEAMS <= 15), "teams map onto named barriers 1..15");
    static_assert(NTEAMS * NBUF * T::WORDS <= kSkStageWords, "staging buffer too small for this resampler instance");
    static_assert(UP - 1 <= kSkGuard, "guard too small");
    const int team = warp / PS, grp = warp - team * PS, tt = tid - team * T::TEAM;
    float* xs_team = xs_all + team * (NBUF * T::WORDS);
    const int f_lo = t0 / UP, f_hi = (t0 + n - 1) / UP;
    const int nblk = (f_hi - f_lo + T::FB) / T::FB;
    float pre[T::NPRE];
    int blk = team, buf = 0;
    if (blk < nblk) T::fetch(pre, xr, (long long)(f_lo + blk * T::FB) * DOWN + off, t_in, tt);
    for (; blk < nblk; blk += NTEAMS) {
      float* xs = xs_team + buf * T::WORDS;
      T::commit(xs, pre, tt);
      if constexpr (PS == 1) __syncwarp(); else sk_team_sync(team, T::TEAM);
      if (blk + NTEAMS < nblk)                            // next block's loads fly while this one is computed
        T::fetch(pre, xr, (long long)(f_lo + (blk + NTEAMS) * T::FB) * DOWN + off, t_in, tt);
      const int f = f_lo + blk * T::FB + lane;
      if (f <= f_hi) T::template dispatch<0>(grp, xs, lane, sig + (f * UP - t0));
      if constexpr (NBUF == 2) buf ^= 1;
      else if constexpr (PS == 1) __syncwarp();
      else sk_team_sync(team, T::TEAM);
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// One sample through both sections (transposed direct form II); coefficients are constant-bank operands.
template <int K>
__device__ __forceinline__ double sk_step(const SkParams& P, double (&z)[4], double xv) {
  const double y0 = fma(P.kinds[K].c[0][0], xv, z[0]);
  z[0] = fma(-P.kinds[K].c[0][3], y0, fma(P.kinds[K].c[0][1], xv, z[1]));
  z[1] = fma(-P.kinds[K].c[0][4], y0, P.kinds[K].c[0][2] * xv);
  const double y1 = fma(P.kinds[K].c[1][0], y0, z[2]);
  z[2] = fma(-P.kinds[K].c[1][3], y1, fma(P.kinds[K].c[1][1], y0, z[3]));
  z[3] = fma(-P.kinds[K].c[1][4], y1, P.kinds[K].c[1][2] * y0);
  return y1;
}
template <int K, int DD>
__device__ __forceinline__ void sk_mv_acc(const SkParams& P, const double (&v)[4], double (&acc)[4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    double a = acc[r];
#pragma unroll
    for (int c = 0; c < 4; ++c) a = fma(P.kinds[K].mp[DD][r * 4 + c], v[c], a);
    acc[r] = a;
  }
}
__device__ __forceinline__ float sk_fix(float v) {        // torch.nan_to_num
  if (!(fabsf(v) <= FLT_MAX)) v = (v != v) ? 0.f : (v > 0.f ? FLT_MAX : -FLT_MAX);
  return v;
}

// Low-pass + high-pass of the tile in place: sig[0, kSkTile) holds the samples (zero beyond the row), of which the
// first n_valid belong to the row.  The state enters and leaves through sm.carry.  (sum, min, max) of the valid
// outputs are added to the caller's running statistics.  Ends WITHOUT a barrier after the last write.
template <int K>
__device__ __forceinline__ void sk_filter_tile(const SkParams& P, SkShared& sm, float* sig, int n_valid, bool fix_nan,
                                               double& lsum, float& lmin, float& lmax) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* mine = sig + tid * kSkL;
  // ---- pass 1: zero-state end state of my chunk, p = sum_j A^(L-1-j) B x[j]
  double p[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int g = 0; g < kSkL / 4; ++g) {
    const float4 v = *reinterpret_cast<const float4*>(mine + 4 * g);
    const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double xd = (double)xv[u];
#pragma unroll
      for (int s = 0; s < 4; ++s) p[s] = fma(P.kinds[K].wt[4 * g + u][s], xd, p[s]);
    }
  }
  // ---- inclusive scan of s_(k+1) = M s_k + p_k inside the warp
  {
    double u[4];
#define MPCG_SK_LEVEL(DD)                                                        \
    _Pragma("unroll") for (int s = 0; s < 4; ++s) u[s] = __shfl_up_sync(kFull, p[s], 1 << DD); \
    if (lane >= (1 << DD)) sk_mv_acc<K, DD>(P, u, p);
    MPCG_SK_LEVEL(0) MPCG_SK_LEVEL(1) MPCG_SK_LEVEL(2) MPCG_SK_LEVEL(3) MPCG_SK_LEVEL(4)
#undef MPCG_SK_LEVEL
  }
  if (lane == 31) {
#pragma unroll
    for (int s = 0; s < 4; ++s) sm.wagg[warp][s] = p[s];
  }
  __syncthreads();
  // ---- warp 0 chains the 16 warp aggregates: S_w = M^32 S_(w-1) + agg_w, S_(-1) = the state carried into the tile
  if (warp == 0) {
    double v[4], c0[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      v[s] = (lane < kSkWarps) ? sm.wagg[lane][s] : 0.0;
      c0[s] = sm.carry[s];
    }
    if (lane == 0) sk_mv_acc<K, 5>(P, c0, v);
    double u[4];
#define MPCG_SK_LEVEL(DD)                                                        \
    _Pragma("unroll") for (int s = 0; s < 4; ++s) u[s] = __shfl_up_sync(kFull, v[s], 1 << DD); \
    if (lane >= (1 << DD)) sk_mv_acc<K, 5 + DD>(P, u, v);
    MPCG_SK_LEVEL(0) MPCG_SK_LEVEL(1) MPCG_SK_LEVEL(2) MPCG_SK_LEVEL(3)
#undef MPCG_SK_LEVEL
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const double e = __shfl_up_sync(kFull, v[s], 1);
      if (lane < kSkWarps) sm.wcar[lane][s] = lane ? e : c0[s];
      if (lane == kSkWarps - 1) sm.carry[s] = v[s];
    }
  }
  __syncthreads();
  // ---- true start state of my chunk = (exclusive scan inside the warp) + M^lane (state at the warp's first chunk)
  double z[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const double e = __shfl_up_sync(kFull, p[s], 1);
    z[s] = lane ? e : 0.0;
  }
  {
    double wc[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) wc[s] = sm.wcar[warp][s];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      double a = z[r];
#pragma unroll
      for (int c = 0; c < 4; ++c) a = fma(sm.mtab[r * 4 + c][lane], wc[c], a);
      z[r] = a;
    }
  }
  // ---- pass 2: re-run my chunk from its true state; statistics of the valid outputs ride along
  int lim = n_valid - tid * kSkL;
  lim = lim < 0 ? 0 : (lim > kSkL ? kSkL : lim);
  if (lim == kSkL) {
    float tmin = INFINITY, tmax = -INFINITY;
    double tsum = 0.0;
    bool finite = true;
#pragma unroll
    for (int g3 = 0; g3 < kSkL / 12; ++g3) {              // fp32 sub-sums of 12 samples, then fp64
      float sacc = 0.f;
#pragma unroll
      for (int g = 3 * g3; g < 3 * g3 + 3; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(mine + 4 * g);
        float4 o;
        o.x = (float)sk_step<K>(P, z, (double)v.x);
        o.y = (float)sk_step<K>(P, z, (double)v.y);
        o.z = (float)sk_step<K>(P, z, (double)v.z);
        o.w = (float)sk_step<K>(P, z, (double)v.w);
        *reinterpret_cast<float4*>(mine + 4 * g) = o;
        sacc += (o.x + o.y) + (o.z + o.w);
        tmin = fminf(tmin, fminf(fminf(o.x, o.y), fminf(o.z, o.w)));
        tmax = fmaxf(tmax, fmaxf(fmaxf(o.x, o.y), fmaxf(o.z, o.w)));
      }
      finite = finite && (fabsf(sacc) < INFINITY);
      tsum += (double)sacc;
    }
    if (fix_nan && !finite) {                             // nan_to_num: a NaN or infinity anywhere in the chunk lands here
      tsum = 0.0; tmin = INFINITY; tmax = -INFINITY;
      for (int j = 0; j < kSkL; ++j) {
        const float v = sk_fix(mine[j]);
        mine[j] = v;
        tsum += (double)v;
        tmin = fminf(tmin, v);
        tmax = fmaxf(tmax, v);
      }
    }
    lsum += tsum; lmin = fminf(lmin, tmin); lmax = fmaxf(lmax, tmax);
  } else if (lim > 0) {                                   // the chunk holding the end of the row
    for (int j = 0; j < lim; ++j) {
      float v = (float)sk_step<K>(P, z, (double)mine[j]);
      if (fix_nan) v = sk_fix(v);
      mine[j] = v;
      lsum += (double)v;
      lmin = fminf(lmin, v);
      lmax = fmaxf(lmax, v);
    }
  }
}

// M^lane of kind `kind` -> sm.mtab (warp 0; binary powers of the scan matrices)
__device__ __forceinline__ void sk_build_mtab(const SkParams& P, SkShared& sm, int kind) {
  const int lane = threadIdx.x & 31;
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = (i % 5 == 0) ? 1.0 : 0.0;
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    if ((lane >> d) & 1) {
      const double* m = P.kinds[kind].mp[d];
      double nxt[16];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < 4; ++k) a = fma(m[r * 4 + k], acc[k * 4 + c], a);
          nxt[r * 4 + c] = a;
        }
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = nxt[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) sm.mtab[i][lane] = acc[i];
}

// Copy the part of every window that intersects samples [s0, s0 + n) of the row (held in sig[0, n)) to `obase`.
// so_j == 1 takes 64-bit stores; the data is read again by the rescale pass, so plain (L2-resident) stores.
__device__ __forceinline__ void sk_store_windows(const float* sig, float* obase, int s0, int n, int start, int win, int hop,
                                                 int nwin, long long so_k, long long so_j) {
  const int tid = threadIdx.x;
  const int s1 = s0 + n;
  int k_first = s0 - start - win + 1;
  k_first = k_first > 0 ? (k_first + hop - 1) / hop : 0;
  int k_last = s1 - 1 - start;
  k_last = k_last < 0 ? -1 : k_last / hop;
  if (k_last > nwin - 1) k_last = nwin - 1;
  for (int k = k_first; k <= k_last; ++k) {
    const int w0 = start + k * hop;
    const int a = w0 > s0 ? w0 : s0;
    const int w1 = w0 + win;
    const int b = w1 < s1 ? w1 : s1;
    const int len = b - a;
    if (len <= 0) continue;
    const float* sp = sig + (a - s0);
    if (so_j == 1) {
      float* dp = obase + k * so_k + (a - w0);
      int head = (int)((reinterpret_cast<uintptr_t>(dp) >> 2) & 1u);
      if (head > len) head = len;
      if (head && tid == 0) dp[0] = sp[0];
      const int npair = (len - head) >> 1;
      const float* sq = sp + head;
      float2* dq = reinterpret_cast<float2*>(dp + head);
      if ((reinterpret_cast<uintptr_t>(sq) & 7u) == 0) {
        const float2* sq2 = reinterpret_cast<const float2*>(sq);
        int i = tid;
        for (; i + kSkThreads < npair; i += 2 * kSkThreads) {
          const float2 p0 = sq2[i], p1 = sq2[i + kSkThreads];
          dq[i] = p0;
          dq[i + kSkThreads] = p1;
        }
        for (; i < npair; i += kSkThreads) dq[i] = sq2[i];
      } else {
        int i = tid;
        for (; i + kSkThreads < npair; i += 2 * kSkThreads) {
          const float a0 = sq[2 * i], a1 = sq[2 * i + 1], b0 = sq[2 * (i + kSkThreads)], b1 = sq[2 * (i + kSkThreads) + 1];
          dq[i] = make_float2(a0, a1);
          dq[i + kSkThreads] = make_float2(b0, b1);
        }
        for (; i < npair; i += kSkThreads) dq[i] = make_float2(sq[2 * i], sq[2 * i + 1]);
      }
      if (((len - head) & 1) && tid == 32) dp[len - 1] = sp[len - 1];
    } else {
      float* dp = obase + k * so_k + (long long)(a - w0) * so_j;
      for (int i = tid; i < len; i += kSkThreads) dp[(long long)i * so_j] = sp[i];
    }
  }
}

template <int UP, int DOWN, int D, int PS>
__global__ void __launch_bounds__(kSkThreads, 2)
fused_stream_kernel(const __grid_constant__ SkParams P) {
  extern __shared__ __align__(16) unsigned char sk_raw[];
  SkShared& sm = *reinterpret_cast<SkShared*>(sk_raw);
  float* sig = reinterpret_cast<float*>(sk_raw + sizeof(SkShared)) + kSkGuard;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* rowbuf = P.work_rows + (long long)blockIdx.x * P.work_stride;
  const unsigned long long total_rows = (unsigned long long)P.recordings * (unsigned)P.channels;
  const bool fix_nan = (P.norm_flags & MPCG_NORM_NAN_TO_NUM) != 0;
  int cur_kind = -1;
#if defined(MPCG_FZ_PHASE_CLOCKS) && MPCG_FZ_PHASE_CLOCKS
  long long ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long ph_last = clock64();
#define SK_STAMP(k) do { const long long now_ = clock64(); ph_acc[k] += now_ - ph_last; ph_last = now_; } while (0)
#else
#define SK_STAMP(k) do { } while (0)
#endif

  for (;;) {
    __syncthreads();
    if (tid == 0) sm.ticket = atomicAdd(P.ticket, 1u);
    __syncthreads();
    const unsigned tk = sm.ticket;
    if ((unsigned long long)tk >= total_rows) break;
    const int ci = (int)(tk / (unsigned long long)P.recordings);
    const long long rec = (long long)(tk - (unsigned long long)ci * (unsigned long long)P.recordings);
    const int ch = P.chan_order[ci];
    const long long row = rec * P.channels + ch;
    const int kind = P.kind_of_channel[ch];
    const int t_in_r = P.row_t_in ? P.row_t_in[rec] : P.t_in;
    const int t_r = P.row_t ? P.row_t[rec] : P.t;
    // window geometry of this row (mpcg_window_count)
    int nwin;
    {
      int rem = t_r - P.start;
      if (rem < P.win) rem = P.win;
      nwin = (rem - P.win) / P.hop + 1;
      if (!P.row_t) nwin = P.n;
    }
    long long so_j, so_k, so_c, base;
    if (P.layout == 1) {
      so_j = P.channels; so_k = (long long)P.win * P.channels; so_c = 1;
      base = P.row_out ? P.row_out[rec] : rec * (long long)nwin * P.win * P.channels;
    } else if (P.layout == 2) {
      so_j = 1; so_k = P.win; so_c = P.plane;
      base = P.row_out ? P.row_out[rec] : rec * (long long)nwin * P.win;
    } else {
      so_j = 1; so_k = P.win; so_c = (long long)nwin * P.win;
      base = P.row_out ? P.row_out[rec] : rec * (long long)P.channels * nwin * P.win;
    }
    float* obase = P.out + base + ch * so_c;
    const float* xr = P.x + row * P.x_stride;
    if (kind != cur_kind) {                               // (uniform) new kind: its M^lane table
      if (warp == 0) sk_build_mtab(P, sm, kind);
      cur_kind = kind;
    }
    const int win_d = P.win_d;
    const int nframes = (P.kinds[kind].despike && win_d >= 1 && t_r >= win_d) ? t_r / win_d : 0;
    const int ntiles = (t_r + kSkTile - 1) / kSkTile;
    const bool single = ntiles <= 1;
    int passes = 0;
    SK_STAMP(0);

    // ------------------------------------------------------------ A. resample, frame maxima, Schmidt despike
    if (nframes > 0) {
      for (int i = tid; i < nframes; i += kSkThreads) sm.tops[i] = 0.f;
      for (int tile = 0; tile < ntiles; ++tile) {
        const int t0 = tile * kSkTile;
        const int n = min(kSkTile, t_r - t0);
        sk_resample_tile<UP, DOWN, D, PS>(sig, sm.xs, xr, t_in_r, P.off, t0, n);
        if (!single) {                                    // park the tile in the row buffer
          const int nv = (n + 3) >> 2;
          const float4* s4 = reinterpret_cast<const float4*>(sig);
          float4* g4 = reinterpret_cast<float4*>(rowbuf + t0);
          for (int i = tid; i < nv; i += kSkThreads) g4[i] = s4[i];
        }
        {                                                 // maxima of the frame pieces inside my warp's share of the tile
          const int ra = warp * (kSkTile / kSkWarps);
          const int rb = min(ra + kSkTile / kSkWarps, n);
          int i = ra;
          while (i < rb) {
            const int f = (t0 + i) / win_d;
            if (f >= nframes) break;
            const int seg_end = min(rb, (f + 1) * win_d - t0);
            float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
            int j = i + lane;
            for (; j + 96 < seg_end; j += 128) {
              m0 = fmaxf(m0, fabsf(sig[j])); m1 = fmaxf(m1, fabsf(sig[j + 32]));
              m2 = fmaxf(m2, fabsf(sig[j + 64])); m3 = fmaxf(m3, fabsf(sig[j + 96]));
            }
            for (; j < seg_end; j += 32) m0 = fmaxf(m0, fabsf(sig[j]));
            const unsigned m = __reduce_max_sync(kFull, __float_as_uint(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3))));
            if (lane == 0) atomicMax(reinterpret_cast<unsigned*>(&sm.tops[f]), m);
            i = seg_end;
          }
        }
        __syncthreads();
      }
      SK_STAMP(1);
      // ---- the Schmidt passes in the reference's order.  Warp 0 decides and flattens; frames it needs that are
      // not resident are fetched by the whole CTA.  A pass that moves nothing is a fixed point (the reference
      // would repeat it until max_iterations) and ends the row.
      const int nblk = (win_d + 31) >> 5;
      const int slot_stride = (win_d + 7) & ~3;           // room for the 16-byte phase match
      int nslots = kSkTile / slot_stride;
      if (nslots > kSkSlots) nslots = kSkSlots;
      if (nslots * nblk > kSkBmWords) nslots = kSkBmWords / nblk;
      if (tid < kSkSlots) sm.slot_frame[tid] = -1;
      if (tid == 0) sm.victim = 0;
      if (single) {                                       // every frame is resident: block maxima of all of them
        for (int q = warp; q < nframes * nblk; q += kSkWarps) {
          const int f = q / nblk, b = q - f * nblk;
          const int i = b * 32 + lane;
          const unsigned m = __reduce_max_sync(kFull, __float_as_uint(fmaxf(i < win_d ? fabsf(sig[f * win_d + i]) : 0.f, 0.f)));
          if (lane == 0) sm.d.bm[q] = __uint_as_float(m);
        }
      }
      __syncthreads();
      const bool sorted_mode = nframes <= 64;
      if (warp == 0 && sorted_mode) spike_sort_init(sm.d.sorted, sm.tops, nframes);
      for (;;) {
        if (warp == 0) {
          int req = -2;                                   // -2: the row is finished
          while (passes < P.max_iter) {
            const SpikeDecision dec = sorted_mode ? spike_sort_decide(sm.d.sorted, nframes, P.threshold, P.median_mode)
                                                  : spike_decide_warp(sm.tops, nframes, P.threshold, P.median_mode);
            if (!dec.active) break;
            float* fr;
            float* bmf;
            if (single) {
              fr = sig + dec.worst * win_d;
              bmf = sm.d.bm + dec.worst * nblk;
            } else {
              int slot = -1;
              for (int s = 0; s < nslots; ++s)
                if (sm.slot_frame[s] == dec.worst) slot = s;
              if (slot < 0) { req = dec.worst; break; }
              fr = sig + slot * slot_stride + phase_of(rowbuf + (long long)dec.worst * win_d);
              bmf = sm.d.bm + slot * nblk;
            }
            const float old_top = sm.tops[dec.worst];
            int peak, lo, hi;
            bool changed;
            float new_top;
            spike_pass_warp(fr, win_d, bmf, nblk, old_top, peak, lo, hi, changed, new_top);
            if (!single) {                                // write the flattened span through to the row buffer
              float* g = rowbuf + (long long)dec.worst * win_d;
              for (int i = lo + lane; i < hi; i += 32) g[i] = kSpikeFill;
            }
            if (lane == 0) {
              sm.tops[dec.worst] = new_top;
              if (P.trace && passes < P.trace_cap) {
                int* tr = P.trace + ((long long)row * P.trace_cap + passes) * 4;
                tr[0] = dec.worst; tr[1] = peak; tr[2] = lo; tr[3] = hi;
              }
            }
            __syncwarp();
            ++passes;
            if (!changed) break;
            if (sorted_mode) {
              if (new_top <= old_top) spike_sort_update(sm.d.sorted, dec.worst, old_top, new_top);
              else spike_sort_init(sm.d.sorted, sm.tops, nframes);      // the fill raised a tiny frame: sort afresh
            }
          }
          if (lane == 0) sm.req = req;
        }
        __syncthreads();
        const int req = sm.req;
        if (req < 0) break;
        {                                                 // fetch frame `req` into the next cache slot
          const int slot = sm.victim;
          const float* g = rowbuf + (long long)req * win_d;
          float* fr = sig + slot * slot_stride + phase_of(g);
          {
            int head = (int)(((16u - ((uintptr_t)g & 15u)) & 15u) >> 2);
            if (head > win_d) head = win_d;
            if (tid < head) fr[tid] = ld_cg(g + tid);
            const int nvec = (win_d - head) >> 2;
            const float4* gv = reinterpret_cast<const float4*>(g + head);
            float4* sv = reinterpret_cast<float4*>(fr + head);
            for (int i = tid; i < nvec; i += kSkThreads) sv[i] = ld_cg4(gv + i);
            const int done = head + (nvec << 2);
            if (tid < win_d - done) fr[done + tid] = ld_cg(g + done + tid);
          }
          __syncthreads();
          for (int b = warp; b < nblk; b += kSkWarps) {
            const int i = b * 32 + lane;
            const unsigned m = __reduce_max_sync(kFull, __float_as_uint(fmaxf(i < win_d ? fabsf(fr[i]) : 0.f, 0.f)));
            if (lane == 0) sm.d.bm[slot * nblk + b] = __uint_as_float(m);
          }
          if (tid == 0) {
            sm.slot_frame[slot] = req;
            sm.victim = (slot + 1) % nslots;
          }
          __syncthreads();
        }
      }
      if (P.edits && tid == 0) P.edits[row] = passes;     // (warp 0 counted them)
      SK_STAMP(2);
    } else if (P.edits && tid == 0) {
      P.edits[row] = 0;
    }

    // ------------------------------------------------------------ B. band filter + statistics, tile by tile
    double lsum = 0.0;
    float lmin = INFINITY, lmax = -INFINITY;
    if (tid < 4) sm.carry[tid] = 0.0;
    for (int tile = 0; tile < ntiles; ++tile) {
      const int t0 = tile * kSkTile;
      const int n = min(kSkTile, t_r - t0);
      if (nframes > 0) {
        if (!single) {                                    // the despiked tile comes back from the row buffer
          __syncthreads();                                // (the previous tile's window stores have left shared memory)
          const int nv = (n + 3) >> 2;
          const float4* g4 = reinterpret_cast<const float4*>(rowbuf + t0);
          float4* s4 = reinterpret_cast<float4*>(sig);
          for (int i = tid; i < nv; i += kSkThreads) s4[i] = ld_cg4(g4 + i);
          __syncthreads();
        }
      } else {
        __syncthreads();
        sk_resample_tile<UP, DOWN, D, PS>(sig, sm.xs, xr, t_in_r, P.off, t0, n);
      }
      for (int i = n + tid; i < kSkTile + kSkGuard; i += kSkThreads) sig[i] = 0.f;    // chunk grid beyond the row
      __syncthreads();
      SK_STAMP(3);
      if (kind == 0) sk_filter_tile<0>(P, sm, sig, n, fix_nan, lsum, lmin, lmax);
      else sk_filter_tile<1>(P, sm, sig, n, fix_nan, lsum, lmin, lmax);
      __syncthreads();
      SK_STAMP(4);
      if (!single) sk_store_windows(sig, obase, t0, n, P.start, P.win, P.hop, nwin, so_k, so_j);
      SK_STAMP(5);
    }

    // ------------------------------------------------------------ row statistics -> the normalising map
    lsum = warp_sum(lsum);
    lmin = warp_min(lmin);
    lmax = warp_max(lmax);
    if (lane == 0) { sm.rsum[warp] = lsum; sm.rlo[warp] = lmin; sm.rhi[warp] = lmax; }
    __syncthreads();
    if (warp == 0) {
      double tot = (lane < kSkWarps) ? sm.rsum[lane] : 0.0;
      float lo_f = (lane < kSkWarps) ? sm.rlo[lane] : INFINITY, hi_f = (lane < kSkWarps) ? sm.rhi[lane] : -INFINITY;
      tot = warp_sum(tot);
      const double lo_all = (double)warp_min(lo_f), hi_all = (double)warp_max(hi_f);
      const double mean = t_r > 0 ? tot / (double)t_r : 0.0;
      const double peak = fmax(hi_all - mean, mean - lo_all);
      double inv_peak;
      if (P.norm_flags & MPCG_NORM_PEAK_GT0) inv_peak = (peak > 0.0) ? 1.0 / peak : 1.0;
      else inv_peak = 1.0 / fmax(peak, 1e-12);
      if (lane == 0) {                                    // y = s * inv - mean * inv as one FFMA
        sm.fscr[0] = (float)inv_peak;
        sm.fscr[1] = (float)(-mean * inv_peak);
      }
    }
    __syncthreads();
    const float inv_f = sm.fscr[0], shift_f = sm.fscr[1];
    auto scaled = [&](float s) { return fminf(fmaxf(fmaf(s, inv_f, shift_f), -1.f), 1.f); };

    // ------------------------------------------------------------ C. normalise the row's windows
    // samples of the row that the windows hold: all of every window, except a short row's single zero-padded window
    const long long span = (long long)nwin * P.win;
    long long valid = span;
    if (P.start + (long long)(nwin - 1) * P.hop + P.win > t_r) {        // (only when nwin == 1)
      valid = t_r - P.start;
      if (valid < 0) valid = 0;
    }
    if (single) {                                         // the filtered row is still in shared memory
      for (int k = 0; k < nwin; ++k) {
        const int w0 = P.start + k * P.hop;
        const int len = (int)min((long long)P.win, (long long)t_r - w0);
        float* dp = obase + k * so_k;
        for (int i = tid; i < len; i += kSkThreads) st_stream(dp + (long long)i * so_j, scaled(sig[w0 + i]));
      }
    } else if (so_j == 1) {
      float* p = obase;                                   // [nwin * win] contiguous
      int head = (int)(((16u - ((uintptr_t)p & 15u)) & 15u) >> 2);
      if ((long long)head > valid) head = (int)valid;
      if (tid < head) st_stream(p + tid, scaled(ld_cg(p + tid)));
      const long long nvec = (valid - head) >> 2;
      float4* p4 = reinterpret_cast<float4*>(p + head);
      long long i = tid;
      for (; i + kSkThreads < nvec; i += 2 * kSkThreads) {
        float4 a = ld_cg4(p4 + i), b = ld_cg4(p4 + i + kSkThreads);
        a.x = scaled(a.x); a.y = scaled(a.y); a.z = scaled(a.z); a.w = scaled(a.w);
        b.x = scaled(b.x); b.y = scaled(b.y); b.z = scaled(b.z); b.w = scaled(b.w);
        st_stream4(p4 + i, a);
        st_stream4(p4 + i + kSkThreads, b);
      }
      for (; i < nvec; i += kSkThreads) {
        float4 a = ld_cg4(p4 + i);
        a.x = scaled(a.x); a.y = scaled(a.y); a.z = scaled(a.z); a.w = scaled(a.w);
        st_stream4(p4 + i, a);
      }
      const long long done = head + (nvec << 2);
      if (tid < valid - done) st_stream(p + done + tid, scaled(ld_cg(p + done + tid)));
    } else {
      for (long long i = tid; i < valid; i += kSkThreads) {
        float* q = obase + i * so_j;
        *q = scaled(ld_cg(q));
      }
    }
    for (long long i = valid + tid; i < span; i += kSkThreads) obase[i * so_j] = 0.f;   // short row: zero padding
    SK_STAMP(6);
  }
#if defined(MPCG_FZ_PHASE_CLOCKS) && MPCG_FZ_PHASE_CLOCKS
  if (P.dbg && tid == 0)
    for (int k = 0; k < 8; ++k) P.dbg[(long long)blockIdx.x * 16 + k] = ph_acc[k];
#endif
#undef SK_STAMP
}

template <int UP, int DOWN, int D, int PS>
int sk_launch(const SkParams& P, size_t smem, int ctas, cudaStream_t stream) {
  auto kern = fused_stream_kernel<UP, DOWN, D, PS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<ctas, kSkThreads, smem, stream>>>(P);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  return MPCG_OK;
}

}  // namespace mpcg
