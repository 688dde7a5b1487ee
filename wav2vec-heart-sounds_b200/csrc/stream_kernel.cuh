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

#ifndef MPCG_SK_THREADS
#define MPCG_SK_THREADS 512                       // threads per CTA: 512 (one tile group, two CTAs per SM) or 1024 (two groups, one CTA per SM)
#endif
constexpr int kSkThreads = MPCG_SK_THREADS;
constexpr int kSkGT = 512;                        // threads of a tile group
constexpr int kSkGW = kSkGT / 32;                 // warps of a tile group
constexpr int kSkGroups = kSkThreads / kSkGT;     // tile groups per CTA: they take alternate tiles of the row
constexpr int kSkWarps = kSkThreads / 32;
constexpr int kSkCtasPerSm = 1024 / kSkThreads;
constexpr int kSkL = 36;                          // samples per filter chunk; 4 x odd: 128-bit shared accesses without bank conflicts
constexpr int kSkTile = kSkGT * kSkL;             // samples per tile (18 432)
constexpr int kSkGuard = 40;                      // floats before and after the tile that edge frames of the resampler spill into
constexpr int kSkBuf = kSkTile + 2 * kSkGuard + 8;    // floats of one group's tile buffer
constexpr int kSkStageWords = 9 * kSkGT + 128;    // resampler input staging of one group: teams x block
constexpr int kSkMaxFrames = 1024;                // despike frames per row (a 30 s recording has 60)
constexpr int kSkSlots = 16;                      // despike frame cache (streamed rows)
constexpr int kSkBmWords = 1536;                  // 32-sample block maxima of the cached frames
constexpr int kSkWorkHeader = 64;                 // floats at the start of the workspace (ticket counter)
constexpr int kSkFastFrames = 64;                 // frames per row the parallel despike rounds take
constexpr int kSkLogCap = 16;                     // passes per frame and round they log before handing over to the serial order
// named barriers: 0 = CTA, 1..8 = resampler teams, 9.. = tile groups, 11.. = carry hand-over between groups
constexpr int kSkBarGroup = 9;
constexpr int kSkBarCarry = 11;
static_assert(kSkL % 4 == 0 && ((kSkL / 4) & 1) == 1, "chunk length must be 4 x odd");
static_assert(kSkGroups == 1 || kSkGroups == 2, "one or two tile groups");

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
  int serial_despike;                             // 1: always take the serial (reference-order) despike path
  unsigned char kind_of_channel[8];
  unsigned char chan_order[8];                    // channels in processing order: despiked kinds first (longest rows first)
  SkKind kinds[2];
};

struct SkGroupShared {                            // private to one tile group
  union {                                         // phases that follow one another share this space:
    float xs[kSkStageWords];                      //   resampler input staging
    struct {                                      //   despike (group 0's copy is the CTA's)
      float bm[kSkBmWords];
      SpikeSorted sorted;
      // parallel rounds (see the kernel): outcome and pass log of every frame above the round's threshold
      float xtop[kSkFastFrames];                  //   maximum the frame ended the round with
      int xmeta[kSkFastFrames];                   //   passes | stuck << 8 | log full << 9
      int xj[kSkFastFrames];                      //   passes that count after a stuck round
      float seq[kSkFastFrames][kSkLogCap + 1];    //   maximum after j logged passes
      unsigned short span[kSkFastFrames][kSkLogCap][2];   // [lo, hi) of every logged pass
      unsigned char act[kSkFastFrames];           //   slot -> frame
      int nact, verdict, total;
      float cutf, lo_mid, hi_mid;
      double cutd;
      unsigned long long kstar;
    } d;
  };
  double wagg[kSkGW][4];                          // warp aggregates
  double wcar[kSkGW][4];                          // state at the start of each warp's first chunk
  float spanmax[kSkTile / 128];                   // maxima of the tile's aligned 128-sample spans
  unsigned long long bulk_bar;                    // mbarriers of this group's bulk (TMA) loads: whole tiles / first halves,
  unsigned long long bulk_bar2;                   //   second halves
};
struct SkShared {
  SkGroupShared grp[kSkGroups];
  float tops[kSkMaxFrames];                       // frame maxima of the current row
  double mtab[16][32];                            // M^lane of the current kind, element-major
  double carry[4];                                // filter state at the start of the next tile
  double rsum[kSkWarps];
  float rlo[kSkWarps], rhi[kSkWarps];
  float fscr[8];
  int slot_frame[kSkSlots];
  int req, victim;
  unsigned int ticket;
  int pad_;
};
static_assert(sizeof(((SkGroupShared*)0)->d) <= sizeof(float) * kSkStageWords, "despike scratch must fit behind the staging area");
static_assert(sizeof(SkShared) % 16 == 0 && sizeof(SkGroupShared) % 16 == 0, "the sample tiles behind SkShared must stay 16-byte aligned");

// What a thread needs to know about its tile group.
struct SkCtx {
  int g, gt, gw, lane;                            // group, thread / warp inside the group, lane
  float* sig;                                     // the group's tile buffer (past its leading guard)
  SkGroupShared* gs;
  __device__ __forceinline__ void sync() const {  // barrier of the group
    if constexpr (kSkGroups == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(kSkBarGroup + g), "r"(kSkGT) : "memory");
  }
};

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
// ---- bulk asynchronous copies (TMA, cp.async.bulk): a whole tile moves with one instruction issued by one thread
__device__ __forceinline__ uint32_t sk_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sk_bulk_load(void* smem_dst, const void* g, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sk_smem(smem_dst)),
               "l"(g), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void sk_bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void sk_bulk_store(void* g, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(sk_smem(smem_src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void sk_bulk_store_wait() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void sk_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sk_fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Ask L2 for the raw samples that the resampler will read for output samples [t0, t0 + n) of a row (one bulk
// prefetch by one thread, issued a tile ahead: the staged loads then find their lines in L2 instead of waiting for HBM).
template <int UP, int DOWN, int D>
__device__ __forceinline__ void sk_prefetch_inputs(const float* xr, int t_in, int off, int t0, int n) {
  if (n <= 0) return;
  long long lo = (long long)(t0 / UP) * DOWN + off, hi = (long long)((t0 + n - 1) / UP) * DOWN + off + D;
  if (lo < 0) lo = 0;
  if (hi > t_in) hi = t_in;
  const uintptr_t a = reinterpret_cast<uintptr_t>(xr + lo) & ~(uintptr_t)15;
  const uintptr_t b = reinterpret_cast<uintptr_t>(xr + hi) & ~(uintptr_t)15;      // (rounded down: stays inside the row)
  if (b > a) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"((uint32_t)(b - a)) : "memory");
}

__device__ __forceinline__ void sk_team_sync(int bar_id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(count) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// Resample samples [t0, t0 + n) of the row into the group's tile.  Teams of PS warps walk the tile in blocks of
// frames: coalesced loads -> registers (prefetched one block ahead) -> staging in shared memory -> each warp of the
// team computes its phase group.  Edge frames spill into the guard floats around the tile.  Rows that start on a
// 16-byte boundary take the vector form (RsVec: 128-bit loads, stores and staged reads), others the scalar one;
// integer up-sampling takes the run form (RsRun).  Ends with a group barrier.  Not inlined: two call sites.
template <int UP, int DOWN, int D, int PS>
__device__ __noinline__ void sk_resample_tile(const SkCtx c, const float* __restrict__ xr, int t_in, int off, int t0, int n) {
  const int gt = c.gt, lane = c.lane, gw = c.gw;
  // (the pointers arrive through a struct in a non-inlined call: tell the compiler again that they are shared memory)
  float* sig = reinterpret_cast<float*>(__cvta_shared_to_generic(__cvta_generic_to_shared(c.sig)));
  float* xs_all = reinterpret_cast<float*>(__cvta_shared_to_generic(__cvta_generic_to_shared(c.gs->xs)));
  if constexpr (UP == DOWN) {
    for (int i = gt; i < n; i += kSkGT) sig[i] = ld_stream(xr + t0 + i);
  } else {
    constexpr int NTEAMS = kSkGW / PS;
    static_assert(kSkGW % PS == 0 && (PS == 1 || kSkGroups * NTEAMS <= 8), "teams map onto named barriers 1..8");
    static_assert(UP - 1 <= kSkGuard, "guard too small");
    const int team = gw / PS, grp = gw - team * PS;
    const int team_bar = 1 + c.g * NTEAMS + team;
    if constexpr (DOWN == 1 && UP % 4 == 0) {             // integer up-sampling: runs of five frames per lane
      using T = RsRun<UP, D>;
      static_assert(kSkGW * T::WORDS <= kSkStageWords, "staging buffer too small for this resampler instance");
      static_assert(T::FR * UP - 1 <= kSkGuard && kSkTile % UP == 0, "guard too small for a run past the tile end");
      float* xs = xs_all + gw * T::WORDS;
      const int f_lo = t0 / UP, f_hi = (t0 + n - 1) / UP;
      const int nblk = (f_hi - f_lo + T::FB) / T::FB;
      float pre[T::NPRE];
      int blk = gw;
      if (blk < nblk) T::fetch(pre, xr, (long long)(f_lo + blk * T::FB), t_in, lane);
      for (; blk < nblk; blk += kSkGW) {
        T::commit(xs, pre, lane);
        __syncwarp();
        if (blk + kSkGW < nblk) T::fetch(pre, xr, (long long)(f_lo + (blk + kSkGW) * T::FB), t_in, lane);
        const int f = f_lo + blk * T::FB + lane * T::FR;
        if (f <= f_hi) T::run(xs, lane, sig + (f * UP - t0));
        __syncwarp();
      }
    } else if ((reinterpret_cast<uintptr_t>(xr) & 15u) == 0) {
      using T = RsVec<UP, DOWN, D, PS>;
      static_assert(NTEAMS * T::WORDS <= kSkStageWords, "staging buffer too small for this resampler instance");
      static_assert((T::ALIGN_F - 1) * UP + UP - 1 <= kSkGuard, "guard too small for the aligned block start");
      const int tt = gt - team * T::TEAM;
      float* xs = xs_all + team * T::WORDS;
      const int f_lo = (t0 / UP) & ~(T::ALIGN_F - 1), f_hi = (t0 + n - 1) / UP;
      const int nblk = (f_hi - f_lo + T::FB) / T::FB;
      float4 pre[T::NPRE];
      int blk = team;
      if (blk < nblk) T::fetch(pre, xr, f_lo + blk * T::FB, t_in, tt);
      for (; blk < nblk; blk += NTEAMS) {
        T::commit(xs, pre, tt);
        if constexpr (PS == 1) __syncwarp(); else sk_team_sync(team_bar, T::TEAM);
        if (blk + NTEAMS < nblk)                          // next block's loads fly while this one is computed
          T::fetch(pre, xr, f_lo + (blk + NTEAMS) * T::FB, t_in, tt);
        const int f = f_lo + blk * T::FB + lane * T::FR;
        if (f <= f_hi) T::template dispatch<0>(grp, xs, lane, sig + (f * UP - t0));
        if constexpr (PS == 1) __syncwarp(); else sk_team_sync(team_bar, T::TEAM);
      }
    } else {
      using T = RsTeam<UP, DOWN, D, PS>;
      static_assert(NTEAMS * T::WORDS <= kSkStageWords, "staging buffer too small for this resampler instance");
      const int tt = gt - team * T::TEAM;
      float* xs = xs_all + team * T::WORDS;
      const int f_lo = t0 / UP, f_hi = (t0 + n - 1) / UP;
      const int nblk = (f_hi - f_lo + T::FB) / T::FB;
      float pre[T::NPRE];
      int blk = team;
      if (blk < nblk) T::fetch(pre, xr, (long long)(f_lo + blk * T::FB) * DOWN + off, t_in, tt);
      for (; blk < nblk; blk += NTEAMS) {
        T::commit(xs, pre, tt);
        if constexpr (PS == 1) __syncwarp(); else sk_team_sync(team_bar, T::TEAM);
        if (blk + NTEAMS < nblk)
          T::fetch(pre, xr, (long long)(f_lo + (blk + NTEAMS) * T::FB) * DOWN + off, t_in, tt);
        const int f = f_lo + blk * T::FB + lane;
        if (f <= f_hi) T::template dispatch<0>(grp, xs, lane, sig + (f * UP - t0));
        if constexpr (PS == 1) __syncwarp(); else sk_team_sync(team_bar, T::TEAM);
      }
    }
  }
  sk_fence_async_smem();                                  // the tile may leave through a bulk (async-proxy) store
  c.sync();
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
// acc += M^(2^DD) v.  The two sections form a CASCADE: the first section's two states never see the second's, so every
// power of the transition matrix is block lower triangular and rows 0-1 need only columns 0-1 (12 products, not 16).
template <int K, int DD>
__device__ __forceinline__ void sk_mv_acc(const SkParams& P, const double (&v)[4], double (&acc)[4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    double a = acc[r];
#pragma unroll
    for (int c = 0; c < (r < 2 ? 2 : 4); ++c) a = fma(P.kinds[K].mp[DD][r * 4 + c], v[c], a);
    acc[r] = a;
  }
}
__device__ __forceinline__ float sk_fix(float v) {        // torch.nan_to_num
  if (!(fabsf(v) <= FLT_MAX)) v = (v != v) ? 0.f : (v > 0.f ? FLT_MAX : -FLT_MAX);
  return v;
}

// Low-pass + high-pass of the group's tile in place: sig[0, kSkTile) holds the samples (zero beyond the row), of
// which the first n_valid belong to the row.  The filter state enters and leaves through sm.carry; with two tile
// groups the tiles of a row alternate between them, and the group working on tile t receives the state from the
// group that owns tile t - 1 over a named barrier (only the two scanning warps meet there: the state is known as
// soon as that tile's chunk aggregates are scanned, long before its samples are written).  (sum, min, max) of the
// valid outputs are added to the caller's running statistics.  Ends WITHOUT a barrier after the last write.
template <int K>
__device__ __forceinline__ void sk_filter_tile(const SkParams& P, SkShared& sm, const SkCtx c, int tile, int ntiles, int tile_warps,
                                               int n_valid, bool fix_nan, double& lsum, float& lmin, float& lmax) {
  const int gt = c.gt, lane = c.lane, gw = c.gw;
  SkGroupShared& gs = *c.gs;
  float* mine = c.sig + gt * kSkL;
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
    for (int s = 0; s < 4; ++s) gs.wagg[gw][s] = p[s];
  }
  c.sync();
  // ---- the group's warp 0 chains the 16 warp aggregates: S_w = M^32 S_(w-1) + agg_w, S_(-1) = the state carried in
  if (gw == 0) {
    double v[4], c0[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) v[s] = (lane < kSkGW) ? gs.wagg[lane][s] : 0.0;
    if (kSkGroups > 1 && tile > 0)                        // the state after tile - 1, from the other group's scanning warp
      asm volatile("bar.sync %0, 64;" ::"r"(kSkBarCarry + (tile & 1)) : "memory");
#pragma unroll
    for (int s = 0; s < 4; ++s) c0[s] = sm.carry[s];
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
      if (lane < kSkGW) gs.wcar[lane][s] = lane ? e : c0[s];
      if (lane == tile_warps - 1) sm.carry[s] = v[s];       // the state where the next tile begins
    }
    if (kSkGroups > 1 && tile + 1 < ntiles) {             // hand the state to the group that owns tile + 1
      __syncwarp();
      asm volatile("bar.arrive %0, 64;" ::"r"(kSkBarCarry + ((tile + 1) & 1)) : "memory");
    }
  }
  c.sync();
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
    for (int s = 0; s < 4; ++s) wc[s] = gs.wcar[gw][s];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      double a = z[r];
#pragma unroll
      for (int cc = 0; cc < (r < 2 ? 2 : 4); ++cc) a = fma(sm.mtab[r * 4 + cc][lane], wc[cc], a);   // (block lower triangular)
      z[r] = a;
    }
  }
  // ---- pass 2: re-run my chunk from its true state; statistics of the valid outputs ride along
  int lim = n_valid - gt * kSkL;
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

// M^lane of kind `kind` -> sm.mtab (one warp; binary powers of the scan matrices)
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

// Normalise and write the part of every window that intersects samples [s0, s0 + n) of the row (held, filtered, in
// sig[0, n)).  so_j == 1: 128-bit streaming stores to 16-byte aligned destinations; the shared-memory side is read as
// the two aligned 16-byte chunks around each group of four and shifted in registers (the shift is uniform per window
// and tile).  `tid` / `nthr`: the calling thread's index among the nthr threads that share the tile.
// kf: first window that may still intersect the tile (carried from tile to tile by the caller: no division here).
template <int SHIFT, class F>
__device__ __forceinline__ void sk_copy_shifted(const float4* __restrict__ s4, float4* __restrict__ d4, int nvec, int tid,
                                                int nthr, const F& f) {
  for (int i = tid; i < nvec; i += nthr) {
    const float4 a = s4[i];
    float4 o;
    if constexpr (SHIFT == 0) {
      o = a;
    } else {
      const float4 b = s4[i + 1];
      if constexpr (SHIFT == 1) o = make_float4(a.y, a.z, a.w, b.x);
      else if constexpr (SHIFT == 2) o = make_float4(a.z, a.w, b.x, b.y);
      else o = make_float4(a.w, b.x, b.y, b.z);
    }
    st_stream4(d4 + i, make_float4(f(o.x), f(o.y), f(o.z), f(o.w)));
  }
}
template <class F>
__device__ __forceinline__ void sk_store_windows(const float* sig, float* obase, int s0, int n, int start, int win, int hop,
                                                 int nwin, long long so_k, long long so_j, int& kf, int tid, int nthr,
                                                 const F& f) {
  const int s1 = s0 + n;
  while (kf < nwin && start + kf * hop + win <= s0) ++kf;
  for (int k = kf; k < nwin; ++k) {
    const int w0 = start + k * hop;
    if (w0 >= s1) break;
    const int a = w0 > s0 ? w0 : s0;
    const int w1 = w0 + win;
    const int b = w1 < s1 ? w1 : s1;
    const int len = b - a;
    if (len <= 0) continue;
    const float* sp = sig + (a - s0);
    if (so_j == 1) {
      float* dp = obase + k * so_k + (a - w0);
      int head = (int)(((16u - ((uintptr_t)dp & 15u)) & 15u) >> 2);      // scalars up to the first aligned destination
      if (head > len) head = len;
      if (tid < head) st_stream(dp + tid, f(sp[tid]));
      const int nvec = (len - head) >> 2;
      const int src = (a - s0) + head;                                   // tile-relative index of the first body sample
      const float4* s4 = reinterpret_cast<const float4*>(sig) + (src >> 2);
      float4* d4 = reinterpret_cast<float4*>(dp + head);
      switch (src & 3) {
        case 0: sk_copy_shifted<0>(s4, d4, nvec, tid, nthr, f); break;
        case 1: sk_copy_shifted<1>(s4, d4, nvec, tid, nthr, f); break;
        case 2: sk_copy_shifted<2>(s4, d4, nvec, tid, nthr, f); break;
        default: sk_copy_shifted<3>(s4, d4, nvec, tid, nthr, f); break;
      }
      const int done = head + (nvec << 2);
      if (tid >= 32 && tid - 32 < len - done) st_stream(dp + done + tid - 32, f(sp[done + tid - 32]));
    } else {
      float* dp = obase + k * so_k + (long long)(a - w0) * so_j;
      for (int i = tid; i < len; i += nthr) dp[(long long)i * so_j] = f(sp[i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Parallel despike rounds, one frame by one warp.  The frame is fetched from the row buffer into the warp's slot and
// flattened there pass after pass while its maximum exceeds the round's threshold; every pass logs its span and the
// maximum it leaves behind.  Which of the logged passes the reference's serial order really performs is decided
// after all frames of the round are done (see the kernel); those spans are then filled in the row buffer.
struct SkCut {                             // "frame maximum exceeds threshold * median" in the oracle's arithmetic
  int mode;
  float cutf;                              // tensor path: fp32 product, fp32 compare
  double cutd;                             // NumPy path: float64 product and compare
  __device__ __forceinline__ bool exceeds(float top) const {
    return mode == MPCG_MEDIAN_LOWER ? (top > cutf) : ((double)top > cutd);
  }
};
template <class D>
__device__ __forceinline__ void sk_fast_frame(D& d, float* slot_buf, float* bm, const float* g, int win, float top, int slot,
                                              const SkCut& cut) {
  const int lane = threadIdx.x & 31;
  float* fr = slot_buf + phase_of(g);
  {                                                          // fetch: 128-bit loads, four in flight
    int head = (int)(((16u - ((uintptr_t)g & 15u)) & 15u) >> 2);
    if (head > win) head = win;
    if (lane < head) fr[lane] = ld_cg(g + lane);
    const int nvec = (win - head) >> 2;
    const float4* gv = reinterpret_cast<const float4*>(g + head);
    float4* sv = reinterpret_cast<float4*>(fr + head);
    int i = lane;
    for (; i + 96 < nvec; i += 128) {
      const float4 a = ld_cg4(gv + i), b = ld_cg4(gv + i + 32), c = ld_cg4(gv + i + 64), e = ld_cg4(gv + i + 96);
      sv[i] = a; sv[i + 32] = b; sv[i + 64] = c; sv[i + 96] = e;
    }
    for (; i < nvec; i += 32) sv[i] = ld_cg4(gv + i);
    const int done = head + (nvec << 2);
    if (lane < win - done) fr[done + lane] = ld_cg(g + done + lane);
  }
  __syncwarp();
  const int nblk = (win + 31) >> 5;
  for (int b0 = 0; b0 < nblk; b0 += 4) {                     // block maxima, four blocks in flight
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = (b0 + u) * 32 + lane;
      v[u] = i < win ? fabsf(fr[i]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned m = __reduce_max_sync(kFull, __float_as_uint(fmaxf(v[u], 0.f)));   // fmaxf drops NaN
      if (lane == 0 && b0 + u < nblk) bm[b0 + u] = __uint_as_float(m);
    }
  }
  __syncwarp();
  if (lane == 0) d.seq[slot][0] = top;
  int k = 0, stuck = 0, over = 0;
  while (cut.exceeds(top)) {
    if (k == kSkLogCap) { over = 1; break; }
    int peak, lo, hi;
    spike_find_span(fr, win, bm, nblk, top, peak, lo, hi);
    bool changed;
    float new_top;
    spike_fill_span(fr, win, bm, nblk, lo, hi, changed, new_top);
    if (!changed) { stuck = 1; break; }      // a pass that moves nothing: the reference repeats it until max_iterations
    // the fill value RAISED the maximum (a frame quieter than 1e-4): the rounds' "maxima only fall" argument does not
    // hold for this row; nothing of this round is committed and the serial order takes over
    if (new_top > top) { over = 1; break; }
    ++k;
    top = new_top;
    if (lane == 0) {
      d.span[slot][k - 1][0] = (unsigned short)lo;
      d.span[slot][k - 1][1] = (unsigned short)hi;
      d.seq[slot][k] = top;
    }
  }
  if (lane == 0) {
    d.xtop[slot] = top;
    d.xmeta[slot] = k | (stuck << 8) | (over << 9);
  }
}
__device__ __forceinline__ unsigned long long sk_warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const unsigned long long u = __shfl_xor_sync(kFull, v, o);
    v = u > v ? u : v;
  }
  return v;
}

template <int UP, int DOWN, int D, int PS>
__global__ void __launch_bounds__(kSkThreads, kSkCtasPerSm)
fused_stream_kernel(const __grid_constant__ SkParams P) {
  extern __shared__ __align__(16) unsigned char sk_raw[];
  SkShared& sm = *reinterpret_cast<SkShared*>(sk_raw);
  float* pool = reinterpret_cast<float*>(sk_raw + sizeof(SkShared));          // the groups' tile buffers, back to back
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  SkCtx cx;
  cx.g = tid / kSkGT; cx.gt = tid - cx.g * kSkGT; cx.gw = cx.gt >> 5; cx.lane = lane;
  cx.sig = pool + cx.g * kSkBuf + kSkGuard;
  cx.gs = &sm.grp[cx.g];
  auto& dsp = sm.grp[0].d;                                // despike scratch (behind group 0's staging area)
  float* rowbuf = P.work_rows + (long long)blockIdx.x * P.work_stride;
  const unsigned long long total_rows = (unsigned long long)P.recordings * (unsigned)P.channels;
  const bool fix_nan = (P.norm_flags & MPCG_NORM_NAN_TO_NUM) != 0;
  const uint32_t bar = sk_smem(&cx.gs->bulk_bar);
  uint32_t bulk_phase = 0, bulk_phase2 = 0;
  int cur_kind = -1;
  if (cx.gt == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sk_smem(&cx.gs->bulk_bar2)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
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
    // tiles: as many as the row needs, rounded up to a multiple of the group count so that the groups get equal
    // shares; the tile length is a whole number of warp shares (32 chunks), so a tile ends on a chunk boundary
    int ntiles = (t_r + kSkTile - 1) / kSkTile;
    int tl = kSkTile;
    if (ntiles > 1 && kSkGroups > 1) {
      ntiles = (ntiles + kSkGroups - 1) / kSkGroups * kSkGroups;
      const int share = 32 * kSkL;
      tl = ((t_r + ntiles - 1) / ntiles + share - 1) / share * share;
      ntiles = (t_r + tl - 1) / tl;
    }
    const int tile_warps = tl / (32 * kSkL);              // warps whose chunks lie inside a full tile
    const bool single = ntiles <= 1;                      // the whole row fits group 0's tile: it never leaves shared memory
    int passes = 0;
    if (UP != DOWN && cx.gt == 32)                        // the raw samples of my first tile
      sk_prefetch_inputs<UP, DOWN, D>(xr, t_in_r, P.off, cx.g * tl, min(tl, t_r - cx.g * tl));
    SK_STAMP(0);

    // ------------------------------------------------------------ A. resample, frame maxima, Schmidt despike
    if (nframes > 0) {
      for (int i = tid; i < nframes; i += kSkThreads) sm.tops[i] = 0.f;
      __syncthreads();
      bool parked = false;
      for (int tile = cx.g; tile < ntiles; tile += kSkGroups) {
        const int t0 = tile * tl;
        const int n = min(tl, t_r - t0);
        if (parked) {                                     // my previous tile's bulk store has read the buffer
          if (cx.gt == 0) sk_bulk_store_wait();
          cx.sync();
        }
        if (cx.gt == 32 && tile + kSkGroups < ntiles)     // the raw samples of my next tile: on their way to L2
          sk_prefetch_inputs<UP, DOWN, D>(xr, t_in_r, P.off, t0 + kSkGroups * tl, min(tl, t_r - t0 - kSkGroups * tl));
        sk_resample_tile<UP, DOWN, D, PS>(cx, xr, t_in_r, P.off, t0, n);
        if (!single && cx.gt == 0)                        // park the tile in the row buffer: one bulk copy
          sk_bulk_store(rowbuf + t0, cx.sig, (uint32_t)((n + 3) >> 2) * 16u);
        parked = !single;
        // frame maxima in two steps: every aligned 128-sample span of the tile (branch-free sweep), then one warp
        // per frame folds its whole spans and rescans only the ragged ends
        {
          const int ra = cx.gw * (kSkTile / kSkGW);
#pragma unroll
          for (int it = 0; it < kSkTile / kSkGW / 128; ++it) {
            const float4 v = *reinterpret_cast<const float4*>(cx.sig + ra + it * 128 + 4 * lane);
            const float m4 = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
            const unsigned m = __reduce_max_sync(kFull, __float_as_uint(fmaxf(m4, 0.f)));   // fmaxf drops NaN
            if (lane == 0) cx.gs->spanmax[(ra >> 7) + it] = __uint_as_float(m);
          }
        }
        cx.sync();
        {
          const int f_first = t0 / win_d;
          int f_last = (t0 + n - 1) / win_d;
          if (f_last > nframes - 1) f_last = nframes - 1;
          for (int f = f_first + cx.gw; f <= f_last; f += kSkGW) {
            const int lo = max(f * win_d - t0, 0), hi = min((f + 1) * win_d - t0, n);
            const int s_lo = (lo + 127) >> 7, s_hi = hi >> 7;
            float m = 0.f;
            if (s_lo < s_hi) {
              for (int q = s_lo + lane; q < s_hi; q += 32) m = fmaxf(m, cx.gs->spanmax[q]);
              for (int j = lo + lane; j < (s_lo << 7); j += 32) m = fmaxf(m, fabsf(cx.sig[j]));
              for (int j = (s_hi << 7) + lane; j < hi; j += 32) m = fmaxf(m, fabsf(cx.sig[j]));
            } else {
              for (int j = lo + lane; j < hi; j += 32) m = fmaxf(m, fabsf(cx.sig[j]));
            }
            const unsigned mu = __reduce_max_sync(kFull, __float_as_uint(m));
            // a frame that straddles two tiles is updated by both groups: integer max of the non-negative float bits
            if (lane == 0) atomicMax(reinterpret_cast<unsigned*>(&sm.tops[f]), mu);
          }
        }
        cx.sync();
      }
      if (parked && cx.gt == 0) sk_bulk_store_wait();     // the parked row is complete and visible
      __syncthreads();
      SK_STAMP(1);
      const int nblk = (win_d + 31) >> 5;
      const int slot_stride = (win_d + 7) & ~3;           // room for the 16-byte phase match
      int nslots = (kSkGroups * kSkBuf) / slot_stride;
      if (nslots > kSkSlots) nslots = kSkSlots;
      if (nslots * nblk > kSkBmWords) nslots = kSkBmWords / nblk;
      float* sig0 = pool + kSkGuard;                      // (single rows live in group 0's tile)
      bool serial = single || P.trace != nullptr || P.serial_despike || nframes > kSkFastFrames || win_d < 2 ||
                    win_d > 65535;
      // ---------------------------------------------------------- parallel rounds
      // The reference flattens one span per pass, always in the frame holding the largest maximum, until no frame
      // exceeds threshold * median.  A pass only changes its own frame, and the threshold can only fall, so the
      // ORDER of the passes does not matter for the samples: every frame above the current threshold must be
      // flattened until it is not, whatever happens elsewhere.  Rounds: (1) warp 0 derives the threshold and the
      // list of frames above it from the frame maxima; (2) one warp per listed frame fetches it from the row buffer
      // and flattens it in its slot, logging each pass; (3) warp 0 judges the outcome and the passes that the
      // serial order really performs are filled in the row buffer.  All of them are, except when some frame got
      // STUCK (a pass that moves nothing, e.g. a one-sample spike between two sign flips: the reference then
      // repeats that pass until max_iterations): the serial order reaches the stuck state with the largest
      // (maximum, first index) key K* and never leaves it, so exactly the logged passes that started from a key
      // above K* happen, and the row is finished.  Anything the logs cannot settle exactly (pass budget in reach,
      // log full) continues on the serial path from the committed state, which is always a state of the serial order.
      if (!serial) {
        for (int round = 0;; ++round) {
          if (warp == 0) {                                  // threshold + frames above it
            const int nf = nframes;
            const float v0 = lane < nf ? sm.tops[lane] : -1.f, v1 = lane + 32 < nf ? sm.tops[lane + 32] : -1.f;
            int less0 = 0, leq0 = 0, less1 = 0, leq1 = 0;
#pragma unroll 4
            for (int j = 0; j < nf; ++j) {
              const float vj = sm.tops[j];
              less0 += (vj < v0); leq0 += (vj <= v0);
              less1 += (vj < v1); leq1 += (vj <= v1);
            }
            const int k_lo = (nf - 1) >> 1, k_hi = nf >> 1;
            float c_lo = -1.f, c_hi = -1.f;                 // maxima are >= 0, so -1 means "not mine"
            if (lane < nf) {
              if (less0 <= k_lo && k_lo < leq0) c_lo = v0;
              if (less0 <= k_hi && k_hi < leq0) c_hi = v0;
            }
            if (lane + 32 < nf) {
              if (less1 <= k_lo && k_lo < leq1) c_lo = fmaxf(c_lo, v1);
              if (less1 <= k_hi && k_hi < leq1) c_hi = fmaxf(c_hi, v1);
            }
            const float lo_mid = warp_max(c_lo), hi_mid = warp_max(c_hi);
            SkCut cut;
            cut.mode = P.median_mode;
            cut.cutf = __fmul_rn((float)P.threshold, lo_mid);
            const double med = ((double)lo_mid + (double)hi_mid) * 0.5;
            cut.cutd = P.threshold * med;
            const bool open = passes < P.max_iter && (P.median_mode == MPCG_MEDIAN_LOWER || med != 0.0);
            const bool a0 = open && lane < nf && cut.exceeds(v0), a1 = open && lane + 32 < nf && cut.exceeds(v1);
            const unsigned m0 = __ballot_sync(kFull, a0), m1 = __ballot_sync(kFull, a1);
            const unsigned below = (1u << lane) - 1u;
            if (a0) dsp.act[__popc(m0 & below)] = (unsigned char)lane;
            if (a1) dsp.act[__popc(m0) + __popc(m1 & below)] = (unsigned char)(lane + 32);
            if (lane == 0) {
              dsp.nact = __popc(m0) + __popc(m1); dsp.cutf = cut.cutf; dsp.cutd = cut.cutd;
              dsp.lo_mid = lo_mid; dsp.hi_mid = hi_mid;
            }
          }
          __syncthreads();
          const int nact = dsp.nact;
          if (nact == 0) break;
          {                                                 // the listed frames, nslots at a time, one warp each
            SkCut cut;
            cut.mode = P.median_mode; cut.cutf = dsp.cutf; cut.cutd = dsp.cutd;
            if (warp < nslots)
              for (int slot = warp; slot < nact; slot += nslots) {
                const int f = dsp.act[slot];
                sk_fast_frame(dsp, pool + warp * slot_stride, dsp.bm + warp * nblk, rowbuf + (long long)f * win_d, win_d,
                              sm.tops[f], slot, cut);
              }
          }
          __syncthreads();
          if (warp == 0) {                                  // the round's outcome
            const bool h0 = lane < nact, h1 = lane + 32 < nact;
            const int me0 = h0 ? dsp.xmeta[lane] : 0, me1 = h1 ? dsp.xmeta[lane + 32] : 0;
            const bool over = __any_sync(kFull, ((me0 | me1) >> 9) & 1);
            unsigned long long ks = 0ull;
            if ((me0 >> 8) & 1) ks = spike_key(dsp.xtop[lane], dsp.act[lane]);
            if ((me1 >> 8) & 1) {
              const unsigned long long k1 = spike_key(dsp.xtop[lane + 32], dsp.act[lane + 32]);
              ks = k1 > ks ? k1 : ks;
            }
            const unsigned long long kstar = sk_warp_max_u64(ks);
            int total = (me0 & 0xff) + (me1 & 0xff);
#pragma unroll
            for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(kFull, total, o);
            // (with a stuck frame fewer passes may count, plus the stuck pass itself: total + 1 bounds both cases)
            const bool hand_over = over || (long long)passes + total + (kstar != 0ull ? 1 : 0) > (long long)P.max_iter;
            // If no flattened frame fell below the old middle value(s), the median and with it the threshold are what
            // they were: nothing else can exceed it and the row is finished without another round.
            const float floor_v = P.median_mode == MPCG_MEDIAN_LOWER ? dsp.lo_mid : dsp.hi_mid;
            const bool same_median = __all_sync(kFull, (!h0 || dsp.xtop[lane] >= floor_v) &&
                                                           (!h1 || dsp.xtop[lane + 32] >= floor_v));
            if (!hand_over && kstar == 0ull) {
              if (h0) sm.tops[dsp.act[lane]] = dsp.xtop[lane];
              if (h1) sm.tops[dsp.act[lane + 32]] = dsp.xtop[lane + 32];
            }
            if (lane == 0) {
              dsp.verdict = hand_over ? 2 : (kstar != 0ull ? 1 : (same_median ? 3 : 0));
              dsp.total = total;
              dsp.kstar = kstar;
            }
          }
          __syncthreads();
          const int verdict = dsp.verdict;
          const unsigned long long kstar = dsp.kstar;
          if (verdict != 2) {
            for (int slot = warp; slot < nact; slot += kSkWarps) {   // fill the spans of the passes that happen
              const int f = dsp.act[slot];
              const int k = dsp.xmeta[slot] & 0xff;
              int jc = k;
              if (verdict == 1) {                             // only passes that started above K* (all of the stuck frame's)
                if (spike_key(dsp.seq[slot][k], f) != kstar)
                  jc = __popc(__ballot_sync(kFull, lane < k && spike_key(dsp.seq[slot][lane < k ? lane : 0], f) > kstar));
                if (lane == 0) dsp.xj[slot] = jc;
              }
              float* g = rowbuf + (long long)f * win_d;
              for (int j = 0; j < jc; ++j) {
                const int lo = dsp.span[slot][j][0], hi = dsp.span[slot][j][1];
                for (int i = lo + lane; i < hi; i += 32) g[i] = kSpikeFill;
              }
            }
          }
          if (verdict == 2) { serial = true; break; }
          if (verdict == 1) {                                 // exact pass count: what really counted + the pass that moves nothing
            __syncthreads();
            if (warp == 0) {
              int t = (lane < nact ? dsp.xj[lane] : 0) + (lane + 32 < nact ? dsp.xj[lane + 32] : 0);
#pragma unroll
              for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(kFull, t, o);
              if (lane == 0) dsp.total = t + 1;
            }
            __syncthreads();
            passes += dsp.total;
            break;
          }
          passes += dsp.total;
          if (verdict == 3) break;
          __syncthreads();
        }
      }
      // ---------------------------------------------------------- serial path (reference order, pass by pass)
      // Warp 0 decides and flattens; frames it needs that are not resident are fetched by the whole CTA.  A pass
      // that moves nothing is a fixed point (the reference would repeat it until max_iterations) and ends the row.
      if (serial) {
        __syncthreads();
        if (tid < kSkSlots) sm.slot_frame[tid] = -1;
        if (tid == 0) sm.victim = 0;
        if (single) {                                     // every frame is resident: block maxima of all of them
          for (int q = warp; q < nframes * nblk; q += kSkWarps) {
            const int f = q / nblk, b = q - f * nblk;
            const int i = b * 32 + lane;
            const unsigned m = __reduce_max_sync(kFull, __float_as_uint(fmaxf(i < win_d ? fabsf(sig0[f * win_d + i]) : 0.f, 0.f)));
            if (lane == 0) dsp.bm[q] = __uint_as_float(m);
          }
        }
        __syncthreads();
        const bool sorted_mode = nframes <= 64;
        if (warp == 0 && sorted_mode) spike_sort_init(dsp.sorted, sm.tops, nframes);
        for (;;) {
          if (warp == 0) {
            int req = -2;                                 // -2: the row is finished
            while (passes < P.max_iter) {
              const SpikeDecision dec = sorted_mode ? spike_sort_decide(dsp.sorted, nframes, P.threshold, P.median_mode)
                                                    : spike_decide_warp(sm.tops, nframes, P.threshold, P.median_mode);
              if (!dec.active) break;
              float* fr;
              float* bmf;
              if (single) {
                fr = sig0 + dec.worst * win_d;
                bmf = dsp.bm + dec.worst * nblk;
              } else {
                int slot = -1;
                for (int s2 = 0; s2 < nslots; ++s2)
                  if (sm.slot_frame[s2] == dec.worst) slot = s2;
                if (slot < 0) { req = dec.worst; break; }
                fr = pool + slot * slot_stride + phase_of(rowbuf + (long long)dec.worst * win_d);
                bmf = dsp.bm + slot * nblk;
              }
              const float old_top = sm.tops[dec.worst];
              int peak, lo, hi;
              bool changed;
              float new_top;
              spike_pass_warp(fr, win_d, bmf, nblk, old_top, peak, lo, hi, changed, new_top);
              if (!single) {                              // write the flattened span through to the row buffer
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
                if (new_top <= old_top) spike_sort_update(dsp.sorted, dec.worst, old_top, new_top);
                else spike_sort_init(dsp.sorted, sm.tops, nframes);     // the fill raised a tiny frame: sort afresh
              }
            }
            if (lane == 0) sm.req = req;
          }
          __syncthreads();
          const int req = sm.req;
          if (req < 0) break;
          {                                               // fetch frame `req` into the next cache slot
            const int slot = sm.victim;
            const float* g = rowbuf + (long long)req * win_d;
            float* fr = pool + slot * slot_stride + phase_of(g);
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
              if (lane == 0) dsp.bm[slot * nblk + b] = __uint_as_float(m);
            }
            if (tid == 0) {
              sm.slot_frame[slot] = req;
              sm.victim = (slot + 1) % nslots;
            }
            __syncthreads();
          }
        }
      }
      if (P.edits && tid == 0) P.edits[row] = passes;     // (thread 0 counted on either path)
      if (!single) sk_fence_async_global();               // the filled spans are read back by bulk (async-proxy) loads
      SK_STAMP(2);
    } else if (P.edits && tid == 0) {
      P.edits[row] = 0;
    }

    // ------------------------------------------------------------ B. band filter + statistics, tile by tile
    // The filtered tile goes back to the row buffer un-normalised (one bulk store per tile, in place).
    double lsum = 0.0;
    float lmin = INFINITY, lmax = -INFINITY;
    if (tid < 4) sm.carry[tid] = 0.0;
    sk_fence_async_smem();                                // (the tile buffers held despike slots: bulk loads overwrite them)
    __syncthreads();
    {
      bool stored = false;
      for (int tile = cx.g; tile < ntiles; tile += kSkGroups) {
        const int t0 = tile * tl;
        const int n = min(tl, t_r - t0);
        if (stored) {                                     // my previous tile's bulk store has read the buffer
          if (cx.gt == 0) sk_bulk_store_wait();
          cx.sync();
        }
        if (nframes > 0) {
          if (!single) {                                  // the despiked tile comes back from the row buffer: one bulk copy
            if (cx.gt == 0) sk_bulk_load(cx.sig, rowbuf + t0, (uint32_t)((n + 3) >> 2) * 16u, bar);
            sk_bar_wait(bar, bulk_phase);
            bulk_phase ^= 1u;
          }
        } else {
          if (cx.gt == 32 && tile + kSkGroups < ntiles)   // the raw samples of my next tile: on their way to L2
            sk_prefetch_inputs<UP, DOWN, D>(xr, t_in_r, P.off, t0 + kSkGroups * tl, min(tl, t_r - t0 - kSkGroups * tl));
          sk_resample_tile<UP, DOWN, D, PS>(cx, xr, t_in_r, P.off, t0, n);
        }
        for (int i = n + cx.gt; i < kSkTile + kSkGuard; i += kSkGT) cx.sig[i] = 0.f;    // chunk grid beyond the row
        cx.sync();
        SK_STAMP(3);
        if (kind == 0) sk_filter_tile<0>(P, sm, cx, tile, ntiles, tile_warps, n, fix_nan, lsum, lmin, lmax);
        else sk_filter_tile<1>(P, sm, cx, tile, ntiles, tile_warps, n, fix_nan, lsum, lmin, lmax);
        SK_STAMP(4);
        if (!single && tile + kSkGroups < ntiles) {       // (my last tile of the row stays in shared memory for pass C)
          sk_fence_async_smem();
          cx.sync();
          if (cx.gt == 0) sk_bulk_store(rowbuf + t0, cx.sig, (uint32_t)((n + 3) >> 2) * 16u);
          stored = true;
        }
        SK_STAMP(5);
      }
      if (stored && cx.gt == 0) sk_bulk_store_wait();     // my tiles have reached the row buffer
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
    auto scaled = [inv_f, shift_f](float v) { return fminf(fmaxf(fmaf(v, inv_f, shift_f), -1.f), 1.f); };

    // ------------------------------------------------------------ C. normalise + write the row's windows
    if (single) {                                         // the filtered row is still in group 0's tile
      int kf = 0;
      sk_store_windows(pool + kSkGuard, obase, 0, t_r, P.start, P.win, P.hop, nwin, so_k, so_j, kf, tid, kSkThreads, scaled);
    } else {
      // my last tile of the row is still in shared memory: its windows leave from there ...
      const int ntiles_mine = cx.g < ntiles ? (ntiles - cx.g + kSkGroups - 1) / kSkGroups : 0;
      const int kept = cx.g + (ntiles_mine - 1) * kSkGroups;
      if (ntiles_mine > 0) {
        int kf0 = 0;
        sk_store_windows(cx.sig, obase, kept * tl, min(tl, t_r - kept * tl), P.start, P.win, P.hop, nwin, so_k, so_j, kf0, cx.gt,
                         kSkGT, scaled);
      }
      // ... the others come back from the row buffer in halves, two bulk copies in flight
      int kf = 0;
      const uint32_t bar2 = sk_smem(&cx.gs->bulk_bar2);
      constexpr int HALF = kSkTile / 2;
      // piece q = half (q & 1) of my (q >> 1)-th tile; piece q lands in half (q & 1) of the tile buffer
      int npieces = 0;
      for (int tile = cx.g; tile < kept; tile += kSkGroups) npieces += (min(tl, t_r - tile * tl) > HALF) ? 2 : 1;
      auto piece = [&](int q, int& s0, int& len, int& half) {
        // tiles are full (two pieces each) except possibly the last one
        const int tile = cx.g + (q >> 1) * kSkGroups;
        const int n = min(tl, t_r - tile * tl);
        half = q & 1;
        s0 = tile * tl + half * HALF;
        len = half ? n - HALF : min(n, HALF);
      };
      sk_fence_async_smem();
      cx.sync();                                          // the group is done with the buffer's previous content
      uint32_t ph0 = bulk_phase, ph1 = bulk_phase2;
      if (cx.gt == 0 && npieces > 0) {
        int s0, len, half;
        piece(0, s0, len, half);
        sk_bulk_load(cx.sig, rowbuf + s0, (uint32_t)((len + 3) >> 2) * 16u, bar);
      }
      for (int q = 0; q < npieces; ++q) {
        int s0, len, half;
        piece(q, s0, len, half);
        if (cx.gt == 0 && q + 1 < npieces) {              // the next piece flies while this one is written out
          int s1, len1, half1;
          piece(q + 1, s1, len1, half1);
          sk_bulk_load(cx.sig + half1 * HALF, rowbuf + s1, (uint32_t)((len1 + 3) >> 2) * 16u, half1 ? bar2 : bar);
        }
        if (half) { sk_bar_wait(bar2, ph1); ph1 ^= 1u; } else { sk_bar_wait(bar, ph0); ph0 ^= 1u; }
        sk_store_windows(cx.sig + half * HALF, obase, s0, len, P.start, P.win, P.hop, nwin, so_k, so_j, kf, cx.gt, kSkGT, scaled);
        sk_fence_async_smem();
        cx.sync();                                        // this half may be overwritten by the piece after next
      }
      bulk_phase = ph0; bulk_phase2 = ph1;
    }
    {                                                     // a short row's single window: zero padding past the row's end
      const long long span = (long long)nwin * P.win;
      if (P.start + (long long)(nwin - 1) * P.hop + P.win > t_r) {        // (only when nwin == 1)
        long long valid = t_r - P.start;
        if (valid < 0) valid = 0;
        for (long long i = valid + tid; i < span; i += kSkThreads) obase[i * so_j] = 0.f;
      }
    }
    SK_STAMP(6);
  }
#if defined(MPCG_FZ_PHASE_CLOCKS) && MPCG_FZ_PHASE_CLOCKS
  if (P.dbg && cx.gt == 0)
    for (int k = 0; k < 8; ++k) P.dbg[((long long)blockIdx.x * kSkGroups + cx.g) * 16 + k] = ph_acc[k];
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
