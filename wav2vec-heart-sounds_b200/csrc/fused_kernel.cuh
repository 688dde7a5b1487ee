// Kernel behind mpcg_preprocess_segment_f32 (fused.cu) -- the whole conditioning chain of one recording channel
//   resample -> Schmidt despike -> low-pass + high-pass -> abs-max normalise -> overlapping windows
// (torchproc.preprocess_pcg / preprocess_ecg + segment, signalproc/torchproc.py:101-129; NumPy twins
// signalproc/preprocess.py:24-37 + segment.py:40-52) in ONE kernel: the raw samples are read from HBM
// once, the windows are written once, every intermediate lives in shared memory.
//
// A row (one channel of one recording, up to ~1.5 MB after resampling) does not fit one SM, so a
// thread-block CLUSTER of ncl CTAs owns a row: CTA `rank` keeps samples [rank*S, rank*S + n) of the
// resampled signal in its shared memory.  S is a whole number of 500 ms despike frames, so a frame never
// straddles two CTAs.  Cross-CTA traffic goes over distributed shared memory and is tiny:
//   * despike: every CTA holds a copy of all frame maxima; per pass the owner of the worst frame flattens
//     it locally and broadcasts one float + one flag;
//   * filter:  each CTA exports E = its end state for a zero start state (4 doubles); start states chain
//     through the constant slice propagator A^S;
//   * normalise: (sum, min, max) per CTA.
// Pipes: the resampler is FFMA-immediate bound, the filter fp64-FMA bound, loads/stores HBM bound; two
// clusters per SM are kept resident so that different rows overlap different pipes.
#pragma once
#include <cooperative_groups.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include "resample.cuh"
#include "biquad.cuh"
#include "despike.cuh"

namespace cg = cooperative_groups;

namespace mpcg {

#ifndef MPCG_FZ_THREADS
#define MPCG_FZ_THREADS 512              // threads per CTA (512 x 2 CTAs/SM measured faster than 1024 x 1)
#endif
#ifndef MPCG_FZ_MINBLOCKS
#define MPCG_FZ_MINBLOCKS 2              // CTAs per SM the register allocation is sized for
#endif
#ifndef MPCG_FZ_FILTER_THREADS
#define MPCG_FZ_FILTER_THREADS 512       // threads that own a filter chunk (the rest idle through the fp64 phases)
#endif
constexpr int kFzThreads = MPCG_FZ_THREADS;
constexpr int kFzWarps = kFzThreads / 32;
constexpr int kFzChunks = MPCG_FZ_FILTER_THREADS;   // one filter chunk per filter thread
constexpr int kFzFW = kFzChunks / 32;               // filter warps
constexpr int kFzLmax = kFzChunks == 1024 ? 41 : (kFzChunks == 512 ? 81 : 163);   // longest chunk (odd)
constexpr int kFzMaxFrames = 64;         // despike frames per row the fused kernel accepts
constexpr int kFzMaxCluster = 8;
constexpr int kFzStageWords = 4608;       // resampler input staging: teams x buffers x block (largest: 8 x 565)
constexpr int kFzGuard = 40;              // floats before and after the slice that edge frames may spill into
// fast despike path (see the kernel): scratch copies of the frames being flattened + per-pass logs
constexpr int kFzScrWords = 2112;         // longest despike frame the fast path takes
constexpr int kFzUndoWords = 3072;        // per-CTA pool for the samples the fast path's passes overwrite
constexpr int kFzLogCap = 16;             // passes per frame and round it logs before handing over to the serial path
constexpr int kFzFastLocal = 16;          // frames per CTA the fast path takes
static_assert(kFzChunks == 1024 || kFzChunks == 512 || kFzChunks == 256, "filter threads: 1024, 512 or 256");
static_assert(kFzChunks <= kFzThreads, "filter threads are a subset of the CTA");

struct FzKind {                           // per channel kind (PCG / ECG): despike on/off + its filter
  int despike;
  int pad_;
  double c[2][5];                         // two sections, b0 b1 b2 a1 a2
  double wt[kFzLmax][4];                  // A^(L-1-j) B
  double mp[10][16];                      // M^(2^d), d = 0..9, M = A^L  (d >= 5 move whole warps)
  double prop_pow[kFzMaxCluster][16];     // (A^S)^j, j = 0..7: state across j full slices
  double prop_part[16];                   // A^nq: state across the valid part of the last chunk of a full slice
  double mlane[32][16];                   // M^lane, lane = 0..31
  double mwarp[kFzFW][16];                // M^(32 w), w = filter warp
};

struct FzParams {
  const float* x;
  float* out;
  int* edits;
  int* trace;
  long long* dbg;                         // optional [ctas, 16] clock64 stamps per phase (tools/ only)
  int trace_cap;
  int channels;                           // rows per recording
  int t_in, t;                            // samples per row before / after resampling
  int off;                                // resampler input offset
  int identity;                           // 1: no resampling (copy)
  int ncl, S, L, cap;                     // cluster size, slice length, chunk length, L * chunks
  int q, nq;                              // chunk holding the last sample of a full slice, valid samples in it
  int win_d, nframes, fpc;                // despike frame length, frames per row, frames per CTA
  double threshold;
  int max_iter, median_mode, norm_flags;
  int serial_despike;                     // 1: always take the serial (reference-order) despike path
  int start, win, hop, n;                 // window geometry
  long long so_b, so_c, so_k, so_j;       // output strides (elements): recording, channel, window, sample
  unsigned char kind_of_channel[8];
  const FzKind* kinds;                    // DEVICE pointer to the (up to two) channel recipes of this call
};

struct FzFilterScratch {                  // this row's recipe, copied from global memory once per CTA
  double mtab[16][32];                    // M^lane, element-major so a warp's loads are conflict-free
  double wt[kFzLmax][4];                  // pass-1 weights
  double mp[10][16];                      // M^(2^d)
  double prop_pow[kFzMaxCluster][16];
  double prop_part[16];
  double mwarp[kFzFW][16];
  double c[2][5];
  double wagg[kFzFW][4];                  // warp aggregates
  double wcar[kFzFW][4];                  // state at the start of each warp's first chunk (zero slice start)
};
constexpr int kFzMaxBlocks = 1344;        // 32-sample blocks in one slice (kFzLmax * kFzChunks / 32 + frames)
struct FzDespikeScratch {                 // serial path
  SpikeSorted sorted;
  float bmax[kFzMaxBlocks];
};
struct FzFastScratch {                    // fast path
  float undo[kFzUndoWords];               // overwritten samples of this round's passes (so that passes can be taken back)
  float bm[kFzFastLocal][kFzScrWords / 32 + 2];      // MY frames being flattened: 32-sample block maxima
  float seq[kFzFastLocal][kFzLogCap + 1];            // MY frames: maximum after j logged passes
  unsigned short span[kFzFastLocal][kFzLogCap][3];   // MY frames: [lo, hi) and undo offset of every logged pass
  float xtop[2][kFzMaxFrames];            // [round parity][slot] maximum the owner ended with   (owner -> every CTA)
  int xmeta[2][kFzMaxFrames];             // passes | stuck << 8 | log full << 9                 (owner -> every CTA)
  int xj[kFzMaxFrames];                   // passes that count, per slot, after a stuck round    (owner -> rank 0)
  unsigned char act_frame[kFzMaxFrames];  // slot -> frame (identical in every CTA)
  int nact, verdict, total, undo_used;
  float cutf, lo_mid, hi_mid;
  double cutd;
  unsigned long long kstar;
};
struct FzWarpStat {
  double sum;
  float lo, hi;
};
struct FzShared {
  union {                                 // phases that follow one another share this space:
    float xs[kFzStageWords];              //   resampler input staging
    FzFastScratch df;                     //   despike, fast path
    FzDespikeScratch d;                   //   despike, serial path: block maxima of my frames + sorted frame maxima
    FzFilterScratch f;                    //   filter recipe tables and scan scratch
  };
  double xE[kFzMaxCluster][4];            // end states exported by each rank
  FzWarpStat xstat[kFzMaxCluster][kFzWarps];   // (sum, min, max) of every warp of every rank (each rank holds a copy)
  float tops[2][kFzMaxFrames];            // serial path: double-buffered frame maxima
  float ftops[kFzMaxFrames];              // fast path: committed frame maxima (identical in every CTA)
  float fscr[40];
  int iscr[64];
  int ctrl[4];                            // (passes, done) published by the round owner, double-buffered
  int decision[2];                        // (active, worst) broadcast by warp 0
};

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

__device__ __forceinline__ void mv4_set(const double* __restrict__ m, const double (&v)[4], double (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    double a = m[r * 4] * v[0];
#pragma unroll
    for (int c = 1; c < 4; ++c) a = fma(m[r * 4 + c], v[c], a);
    out[r] = a;
  }
}

// acc += (M^lane) v with the element-major table: element (r, c) of lane's matrix sits at tab[r*4+c][lane].
__device__ __forceinline__ void mv4_lane_acc(const double (*tab)[32], int lane, const double (&v)[4], double (&acc)[4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    double a = acc[r];
#pragma unroll
    for (int c = 0; c < 4; ++c) a = fma(tab[r * 4 + c][lane], v[c], a);
    acc[r] = a;
  }
}

struct FzCoef {
  double b00, b01, b02, a01, a02, b10, b11, b12, a11, a12;
};
// One sample through both sections (transposed direct form II).
__device__ __forceinline__ double fz_step(const FzCoef& k, double (&z)[4], double xv) {
  const double y0 = fma(k.b00, xv, z[0]);
  z[0] = fma(-k.a01, y0, fma(k.b01, xv, z[1]));
  z[1] = fma(-k.a02, y0, k.b02 * xv);
  const double y1 = fma(k.b10, y0, z[2]);
  z[2] = fma(-k.a11, y1, fma(k.b11, y0, z[3]));
  z[3] = fma(-k.a12, y1, k.b12 * y0);
  return y1;
}
__device__ __forceinline__ float fz_round(double y, bool fix_nan) {
  float v = (float)y;
  if (fix_nan && !(fabsf(v) <= FLT_MAX)) v = (v != v) ? 0.f : (v > 0.f ? FLT_MAX : -FLT_MAX);
  return v;
}

__device__ __forceinline__ void fz_team_sync(int team, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(count) : "memory");
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const unsigned long long u = __shfl_xor_sync(kFull, v, o);
    v = u > v ? u : v;
  }
  return v;
}

// ---------------------------------------------------------------------------------------------------------
// Fast despike path, one frame by one warp.  The frame is flattened IN PLACE pass after pass while its maximum
// exceeds the round's threshold; every pass logs its span, the maximum it leaves behind and the samples it
// overwrote.  Which of the logged passes the reference's serial order really performs is decided after the
// cluster has exchanged the outcomes (see the kernel); the others are taken back from the undo log.
struct FzCut {                             // "frame maximum exceeds threshold * median" in the oracle's arithmetic
  int mode;
  float cutf;                              // tensor path: fp32 product, fp32 compare
  double cutd;                             // NumPy path: float64 product and compare
  __device__ __forceinline__ bool exceeds(float top) const {
    return mode == MPCG_MEDIAN_LOWER ? (top > cutf) : ((double)top > cutd);
  }
};
__device__ __forceinline__ void fz_fast_frame(FzShared& sm, cg::cluster_group& cluster, float* fr, int win, int fl,
                                              int gframe, int slot, int par, int ncl, const FzCut& cut) {
  const int lane = threadIdx.x & 31;
  float* bm = sm.df.bm[fl];
  const int nblk = (win + 31) >> 5;
  for (int b0 = 0; b0 < nblk; b0 += 4) {                    // block maxima, four blocks in flight
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
  float top = sm.ftops[gframe];
  if (lane == 0) sm.df.seq[fl][0] = top;
  int k = 0, stuck = 0, over = 0;
  while (cut.exceeds(top)) {
    if (k == kFzLogCap) { over = 1; break; }
    int peak, lo, hi;
    spike_find_span(fr, win, bm, nblk, top, peak, lo, hi);
    int off = 0;
    if (lane == 0) off = atomicAdd(&sm.df.undo_used, hi - lo);
    off = __shfl_sync(kFull, off, 0);
    if (off + (hi - lo) > kFzUndoWords) { over = 1; break; }   // pool exhausted: the pass is not performed
    bool changed;
    float new_top;
    spike_fill_span(fr, win, bm, nblk, lo, hi, changed, new_top, sm.df.undo + off);
    if (!changed) { stuck = 1; break; }      // a pass that moves nothing: the reference repeats it until max_iterations
    ++k;
    top = new_top;
    if (lane == 0) {
      sm.df.span[fl][k - 1][0] = (unsigned short)lo;
      sm.df.span[fl][k - 1][1] = (unsigned short)hi;
      sm.df.span[fl][k - 1][2] = (unsigned short)off;
      sm.df.seq[fl][k] = top;
    }
  }
  if (lane < ncl) {
    *cluster.map_shared_rank(&sm.df.xtop[par][slot], lane) = top;
    *cluster.map_shared_rank(&sm.df.xmeta[par][slot], lane) = k | (stuck << 8) | (over << 9);
  }
}

template <int UP, int DOWN, int D, int PS>
__global__ void __launch_bounds__(kFzThreads, MPCG_FZ_MINBLOCKS)
fused_preprocess_kernel(const __grid_constant__ FzParams P) {
  extern __shared__ __align__(16) unsigned char fz_raw[];
  FzShared& sm = *reinterpret_cast<FzShared*>(fz_raw);
  float* sig = reinterpret_cast<float*>(fz_raw + sizeof(FzShared)) + kFzGuard;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster.block_rank();
  const unsigned row = blockIdx.x / (unsigned)P.ncl;     // (the launcher keeps rows * cluster below 2^31)
  const int ch = (int)(row % (unsigned)P.channels);
  const long long rec = row / (unsigned)P.channels;
  const FzKind* __restrict__ Kg = P.kinds + P.kind_of_channel[ch];
  const int k_despike = Kg->despike;
  const int s0 = rank * P.S;
  int n = (rank == P.ncl - 1) ? (P.t - s0) : P.S;
  if (n < 0) n = 0;
  const int L = P.L;
  // per-phase clock stamps: compiled in only for tools/phase_clocks.py (make FZFLAGS=-DMPCG_FZ_PHASE_CLOCKS=1)
#if defined(MPCG_FZ_PHASE_CLOCKS) && MPCG_FZ_PHASE_CLOCKS
  int dbg_k = 0;
  auto stamp = [&]() {
    if (P.dbg && tid == 0) P.dbg[(long long)blockIdx.x * 16 + dbg_k] = clock64();
    ++dbg_k;
  };
#else
  auto stamp = [] {};
#endif
  stamp();                                                // 0: start

  // ---------------------------------------------------------------- 1. resample my slice into shared memory
  // Teams of PS warps walk the slice in blocks of 32 frames: coalesced loads -> registers (prefetched one block
  // ahead) -> skewed staging -> each warp of the team computes its phase group -> the slice.  Edge frames spill
  // into the guard floats around the slice instead of being bounds-checked.
  const float* xr = P.x + (long long)row * P.t_in;
  if constexpr (UP == DOWN) {                             // no resampling: plain copy of my slice
    for (int i = tid; i < n; i += kFzThreads) sig[i] = ld_stream(xr + s0 + i);
  } else if (n > 0) {
    using T = RsTeam<UP, DOWN, D, PS>;
    constexpr int NTEAMS = kFzWarps / PS;
    constexpr int NBUF = (NTEAMS * 2 * T::WORDS <= kFzStageWords) ? 2 : 1;
    static_assert(kFzWarps % PS == 0 && (PS == 1 || NTEAMS <= 15), "teams map onto named barriers 1..15");
    static_assert(NTEAMS * NBUF * T::WORDS <= kFzStageWords, "staging buffer too small for this resampler instance");
    static_assert(UP - 1 <= kFzGuard, "guard too small");
    const int team = warp / PS, grp = warp - team * PS, tt = tid - team * T::TEAM;
    float* xs_team = sm.xs + team * (NBUF * T::WORDS);
    const int f_lo = s0 / UP, f_hi = (s0 + n - 1) / UP;
    const int nblk = (f_hi - f_lo + T::FB) / T::FB;
    float pre[T::NPRE];
    int blk = team, buf = 0;
    if (blk < nblk) T::fetch(pre, xr, (long long)(f_lo + blk * T::FB) * DOWN + P.off, P.t_in, tt);
    for (; blk < nblk; blk += NTEAMS) {
      float* xs = xs_team + buf * T::WORDS;
      T::commit(xs, pre, tt);
      if constexpr (PS == 1) __syncwarp(); else fz_team_sync(team, T::TEAM);
      if (blk + NTEAMS < nblk)                            // next block's loads fly while this one is computed
        T::fetch(pre, xr, (long long)(f_lo + (blk + NTEAMS) * T::FB) * DOWN + P.off, P.t_in, tt);
      const int f = f_lo + blk * T::FB + lane;
      if (f <= f_hi) T::template dispatch<0>(grp, xs, lane, sig + (f * UP - s0));
      if constexpr (NBUF == 2) buf ^= 1;
      else if constexpr (PS == 1) __syncwarp();
      else fz_team_sync(team, T::TEAM);
    }
  }
  __syncthreads();
  stamp();                                                // 1: slice resampled
  for (int i = n + tid; i < P.cap + kFzGuard; i += kFzThreads) sig[i] = 0.f;      // chunk grid beyond the slice
  __syncthreads();
  stamp();                                                // 2: tail cleared

  // ---------------------------------------------------------------- 2. Schmidt despike (cluster-wide)
  int passes = 0;
  if (k_despike && P.nframes > 0) {
    const int gf0 = rank * P.fpc;
    int nloc = P.nframes - gf0;
    nloc = nloc < 0 ? 0 : (nloc > P.fpc ? P.fpc : nloc);
    bool serial = P.serial_despike || P.trace != nullptr || P.win_d > kFzScrWords || P.win_d < 2;
    // ------------------------------------------------------------ 2a. fast path
    // The reference flattens one span per pass, always in the frame holding the largest maximum, until no frame
    // exceeds threshold * median.  A pass only changes its own frame, and the threshold can only fall, so the
    // ORDER of the passes does not matter for the samples: every frame above the current threshold must be
    // flattened until it is not, whatever happens elsewhere.  Rounds: (1) every CTA derives the threshold and the
    // list of frames above it from its copy of the frame maxima; (2) the owners flatten those frames on scratch
    // copies, in parallel, logging each pass; (3) the outcomes are exchanged (one cluster barrier) and the passes
    // that the serial order really performs are committed to the signal.  All of them are, except when some frame
    // got STUCK (a pass that moves nothing, e.g. a one-sample spike between two sign flips: the reference then
    // repeats that pass until max_iterations): the serial order reaches the stuck state with the largest
    // (maximum, first index) key K* and never leaves it, so exactly the logged passes that started from a key
    // above K* happen, and the row is finished.  Anything the logs cannot settle exactly (pass budget in reach,
    // log full) continues on the serial path from the committed state, which is always a state of the serial order.
    if (!serial && P.fpc > kFzFastLocal) serial = true;
    if (!serial) {
      for (int f = warp; f < nloc; f += kFzWarps) {         // maxima of my frames -> every CTA
        const float* fr = sig + f * P.win_d;
        float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
        int i = lane;
        for (; i + 96 < P.win_d; i += 128) {
          m0 = fmaxf(m0, fabsf(fr[i])); m1 = fmaxf(m1, fabsf(fr[i + 32]));
          m2 = fmaxf(m2, fabsf(fr[i + 64])); m3 = fmaxf(m3, fabsf(fr[i + 96]));
        }
        for (; i < P.win_d; i += 32) m0 = fmaxf(m0, fabsf(fr[i]));
        const float m = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)))));
        if (lane < P.ncl) *cluster.map_shared_rank(&sm.ftops[gf0 + f], lane) = m;
      }
      cluster_arrive();
      cluster_wait();
      for (int round = 0;; ++round) {
        const int par = round & 1;
        if (warp == 0) {                                    // threshold + frames above it (same bits in every CTA)
          const int nf = P.nframes;
          const float v0 = lane < nf ? sm.ftops[lane] : -1.f, v1 = lane + 32 < nf ? sm.ftops[lane + 32] : -1.f;
          int less0 = 0, leq0 = 0, less1 = 0, leq1 = 0;
#pragma unroll 4
          for (int j = 0; j < nf; ++j) {
            const float vj = sm.ftops[j];
            less0 += (vj < v0); leq0 += (vj <= v0);
            less1 += (vj < v1); leq1 += (vj <= v1);
          }
          const int k_lo = (nf - 1) >> 1, k_hi = nf >> 1;
          float c_lo = -1.f, c_hi = -1.f;                   // maxima are >= 0, so -1 means "not mine"
          if (lane < nf) {
            if (less0 <= k_lo && k_lo < leq0) c_lo = v0;
            if (less0 <= k_hi && k_hi < leq0) c_hi = v0;
          }
          if (lane + 32 < nf) {
            if (less1 <= k_lo && k_lo < leq1) c_lo = fmaxf(c_lo, v1);
            if (less1 <= k_hi && k_hi < leq1) c_hi = fmaxf(c_hi, v1);
          }
          const float lo_mid = warp_max(c_lo), hi_mid = warp_max(c_hi);
          FzCut cut;
          cut.mode = P.median_mode;
          cut.cutf = __fmul_rn((float)P.threshold, lo_mid);
          const double med = ((double)lo_mid + (double)hi_mid) * 0.5;
          cut.cutd = P.threshold * med;
          const bool open = passes < P.max_iter && (P.median_mode == MPCG_MEDIAN_LOWER || med != 0.0);
          const bool a0 = open && lane < nf && cut.exceeds(v0), a1 = open && lane + 32 < nf && cut.exceeds(v1);
          const unsigned m0 = __ballot_sync(kFull, a0), m1 = __ballot_sync(kFull, a1);
          const unsigned below = (1u << lane) - 1u;
          if (a0) sm.df.act_frame[__popc(m0 & below)] = (unsigned char)lane;
          if (a1) sm.df.act_frame[__popc(m0) + __popc(m1 & below)] = (unsigned char)(lane + 32);
          if (lane == 0) {
            sm.df.nact = __popc(m0) + __popc(m1); sm.df.cutf = cut.cutf; sm.df.cutd = cut.cutd;
            sm.df.lo_mid = lo_mid; sm.df.hi_mid = hi_mid; sm.df.undo_used = 0;
          }
        }
        __syncthreads();
        const int nact = sm.df.nact;
        if (nact == 0) break;
        {                                                   // my frames of the list, one warp each
          FzCut cut;
          cut.mode = P.median_mode; cut.cutf = sm.df.cutf; cut.cutd = sm.df.cutd;
          int ord = 0;
          for (int slot = 0; slot < nact; ++slot) {
            const int f = sm.df.act_frame[slot];
            if (f < gf0 || f >= gf0 + nloc) continue;
            if ((ord++ % kFzWarps) != warp) continue;
            fz_fast_frame(sm, cluster, sig + (f - gf0) * P.win_d, P.win_d, f - gf0, f, slot, par, P.ncl, cut);
          }
        }
        cluster_arrive();
        cluster_wait();
        if (warp == 0) {                                    // the round's outcome (same bits in every CTA)
          const bool h0 = lane < nact, h1 = lane + 32 < nact;
          const int me0 = h0 ? sm.df.xmeta[par][lane] : 0, me1 = h1 ? sm.df.xmeta[par][lane + 32] : 0;
          const bool over = __any_sync(kFull, ((me0 | me1) >> 9) & 1);
          unsigned long long ks = 0ull;
          if ((me0 >> 8) & 1) ks = spike_key(sm.df.xtop[par][lane], sm.df.act_frame[lane]);
          if ((me1 >> 8) & 1) {
            const unsigned long long k1 = spike_key(sm.df.xtop[par][lane + 32], sm.df.act_frame[lane + 32]);
            ks = k1 > ks ? k1 : ks;
          }
          const unsigned long long kstar = warp_max_u64(ks);
          int total = (me0 & 0xff) + (me1 & 0xff);
#pragma unroll
          for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(kFull, total, o);
          // (with a stuck frame fewer passes may count, plus the stuck pass itself: total + 1 bounds both cases)
          const bool hand_over = over || (long long)passes + total + (kstar != 0ull ? 1 : 0) > (long long)P.max_iter;
          // If no flattened frame fell below the old middle value(s), the median and with it the threshold are what
          // they were: nothing else can exceed it and the row is finished without another round.
          const float floor_v = P.median_mode == MPCG_MEDIAN_LOWER ? sm.df.lo_mid : sm.df.hi_mid;
          const bool same_median = __all_sync(kFull, (!h0 || sm.df.xtop[par][lane] >= floor_v) &&
                                                         (!h1 || sm.df.xtop[par][lane + 32] >= floor_v));
          if (!hand_over && kstar == 0ull) {
            if (h0) sm.ftops[sm.df.act_frame[lane]] = sm.df.xtop[par][lane];
            if (h1) sm.ftops[sm.df.act_frame[lane + 32]] = sm.df.xtop[par][lane + 32];
          }
          if (lane == 0) {
            sm.df.verdict = hand_over ? 2 : (kstar != 0ull ? 1 : (same_median ? 3 : 0));
            sm.df.total = total;
            sm.df.kstar = kstar;
          }
        }
        __syncthreads();
        const int verdict = sm.df.verdict;
        const unsigned long long kstar = sm.df.kstar;
        for (int slot = warp; slot < nact; slot += kFzWarps) {     // take back my frames' passes that do not happen
          const int f = sm.df.act_frame[slot];
          if (f < gf0 || f >= gf0 + nloc) continue;
          const int fl = f - gf0;
          float* fr = sig + fl * P.win_d;
          const int k = sm.df.xmeta[par][slot] & 0xff;
          int jc = k;
          if (verdict == 2) jc = 0;                         // the serial path restarts from the committed state
          if (verdict == 1) {                               // only passes that started above K* (all of the stuck frame's)
            if (spike_key(sm.df.seq[fl][k], f) != kstar)
              jc = __popc(__ballot_sync(kFull, lane < k && spike_key(sm.df.seq[fl][lane < k ? lane : 0], f) > kstar));
            if (P.edits && lane == 0) *cluster.map_shared_rank(&sm.df.xj[slot], 0) = jc;
          }
          for (int j = k - 1; j >= jc; --j) {               // newest first: spans of later passes may cover earlier fills
            const int lo = sm.df.span[fl][j][0], hi = sm.df.span[fl][j][1], off = sm.df.span[fl][j][2];
            for (int i = lo + lane; i < hi; i += 32) fr[i] = sm.df.undo[off + i - lo];
            __syncwarp();
          }
        }
        if (verdict == 2) { serial = true; break; }
        if (verdict == 1) {
          if (P.edits) {                                    // exact pass count: rank 0 adds up what really counted
            cluster_arrive();
            cluster_wait();
            if (rank == 0 && warp == 0) {
              int t = (lane < nact ? sm.df.xj[lane] : 0) + (lane + 32 < nact ? sm.df.xj[lane + 32] : 0);
#pragma unroll
              for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(kFull, t, o);
              if (lane == 0) sm.df.total = t + 1;           // + the pass that moves nothing
            }
            __syncthreads();
            if (rank == 0) passes += sm.df.total;
          }
          break;
        }
        passes += sm.df.total;
        if (verdict == 3) break;
        __syncthreads();
      }
    }
    // ------------------------------------------------------------ 2b. serial path (reference order, pass by pass)
    // Every CTA keeps two copies of all frame maxima.  Round k reads copy k&1.  Warp 0 of every CTA sorts the
    // maxima (identical data, identical verdict).  The CTA owning the worst frame then keeps going ON ITS OWN:
    // flatten, update its local maxima, decide again -- pass after pass with no barrier at all, as long as the
    // worst frame stays one of its own.  When ownership moves (or nothing is left to do) it publishes its frames'
    // maxima and the pass counter into copy (k+1)&1 of every CTA; one cluster barrier per round.
    if (serial) {
      __syncthreads();
      const int nblk = (P.win_d + 31) >> 5;                   // 32-sample blocks per frame
      for (int q = warp; q < nloc * nblk; q += kFzWarps) {    // block maxima of my frames
        const int f = q / nblk, b = q - f * nblk;
        const int i = b * 32 + lane;
        const float m = warp_max(i < P.win_d ? fabsf(sig[f * P.win_d + i]) : 0.f);
        if (lane == 0) sm.d.bmax[q] = m;
      }
      __syncthreads();
      for (int f = warp; f < nloc; f += kFzWarps) {           // frame maxima -> every CTA's copy 0
        float m = 0.f;
        for (int b = lane; b < nblk; b += 32) m = fmaxf(m, sm.d.bmax[f * nblk + b]);
        m = warp_max(m);
        if (lane < P.ncl) *cluster.map_shared_rank(&sm.tops[0][gf0 + f], lane) = m;
      }
      cluster_arrive();
      cluster_wait();
      for (int round = 0;; ++round) {
        float* cur = sm.tops[round & 1];
        float* nxt = sm.tops[(round + 1) & 1];
        if (warp == 0) {
          bool idle = passes >= P.max_iter;                   // budget spent, or (below) nothing exceeds the threshold
          SpikeDecision dec;
          dec.active = false; dec.worst = 0;
          if (!idle) {
            spike_sort_init(sm.d.sorted, cur, P.nframes);
            dec = spike_sort_decide(sm.d.sorted, P.nframes, P.threshold, P.median_mode);
            idle = !dec.active;
          }
          if (lane == 0) sm.decision[round & 1] = idle ? 1 : 0;   // same verdict in every CTA
          if (!idle) {
            const int owner = dec.worst / P.fpc;
            const int of0 = owner * P.fpc;
            for (int i = lane; i < P.nframes; i += 32)        // carry the other CTAs' entries over locally
              if (i < of0 || i >= of0 + P.fpc) nxt[i] = cur[i];
            if (rank == owner) {
              bool done = false;
              while (true) {                                  // local passes on my own frames, no barrier
                const int fl = dec.worst - gf0;
                const float old_top = cur[dec.worst];
                int peak, lo, hi;
                bool changed;
                float new_top;
                spike_pass_warp(sig + fl * P.win_d, P.win_d, sm.d.bmax + fl * nblk, nblk, old_top, peak, lo, hi, changed,
                                new_top);
                if (lane == 0) {
                  cur[dec.worst] = new_top;
                  if (P.trace && passes < P.trace_cap) {
                    int* tr = P.trace + ((long long)row * P.trace_cap + passes) * 4;
                    tr[0] = dec.worst; tr[1] = peak; tr[2] = lo; tr[3] = hi;
                  }
                }
                __syncwarp();
                ++passes;
                if (!changed || passes >= P.max_iter) { done = true; break; }   // fixed point / budget
                spike_sort_update(sm.d.sorted, dec.worst, old_top, new_top);
                dec = spike_sort_decide(sm.d.sorted, P.nframes, P.threshold, P.median_mode);
                if (!dec.active) { done = true; break; }
                if (dec.worst / P.fpc != rank) break;         // somebody else's frame: hand over
              }
              for (int f = lane; f < nloc * P.ncl; f += 32) { // publish my frames' maxima to every CTA
                const int fi = f % nloc, rk = f / nloc;
                *cluster.map_shared_rank(&nxt[gf0 + fi], rk) = cur[gf0 + fi];
              }
              if (lane < P.ncl) {
                int* c = cluster.map_shared_rank(&sm.ctrl[2 * ((round + 1) & 1)], lane);
                c[0] = passes;
                c[1] = done ? 1 : 0;
              }
            }
          }
        }
        cluster_arrive();
        cluster_wait();
        if (sm.decision[round & 1]) break;                    // nothing to do this round: every CTA stops together
        passes = sm.ctrl[2 * ((round + 1) & 1)];
        if (sm.ctrl[2 * ((round + 1) & 1) + 1]) break;
      }
    }
    if (P.edits && rank == 0 && tid == 0) P.edits[row] = passes;
    __syncthreads();
  } else if (P.edits && rank == 0 && tid == 0) {
    P.edits[row] = 0;
  }

  stamp();                                                // 3: despiked
  // recipe tables: global (L2-resident, ~11 KB) -> shared, coalesced; everything below reads shared memory
  for (int i = tid; i < 512; i += kFzThreads) sm.f.mtab[i & 15][i >> 4] = (&Kg->mlane[0][0])[i];
  for (int i = tid; i < L * 4; i += kFzThreads) (&sm.f.wt[0][0])[i] = (&Kg->wt[0][0])[i];
  for (int i = tid; i < 160; i += kFzThreads) (&sm.f.mp[0][0])[i] = (&Kg->mp[0][0])[i];
  for (int i = tid; i < kFzMaxCluster * 16; i += kFzThreads) (&sm.f.prop_pow[0][0])[i] = (&Kg->prop_pow[0][0])[i];
  for (int i = tid; i < kFzFW * 16; i += kFzThreads) (&sm.f.mwarp[0][0])[i] = (&Kg->mwarp[0][0])[i];
  if (tid < 16) sm.f.prop_part[tid] = Kg->prop_part[tid];
  if (tid < 10) (&sm.f.c[0][0])[tid] = (&Kg->c[0][0])[tid];
  __syncthreads();
  // ---------------------------------------------------------------- 3. low-pass + high-pass as one 4-state scan
  // (threads beyond kFzChunks own no chunk: their loops are empty and their scan inputs zero)
  const bool filt = tid < kFzChunks;
  float* mine = sig + (filt ? tid : 0) * L;
  const int Lm = filt ? L : 0;
  double p[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
  for (int j = 0; j < Lm; ++j) {
    const double2 w01 = *reinterpret_cast<const double2*>(&sm.f.wt[j][0]);
    const double2 w23 = *reinterpret_cast<const double2*>(&sm.f.wt[j][2]);
    const double xv = (double)mine[j];
    p[0] = fma(w01.x, xv, p[0]);
    p[1] = fma(w01.y, xv, p[1]);
    p[2] = fma(w23.x, xv, p[2]);
    p[3] = fma(w23.y, xv, p[3]);
  }
  stamp();                                                // 4: pass 1 done
  // zero-state response of the partial last chunk of a full slice (the piece of E that the scan does not give):
  // the exporting thread's whole warp shares the dot product
  const bool exporter = (P.ncl > 1) && (rank < P.ncl - 1) && (tid == P.q);
  double pp[4] = {0.0, 0.0, 0.0, 0.0};
  if ((P.ncl > 1) && (rank < P.ncl - 1) && (warp == (P.q >> 5))) {
    const float* qs = sig + P.q * L;
    const int shift = L - P.nq;
    for (int j = lane; j < P.nq; j += 32) {
      const double xv = (double)qs[j];
#pragma unroll
      for (int s = 0; s < 4; ++s) pp[s] = fma(sm.f.wt[j + shift][s], xv, pp[s]);
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) pp[s] = warp_sum(pp[s]);
  }
#pragma unroll
  for (int d = 0; d < 5; ++d) {                           // inclusive scan inside the warp
    double u[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) u[s] = __shfl_up_sync(kFull, p[s], 1 << d);
    if (lane >= (1 << d)) mv4_acc(sm.f.mp[d], u, p);
  }
  if (lane == 31 && filt) {
#pragma unroll
    for (int s = 0; s < 4; ++s) sm.f.wagg[warp][s] = p[s];
  }
  __syncthreads();
  if (warp == 0) {                                        // scan the warp aggregates in one warp
    double v[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) v[s] = (lane < kFzFW) ? sm.f.wagg[lane][s] : 0.0;
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      double u[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) u[s] = __shfl_up_sync(kFull, v[s], 1 << d);
      if (lane >= (1 << d)) mv4_acc(sm.f.mp[5 + d], u, v);
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const double e = __shfl_up_sync(kFull, v[s], 1);
      if (lane < kFzFW) sm.f.wcar[lane][s] = lane ? e : 0.0;
    }
  }
  __syncthreads();
  stamp();                                                // 5: intra-CTA scan done
  double z[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const double e = __shfl_up_sync(kFull, p[s], 1);
    z[s] = lane ? e : 0.0;
  }
  {
    double wc[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) wc[s] = sm.f.wcar[filt ? warp : 0][s];
    mv4_lane_acc(sm.f.mtab, lane, wc, z);                     // chunk start state for a zero slice start
  }
  if (P.ncl > 1) {
    if (exporter) {
      double e[4] = {pp[0], pp[1], pp[2], pp[3]};
      mv4_acc(sm.f.prop_part, z, e);                         // E = A^nq start_q + partial response
      for (int rk = 0; rk < P.ncl; ++rk) {
        double* dst = cluster.map_shared_rank(&sm.xE[rank][0], rk);
#pragma unroll
        for (int s = 0; s < 4; ++s) dst[s] = e[s];
      }
    }
    cluster_arrive();
    cluster_wait();
    // true slice start state c_rank = sum_k (A^S)^(rank-1-k) E_k (independent products, two accumulators),
    // then carried to my chunk: M^lane (M^(32 warp) c)
    double c0[4] = {0.0, 0.0, 0.0, 0.0}, c1[4] = {0.0, 0.0, 0.0, 0.0};
    for (int r = 0; r < rank; r += 2) {
      double e0[4], e1[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) { e0[s] = sm.xE[r][s]; e1[s] = (r + 1 < rank) ? sm.xE[r + 1][s] : 0.0; }
      mv4_acc(sm.f.prop_pow[rank - 1 - r], e0, c0);
      if (r + 1 < rank) mv4_acc(sm.f.prop_pow[rank - 2 - r], e1, c1);
    }
    double c[4];
    {
      double cs[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) cs[s] = c0[s] + c1[s];
      mv4_set(sm.f.mwarp[filt ? warp : 0], cs, c);
    }
    mv4_lane_acc(sm.f.mtab, lane, c, z);
  }
  stamp();                                                // 6: cluster carry applied
  // pass 2; statistics of the valid outputs ride along
  double lsum = 0.0;
  float lmin = INFINITY, lmax = -INFINITY;
  {
    FzCoef kc;
    kc.b00 = sm.f.c[0][0]; kc.b01 = sm.f.c[0][1]; kc.b02 = sm.f.c[0][2]; kc.a01 = sm.f.c[0][3]; kc.a02 = sm.f.c[0][4];
    kc.b10 = sm.f.c[1][0]; kc.b11 = sm.f.c[1][1]; kc.b12 = sm.f.c[1][2]; kc.a11 = sm.f.c[1][3]; kc.a12 = sm.f.c[1][4];
    const bool fix_nan = (P.norm_flags & MPCG_NORM_NAN_TO_NUM) != 0;
    int lim = n - tid * L;
    lim = (lim < 0 || !filt) ? 0 : (lim > L ? L : lim);
    // The non-finite fix-up of nan_to_num is kept out of the loop: a NaN or infinity anywhere in the chunk turns the
    // chunk's sum non-finite, which sends this thread through a second sweep below.
    int j = 0;
    for (; j + 8 <= lim; j += 8) {
      float sacc = 0.f;
#pragma unroll
      for (int u = 0; u < 8; u += 2) {
        const float v0 = (float)fz_step(kc, z, (double)mine[j + u]);
        const float v1 = (float)fz_step(kc, z, (double)mine[j + u + 1]);
        mine[j + u] = v0;
        mine[j + u + 1] = v1;
        sacc += v0;
        sacc += v1;
        lmin = fminf(lmin, fminf(v0, v1));                // three-input min / max
        lmax = fmaxf(lmax, fmaxf(v0, v1));
      }
      lsum += (double)sacc;
    }
    {
      float sacc = 0.f;
      for (; j < lim; ++j) {
        const float v = (float)fz_step(kc, z, (double)mine[j]);
        mine[j] = v;
        sacc += v;
        lmin = fminf(lmin, v);
        lmax = fmaxf(lmax, v);
      }
      lsum += (double)sacc;
    }
    if (fix_nan && !(fabs(lsum) < (double)INFINITY)) {
      lsum = 0.0; lmin = INFINITY; lmax = -INFINITY;
      for (j = 0; j < lim; j += 8) {
        float sacc = 0.f;
        for (int u = 0; u < 8 && j + u < lim; ++u) {
          const float v = fz_round((double)mine[j + u], true);
          mine[j + u] = v;
          sacc += v;
          lmin = fminf(lmin, v);
          lmax = fmaxf(lmax, v);
        }
        lsum += (double)sacc;
      }
    }
  }

  stamp();                                                // 7: pass 2 done
  // ---------------------------------------------------------------- 4. row statistics across the cluster
  // every warp hands its (sum, min, max) straight to every rank; after ONE cluster barrier warp 0 folds the
  // cluster's ncl x warps partials (fixed order: bit-reproducible)
  lsum = warp_sum(lsum);
  lmin = warp_min(lmin);
  lmax = warp_max(lmax);
  if (lane < P.ncl) {
    FzWarpStat* dst = cluster.map_shared_rank(&sm.xstat[rank][warp], lane);
    dst->sum = lsum; dst->lo = lmin; dst->hi = lmax;
  }
  cluster_arrive();
  cluster_wait();
  if (warp == 0) {                                        // one warp turns the cluster's statistics into the map
    double tot = 0.0;
    float lo_f = INFINITY, hi_f = -INFINITY;
    const FzWarpStat* all = &sm.xstat[0][0];
    for (int e = lane; e < P.ncl * kFzWarps; e += 32) {
      tot += all[e].sum;
      lo_f = fminf(lo_f, all[e].lo);
      hi_f = fmaxf(hi_f, all[e].hi);
    }
    tot = warp_sum(tot);
    const double lo_all = (double)warp_min(lo_f), hi_all = (double)warp_max(hi_f);
    const double mean = tot / (double)P.t;
    const double peak = fmax(hi_all - mean, mean - lo_all);
    double inv_peak;
    if (P.norm_flags & MPCG_NORM_PEAK_GT0) inv_peak = (peak > 0.0) ? 1.0 / peak : 1.0;
    else inv_peak = 1.0 / fmax(peak, 1e-12);
    // fp32 map: the mean is split hi + lo so (s - hi) - lo carries no cancellation error
    const float mh = (float)mean;
    if (lane == 0) { sm.fscr[0] = mh; sm.fscr[1] = (float)(mean - (double)mh); sm.fscr[2] = (float)inv_peak; }
  }
  __syncthreads();
  const float mean_hi = sm.fscr[0], mean_lo = sm.fscr[1], inv_f = sm.fscr[2];

  stamp();                                                // 8: statistics exchanged
  // ---------------------------------------------------------------- 5. normalise + write my share of every window
  float* obase = P.out + rec * P.so_b + ch * P.so_c;
  const int s1 = s0 + n;
  auto scaled = [&](float s) { return fminf(fmaxf(((s - mean_hi) - mean_lo) * inv_f, -1.f), 1.f); };
  // windows that intersect my slice: k_first .. k_last
  int k_first = s0 - P.start - P.win + 1;
  k_first = k_first > 0 ? (k_first + P.hop - 1) / P.hop : 0;
  int k_last = s1 - 1 - P.start;
  k_last = k_last < 0 ? -1 : k_last / P.hop;
  if (k_last > P.n - 1) k_last = P.n - 1;
  for (int k = k_first; k <= k_last; ++k) {
    const int w0 = P.start + k * P.hop;
    const int a = w0 > s0 ? w0 : s0;
    const int w1 = w0 + P.win;
    const int b = w1 < s1 ? w1 : s1;
    const int len = b - a;
    const float* sp = sig + (a - s0);
    if (P.so_j == 1) {
      float* dp = obase + k * P.so_k + (a - w0);
      // 64-bit stores: one head element if dp is odd, pairs, one tail element
      int head = (int)((reinterpret_cast<uintptr_t>(dp) >> 2) & 1u);
      if (head > len) head = len;
      if (head && tid == 0) st_stream(dp, scaled(sp[0]));
      const int npair = (len - head) >> 1;
      const float* sq = sp + head;
      float2* dq = reinterpret_cast<float2*>(dp + head);
      if ((reinterpret_cast<uintptr_t>(sq) & 7u) == 0) {    // shared side aligned too: 64-bit loads
        const float2* sq2 = reinterpret_cast<const float2*>(sq);
        int i = tid;
        for (; i + kFzThreads < npair; i += 2 * kFzThreads) {
          const float2 p0 = sq2[i], p1 = sq2[i + kFzThreads];
          st_stream2(dq + i, make_float2(scaled(p0.x), scaled(p0.y)));
          st_stream2(dq + i + kFzThreads, make_float2(scaled(p1.x), scaled(p1.y)));
        }
        for (; i < npair; i += kFzThreads) {
          const float2 p0 = sq2[i];
          st_stream2(dq + i, make_float2(scaled(p0.x), scaled(p0.y)));
        }
      } else {
        int i = tid;
        for (; i + kFzThreads < npair; i += 2 * kFzThreads) {
          const float a0 = sq[2 * i], a1 = sq[2 * i + 1], b0 = sq[2 * (i + kFzThreads)], b1 = sq[2 * (i + kFzThreads) + 1];
          st_stream2(dq + i, make_float2(scaled(a0), scaled(a1)));
          st_stream2(dq + i + kFzThreads, make_float2(scaled(b0), scaled(b1)));
        }
        for (; i < npair; i += kFzThreads) st_stream2(dq + i, make_float2(scaled(sq[2 * i]), scaled(sq[2 * i + 1])));
      }
      if (((len - head) & 1) && tid == 32) st_stream(dp + len - 1, scaled(sp[len - 1]));
    } else {
      float* dp = obase + k * P.so_k + (long long)(a - w0) * P.so_j;
      for (int i = tid; i < len; i += kFzThreads) dp[(long long)i * P.so_j] = scaled(sp[i]);
    }
  }
  if (rank == P.ncl - 1 && P.n > 0 && P.start + (P.n - 1) * P.hop + P.win > P.t) {
    for (int k = 0; k < P.n; ++k) {                       // short recording: zero-fill past its end
      const int w0 = P.start + k * P.hop, w1 = w0 + P.win;
      if (w1 <= P.t) continue;
      float* dst = obase + k * P.so_k;
      const int z0 = (P.t > w0 ? P.t : w0);
      for (int i = z0 + tid; i < w1; i += kFzThreads) dst[(long long)(i - w0) * P.so_j] = 0.f;
    }
  }
  stamp();                                                // 9: windows stored (issued)
  // No trailing cluster barrier: every remote shared-memory access of this kernel happens before the
  // statistics barrier above, so a CTA may retire while its peers are still storing their windows.
}

template <int UP, int DOWN, int D, int PS>
int fz_launch(const FzParams& P, size_t smem, long long rows, cudaStream_t stream) {
  auto kern = fused_preprocess_kernel<UP, DOWN, D, PS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(rows * P.ncl));
  cfg.blockDim = dim3(kFzThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)P.ncl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (const char* e = getenv("MPCG_FZ_CLUSTER_POLICY")) {   // experiments: 1 = spread, 2 = load balancing
    const int v = atoi(e);
    if (v == 1 || v == 2) {
      attr[1].id = cudaLaunchAttributeClusterSchedulingPolicyPreference;
      attr[1].val.clusterSchedulingPolicyPreference = v == 1 ? cudaClusterSchedulingPolicySpread : cudaClusterSchedulingPolicyLoadBalancing;
      cfg.numAttrs = 2;
    }
  }
  e = cudaLaunchKernelEx(&cfg, kern, P);
  if (e != cudaSuccess) return (int)e;
  return MPCG_OK;
}


}  // namespace mpcg
