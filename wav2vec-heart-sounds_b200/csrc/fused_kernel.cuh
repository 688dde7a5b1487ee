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
constexpr int kFzThreads = MPCG_FZ_THREADS;
constexpr int kFzWarps = kFzThreads / 32;
constexpr int kFzChunks = kFzThreads;     // one filter chunk per thread
constexpr int kFzLmax = 81;               // longest chunk (odd)
constexpr int kFzMaxFrames = 64;         // despike frames per row the fused kernel accepts
constexpr int kFzMaxCluster = 8;
constexpr int kFzStageWords = 4400;       // resampler input staging (largest instance: 4116 + skew)

struct FzKind {                           // per channel kind (PCG / ECG): despike on/off + its filter
  int despike;
  int pad_;
  double c[2][5];                         // two sections, b0 b1 b2 a1 a2
  double wt[kFzLmax][4];                  // A^(L-1-j) B
  double mp[10][16];                      // M^(2^d), d = 0..9, M = A^L  (d >= 5 move whole warps)
  double prop_slice[16];                  // A^S: state across one full slice
  double prop_part[16];                   // A^nq: state across the valid part of the last chunk of a full slice
  double mlane[32][16];                   // M^lane, lane = 0..31
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
  int start, win, hop, n;                 // window geometry
  long long so_b, so_c, so_k, so_j;       // output strides (elements): recording, channel, window, sample
  unsigned char kind_of_channel[8];
  const FzKind* kinds;                    // DEVICE pointer to the (up to two) channel recipes of this call
};

struct FzFilterScratch {                  // this row's recipe, copied from global memory once per CTA
  double mtab[16][32];                    // M^lane, element-major so a warp's loads are conflict-free
  double wt[kFzLmax][4];                  // pass-1 weights
  double mp[10][16];                      // M^(2^d)
  double prop_slice[16];
  double prop_part[16];
  double c[2][5];
  double wagg[kFzWarps][4];               // warp aggregates
  double wcar[kFzWarps][4];               // state at the start of each warp's first chunk (zero slice start)
};
constexpr int kFzMaxBlocks = 1344;        // 32-sample blocks in one slice (kFzLmax * kFzChunks / 32 + frames)
struct FzDespikeScratch {
  SpikeSorted sorted;
  float bmax[kFzMaxBlocks];
};
struct FzShared {
  union {                                 // three phases, one after the other, share this space:
    float xs[kFzStageWords];              //   resampler input staging
    FzDespikeScratch d;                   //   despike: block maxima of my frames + sorted frame maxima
    FzFilterScratch f;                    //   filter recipe tables and scan scratch
  };
  double xE[kFzMaxCluster][4];            // end states exported by each rank
  double xstat[kFzMaxCluster][4];         // (sum, min, max, -) exported by each rank
  double wstat[kFzWarps][4];
  float tops[2][kFzMaxFrames];            // double-buffered frame maxima (see the despike loop)
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

template <int UP, int DOWN, int D, int FR, int PS>
__global__ void __launch_bounds__(kFzThreads, MPCG_FZ_MINBLOCKS)
fused_preprocess_kernel(const __grid_constant__ FzParams P) {
  extern __shared__ __align__(16) unsigned char fz_raw[];
  FzShared& sm = *reinterpret_cast<FzShared*>(fz_raw);
  float* sig = reinterpret_cast<float*>(fz_raw + sizeof(FzShared));
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster.block_rank();
  const long long row = blockIdx.x / P.ncl;
  const int ch = (int)(row % P.channels);
  const long long rec = row / P.channels;
  const FzKind* __restrict__ Kg = P.kinds + P.kind_of_channel[ch];
  const int k_despike = Kg->despike;
  const int s0 = rank * P.S;
  int n = (rank == P.ncl - 1) ? (P.t - s0) : P.S;
  if (n < 0) n = 0;
  const int L = P.L;
  int dbg_k = 0;
  auto stamp = [&]() {
    if (P.dbg && tid == 0) P.dbg[(long long)blockIdx.x * 16 + dbg_k] = clock64();
    ++dbg_k;
  };
  stamp();                                                // 0: start

  // ---------------------------------------------------------------- 1. resample my slice into shared memory
  const float* xr = P.x + row * (long long)P.t_in;
  if constexpr (UP == DOWN) {                             // no resampling: plain copy of my slice
    for (int i = tid; i < n; i += kFzThreads) sig[i] = ld_stream(xr + s0 + i);
  } else if (n > 0) {
    using T = RsTile<UP, DOWN, D, FR, PS, kFzThreads>;
    static_assert(T::IN_WORDS <= kFzStageWords, "staging buffer too small for this resampler instance");
    const int f_lo = s0 / UP, f_hi = (s0 + n - 1) / UP;
    float pre[T::NPRE];
    T::fetch(pre, xr, (long long)f_lo * DOWN + P.off, P.t_in);
    for (int fb = f_lo; fb <= f_hi; fb += T::NF) {
      T::commit(sm.xs, pre);
      __syncthreads();
      if (fb + T::NF <= f_hi) T::fetch(pre, xr, (long long)(fb + T::NF) * DOWN + P.off, P.t_in);   // next tile in flight
      const int obase = fb * UP - s0;
      auto sink = [&](int frame, int p, float v) {
        const int o = obase + frame * UP + p;
        if ((unsigned)o < (unsigned)n) sig[o] = v;
      };
      T::compute(sm.xs, sink);
      __syncthreads();
    }
  }
  stamp();                                                // 1: slice resampled
  for (int i = n + tid; i < P.cap; i += kFzThreads) sig[i] = 0.f;      // chunk grid beyond the slice

  // ---------------------------------------------------------------- filter tables (independent of the samples)
  __syncthreads();

  stamp();                                                // 2: tables ready
  // ---------------------------------------------------------------- 2. Schmidt despike (cluster-wide)
  // Every CTA keeps two copies of all frame maxima.  Round k reads copy k&1.  Warp 0 of every CTA sorts the
  // maxima (identical data, identical verdict).  The CTA owning the worst frame then keeps going ON ITS OWN:
  // flatten, update its local maxima, decide again -- pass after pass with no barrier at all, as long as the
  // worst frame stays one of its own (a burst is flattened half-cycle by half-cycle, so it usually does).  When
  // ownership moves (or nothing is left to do) it publishes its frames' maxima and the pass counter into copy
  // (k+1)&1 of every CTA; the other CTAs carried the remaining entries over locally; one cluster barrier per round.
  // All other warps only wait at that barrier.
  int passes = 0;
  if (k_despike && P.nframes > 0) {
    const int gf0 = rank * P.fpc;
    int nloc = P.nframes - gf0;
    nloc = nloc < 0 ? 0 : (nloc > P.fpc ? P.fpc : nloc);
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
                  int* tr = P.trace + (row * P.trace_cap + passes) * 4;
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
  if (tid < 16) { sm.f.prop_slice[tid] = Kg->prop_slice[tid]; sm.f.prop_part[tid] = Kg->prop_part[tid]; }
  if (tid < 10) (&sm.f.c[0][0])[tid] = (&Kg->c[0][0])[tid];
  __syncthreads();
  // ---------------------------------------------------------------- 3. low-pass + high-pass as one 4-state scan
  float* mine = sig + tid * L;
  double p[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
  for (int j = 0; j < L; ++j) {
    const double2 w01 = *reinterpret_cast<const double2*>(&sm.f.wt[j][0]);
    const double2 w23 = *reinterpret_cast<const double2*>(&sm.f.wt[j][2]);
    const double xv = (double)mine[j];
    p[0] = fma(w01.x, xv, p[0]);
    p[1] = fma(w01.y, xv, p[1]);
    p[2] = fma(w23.x, xv, p[2]);
    p[3] = fma(w23.y, xv, p[3]);
  }
  stamp();                                                // 4: pass 1 done
  const bool exporter = (P.ncl > 1) && (rank < P.ncl - 1) && (tid == P.q);
  double pp[4] = {0.0, 0.0, 0.0, 0.0};
  if (exporter) {                                         // zero-state response of the partial last chunk
    const int shift = L - P.nq;
    for (int j = 0; j < P.nq; ++j) {
      const double xv = (double)mine[j];
#pragma unroll
      for (int s = 0; s < 4; ++s) pp[s] = fma(sm.f.wt[j + shift][s], xv, pp[s]);
    }
  }
#pragma unroll
  for (int d = 0; d < 5; ++d) {                           // inclusive scan inside the warp
    double u[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) u[s] = __shfl_up_sync(kFull, p[s], 1 << d);
    if (lane >= (1 << d)) mv4_acc(sm.f.mp[d], u, p);
  }
  if (lane == 31) {
#pragma unroll
    for (int s = 0; s < 4; ++s) sm.f.wagg[warp][s] = p[s];
  }
  __syncthreads();
  if (warp == 0) {                                        // scan the warp aggregates in one warp
    double v[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) v[s] = (lane < kFzWarps) ? sm.f.wagg[lane][s] : 0.0;
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
      if (lane < kFzWarps) sm.f.wcar[lane][s] = lane ? e : 0.0;
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
    for (int s = 0; s < 4; ++s) wc[s] = sm.f.wcar[warp][s];
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
    // true slice start state c_(r+1) = A^S c_r + E_r, then carried to my chunk: M^(32*warp + lane)
    double c[4] = {0.0, 0.0, 0.0, 0.0};
    for (int r = 0; r < rank; ++r) {
      double nx[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) nx[s] = sm.xE[r][s];
      mv4_acc(sm.f.prop_slice, c, nx);
#pragma unroll
      for (int s = 0; s < 4; ++s) c[s] = nx[s];
    }
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      if ((warp >> d) & 1) {
        double nx[4];
        mv4_set(sm.f.mp[5 + d], c, nx);
#pragma unroll
        for (int s = 0; s < 4; ++s) c[s] = nx[s];
      }
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
    lim = lim < 0 ? 0 : (lim > L ? L : lim);
    float sacc = 0.f;
#pragma unroll 4
    for (int j = 0; j < lim; ++j) {
      const float v = fz_round(fz_step(kc, z, (double)mine[j]), fix_nan);
      mine[j] = v;
      sacc += v;
      lmin = fminf(lmin, v);
      lmax = fmaxf(lmax, v);
      if ((j & 7) == 7) { lsum += (double)sacc; sacc = 0.f; }
    }
    lsum += (double)sacc;
  }

  stamp();                                                // 7: pass 2 done
  // ---------------------------------------------------------------- 4. row statistics across the cluster
  lsum = warp_sum(lsum);
  lmin = warp_min(lmin);
  lmax = warp_max(lmax);
  if (lane == 0) { sm.wstat[warp][0] = lsum; sm.wstat[warp][1] = (double)lmin; sm.wstat[warp][2] = (double)lmax; }
  __syncthreads();
  if (tid < P.ncl) {
    double a = 0.0, b = INFINITY, c = -INFINITY;
#pragma unroll
    for (int w = 0; w < kFzWarps; ++w) { a += sm.wstat[w][0]; b = fmin(b, sm.wstat[w][1]); c = fmax(c, sm.wstat[w][2]); }
    double* dst = cluster.map_shared_rank(&sm.xstat[rank][0], tid);
    dst[0] = a; dst[1] = b; dst[2] = c;
  }
  cluster_arrive();
  cluster_wait();
  double tot = 0.0, lo_all = INFINITY, hi_all = -INFINITY;
  for (int r = 0; r < P.ncl; ++r) {
    tot += sm.xstat[r][0];
    lo_all = fmin(lo_all, sm.xstat[r][1]);
    hi_all = fmax(hi_all, sm.xstat[r][2]);
  }
  const double mean = tot / (double)P.t;
  const double peak = fmax(hi_all - mean, mean - lo_all);
  double inv_peak;
  if (P.norm_flags & MPCG_NORM_PEAK_GT0) inv_peak = (peak > 0.0) ? 1.0 / peak : 1.0;
  else inv_peak = 1.0 / fmax(peak, 1e-12);
  // fp32 map: the mean is split hi + lo so (s - hi) - lo carries no cancellation error
  const float mean_hi = (float)mean, mean_lo = (float)(mean - (double)mean_hi), inv_f = (float)inv_peak;

  stamp();                                                // 8: statistics exchanged
  // ---------------------------------------------------------------- 5. normalise + write my share of every window
  float* obase = P.out + rec * P.so_b + ch * P.so_c;
  const int s1 = s0 + n;
  for (int k = 0; k < P.n; ++k) {
    const int w0 = P.start + k * P.hop;
    const int a = w0 > s0 ? w0 : s0;
    const int w1 = w0 + P.win;
    const int b = w1 < s1 ? w1 : s1;
    float* dst = obase + k * P.so_k;
    if (P.so_j == 1) {
      for (int i = a + tid; i < b; i += kFzThreads) {
        const float u = ((sig[i - s0] - mean_hi) - mean_lo) * inv_f;
        st_stream(dst + (i - w0), fminf(fmaxf(u, -1.f), 1.f));
      }
    } else {
      for (int i = a + tid; i < b; i += kFzThreads) {
        const float u = ((sig[i - s0] - mean_hi) - mean_lo) * inv_f;
        dst[(long long)(i - w0) * P.so_j] = fminf(fmaxf(u, -1.f), 1.f);
      }
    }
    if (rank == P.ncl - 1 && w1 > P.t) {                  // short recording: zero-fill past its end
      const int z0 = (P.t > w0 ? P.t : w0);
      for (int i = z0 + tid; i < w1; i += kFzThreads) dst[(long long)(i - w0) * P.so_j] = 0.f;
    }
  }
  stamp();                                                // 9: windows stored (issued)
  // No trailing cluster barrier: every remote shared-memory access of this kernel happens before the
  // statistics barrier above, so a CTA may retire while its peers are still storing their windows.
}

template <int UP, int DOWN, int D, int FR, int PS>
int fz_launch(const FzParams& P, size_t smem, long long rows, cudaStream_t stream) {
  auto kern = fused_preprocess_kernel<UP, DOWN, D, FR, PS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(rows * P.ncl));
  cfg.blockDim = dim3(kFzThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)P.ncl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, P);
  if (e != cudaSuccess) return (int)e;
  return MPCG_OK;
}


}  // namespace mpcg
