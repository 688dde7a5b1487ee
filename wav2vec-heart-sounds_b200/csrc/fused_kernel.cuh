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

constexpr int kFzThreads = 256;
constexpr int kFzWarps = kFzThreads / 32;
constexpr int kFzLmax = 161;              // longest per-thread chunk (odd)
constexpr int kFzMaxFrames = 1024;        // despike frames per row the fused kernel accepts
constexpr int kFzMaxCluster = 8;
constexpr int kFzStageWords = 4400;       // resampler input staging (largest instance: 4110 + skew)

struct FzKind {                           // per channel kind (PCG / ECG): despike on/off + its filter
  int despike;
  int pad_;
  double c[2][5];                         // two sections, b0 b1 b2 a1 a2
  double wt[kFzLmax][4];                  // A^(L-1-j) B
  double mp[8][16];                       // M^(2^d), d = 0..7, M = A^L  (d >= 5 move whole warps)
  double prop_slice[16];                  // A^S: state across one full slice
  double prop_part[16];                   // A^nq: state across the valid part of the last chunk of a full slice
};

struct FzParams {
  const float* x;
  float* out;
  int* edits;
  int* trace;
  int trace_cap;
  int channels;                           // rows per recording
  int t_in, t;                            // samples per row before / after resampling
  int off;                                // resampler input offset
  int identity;                           // 1: no resampling (copy)
  int ncl, S, L, cap;                     // cluster size, slice length, chunk length, L * threads
  int q, nq;                              // chunk holding the last sample of a full slice, valid samples in it
  int win_d, nframes, fpc;                // despike frame length, frames per row, frames per CTA
  double threshold;
  int max_iter, median_mode, norm_flags;
  int start, win, hop, n;                 // window geometry
  long long so_b, so_c, so_k, so_j;       // output strides (elements): recording, channel, window, sample
  unsigned char kind_of_channel[8];
  FzKind kind[2];
};

struct FzShared {
  double mtab[32][16];                    // M^lane
  double wagg[kFzWarps][4];
  double wcar[kFzWarps][4];
  double xE[kFzMaxCluster][4];            // end states exported by each rank
  double xstat[kFzMaxCluster][4];         // (sum, min, max, -) exported by each rank
  double dscr[32];
  float tops[kFzMaxFrames];
  float fscr[40];
  int iscr[32];
  int ctrl[4];
  float xs[kFzStageWords];
};

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int UP, int DOWN, int D, int FR, int PS>
__global__ void __launch_bounds__(kFzThreads, 2)
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
  const FzKind& K = P.kind[P.kind_of_channel[ch]];
  const int s0 = rank * P.S;
  int n = (rank == P.ncl - 1) ? (P.t - s0) : P.S;
  if (n < 0) n = 0;

  // ---------------------------------------------------------------- M^lane table (warp 0), overlaps the loads
  if (tid < 32) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (i % 5 == 0) ? 1.0 : 0.0;
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      if ((tid >> d) & 1) {
        double nxt[16];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) a = fma(K.mp[d][r * 4 + k], acc[k * 4 + c], a);
            nxt[r * 4 + c] = a;
          }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = nxt[i];
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) sm.mtab[tid][i] = acc[i];
  }

  // ---------------------------------------------------------------- 1. resample my slice into shared memory
  const float* xr = P.x + row * (long long)P.t_in;
  if constexpr (UP == DOWN) {                             // no resampling: plain copy of my slice
    for (int i = tid; i < n; i += kFzThreads) sig[i] = ld_stream(xr + s0 + i);
  } else if (n > 0) {
    using T = RsTile<UP, DOWN, D, FR, PS, kFzThreads>;
    static_assert(T::IN_WORDS <= kFzStageWords, "staging buffer too small for this resampler instance");
    const int f_lo = s0 / UP, f_hi = (s0 + n - 1) / UP;
    for (int fb = f_lo; fb <= f_hi; fb += T::NF) {
      T::stage(sm.xs, xr, (long long)fb * DOWN + P.off, P.t_in);
      __syncthreads();
      const int obase = fb * UP - s0;
      auto sink = [&](int frame, int p, float v) {
        const int o = obase + frame * UP + p;
        if (o >= 0 && o < n) sig[o] = v;
      };
      T::compute(sm.xs, sink);
      __syncthreads();
    }
  }
  for (int i = n + tid; i < P.cap; i += kFzThreads) sig[i] = 0.f;      // chunk grid beyond the slice
  __syncthreads();

  // ---------------------------------------------------------------- 2. Schmidt despike (cluster-wide)
  int passes = 0;
  if (K.despike && P.nframes > 0) {
    const int gf0 = rank * P.fpc;
    int nloc = P.nframes - gf0;
    nloc = nloc < 0 ? 0 : (nloc > P.fpc ? P.fpc : nloc);
    for (int f = warp; f < nloc; f += kFzWarps) {
      const float* p = sig + f * P.win_d;
      float m = 0.f;
      for (int i = lane; i < P.win_d; i += 32) m = fmaxf(m, fabsf(p[i]));
      m = warp_max(m);
      if (lane < P.ncl) *cluster.map_shared_rank(&sm.tops[gf0 + f], lane) = m;
    }
    cluster_arrive();
    cluster_wait();
    for (; passes < P.max_iter; ++passes) {
      const SpikeDecision dec =
          spike_decide<kFzThreads>(sm.tops, P.nframes, P.threshold, P.median_mode, sm.fscr, sm.iscr);
      if (!dec.active) break;
      const int owner = dec.worst / P.fpc;
      cluster_arrive();                                   // every CTA has finished reading tops
      int peak = 0, lo = 0, hi = 0;
      bool changed = false;
      float new_top = 0.f;
      if (rank == owner) {
        spike_flatten<kFzThreads>(sig + (dec.worst - gf0) * P.win_d, P.win_d, peak, lo, hi, changed, new_top,
                                  sm.fscr, sm.iscr);
        if (tid == 0 && P.trace && passes < P.trace_cap) {
          int* tr = P.trace + (row * P.trace_cap + passes) * 4;
          tr[0] = dec.worst; tr[1] = peak; tr[2] = lo; tr[3] = hi;
        }
      }
      cluster_wait();
      if (rank == owner && tid < P.ncl) {
        *cluster.map_shared_rank(&sm.tops[dec.worst], tid) = new_top;
        *cluster.map_shared_rank(&sm.ctrl[0], tid) = changed ? 1 : 0;
      }
      cluster_arrive();
      cluster_wait();
      if (!sm.ctrl[0]) { ++passes; break; }               // fixed point: the reference would only repeat it
    }
    if (P.edits && rank == 0 && tid == 0) P.edits[row] = passes;
    __syncthreads();
  } else if (P.edits && rank == 0 && tid == 0) {
    P.edits[row] = 0;
  }

  // ---------------------------------------------------------------- 3. low-pass + high-pass as one 4-state scan
  const int L = P.L;
  float* mine = sig + tid * L;
  double p[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
  for (int j = 0; j < L; ++j) {
    const double xv = (double)mine[j];
#pragma unroll
    for (int s = 0; s < 4; ++s) p[s] = fma(K.wt[j][s], xv, p[s]);
  }
  const bool exporter = (P.ncl > 1) && (rank < P.ncl - 1) && (tid == P.q);
  double pp[4] = {0.0, 0.0, 0.0, 0.0};
  if (exporter) {                                         // zero-state response of the partial last chunk
    const int shift = L - P.nq;
    for (int j = 0; j < P.nq; ++j) {
      const double xv = (double)mine[j];
#pragma unroll
      for (int s = 0; s < 4; ++s) pp[s] = fma(K.wt[j + shift][s], xv, pp[s]);
    }
  }
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    double u[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) u[s] = __shfl_up_sync(kFull, p[s], 1 << d);
    if (lane >= (1 << d)) mv4_acc(K.mp[d], u, p);
  }
  if (lane == 31) {
#pragma unroll
    for (int s = 0; s < 4; ++s) sm.wagg[warp][s] = p[s];
  }
  __syncthreads();
  if (tid == 0) {                                         // chain the warp aggregates from a zero slice start
    double c[4] = {0.0, 0.0, 0.0, 0.0};
    for (int w = 0; w < kFzWarps; ++w) {
      double nx[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) { sm.wcar[w][s] = c[s]; nx[s] = sm.wagg[w][s]; }
      mv4_acc(K.mp[5], c, nx);
#pragma unroll
      for (int s = 0; s < 4; ++s) c[s] = nx[s];
    }
  }
  __syncthreads();
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
    mv4_acc(sm.mtab[lane], wc, z);                       // start state of my chunk for a zero slice start
  }
  if (P.ncl > 1) {
    if (exporter) {
      double e[4] = {pp[0], pp[1], pp[2], pp[3]};
      mv4_acc(K.prop_part, z, e);                         // E = A^nq * start_q + partial response
      for (int rk = 0; rk < P.ncl; ++rk) {
        double* dst = cluster.map_shared_rank(&sm.xE[rank][0], rk);
#pragma unroll
        for (int s = 0; s < 4; ++s) dst[s] = e[s];
      }
    }
    cluster_arrive();
    cluster_wait();
    // true slice start state: c_(r+1) = A^S c_r + E_r ; then move it to my chunk: M^(32*warp + lane)
    double c[4] = {0.0, 0.0, 0.0, 0.0};
    for (int r = 0; r < rank; ++r) {
      double nx[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) nx[s] = sm.xE[r][s];
      mv4_acc(K.prop_slice, c, nx);
#pragma unroll
      for (int s = 0; s < 4; ++s) c[s] = nx[s];
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {                         // whole warps: M^(32 * 2^d)
      if ((warp >> d) & 1) {
        double nx[4] = {0.0, 0.0, 0.0, 0.0};
        mv4_acc(K.mp[5 + d], c, nx);
#pragma unroll
        for (int s = 0; s < 4; ++s) c[s] = nx[s];
      }
    }
    mv4_acc(sm.mtab[lane], c, z);
  }
  // pass 2: transposed direct form II; statistics of the valid outputs ride along
  double lsum = 0.0;
  float lmin = INFINITY, lmax = -INFINITY;
  {
    const double b00 = K.c[0][0], b01 = K.c[0][1], b02 = K.c[0][2], a01 = K.c[0][3], a02 = K.c[0][4];
    const double b10 = K.c[1][0], b11 = K.c[1][1], b12 = K.c[1][2], a11 = K.c[1][3], a12 = K.c[1][4];
    const bool fix_nan = (P.norm_flags & MPCG_NORM_NAN_TO_NUM) != 0;
    const int valid = n - tid * L;                        // samples of my chunk that belong to the row
#pragma unroll 4
    for (int j = 0; j < L; ++j) {
      const double xv = (double)mine[j];
      const double y0 = fma(b00, xv, z[0]);
      z[0] = fma(-a01, y0, fma(b01, xv, z[1]));
      z[1] = fma(-a02, y0, b02 * xv);
      const double y1 = fma(b10, y0, z[2]);
      z[2] = fma(-a11, y1, fma(b11, y0, z[3]));
      z[3] = fma(-a12, y1, b12 * y0);
      float v = (float)y1;
      if (fix_nan) {
        if (v != v) v = 0.f;
        else if (v == INFINITY) v = FLT_MAX;
        else if (v == -INFINITY) v = -FLT_MAX;
      }
      mine[j] = v;
      if (j < valid) {
        lsum += (double)v;
        lmin = fminf(lmin, v);
        lmax = fmaxf(lmax, v);
      }
    }
  }

  // ---------------------------------------------------------------- 4. row statistics across the cluster
  lsum = block_sum<kFzThreads>(lsum, sm.dscr);
  lmin = block_min<kFzThreads>(lmin, sm.fscr);
  lmax = block_max<kFzThreads>(lmax, sm.fscr);
  if (tid < P.ncl) {
    double* dst = cluster.map_shared_rank(&sm.xstat[rank][0], tid);
    dst[0] = lsum; dst[1] = (double)lmin; dst[2] = (double)lmax;
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
        const double u = ((double)sig[i - s0] - mean) * inv_peak;
        st_stream(dst + (i - w0), (float)fmin(fmax(u, -1.0), 1.0));
      }
    } else {
      for (int i = a + tid; i < b; i += kFzThreads) {
        const double u = ((double)sig[i - s0] - mean) * inv_peak;
        dst[(long long)(i - w0) * P.so_j] = (float)fmin(fmax(u, -1.0), 1.0);
      }
    }
    if (rank == P.ncl - 1 && w1 > P.t) {                  // short recording: zero-fill past its end
      const int z0 = (P.t > w0 ? P.t : w0);
      for (int i = z0 + tid; i < w1; i += kFzThreads) dst[(long long)(i - w0) * P.so_j] = 0.f;
    }
  }
  cluster_arrive();                                       // peers may still be reading my shared memory
  cluster_wait();
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
