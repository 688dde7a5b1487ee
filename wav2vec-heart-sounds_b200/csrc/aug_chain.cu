// extern "C" entry: mpcg_aug_chain_f32 -- the whole of augment_pcg_batch (augment/torchaug.py:103-111) in ONE kernel:
//   x = N(x);  x = N(m1 ? x + s1 z1 : x);  x = N(m2 ? x (1 + wander) : x);  x = N(m3 ? EQ(x) : x);  x = N(m4 ? x + s4 z4 : x)
// with  N = subtract the row mean, divide by the row's max |.|, clamp  (torchaug.py:24-27),
//       EQ(x) = N( N(cascade(x)) / 50 + N(x) ),  cascade = the five first-order band-pass sections (torchaug.py:88-100).
// The stage-per-kernel path (aug.cu) reads and writes every row once per stage (ten HBM sweeps and more for the EQ);
// here a row is read once and written once.
//
// A row (a 4 s window at 16 kHz is 256 KB) is spread over a thread-block CLUSTER: CTA `rank` owns samples
// [rank*S, rank*S + S), S = 512 * L, two CTAs per SM.  The slice arrives in shared memory as one bulk asynchronous copy
// (TMA) and from there every thread takes its chunk of L CONSECUTIVE samples into REGISTERS (L is a template parameter:
// 9, 17, 25 or 33; odd, so the lanes' strided shared-memory accesses are conflict-free).  All stages are straight-line
// code over that register chunk -- no index arithmetic, no shared-memory traffic between stages -- and every N is the
// thread's (sum, min, max) plus one small exchange over (distributed) shared memory; the normalising map of stage k is
// applied on the fly by stage k + 1 as one FFMA + clamp.  Shared memory is free in between: it stages the Philox
// normals of the noise stages (drawn in aligned groups of four, like aug.cu draws them) and holds the EQ cascade's
// intermediate signal while the EQ input stays in registers.  The cascade runs as the chunked linear-recurrence scan
// of the preprocessing kernel (fp64 state, two sections per 4-state group, recipe as constant-bank operands, slices
// chained through (A^S)^j); rows whose Bernoulli mask is off skip it.  The wandering-volume sinusoids are evaluated
// once per thread with the reference's rounding sequence and advanced sample by sample with the two-multiply
// recurrence  c -= k s; s += k c  (k = 2 sin(theta / 2): an exact sinusoid in exact arithmetic, 33 steps at most).
// Transforms follow aug.cu's rounding sequence; the final slice leaves as one bulk copy shared -> global.
#include <cooperative_groups.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include "aug.cuh"
#include "biquad.cuh"

namespace cg = cooperative_groups;

namespace mpcg {

#ifndef MPCG_AC_P1_WEIGHTS
#define MPCG_AC_P1_WEIGHTS 1            // 1: pass 1 as the weighted sum (a fresh constant per DFMA); 0: as the recurrence
#endif
#ifndef MPCG_AC_FOLD_ONE
#define MPCG_AC_FOLD_ONE 1              // 1: one warp folds the exchanged statistics (second barrier); 0: every warp does
                                        // (the kernel is issue bound: sixteen redundant folds cost more than the barrier --
                                        //  1.245 -> 1.184 ms at configs[2], 0.499 -> 0.481 ms at the bench shape)
#endif
#ifndef MPCG_AC_UNROLL_SMEM
#define MPCG_AC_UNROLL_SMEM 3          // unroll factor of the filter loops that read shared memory (0: full)
#endif
constexpr int kAcUnrollSmem = MPCG_AC_UNROLL_SMEM;
constexpr int kAcThreads = 512;
constexpr int kAcWarps = kAcThreads / 32;
constexpr int kAcLmax = 33;                 // longest chunk (odd)
constexpr int kAcMaxCluster = 8;
constexpr int kAcMaxGroups = 3;             // up to six sections

struct AcGroupK {                           // one 4-state group (two sections) of the EQ cascade: kernel parameter
  double c[2][5];                           // b0 b1 b2 a1 a2 per section
#if MPCG_AC_P1_WEIGHTS
  double wt[kAcLmax][4];                    // A^(L-1-j) B
#endif
  double mp[9][16];                         // M^(2^d), M = A^L (block lower triangular: a cascade)
};

struct AcParams {
  const float* x;
  float* y;
  int t;
  float fs;
  int ncl, S;
  int ahead;                                // the slice of block b + ahead is requested into L2 by block b
  const float *rowp1, *noise1, *mask1;      // stage 1: noise
  const float *rowp2, *mask2;               // stage 2: wandering volume
  const float* mask3;                       // stage 3: EQ
  const float *rowp4, *noise4, *mask4;      // stage 4: noise
  unsigned long long seed1, sid1, seed4, sid4;
  int ngroups, odd_tail;                    // odd_tail: the last group holds ONE section
  int collapse;                             // 1: a stage whose mask is off keeps the pending map (N(N(x)) == N(x) up to rounding)
  const double* mlane;                      // device (the caller's workspace): AcTables
  AcGroupK k[kAcMaxGroups];
};

// The part of the recipe that is indexed by lane, warp or rank lives in the caller's workspace (a big parameter block
// costs every block's launch: 19 KB of parameters made the plain load-normalise-store path 7 % slower than 16 KB).
struct AcTables {
  double mlane[kAcMaxGroups][16][32];       // M^lane, element-major
  double mwarp[kAcMaxGroups][kAcWarps][16]; // M^(32 w)
  double prop_pow[kAcMaxGroups][kAcMaxCluster][16];   // (A^S)^j
};

struct AcStat {
  double sum;
  float lo, hi;
};
struct AcShared {
  double mtab[kAcMaxGroups][16][32];            // M^lane per group, element-major
  double wagg[kAcWarps][4];
  double wcar[kAcWarps][4];
  AcStat xstat[3][kAcMaxCluster][kAcWarps];      // [exchange parity (set 0) | 2 (set 1)][rank][warp]
  double xE[2][kAcMaxCluster][4];               // [group parity][rank]: end states exported by each rank
  unsigned long long load_bar;                  // mbarrier of the bulk load
  unsigned long long pad_;
  float mapv[2][4];                             // (MPCG_AC_FOLD_ONE) the maps of the exchange in flight
  float volc[2][4];                             // wandering volume, per band: k = 2 sin(theta/2), cos(theta/2), sin(theta/2)
  float rowc[8];                                // the row's parameters: 0-5 wandering volume (amp, freq, phase) x 2, 6 / 7 noise scales
};

static_assert(sizeof(AcShared) % 16 == 0, "the sample slice behind AcShared must stay 16-byte aligned");

__device__ __forceinline__ void ac_cluster_sync(int ncl) {
  if (ncl > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
}

// The pending N of the previous stage:  clamp(v * inv + c),  c = -mean * inv split hi + lo.  The FMA cancels exactly
// (one rounding, of the result), so (v * inv + hi) + lo is within an ulp of the float64 map whatever the row's offset;
// once a row has been normalised its mean is ~0 and `lo` is below 1e-7 of the output range: the one-FFMA form `fast`.
struct AcMap {
  float inv, chi, clo;
  __device__ __forceinline__ float fast(float v) const { return fminf(fmaxf(fmaf(v, inv, chi), -1.f), 1.f); }
  __device__ __forceinline__ float exact(float v) const { return fminf(fmaxf(fmaf(v, inv, chi) + clo, -1.f), 1.f); }
  __device__ __forceinline__ bool needs_exact() const { return !(fabsf(chi) <= 1.f); }
};
struct AcAcc {                              // (sum, min, max) of a thread's chunk
  double sum;
  float lo, hi;
  __device__ __forceinline__ void init() { sum = 0.0; lo = INFINITY; hi = -INFINITY; }
};

// (sum, min, max) of the first `valid` of the L register samples: float32 partial sums of every fourth sample, added
// up in float64 (the float64 <-> float32 conversion unit runs at a quarter of the FMA rate).  A chunk whose offset
// dwarfs its swing (raw rows that were never centred: |v| > 4 (max - min)) would lose the swing in those float32
// partial sums -- the row mean came out 5e-5 of the range off for values around -300 +- 0.003 -- so such a chunk is
// summed again as  n * min + sum (v - min):  the differences are exact (Sterbenz) and small.
template <int L, typename V>
__device__ __forceinline__ void ac_acc_chunk(const V& v, int valid, AcAcc& a) {
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  float lo = INFINITY, hi = -INFINITY;
  if (valid == L) {
#pragma unroll
    for (int j = 0; j < L; ++j) {
      s[j & 3] += v[j];
      lo = fminf(lo, v[j]);
      hi = fmaxf(hi, v[j]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < L; ++j) {
      if (j < valid) {
        s[j & 3] += v[j];
        lo = fminf(lo, v[j]);
        hi = fmaxf(hi, v[j]);
      }
    }
  }
  double sum = ((double)s[0] + (double)s[1]) + ((double)s[2] + (double)s[3]);
  if (valid > 0 && fmaxf(fabsf(lo), fabsf(hi)) > 4.f * (hi - lo)) {
    float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < L; ++j)
      if (j < valid) d[j & 3] += v[j] - lo;
    sum = (double)valid * (double)lo + (((double)d[0] + (double)d[1]) + ((double)d[2] + (double)d[3]));
  }
  a.lo = fminf(a.lo, lo);
  a.hi = fmaxf(a.hi, hi);
  a.sum += sum;
}

// min / max of a float over the warp as ONE integer reduction (REDUX) on the order-preserving image of the bits
// (five shuffle + compare rounds otherwise; the statistics never hold NaNs that matter: a NaN row is NaN either way).
__device__ __forceinline__ int ac_ord(float f) {
  const int b = __float_as_int(f);
  return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ac_warp_min(float v) {
  const int o = __reduce_min_sync(kFull, ac_ord(v));
  return __int_as_float(o ^ ((o >> 31) & 0x7fffffff));
}
__device__ __forceinline__ float ac_warp_max(float v) {
  const int o = __reduce_max_sync(kFull, ac_ord(v));
  return __int_as_float(o ^ ((o >> 31) & 0x7fffffff));
}

// Exchange NSETS row statistics across the cluster; every thread receives the maps.  Each warp leaves its partials in
// every rank's table; after the (cluster) barrier warp 0 folds the table and the maps reach the others through shared
// memory behind a CTA barrier.
template <int NSETS>
__device__ __forceinline__ void ac_exchange(AcShared& sm, cg::cluster_group& cluster, int ncl, int rank, double inv_t, int& parity,
                                            AcAcc (&st)[NSETS], AcMap (&out)[NSETS]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < NSETS; ++s) {
    const double sum = warp_sum(st[s].sum);
    const float lo = ac_warp_min(st[s].lo), hi = ac_warp_max(st[s].hi);
    if (lane < ncl) {
      AcStat* slot = &sm.xstat[s == 0 ? parity : 2][rank][warp];
      AcStat* dst = ncl > 1 ? cluster.map_shared_rank(slot, lane) : slot;
      dst->sum = sum; dst->lo = lo; dst->hi = hi;
    }
  }
  ac_cluster_sync(ncl);
#if MPCG_AC_FOLD_ONE
  if (warp == 0) {                                          // one warp folds; the others wait at a second (CTA) barrier
#endif
#pragma unroll
  for (int s = 0; s < NSETS; ++s) {
    const AcStat* all = &sm.xstat[s == 0 ? parity : 2][0][0];
    double tot = 0.0;
    float lo = INFINITY, hi = -INFINITY;
    for (int e = lane; e < ncl * kAcWarps; e += 32) {
      tot += all[e].sum;
      lo = fminf(lo, all[e].lo);
      hi = fmaxf(hi, all[e].hi);
    }
    tot = warp_sum(tot);
    lo = ac_warp_min(lo);
    hi = ac_warp_max(hi);
    const double mean = tot * inv_t;
    const double peak = fmax((double)hi - mean, mean - (double)lo);
    const float inv = __frcp_rn((float)fmax(peak, 1e-12));
    const double c = -mean * (double)inv;                   // consistent with the ROUNDED scale: the FMA then cancels exactly
#if MPCG_AC_FOLD_ONE
    if (lane == 0) { sm.mapv[s][0] = inv; sm.mapv[s][1] = (float)c; sm.mapv[s][2] = (float)(c - (double)(float)c); }
  }
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < NSETS; ++s) {
    out[s].inv = sm.mapv[s][0];
    out[s].chi = sm.mapv[s][1];
    out[s].clo = sm.mapv[s][2];
  }
#else
    out[s].inv = inv;
    out[s].chi = (float)c;
    out[s].clo = (float)(c - (double)out[s].chi);
  }
#endif
  parity ^= 1;
}

// x / 50 correctly rounded (the reference divides): one Newton step on the product with the rounded reciprocal.
__device__ __forceinline__ float ac_div50(float x) {
  const float q = x * 0.02f;
  return fmaf(fmaf(-50.f, q, x), 0.02f, q);
}

// One group G of the cascade over the slice: NS = 2 sections (four states) or, for the odd section at the end of the
// cascade, NS = 1 (two states: half the work; its transition matrices are the upper-left 2x2 blocks of the group's
// tables, because the missing second section is the identity).  Input: the register chunk (G == 0) or the thread's
// chunk of `mine` (shared memory, G > 0); output: `mine`.  Chunks beyond the row's end see their own finite leftovers,
// whose outputs nobody reads (the filter is causal).  All recurrence state is float64: a float32 first pass was tried
// on the host (tools/eq_precision_study.py: 1e-4 .. 6e-4 of the coloured signal's peak on tonal rows with narrow low
// bands, against 5e-8) and dropped.
//   pass 1  the chunk's end state from a zero start, p = sum_j A^(L-1-j) B x_j, the weights as constant-bank operands.
//           (In isolation this form is bound by the constant loads, a third of the DFMA rate, and the recurrence
//           itself -- 10 DFMA per sample, ten resident constants -- is faster: tools/ubench_iir.cu.  Inside the kernel
//           the weighted sum wins, 1.26 against 1.34 ms at configs[2]: what it leaves of the fp64 pipe serves the
//           co-resident CTA.  -DMPCG_AC_P1_WEIGHTS=0 builds the other form.)
//   scan    Hillis-Steele inside the warp with M^(2^d), warp 0 chains the warp aggregates, M^lane fixes up each lane
//   pass 2  the chunk again from its true start state, transposed direct form II.
template <int NS>
__device__ __forceinline__ void ac_mv(const double* __restrict__ m, const double (&v)[2 * NS], double (&acc)[2 * NS]) {
#pragma unroll
  for (int r = 0; r < 2 * NS; ++r) {
    double a = acc[r];
#pragma unroll
    for (int c = 0; c < (r < 2 ? 2 : 4); ++c) a = fma(m[r * 4 + c], v[c], a);
    acc[r] = a;
  }
}
template <int NS>
__device__ __forceinline__ void ac_mv_lane(const double (*tab)[32], int lane, const double (&v)[2 * NS], double (&acc)[2 * NS]) {
#pragma unroll
  for (int r = 0; r < 2 * NS; ++r) {
    double a = acc[r];
#pragma unroll
    for (int c = 0; c < (r < 2 ? 2 : 4); ++c) a = fma(tab[r * 4 + c][lane], v[c], a);
    acc[r] = a;
  }
}
// one sample through the group's sections (transposed direct form II); returns the last section's output
template <int NS>
__device__ __forceinline__ double ac_step(const AcGroupK& K, double xv, double (&z)[2 * NS]) {
  const double y0 = fma(K.c[0][0], xv, z[0]);
  z[0] = fma(-K.c[0][3], y0, fma(K.c[0][1], xv, z[1]));
  z[1] = fma(-K.c[0][4], y0, K.c[0][2] * xv);
  if (NS == 1) return y0;
  const double y1 = fma(K.c[1][0], y0, z[2 * NS - 2]);
  z[2 * NS - 2] = fma(-K.c[1][3], y1, fma(K.c[1][1], y0, z[2 * NS - 1]));
  z[2 * NS - 1] = fma(-K.c[1][4], y1, K.c[1][2] * y0);
  return y1;
}

template <int L, int G, int NS>
__device__ __forceinline__ void ac_filter_group(const AcParams& P, AcShared& sm, cg::cluster_group& cluster, float* mine,
                                                const float (&v)[L], int ncl, int rank) {
  constexpr int NST = 2 * NS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const AcGroupK& K = P.k[G];
  double p[NST];
#pragma unroll
  for (int s = 0; s < NST; ++s) p[s] = 0.0;
#if MPCG_AC_P1_WEIGHTS
#pragma unroll
  for (int j = 0; j < L; ++j) {
    const double xv = (double)(G == 0 ? v[j] : mine[j]);
#pragma unroll
    for (int s = 0; s < NST; ++s) p[s] = fma(K.wt[j][s], xv, p[s]);
  }
#else
  if (G == 0 || MPCG_AC_UNROLL_SMEM == 0) {               // straight-line code
#pragma unroll
    for (int j = 0; j < L; ++j) ac_step<NS>(K, (double)(G == 0 ? v[j] : mine[j]), p);
  } else {                                                // from shared memory: a short loop body keeps the code small
#pragma unroll(kAcUnrollSmem > 0 ? kAcUnrollSmem : 1)
    for (int j = 0; j < L; ++j) ac_step<NS>(K, (double)mine[j], p);
  }
#endif
  {
    double u[NST];
#define MPCG_AC_LEVEL(DD)                                                                       \
    _Pragma("unroll") for (int s = 0; s < NST; ++s) u[s] = __shfl_up_sync(kFull, p[s], 1 << DD); \
    if (lane >= (1 << DD)) ac_mv<NS>(K.mp[DD], u, p);
    MPCG_AC_LEVEL(0) MPCG_AC_LEVEL(1) MPCG_AC_LEVEL(2) MPCG_AC_LEVEL(3) MPCG_AC_LEVEL(4)
#undef MPCG_AC_LEVEL
  }
  if (lane == 31) {
#pragma unroll
    for (int s = 0; s < NST; ++s) sm.wagg[warp][s] = p[s];
  }
  __syncthreads();
  if (warp == 0) {                                        // scan the sixteen warp aggregates
    double a[NST], u[NST];
#pragma unroll
    for (int s = 0; s < NST; ++s) a[s] = (lane < kAcWarps) ? sm.wagg[lane][s] : 0.0;
#define MPCG_AC_LEVEL(DD)                                                                       \
    _Pragma("unroll") for (int s = 0; s < NST; ++s) u[s] = __shfl_up_sync(kFull, a[s], 1 << DD); \
    if (lane >= (1 << DD)) ac_mv<NS>(K.mp[5 + DD], u, a);
    MPCG_AC_LEVEL(0) MPCG_AC_LEVEL(1) MPCG_AC_LEVEL(2) MPCG_AC_LEVEL(3)
#undef MPCG_AC_LEVEL
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      const double e = __shfl_up_sync(kFull, a[s], 1);
      if (lane < kAcWarps) sm.wcar[lane][s] = lane ? e : 0.0;
    }
    if (ncl > 1 && lane == kAcWarps - 1) {                 // the slice's end state for a zero slice start, to every rank
      for (int rk = 0; rk < ncl; ++rk) {
        double* dst = cluster.map_shared_rank(&sm.xE[G & 1][rank][0], rk);
#pragma unroll
        for (int s = 0; s < NST; ++s) dst[s] = a[s];
      }
    }
  }
  __syncthreads();
  double z[NST];
#pragma unroll
  for (int s = 0; s < NST; ++s) {
    const double e = __shfl_up_sync(kFull, p[s], 1);
    z[s] = lane ? e : 0.0;
  }
  {
    double wc[NST];
#pragma unroll
    for (int s = 0; s < NST; ++s) wc[s] = sm.wcar[warp][s];
    ac_mv_lane<NS>(sm.mtab[G], lane, wc, z);                // chunk start state for a zero slice start
  }
  if (ncl > 1) {
    ac_cluster_sync(ncl);
    double c0[NST];
#pragma unroll
    for (int s = 0; s < NST; ++s) c0[s] = 0.0;
    for (int r = 0; r < rank; ++r) {
      double e0[NST];
#pragma unroll
      for (int s = 0; s < NST; ++s) e0[s] = sm.xE[G & 1][r][s];
      double m[16];
      const double* src = reinterpret_cast<const AcTables*>(P.mlane)->prop_pow[G][rank - 1 - r];
#pragma unroll
      for (int i = 0; i < 4 * NST; ++i) m[i] = __ldg(src + i);
      ac_mv<NS>(m, e0, c0);
    }
    double c[NST];
#pragma unroll
    for (int s = 0; s < NST; ++s) c[s] = 0.0;
    {
      double m[16];
      const double* src = reinterpret_cast<const AcTables*>(P.mlane)->mwarp[G][warp];
#pragma unroll
      for (int i = 0; i < 4 * NST; ++i) m[i] = __ldg(src + i);
      ac_mv<NS>(m, c0, c);
    }
    ac_mv_lane<NS>(sm.mtab[G], lane, c, z);
  }
  if (G == 0 || MPCG_AC_UNROLL_SMEM == 0) {
#pragma unroll
    for (int j = 0; j < L; ++j) mine[j] = (float)ac_step<NS>(K, (double)(G == 0 ? v[j] : mine[j]), z);
  } else {
#pragma unroll(kAcUnrollSmem > 0 ? kAcUnrollSmem : 1)
    for (int j = 0; j < L; ++j) mine[j] = (float)ac_step<NS>(K, (double)mine[j], z);
  }
}

template <int L>
__global__ void __launch_bounds__(kAcThreads, 2)
aug_chain_kernel(const __grid_constant__ AcParams P) {
  extern __shared__ __align__(16) unsigned char ac_raw[];
  AcShared& sm = *reinterpret_cast<AcShared*>(ac_raw);
  float* buf = reinterpret_cast<float*>(ac_raw + sizeof(AcShared));
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x;
  const int ncl = P.ncl;
  const int rank = ncl > 1 ? (int)cluster.block_rank() : 0;
  const unsigned row = blockIdx.x / (unsigned)ncl;
  const int s0 = rank * P.S;
  int n = P.t - s0;
  n = n < 0 ? 0 : (n > P.S ? P.S : n);
  const float* xr = P.x + (long long)row * P.t + s0;
  float* yr = P.y + (long long)row * P.t + s0;
  float* mine = buf + tid * L;
  int valid = n - tid * L;
  valid = valid < 0 ? 0 : (valid > L ? L : valid);
  const bool aligned = ((reinterpret_cast<uintptr_t>(xr) | reinterpret_cast<uintptr_t>(yr)) & 15u) == 0 && n > 0 && (n & 3) == 0;
  // everything the stages need from global memory is requested here, before the slice is waited for; the per-row
  // parameters go to shared memory (sm.rowc: 0-5 wandering volume, 6 / 7 the noise scales)
  const float mk1 = P.mask1 ? P.mask1[row] : 1.f, mk2 = P.mask2 ? P.mask2[row] : 1.f;
  const float mk3 = P.mask3 ? P.mask3[row] : 1.f, mk4 = P.mask4 ? P.mask4[row] : 1.f;
  if (tid < 8) sm.rowc[tid] = tid < 6 ? P.rowp2[(long long)row * 8 + tid] : (tid == 6 ? P.rowp1 : P.rowp4)[(long long)row * 8];
  const bool vol_on = mk2 != 0.f;
  const bool eq_on = P.ngroups > 0 && mk3 != 0.f;
  const unsigned on_bits = (mk1 != 0.f ? 1u : 0u) | (vol_on ? 2u : 0u) | (eq_on ? 4u : 0u) | (mk4 != 0.f ? 16u : 0u);
  int parity = 0;

  // ---- the slice: ONE bulk asynchronous copy (TMA, cp.async.bulk global -> shared, completion counted on an mbarrier)
  //      when it is 16-byte aligned, plain loads otherwise.  While it flies: zero the slice's tail, fetch the M^lane
  //      tables of an EQ row, derive the per-row constants of the wandering volume.
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&sm.load_bar);
  if (aligned && tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = (uint32_t)n * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(buf)),
                 "l"(xr), "r"(bytes), "r"(bar)
                 : "memory");
  }
  if (tid == 32 && P.ahead > 0 && blockIdx.x + (unsigned)P.ahead < gridDim.x) {
    // the slice of the block that will take this block's place: ask L2 for it now, so that its bulk copy finds it there
    const unsigned nb = blockIdx.x + (unsigned)P.ahead;
    const int sb = (int)(nb % (unsigned)ncl) * P.S;
    int n2 = P.t - sb;
    n2 = n2 > P.S ? P.S : n2;
    const float* src = P.x + (long long)(nb / (unsigned)ncl) * P.t + sb;
    if (n2 > 0 && (n2 & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)n2 * 4u) : "memory");
  }
  for (int i = n + tid; i < P.S; i += kAcThreads) buf[i] = 0.f;
  if (!aligned)
    for (int i = tid; i < n; i += kAcThreads) buf[i] = ld_stream(xr + i);
  if (eq_on) {
    const double* src = P.mlane;
    double* dst = &sm.mtab[0][0][0];
    for (int i = tid; i < P.ngroups * 512; i += kAcThreads) dst[i] = src[i];
  }
  if (vol_on && tid < 2) {
    const float f = P.rowp2[(long long)row * 8 + 3 * tid + 1];
    const double half = 3.141592653589793 * (double)f / (double)P.fs;      // theta / 2
    double sh, ch;
    sincos(half, &sh, &ch);
    sm.volc[tid][0] = (float)(2.0 * sh);
    sm.volc[tid][1] = (float)ch;
    sm.volc[tid][2] = (float)sh;
  }
  __syncthreads();                                                       // the barrier is initialised, the tail zeroed
  if (aligned) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "AC_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra AC_DONE;\n\t"
        "bra AC_WAIT;\n\t"
        "AC_DONE:\n\t"
        "}\n" ::"r"(bar)
        : "memory");
  }
  float v[L];
#pragma unroll
  for (int j = 0; j < L; ++j) v[j] = mine[j];

  const double inv_t = 1.0 / (double)P.t;
  AcAcc st[1];
  AcMap map[1];
  st[0].init();
  ac_acc_chunk<L>(v, valid, st[0]);
  ac_exchange<1>(sm, cluster, ncl, rank, inv_t, parity, st, map);
  // A raw row whose offset dwarfs its swing needs the two-term map: apply it now, once.  From here on every pending
  // map belongs to a row that was centred one stage ago (|mean| / peak <= ~4 at the very worst: the one-FFMA form is
  // within 1.2e-7 of the output range, typically 1e-9).
  if (map[0].needs_exact()) {
    const AcMap m = map[0];
#pragma unroll
    for (int j = 0; j < L; ++j) v[j] = m.exact(v[j]);
    map[0].inv = 1.f; map[0].chi = 0.f; map[0].clo = 0.f;
  }

  // ---- the stages, one after the other: 0 noise, 1 wandering volume, 2 EQ, 3 the N that follows the EQ's own N, 4 noise.
  //      Every pass ends with the statistics of the chunk, the exchange and the next pending map.
#pragma unroll 1
  for (int stage = 0; stage < 5; ++stage) {
    bool on = (on_bits >> stage) & 1u;
    if (stage == 3) {
      if (!eq_on) continue;                                   // no EQ: nothing of it to re-normalise
      on = false;
    }
    const AcMap m = map[0];
    if (!on) {
      // A stage that leaves the row alone would only normalise an already normalised row again: mean 0 and peak 1 up
      // to float32 rounding, i.e. the identity to ~1e-7.  With P.collapse the pending map simply stays pending (unless
      // the row is degenerate: a peak below 1e-6 means the previous map amplified rounding residue).
      if (P.collapse && m.inv < 1e6f) continue;
#pragma unroll
      for (int j = 0; j < L; ++j) v[j] = m.fast(v[j]);
    } else if (stage == 1) {
      // ---- wandering volume
      const float two_pi = 6.283185307179586f;
      const float tt0 = __fdiv_rn((float)(s0 + tid * L), P.fs);
      const float a0 = sm.rowc[0], a1 = sm.rowc[3];
      float sn[2], cs[2], kk[2];
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        float s_, c_;
        sincosf(__fmul_rn(two_pi, __fadd_rn(__fmul_rn(sm.rowc[3 * b + 1], tt0), sm.rowc[3 * b + 2])), &s_, &c_);
        kk[b] = sm.volc[b][0];
        sn[b] = s_;
        cs[b] = fmaf(c_, sm.volc[b][1], s_ * sm.volc[b][2]);     // cos(phase - theta/2)
      }
#pragma unroll
      for (int j = 0; j < L; ++j) {
        const float mod = __fadd_rn(__fmul_rn(a0, sn[0]), __fmul_rn(a1, sn[1]));
        v[j] = __fmul_rn(m.fast(v[j]), __fadd_rn(1.f, mod));
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          cs[b] = fmaf(-kk[b], sn[b], cs[b]);
          sn[b] = fmaf(kk[b], cs[b], sn[b]);
        }
      }
    } else if (stage == 2) {
      // ---- EQ: v <- N(cascade(x2)) / 50 + N(x2), x2 = N(previous) kept in registers, the cascade's signal in `mine`
      AcAcc s2[2];
      AcMap m2[2];
      s2[0].init(); s2[1].init();
#pragma unroll
      for (int j = 0; j < L; ++j) v[j] = m.fast(v[j]);
      ac_acc_chunk<L>(v, valid, s2[1]);
      // (the odd section at the end of the cascade is a group of its own with half the work)
#define MPCG_AC_GROUP(GG)                                                                                 \
      if (P.ngroups > GG) {                                                                              \
        if (P.ngroups == GG + 1 && P.odd_tail) ac_filter_group<L, GG, 1>(P, sm, cluster, mine, v, ncl, rank); \
        else ac_filter_group<L, GG, 2>(P, sm, cluster, mine, v, ncl, rank);                              \
      }
      MPCG_AC_GROUP(0) MPCG_AC_GROUP(1) MPCG_AC_GROUP(2)
#undef MPCG_AC_GROUP
      ac_acc_chunk<L>(mine, valid, s2[0]);                                // statistics of the coloured signal
      ac_exchange<2>(sm, cluster, ncl, rank, inv_t, parity, s2, m2);      // N(c) and N(x2)
#pragma unroll
      for (int j = 0; j < L; ++j) v[j] = __fadd_rn(ac_div50(m2[0].exact(mine[j])), m2[1].exact(v[j]));
    } else {
      // ---- noise: the normals of the slice go through shared memory, drawn (or loaded) in aligned groups of four and
      //      read back chunk-wise (whoever read the buffer last did so before the barrier of the exchange that followed)
      const float* noise = stage == 0 ? P.noise1 : P.noise4;
      if (noise == nullptr) {
        const unsigned long long seed = stage == 0 ? P.seed1 : P.seed4, sid = stage == 0 ? P.sid1 : P.sid4;
        float4* buf4 = reinterpret_cast<float4*>(buf);
        const int nq = (n + 3) >> 2;
        for (int q = tid; q < nq; q += kAcThreads) buf4[q] = philox_normal4(seed, sid, row, (long long)((s0 >> 2) + q));
      } else {
        const float* nz = noise + (long long)row * P.t + s0;
        for (int i = tid; i < n; i += kAcThreads) buf[i] = ld_stream(nz + i);
      }
      __syncthreads();
      const float scale = sm.rowc[stage == 0 ? 6 : 7];
#pragma unroll
      for (int j = 0; j < L; ++j) v[j] = __fadd_rn(m.fast(v[j]), __fmul_rn(scale, mine[j]));      // x + (scale*std) * noise
    }
    st[0].init();
    ac_acc_chunk<L>(v, valid, st[0]);
    ac_exchange<1>(sm, cluster, ncl, rank, inv_t, parity, st, map);
  }

  // ---- final map + store: the slice goes back through shared memory and leaves as one bulk copy
  {
    const AcMap m = map[0];
#pragma unroll
    for (int j = 0; j < L; ++j) mine[j] = m.fast(v[j]);
    if (aligned) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(yr),
                     "r"((uint32_t)__cvta_generic_to_shared(buf)), "r"((uint32_t)n * 4u)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    } else {
      __syncthreads();
      for (int i = tid; i < n; i += kAcThreads) st_stream(yr + i, buf[i]);
    }
  }
}

template <int L>
static int ac_launch(const AcParams& P, long long rows, cudaStream_t stream) {
  const size_t smem = sizeof(AcShared) + (size_t)P.S * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(aug_chain_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(rows * P.ncl));
  cfg.blockDim = dim3(kAcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)P.ncl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, aug_chain_kernel<L>, P);
  if (e != cudaSuccess) return (int)e;
  return MPCG_OK;
}

}  // namespace mpcg

// The lane-indexed part of the EQ recipe (M^lane per group) travels through the caller's workspace (the ABI never
// allocates and keeps no device state between calls): uploaded on the caller's stream right before the launch, so calls
// on different streams or from different host threads cannot disturb each other as long as each brings its own
// workspace.  Everything else of the recipe is a kernel parameter.
extern "C" int64_t mpcg_aug_chain_work_bytes(void) { return (int64_t)sizeof(mpcg::AcTables); }

extern "C" int mpcg_aug_chain_f32(const float* x, float* y, int64_t rows, int64_t t, float fs, const float* rowp1,
                                  const float* noise1, const float* mask1, uint64_t seed1, uint64_t sid1,
                                  const float* rowp2, const float* mask2, const double* eq_sos, int eq_sections,
                                  const float* mask3, const float* rowp4, const float* noise4, const float* mask4,
                                  uint64_t seed4, uint64_t sid4, int flags, void* work, int64_t work_bytes, void* stream_) {
  using namespace mpcg;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rows < 0 || t < 0 || !(fs > 0.f) || eq_sections < 0) return MPCG_EINVAL;
  if (eq_sections > 2 * kAcMaxGroups) return MPCG_EUNSUPPORTED;
  if (rows == 0 || t == 0) return MPCG_OK;
  if (!x || !y || !rowp1 || !rowp2 || !rowp4 || (eq_sections > 0 && !eq_sos)) return MPCG_EINVAL;
  if (x == y) return MPCG_EINVAL;
  if (t > 0x7fffffffLL) return MPCG_ERANGE;
  // geometry: slices of S = threads * L samples; the smallest power-of-two cluster whose chunk length fits, and the
  // smallest compiled chunk length (9 / 17 / 25 / 33) that covers it.  MPCG_AC_CLUSTER overrides (experiments).
  int ncl = 0, L = 0;
  {
    const char* env = getenv("MPCG_AC_CLUSTER");
    const int forced = env ? atoi(env) : 0;
    for (int c = 1; c <= kAcMaxCluster && ncl == 0; c *= 2) {
      if (forced > 0 && c != forced) continue;
      const long long per = (t + c - 1) / c;
      const long long l = (per + kAcThreads - 1) / kAcThreads;
      for (int cand = 9; cand <= kAcLmax; cand += 8)
        if (l <= cand) { ncl = c; L = cand; break; }
    }
  }
  if (ncl == 0) return MPCG_EUNSUPPORTED;
  const int S = kAcThreads * L;
  if (rows * ncl > 0x7fffffffLL) return MPCG_ERANGE;
  const int ngroups = (eq_sections + 1) / 2;

  AcParams P;
  memset(&P, 0, sizeof(P));
  static thread_local AcTables tab;                            // (21 KB: off the stack)
  for (int g = 0; g < ngroups; ++g) {
    AcGroupK& k = P.k[g];
    bool ok;
    bq_group_coeffs(eq_sos, eq_sections, 2 * g, k.c, &ok);
    if (!ok) return MPCG_EINVAL;
    double A[16], B[4];
    bq_group_AB(k.c, A, B);
#if MPCG_AC_P1_WEIGHTS
    {
      double v[4] = {B[0], B[1], B[2], B[3]};
      for (int j = L - 1; j >= 0; --j) {
        for (int s = 0; s < 4; ++s) k.wt[j][s] = v[s];
        bq_group_step(k.c, v, 0.0);
      }
    }
#endif
    bq_mat_pow(A, L, k.mp[0]);
    for (int d = 1; d < 9; ++d) bq_mat_mul(k.mp[d - 1], k.mp[d - 1], k.mp[d]);
    for (int l = 0; l < 32; ++l) {
      double m[16];
      bq_mat_pow(k.mp[0], l, m);
      for (int i = 0; i < 16; ++i) tab.mlane[g][i][l] = m[i];
    }
    for (int w = 0; w < kAcWarps; ++w) bq_mat_pow(k.mp[0], 32LL * w, tab.mwarp[g][w]);
    for (int j = 0; j < kAcMaxCluster; ++j) bq_mat_pow(A, (long long)S * j, tab.prop_pow[g][j]);
  }
  if (ngroups > 0) {
    if (!work || ((uintptr_t)work & 15u) || work_bytes < (int64_t)sizeof(AcTables)) return MPCG_EINVAL;
    // (pageable source: the runtime stages the bytes before returning, so the next call on this thread may refill `tab`)
    cudaError_t e = cudaMemcpyAsync(work, &tab, sizeof(AcTables), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return (int)e;
    P.mlane = reinterpret_cast<const double*>(work);
  }

  P.x = x; P.y = y; P.t = (int)t; P.fs = fs; P.ncl = ncl; P.S = S;
  {
    static int sms = 0;                                       // two blocks per SM are resident; half a wave ahead measured best
                                                              // (0 / 148 / 296 / 592 blocks: 1.257 / 1.235 / 1.239 / 1.269 ms at configs[2])
    if (sms == 0) {
      int dev = 0, v = 0;
      if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) v = 148;
      sms = v;
    }
    const char* env = getenv("MPCG_AC_AHEAD");
    P.ahead = env ? atoi(env) : sms / ncl * ncl;
  }
  P.rowp1 = rowp1; P.noise1 = noise1; P.mask1 = mask1; P.seed1 = seed1; P.sid1 = sid1;
  P.rowp2 = rowp2; P.mask2 = mask2; P.mask3 = mask3;
  P.rowp4 = rowp4; P.noise4 = noise4; P.mask4 = mask4; P.seed4 = seed4; P.sid4 = sid4;
  P.ngroups = ngroups; P.odd_tail = eq_sections & 1; P.collapse = (flags & MPCG_AUG_CHAIN_COLLAPSE) ? 1 : 0;
  switch (L) {
    case 9: return ac_launch<9>(P, rows, stream);
    case 17: return ac_launch<17>(P, rows, stream);
    case 25: return ac_launch<25>(P, rows, stream);
    default: return ac_launch<33>(P, rows, stream);
  }
}
