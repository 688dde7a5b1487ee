// extern "C" entry: mpcg_aug_chain_f32 -- the whole of augment_pcg_batch (augment/torchaug.py:103-111) in ONE kernel:
//   x = N(x);  x = N(m1 ? x + s1 z1 : x);  x = N(m2 ? x (1 + wander) : x);  x = N(m3 ? EQ(x) : x);  x = N(m4 ? x + s4 z4 : x)
// with  N = subtract the row mean, divide by the row's max |.|, clamp  (torchaug.py:24-27),
//       EQ(x) = N( N(cascade(x)) / 50 + N(x) ),  cascade = the five first-order band-pass sections (torchaug.py:88-100).
// The stage-per-kernel path (aug.cu) reads and writes every row once per stage (ten HBM sweeps and more for the EQ);
// here a row is read once and written once.
//
// A row (a 4 s window at 16 kHz is 256 KB) is spread over a thread-block CLUSTER: CTA `rank` keeps samples
// [rank*S, rank*S + S) in shared memory, S = 256 * L, three CTAs per SM so that the clusters of different rows
// overlap their phases.  Every N is a sweep over the resident slice plus one small exchange of (sum, min, max)
// over distributed shared memory; the normalising map of stage k is applied on the fly by the sweep of stage k+1.
// The EQ cascade runs in place as the chunked linear-recurrence scan of the preprocessing kernel (fp64 state, two
// sections per 4-state group), slices chained through (A^S)^j tables; rows whose Bernoulli mask is off skip it.
// While the cascade overwrites the slice, the stage's input is parked in the row's OUTPUT buffer (re-read from L2).
// Transforms follow aug.cu's rounding sequence.  The normalising map runs in float32 with the mean split hi + lo
// ((v - hi) - lo) * inv -- no cancellation error, within 1-2 ulp of aug.cu's float64 map -- and row sums are formed as
// float32 partial sums of four samples added up in float64: the float64 <-> float32 conversion unit (16 lanes per
// clock per SM) would otherwise bound a kernel that sweeps its row six times.
#include <cooperative_groups.h>
#include <string.h>
#include <stdlib.h>
#include "aug.cuh"
#include "biquad.cuh"

namespace cg = cooperative_groups;

namespace mpcg {

constexpr int kAcThreads = 512;
constexpr int kAcWarps = kAcThreads / 32;
constexpr int kAcLmax = 33;                 // longest filter chunk (odd)
constexpr int kAcMaxCluster = 8;
constexpr int kAcMaxGroups = 3;             // up to six sections

struct AcGroup {                            // one 4-state group (two sections) of the EQ cascade
  double c[2][5];                           // b0 b1 b2 a1 a2 per section
  double wt[kAcLmax][4];                    // A^(L-1-j) B
  double mp[9][16];                         // M^(2^d), M = A^L
  double mlane[32][16];                     // M^lane
  double mwarp[kAcWarps][16];               // M^(32 w)
  double prop_pow[kAcMaxCluster][16];       // (A^S)^j
};

struct AcParams {
  const float* x;
  float* y;
  int t;
  float fs;
  int ncl, S, L;
  const float *rowp1, *noise1, *mask1;      // stage 1: noise
  const float *rowp2, *mask2;               // stage 2: wandering volume
  const float* mask3;                       // stage 3: EQ
  const float *rowp4, *noise4, *mask4;      // stage 4: noise
  unsigned long long seed1, sid1, seed4, sid4;
  int ngroups;
  int collapse;                             // 1: a stage whose mask is off keeps the pending map (N(N(x)) == N(x) up to rounding)
  const AcGroup* groups;                    // device
};

struct AcStat {
  double sum;
  float lo, hi;
};
struct AcTables {                           // current group's recipe in shared memory
  double mtab[16][32];                      // M^lane, element-major
  double wt[kAcLmax][4];
  double mp[9][16];
  double mwarp[kAcWarps][16];
  double prop_pow[kAcMaxCluster][16];
  double c[2][5];
  double wagg[kAcWarps][4];
  double wcar[kAcWarps][4];
};
struct AcShared {
  AcTables f;
  AcStat xstat[3][kAcMaxCluster][kAcWarps];      // [exchange parity (set 0) | 2 (set 1)][rank][warp]
  double xE[2][kAcMaxCluster][4];               // [group parity][rank]: end states exported by each rank
  unsigned long long load_bar;                  // mbarrier of the bulk load
  unsigned long long pad_;
  float mapv[2][4];                             // (mean hi, mean lo, 1 / peak) of the exchange in flight
};

__device__ __forceinline__ void ac_cluster_sync(int ncl) {
  if (ncl > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
}

struct AcMap {                              // the pending N of the previous stage
  float mh, ml, inv;
  __device__ __forceinline__ float operator()(float v) const {
    return fminf(fmaxf(((v - mh) - ml) * inv, -1.f), 1.f);
  }
};
struct AcAcc {                              // (sum, min, max) of a thread's share: float32 partial sums -> float64
  double sum;
  float part, lo, hi;
  __device__ __forceinline__ void init() { sum = 0.0; part = 0.f; lo = INFINITY; hi = -INFINITY; }
  __device__ __forceinline__ void add(float v) { part += v; lo = fminf(lo, v); hi = fmaxf(hi, v); }
  __device__ __forceinline__ void flush() { sum += (double)part; part = 0.f; }
};

// Exchange NSETS row statistics across the cluster; every thread receives the maps.
template <int NSETS>
__device__ __forceinline__ void ac_exchange(AcShared& sm, cg::cluster_group& cluster, int ncl, int rank, int t, int& parity,
                                            AcAcc (&st)[NSETS], AcMap (&out)[NSETS]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < NSETS; ++s) {
    st[s].flush();
    const double sum = warp_sum(st[s].sum);
    const float lo = warp_min(st[s].lo), hi = warp_max(st[s].hi);
    if (lane < ncl) {
      AcStat* slot = &sm.xstat[s == 0 ? parity : 2][rank][warp];
      AcStat* dst = ncl > 1 ? cluster.map_shared_rank(slot, lane) : slot;
      dst->sum = sum; dst->lo = lo; dst->hi = hi;
    }
  }
  ac_cluster_sync(ncl);
  if (warp == 0) {                                          // one warp folds the partials and forms the maps
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
      lo = warp_min(lo);
      hi = warp_max(hi);
      const double mean = tot / (double)t;
      const double peak = fmax((double)hi - mean, mean - (double)lo);
      if (lane == 0) {
        const float mh = (float)mean;
        sm.mapv[s][0] = mh;
        sm.mapv[s][1] = (float)(mean - (double)mh);
        sm.mapv[s][2] = (float)(1.0 / fmax(peak, 1e-12));
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < NSETS; ++s) {
    out[s].mh = sm.mapv[s][0];
    out[s].ml = sm.mapv[s][1];
    out[s].inv = sm.mapv[s][2];
  }
  parity ^= 1;
}

__device__ __forceinline__ void ac_mv4_lane_acc(const double (*tab)[32], int lane, const double (&v)[4], double (&acc)[4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    double a = acc[r];
#pragma unroll
    for (int c = 0; c < 4; ++c) a = fma(tab[r * 4 + c][lane], v[c], a);
    acc[r] = a;
  }
}
__device__ __forceinline__ void ac_mv4_set(const double* __restrict__ m, const double (&v)[4], double (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    double a = m[r * 4] * v[0];
#pragma unroll
    for (int c = 1; c < 4; ++c) a = fma(m[r * 4 + c], v[c], a);
    out[r] = a;
  }
}

// One 4-state group of the cascade, in place over the resident slice (all S samples; beyond the row's end the slice
// holds zeros, whose outputs nobody reads).
__device__ __forceinline__ void ac_filter_group(AcShared& sm, cg::cluster_group& cluster, float* buf, const AcGroup* G,
                                                int L, int ncl, int rank, int gpar) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __syncthreads();                                        // previous users of the tables / of buf are done
  for (int i = tid; i < 512; i += kAcThreads) sm.f.mtab[i & 15][i >> 4] = (&G->mlane[0][0])[i];
  for (int i = tid; i < L * 4; i += kAcThreads) (&sm.f.wt[0][0])[i] = (&G->wt[0][0])[i];
  for (int i = tid; i < 144; i += kAcThreads) (&sm.f.mp[0][0])[i] = (&G->mp[0][0])[i];
  for (int i = tid; i < kAcWarps * 16; i += kAcThreads) (&sm.f.mwarp[0][0])[i] = (&G->mwarp[0][0])[i];
  for (int i = tid; i < kAcMaxCluster * 16; i += kAcThreads) (&sm.f.prop_pow[0][0])[i] = (&G->prop_pow[0][0])[i];
  if (tid < 10) (&sm.f.c[0][0])[tid] = (&G->c[0][0])[tid];
  __syncthreads();
  float* mine = buf + tid * L;
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
  const double raw[4] = {p[0], p[1], p[2], p[3]};
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
  if (warp == 0) {                                        // scan the sixteen warp aggregates
    double v[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) v[s] = (lane < kAcWarps) ? sm.f.wagg[lane][s] : 0.0;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      double u[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) u[s] = __shfl_up_sync(kFull, v[s], 1 << d);
      if (lane >= (1 << d)) mv4_acc(sm.f.mp[5 + d], u, v);
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const double e = __shfl_up_sync(kFull, v[s], 1);
      if (lane < kAcWarps) sm.f.wcar[lane][s] = lane ? e : 0.0;
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
    for (int s = 0; s < 4; ++s) wc[s] = sm.f.wcar[warp][s];
    ac_mv4_lane_acc(sm.f.mtab, lane, wc, z);                 // chunk start state for a zero slice start
  }
  if (ncl > 1) {
    if (tid == kAcThreads - 1) {                            // S = threads * L: the slice ends with the last chunk
      double e[4] = {raw[0], raw[1], raw[2], raw[3]};
      mv4_acc(sm.f.mp[0], z, e);                            // E = M z + p
      for (int rk = 0; rk < ncl; ++rk) {
        double* dst = cluster.map_shared_rank(&sm.xE[gpar][rank][0], rk);
#pragma unroll
        for (int s = 0; s < 4; ++s) dst[s] = e[s];
      }
    }
    ac_cluster_sync(ncl);
    double c0[4] = {0.0, 0.0, 0.0, 0.0}, c1[4] = {0.0, 0.0, 0.0, 0.0};
    for (int r = 0; r < rank; r += 2) {
      double e0[4], e1[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) { e0[s] = sm.xE[gpar][r][s]; e1[s] = (r + 1 < rank) ? sm.xE[gpar][r + 1][s] : 0.0; }
      mv4_acc(sm.f.prop_pow[rank - 1 - r], e0, c0);
      if (r + 1 < rank) mv4_acc(sm.f.prop_pow[rank - 2 - r], e1, c1);
    }
    double cs[4], c[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) cs[s] = c0[s] + c1[s];
    ac_mv4_set(sm.f.mwarp[warp], cs, c);
    ac_mv4_lane_acc(sm.f.mtab, lane, c, z);
  }
  const double b00 = sm.f.c[0][0], b01 = sm.f.c[0][1], b02 = sm.f.c[0][2], a01 = sm.f.c[0][3], a02 = sm.f.c[0][4];
  const double b10 = sm.f.c[1][0], b11 = sm.f.c[1][1], b12 = sm.f.c[1][2], a11 = sm.f.c[1][3], a12 = sm.f.c[1][4];
#pragma unroll 4
  for (int j = 0; j < L; ++j) {
    const double xv = (double)mine[j];
    const double y0 = fma(b00, xv, z[0]);
    z[0] = fma(-a01, y0, fma(b01, xv, z[1]));
    z[1] = fma(-a02, y0, b02 * xv);
    const double y1 = fma(b10, y0, z[2]);
    z[2] = fma(-a11, y1, fma(b11, y0, z[3]));
    z[3] = fma(-a12, y1, b12 * y0);
    mine[j] = (float)y1;
  }
  __syncthreads();
}

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
  const int nq = (n + 3) >> 2;                               // groups of four samples (s0 is a multiple of 4)
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(xr) | reinterpret_cast<uintptr_t>(yr)) & 15u) == 0;
  int parity = 0;

  // ---- load + statistics of the raw row.  A 16-byte aligned slice arrives as ONE bulk asynchronous copy (TMA,
  //      cp.async.bulk global -> shared, completion counted on an mbarrier): no load instructions in the sweep and the
  //      whole HBM latency paid once; otherwise plain loads.
  AcAcc st[1];
  AcMap map[1];
  st[0].init();
  const bool bulk = vec_ok && n > 0 && (n & 3) == 0;
  if (bulk) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&sm.load_bar);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const uint32_t bytes = (uint32_t)n * 4u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       (uint32_t)__cvta_generic_to_shared(buf)),
                   "l"(xr), "r"(bytes), "r"(bar)
                   : "memory");
    }
    for (int i = n + tid; i < P.S; i += kAcThreads) buf[i] = 0.f;       // the slice's tail while the copy flies
    __syncthreads();                                                     // the barrier is initialised for everybody
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
    const float4* buf4 = reinterpret_cast<const float4*>(buf);
    for (int q = tid; q < nq; q += kAcThreads) {
      const float4 w = buf4[q];
      st[0].add(w.x); st[0].add(w.y); st[0].add(w.z); st[0].add(w.w);
      st[0].flush();
    }
  } else {
    for (int q = tid; q < (P.S >> 2); q += kAcThreads) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int i = 4 * q;
      if (i + 3 < n && vec_ok) {
        const float4 w = ld_stream4(reinterpret_cast<const float4*>(xr + i));
        v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (i + k < n) v[k] = ld_stream(xr + i + k);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        buf[i + k] = v[k];
        if (i + k < n) st[0].add(v[k]);
      }
      st[0].flush();
    }
  }
  ac_exchange<1>(sm, cluster, ncl, rank, P.t, parity, st, map);

  // ---- one elementwise stage: buf <- blend(transform(N_prev(buf))), statistics of the result
  auto stage = [&](int op, const float* rowp, const float* noise, const float* mask, unsigned long long seed,
                   unsigned long long sid) {
    StageArgs a;
    a.op = op; a.fs = P.fs; a.rowp = rowp; a.noise = noise; a.mask = mask; a.seed = seed; a.stream = sid;
    const float* p = rowp ? rowp + (long long)row * 8 : nullptr;
    const float* nz = noise ? noise + (long long)row * P.t : nullptr;
    const bool on = op != MPCG_AUG_IDENTITY && (mask == nullptr || mask[row] != 0.f);
    const bool philox = on && op == MPCG_AUG_NOISE && nz == nullptr;
    // A stage that leaves the row alone would only normalise an already normalised row again: mean 0 and peak 1 up to
    // float32 rounding, i.e. the identity to ~1e-7.  With P.collapse the pending map simply stays pending (unless the
    // row is degenerate: a peak below 1e-6 means the previous map amplified rounding residue).
    if (!on && P.collapse && map[0].inv < 1e6f) return;
    st[0].init();
    const AcMap m = map[0];
    float4* buf4 = reinterpret_cast<float4*>(buf);
    if (on && op == MPCG_AUG_SINE_MUL) {
      // wandering volume: each thread evaluates the two sinusoids exactly once (at its first sample, with the
      // reference's rounding sequence) and then ROTATES: by one sample inside a group of four, by kAcThreads groups from
      // one of its groups to the next.  At most ~40 rotations per thread: ~3e-6 rad of phase, < 1e-6 of the output.
      const float two_pi = 6.283185307179586f;
      float cd[2], sd[2], cj[2], sj[2], sg[2], cg[2];
      const float tt0 = __fdiv_rn((float)(s0 + 4 * tid), P.fs);
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const float dw = __fdiv_rn(__fmul_rn(two_pi, p[3 * b + 1]), P.fs);          // phase step per sample
        sincosf(dw, &sd[b], &cd[b]);
        sincosf(dw * (float)(4 * kAcThreads), &sj[b], &cj[b]);
        sincosf(__fmul_rn(two_pi, __fadd_rn(__fmul_rn(p[3 * b + 1], tt0), p[3 * b + 2])), &sg[b], &cg[b]);
      }
      for (int q = tid; q < nq; q += kAcThreads) {
        const float4 b4 = buf4[q];
        float w[4] = {b4.x, b4.y, b4.z, b4.w};
        float sn[2] = {sg[0], sg[1]}, cs[2] = {cg[0], cg[1]};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float mod = __fadd_rn(__fadd_rn(0.f, __fmul_rn(p[0], sn[0])), __fmul_rn(p[3], sn[1]));
          const float v = m(w[k]);
          w[k] = __fmul_rn(v, __fadd_rn(1.f, mod));
          if (4 * q + k < n) st[0].add(w[k]);
          if (k < 3) {
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              const float s_n = fmaf(sn[b], cd[b], cs[b] * sd[b]), c_n = fmaf(cs[b], cd[b], -sn[b] * sd[b]);
              sn[b] = s_n; cs[b] = c_n;
            }
          }
        }
        buf4[q] = make_float4(w[0], w[1], w[2], w[3]);
        st[0].flush();
#pragma unroll
        for (int b = 0; b < 2; ++b) {                         // on to this thread's next group
          const float s_n = fmaf(sg[b], cj[b], cg[b] * sj[b]), c_n = fmaf(cg[b], cj[b], -sg[b] * sj[b]);
          sg[b] = s_n; cg[b] = c_n;
        }
      }
    } else {
    for (int q = tid; q < nq; q += kAcThreads) {
      float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (philox) z4 = philox_normal4(seed, sid, row, (long long)((s0 >> 2) + q));
      const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
      const float4 b4 = buf4[q];
      float w[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = 4 * q + k;
        const float v = m(w[k]);
        w[k] = (on && i < n) ? stage_value(a, p, nz, row, (long long)s0 + i, v, zz[k]) : v;
        if (i < n) st[0].add(w[k]);
      }
      buf4[q] = make_float4(w[0], w[1], w[2], w[3]);
      st[0].flush();
    }
    }
    ac_exchange<1>(sm, cluster, ncl, rank, P.t, parity, st, map);
  };

  stage(MPCG_AUG_NOISE, P.rowp1, P.noise1, P.mask1, P.seed1, P.sid1);
  stage(MPCG_AUG_SINE_MUL, P.rowp2, nullptr, P.mask2, 0, 0);

  // ---- EQ
  const bool eq_on = P.ngroups > 0 && (P.mask3 == nullptr || P.mask3[row] != 0.f);
  if (!eq_on) {
    stage(MPCG_AUG_IDENTITY, nullptr, nullptr, nullptr, 0, 0);
  } else {
    // x2 = N(previous) replaces the slice and is parked in the output row; its statistics ride along
    AcAcc s2[2];
    AcMap m2[2];
    s2[0].init(); s2[1].init();
    {
      const AcMap m = map[0];
      for (int i = tid; i < P.S; i += kAcThreads) {
        float v = 0.f;
        if (i < n) {
          v = m(buf[i]);
          yr[i] = v;
          s2[1].add(v);
        }
        buf[i] = v;                                          // zeros beyond the row's end feed the filter
        if ((i / kAcThreads & 3) == 3) s2[1].flush();
      }
    }
    for (int g = 0; g < P.ngroups; ++g) ac_filter_group(sm, cluster, buf, P.groups + g, P.L, ncl, rank, g & 1);
    for (int i = tid; i < n; i += kAcThreads) {
      s2[0].add(buf[i]);
      if ((i / kAcThreads & 3) == 3) s2[0].flush();
    }
    ac_exchange<2>(sm, cluster, ncl, rank, P.t, parity, s2, m2);        // N(c) and N(x2)
    st[0].init();
    for (int i = tid; i < n; i += kAcThreads) {
      const float v = __fadd_rn(__fdiv_rn(m2[0](buf[i]), 50.f), m2[1](yr[i]));
      buf[i] = v;
      st[0].add(v);
      if ((i / kAcThreads & 3) == 3) st[0].flush();
    }
    ac_exchange<1>(sm, cluster, ncl, rank, P.t, parity, st, map);       // e = N(v)
    stage(MPCG_AUG_IDENTITY, nullptr, nullptr, nullptr, 0, 0);         // x3 = N(e)
  }

  stage(MPCG_AUG_NOISE, P.rowp4, P.noise4, P.mask4, P.seed4, P.sid4);

  // ---- final map + store
  {
    const AcMap m = map[0];
    for (int q = tid; q < nq; q += kAcThreads) {
      const int i = 4 * q;
      if (i + 3 < n && vec_ok) {
        const float4 b4 = reinterpret_cast<const float4*>(buf)[q];
        st_stream4(reinterpret_cast<float4*>(yr + i), make_float4(m(b4.x), m(b4.y), m(b4.z), m(b4.w)));
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (i + k < n) yr[i + k] = m(buf[i + k]);
      }
    }
  }
}

}  // namespace mpcg

// EQ recipes travel through the caller's workspace (the ABI never allocates and keeps no device state between calls):
// uploaded on the caller's stream right before the launch, so calls on different streams or from different host
// threads cannot disturb each other as long as each brings its own workspace.
extern "C" int64_t mpcg_aug_chain_work_bytes(void) { return (int64_t)sizeof(mpcg::AcGroup) * mpcg::kAcMaxGroups; }

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
  if (x == y) return MPCG_EINVAL;                            // the output row is scratch while the EQ runs
  // geometry: slices of S = threads * L samples, the smallest power-of-two cluster with L <= kAcLmax (a 66 KB slice,
  // two CTAs per SM).  MPCG_AC_CLUSTER overrides (experiments).
  int ncl = 0, L = 0;
  {
    const char* env = getenv("MPCG_AC_CLUSTER");
    const int forced = env ? atoi(env) : 0;
    for (int c = 1; c <= kAcMaxCluster; c *= 2) {
      if (forced > 0 && c != forced) continue;
      const long long per = (t + c - 1) / c;
      long long l = (per + kAcThreads - 1) / kAcThreads;
      l |= 1;
      if (l <= kAcLmax) { ncl = c; L = (int)l; break; }
    }
  }
  if (ncl == 0) return MPCG_EUNSUPPORTED;
  const int S = kAcThreads * L;
  if ((long long)(ncl - 1) * S >= t && ncl > 1) return MPCG_EUNSUPPORTED;   // every rank must own part of the row
  if (rows * ncl > 0x7fffffffLL) return MPCG_ERANGE;
  const int ngroups = (eq_sections + 1) / 2;

  AcGroup groups[kAcMaxGroups];
  memset(groups, 0, sizeof(groups));
  for (int g = 0; g < ngroups; ++g) {
    AcGroup& k = groups[g];
    bool ok;
    bq_group_coeffs(eq_sos, eq_sections, 2 * g, k.c, &ok);
    if (!ok) return MPCG_EINVAL;
    double A[16], B[4];
    bq_group_AB(k.c, A, B);
    double v[4] = {B[0], B[1], B[2], B[3]};
    for (int j = L - 1; j >= 0; --j) {
      for (int s = 0; s < 4; ++s) k.wt[j][s] = v[s];
      bq_group_step(k.c, v, 0.0);
    }
    bq_mat_pow(A, L, k.mp[0]);
    for (int d = 1; d < 9; ++d) bq_mat_mul(k.mp[d - 1], k.mp[d - 1], k.mp[d]);
    for (int l = 0; l < 32; ++l) bq_mat_pow(k.mp[0], l, k.mlane[l]);
    for (int w = 0; w < kAcWarps; ++w) bq_mat_pow(k.mp[0], 32LL * w, k.mwarp[w]);
    for (int j = 0; j < kAcMaxCluster; ++j) bq_mat_pow(A, (long long)S * j, k.prop_pow[j]);
  }
  const AcGroup* dev_groups = nullptr;
  if (ngroups > 0) {
    if (!work || ((uintptr_t)work & 15u) || work_bytes < (int64_t)sizeof(groups)) return MPCG_EINVAL;
    // (pageable source: the runtime stages the bytes before returning, so `groups` may leave scope)
    cudaError_t e = cudaMemcpyAsync(work, groups, sizeof(AcGroup) * ngroups, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return (int)e;
    dev_groups = reinterpret_cast<const AcGroup*>(work);
  }

  AcParams P;
  memset(&P, 0, sizeof(P));
  P.x = x; P.y = y; P.t = (int)t; P.fs = fs; P.ncl = ncl; P.S = S; P.L = L;
  P.rowp1 = rowp1; P.noise1 = noise1; P.mask1 = mask1; P.seed1 = seed1; P.sid1 = sid1;
  P.rowp2 = rowp2; P.mask2 = mask2; P.mask3 = mask3;
  P.rowp4 = rowp4; P.noise4 = noise4; P.mask4 = mask4; P.seed4 = seed4; P.sid4 = sid4;
  P.ngroups = ngroups; P.groups = dev_groups; P.collapse = (flags & MPCG_AUG_CHAIN_COLLAPSE) ? 1 : 0;
  if (t > 0x7fffffffLL) return MPCG_ERANGE;

  const size_t smem = sizeof(AcShared) + (size_t)S * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(aug_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(rows * ncl));
  cfg.blockDim = dim3(kAcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)ncl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, aug_chain_kernel, P);
  if (e != cudaSuccess) return (int)e;
  return MPCG_OK;
}
