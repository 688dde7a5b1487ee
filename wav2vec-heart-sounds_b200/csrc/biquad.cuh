// Causal biquad cascades as a chunked linear-recurrence scan (replaces torchaudio's one-thread-
// per-row iir_cu_kernel and SciPy's sosfilt for torchproc.lowpass/highpass/bandpass_cascade and
// torchaug.parametric_eq; reference call sites signalproc/torchproc.py:38-53,
// signalproc/filters.py:25-39, augment/torchaug.py:88-100).
//
// Scheme (per row, per tile of THREADS*L samples held in shared memory, sections taken two at a
// time as one 4-state linear system  s' = A s + B x ):
//   pass 1  each thread: zero-state end state of its L-sample chunk  p = sum_j A^(L-1-j) B x[j]
//   scan    s_(k+1) = M s_k + p_k  with the constant  M = A^L  (warp Hillis-Steele with M^(2^d),
//           one thread chains the warp aggregates, M^lane fixes up each lane)
//   pass 2  each thread re-runs its chunk from its true start state in transposed direct form II
//           and overwrites the chunk with the filtered samples.
// All recurrence state is fp64 (pole radii reach 0.9997); samples are fp32 in memory.
#pragma once
#include "common.cuh"

namespace mpcg {

constexpr int kBqL = 61;                 // samples per thread-chunk; odd => conflict-free smem strides
constexpr int kBqMaxGroups = 3;          // up to 6 second-order sections per call
constexpr int kBqStates = 4;             // two sections per group

struct BqGroup {
  double c[2][5];                        // per section: b0 b1 b2 a1 a2 (a0 == 1)
  double wt[kBqL][kBqStates];            // pass-1 weights  A^(L-1-j) B
  double mp[6][kBqStates * kBqStates];   // M, M^2, M^4, M^8, M^16, M^32 (row-major)
};

// Host helpers shared with the fused kernel's planner (biquad.cu).
void bq_group_coeffs(const double* sos, int n_sections, int first, double c[2][5], bool* ok);
void bq_group_AB(const double c[2][5], double A[16], double B[4]);
void bq_mat_pow(const double A[16], long long n, double out[16]);
void bq_mat_mul(const double* a, const double* b, double* out);
void bq_group_step(const double c[2][5], double z[4], double x);
struct BqPlan {
  int ngroups;
  int pad_;
  BqGroup g[kBqMaxGroups];
};

// Host: fill a plan from n_sections rows of SciPy-layout sos [b0 b1 b2 a0 a1 a2]; returns 0 or MPCG_E*.
int bq_make_plan(const double* sos, int n_sections, BqPlan* plan);

// 4x4 row-major matrix times vector, accumulated into acc.
__device__ __forceinline__ void mv4_acc(const double* __restrict__ m, const double (&v)[4], double (&acc)[4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    double a = acc[r];
#pragma unroll
    for (int c = 0; c < 4; ++c) a = fma(m[r * 4 + c], v[c], a);
    acc[r] = a;
  }
}

// Shared-memory scratch one CTA needs for the scan (besides the sample tile itself).
template <int THREADS>
struct BqScratch {
  double mtab[kBqMaxGroups][32][16];     // M^lane per group, built once per CTA
  double wagg[THREADS / 32][4];          // warp aggregates
  double wcar[THREADS / 32][4];          // state at the start of each warp's first chunk
  double carry[kBqMaxGroups][4];         // state at the start of the next tile, per group
};

template <int THREADS>
__device__ __forceinline__ void bq_init_scratch(BqScratch<THREADS>& sc, const BqPlan& plan) {
  const int tid = threadIdx.x;
  // M^lane by binary powers: every lane multiplies in the M^(2^d) whose bit is set in its index.
  if (tid < 32) {
    for (int g = 0; g < plan.ngroups; ++g) {
      double acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = (i % 5 == 0) ? 1.0 : 0.0;
#pragma unroll
      for (int d = 0; d < 5; ++d) {
        if ((tid >> d) & 1) {
          const double* m = plan.g[g].mp[d];
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
      for (int i = 0; i < 16; ++i) sc.mtab[g][tid][i] = acc[i];
    }
  }
  if (tid < kBqMaxGroups * 4) sc.carry[tid >> 2][tid & 3] = 0.0;
}

// Filter THREADS*kBqL samples in place in shared memory (sh[0 .. THREADS*kBqL)); samples past the
// end of the row must be zero.  Carries the state between calls in sc.carry.  Ends with all
// threads past their last write but NOT synchronised: the caller syncs before reading the tile.
template <int THREADS>
__device__ __forceinline__ void bq_filter_tile(float* sh, BqScratch<THREADS>& sc, const BqPlan& plan) {
  constexpr int NW = THREADS / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* mine = sh + tid * kBqL;
  for (int g = 0; g < plan.ngroups; ++g) {
    const BqGroup& G = plan.g[g];
    // ---- pass 1: zero-state end state of this thread's chunk
    double p[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < kBqL; ++j) {
      const double xv = (double)mine[j];
#pragma unroll
      for (int s = 0; s < 4; ++s) p[s] = fma(G.wt[j][s], xv, p[s]);
    }
    // ---- inclusive warp scan of  s_(k+1) = M s_k + p_k
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      double u[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) u[s] = __shfl_up_sync(kFull, p[s], 1 << d);
      if (lane >= (1 << d)) mv4_acc(G.mp[d], u, p);
    }
    if (lane == 31) {
#pragma unroll
      for (int s = 0; s < 4; ++s) sc.wagg[warp][s] = p[s];
    }
    __syncthreads();
    if (tid == 0) {
      double c[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) c[s] = sc.carry[g][s];
      for (int w = 0; w < NW; ++w) {
        double nx[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) { sc.wcar[w][s] = c[s]; nx[s] = sc.wagg[w][s]; }
        mv4_acc(G.mp[5], c, nx);
#pragma unroll
        for (int s = 0; s < 4; ++s) c[s] = nx[s];
      }
#pragma unroll
      for (int s = 0; s < 4; ++s) sc.carry[g][s] = c[s];
    }
    __syncthreads();
    // ---- true start state of this chunk = (exclusive scan) + M^lane * (warp start state)
    double z[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const double e = __shfl_up_sync(kFull, p[s], 1);
      z[s] = lane ? e : 0.0;
    }
    {
      double wc[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) wc[s] = sc.wcar[warp][s];
      mv4_acc(sc.mtab[g][lane], wc, z);
    }
    // ---- pass 2: transposed direct form II, two sections back to back
    const double b00 = G.c[0][0], b01 = G.c[0][1], b02 = G.c[0][2], a01 = G.c[0][3], a02 = G.c[0][4];
    const double b10 = G.c[1][0], b11 = G.c[1][1], b12 = G.c[1][2], a11 = G.c[1][3], a12 = G.c[1][4];
#pragma unroll 4
    for (int j = 0; j < kBqL; ++j) {
      const double xv = (double)mine[j];
      const double y0 = fma(b00, xv, z[0]);
      z[0] = fma(-a01, y0, fma(b01, xv, z[1]));
      z[1] = fma(-a02, y0, b02 * xv);
      const double y1 = fma(b10, y0, z[2]);
      z[2] = fma(-a11, y1, fma(b11, y0, z[3]));
      z[3] = fma(-a12, y1, b12 * y0);
      mine[j] = (float)y1;
    }
  }
}

}  // namespace mpcg
