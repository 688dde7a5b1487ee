// extern "C" entry: mpcg_mel_tc_f32 -- mel framing as a DFT GEMM on the 5th-generation tensor cores (tcgen05).
//
// Reformulation that makes the contraction dense AND non-redundant.  With n_fft = Q * hop the padded signal is a
// matrix of non-overlapping hop rows  Xh[g][j] = xp[g*hop + j].  The RECTANGULAR-window partial DFT of a hop row,
//       P[g][k] = sum_{j<hop} Xh[g][j] * exp(-2 pi i k j / n_fft),
// is computed once per hop row and shared by the Q frames that contain it:
//       Xrect[f][k] = sum_{q<Q} exp(-2 pi i k q / Q) * P[f+q][k]          (Q = 4: multiples of 90 degrees)
// and the periodic Hann window is applied in the frequency domain,
//       X[f][k] = 0.5 Xrect[f][k] - 0.25 (Xrect[f][k-1] + Xrect[f][k+1]).
// So the tensor cores run ONE GEMM  [hop rows x hop] . [hop x 2*(bins+2)]  with a basis that does not depend on q
// and stays resident in shared memory: 4x fewer flops than framing first, and no im2col.
//
// Precision: operands are split fp16 pairs (x = hi + lo, e = hi + lo), three MMAs per k-step (hi*hi + hi*lo + lo*hi)
// accumulate in fp32 in TMEM: ~2^-22 relative, i.e. fp32-class.  This is the `fast` tier of MelConfig.build():
// like the fp32 FMA path it resolves leakage skirts only down to ~1e-6 of a frame's largest rectangular-window bin.
//
// Tiles are cut from the CONCATENATED hop rows of all signals (frames + Q - 1 per signal), fpt = 128 - (Q - 1)
// frames each, so a signal's last frames share a tile with the next signal's first ones instead of leaving a
// nearly empty tile per signal; frames that would straddle two signals are simply not stored.
// One persistent CTA per SM (512 threads).  Per tile of 128 hop rows: threads load + split the samples straight
// into the canonical K-major (no swizzle) core-matrix layout, one elected thread issues 3 * hop/16 tcgen05.mma
// (M=128, N<=256, K=16) into a TMEM accumulator, a tcgen05.commit on an mbarrier signals completion, four warps
// pull the accumulator back with tcgen05.ld, and the epilogue (twiddle sum, Hann, magnitude, mel projection, dB
// map) runs out of shared memory.
#include "common.cuh"
#include <cuda_fp16.h>

namespace mpcg {

constexpr int kTcThreads = 512;            // 16 warps: TMEM lane quarter = warp & 3, column share = warp >> 2
constexpr int kTcRows = 128;                 // hop rows per tile = UMMA M

struct MelTcArgs {
  const float* x;          // [rows, t]
  float* out;              // [rows, n_mels, frames]
  const __half* basis;     // [2 (hi, lo)][N * hop] in canonical K-major core-matrix order (see host packer)
  const float* fb;         // [nbins][n_mels]
  const float* twq;        // [Q][2] cos, sin of -2 pi m / Q
  long long t, rows;
  int n_fft, hop, Q, k0, nbins, N, n_mels, frames, log_map, rps;   // rps: hop rows per signal
  long long total_tiles;
  float inv_norm;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // SM100 shared-memory matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
  // layout_type [61,64) = 0 (no swizzle)
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float mel_log_map(float mel) {
  // (20 log10(max(x, 1e-5)) - 20 + 100) / 100 with log10 = log2 * log10(2) on the SFU: |error| of lg2.approx is
  // ~2^-22 absolute, i.e. ~1.4e-8 after the scaling -- three orders inside the 1e-5 tolerance of the map's [0, 1] range
  const float db = 6.020599913279624f * __log2f(fmaxf(mel, 1e-5f)) - 20.f;
  return fminf(fmaxf((db + 100.f) * 0.01f, 0.f), 1.f);
}

__global__ void __launch_bounds__(kTcThreads, 1)
mel_tc_kernel(const MelTcArgs a) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  const int hop = a.hop, N = a.N, Q = a.Q;
  const int a_bytes = kTcRows * hop * 2;                      // one half (hi or lo) of the A tile
  const int b_bytes = N * hop * 2;
  unsigned char* A_hi = tc_smem;
  unsigned char* A_lo = tc_smem + a_bytes;
  unsigned char* B_hi = tc_smem + 2 * a_bytes;
  unsigned char* B_lo = B_hi + b_bytes;
  unsigned char* tail = B_lo + b_bytes;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(tail);         // 8 B
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 8);
  float* twq = reinterpret_cast<float*>(tail + 16);           // [Q][2]
  int* row_sig = reinterpret_cast<int*>(tail + 16 + 64);      // [128] signal of each hop row of the tile (-1: none)
  int* row_loc = row_sig + kTcRows;                           // [128] its hop-row index inside that signal
  float* fbs = reinterpret_cast<float*>(row_loc + kTcRows);   // [nbins][n_mels4] filterbank, rows padded to 4 mels
  const int nm4_ = (a.n_mels + 3) & ~3;
  float* twb = fbs + a.nbins * nm4_;                          // [nbins + 2][Q][2] twiddle of (bin, hop row): exp(-2 pi i k q / Q)
  int* mk0 = reinterpret_cast<int*>(twb + (a.nbins + 2) * a.Q * 2);   // [n_mels] first / last bin with weight
  int* mk1 = mk0 + a.n_mels;
  float* stage = reinterpret_cast<float*>(tc_smem);           // [128][N + 1], reuses the A region after the MMAs
  const int nb2 = a.nbins + 2;
  const int srow = N + 1;
  const int mstride = (a.nbins + 1) | 1;                      // odd: frames run across lanes
  float* mags = stage + kTcRows * srow;                       // [FPT][mstride]
  const int rstride = nb2 | 1;
  float* rect_re = mags + kTcRows * mstride;                  // [FPT][rstride] rectangular-window spectra of the frames
  float* rect_im = rect_re + kTcRows * rstride;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int fpt = kTcRows - (Q - 1);                          // frames per tile

  // ---- one-time setup: basis to shared memory, mbarrier, TMEM
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.basis);
    uint4* dst = reinterpret_cast<uint4*>(B_hi);
    for (int i = tid; i < (2 * b_bytes) / 16; i += kTcThreads) dst[i] = __ldg(src + i);
  }
  if (tid < 2 * Q) twq[tid] = a.twq[tid];
  const int nm4 = (a.n_mels + 3) & ~3;
  for (int i = tid; i < a.nbins * nm4; i += kTcThreads) {
    const int k = i / nm4, m = i - k * nm4;
    fbs[i] = m < a.n_mels ? a.fb[(long long)k * a.n_mels + m] : 0.f;
  }
  for (int i = tid; i < nb2 * Q; i += kTcThreads) {           // twiddles per (bin, hop row) from the Q-point table
    const int b = i / Q, q = i - b * Q;
    const int k = a.k0 - 1 + b;
    const int m = (int)((((long long)k * q) % Q + Q) % Q);
    twb[2 * i] = a.twq[2 * m];
    twb[2 * i + 1] = a.twq[2 * m + 1];
  }
  __syncthreads();                                            // fbs complete
  for (int m = tid; m < a.n_mels; m += kTcThreads) {          // each mel filter's support (a short run of bins)
    int k0m = a.nbins, k1m = -1;
    for (int k = 0; k < a.nbins; ++k)
      if (fbs[k * nm4 + m] != 0.f) { k0m = k < k0m ? k : k0m; k1m = k; }
    mk0[m] = k0m; mk1[m] = k1m;
  }
  if (tid == 0) mbar_init(mbar, 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // instruction descriptor: D fp32, A/B fp16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);
  const uint32_t sbo = (uint32_t)hop * 16u;                   // bytes between 8-row groups: (hop/8) core matrices of 128 B
  const int pad = a.n_fft / 2;
  const long long total_tiles = a.total_tiles;
  const bool q_pow2 = (Q & (Q - 1)) == 0;
  uint32_t phase = 0;

  for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const long long G0 = tile * fpt;                          // first global hop row (= first frame slot) of this tile
    if (tid < kTcRows) {
      const long long G = G0 + tid;
      const long long sg = G / a.rps;
      row_sig[tid] = sg < a.rows ? (int)sg : -1;
      row_loc[tid] = (int)(G - sg * a.rps);
    }
    __syncthreads();
    // ---- A tile: 128 hop rows x hop samples, split into fp16 hi / lo, canonical core-matrix order
    const int chunks_per_row = hop >> 3;
    const int nchunks = kTcRows * chunks_per_row;
    const bool cpr_pow2 = (chunks_per_row & (chunks_per_row - 1)) == 0;
    const int cpr_shift = 31 - __clz(chunks_per_row);
    // Four chunks per thread in flight: all global loads of a batch are issued before the first conversion.
    constexpr int kBatch = 4;
    for (int idx0 = tid; idx0 < nchunks; idx0 += kBatch * kTcThreads) {
      float4 p0[kBatch], p1[kBatch];
      int off[kBatch], kind[kBatch];                          // kind: 0 = zeros, 1 = loaded (aligned interior), 2 = edge
      const float* xrs[kBatch];
      long long s0s[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int idx = idx0 + u * kTcThreads;
        kind[u] = -1;
        if (idx < nchunks) {
          const int r8 = idx & 7;
          const int rest = idx >> 3;
          const int g8 = cpr_pow2 ? (rest >> cpr_shift) : rest / chunks_per_row;
          const int jc = rest - g8 * chunks_per_row;
          const int g = g8 * 8 + r8;
          const int sg = row_sig[g];
          const float* xr = a.x + (long long)(sg < 0 ? 0 : sg) * a.t;
          const long long s0 = (long long)row_loc[g] * hop - pad + jc * 8;   // first un-padded sample index of the chunk
          off[u] = g8 * (int)sbo + jc * 128 + r8 * 16;
          xrs[u] = xr; s0s[u] = s0;
          if (sg < 0 || s0 >= a.t + pad) {                    // past the last signal / past the reflected tail
            kind[u] = 0;
          } else if (s0 >= 0 && s0 + 7 < a.t && ((((uintptr_t)(xr + s0)) & 15u) == 0)) {
            kind[u] = 1;
            p0[u] = __ldg(reinterpret_cast<const float4*>(xr + s0));
            p1[u] = __ldg(reinterpret_cast<const float4*>(xr + s0 + 4));
          } else {
            kind[u] = 2;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        if (kind[u] < 0) continue;
        float v[8];
        if (kind[u] == 1) {
          v[0] = p0[u].x; v[1] = p0[u].y; v[2] = p0[u].z; v[3] = p0[u].w;
          v[4] = p1[u].x; v[5] = p1[u].y; v[6] = p1[u].z; v[7] = p1[u].w;
        } else if (kind[u] == 0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = 0.f;
        } else {
          long long js[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            long long j = s0s[u] + e;
            if (j < 0) j = -j;                                // reflect padding (no edge repeat)
            if (j >= a.t) j = 2 * (a.t - 1) - j;
            js[e] = j;
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = (js[e] >= 0 && js[e] < a.t) ? __ldg(xrs[u] + js[e]) : 0.f;
        }
        __half2 hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __half h0 = __float2half_rn(v[2 * e]), h1 = __float2half_rn(v[2 * e + 1]);
          hi[e] = __halves2half2(h0, h1);
          lo[e] = __halves2half2(__float2half_rn(v[2 * e] - __half2float(h0)), __float2half_rn(v[2 * e + 1] - __half2float(h1)));
        }
        *reinterpret_cast<uint4*>(A_hi + off[u]) = *reinterpret_cast<uint4*>(hi);
        *reinterpret_cast<uint4*>(A_lo + off[u]) = *reinterpret_cast<uint4*>(lo);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
    __syncthreads();
    // ---- MMAs: one elected thread
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ah = smem_u32(A_hi), al = smem_u32(A_lo), bh = smem_u32(B_hi), bl = smem_u32(B_lo);
      const int ksteps = hop >> 4;
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint32_t ko = (uint32_t)ks * 256u;              // two 128-byte core matrices per K = 16 step
        const uint64_t dah = umma_desc(ah + ko, 128, sbo), dal = umma_desc(al + ko, 128, sbo);
        const uint64_t dbh = umma_desc(bh + ko, 128, sbo), dbl = umma_desc(bl + ko, 128, sbo);
        umma_f16(tmem_base, dah, dbh, idesc, ks > 0 ? 1u : 0u);
        umma_f16(tmem_base, dah, dbl, idesc, 1u);
        umma_f16(tmem_base, dal, dbh, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar))
                   : "memory");
    }
    mbar_wait(mbar, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- accumulator -> shared staging (A is dead now): row = TMEM lane, N columns.  A warp reaches the 32 TMEM
    //      lanes of quarter (warp & 3); the four warps of a quarter share its columns in chunks of eight.
    {
      const int quarter = warp & 3;
      const int r = quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      for (int c0 = (warp >> 2) * 8; c0 < N; c0 += 8 * (kTcThreads / 128)) {
        float v[8];
        tmem_ld8(taddr + (uint32_t)c0, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) stage[r * srow + c0 + e] = v[e];
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---- epilogue thread map: thread = (frame f = tid % fpt, lane group gi = tid / fpt); groups stride the bins / mels
    const int ngrp = kTcThreads / fpt;                        // whole groups of fpt threads (the few left over idle here)
    const int gi = tid / fpt, f = tid - gi * fpt;
    const bool epi = gi < ngrp;
    // ---- frame spectra, pass A: rectangular-window bins of every frame = twiddle sum over its Q hop rows
    if (epi) {
      for (int b = gi; b < nb2; b += ngrp) {
        float sr = 0.f, si = 0.f;
        for (int q = 0; q < Q; ++q) {
          const float wr = twb[2 * (b * Q + q)], wi = twb[2 * (b * Q + q) + 1];
          const float pr = stage[(f + q) * srow + b], pi = stage[(f + q) * srow + nb2 + b];
          sr = fmaf(wr, pr, fmaf(-wi, pi, sr));
          si = fmaf(wr, pi, fmaf(wi, pr, si));
        }
        rect_re[f * rstride + b] = sr;
        rect_im[f * rstride + b] = si;
      }
    }
    __syncthreads();
    // ---- pass B: Hann in the frequency domain (0.5 X[k] - 0.25 (X[k-1] + X[k+1])), magnitude
    if (epi) {
      const float* rr = rect_re + f * rstride;
      const float* ri = rect_im + f * rstride;
      for (int kb = gi; kb < a.nbins; kb += ngrp) {
        const float xr2 = 0.5f * rr[kb + 1] - 0.25f * (rr[kb] + rr[kb + 2]);
        const float xi2 = 0.5f * ri[kb + 1] - 0.25f * (ri[kb] + ri[kb + 2]);
        mags[f * mstride + kb] = sqrtf(xr2 * xr2 + xi2 * xi2) * a.inv_norm;
      }
    }
    __syncthreads();
    // ---- mel projection over each filter's own bins, dB map, store (lanes run over frames: contiguous stores)
    if (epi) {
      const int sg = row_sig[f], fg = row_loc[f];             // frame slot f = hop row f of the tile
      if (sg >= 0 && fg < a.frames) {
        const float* mg = mags + f * mstride;
        float* op = a.out + (long long)sg * a.n_mels * a.frames + fg;
        for (int m = gi; m < a.n_mels; m += ngrp) {
          float acc = 0.f;
          for (int k = mk0[m]; k <= mk1[m]; ++k) acc = fmaf(fbs[k * nm4 + m], mg[k], acc);
          if (a.log_map & 2) acc += op[(long long)m * a.frames];   // partial sums of the earlier bin ranges
          op[(long long)m * a.frames] = (a.log_map & 1) ? mel_log_map(acc) : acc;
        }
      }
    }
    __syncthreads();                                          // staging / mags are reused by the next tile's A
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

}  // namespace mpcg

extern "C" int mpcg_mel_tc_f32(const float* x, float* out, int64_t rows, int64_t t, int n_fft, int hop, int k0, int nbins,
                               int ncols, const void* basis_f16, const float* fb, const float* twq, float inv_norm,
                               int n_mels, int64_t frames, int log_map, void* stream) {
  using namespace mpcg;
  if (rows < 0 || t < 0 || n_fft < 2 || hop < 16 || n_fft % hop != 0 || (hop & 15) != 0 || nbins < 1 || n_mels < 1)
    return MPCG_EINVAL;
  const int Q = n_fft / hop;
  if (Q < 1 || Q > 8 || ncols < 2 * (nbins + 2) || (ncols & 15) != 0 || ncols > 256) return MPCG_EUNSUPPORTED;
  if (frames != 1 + t / hop) return MPCG_EINVAL;
  if (rows == 0 || frames == 0) return MPCG_OK;
  if (!x || !out || !basis_f16 || !fb || !twq) return MPCG_EINVAL;
  if (t <= n_fft / 2) return MPCG_EINVAL;
  const int fpt = kTcRows - (Q - 1);
  const size_t a_bytes = (size_t)kTcRows * hop * 2, b_bytes = (size_t)ncols * hop * 2;
  const size_t stage_bytes = (size_t)kTcRows * (ncols + 1) * 4 + (size_t)kTcRows * ((nbins + 1) | 1) * 4 +
                             2 * (size_t)kTcRows * ((nbins + 2) | 1) * 4;
  if (stage_bytes > 2 * a_bytes) return MPCG_EUNSUPPORTED;
  const size_t smem = 2 * a_bytes + 2 * b_bytes + 16 + 64 + 2 * kTcRows * sizeof(int) +
                      (size_t)nbins * ((n_mels + 3) & ~3) * sizeof(float) + (size_t)(nbins + 2) * Q * 2 * sizeof(float) +
                      2 * (size_t)n_mels * sizeof(int) + 64;
  if (smem > 227 * 1024) return MPCG_EUNSUPPORTED;
  MelTcArgs a;
  a.x = x; a.out = out; a.basis = (const __half*)basis_f16; a.fb = fb; a.twq = twq; a.t = t; a.rows = rows;
  a.n_fft = n_fft; a.hop = hop; a.Q = Q; a.k0 = k0; a.nbins = nbins; a.N = ncols; a.n_mels = n_mels;
  a.frames = (int)frames; a.log_map = log_map; a.inv_norm = inv_norm;
  a.rps = (int)frames + Q - 1;
  a.total_tiles = ((long long)rows * a.rps + fpt - 1) / fpt;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = cudaFuncSetAttribute(mel_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const long long total = a.total_tiles;
  const int grid = (int)(total < sms ? total : sms);
  mel_tc_kernel<<<grid, kTcThreads, smem, (cudaStream_t)stream>>>(a);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}
