// extern "C" entry: mpcg_mel_dm_f32 -- the exact (float64) mel tier on the fp64 tensor path.
// Same contract as mpcg_mel_f32 with basis_f64 = 1 (signalproc/spectrogram.py:13-45: framing, windowed DFT of the bins
// that carry mel weight, magnitude, HTK mel projection [, dB map]), within 1e-5 of the float64 reference on any input.
//
// Framing + windowed DFT is a GEMM  [frames x n] . [n x 2 bins]  in float64.  mel_kernel<double> issues it as scalar
// DFMAs whose operands all come through the load/store unit (a broadcast sample and a float -> double conversion per
// two DFMAs): 28 % of the DFMA peak.  Here it runs as  mma.sync.m8n8k4.f64  (DMMA: the same fp64 pipe, but one
// instruction per 256 FMAs and operands that are loaded once per fragment): a warp owns 16 frames x 16 bins (two by
// four 8x8 accumulator tiles, cos / -sin interleaved so that a thread holds re and im of its bins); per four samples
// it loads two A fragments (converted once) and four B fragments for eight DMMAs.  Measured at configs[3]
// (8192 x 64000 @16 kHz): 10.4 ms against 26.2 ms, 80 % of the DMMA rate; warp tiles 8x32 / 8x16 / 16x32 / 32x16:
// 11.2 / 11.3 / 14.4 / 14.7 ms; a four-slot ring (two CTAs per SM instead of three): 14.8 ms.
//   * samples: the CTA's 32 frames overlap; their span is staged once, float32, as rows of `hop` samples with a row
//     stride = 4 (mod 32) words, so the A fragment's eight frames x four samples hit 32 different banks;
//   * basis: prepared by the host in the fragment order, [bin block][slice of 16 samples][64 columns][16 + 4 pad]
//     float64, so that a slice is ONE bulk asynchronous copy (cp.async.bulk + mbarrier) into a two-slot ring and the
//     B fragments (column stride 20 doubles = 8 banks) load without conflicts;
//   * magnitudes go to shared memory, the mel projection and the dB map follow as in mel.cu.
#include "common.cuh"

namespace mpcg {

#ifndef MPCG_DM_MT
#define MPCG_DM_MT 2                  // 8-frame tiles per warp
#endif
#ifndef MPCG_DM_NT
#define MPCG_DM_NT 4                  // 4-bin tiles per warp
#endif
constexpr int kDmMT = MPCG_DM_MT, kDmNT = MPCG_DM_NT;
#ifndef MPCG_DM_FT
#define MPCG_DM_FT 32
#endif
constexpr int kDmFT = MPCG_DM_FT;                // frames per CTA (x 32 bins per pass over the samples)
constexpr int kDmWarps = (kDmFT / 8 / kDmMT) * (8 / kDmNT);
constexpr int kDmThreads = 32 * kDmWarps;
#ifndef MPCG_DM_SLOTS
#define MPCG_DM_SLOTS 2
#endif
constexpr int kDmSlots = MPCG_DM_SLOTS;          // ring of basis slices (a power of two)
constexpr int kDmSliceK = 16;                    // samples per slice
constexpr int kDmColStride = 20;                 // doubles per basis column in a slice (16 + 4 pad)
constexpr int kDmSliceDoubles = 64 * kDmColStride;
constexpr int kDmSliceBytes = kDmSliceDoubles * 8;

struct DmArgs {
  const float* x;         // [rows, t]
  float* out;             // [rows, n_mels, frames]
  const double* basis;    // [kpad / 32][nslices][64][20]
  const float* fb;        // [nbins][n_mels]
  const int* mel_range;   // [n_mels][2]: the bins [lo, hi) a mel filter has weight on (triangles: a contiguous run)
  long long t;
  int n_fft, hop, n_lo, nn16, nbins, kpad, n_mels, frames, log_map;
  int segw, nsegs, nslices;
};

__device__ __forceinline__ float dm_log_map(float mel) {
  const float db = 20.f * log10f(fmaxf(mel, 1e-5f)) - 20.f;
  return fminf(fmaxf((db + 100.f) / 100.f, 0.f), 1.f);
}

__global__ void __launch_bounds__(kDmThreads)
mel_dm_kernel(const DmArgs a) {
  extern __shared__ __align__(16) unsigned char dm_raw[];
  double* ring = reinterpret_cast<double*>(dm_raw);                                  // [slots][64][20]
  unsigned long long* full = reinterpret_cast<unsigned long long*>(dm_raw + kDmSlots * kDmSliceBytes);
  float* xs = reinterpret_cast<float*>(dm_raw + kDmSlots * kDmSliceBytes + 64);      // [nsegs][segw]
  float* mag = xs + a.nsegs * a.segw;                                                // [kDmFT][kpad + 1]
  const int ms = a.kpad + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
  const long long row = blockIdx.y;
  const int f0 = blockIdx.x * kDmFT;
  const float* xr = a.x + row * a.t;
  const int nblocks = a.kpad >> 5;
  const int total = nblocks * a.nslices;

  auto issue = [&](int item) {                                                       // thread 0 only
    const int slot = item & (kDmSlots - 1);
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(full + slot);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)kDmSliceBytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(ring + slot * kDmSliceDoubles)),
                 "l"(a.basis + (long long)item * kDmSliceDoubles), "r"((uint32_t)kDmSliceBytes), "r"(bar)
                 : "memory");
  };
  if (tid == 0) {
    for (int s = 0; s < kDmSlots; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(full + s)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int it = 0; it < kDmSlots && it < total; ++it) issue(it);
  }
  // stage the span of the CTA's frames: un-padded index = f0*hop + n_lo - n_fft/2 + m, reflected at both ends
  {
    const long long base = (long long)f0 * a.hop + a.n_lo - a.n_fft / 2;
    for (int seg = warp; seg < a.nsegs; seg += kDmWarps) {
      const long long j0 = base + (long long)seg * a.hop;
      float* dst = xs + seg * a.segw;
      if (j0 >= 0 && j0 + a.hop <= a.t && (a.hop & 3) == 0 && (a.segw & 3) == 0 &&
          ((reinterpret_cast<uintptr_t>(xr + j0)) & 15u) == 0) {            // interior row: 128-bit loads and stores
        const float4* src4 = reinterpret_cast<const float4*>(xr + j0);
        float4* dst4 = reinterpret_cast<float4*>(dst);
        for (int q = lane; q < (a.hop >> 2); q += 32) dst4[q] = __ldg(src4 + q);
      } else {
        for (int jj = lane; jj < a.hop; jj += 32) {
          long long j = j0 + jj;
          if (j < 0) j = -j;
          if (j >= a.t) j = 2 * (a.t - 1) - j;
          dst[jj] = (j >= 0 && j < a.t) ? __ldg(xr + j) : 0.f;
        }
      }
    }
  }
  __syncthreads();

  // warp -> (frame tiles, bin tiles): wf-th group of kDmMT frame tiles, wb-th group of kDmNT bin tiles
  const int wf = warp / (8 / kDmNT), wb = warp - wf * (8 / kDmNT);
  const int fl = wf * 8 * kDmMT + g;                                                 // my frame of the first A fragment
  double acc[kDmMT][kDmNT][2];
  for (int it = 0; it < total; ++it) {
    const int slot = it & (kDmSlots - 1);
    const int sl = it % a.nslices;
    if (sl == 0) {
#pragma unroll
      for (int i = 0; i < kDmMT; ++i)
#pragma unroll
        for (int j = 0; j < kDmNT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    }
    {
      const uint32_t bar = (uint32_t)__cvta_generic_to_shared(full + slot);
      const uint32_t parity = (uint32_t)((it / kDmSlots) & 1);
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "DM_WAIT:\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
          "@p bra DM_DONE;\n\t"
          "bra DM_WAIT;\n\t"
          "DM_DONE:\n\t"
          "}\n" ::"r"(bar), "r"(parity)
          : "memory");
    }
    const double* B = ring + slot * kDmSliceDoubles + (wb * kDmNT * 8 + g) * kDmColStride + tq;
    int nn = sl * kDmSliceK + tq;
    int seg = nn / a.hop;
    int jj = nn - seg * a.hop;
    const float* xf = xs + fl * a.segw;
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      double av[kDmMT];
#pragma unroll
      for (int i = 0; i < kDmMT; ++i) av[i] = (double)xf[(seg + 8 * i) * a.segw + jj];
      jj += 4;
      if (jj >= a.hop) { jj -= a.hop; ++seg; }
#pragma unroll
      for (int j = 0; j < kDmNT; ++j) {
        const double bv = B[j * 8 * kDmColStride + k4 * 4];
#pragma unroll
        for (int i = 0; i < kDmMT; ++i)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                       : "d"(av[i]), "d"(bv));
      }
    }
    __syncthreads();                                                                 // the slot is free again
    if (tid == 0 && it + kDmSlots < total) issue(it + kDmSlots);
    if (sl == a.nslices - 1) {                                                       // |X| of this bin block
      const int bb = it / a.nslices;
#pragma unroll
      for (int i = 0; i < kDmMT; ++i)
#pragma unroll
        for (int j = 0; j < kDmNT; ++j)
          mag[(fl + 8 * i) * ms + bb * 32 + (wb * kDmNT + j) * 4 + tq] =
              (float)sqrt(acc[i][j][0] * acc[i][j][0] + acc[i][j][1] * acc[i][j][1]);
    }
  }
  __syncthreads();
  // mel projection: out[m, frame] = sum_k fb[k, m] * mag[frame, k]; lanes run over frames so stores are contiguous
  for (int o = tid; o < a.n_mels * kDmFT; o += kDmThreads) {
    const int m = o / kDmFT, f = o - m * kDmFT;
    if (f0 + f >= a.frames) continue;
    const float* mg = mag + f * ms;
    float s = 0.f;
    const int k_hi = __ldg(a.mel_range + 2 * m + 1);
    for (int k = __ldg(a.mel_range + 2 * m); k < k_hi; ++k) s = fmaf(__ldg(a.fb + (long long)k * a.n_mels + m), mg[k], s);
    a.out[(row * a.n_mels + m) * a.frames + f0 + f] = a.log_map ? dm_log_map(s) : s;
  }
}

}  // namespace mpcg

extern "C" int mpcg_mel_dm_f32(const float* x, float* out, int64_t rows, int64_t t, int n_fft, int hop, int n_lo, int n_hi,
                               int nbins, int kpad, const double* basis_dm, const float* fb, const int* mel_range, int n_mels,
                               int64_t frames, int log_map, void* stream) {
  using namespace mpcg;
  if (rows < 0 || t < 0 || n_fft < 2 || hop < 1 || n_lo < 0 || n_hi <= n_lo || n_hi > n_fft || nbins < 1 ||
      kpad < nbins || (kpad & 31) != 0 || n_mels < 1 || frames < 0)
    return MPCG_EINVAL;
  if (frames != 1 + t / hop) return MPCG_EINVAL;
  if (rows == 0 || frames == 0) return MPCG_OK;
  if (!x || !out || !basis_dm || !fb || !mel_range) return MPCG_EINVAL;
  if (t <= n_fft / 2) return MPCG_EINVAL;                    // reflect padding needs pad < length (torch raises too)
  if (rows > 65535 || frames > 0x3fffffffLL) return MPCG_ERANGE;
  if (hop < 4 || ((uintptr_t)basis_dm & 15u)) return MPCG_EUNSUPPORTED;
  DmArgs a;
  a.x = x; a.out = out; a.basis = basis_dm; a.fb = fb; a.mel_range = mel_range; a.t = t; a.n_fft = n_fft; a.hop = hop; a.n_lo = n_lo;
  a.nn16 = (n_hi - n_lo + kDmSliceK - 1) / kDmSliceK * kDmSliceK;
  a.nbins = nbins; a.kpad = kpad; a.n_mels = n_mels; a.frames = (int)frames; a.log_map = log_map;
  a.nslices = a.nn16 / kDmSliceK;
  const int span = (kDmFT - 1) * hop + a.nn16 + 4;           // (+4: the last A fragment's lanes may look one group ahead)
  a.nsegs = (span + hop - 1) / hop;
  a.segw = hop + ((4 - hop % 32) + 32) % 32;                 // row stride = 4 (mod 32) words
  const size_t smem = (size_t)kDmSlots * kDmSliceBytes + 64 + ((size_t)a.nsegs * a.segw + (size_t)kDmFT * (kpad + 1)) * sizeof(float);
  if (smem > 200 * 1024) return MPCG_EUNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(mel_dm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((unsigned)((frames + kDmFT - 1) / kDmFT), (unsigned)rows);
  mel_dm_kernel<<<grid, kDmThreads, smem, (cudaStream_t)stream>>>(a);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}
