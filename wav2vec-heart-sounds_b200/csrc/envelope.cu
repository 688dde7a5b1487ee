// extern "C" entries: mpcg_hilbert_work_bytes, mpcg_hilbert_envelope_f32 -- amplitude envelope of the analytic signal
// (reference signalproc/envelopes.py:11-13: abs(scipy.signal.hilbert(x))).
//
// SciPy builds the analytic signal with an N-point DFT of the whole row (N = the row length, any integer), masks the
// negative frequencies (h[0] = 1, h[1 .. ceil(N/2)-1] = 2, h[N/2] = 1 for even N) and transforms back.  Rows here are
// 8 k .. 500 k samples with arbitrary factorisations, so both N-point transforms run as Bluestein chirp convolutions
// of a power-of-two length M >= 2N - 1 in fp64:
//   X[k] = w[k] * sum_n (x[n] w[n]) conj(w[k - n]),   w[n] = exp(-i pi n^2 / N)
// and because |w| = 1 the chirps between the two transforms cancel: the second convolution's input is simply
// h[k] * conj(c[k]) (c = the first convolution), and the envelope is |second convolution| / (M^2 N).
// A length-M FFT is a four-step transform, M = M1 * M2: column FFTs of length M1 (stride M2), twiddles, row FFTs of
// length M2; spectra stay in the transposed order they come out in, the convolution kernel (the chirp's spectrum, built
// once per call with the same kernels) is stored the same way.  Five launches per batch, each a read + write of the
// [rows, M] complex workspace:
//   columns (x w -> FFT -> twiddle) | rows (FFT * B -> IFFT) | columns (IFFT -> h conj -> FFT) | rows | columns (IFFT -> |.|)
// Shared-memory FFTs with radix-8 register butterflies: decimation in time from a bit-reversed scatter, decimation in frequency into a
// bit-reversed gather, so no separate permutation pass.  flags & MPCG_ENV_LOG writes log(max(envelope, eps)), the input
// of the homomorphic envelope's low-pass (envelopes.py:21-22).
#include "common.cuh"

namespace mpcg {

constexpr int kEvThreads = 512;
constexpr int kEvLo = 1024;                          // two-level root table: exp(-2 pi i j / M) = hi[j >> 10] * lo[j & 1023]

struct EvPlan {
  long long n, m;
  int log1, log2;                                    // M1 = 1 << log1 (columns, stride M2), M2 = 1 << log2 (rows)
  int cols;                                          // columns per CTA in the column kernels
  int rws;                                           // rows per CTA in the row kernels
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }

// Shared-memory index padding (in 16-byte elements): one pad per 8 and per 64 elements, so that strides of 8 and of 64
// elements -- the first merged pass and the bit-reversed scatter / gather -- fall on different banks within a quarter warp.
__host__ __device__ __forceinline__ int ev_pad(int i) { return i + (i >> 3) + (i >> 6); }
__host__ __device__ __forceinline__ int ev_seq_stride(int n) { return ev_pad(n) | 1; }

__device__ __forceinline__ double2 ev_root(const double2* __restrict__ hi, const double2* __restrict__ lo, long long idx) {
  return cmul(hi[idx >> 10], lo[idx & (kEvLo - 1)]);
}

// tw[j] = exp(-2 pi i j / n), j < n / 2, from the root tables of the length-m transform
__device__ __forceinline__ void ev_fill_tw(double2* tw, int logn, const EvPlan& p, const double2* hi, const double2* lo) {
  const int half = 1 << (logn - 1);
  const long long step = p.m >> logn;
  for (int j = threadIdx.x; j < half; j += kEvThreads) tw[ev_pad(j)] = ev_root(hi, lo, (long long)j * step);
}

// In-place FFTs of `nf` sequences of length 1 << logn, sequence f at d + f * stride.
// DIT: input in bit-reversed positions, output in natural order.  DIF: natural in, bit-reversed out.
// Up to three radix-2 stages are merged per pass (a radix-8 butterfly held in registers): shared-memory traffic, not the
// fp64 pipe, bounds a stage-per-pass transform.  Twiddles of a merged group: stage s+u needs
// exp(-2 pi i (p + k 2^s) / 2^(s+u+1)) = base_u * (8th root of unity)^(k << (2-u)), base_u = omega_(s+u+1)^p, and
// base_u = base_(u+1)^2 -- one table load per group.
template <int E>
__device__ __forceinline__ double2 ev_mul_root8(double2 z) {       // z * exp(-2 pi i E / 8)
  constexpr double h = 0.70710678118654752440;
  if (E == 0) return z;
  if (E == 1) return make_double2((z.x + z.y) * h, (z.y - z.x) * h);
  if (E == 2) return make_double2(z.y, -z.x);
  return make_double2((z.y - z.x) * h, -(z.x + z.y) * h);            // E == 3
}

template <int Q, bool DIT, int U, int M>
struct EvStage {                                                     // butterflies of merged stage U on local indices >= M
  static __device__ __forceinline__ void run(double2 (&r)[1 << Q], const double2 (&base)[Q]) {
    if constexpr (M < (1 << Q)) {
      if constexpr ((M & (1 << U)) == 0) {
        constexpr int K = M & ((1 << U) - 1);
        const double2 w = ev_mul_root8<(K << (2 - U)) & 3>(base[U]);
        if constexpr (DIT) {
          const double2 a = r[M], b = cmul(r[M + (1 << U)], w);
          r[M] = make_double2(a.x + b.x, a.y + b.y);
          r[M + (1 << U)] = make_double2(a.x - b.x, a.y - b.y);
        } else {
          const double2 a = r[M], b = r[M + (1 << U)];
          r[M] = make_double2(a.x + b.x, a.y + b.y);
          r[M + (1 << U)] = cmul(make_double2(a.x - b.x, a.y - b.y), w);
        }
      }
      EvStage<Q, DIT, U, M + 1>::run(r, base);
    }
  }
};

// One pass over stages s .. s+Q-1 of every sequence.
template <int Q, bool DIT>
__device__ __forceinline__ void ev_pass(double2* d, int logn, int nf, int stride, const double2* tw, int s) {
  constexpr int R = 1 << Q;
  const int hl = logn - 1, lg = logn - Q;                            // groups per sequence = 1 << lg
  for (int b = threadIdx.x; b < nf << lg; b += kEvThreads) {
    const int f = b >> lg, j = b & ((1 << lg) - 1);
    const int p = j & ((1 << s) - 1);
    double2* q = d + f * stride;
    const int i0 = ((j >> s) << (s + Q)) + p;
    double2 r[R], base[Q];
#pragma unroll
    for (int m = 0; m < R; ++m) r[m] = q[ev_pad(i0 + (m << s))];
    base[Q - 1] = tw[ev_pad(p << (hl - (s + Q - 1)))];
#pragma unroll
    for (int u = Q - 2; u >= 0; --u) base[u] = cmul(base[u + 1], base[u + 1]);
    if constexpr (DIT) {
      EvStage<Q, true, 0, 0>::run(r, base);
      if constexpr (Q > 1) EvStage<Q, true, 1, 0>::run(r, base);
      if constexpr (Q > 2) EvStage<Q, true, 2, 0>::run(r, base);
    } else {
      if constexpr (Q > 2) EvStage<Q, false, 2, 0>::run(r, base);
      if constexpr (Q > 1) EvStage<Q, false, 1, 0>::run(r, base);
      EvStage<Q, false, 0, 0>::run(r, base);
    }
#pragma unroll
    for (int m = 0; m < R; ++m) q[ev_pad(i0 + (m << s))] = r[m];
  }
  __syncthreads();
}

__device__ __forceinline__ void ev_fft_dit(double2* d, int logn, int nf, int stride, const double2* tw) {
  int s = 0;
  for (; s + 3 <= logn; s += 3) ev_pass<3, true>(d, logn, nf, stride, tw, s);
  if (logn - s == 2) ev_pass<2, true>(d, logn, nf, stride, tw, s);
  else if (logn - s == 1) ev_pass<1, true>(d, logn, nf, stride, tw, s);
}
__device__ __forceinline__ void ev_fft_dif(double2* d, int logn, int nf, int stride, const double2* tw) {
  // the mirror image: the passes of the DIT transform in reverse order, each run backwards
  const int rem = logn % 3, full = logn - rem;
  if (rem == 2) ev_pass<2, false>(d, logn, nf, stride, tw, full);
  else if (rem == 1) ev_pass<1, false>(d, logn, nf, stride, tw, full);
  for (int s = full - 3; s >= 0; s -= 3) ev_pass<3, false>(d, logn, nf, stride, tw, s);
}
__device__ __forceinline__ int ev_rev(int i, int logn) { return (int)(__brev((unsigned)i) >> (32 - logn)); }

// chirp w[n] = exp(-i pi n^2 / N) (n < N), root tables of the length-M transform, and the convolution kernel
// b[j] = conj(w[|j|]) for |j| < N wrapped to length M (zero elsewhere)
__global__ void __launch_bounds__(256)
ev_setup_kernel(EvPlan p, double2* __restrict__ chirp, double2* __restrict__ hi, double2* __restrict__ lo, double2* __restrict__ bk) {
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x, gsz = (long long)gridDim.x * 256;
  for (long long j = gid; j < p.m; j += gsz) {
    const long long a = j < p.n ? j : (j > p.m - p.n ? p.m - j : -1);
    double2 v = make_double2(0.0, 0.0);
    if (a >= 0) {
      const unsigned long long r = ((unsigned long long)a * (unsigned long long)a) % (unsigned long long)(2 * p.n);
      double sn, cs;
      sincospi((double)r / (double)p.n, &sn, &cs);
      v = make_double2(cs, sn);                                     // conj(w[a])
      if (j < p.n) chirp[j] = make_double2(cs, -sn);
    }
    bk[j] = v;
  }
  const long long nhi = p.m > kEvLo ? p.m >> 10 : 1;
  for (long long j = gid; j < nhi + kEvLo; j += gsz) {
    const long long idx = j < nhi ? j << 10 : j - nhi;               // exponent of the root, out of m
    double sn, cs;
    sincospi(2.0 * (double)idx / (double)p.m, &sn, &cs);
    const double2 v = idx < p.m ? make_double2(cs, -sn) : make_double2(1.0, 0.0);
    if (j < nhi) hi[j] = v; else lo[j - nhi] = v;
  }
}

enum { EV_COL_PLAIN = 0, EV_COL_FIRST = 1, EV_COL_MID = 2, EV_COL_LAST = 3 };

// Column transforms: CTA (blockIdx.x, blockIdx.y) owns columns n2 = blockIdx.x * cols .. + cols of row blockIdx.y.
template <int MODE>
__global__ void __launch_bounds__(kEvThreads, 2)
ev_col_kernel(const __grid_constant__ EvPlan p, const float* __restrict__ x, float* __restrict__ y, double2* __restrict__ u,
              const double2* __restrict__ chirp, const double2* __restrict__ hi, const double2* __restrict__ lo, int flags) {
  extern __shared__ __align__(16) unsigned char ev_raw[];
  double2* tw = reinterpret_cast<double2*>(ev_raw);
  const int m1 = 1 << p.log1, m2 = 1 << p.log2;
  double2* d = tw + ev_pad(m1 >> 1) + 1;
  const int stride = ev_seq_stride(m1);
  const long long row = blockIdx.y;
  const int c0 = blockIdx.x * p.cols;
  double2* ur = u + row * p.m;
  const int lc = 31 - __clz(p.cols);
  ev_fill_tw(tw, p.log1, p, hi, lo);
  // load (scatter to bit-reversed positions for the decimation-in-time transform)
  for (int e = threadIdx.x; e < m1 << lc; e += kEvThreads) {
    const int n1 = e >> lc, c = e & (p.cols - 1);
    const long long n = (long long)n1 * m2 + c0 + c;
    double2 v;
    if (MODE == EV_COL_FIRST) {
      v = make_double2(0.0, 0.0);
      if (n < p.n) { const double s = (double)x[row * p.n + n]; const double2 w = chirp[n]; v = make_double2(s * w.x, s * w.y); }
    } else if (MODE == EV_COL_PLAIN) {
      v = ur[n];
    } else {
      // inverse transform: conj(twiddle) on the way in, and IFFT(z) = conj(FFT(conj(z)))
      const double2 t = cmul(ur[n], cconj(ev_root(hi, lo, ((long long)n1 * (c0 + c)) & (p.m - 1))));
      v = cconj(t);
    }
    d[c * stride + ev_pad(ev_rev(n1, p.log1))] = v;
  }
  __syncthreads();
  ev_fft_dit(d, p.log1, p.cols, stride, tw);
  if (MODE == EV_COL_LAST) {
    const double scale = 1.0 / ((double)p.m * (double)p.m * (double)p.n);
    for (int e = threadIdx.x; e < m1 << lc; e += kEvThreads) {
      const int n1 = e >> lc, c = e & (p.cols - 1);
      const long long n = (long long)n1 * m2 + c0 + c;
      if (n < p.n) {
        const double2 v = d[c * stride + ev_pad(n1)];
        double env = sqrt(v.x * v.x + v.y * v.y) * scale;
        if (flags & MPCG_ENV_LOG) env = log(fmax(env, 2.220446049250313e-16));
        y[row * p.n + n] = (float)env;
      }
    }
    return;
  }
  if (MODE == EV_COL_MID) {
    // d holds conj(c[n]) (the conjugate of the inverse transform was never taken): mask, then transform forward again
    const long long nyq = (p.n & 1) ? -1 : p.n >> 1;
    const long long top = (p.n + 1) >> 1;                             // h = 2 for 0 < n < top
    for (int e = threadIdx.x; e < m1 << lc; e += kEvThreads) {
      const int n1 = e >> lc, c = e & (p.cols - 1);
      const long long n = (long long)n1 * m2 + c0 + c;
      const double h = n == 0 || n == nyq ? 1.0 : (n < top ? 2.0 : 0.0);
      double2& v = d[c * stride + ev_pad(n1)];
      v = make_double2(v.x * h, v.y * h);
    }
    __syncthreads();
    ev_fft_dif(d, p.log1, p.cols, stride, tw);
  }
  // store with the forward twiddle exp(-2 pi i n2 k1 / M); after the DIF transform element k1 sits at rev(k1)
  for (int e = threadIdx.x; e < m1 << lc; e += kEvThreads) {
    const int k1 = e >> lc, c = e & (p.cols - 1);
    const double2 v = d[c * stride + ev_pad(MODE == EV_COL_MID ? ev_rev(k1, p.log1) : k1)];
    ur[(long long)k1 * m2 + c0 + c] = cmul(v, ev_root(hi, lo, ((long long)k1 * (c0 + c)) & (p.m - 1)));
  }
}

// Row transforms: CTA (blockIdx.x, blockIdx.y) owns rows k1 = blockIdx.x * rws .. + rws of signal blockIdx.y.
// conv != 0: FFT, times the kernel spectrum, inverse FFT (unnormalised); conv == 0: FFT only (building the kernel spectrum).
__global__ void __launch_bounds__(kEvThreads, 2)
ev_row_kernel(const __grid_constant__ EvPlan p, double2* __restrict__ u, const double2* __restrict__ bk,
              const double2* __restrict__ hi, const double2* __restrict__ lo, int conv) {
  extern __shared__ __align__(16) unsigned char ev_raw[];
  double2* tw = reinterpret_cast<double2*>(ev_raw);
  const int m2 = 1 << p.log2;
  double2* d = tw + ev_pad(m2 >> 1) + 1;
  const int stride = ev_seq_stride(m2);
  const long long base = (long long)blockIdx.x * p.rws * m2;
  double2* ur = u + (long long)blockIdx.y * p.m + base;
  ev_fill_tw(tw, p.log2, p, hi, lo);
  for (int e = threadIdx.x; e < p.rws << p.log2; e += kEvThreads) {
    const int r = e >> p.log2, i = e & (m2 - 1);
    d[r * stride + ev_pad(ev_rev(i, p.log2))] = ur[e];
  }
  __syncthreads();
  ev_fft_dit(d, p.log2, p.rws, stride, tw);
  if (!conv) {
    for (int e = threadIdx.x; e < p.rws << p.log2; e += kEvThreads) ur[e] = d[(e >> p.log2) * stride + ev_pad(e & (m2 - 1))];
    return;
  }
  for (int e = threadIdx.x; e < p.rws << p.log2; e += kEvThreads) {
    double2& v = d[(e >> p.log2) * stride + ev_pad(e & (m2 - 1))];
    v = cconj(cmul(v, bk[base + e]));
  }
  __syncthreads();
  ev_fft_dif(d, p.log2, p.rws, stride, tw);
  for (int e = threadIdx.x; e < p.rws << p.log2; e += kEvThreads) {
    const int r = e >> p.log2, i = e & (m2 - 1);
    ur[e] = cconj(d[r * stride + ev_pad(ev_rev(i, p.log2))]);
  }
}

static bool ev_make_plan(long long n, EvPlan* p) {
  if (n < 1 || n > (1LL << 19)) return false;
  int lg = 8;                                                       // at least 16 x 16
  while ((1LL << lg) < 2 * n - 1) ++lg;
  p->n = n;
  p->m = 1LL << lg;
  p->log1 = lg / 2;
  p->log2 = lg - p->log1;
  // 4096 points per CTA (64 KB of fp64 pairs), at least 4 adjacent columns (64-byte runs) per column transform
  p->cols = 1 << (p->log1 >= 10 ? 2 : (12 - p->log1 < p->log2 ? 12 - p->log1 : p->log2));
  p->rws = 1 << (12 - p->log2 < p->log1 ? 12 - p->log2 : p->log1);
  return true;
}
static size_t ev_fixed_bytes(const EvPlan& p) {                     // chirp + root tables + kernel spectrum
  const long long nhi = p.m > kEvLo ? p.m >> 10 : 1;
  return (size_t)(p.n + nhi + kEvLo + p.m) * sizeof(double2);
}

}  // namespace mpcg

extern "C" int64_t mpcg_hilbert_work_bytes(int64_t rows, int64_t t) {
  mpcg::EvPlan p;
  if (rows < 0 || !mpcg::ev_make_plan(t, &p)) return -1;
  return (int64_t)(mpcg::ev_fixed_bytes(p) + (size_t)rows * (size_t)p.m * sizeof(double2));
}

extern "C" int mpcg_hilbert_envelope_f32(const float* x, float* y, void* work, int64_t work_bytes, int64_t rows, int64_t t,
                                         int flags, void* stream) {
  using namespace mpcg;
  if (rows < 0 || t < 0) return MPCG_EINVAL;
  if (rows == 0 || t == 0) return MPCG_OK;
  EvPlan p;
  if (!ev_make_plan(t, &p)) return MPCG_ERANGE;
  if (!x || !y || !work) return MPCG_EINVAL;
  if (rows > 65535) return MPCG_ERANGE;                              // gridDim.y; the host side chunks long before this
  if (work_bytes < mpcg_hilbert_work_bytes(rows, t)) return MPCG_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const long long nhi = p.m > kEvLo ? p.m >> 10 : 1;
  double2* chirp = reinterpret_cast<double2*>(work);
  double2* hi = chirp + p.n;
  double2* lo = hi + nhi;
  double2* bk = lo + kEvLo;
  double2* u = bk + p.m;
  const int m1 = 1 << p.log1, m2 = 1 << p.log2;
  const size_t col_smem = ((size_t)ev_pad(m1 >> 1) + 1 + (size_t)p.cols * ev_seq_stride(m1)) * sizeof(double2);
  const size_t row_smem = ((size_t)ev_pad(m2 >> 1) + 1 + (size_t)p.rws * ev_seq_stride(m2)) * sizeof(double2);
  cudaError_t e;
#define EV_ATTR(k, bytes)                                                                      \
  e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));       \
  if (e != cudaSuccess) return (int)e;
  EV_ATTR(ev_col_kernel<EV_COL_PLAIN>, col_smem)
  EV_ATTR(ev_col_kernel<EV_COL_FIRST>, col_smem)
  EV_ATTR(ev_col_kernel<EV_COL_MID>, col_smem)
  EV_ATTR(ev_col_kernel<EV_COL_LAST>, col_smem)
  EV_ATTR(ev_row_kernel, row_smem)
#undef EV_ATTR
  const dim3 gcol1((unsigned)(m2 / p.cols), 1), grow1((unsigned)(m1 / p.rws), 1);
  const dim3 gcol((unsigned)(m2 / p.cols), (unsigned)rows), grow((unsigned)(m1 / p.rws), (unsigned)rows);
  ev_setup_kernel<<<148 * 4, 256, 0, st>>>(p, chirp, hi, lo, bk);
  MPCG_LAUNCH_CHECK();
  ev_col_kernel<EV_COL_PLAIN><<<gcol1, kEvThreads, col_smem, st>>>(p, nullptr, nullptr, bk, chirp, hi, lo, flags);
  MPCG_LAUNCH_CHECK();
  ev_row_kernel<<<grow1, kEvThreads, row_smem, st>>>(p, bk, nullptr, hi, lo, 0);
  MPCG_LAUNCH_CHECK();
  ev_col_kernel<EV_COL_FIRST><<<gcol, kEvThreads, col_smem, st>>>(p, x, y, u, chirp, hi, lo, flags);
  MPCG_LAUNCH_CHECK();
  ev_row_kernel<<<grow, kEvThreads, row_smem, st>>>(p, u, bk, hi, lo, 1);
  MPCG_LAUNCH_CHECK();
  ev_col_kernel<EV_COL_MID><<<gcol, kEvThreads, col_smem, st>>>(p, x, y, u, chirp, hi, lo, flags);
  MPCG_LAUNCH_CHECK();
  ev_row_kernel<<<grow, kEvThreads, row_smem, st>>>(p, u, bk, hi, lo, 1);
  MPCG_LAUNCH_CHECK();
  ev_col_kernel<EV_COL_LAST><<<gcol, kEvThreads, col_smem, st>>>(p, x, y, u, chirp, hi, lo, flags);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}
