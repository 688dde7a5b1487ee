// extern "C" entry: mpcg_biquad_cascade_f32  (see include/mpcg_b200.h)
#include "biquad.cuh"
#include <string.h>
#include <math.h>

namespace mpcg {

// One time step of a two-section group in transposed direct form II.
static inline double group_step(const double c[2][5], double z[4], double x) {
  const double y0 = c[0][0] * x + z[0];
  z[0] = c[0][1] * x - c[0][3] * y0 + z[1];
  z[1] = c[0][2] * x - c[0][4] * y0;
  const double y1 = c[1][0] * y0 + z[2];
  z[2] = c[1][1] * y0 - c[1][3] * y1 + z[3];
  z[3] = c[1][2] * y0 - c[1][4] * y1;
  return y1;
}

void bq_mat_mul(const double* a, const double* b, double* out) {
  double t[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += a[r * 4 + k] * b[k * 4 + c];
      t[r * 4 + c] = s;
    }
  memcpy(out, t, sizeof(t));
}

void bq_group_step(const double c[2][5], double z[4], double x) { group_step(c, z, x); }

void bq_group_coeffs(const double* sos, int n_sections, int first, double c[2][5], bool* ok) {
  *ok = true;
  for (int s = 0; s < 2; ++s) {
    for (int k = 0; k < 5; ++k) c[s][k] = 0.0;
    const int idx = first + s;
    if (idx < n_sections) {
      const double* r = sos + 6 * idx;
      const double a0 = r[3];
      if (a0 == 0.0 || !isfinite(a0)) { *ok = false; return; }
      c[s][0] = r[0] / a0; c[s][1] = r[1] / a0; c[s][2] = r[2] / a0;
      c[s][3] = r[4] / a0; c[s][4] = r[5] / a0;
    } else {
      c[s][0] = 1.0;                           // pass-through padding section
    }
  }
}

void bq_group_AB(const double c[2][5], double A[16], double B[4]) {
  for (int col = 0; col < 4; ++col) {
    double z[4] = {0, 0, 0, 0};
    z[col] = 1.0;
    group_step(c, z, 0.0);
    for (int r = 0; r < 4; ++r) A[r * 4 + col] = z[r];
  }
  double z[4] = {0, 0, 0, 0};
  group_step(c, z, 1.0);
  for (int r = 0; r < 4; ++r) B[r] = z[r];
}

void bq_mat_pow(const double A[16], long long n, double out[16]) {
  double acc[16], base[16];
  for (int i = 0; i < 16; ++i) { acc[i] = (i % 5 == 0) ? 1.0 : 0.0; base[i] = A[i]; }
  while (n > 0) {
    if (n & 1) bq_mat_mul(base, acc, acc);
    bq_mat_mul(base, base, base);
    n >>= 1;
  }
  for (int i = 0; i < 16; ++i) out[i] = acc[i];
}

int bq_make_plan(const double* sos, int n_sections, BqPlan* plan) {
  if (!sos || !plan || n_sections < 1) return MPCG_EINVAL;
  if (n_sections > 2 * kBqMaxGroups) return MPCG_ERANGE;
  memset(plan, 0, sizeof(*plan));
  plan->ngroups = (n_sections + 1) / 2;
  for (int g = 0; g < plan->ngroups; ++g) {
    BqGroup& G = plan->g[g];
    for (int s = 0; s < 2; ++s) {
      const int idx = 2 * g + s;
      if (idx < n_sections) {
        const double* r = sos + 6 * idx;
        const double a0 = r[3];
        if (a0 == 0.0 || !isfinite(a0)) return MPCG_EINVAL;
        G.c[s][0] = r[0] / a0; G.c[s][1] = r[1] / a0; G.c[s][2] = r[2] / a0;
        G.c[s][3] = r[4] / a0; G.c[s][4] = r[5] / a0;
      } else {                               // pad an odd cascade with a pass-through section
        G.c[s][0] = 1.0;
      }
    }
    // A (one step, zero input) column by column, and B (unit input from rest).
    double A[16], B[4];
    for (int col = 0; col < 4; ++col) {
      double z[4] = {0, 0, 0, 0};
      z[col] = 1.0;
      group_step(G.c, z, 0.0);
      for (int r = 0; r < 4; ++r) A[r * 4 + col] = z[r];
    }
    {
      double z[4] = {0, 0, 0, 0};
      group_step(G.c, z, 1.0);
      for (int r = 0; r < 4; ++r) B[r] = z[r];
    }
    // wt[j] = A^(L-1-j) B : walk B forward under zero input.
    double v[4] = {B[0], B[1], B[2], B[3]};
    for (int j = kBqL - 1; j >= 0; --j) {
      for (int s = 0; s < 4; ++s) G.wt[j][s] = v[s];
      group_step(G.c, v, 0.0);
    }
    // M = A^L, then squarings.
    double M[16];
    for (int i = 0; i < 16; ++i) M[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int k = 0; k < kBqL; ++k) bq_mat_mul(A, M, M);
    memcpy(G.mp[0], M, sizeof(M));
    for (int d = 1; d < 6; ++d) bq_mat_mul(G.mp[d - 1], G.mp[d - 1], G.mp[d]);
  }
  return MPCG_OK;
}

constexpr int kBqThreads = 256;
constexpr int kBqTile = kBqThreads * kBqL;

struct BqSmem {
  BqScratch<kBqThreads> sc;
  float tile[kBqTile + 8];
};

__global__ void __launch_bounds__(kBqThreads)
biquad_rows_kernel(const float* __restrict__ x, float* __restrict__ y, long long t,
                   const float* __restrict__ row_mask, const __grid_constant__ BqPlan plan) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BqSmem& sm = *reinterpret_cast<BqSmem*>(smem_raw);
  const long long row = blockIdx.x;
  if (row_mask && row_mask[row] == 0.f) return;              // rows not selected are left untouched in y
  const float* xr = x + row * t;
  float* yr = y + row * t;
  bq_init_scratch<kBqThreads>(sm.sc, plan);
  __syncthreads();
  for (long long t0 = 0; t0 < t; t0 += kBqTile) {
    const int n = (int)((t - t0) < (long long)kBqTile ? (t - t0) : (long long)kBqTile);
    const float* src = xr + t0;
    float* dst = yr + t0;
    float* sh = sm.tile + phase_of(src);
    copy_g2s<kBqThreads>(sh, src, n);
    for (int i = n + threadIdx.x; i < kBqTile; i += kBqThreads) sh[i] = 0.f;
    __syncthreads();
    bq_filter_tile<kBqThreads>(sh, sm.sc, plan);
    __syncthreads();
    if (phase_of(dst) == phase_of(src)) {
      copy_s2g<kBqThreads>(dst, sh, n);
    } else {
      for (int i = threadIdx.x; i < n; i += kBqThreads) dst[i] = sh[i];
    }
    __syncthreads();
  }
}

}  // namespace mpcg

extern "C" int mpcg_biquad_cascade_masked_f32(const float* x, float* y, int64_t rows, int64_t t, const double* sos,
                                              int n_sections, const float* row_mask, void* stream);

extern "C" int mpcg_biquad_cascade_f32(const float* x, float* y, int64_t rows, int64_t t, const double* sos,
                                       int n_sections, void* stream) {
  return mpcg_biquad_cascade_masked_f32(x, y, rows, t, sos, n_sections, nullptr, stream);
}

extern "C" int mpcg_biquad_cascade_masked_f32(const float* x, float* y, int64_t rows, int64_t t, const double* sos,
                                              int n_sections, const float* row_mask, void* stream) {
  using namespace mpcg;
  if (rows < 0 || t < 0) return MPCG_EINVAL;
  BqPlan plan;
  const int rc = bq_make_plan(sos, n_sections, &plan);
  if (rc != MPCG_OK) return rc;
  if (rows == 0 || t == 0) return MPCG_OK;
  if (!x || !y) return MPCG_EINVAL;
  if (rows > 0x7fffffffLL) return MPCG_ERANGE;
  cudaError_t e = cudaFuncSetAttribute(biquad_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(BqSmem));
  if (e != cudaSuccess) return (int)e;
  biquad_rows_kernel<<<(unsigned)rows, kBqThreads, sizeof(BqSmem), (cudaStream_t)stream>>>(x, y, (long long)t, row_mask, plan);
  MPCG_LAUNCH_CHECK();
  return MPCG_OK;
}
