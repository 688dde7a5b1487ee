// tools-only entry: mpcg_preprocess_segment_cluster_f32 -- the round-1 cluster kernel, kept for A/B timing against stream.cu.  The kernel template lives in
// fused_kernel.cuh; each resampler instance is compiled in its own translation unit (fused_inst_*.cu) so the
// seven heavily unrolled instantiations build in parallel.
#include "fused_kernel.cuh"
#include <stdlib.h>

namespace mpcg {
#include "fused_instances.h"
#define MPCG_FZ_EXTERN_(...) MPCG_FZ_EXTERN(__VA_ARGS__)
#define MPCG_FZ_EXTERN(U, DN, DD, PS) \
  extern template int fz_launch<U, DN, DD, PS>(const FzParams&, size_t, long long, cudaStream_t);
MPCG_FZ_EXTERN_(1, 1, 1, 1)
MPCG_FZ_EXTERN_(FZ_I8)
MPCG_FZ_EXTERN_(FZ_I16)
MPCG_FZ_EXTERN_(FZ_I32)
MPCG_FZ_EXTERN_(FZ_I8N)
MPCG_FZ_EXTERN_(FZ_I16N)
MPCG_FZ_EXTERN_(FZ_I32N)
#undef MPCG_FZ_EXTERN
#undef MPCG_FZ_EXTERN_

struct FzGeometry {
  int ncl, S, L, cap, q, nq, fpc, nframes;
  size_t smem;
};

// Shared-memory budget per CTA the planner aims for (MPCG_FZ_SMEM_KB overrides, for experiments).
static size_t fz_smem_target() {
  const char* e = getenv("MPCG_FZ_SMEM_KB");
  const long dflt = (MPCG_FZ_MINBLOCKS >= 2) ? 112 : 160;
  const long kb = e ? atol(e) : dflt;
  return (size_t)(kb > 0 ? kb : dflt) * 1024;
}

// Smallest cluster whose slice leaves room for two CTAs per SM; otherwise the smallest that fits at all.
static bool fz_plan_geometry(int t, int win_d, bool frames, FzGeometry* g) {
  FzGeometry best{};
  bool have = false;
  for (int ncl = 1; ncl <= kFzMaxCluster; ++ncl) {
    FzGeometry c{};
    c.ncl = ncl;
    if (frames) {
      c.nframes = t / win_d;
      if (c.nframes > kFzMaxFrames || c.nframes < 1) return false;
      c.fpc = (c.nframes + ncl - 1) / ncl;
      c.S = c.fpc * win_d;
    } else {
      c.nframes = 0;
      c.fpc = 1;
      c.S = (((t + ncl - 1) / ncl) + 3) & ~3;
    }
    const long long last = (long long)t - (long long)(ncl - 1) * c.S;
    if (last <= 0) continue;                              // the last CTA must own the end of the row
    const int need = (int)(last > c.S ? last : c.S);
    int L = (need + kFzChunks - 1) / kFzChunks;
    L |= 1;
    if (L > kFzLmax) continue;
    c.L = L;
    c.cap = L * kFzChunks;
    c.q = (c.S - 1) / L;
    c.nq = c.S - c.q * L;
    c.smem = sizeof(FzShared) + (size_t)(c.cap + 2 * kFzGuard + 8) * sizeof(float);
    if (c.smem > 225 * 1024) continue;
    if (c.smem <= fz_smem_target()) { *g = c; return true; }
    if (!have) { best = c; have = true; }
  }
  if (have) *g = best;
  return have;
}

static int fz_fill_kind(const mpcg_chain_kind& in, const FzGeometry& g, FzKind* k) {
  if (in.n_sections < 1 || in.n_sections > 2) return MPCG_EUNSUPPORTED;
  memset(k, 0, sizeof(*k));
  k->despike = in.despike ? 1 : 0;
  bool ok;
  bq_group_coeffs(&in.sos[0][0], in.n_sections, 0, k->c, &ok);
  if (!ok) return MPCG_EINVAL;
  double A[16], B[4];
  bq_group_AB(k->c, A, B);
  double v[4] = {B[0], B[1], B[2], B[3]};
  for (int j = g.L - 1; j >= 0; --j) {
    for (int s = 0; s < 4; ++s) k->wt[j][s] = v[s];
    bq_group_step(k->c, v, 0.0);
  }
  bq_mat_pow(A, g.L, k->mp[0]);
  for (int d = 1; d < 10; ++d) bq_mat_mul(k->mp[d - 1], k->mp[d - 1], k->mp[d]);
  for (int l = 0; l < 32; ++l) bq_mat_pow(k->mp[0], l, k->mlane[l]);
  for (int j = 0; j < kFzMaxCluster; ++j) bq_mat_pow(A, (long long)g.S * j, k->prop_pow[j]);
  for (int w = 0; w < kFzFW; ++w) bq_mat_pow(k->mp[0], 32LL * w, k->mwarp[w]);
  bq_mat_pow(A, g.nq, k->prop_part);
  return MPCG_OK;
}

}  // namespace mpcg

// Channel recipes live in a small ring of device-global slots (the ABI never allocates).  A slot is rewritten
// only when a call brings a recipe pair that is not already resident; the upload is ordered on the caller's
// stream.  Calls that use different recipes concurrently on different streams should stay within the ring
// (8 distinct recipe pairs in flight).
constexpr int kFzPlanSlots = 8;
__device__ mpcg::FzKind g_fz_plan_dev[kFzPlanSlots][2];
static mpcg::FzKind g_fz_plan_host[kFzPlanSlots][2];
static bool g_fz_plan_valid[kFzPlanSlots];
static int g_fz_plan_next = 0;
static int g_fz_plan_device = -1;

static int fz_resident_plan(const mpcg::FzKind (&kinds)[2], cudaStream_t stream, const mpcg::FzKind** dev_out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev != g_fz_plan_device) {                          // symbols are per device: start over
    for (int i = 0; i < kFzPlanSlots; ++i) g_fz_plan_valid[i] = false;
    g_fz_plan_device = dev;
  }
  mpcg::FzKind* base = nullptr;
  e = cudaGetSymbolAddress((void**)&base, g_fz_plan_dev);
  if (e != cudaSuccess) return (int)e;
  for (int i = 0; i < kFzPlanSlots; ++i)
    if (g_fz_plan_valid[i] && memcmp(g_fz_plan_host[i], kinds, sizeof(kinds)) == 0) {
      *dev_out = base + 2 * i;
      return MPCG_OK;
    }
  const int slot = g_fz_plan_next;
  g_fz_plan_next = (g_fz_plan_next + 1) % kFzPlanSlots;
  memcpy(g_fz_plan_host[slot], kinds, sizeof(kinds));
  g_fz_plan_valid[slot] = true;
  e = cudaMemcpyAsync(base + 2 * slot, g_fz_plan_host[slot], sizeof(kinds), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) { g_fz_plan_valid[slot] = false; return (int)e; }
  *dev_out = base + 2 * slot;
  return MPCG_OK;
}

// tools/ only: device buffer [ctas, 16] that receives clock64 stamps per phase (NULL = off).
static void* g_fz_debug = nullptr;
extern "C" void mpcg_debug_set_phase_clock_buffer_cluster(void* dev_ptr) { g_fz_debug = dev_ptr; }

extern "C" int mpcg_preprocess_segment_cluster_f32(const float* x, float* out, int64_t recordings, int channels,
                                           const mpcg_chain_desc* d, int32_t* edits, int32_t* trace, int trace_cap,
                                           void* stream_) {
  using namespace mpcg;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!d || recordings < 0 || channels < 1 || trace_cap < 0) return MPCG_EINVAL;
  if (channels > 8 || d->n_kinds < 1 || d->n_kinds > 2) return MPCG_EUNSUPPORTED;
  if (d->t_in < 1 || d->t_out < 1 || d->seg_win < 1 || d->seg_hop < 1 || d->seg_start < 0) return MPCG_EINVAL;
  if (d->t_in > 0x3fffffff || d->t_out > 0x3fffffff || d->seg_win > 0x3fffffff || d->seg_hop > 0x3fffffff ||
      d->seg_start > 0x3fffffff)
    return MPCG_ERANGE;
  if (d->seg_n != mpcg_window_count(d->t_out, d->seg_start, d->seg_win, d->seg_hop)) return MPCG_EINVAL;
  if (d->seg_n > 4096) return MPCG_EUNSUPPORTED;
  for (int c = 0; c < channels; ++c)
    if (d->kind_of_channel[c] >= d->n_kinds) return MPCG_EINVAL;
  const bool identity = (d->up == d->down);
  if (identity && d->t_in != d->t_out) return MPCG_EINVAL;
  if (!identity && !d->taps) return MPCG_EINVAL;
  if (recordings == 0) return MPCG_OK;
  if (!x || !out) return MPCG_EINVAL;
  const long long rows = (long long)recordings * channels;
  if (rows * kFzMaxCluster > 0x7fffffffLL) return MPCG_ERANGE;

  bool any_despike = false;
  for (int k = 0; k < d->n_kinds; ++k) any_despike |= (d->kinds[k].despike != 0);
  const bool frames = any_despike && d->despike_win >= 1 && d->t_out >= d->despike_win;
  if (frames && d->despike_win > 0x3fffffff) return MPCG_ERANGE;
  FzGeometry g;
  if (!fz_plan_geometry((int)d->t_out, (int)d->despike_win, frames, &g)) return MPCG_EUNSUPPORTED;

  FzParams P;
  memset(&P, 0, sizeof(P));
  P.x = x; P.out = out; P.dbg = (long long*)g_fz_debug; P.edits = edits; P.trace = trace; P.trace_cap = trace ? trace_cap : 0;
  P.channels = channels;
  P.t_in = (int)d->t_in; P.t = (int)d->t_out; P.off = (int)d->offset; P.identity = identity ? 1 : 0;
  P.ncl = g.ncl; P.S = g.S; P.L = g.L; P.cap = g.cap; P.q = g.q; P.nq = g.nq;
  P.win_d = frames ? (int)d->despike_win : 1; P.nframes = frames ? g.nframes : 0; P.fpc = g.fpc;
  P.threshold = d->despike_threshold; P.max_iter = d->despike_max_iterations; P.median_mode = d->median_mode;
  P.norm_flags = d->norm_flags;
  {
    const char* e = getenv("MPCG_FZ_DESPIKE_SERIAL");       // tests: force the reference-order despike path
    P.serial_despike = (e && atoi(e) != 0) ? 1 : 0;
  }
  P.start = (int)d->seg_start; P.win = (int)d->seg_win; P.hop = (int)d->seg_hop; P.n = (int)d->seg_n;
  if (d->channels_last == 2) {                             // channel-major: out[channel, recording, window, sample]
    P.so_j = 1; P.so_k = d->seg_win; P.so_b = (long long)d->seg_n * d->seg_win;
    P.so_c = (long long)recordings * d->seg_n * d->seg_win;
  } else if (d->channels_last) {
    P.so_j = channels; P.so_c = 1; P.so_k = (long long)d->seg_win * channels;
    P.so_b = (long long)d->seg_n * d->seg_win * channels;
  } else {
    P.so_j = 1; P.so_k = d->seg_win; P.so_c = (long long)d->seg_n * d->seg_win;
    P.so_b = (long long)channels * d->seg_n * d->seg_win;
  }
  for (int c = 0; c < 8; ++c) P.kind_of_channel[c] = c < channels ? d->kind_of_channel[c] : 0;
  FzKind kinds[2];
  memset(kinds, 0, sizeof(kinds));
  for (int k = 0; k < d->n_kinds; ++k) {
    const int rc = fz_fill_kind(d->kinds[k], g, &kinds[k]);
    if (rc != MPCG_OK) return rc;
  }
  {
    const int rc = fz_resident_plan(kinds, stream, &P.kinds);
    if (rc != MPCG_OK) return rc;
  }
  if (P.max_iter < 0 || (P.median_mode != MPCG_MEDIAN_LOWER && P.median_mode != MPCG_MEDIAN_MEAN)) return MPCG_EINVAL;

  if (identity) return fz_launch<1, 1, 1, 1>(P, g.smem, rows, stream);
  const int D = d->taps_per_phase;
#define MPCG_FZ_CASE_(...) MPCG_FZ_CASE(__VA_ARGS__)
#define MPCG_FZ_CASE(U, DN, DD, PS)                                                                  \
  if (d->up == U && d->down == DN && D == DD && rs_taps_match<U, DN, DD>(d->taps, d->offset))       \
    return fz_launch<U, DN, DD, PS>(P, g.smem, rows, stream);
  MPCG_FZ_CASE_(FZ_I8)
  MPCG_FZ_CASE_(FZ_I16)
  MPCG_FZ_CASE_(FZ_I32)
  MPCG_FZ_CASE_(FZ_I8N)
  MPCG_FZ_CASE_(FZ_I16N)
  MPCG_FZ_CASE_(FZ_I32N)
#undef MPCG_FZ_CASE
  return MPCG_EUNSUPPORTED;                                // caller composes the stand-alone kernels instead
}
