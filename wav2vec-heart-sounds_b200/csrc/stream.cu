// extern "C" entries: mpcg_preprocess_segment_f32, mpcg_preprocess_segment_work_bytes -- launcher of the
// row-streaming fused kernel (stream_kernel.cuh).  Each resampler instance is compiled in its own translation
// unit (stream_inst_*.cu) so the heavily unrolled instantiations build in parallel.
#include "stream_kernel.cuh"
#include <stdlib.h>

namespace mpcg {
#include "stream_instances.h"
#define MPCG_SK_EXTERN_(...) MPCG_SK_EXTERN(__VA_ARGS__)
#define MPCG_SK_EXTERN(U, DN, DD, PS) extern template int sk_launch<U, DN, DD, PS>(const SkParams&, size_t, int, cudaStream_t);
MPCG_SK_EXTERN_(1, 1, 1, 1)
MPCG_SK_EXTERN_(FZ_I8)
MPCG_SK_EXTERN_(FZ_I16)
MPCG_SK_EXTERN_(FZ_I32)
MPCG_SK_EXTERN_(FZ_I8N)
MPCG_SK_EXTERN_(FZ_I16N)
MPCG_SK_EXTERN_(FZ_I32N)
#undef MPCG_SK_EXTERN
#undef MPCG_SK_EXTERN_

static int sk_fill_kind(const mpcg_chain_kind& in, SkKind* k) {
  if (in.n_sections < 1 || in.n_sections > 2) return MPCG_EUNSUPPORTED;
  memset(k, 0, sizeof(*k));
  k->despike = in.despike ? 1 : 0;
  bool ok;
  bq_group_coeffs(&in.sos[0][0], in.n_sections, 0, k->c, &ok);
  if (!ok) return MPCG_EINVAL;
  double A[16], B[4];
  bq_group_AB(k->c, A, B);
  double v[4] = {B[0], B[1], B[2], B[3]};
  for (int j = kSkL - 1; j >= 0; --j) {
    for (int s = 0; s < 4; ++s) k->wt[j][s] = v[s];
    bq_group_step(k->c, v, 0.0);
  }
  bq_mat_pow(A, kSkL, k->mp[0]);
  for (int d = 1; d < 10; ++d) bq_mat_mul(k->mp[d - 1], k->mp[d - 1], k->mp[d]);
  return MPCG_OK;
}

// Persistent grid: as many CTAs per SM as the kernel's shared memory and register budget allows, cached per device.
static int sk_grid_ctas(int* out) {
  static int cached_dev = -1, cached = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev != cached_dev) {
    int sms = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
    cached = kSkCtasPerSm * sms;
    cached_dev = dev;
  }
  if (const char* env = getenv("MPCG_SK_CTAS")) {           // experiments: override the grid size
    const int v = atoi(env);
    if (v > 0) { *out = v; return MPCG_OK; }
  }
  *out = cached;
  return MPCG_OK;
}

static long long sk_row_stride(long long t_max) { return ((t_max + 3) & ~3LL) + 8; }

}  // namespace mpcg

// tools/ only: device buffer [ctas, 16] that receives per-phase cycle counts (NULL = off).
static void* g_sk_debug = nullptr;
extern "C" void mpcg_debug_set_phase_clock_buffer(void* dev_ptr) { g_sk_debug = dev_ptr; }

extern "C" int64_t mpcg_preprocess_segment_work_bytes(int64_t t_out_max) {
  using namespace mpcg;
  if (t_out_max < 1 || t_out_max > 0x3fffffff) return MPCG_EINVAL;
  int ctas = 0;
  if (sk_grid_ctas(&ctas) != MPCG_OK) return MPCG_EINVAL;
  return (int64_t)sizeof(float) * (kSkWorkHeader + (int64_t)ctas * sk_row_stride(t_out_max));
}

extern "C" int mpcg_preprocess_segment_f32(const float* x, float* out, int64_t recordings, int channels,
                                           const mpcg_chain_desc* d, void* work, int64_t work_bytes, int32_t* edits,
                                           int32_t* trace, int trace_cap, void* stream_) {
  using namespace mpcg;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!d || recordings < 0 || channels < 1 || trace_cap < 0) return MPCG_EINVAL;
  if (channels > 8 || d->n_kinds < 1 || d->n_kinds > 2) return MPCG_EUNSUPPORTED;
  if (d->t_in < 1 || d->t_out < 1 || d->seg_win < 1 || d->seg_hop < 1 || d->seg_start < 0) return MPCG_EINVAL;
  if (d->t_in > 0x3fffffff || d->t_out > 0x3fffffff || d->seg_win > 0x3fffffff || d->seg_hop > 0x3fffffff ||
      d->seg_start > 0x3fffffff)
    return MPCG_ERANGE;
  const bool ragged = d->row_t_in != nullptr;
  if (ragged && (!d->row_t_out || !d->row_out_offset)) return MPCG_EINVAL;
  if (!ragged && d->seg_n != mpcg_window_count(d->t_out, d->seg_start, d->seg_win, d->seg_hop)) return MPCG_EINVAL;
  for (int c = 0; c < channels; ++c)
    if (d->kind_of_channel[c] >= d->n_kinds) return MPCG_EINVAL;
  const bool identity = (d->up == d->down);
  if (identity && !ragged && d->t_in != d->t_out) return MPCG_EINVAL;
  if (!identity && !d->taps) return MPCG_EINVAL;
  if (d->despike_max_iterations < 0 || (d->median_mode != MPCG_MEDIAN_LOWER && d->median_mode != MPCG_MEDIAN_MEAN))
    return MPCG_EINVAL;
  if (d->channels_last < 0 || d->channels_last > 2) return MPCG_EINVAL;
  if (recordings == 0) return MPCG_OK;
  if (!x || !out || !work) return MPCG_EINVAL;
  const long long rows = (long long)recordings * channels;
  if (rows > 0x7fffffffLL) return MPCG_ERANGE;

  bool any_despike = false;
  for (int c = 0; c < channels; ++c) any_despike |= (d->kinds[d->kind_of_channel[c]].despike != 0);
  if (any_despike && d->despike_win >= 1) {
    if (d->despike_win > 0x3fffffff) return MPCG_ERANGE;
    // the frame cache holds whole frames; the frame-maxima table holds kSkMaxFrames entries
    if (d->t_out > kSkTile && d->despike_win + 7 > kSkGroups * kSkBuf) return MPCG_EUNSUPPORTED;
    if (d->t_out / d->despike_win > kSkMaxFrames) return MPCG_EUNSUPPORTED;
    if ((d->despike_win + 31) / 32 > kSkBmWords) return MPCG_EUNSUPPORTED;
  }
  int ctas = 0;
  {
    const int rc = sk_grid_ctas(&ctas);
    if (rc != MPCG_OK) return rc;
  }
  const long long stride = sk_row_stride(d->t_out);           // (ragged batches: t_out is the longest row)
  if (work_bytes < (int64_t)sizeof(float) * (kSkWorkHeader + (long long)ctas * stride)) return MPCG_EINVAL;
  if ((uintptr_t)work & 15u) return MPCG_EINVAL;

  SkParams P;
  memset(&P, 0, sizeof(P));
  P.x = x; P.out = out; P.dbg = (long long*)g_sk_debug; P.edits = edits; P.trace = trace; P.trace_cap = trace ? trace_cap : 0;
  P.ticket = (unsigned int*)work;
  P.work_rows = (float*)work + kSkWorkHeader;
  P.work_stride = stride;
  P.row_t_in = d->row_t_in; P.row_t = d->row_t_out; P.row_out = (const long long*)d->row_out_offset;
  P.x_stride = d->t_in;
  P.recordings = recordings;
  P.channels = channels;
  P.t_in = (int)d->t_in; P.t = (int)d->t_out; P.off = (int)d->offset;
  P.win_d = (any_despike && d->despike_win >= 1) ? (int)d->despike_win : 0;
  P.threshold = d->despike_threshold; P.max_iter = d->despike_max_iterations; P.median_mode = d->median_mode;
  P.norm_flags = d->norm_flags;
  P.start = (int)d->seg_start; P.win = (int)d->seg_win; P.hop = (int)d->seg_hop; P.n = (int)d->seg_n;
  P.layout = d->channels_last;
  {
    const char* e = getenv("MPCG_FZ_DESPIKE_SERIAL");       // tests: force the reference-order despike path
    P.serial_despike = (e && atoi(e) != 0) ? 1 : 0;
  }
  P.plane = d->plane_elems > 0 ? d->plane_elems : (long long)recordings * d->seg_n * d->seg_win;
  {                                                           // despiked channels first: the longest rows start first
    int k = 0;
    for (int pass = 0; pass < 2; ++pass)
      for (int c = 0; c < channels; ++c)
        if ((d->kinds[d->kind_of_channel[c]].despike != 0) == (pass == 0)) P.chan_order[k++] = (unsigned char)c;
  }
  for (int c = 0; c < 8; ++c) P.kind_of_channel[c] = c < channels ? d->kind_of_channel[c] : 0;
  for (int k = 0; k < d->n_kinds; ++k) {
    const int rc = sk_fill_kind(d->kinds[k], &P.kinds[k]);
    if (rc != MPCG_OK) return rc;
  }
  if (d->n_kinds == 1) P.kinds[1] = P.kinds[0];

  cudaError_t e = cudaMemsetAsync(work, 0, sizeof(float) * kSkWorkHeader, stream);
  if (e != cudaSuccess) return (int)e;
  if ((long long)ctas > rows) ctas = (int)rows;
  const size_t smem = sizeof(SkShared) + (size_t)kSkGroups * kSkBuf * sizeof(float);

  if (identity) return sk_launch<1, 1, 1, 1>(P, smem, ctas, stream);
  const int D = d->taps_per_phase;
#define MPCG_SK_CASE_(...) MPCG_SK_CASE(__VA_ARGS__)
#define MPCG_SK_CASE(U, DN, DD, PS)                                                                  \
  if (d->up == U && d->down == DN && D == DD && rs_taps_match<U, DN, DD>(d->taps, d->offset))       \
    return sk_launch<U, DN, DD, PS>(P, smem, ctas, stream);
  MPCG_SK_CASE_(FZ_I8)
  MPCG_SK_CASE_(FZ_I16)
  MPCG_SK_CASE_(FZ_I32)
  MPCG_SK_CASE_(FZ_I8N)
  MPCG_SK_CASE_(FZ_I16N)
  MPCG_SK_CASE_(FZ_I32N)
#undef MPCG_SK_CASE
  return MPCG_EUNSUPPORTED;                                // a ratio without a baked tap set: the caller chains the stand-alone kernels
}
