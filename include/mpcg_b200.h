/* mpcg_b200.h -- C ABI of libmpcg_b200.so: the B200 (sm_100a) signal-conditioning hot path of
 * mpcg_wav2vec (MilanMarocchi/wav2vec-heart-sounds).
 *
 * The reference has no FFI of its own: its boundary for this path is a set of module-level Python
 * functions (signalproc/torchproc.py, augment/torchaug.py, signalproc/spectrogram.py).  Each entry
 * point below replaces the body of one of them; the file:line it stands in for is cited.  The Python
 * mirror in wav2vec-heart-sounds_b200/ binds these with ctypes; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *  - every pointer named x / y / out / work is a DEVICE pointer to fp32 unless its comment says "host";
 *  - rows are contiguous, time last: element (r, i) of a [rows, t] block is p[r * t + i];
 *  - nothing here allocates, synchronises or owns memory; kernels are enqueued on `stream`
 *    (a cudaStream_t passed as void*; NULL = legacy default stream);
 *  - return value: 0 success; < 0 argument error (MPCG_E*); > 0 a cudaError_t from the launch.
 *  - there is no host/CPU fallback anywhere in this library.
 */
#ifndef MPCG_B200_H
#define MPCG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPCG_ABI_VERSION 2

#define MPCG_OK 0
#define MPCG_EINVAL (-1)
#define MPCG_ERANGE (-2)
#define MPCG_EUNSUPPORTED (-3)

/* despike median rule */
#define MPCG_MEDIAN_LOWER 0 /* torch.median: lower of the two middle values (torchproc.py:82)      */
#define MPCG_MEDIAN_MEAN 1  /* numpy.median: mean of the two middle values, stop when 0 (despike.py:45-46) */

/* normalise flags */
#define MPCG_NORM_NAN_TO_NUM 1 /* torch.nan_to_num first (torchproc.py:63); without it = torchaug._normalise */
#define MPCG_NORM_PEAK_GT0 2   /* numpy rule: divide only if peak > 0 (normalize.py:28-29) instead of clamp_min(1e-12) */

/* mpcg_row_normalise_f32 modes and flags */
#define MPCG_RN_MINMAX 0 /* (x - min) / span * (hi - lo) + lo                          (normalize.py:33-44) */
#define MPCG_RN_ZSCORE 1 /* (x - mean) / (population std + 1e-8)                       (normalize.py:47-56) */
#define MPCG_RN_KPEAK 2  /* lo + (x - lo_ref) / span * (hi - lo), refs = mean of the k smallest / largest (normalize.py:59-78) */
#define MPCG_RN_EPS 1    /* tensor rule: span + 1e-8 in the denominator; without it the NumPy rule: span <= 0 -> (lo + hi) / 2 */
#define MPCG_RN_GLOBAL 2 /* one range for the whole [rows, t] tensor, as the reference's tensor functions compute it */

int mpcg_abi_version(void);
const char* mpcg_error_string(int code);

/* Causal cascade of second-order sections, zero initial state.
 * Replaces torchproc._causal / lowpass / highpass / bandpass_cascade (signalproc/torchproc.py:38-53),
 * filters.lowpass/highpass/bandpass_cascade (signalproc/filters.py:25-39) and the five lfilter calls of
 * torchaug.parametric_eq (augment/torchaug.py:92-99).
 * sos: HOST pointer, n_sections rows of SciPy layout [b0 b1 b2 a0 a1 a2] (1 <= n_sections <= 6). */
int mpcg_biquad_cascade_f32(const float* x, float* y, int64_t rows, int64_t t, const double* sos, int n_sections,
                            void* stream);

/* Same, but rows whose row_mask[r] == 0 are skipped (their y is not written); row_mask: device [rows] or NULL.
 * Used by the EQ stage of augment_pcg_batch, where only the Bernoulli-selected rows need the coloured signal. */
int mpcg_biquad_cascade_masked_f32(const float* x, float* y, int64_t rows, int64_t t, const double* sos,
                                   int n_sections, const float* row_mask, void* stream);

/* Rational polyphase resampler in dense frame form:
 *   y[i*up + p] = sum_{d<taps_per_phase} x[i*down + offset + d] * taps[p*taps_per_phase + d],  x = 0 outside [0, t_in)
 * for 0 <= i*up + p < t_out.  The host shim builds `taps` for either oracle:
 * torchaudio sinc/Hann (torchproc.resample, signalproc/torchproc.py:56-59) or SciPy Kaiser resample_poly
 * (signalproc/resample.py:11-22).  taps: HOST pointer, up * taps_per_phase floats. */
int mpcg_resample_f32(const float* x, float* y, int64_t rows, int64_t t_in, int64_t t_out, const float* taps, int up,
                      int down, int taps_per_phase, int64_t offset, void* stream);

/* Schmidt spike removal, in place on x[rows, t] (give it a copy: the reference clones first,
 * torchproc.py:72 / despike.py:34).  Replaces torchproc.remove_spikes (signalproc/torchproc.py:69-98) and
 * despike.remove_spikes (signalproc/despike.py:31-54).  win = round(fs/2) is computed by the caller.
 * edits (optional, device, [rows] int32): number of flattening passes each row ran.
 * trace (optional, device, [rows, trace_cap, 4] int32): (frame, peak, lo, hi) of each pass, for parity tests. */
int mpcg_despike_f32(float* x, int64_t rows, int64_t t, int64_t win, double threshold, int max_iterations,
                     int median_mode, int32_t* edits, int32_t* trace, int trace_cap, void* stream);

/* Row-wise  clip((x - mean) / max|x - mean|, -1, 1).
 * Replaces torchproc.abs_max_normalise (signalproc/torchproc.py:62-66), torchaug._normalise
 * (augment/torchaug.py:24-27) and normalize.abs_max_normalise (signalproc/normalize.py:20-30, without its
 * NaN interpolation).  flags: MPCG_NORM_*. */
int mpcg_absmax_norm_f32(const float* x, float* y, int64_t rows, int64_t t, int flags, void* stream);

/* Bridge NaN runs by linear interpolation between the nearest valid samples of the row; edges hold the first / last valid
 * value; a row without a valid sample stays as it is.  Replaces normalize.interpolate_nans (signalproc/normalize.py:11-17),
 * the first step of the NumPy chains (signalproc/preprocess.py:25,34).  In place on x[recordings, channels, t]; row_len
 * (optional, device, [recordings] int32): valid samples of each recording's rows (the rest of the pitch is ignored). */
int mpcg_fill_nans_f32(float* x, int64_t recordings, int channels, int64_t t, const int32_t* row_len, void* stream);

/* Overlapping-window gather.  x is [rows, channels, t]; window k of a row starts at start + k*hop and is
 * zero-filled past t.  channels_last = 0: out[rows, channels, n, win] (torchproc.segment on [B,C,T],
 * signalproc/torchproc.py:119-129, returned there as a view);  channels_last = 1: out[rows, n, win, channels]
 * (loader layout of segment.segment on [T,C], signalproc/segment.py:40-52).  n is computed by the caller
 * with mpcg_window_count. */
int mpcg_segment_f32(const float* x, float* out, int64_t rows, int64_t channels, int64_t t, int64_t start,
                     int64_t win, int64_t hop, int64_t n, int channels_last, void* stream);

/* Window count of the tensor path (torchproc.py:124-128): after dropping `start` samples and zero-padding
 * to one window if short, (len - win) / hop + 1.  Pure host arithmetic. */
int64_t mpcg_window_count(int64_t t, int64_t start, int64_t win, int64_t hop);

/* ---- fused chain ---------------------------------------------------------------------------------
 * One launch for  resample -> [despike] -> low-pass + high-pass -> abs-max normalise -> windows, i.e.
 * torchproc.preprocess_pcg / preprocess_ecg followed by torchproc.segment (signalproc/torchproc.py:101-129;
 * NumPy twins signalproc/preprocess.py:24-37 + signalproc/segment.py:40-52; loader call sites
 * datasets/cinc.py:86-94,115 and datasets/vest.py:50-51,84).  Persistent CTAs stream whole rows through shared
 * memory tile by tile: raw samples are read from HBM once, windows leave once; a row may have any length
 * (30 s at 16 kHz = 480 000 samples) and, with the row tables below, every recording its own length. */
typedef struct mpcg_chain_kind {
  int despike;        /* 1: Schmidt despike before the band filter (PCG); 0: none (ECG)            */
  int n_sections;     /* 1 or 2 second-order sections, applied in order                            */
  double sos[2][6];   /* SciPy layout [b0 b1 b2 a0 a1 a2]                                          */
} mpcg_chain_kind;

typedef struct mpcg_chain_desc {
  int64_t t_in, t_out;          /* samples per row before / after resampling (ragged batch: the row pitch */
                                /*   of x and the longest resampled row)                              */
  int up, down, taps_per_phase; /* resampler in dense frame form (see mpcg_resample_f32); up == down  */
  int64_t offset;               /*   means "no resampling" and taps may be NULL                       */
  const float* taps;            /* HOST pointer, up * taps_per_phase floats                           */
  int64_t despike_win;          /* round(fs_out / 2)                                                  */
  double despike_threshold;
  int despike_max_iterations;
  int median_mode;              /* MPCG_MEDIAN_*                                                      */
  int norm_flags;               /* MPCG_NORM_*                                                        */
  int64_t seg_start, seg_win, seg_hop, seg_n;   /* seg_n = mpcg_window_count(t_out, start, win, hop)  */
  int channels_last;            /* 0: out[rec, ch, n, win]   1: out[rec, n, win, ch]   2: out[ch, rec, n, win] */
  int n_kinds;                  /* 1 or 2 distinct channel recipes                                    */
  mpcg_chain_kind kinds[2];
  uint8_t kind_of_channel[8];   /* recipe index of each channel (channels <= 8)                       */
  /* Ragged batch (all three NULL = every recording has t_in / t_out samples): DEVICE tables indexed by
   * recording -- valid samples of its rows in x, their resampled length, and the element offset in `out` of
   * the recording's block (its windows in the chosen layout; layout 2: offset inside a channel plane).  The
   * kernel derives each recording's window count with mpcg_window_count's rule; seg_n is ignored. */
  const int32_t* row_t_in;
  const int32_t* row_t_out;
  const int64_t* row_out_offset;
  int64_t plane_elems;          /* layout 2 only: elements per channel plane (0 = recordings * seg_n * seg_win) */
} mpcg_chain_desc;

/* Bytes of device scratch mpcg_preprocess_segment_f32 needs for rows of up to t_out_max resampled samples on the
 * current device: a ticket counter plus one row buffer per persistent CTA (despiked rows are parked there between
 * the frame-maxima pass and the filter pass; it is reused row after row and therefore lives in L2).  < 0 = error. */
int64_t mpcg_preprocess_segment_work_bytes(int64_t t_out_max);

/* x: [recordings, channels, t_in] -> out (layout above).  work: 16-byte aligned device scratch of at least
 * mpcg_preprocess_segment_work_bytes(desc->t_out) bytes, private to this call until it has finished (launches that
 * may overlap on different streams need a workspace each).  edits / trace as in mpcg_despike_f32, indexed by
 * row = recording * channels + channel; the passes always run in the reference's order.
 * Returns MPCG_EUNSUPPORTED for a resampling ratio without a baked tap set, > 2 sections, a despike frame longer
 * than one tile or more than 1024 despike frames per row: the caller then chains the stand-alone entry points above,
 * which accept everything. */
int mpcg_preprocess_segment_f32(const float* x, float* out, int64_t recordings, int channels,
                                const mpcg_chain_desc* desc, void* work, int64_t work_bytes, int32_t* edits,
                                int32_t* trace, int trace_cap, void* stream);

/* ---- augmentation (augment/torchaug.py) ----------------------------------------------------------
 * Random draws are made by the caller (the Python mirror draws them with torch in the reference's call
 * order) and handed in; `noise == NULL` selects an in-kernel counter-based Philox stream instead. */
#define MPCG_AUG_IDENTITY 0 /* plain _normalise (torchaug.py:24-27)                                          */
#define MPCG_AUG_NOISE 1    /* x + rowp[0] * z,  rowp[0] = scale*std (add_white_noise, torchaug.py:39-42)     */
#define MPCG_AUG_SINE_MUL 2 /* x * (1 + sum_k a_k sin(2 pi (f_k i/fs + p_k)))  (sinusoidal_envelope, :45-54)  */
#define MPCG_AUG_SINE_ADD 3 /* x + sum_k a_k sin(...)                          (baseline_wander, :57-66)      */
#define MPCG_AUG_SELECT 4   /* take the sample from `noise` (a second tensor): the blend of _apply (:34-36)   */

/* One stage: w = transform(x) on rows whose mask is non-zero (mask NULL: every row), w = x elsewhere; then
 * y = w, or with normalise != 0  y = _normalise(w)  -- i.e. torchaug._apply (torchaug.py:34-36) with the
 * Bernoulli mask given.  normalise == 2 re-normalises only the rows whose mask is on and passes the others through
 * untouched: the NumPy primitives normalise inside the transform (augment/primitives.py:44-70), so a stage that the
 * pipeline skips leaves its signal as it was (augment/pipelines.py:47-60).  rowp: device [rows, 8] floats = (a0, f0, p0, a1, f1, p1, -, -) for the sine ops,
 * (scale*std, ...) for noise.  noise: device [rows, t] standard normals or NULL (Philox keyed by seed/stream_id). */
int mpcg_aug_stage_f32(const float* x, float* y, int64_t rows, int64_t t, int op, float fs, const float* rowp,
                       const float* noise, const float* mask, int normalise, uint64_t seed, uint64_t stream_id,
                       void* stream);

/* amplitude_warp (torchaug.py:69-85): y[r, i] = sum_k curves[r, k] * xpad[r, i + k], xpad = reflect pad ntaps/2.
 * curves: device [rows, ntaps] (ntaps odd, <= 257), built by the caller from the control gains. */
int mpcg_aug_warp_f32(const float* x, float* y, int64_t rows, int64_t t, const float* curves, int ntaps,
                      void* stream);

/* Tail of parametric_eq (torchaug.py:100) after the five band-pass sections produced `coloured`
 * (mpcg_biquad_cascade_f32):  e = N(N(coloured)/50 + N(x)).  mix_only != 0: y = e.  Otherwise the stage of
 * augment_pcg_batch (torchaug.py:109): y = N(mask ? e : x). */
int mpcg_aug_eq_mix_f32(const float* x, const float* coloured, float* y, int64_t rows, int64_t t, const float* mask,
                        int mix_only, void* stream);

/* ---- mel conditioning (signalproc/spectrogram.py:13-45) -------------------------------------------
 * x [rows, t] -> out [rows, n_mels, frames], frames = 1 + t / hop (centred, reflect-padded frames).
 * basis: device [n_hi - n_lo][2][kpad], fp64 (basis_f64 != 0: exact path, fp64 accumulation) or fp32 (fast path)
 * = w[n] cos(2 pi k n / n_fft) / sqrt(sum w^2)  |  -w[n] sin(...) / ...  for
 * the nbins DFT bins k0 .. k0+nbins-1 that carry mel weight (zero-padded to kpad, a multiple of 32) and the
 * window's non-zero span [n_lo, n_hi); fb: device [nbins][n_mels] HTK triangles of those bins.  Both are built
 * once per MelConfig by the host mirror.  log_map != 0 fuses log_mel's dB map (spectrogram.py:44-45).
 * A CTA stages the sample span of 32 frames (8 when 31 hops + a window exceed shared memory); MPCG_ERANGE when even
 * 7 hops + a window exceed ~55 000 samples. */
int mpcg_mel_f32(const float* x, float* out, int64_t rows, int64_t t, int n_fft, int hop, int n_lo, int n_hi, int nbins,
                 int kpad, const void* basis, int basis_f64, const float* fb, int n_mels, int64_t frames, int log_map,
                 void* stream);

/* Per-row random draws of augment_pcg_batch's throughput mode in ONE launch (its own Philox stream; the reference draws
 * the same quantities with ~40 small tensor operations, augment/torchaug.py:39-54,103-111): tab[i][row][j] = offset[i][j] +
 * U * scale[i][j], an independent U per entry (scale 0 leaves offset[i][j]; i = noise 1 | wandering volume |
 * noise 2, the [rows, 8] parameter tables of mpcg_aug_chain_f32); masks[i][row] = (U < prob[i]) as 0 / 1 floats.
 * scale / offset: HOST [3][8]; prob: HOST [4]; tab: device [3][rows][8], 16-byte aligned; masks: device [4][rows]. */
int mpcg_aug_draw_f32(float* tab, float* masks, int64_t rows, const float* scale, const float* offset, const float* prob,
                      uint64_t seed, uint64_t sid, void* stream);

/* The whole of augment_pcg_batch (augment/torchaug.py:103-111) in one kernel: N(x), noise, wandering volume, EQ, noise,
 * every stage behind its per-row mask and followed by the row re-normalisation, rows resident in the registers of a thread-block cluster
 * (read once, written once).  rowp1 / rowp4: [rows, 8], [0] = scale*std of the noise stages; noise1 / noise4: injected
 * standard normals [rows, t] or NULL (in-kernel Philox keyed by (seedN, sidN, row, sample)); rowp2: [rows, 8] =
 * (amp, freq, phase) x 2 of sinusoidal_envelope; eq_sos: [eq_sections, 6] SciPy-layout sections shared by the batch
 * (0 sections = no EQ stage transform); maskN: [rows] 0/1 or NULL (= all rows).  y must not alias x.
 * flags: MPCG_AUG_CHAIN_COLLAPSE = a stage whose mask is off for a row does not re-normalise that (already normalised)
 * row again -- N(N(x)) equals N(x) to ~1e-7 -- which saves its sweep and its exchange; 0 = re-normalise every time.
 * MPCG_EUNSUPPORTED when a row does not fit an 8-CTA cluster (t > 135168): compose mpcg_aug_stage_f32 instead. */
int mpcg_aug_chain_f32(const float* x, float* y, int64_t rows, int64_t t, float fs, const float* rowp1,
                       const float* noise1, const float* mask1, uint64_t seed1, uint64_t sid1, const float* rowp2,
                       const float* mask2, const double* eq_sos, int eq_sections, const float* mask3,
                       const float* rowp4, const float* noise4, const float* mask4, uint64_t seed4, uint64_t sid4,
                       int flags, void* work, int64_t work_bytes, void* stream);
/* work: 16-byte aligned device scratch of mpcg_aug_chain_work_bytes() bytes that receives the EQ recipe of this call (only
 * read when eq_sections > 0), private to the call until it has finished: the library keeps no device state between calls. */
int64_t mpcg_aug_chain_work_bytes(void);
#define MPCG_AUG_CHAIN_COLLAPSE 1

/* Zero-phase IIR filtering, scipy.signal.sosfiltfilt arithmetic (reference signalproc/filters.py:44-90): odd extension by
 * `edge` samples, forward and backward cascade passes started from zi * (first sample).  sos: HOST [n_sections, 6]
 * (<= 6 sections); zi: HOST [n_sections, 2] = scipy.signal.sosfilt_zi(sos); work: device scratch [rows, t + 2 edge];
 * t must exceed edge (SciPy raises ValueError otherwise -> MPCG_EINVAL). */
int mpcg_sosfiltfilt_f32(const float* x, float* y, float* work, int64_t rows, int64_t t, const double* sos, int n_sections,
                         const double* zi, int64_t edge, void* stream);
/* The same with an epilogue on the stored samples: MPCG_EPI_EXP writes exp(y) -- the last step of the homomorphic
 * envelope (reference signalproc/envelopes.py:22-23: exp(butter_lowpass(log(envelope)))). */
#define MPCG_EPI_NONE 0
#define MPCG_EPI_EXP 1
int mpcg_sosfiltfilt_epi_f32(const float* x, float* y, float* work, int64_t rows, int64_t t, const double* sos,
                             int n_sections, const double* zi, int64_t edge, int epilogue, void* stream);

/* Hilbert amplitude envelope of rows x [rows, t] (reference signalproc/envelopes.py:11-13: abs(scipy.signal.hilbert(x)),
 * the analytic signal of the t-point DFT, any t <= 2^19): two fp64 Bluestein chirp convolutions of a power-of-two length
 * M >= 2t - 1 through four-step shared-memory FFTs.  y [rows, t] = the envelope, or with MPCG_ENV_LOG
 * log(max(envelope, DBL_EPSILON)) (the input of the homomorphic envelope's low-pass, envelopes.py:21-22).
 * work: device scratch of mpcg_hilbert_work_bytes(rows, t) bytes (16-byte aligned; -1 = t out of range); rows <= 65535. */
#define MPCG_ENV_LOG 1
int64_t mpcg_hilbert_work_bytes(int64_t rows, int64_t t);
int mpcg_hilbert_envelope_f32(const float* x, float* y, void* work, int64_t work_bytes, int64_t rows, int64_t t, int flags,
                              void* stream);

/* Generator-dataset conditioning of one batch (reference datasets/generative.py:77-115 with
 * signalproc/preprocess.py:45-64): y [rows, crop] = fit_length(fade(abs_max_normalise(x [rows, t])), crop) with 128-sample
 * linear fade ramps (fade_n) at both ends of the length-t signal (none when t < 2 fade_n); chirp (optional, [rows, crop]) =
 * add_chirp(y, fs): y plus a full-band linear chirp scaled to max(0.5, max|y|).  norm_flags as in mpcg_absmax_norm_f32. */
int mpcg_gen_condition_f32(const float* x, float* y, float* chirp, int64_t rows, int64_t t, int64_t crop, int fade_n,
                           double fs, int norm_flags, void* stream);
/* The same for rows of different valid lengths: x [rows, t] with row r holding row_len[r] <= t samples (device int64, or
 * NULL = all t).  MPCG_GEN_NO_NORM in norm_flags skips the normalisation: rows rebuilt from cycles of an already
 * normalised signal are only faded and fitted (datasets/generative.py:80-91). */
#define MPCG_GEN_NO_NORM 4
int mpcg_gen_condition_rows_f32(const float* x, float* y, float* chirp, const int64_t* row_len, int64_t rows, int64_t t,
                                int64_t crop, int fade_n, double fs, int norm_flags, void* stream);

/* Rebuild a signal from rearranged cardiac cycles (reference datasets/heart_cycles.py:38-69, _crossfade + rebuild).
 * x [rows, t]: source rows (already normalised); row r has counts[r] cycles, cycle c = x[r, starts[r, c] : starts[r, c] +
 * lens[r, c]] in the order they are to be joined (device int32 [rows, kmax] / [rows]).  y [rows, cap], out_len [rows]
 * (device int64): out = cycle 0, then crossfade-append cycle (i mod count) over fade_n samples until out_len >= target_len
 * (at most 10 count + 5 joins, as the reference's guard allows); counts[r] == 0 copies the row unchanged.  cap must hold
 * target_len + the longest cycle (and t for pass-through rows); writes never pass cap. */
int mpcg_cycle_rebuild_f32(const float* x, float* y, int64_t* out_len, const int32_t* starts, const int32_t* lens,
                           const int32_t* counts, int64_t rows, int64_t t, int64_t cap, int kmax, int64_t target_len,
                           int fade_n, void* stream);

/* The amplitude normalisers other than abs-max (reference signalproc/normalize.py:33-78), per row of x [rows, t] or, with
 * MPCG_RN_GLOBAL, with one range for the whole tensor (min / max over every element; k-peak: the mean over all rows' k
 * largest / smallest values -- what minmax_normalise_torch / kpeak_normalise_torch return for a batched input).
 * k (k-peak only) is clamped to t.  stats: optional device [rows, 8] doubles receiving (min, max, mean, std, hi_ref, lo_ref,
 * 0, 0) per row; REQUIRED with MPCG_RN_GLOBAL and then [rows + 1, 8] (the last row holds the combined map). */
int mpcg_row_normalise_f32(const float* x, float* y, double* stats, int64_t rows, int64_t t, int mode, int k, double lo,
                           double hi, int flags, void* stream);

/* Tensor-core tier of the same transform (tcgen05, split-fp16 operands, fp32 accumulation in TMEM) for configurations
 * with n_fft = Q * hop (Q <= 8) and win_length == n_fft.  The GEMM multiplies every hop row by Q windowed bases (one per
 * position the row can take inside a frame: window segment and phase folded in, i.e. the window is applied in the time
 * domain as the reference applies it); a frame's bins are the sum of its Q rows' partial sums.  Bins per call:
 * ncols_q / 2 >= nbins with Q * ncols_q <= 256 and a multiple of 16.  basis_f16: device, 16-byte aligned,
 * [hop / 16 chunks][2 (hi, lo)][Q * ncols_q * 16] fp16, each half chunk in the canonical K-major core-matrix order
 * (byte offset of element (n, j), j < 16: (n>>3)*256 + (j>>3)*128 + (n&7)*16 + (j&7)*2); column n = q * ncols_q + c holds
 * w[q hop + j] cos(2 pi k (q hop + j) / n_fft) for c = k - k0 < nbins and -w[..] sin(..) for c = ncols_q / 2 + k - k0; the
 * chunks stream through shared memory (they do not fit beside the sample tile), so they should stay L2-resident.
 * Returns MPCG_EUNSUPPORTED when the shape does not fit; the caller then uses mpcg_mel_f32.
 * log_map: bit 0 = fuse log_mel's dB map, bit 1 = add to the values already in `out` (presets with more weighted bins than
 * one call takes run as several calls over consecutive bin ranges; the last one applies the map). */
int mpcg_mel_tc_f32(const float* x, float* out, int64_t rows, int64_t t, int n_fft, int hop, int k0, int nbins, int ncols_q,
                    const void* basis_f16, const float* fb, float inv_norm, int n_mels, int64_t frames, int log_map,
                    void* stream);

/* mpcg_mel_f32's float64 tier on the fp64 tensor path (mma.sync.m8n8k4.f64; csrc/mel_dm.cu): same arguments and the same
 * 1e-5 bound against the float64 reference on any input (signalproc/spectrogram.py:13-45), 2-3x faster.  basis_dm: the
 * windowed, normalised basis in fragment order, float64 [kpad / 32][ceil((n_hi - n_lo) / 16)][64][20]: block b, slice s,
 * column c = 2 (bin - 32 b) + part (0: cos, 1: -sin), entry j < 16 = sample n_lo + 16 s + j (zero beyond the window / the
 * weighted bins; entries 16..19 are padding).  mel_range: device int32 [n_mels][2], the run of bins [lo, hi) (relative to the
 * first weighted bin) on which filter m is non-zero (HTK triangles; lo = hi for an empty filter): the projection skips the
 * zeros.  MPCG_EUNSUPPORTED (hop < 4, shared memory) -> use mpcg_mel_f32. */
int mpcg_mel_dm_f32(const float* x, float* out, int64_t rows, int64_t t, int n_fft, int hop, int n_lo, int n_hi, int nbins,
                    int kpad, const double* basis_dm, const float* fb, const int* mel_range, int n_mels, int64_t frames,
                    int log_map, void* stream);

/* y = clamp((20 log10(max(x, 1e-5)) - 20 + 100) / 100, 0, 1) elementwise (log_mel on a foreign transform's output). */
int mpcg_logmap_f32(const float* x, float* y, int64_t n, void* stream);

/* ---- HPSS augmentation (augment/primitives.py:88-123; librosa stft / decompose.hpss / softmask / istft) -------
 * Spectra are frame-major complex64: spec[row][frame][bin], bins = n_fft/2 + 1, frames = 1 + t/hop.
 * window: device [n_fft] periodic Hann; twiddle: device [n_fft/2] complex64 exp(-2 pi i k / n_fft). n_fft = 2^m. */
int mpcg_hpss_stft_f32(const float* x, float* spec, int64_t rows, int64_t t, int n_fft, int hop, int64_t frames,
                       const float* window, const float* twiddle, void* stream);
/* out[row][frame][bin] = median of |spec| over k neighbours along time (along_time != 0: librosa's harmonic
 * filter, size (1, k)) or along frequency (percussive, size (k, 1)); scipy.ndimage 'reflect' boundary, rank k/2. */
int mpcg_hpss_median_f32(const float* spec, float* out, int64_t rows, int64_t frames, int bins, int k, int along_time,
                         void* stream);
/* Soft masks (power 2, margins as in decompose.hpss) -> harmonic and percussive spectra -> ONE packed inverse FFT per frame
 * (ifft(H + iP) = h + ip) -> windowed overlap-add into acc[row][0..1][n_fft + hop*(frames-1)] (acc[row][3][..] zeroed here;
 * the residual plane stays zero: mpcg_hpss_finish3_f32 forms the residual as x - harmonic - percussive). */
int mpcg_hpss_istft_f32(const float* spec, const float* harm, const float* perc, float* acc, int64_t rows, int n_fft,
                        int hop, int64_t frames, float margin_h, float margin_p, const float* window,
                        const float* twiddle, void* stream);
/* y[line][i] = acc[line][i + n_fft/2] / wsum[i + n_fft/2] (where wsum > tiny), i < hop*(frames-1): istft's
 * normalisation and centre trim.  wsum: device [n_fft + hop*(frames-1)] window sum-of-squares. */
int mpcg_hpss_finish_f32(const float* acc, const float* wsum, float* y, int64_t lines, int n_fft, int hop,
                         int64_t frames, void* stream);
/* y[row][c][i], c = 0 harmonic, 1 percussive: acc / wsum as above; c = 2 residual: x[row][i] - (harmonic + percussive)
 * (istft is linear and istft(stft(x)) = x, so this equals istft(S - (H + P)), primitives.py:92).  x: device [rows, t]. */
int mpcg_hpss_finish3_f32(const float* acc, const float* wsum, const float* x, float* y, int64_t rows, int64_t t, int n_fft,
                          int hop, int64_t frames, void* stream);
/* hpss_recombine's tail (primitives.py:117-123): out = N(N(sum w1_p part_p) + wmix * N(sum w2_p N(part_p))), N =
 * NumPy abs_max_normalise.  parts: device [nparts][rows][n]; w1, w2: HOST [nparts]. */
int mpcg_hpss_mix_f32(const float* parts, float* out, int64_t rows, int64_t n, int nparts, const float* w1,
                      const float* w2, float wmix, void* stream);

/* ---- time warp and recorded-noise mixing (NumPy-only in the reference; parity unpinned) ------------------------
 * y[r, j] = x_r(j * rate), 4-point Catmull-Rom, edges clamped; n_out is chosen by the caller (round(t / rate)).
 * Stands in for primitives.time_stretch (augment/primitives.py:30-34), whose rubberband arithmetic is external. */
int mpcg_time_warp_f32(const float* x, float* y, int64_t rows, int64_t t, int64_t n_out, double rate, void* stream);
/* y[r] = N(x[r] + scale[r] * N(bank[src_row[r], src_start[r] : src_start[r] + t])), N = NumPy abs_max_normalise
 * (noise_sources.py:53-64 + pipelines.py:59-60).  bank: device [bank_rows, bank_len]; src_row / src_start: device
 * int64 [rows]; scale: device [rows].  The caller guarantees src_start + t <= bank_len. */
int mpcg_mix_noise_f32(const float* x, const float* bank, float* y, int64_t rows, int64_t t, int64_t bank_rows,
                       int64_t bank_len, const int64_t* src_row, const int64_t* src_start, const float* scale,
                       void* stream);

/* Recorded clinical noise for a batch (reference augment/noise_sources.py:33-64 after its file reads): for row r
 *   out[r] = sum_c scale[r, c] * N(bank[src_row[r, c], src_start[r, c] : src_start[r, c] + t]),   c < ncomp <= 4,
 * N = NumPy abs_max_normalise of the cropped record; with normalise_sum != 0 the sum is normalised again when its peak is
 * non-zero (pcg_noise, :48-50), otherwise it is returned as it is (ecg_noise, :61).  bank: device [bank_rows, bank_len]
 * records already resampled to the signal rate (mpcg_resample_f32 with the SciPy taps = the reference's resample_poly);
 * src_row / src_start: device int64 [rows, ncomp]; scale: device [rows, ncomp] (a zero scale skips the component's
 * statistics, as the reference's choice([0, U]) does).  Entries are clamped into the bank. */
int mpcg_noise_combine_f32(const float* bank, float* out, int64_t rows, int64_t t, int64_t bank_rows, int64_t bank_len,
                           int ncomp, const int64_t* src_row, const int64_t* src_start, const float* scale,
                           int normalise_sum, void* stream);

/* Time-varying sinc delay-and-sum of a multichannel batch (reference classify/beamformer.py:41-55):
 *   out[b, t] = sum_m ( sum_k xp[b, m, t + k] * kern_k(delays[b, m, t]) )^2,   kern_k(d) = sinc(k - K/2 - d) window[k] / sum_k(..),
 * xp = x reflect-padded by K/2.  x, delays: device [batch, mics, t]; window: HOST [kernel_size] (Hamming in the reference,
 * kernel_size odd, <= 129).  aux (optional, device, [batch, mics, t, 4] floats) receives what the backward pass needs.
 * The backward entry turns grad_out [batch, t] into grad_x and / or grad_delays [batch, mics, t] (either may be NULL). */
int mpcg_beamform_fwd_f32(const float* x, const float* delays, float* out, float* aux, int64_t batch, int mics, int64_t t,
                          const float* window, int kernel_size, void* stream);
int mpcg_beamform_bwd_f32(const float* delays, const float* aux, const float* grad_out, float* grad_x, float* grad_delays,
                          int64_t batch, int mics, int64_t t, const float* window, int kernel_size, void* stream);

/* Profiling aid (tools/ only): device buffer [ctas, 16] of int64 that the fused kernel fills with clock64 stamps at
 * its phase boundaries; NULL switches it off.  Not part of the data path. */
void mpcg_debug_set_phase_clock_buffer(void* dev_ptr);

#ifdef __cplusplus
}
#endif
#endif /* MPCG_B200_H */
