"""CPU oracle, float64 NumPy/SciPy leg -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, in this repo's own words, the arithmetic of the reference's NumPy conditioning
chain so the CUDA path can be checked against it on a box that has no copy of the
reference.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package; nothing under
``wav2vec-heart-sounds_b200/`` does.

What is restated (reference file:line -> function here):

* ``signalproc/normalize.py:11-17``   -> :func:`fill_nans`
* ``signalproc/normalize.py:20-30``   -> :func:`abs_max_normalise`
* ``signalproc/resample.py:11-22``    -> :func:`resample` (SciPy ``resample_poly`` does the work)
* ``signalproc/despike.py:16-54``     -> :func:`remove_spikes`
* ``signalproc/filters.py:25-39``     -> :func:`lowpass` / :func:`highpass` / :func:`bandpass_cascade`
* ``signalproc/preprocess.py:24-37``  -> :func:`preprocess_pcg` / :func:`preprocess_ecg`
* ``signalproc/segment.py:17-52``     -> :class:`WindowSpec`, :func:`window_starts`, :func:`segment`
* ``signalproc/normalize.py:33-38,47-49,59-72`` -> :func:`minmax_normalise`, :func:`z_normalise`, :func:`kpeak_normalise`
* ``signalproc/envelopes.py:11-23``   -> :func:`hilbert_envelope`, :func:`homomorphic_envelope`
* ``datasets/heart_cycles.py:32-69``  -> :func:`split_cycles`, :func:`crossfade`, :func:`rebuild`

Third-party arithmetic (SciPy ``resample_poly``/``butter``/``sosfilt``) is called, not
restated: SciPy is installed on the GPU box too, and it is the very code the reference runs.

Pinning: ``tests/test_oracle_pinned.py`` checks every function here (a) bit-for-bit against
the reference itself when ``/root/reference`` is importable and (b) against the committed
fixtures in ``tests/golden/`` that ``oracle/make_golden.py`` produced by running the
reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
from scipy import signal as _sig

PCG_BAND = (25.0, 450.0)   # Hz, divided by fs (not Nyquist) when designing -- reference convention
ECG_BAND = (2.0, 40.0)
SPIKE_FILL = 1e-4


# --------------------------------------------------------------------------- NaN repair
def fill_nans(x) -> np.ndarray:
    """Copy ``x`` to float64 and bridge NaN runs by linear interpolation (edges hold)."""
    out = np.array(x, dtype=np.float64)
    bad = np.isnan(out)
    if bad.any() and not bad.all():
        good_idx = np.flatnonzero(~bad)
        out[bad] = np.interp(np.flatnonzero(bad), good_idx, out[good_idx])
    return out


# --------------------------------------------------------------------------- normalise
def abs_max_normalise(x) -> np.ndarray:
    """Remove the mean, scale by the largest magnitude (when non-zero), clip to [-1, 1]."""
    v = fill_nans(x)
    v = v - np.mean(v)
    top = np.max(np.abs(v))
    if top > 0:
        v = v / top
    return np.clip(v, -1.0, 1.0)


# --------------------------------------------------------------------------- resample
def rational_ratio(fs_in: float, fs_out: float) -> tuple[int, int]:
    up, down = int(round(fs_out)), int(round(fs_in))
    g = math.gcd(up, down)
    return up // g, down // g


def resample(x, fs_in: float, fs_out: float) -> np.ndarray:
    if fs_in == fs_out:
        return np.asarray(x)
    up, down = rational_ratio(fs_in, fs_out)
    return _sig.resample_poly(x, up, down)


# --------------------------------------------------------------------------- Schmidt despike
def _flip_positions(frame: np.ndarray) -> np.ndarray:
    """Indices i where sign(frame[i]) and sign(frame[i+1]) are strictly opposite (+1/-1)."""
    sg = np.sign(frame)
    return np.flatnonzero(np.abs(sg[1:] - sg[:-1]) > 1)


def spike_span(frame: np.ndarray, peak: int) -> tuple[int, int]:
    """Half-open sample range flattened around ``peak``: after the last flip before the
    peak, up to the first flip at or after it (``len-1`` when there is none)."""
    flips = _flip_positions(frame)
    left = flips[flips < peak]
    right = flips[flips >= peak]
    lo = int(left[-1]) + 1 if left.size else 0
    hi = int(right[0]) if right.size else frame.size - 1
    return lo, hi


def remove_spikes(x, fs: float, threshold: float = 3.0, max_iterations: int = 1000,
                  trace: list | None = None) -> np.ndarray:
    """Schmidt spike removal on one recording, 500 ms frames, float64, mean-of-middle median.

    ``trace`` (optional) collects ``(frame, peak, lo, hi)`` per iteration so tests can check
    the integer decisions bit-exactly.
    """
    v = np.array(x, dtype=np.float64)
    win = round(float(fs) / 2.0)
    if win < 1 or v.size < win:
        return v
    covered = v.size - v.size % win
    frames = v[:covered].reshape(-1, win)            # view: row w = samples [w*win, (w+1)*win)
    for _ in range(max_iterations):
        tops = np.abs(frames).max(axis=1)
        mid = np.median(tops)
        if mid == 0 or not (tops > threshold * mid).any():
            break
        w = int(np.argmax(tops))
        peak = int(np.argmax(np.abs(frames[w])))
        lo, hi = spike_span(frames[w], peak)
        if trace is not None:
            trace.append((w, peak, lo, hi))
        frames[w, lo:hi] = SPIKE_FILL
    return v


# --------------------------------------------------------------------------- band limiting
def _sos(cutoff: float, fs: float, kind: str, order: int) -> np.ndarray:
    return _sig.butter(order, cutoff / fs, btype=kind, output="sos")


def lowpass(x, fs: float, cutoff: float, order: int = 2) -> np.ndarray:
    return _sig.sosfilt(_sos(cutoff, fs, "lowpass", order), np.asarray(x, dtype=np.float64))


def highpass(x, fs: float, cutoff: float, order: int = 2) -> np.ndarray:
    return _sig.sosfilt(_sos(cutoff, fs, "highpass", order), np.asarray(x, dtype=np.float64))


def bandpass_cascade(x, fs: float, low: float, high: float, order: int = 2) -> np.ndarray:
    """Low-pass at the upper edge first, then high-pass at the lower edge (both causal)."""
    return highpass(lowpass(x, fs, high, order), fs, low, order)


# --------------------------------------------------------------------------- chains
def preprocess_pcg(x, fs_in: float, fs_out: float, *, despike: bool = True) -> np.ndarray:
    v = resample(fill_nans(x), fs_in, fs_out)
    if despike:
        v = remove_spikes(v, fs_out)
    return abs_max_normalise(bandpass_cascade(v, fs_out, *PCG_BAND, order=2))


def preprocess_ecg(x, fs_in: float, fs_out: float) -> np.ndarray:
    v = resample(fill_nans(x), fs_in, fs_out)
    return abs_max_normalise(bandpass_cascade(v, fs_out, *ECG_BAND, order=2))


# --------------------------------------------------------------------------- segmentation
@dataclass(frozen=True)
class WindowSpec:
    window_s: float
    overlap_s: float = 0.25
    start_pad_s: float = 0.3

    def window_len(self, fs: float) -> int:
        return int(round(self.window_s * fs))

    def hop_len(self, fs: float) -> int:
        return max(1, int(round((self.window_s - self.overlap_s) * fs)))


def window_starts(n_samples: int, fs: float, spec) -> list[int]:
    """First-sample index of every window; ``[]`` when the start pad eats the recording,
    a single (to be zero-padded) window when fewer than ``win`` samples remain."""
    first = int(round(spec.start_pad_s * fs))
    if n_samples <= first:
        return []
    final = max(first, n_samples - spec.window_len(fs))
    return list(range(first, final + 1, spec.hop_len(fs)))


def segment(x, fs: float, spec) -> np.ndarray:
    """``[T]`` -> ``[N, win]`` or ``[T, C]`` -> ``[N, win, C]`` (copy, zero-padded if short)."""
    v = np.asarray(x)
    win = spec.window_len(fs)
    starts = window_starts(v.shape[0], fs, spec)
    out = np.zeros((len(starts), win) + v.shape[1:], dtype=v.dtype)
    for k, s in enumerate(starts):
        piece = v[s:s + win]
        out[k, :piece.shape[0]] = piece
    return out


# --------------------------------------------------------------------------- generator conditioning
def fade(x, n: int = 128) -> np.ndarray:
    """``datasets/generative.py:36-42`` (``_fade``): linear ramps over the first and last ``n`` samples."""
    x = np.asarray(x, dtype=np.float64)
    if len(x) < 2 * n:
        return x
    x = x.copy()
    x[:n] *= np.linspace(0.0, 1.0, n)
    x[-n:] *= np.linspace(1.0, 0.0, n)
    return x


def fit_length(x, length: int):
    """``signalproc/preprocess.py:45-64``: zero-pad or crop along axis 0; returns ``(array, valid_length)``."""
    x = np.asarray(x)
    orig = x.shape[0]
    if orig < length:
        x = np.pad(x, ((0, length - orig),) + tuple((0, 0) for _ in range(x.ndim - 1)), mode="constant")
    elif orig > length:
        x = x[:length]
    return x, min(orig, length)


def add_chirp(x, fs: float) -> np.ndarray:
    """``signalproc/preprocess.py`` (``add_chirp``): full-band linear chirp scaled to ``max(0.5, max|x|)``."""
    x = np.asarray(x, dtype=np.float64)
    t = np.arange(len(x)) / fs
    wave = np.asarray(_sig.chirp(t, f0=0, f1=fs / 2, t1=t[-1] if len(t) else 1.0, method="linear"))
    peak = np.max(np.abs(wave)) or 1.0
    wave = wave / peak * max(0.5, float(np.max(np.abs(x))) if len(x) else 0.5)
    return x + wave


def generator_item(reference, conditioning, fs: float, crop: int):
    """The waveform part of ``GenerativeDataset.__getitem__`` without cycle rearrangement
    (``datasets/generative.py:88-104,112``): returns ``(ref, con, chirp)``, each ``[crop]``."""
    ref = fit_length(fade(abs_max_normalise(reference)), crop)[0]
    con = fit_length(fade(abs_max_normalise(conditioning)), crop)[0]
    return ref, con, add_chirp(ref, fs)


# --------------------------------------------------------------------------- zero-phase filters
def butter_bandpass_zp(x, fs: float, low: float, high: float, order: int = 4) -> np.ndarray:
    """``signalproc/filters.py:44-48``."""
    nyq = 0.5 * fs
    return _sig.sosfiltfilt(_sig.butter(order, [low / nyq, high / nyq], btype="bandpass", output="sos"), np.asarray(x, dtype=np.float64))


def butter_lowpass_zp(x, fs: float, cutoff: float, order: int = 4) -> np.ndarray:
    """``signalproc/filters.py:51-53``."""
    return _sig.sosfiltfilt(_sig.butter(order, cutoff / (0.5 * fs), btype="lowpass", output="sos"), np.asarray(x, dtype=np.float64))


def butter_highpass_zp(x, fs: float, cutoff: float, order: int = 4) -> np.ndarray:
    """``signalproc/filters.py:56-58``."""
    return _sig.sosfiltfilt(_sig.butter(order, cutoff / (0.5 * fs), btype="highpass", output="sos"), np.asarray(x, dtype=np.float64))


def band_stop_zp(x, fs: float, low: float, high: float, order: int = 4) -> np.ndarray:
    """``signalproc/filters.py:78-82``."""
    nyq = 0.5 * fs
    return _sig.sosfiltfilt(_sig.butter(order, [low / nyq, high / nyq], btype="bandstop", output="sos"), np.asarray(x, dtype=np.float64))


def notch_zp(x, fs: float, freq: float, q: float = 30.0) -> np.ndarray:
    """``signalproc/filters.py:62-65``."""
    b, a = _sig.iirnotch(freq / (0.5 * fs), q)
    return _sig.filtfilt(b, a, np.asarray(x, dtype=np.float64))


def notch_chain_zp(x, fs: float, freqs, q: float = 55.0) -> np.ndarray:
    """``signalproc/filters.py:68-74``."""
    y = np.asarray(x, dtype=np.float64)
    for f in freqs:
        if f < 0.5 * fs:
            y = notch_zp(y, fs, f, q)
    return y


def fir_subbands(fs: float, taps: int = 61, edges=(45.0, 80.0, 200.0)):
    """``signalproc/filters.py:85-95``."""
    nyq = 0.5 * fs
    e0, e1, e2 = edges
    return [_sig.firwin(taps, e0 / nyq, window="hamming", pass_zero="lowpass"),
            _sig.firwin(taps, [e0 / nyq, e1 / nyq], window="hamming", pass_zero="bandpass"),
            _sig.firwin(taps, [e1 / nyq, e2 / nyq], window="hamming", pass_zero="bandpass"),
            _sig.firwin(taps, e2 / nyq, window="hamming", pass_zero="highpass")]


def decompose_bands(x, fs: float, **kwargs) -> np.ndarray:
    """``signalproc/filters.py:98-101``: ``[4, T]`` zero-phase FIR sub-bands."""
    return np.stack([_sig.filtfilt(b, [1.0], np.asarray(x, dtype=np.float64)) for b in fir_subbands(fs, **kwargs)], axis=0)


# --------------------------------------------------------------------------- other normalisers
RANGE_EPS = 1e-8


def minmax_normalise(x, lo: float = -1.0, hi: float = 1.0) -> np.ndarray:
    """``signalproc/normalize.py:33-38``: whole-array range; a flat signal becomes the midpoint."""
    v = np.asarray(x, dtype=np.float64)
    bottom, top = v.min(), v.max()
    if top - bottom <= 0:
        return np.full_like(v, 0.5 * (lo + hi))
    return (v - bottom) / (top - bottom) * (hi - lo) + lo


def z_normalise(x, axis: int = 0) -> np.ndarray:
    """``signalproc/normalize.py:47-49``: population standard deviation plus 1e-8."""
    v = np.asarray(x, dtype=np.float64)
    return (v - v.mean(axis=axis)) / (v.std(axis=axis) + RANGE_EPS)


def kpeak_normalise(x, k: int = 3, lo: float = -1.0, hi: float = 1.0) -> np.ndarray:
    """``signalproc/normalize.py:59-72``: range = mean of the k smallest .. mean of the k largest samples."""
    v = np.asarray(x, dtype=np.float64)
    ranked = np.sort(v)
    bottom, top = ranked[:k].mean(), ranked[-k:].mean()
    if top - bottom <= 0:
        return np.full_like(v, 0.5 * (lo + hi))
    return lo + (v - bottom) / (top - bottom) * (hi - lo)


# --------------------------------------------------------------------------- envelopes
def hilbert_envelope(x) -> np.ndarray:
    """``signalproc/envelopes.py:11-13``: magnitude of the analytic signal (SciPy's N-point DFT construction)."""
    return np.abs(_sig.hilbert(np.asarray(x, dtype=np.float64)))


def homomorphic_envelope(x, fs: float, cutoff: float = 8.0, order: int = 6) -> np.ndarray:
    """``signalproc/envelopes.py:16-23``: exp(zero-phase low-pass(log(max(envelope, eps))))."""
    if cutoff >= 0.5 * fs:
        raise ValueError(f"cutoff {cutoff} Hz is above Nyquist for fs={fs}")
    floor = np.finfo(float).eps
    return np.exp(butter_lowpass_zp(np.log(np.maximum(hilbert_envelope(x), floor)), fs, cutoff, order=order))


# --------------------------------------------------------------------------- cardiac-cycle rebuild
def split_cycles(signal, joins) -> list:
    """``datasets/heart_cycles.py:32-35``: pieces between consecutive joins that lie strictly inside the signal."""
    cuts = [j for j in joins if 0 < j < len(signal)]
    return [signal[lo:hi] for lo, hi in zip(cuts[:-1], cuts[1:]) if hi > lo]


def crossfade(a, b, n: int) -> np.ndarray:
    """``datasets/heart_cycles.py:38-52``: join ``a`` and ``b`` over ``n`` samples; linear ramp for flat pieces, else the
    correlation-aware odd/even fade pair."""
    if n <= 1 or len(a) < n or len(b) < n:
        return np.concatenate([a, b])
    out_tail, in_head = a[-n:], b[:n]
    if np.var(out_tail) < 1e-5 or np.var(in_head) < 1e-5:
        gain = np.linspace(0.0, 1.0, n)
    else:
        rho = np.corrcoef(out_tail, in_head)[0, 1]
        rho = 0.0 if np.isnan(rho) else abs(rho)
        u = np.linspace(-1.0, 1.0, n)
        odd = (9 / 16) * np.sin(np.pi / 2 * u) + (1 / 16) * np.sin(3 * np.pi / 2 * u)
        even = np.sqrt(np.clip(0.5 / (1 + rho) - ((1 - rho) / (1 + rho)) * odd ** 2, 0.0, None))
        gain = np.clip(even + odd, 0.0, 1.0)
    return np.concatenate([a[:-n], out_tail * (1.0 - gain) + in_head * gain, b[n:]])


def rebuild(cycles, target_len: int, fade_samples: int) -> np.ndarray:
    """``datasets/heart_cycles.py:55-69``: append cycles cyclically until ``target_len`` is reached; at most
    ``10 * len(cycles) + 5`` joins."""
    if not cycles:
        return np.zeros(target_len)
    out, joins = cycles[0], 0
    while len(out) < target_len:
        out = crossfade(out, cycles[(joins + 1) % len(cycles)], fade_samples)
        joins += 1
        if joins > 10 * len(cycles) + 4:
            break
    return out
