"""CPU oracle for the HPSS augmentation -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  **Parity unpinned.**

The reference's HPSS (``augment/primitives.py:88-123``) delegates every number to librosa 0.11.0
(``uv.lock:1110-1111``: ``stft``, ``istft``, ``decompose.hpss``, ``util.softmask``, ``magphase``), which is not
installed here and cannot be (no network); no reference test exercises it either (every test sets
``prob_hpss=0``).  This module therefore restates librosa's *published* algorithm in NumPy, and calls
``scipy.ndimage.median_filter`` -- the routine librosa itself calls for the selection step -- for the medians.
It is pinned only by internal identities (``tests/test_oracle_hpss.py``: H + P + R == S, ISTFT(STFT(x)) == x on the
trimmed span, median rank/reflection rules against hand-worked cases) and by agreement of its transforms with
``torch.stft`` / ``torch.istft`` (same centring, window and normalisation) and of its medians with
``scipy.ndimage.median_filter``, not by reference outputs.

librosa semantics restated:
* ``stft``: periodic Hann of length n_fft, ``center=True`` with ZERO padding of n_fft//2, frames = 1 + len//hop,
  one-sided spectrum ``[1 + n_fft/2, frames]``.
* ``decompose.hpss(S, kernel_size=(k_h, k_p), margin=(m_h, m_p), power=2)``: ``harm`` = median of |S| along TIME
  (size (1, k_h)), ``perc`` = median along FREQUENCY (size (k_p, 1)), both ``mode='reflect'``; soft masks
  ``softmask(harm, perc*m_h)`` and ``softmask(perc, harm*m_p)``; outputs ``|S| * mask * phase``.
* ``softmask(X, R, power=2)``: ``Z = max(X, R)``; where ``Z < tiny`` the mask is 0 (0.5 only when both margins
  are exactly 1); elsewhere ``(X/Z)^2 / ((X/Z)^2 + (R/Z)^2)``.
* ``istft``: irfft of every frame times the window, overlap-added, divided by the window sum-of-squares where it
  exceeds tiny, trimmed by n_fft//2 on both sides -> length hop * (len // hop).
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage as _ndi
from scipy.signal import get_window as _get_window


def hann(n_fft: int) -> np.ndarray:
    return _get_window("hann", n_fft, fftbins=True).astype(np.float64)


def stft(y: np.ndarray, n_fft: int, hop: int) -> np.ndarray:
    y = np.asarray(y, dtype=np.float64)
    w = hann(n_fft)
    pad = n_fft // 2
    yp = np.concatenate([np.zeros(pad), y, np.zeros(pad)])
    frames = 1 + len(y) // hop
    idx = (np.arange(frames) * hop)[:, None] + np.arange(n_fft)[None]
    return np.fft.rfft(yp[idx] * w[None], axis=1).T               # [1 + n_fft/2, frames]


def istft(spec: np.ndarray, n_fft: int, hop: int) -> np.ndarray:
    w = hann(n_fft)
    frames = spec.shape[1]
    total = n_fft + hop * (frames - 1)
    y = np.zeros(total)
    wsum = np.zeros(total)
    seg = np.fft.irfft(spec.T, n=n_fft, axis=1) * w[None]
    for t in range(frames):
        y[t * hop:t * hop + n_fft] += seg[t]
        wsum[t * hop:t * hop + n_fft] += w ** 2
    ok = wsum > np.finfo(wsum.dtype).tiny
    y[ok] /= wsum[ok]
    return y[n_fft // 2: total - n_fft // 2]


def softmask(x: np.ndarray, ref: np.ndarray, power: float = 2.0, split_zeros: bool = False) -> np.ndarray:
    z = np.maximum(x, ref)
    bad = z < np.finfo(np.float32).tiny
    z = np.where(bad, 1.0, z)
    m = (x / z) ** power
    r = (ref / z) ** power
    out = np.where(bad, 0.5 if split_zeros else 0.0, m / np.where(bad, 1.0, m + r))
    return out


def median_time(mag: np.ndarray, k: int) -> np.ndarray:
    return _ndi.median_filter(mag, size=(1, k), mode="reflect")


def median_freq(mag: np.ndarray, k: int) -> np.ndarray:
    return _ndi.median_filter(mag, size=(k, 1), mode="reflect")


def hpss_spectra(spec: np.ndarray, margin, kernel):
    """(H, P, R) complex spectra of one split; kernel = (k_harm_time, k_perc_freq), margin = (m_harm, m_perc)."""
    mag = np.abs(spec)
    phase = np.where(mag > 0, spec / np.where(mag > 0, mag, 1.0), 1.0 + 0j)   # librosa.magphase: exp(i angle), 1 at 0
    harm = median_time(mag, kernel[0])
    perc = median_freq(mag, kernel[1])
    split = (margin[0] == 1 and margin[1] == 1)
    mh = softmask(harm, perc * margin[0], split_zeros=split)
    mp = softmask(perc, harm * margin[1], split_zeros=split)
    h = mag * mh * phase
    p = mag * mp * phase
    return h, p, spec - (h + p)


def hpss_split(y, n_fft: int, hop: int, margin, kernel):
    """(harmonic, percussive, residual) waveforms, each of length hop * (len(y) // hop)."""
    s = stft(y, n_fft, hop)
    h, p, r = hpss_spectra(s, margin, kernel)
    return istft(h, n_fft, hop), istft(p, n_fft, hop), istft(r, n_fft, hop)


def _norm(x):
    x = x - np.mean(x)
    top = np.max(np.abs(x))
    return np.clip(x / top if top > 0 else x, -1.0, 1.0)


def hpss_recombine(x, params: dict, include_residual: bool = True):
    """Two-stage split and random re-weighting (reference primitives.py:96-123) with every draw injected:
    params = {n_fft1, hop1, n_fft2, hop2, margin1 (2), margin2 (2), kernel1 (2), kernel2 (2),
              w1 (len parts), w2 (len parts), w_mix}.  Returns (signal, length)."""
    x = np.asarray(x, dtype=np.float64)
    harm, perc, resid = hpss_split(x, params["n_fft1"], params["hop1"], params["margin1"], params["kernel1"])
    h1, p1, r1 = hpss_split(harm, params["n_fft2"], params["hop2"], params["margin2"], params["kernel2"])
    h2, p2, r2 = hpss_split(perc, params["n_fft2"], params["hop2"], params["margin2"], params["kernel2"])
    parts = [h1, p1, r1, h2, p2, r2, resid] if include_residual else [h1, p1, h2, p2]
    n = min(len(p) for p in parts)
    parts = [p[:n] for p in parts]
    mix1 = _norm(sum(w * p for w, p in zip(params["w1"], parts)))
    mix2 = _norm(sum(w * _norm(p) for w, p in zip(params["w2"], parts)))
    return _norm(mix1 + params["w_mix"] * mix2), n
