"""Host-side filter design (not on the hot path; results are cached).

* Butterworth sections come from SciPy, exactly as the reference designs them
  (``signalproc/torchproc.py:32-35``: cut-off divided by fs; ``augment/torchaug.py:96``: EQ
  bands divided by Nyquist).
* Resampler taps are produced here in float64 for both oracles and handed to the kernel in
  the dense frame form  ``y[i*up+p] = sum_d x[i*down + offset + d] * G[p, d]``:
  - ``sinc_hann_frames``  follows torchaudio's ``_get_sinc_resample_kernel``
    (site-packages/torchaudio/functional/functional.py:1340-1404), the arithmetic behind
    ``torchproc.resample`` (``signalproc/torchproc.py:56-59``);
  - ``kaiser_poly_frames`` follows SciPy ``resample_poly`` + ``upfirdn``
    (site-packages/scipy/signal/_signaltools.py:4025-4100), the arithmetic behind
    ``signalproc/resample.py:11-22``.
"""
from __future__ import annotations

import functools
import math

import numpy as np
from scipy import signal as _sig


@functools.lru_cache(maxsize=256)
def butter_sos(cutoff_over_fs: float, kind: str, order: int) -> np.ndarray:
    """[n_sections, 6] SciPy-layout sections of the causal Butterworth the reference applies.

    order 2 is returned as the single (b, a) section verbatim, which is what ``lfilter`` sees.
    """
    if order == 2:
        b, a = _sig.butter(2, cutoff_over_fs, btype=kind)
        return np.concatenate([b, a])[None].astype(np.float64)
    return np.asarray(_sig.butter(order, cutoff_over_fs, btype=kind, output="sos"), dtype=np.float64)


def eq_band_sos(fs: float, bands) -> np.ndarray:
    """First-order Butterworth band-pass per (lo, hi) Hz pair -> one biquad each, [n, 6].

    The bands of ``parametric_eq`` are drawn afresh on every call (reference ``augment/torchaug.py:94-96``), so this
    sits on the host side of every augmentation batch; ``scipy.signal.butter`` costs ~0.3 ms per band.  The section is
    written out in closed form instead -- SciPy's own recipe for ``butter(1, [lo/nyq, hi/nyq], 'band')``: pre-warp both
    edges (``fs = 2``), analog band-pass ``H(s) = bw s / (s^2 + bw s + w0^2)``, bilinear transform -- and
    ``tests/test_host_logic.py`` pins it to ``butter`` within one unit in the last place.
    """
    nyq = fs / 2.0
    rows = []
    for lo, hi in bands:
        w1 = 4.0 * math.tan(math.pi * (lo / nyq) / 2.0)
        w2 = 4.0 * math.tan(math.pi * (hi / nyq) / 2.0)
        bw, w0sq, k = w2 - w1, w1 * w2, 4.0
        a0 = k * k + bw * k + w0sq
        rows.append([bw * k / a0, 0.0, -bw * k / a0, 1.0, (2.0 * w0sq - 2.0 * k * k) / a0, (k * k - bw * k + w0sq) / a0])
    return np.asarray(rows, dtype=np.float64)


def eq_band_sos_scipy(fs: float, bands) -> np.ndarray:
    """The same sections straight from SciPy (the reference's call, ``augment/torchaug.py:96``); test oracle."""
    nyq = fs / 2.0
    rows = []
    for lo, hi in bands:
        b, a = _sig.butter(1, [lo / nyq, hi / nyq], btype="band")
        rows.append(np.concatenate([b, a]))
    return np.asarray(rows, dtype=np.float64)


def reduce_ratio(fs_in: float, fs_out: float) -> tuple[int, int]:
    """(up, down) after removing the common factor of the rounded rates."""
    up, down = int(round(fs_out)), int(round(fs_in))
    g = math.gcd(up, down)
    return up // g, down // g


@functools.lru_cache(maxsize=64)
def sinc_hann_frames(up: int, down: int, zeros: int = 6, rolloff: float = 0.99):
    """torchaudio 'sinc_interp_hann' taps in float64.  Returns (G[up, D], offset, D)."""
    base = min(up, down) * rolloff
    width = math.ceil(zeros * down / base)
    idx = np.arange(-width, width + down, dtype=np.float64)[None, :] / down
    t = np.arange(0, -up, -1, dtype=np.float64)[:, None] / up + idx
    t = np.clip(t * base, -zeros, zeros)
    window = np.cos(t * math.pi / zeros / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        core = np.where(t == 0, 1.0, np.sin(t) / t)
    g = core * window * (base / down)
    return np.ascontiguousarray(g), -width, g.shape[1]


def sinc_out_len(t_in: int, up: int, down: int) -> int:
    """torchaudio truncates to ceil(float32(up * t_in / down)) (functional.py:1427)."""
    return int(math.ceil(np.float32(up * t_in / down)))


@functools.lru_cache(maxsize=64)
def kaiser_poly_frames(up: int, down: int, t_in: int):
    """SciPy resample_poly (Kaiser 5.0, 20*max(up,down)+1 taps) in dense frame form, float64."""
    from scipy.signal._upfirdn import _output_len
    max_rate = max(up, down)
    half = 10 * max_rate
    h = _sig.firwin(2 * half + 1, 1.0 / max_rate, window=("kaiser", 5.0)) * up
    pre = down - half % down
    skip = (half + pre) // down                       # outputs dropped at the front
    t_out = kaiser_out_len(t_in, up, down)
    post = 0
    while _output_len(len(h) + pre + post, t_in, up, down) < t_out + skip:
        post += 1
    hp = np.concatenate([np.zeros(pre), h, np.zeros(post)])
    # output i*up+p reads x[i*down + c_p - k] * hp[phi_p + k*up]
    c = [((p + skip) * down) // up for p in range(up)]
    phi = [((p + skip) * down) % up for p in range(up)]
    cnt = [len(range(ph, len(hp), up)) for ph in phi]
    first = min(ci - (ki - 1) for ci, ki in zip(c, cnt))
    depth = max(c) - first + 1
    g = np.zeros((up, depth))
    for p in range(up):
        for k in range(cnt[p]):
            g[p, c[p] - k - first] = hp[phi[p] + k * up]
    return g, first, depth


def kaiser_out_len(t_in: int, up: int, down: int) -> int:
    n = t_in * up
    return n // down + bool(n % down)
