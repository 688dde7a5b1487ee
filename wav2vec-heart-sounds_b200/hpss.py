"""Batched HPSS augmentation on the device.  The reference only has this on its NumPy path
(``augment/primitives.py:88-123``: ``_hpss_split`` and ``hpss_recombine``, arithmetic by librosa 0.11); the names
and argument order follow it, operating on ``[B, T]`` batches with one set of parameters per call.

Parity note: librosa is not installable here and no reference test touches HPSS, so this row is checked against
``oracle/hpss_path.py`` (a restatement of librosa's published algorithm that calls ``scipy.ndimage.median_filter``
for the selection step) -- "parity unpinned" in DESIGN.md.  Median selections are bit-exact for identical
magnitudes; transforms and masks are float32 with float32-level error.
"""
from __future__ import annotations

import functools
import math
import random

import numpy as np
import torch

from . import _lib


@functools.lru_cache(maxsize=32)
def _tables_host(n_fft: int):
    from scipy.signal import get_window
    w = get_window("hann", n_fft, fftbins=True).astype(np.float64)
    k = np.arange(n_fft // 2, dtype=np.float64)
    tw = np.stack([np.cos(2 * np.pi * k / n_fft), -np.sin(2 * np.pi * k / n_fft)], axis=1)
    return w, tw.astype(np.float32)


_DEV_TABLES: dict = {}


def _tables(n_fft: int, hop: int, frames: int, device):
    key = (n_fft, str(device))
    if key not in _DEV_TABLES:
        w, tw = _tables_host(n_fft)
        _DEV_TABLES[key] = (torch.from_numpy(w.astype(np.float32)).to(device), torch.from_numpy(tw).to(device))
    wkey = (n_fft, hop, frames, str(device))
    if wkey not in _DEV_TABLES:
        w, _ = _tables_host(n_fft)
        total = n_fft + hop * (frames - 1)
        ws = np.zeros(total)
        for t in range(frames):
            ws[t * hop:t * hop + n_fft] += w ** 2
        _DEV_TABLES[wkey] = torch.from_numpy(ws.astype(np.float32)).to(device)
    return _DEV_TABLES[key] + (_DEV_TABLES[wkey],)


def stft(x: torch.Tensor, n_fft: int, hop: int) -> torch.Tensor:
    """``[B, T]`` -> complex64 ``[B, frames, 1 + n_fft/2]`` (frame-major; librosa.stft(center=True, zero pad))."""
    x = _lib.require_cuda_f32(x)
    b, t = x.shape
    frames = 1 + t // hop
    win, tw, _ = _tables(n_fft, hop, frames, x.device)
    spec = torch.empty((b, frames, n_fft // 2 + 1, 2), device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().mpcg_hpss_stft_f32(x.data_ptr(), spec.data_ptr(), b, t, n_fft, hop, frames, win.data_ptr(),
                                             tw.data_ptr(), _lib.stream_ptr(x)), "hpss stft")
    return torch.view_as_complex(spec)


def median_magnitude(spec: torch.Tensor, k: int, along_time: bool) -> torch.Tensor:
    """Running median of ``|spec|`` (``[B, frames, bins]`` complex64) over ``k`` neighbours along time or frequency."""
    if not (spec.is_cuda and spec.dtype == torch.complex64 and spec.is_contiguous() and spec.dim() == 3):
        raise ValueError("spec must be a contiguous CUDA complex64 [B, frames, bins] tensor")
    b, frames, bins = spec.shape
    out = torch.empty((b, frames, bins), device=spec.device, dtype=torch.float32)
    _lib.check(_lib.lib().mpcg_hpss_median_f32(spec.data_ptr(), out.data_ptr(), b, frames, bins, int(k),
                                               1 if along_time else 0, _lib.stream_ptr(spec)), "hpss median")
    return out


def hpss_split(y: torch.Tensor, n_fft: int, hop: int, margin, kernel):
    """``_hpss_split`` (reference primitives.py:88-93) for a batch: returns (harmonic, percussive, residual)
    waveforms ``[B, hop * (T // hop)]``.  ``kernel = (k_harmonic_time, k_percussive_freq)``, ``margin = (m_h, m_p)``."""
    y = _lib.require_cuda_f32(y)
    if y.dim() != 2:
        raise ValueError("hpss_split takes a [B, T] batch")
    b, t = y.shape
    frames = 1 + t // hop
    n_out = hop * (frames - 1)
    win, tw, wsum = _tables(n_fft, hop, frames, y.device)
    out = torch.empty((b, 3, n_out), device=y.device, dtype=torch.float32)
    bins = n_fft // 2 + 1
    per_row = frames * bins * 16 + 3 * (n_fft + n_out) * 4
    chunk = max(1, min(b, int(6e9 // per_row)))                  # bound the scratch to ~6 GB
    for lo in range(0, b, chunk):
        rows = y[lo:lo + chunk]
        nb = rows.shape[0]
        spec = stft(rows, n_fft, hop)
        harm = median_magnitude(spec, kernel[0], True)
        perc = median_magnitude(spec, kernel[1], False)
        acc = torch.empty((nb, 3, n_fft + n_out), device=y.device, dtype=torch.float32)
        _lib.check(_lib.lib().mpcg_hpss_istft_f32(spec.data_ptr(), harm.data_ptr(), perc.data_ptr(), acc.data_ptr(), nb,
                                                  n_fft, hop, frames, float(margin[0]), float(margin[1]), win.data_ptr(),
                                                  tw.data_ptr(), _lib.stream_ptr(y)), "hpss istft")
        _lib.check(_lib.lib().mpcg_hpss_finish3_f32(acc.data_ptr(), wsum.data_ptr(), rows.contiguous().data_ptr(),
                                                    out[lo:lo + nb].data_ptr(), nb, t, n_fft, hop, frames,
                                                    _lib.stream_ptr(y)), "hpss finish")
    return out[:, 0], out[:, 1], out[:, 2]


def draw_recombine_params(include_residual: bool = True) -> dict:
    """The reference's draw order (primitives.py:104-111, 120-122) with Python's ``random``."""
    rf = lambda lo, hi: lo + random.random() * (hi - lo)
    p = {"n_fft1": random.choice([512, 1024, 2048]), "hop1": random.choice([16, 32, 64, 128]),
         "n_fft2": random.choice([512, 1024, 2048]), "hop2": random.choice([16, 32, 64, 128])}
    p["margin1"] = (rf(1.0, 2.0), rf(1.0, 2.0))
    p["margin2"] = (rf(1.0, 4.0), rf(1.0, 4.0))
    p["kernel1"] = (random.randint(5, 30), random.randint(5, 30))
    p["kernel2"] = (random.randint(5, 30), random.randint(5, 30))
    n = 7 if include_residual else 4
    p["w1"] = [rf(0.01, 10) for _ in range(n)]
    p["w2"] = [rf(0.01, 10) for _ in range(n)]
    p["w_mix"] = rf(0.01, 0.05)
    return p


def hpss_recombine(x: torch.Tensor, include_residual: bool = True, *, params: dict | None = None):
    """Two-stage HPSS split and random re-weighting (reference primitives.py:96-123) of a ``[B, T]`` batch.
    Returns ``(signal [B, n], n)``; parameters are drawn once per call (or injected through ``params``)."""
    x = _lib.require_cuda_f32(x)
    p = params or draw_recombine_params(include_residual)
    harm, perc, resid = hpss_split(x, p["n_fft1"], p["hop1"], p["margin1"], p["kernel1"])
    h1, p1, r1 = hpss_split(harm.contiguous(), p["n_fft2"], p["hop2"], p["margin2"], p["kernel2"])
    h2, p2, r2 = hpss_split(perc.contiguous(), p["n_fft2"], p["hop2"], p["margin2"], p["kernel2"])
    parts = [h1, p1, r1, h2, p2, r2, resid] if include_residual else [h1, p1, h2, p2]
    n = min(q.shape[1] for q in parts)
    stacked = torch.stack([q[:, :n] for q in parts], dim=0).contiguous()           # [P, B, n]
    out = torch.empty((x.shape[0], n), device=x.device, dtype=torch.float32)
    w1 = np.ascontiguousarray(p["w1"], dtype=np.float32)
    w2 = np.ascontiguousarray(p["w2"], dtype=np.float32)
    _lib.check(_lib.lib().mpcg_hpss_mix_f32(stacked.data_ptr(), out.data_ptr(), x.shape[0], n, len(parts), w1.ctypes.data,
                                            w2.ctypes.data, float(p["w_mix"]), _lib.stream_ptr(x)), "hpss mix")
    return out, n
