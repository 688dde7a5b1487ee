"""Zero-phase filters of ``signalproc/filters.py:44-90`` on the device: same names,
argument order and Nyquist-normalised cut-offs, ``[..., T]`` float32 CUDA tensors in and out.  The design stays with
SciPy on the host (cached); ``mpcg_sosfiltfilt_f32`` runs SciPy's ``sosfiltfilt`` recipe (odd extension, steady-state
initial conditions, forward and backward cascade passes) with float64 recurrence state."""
from __future__ import annotations

import functools

import numpy as np
import torch
from scipy import signal as _sig

from . import _lib


@functools.lru_cache(maxsize=256)
def _design(kind: str, args: tuple):
    if kind == "notch":
        b, a = _sig.iirnotch(*args)
        sos = np.concatenate([b, a])[None].astype(np.float64)
        edge = 3 * max(len(a), len(b))                       # scipy.signal.filtfilt default padlen
    else:
        order, wn, btype = args
        sos = np.asarray(_sig.butter(order, list(wn) if isinstance(wn, tuple) else wn, btype=btype, output="sos"), dtype=np.float64)
        ntaps = 2 * sos.shape[0] + 1
        ntaps -= min(int((sos[:, 2] == 0).sum()), int((sos[:, 5] == 0).sum()))
        edge = 3 * ntaps                                      # scipy.signal.sosfiltfilt default padlen
    zi = np.ascontiguousarray(_sig.sosfilt_zi(sos), dtype=np.float64)
    return np.ascontiguousarray(sos), zi, int(edge)


def _filtfilt(x: torch.Tensor, sos: np.ndarray, zi: np.ndarray, edge: int, epilogue: int = 0) -> torch.Tensor:
    x = _lib.require_cuda_f32(x)
    lead, t = x.shape[:-1], x.shape[-1]
    if t <= edge:
        raise ValueError(f"The length of the input vector x must be greater than padlen, which is {edge}.")
    rows = x.reshape(-1, t).contiguous()
    if sos.shape[0] > 6:
        raise ValueError("at most 6 second-order sections per call")
    out = torch.empty_like(rows)
    work = torch.empty((rows.shape[0], t + 2 * edge), device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().mpcg_sosfiltfilt_epi_f32(rows.data_ptr(), out.data_ptr(), work.data_ptr(), rows.shape[0], t,
                                                   sos.ctypes.data, sos.shape[0], zi.ctypes.data, edge, int(epilogue),
                                                   _lib.stream_ptr(x)), "zero-phase filter")
    return out.reshape(*lead, t)


def butter_bandpass(x: torch.Tensor, fs: float, low: float, high: float, order: int = 4) -> torch.Tensor:
    """Zero-phase Butterworth band-pass between ``low`` and ``high`` Hz (reference filters.py:44-48)."""
    nyq = 0.5 * fs
    return _filtfilt(x, *_design("butter", (int(order), (low / nyq, high / nyq), "bandpass")))


def butter_lowpass(x: torch.Tensor, fs: float, cutoff: float, order: int = 4) -> torch.Tensor:
    return _filtfilt(x, *_design("butter", (int(order), cutoff / (0.5 * fs), "lowpass")))


def butter_highpass(x: torch.Tensor, fs: float, cutoff: float, order: int = 4) -> torch.Tensor:
    return _filtfilt(x, *_design("butter", (int(order), cutoff / (0.5 * fs), "highpass")))


def band_stop(x: torch.Tensor, fs: float, low: float, high: float, order: int = 4) -> torch.Tensor:
    """Zero-phase Butterworth band-stop (reference filters.py:78-82)."""
    nyq = 0.5 * fs
    return _filtfilt(x, *_design("butter", (int(order), (low / nyq, high / nyq), "bandstop")))


def notch(x: torch.Tensor, fs: float, freq: float, q: float = 30.0) -> torch.Tensor:
    """Zero-phase IIR notch at ``freq`` Hz with quality factor ``q`` (reference filters.py:62-65)."""
    return _filtfilt(x, *_design("notch", (freq / (0.5 * fs), float(q))))


def notch_chain(x: torch.Tensor, fs: float, freqs, q: float = 55.0) -> torch.Tensor:
    """Several notches in sequence, e.g. mains hum + harmonics (reference filters.py:68-74)."""
    y = x
    for f in freqs:
        if f < 0.5 * fs:
            y = notch(y, fs, f, q)
    return y


# ------------------------------------------------------------------------------------------------ FIR band split
@functools.lru_cache(maxsize=32)
def _subband_kernels(fs: float, taps: int, edges: tuple):
    """The four Hamming-window FIR band filters of ``fir_subbands`` (reference filters.py:85-95) and, per band, the
    zero-phase kernel ``g = b (*) reversed(b)`` that ``filtfilt(b, [1.0])`` applies to the interior of a row."""
    nyq = 0.5 * fs
    e0, e1, e2 = edges
    bs = [_sig.firwin(taps, e0 / nyq, window="hamming", pass_zero="lowpass"),
          _sig.firwin(taps, [e0 / nyq, e1 / nyq], window="hamming", pass_zero="bandpass"),
          _sig.firwin(taps, [e1 / nyq, e2 / nyq], window="hamming", pass_zero="bandpass"),
          _sig.firwin(taps, e2 / nyq, window="hamming", pass_zero="highpass")]
    return [np.convolve(b, b[::-1]).astype(np.float32) for b in bs]


def decompose_bands(x: torch.Tensor, fs: float, taps: int = 61, edges=(45.0, 80.0, 200.0)) -> torch.Tensor:
    """``[..., T]`` -> ``[..., 4, T]``: zero-phase FIR sub-bands (LP / BP / BP / HP), reference filters.py:98-101
    (``scipy.signal.filtfilt(b, [1.0], x)`` per band).  ``filtfilt`` pads the row by odd extension (3 * taps samples a
    side), filters forward and backward and drops the padding; a FIR's memory (taps - 1) is shorter than the padding, so
    the interior equals ONE correlation of the extended row with the symmetric kernel ``b (*) reversed(b)`` -- which is
    what runs here (the per-row FIR kernel of ``amplitude_warp``, 2 * taps - 1 = 121 taps)."""
    x = _lib.require_cuda_f32(x)
    lead, t = x.shape[:-1], x.shape[-1]
    edge = 3 * taps
    if t <= edge:
        raise ValueError(f"The length of the input vector x must be greater than padlen, which is {edge}.")
    rows = x.reshape(-1, t)
    left = 2.0 * rows[:, :1] - rows[:, 1:edge + 1].flip(-1)
    right = 2.0 * rows[:, -1:] - rows[:, -edge - 1:-1].flip(-1)
    ext = torch.cat([left, rows, right], dim=-1).contiguous()                      # odd extension (host-side plumbing)
    out = torch.empty((rows.shape[0], 4, t), device=x.device, dtype=torch.float32)
    tmp = torch.empty_like(ext)
    for k, g in enumerate(_subband_kernels(float(fs), int(taps), tuple(float(e) for e in edges))):
        curves = torch.from_numpy(g).to(x.device).expand(rows.shape[0], -1).contiguous()
        _lib.check(_lib.lib().mpcg_aug_warp_f32(ext.data_ptr(), tmp.data_ptr(), ext.shape[0], ext.shape[1], curves.data_ptr(),
                                                g.shape[0], _lib.stream_ptr(x)), "FIR band split")
        out[:, k] = tmp[:, edge:edge + t]
    return out.reshape(*lead, 4, t)


def preprocess_four_bands(pcg: torch.Tensor, fs: float) -> torch.Tensor:
    """``[T]`` -> ``[T, 4]`` four-band PCG split (reference signalproc/preprocess.py:40-42)."""
    return decompose_bands(pcg.reshape(-1), fs).transpose(0, 1).contiguous()
