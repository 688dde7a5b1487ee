"""Drop-in for ``mpcg_wav2vec.signalproc.torchproc`` on B200: same names, argument order,
tensor layouts and degenerate-size behaviour (reference ``signalproc/torchproc.py:22-129``), with
every function body replaced by one call into ``libmpcg_b200.so``.

Differences a caller can observe, all deliberate:

* inputs must be CUDA float32 tensors (anything else raises: there is no CPU fallback);
* ``segment`` returns a contiguous copy instead of an ``unfold`` view (callers only read it);
* a keyword-only ``mode`` selects which of the reference's two disagreeing oracles is followed:
  ``"torch"`` (default; ``torchproc`` run in float64: sinc/Hann resampler, lower-median despike,
  ``nan_to_num`` + ``clamp_min(1e-12)`` normalise) or ``"numpy"`` (``signalproc.preprocess``:
  SciPy Kaiser ``resample_poly``, mean-of-middles median with the zero-median stop, ``peak > 0``).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, design
from .segment import WindowSpec, start_index

PCG_BAND = (25.0, 450.0)
ECG_BAND = (2.0, 40.0)

_MODES = ("torch", "numpy")


def _check_mode(mode: str) -> str:
    if mode not in _MODES:
        raise ValueError(f"mode must be one of {_MODES}, got {mode!r}")
    return mode


def _as_rows(x: torch.Tensor):
    """View ``[..., T]`` as contiguous ``[rows, T]``; remember the leading shape."""
    x = _lib.require_cuda_f32(x)
    lead = x.shape[:-1]
    return x.reshape(-1, x.shape[-1]) if x.dim() != 2 else x, lead


def fill_nans(x: torch.Tensor, lengths: torch.Tensor | None = None) -> torch.Tensor:
    """NaN runs bridged by linear interpolation between the nearest valid samples, edges held (reference
    ``signalproc/normalize.py:11-17``, the first step of the NumPy chains).  Returns ``x`` itself when it holds no NaN
    (one device reduction), else a repaired copy.  ``lengths``: int32 ``[B]`` valid samples per recording of ``[B, C, T]``."""
    x = _lib.require_cuda_f32(x)
    if x.numel() == 0 or not bool(torch.isnan(x).any()):
        return x
    y = x.clone()
    rows = y.reshape(-1, y.shape[-1])
    c = 1 if lengths is None else int(rows.shape[0] // max(int(lengths.numel()), 1))
    b = rows.shape[0] // c
    _lib.check(_lib.lib().mpcg_fill_nans_f32(rows.data_ptr(), b, c, rows.shape[1], _lib.ptr(lengths), _lib.stream_ptr(rows)),
               "fill NaNs")
    return y


def _host_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _cascade(x: torch.Tensor, sos: np.ndarray) -> torch.Tensor:
    rows, lead = _as_rows(x)
    out = torch.empty_like(rows)
    sos = _host_f64(sos)
    _lib.check(_lib.lib().mpcg_biquad_cascade_f32(rows.data_ptr(), out.data_ptr(), rows.shape[0], rows.shape[1],
                                                  sos.ctypes.data, sos.shape[0], _lib.stream_ptr(rows)),
               "biquad cascade")
    return out.reshape(*lead, rows.shape[1])


def lowpass(x: torch.Tensor, fs: float, cutoff: float, order: int = 2) -> torch.Tensor:
    """Causal Butterworth low-pass, cut-off normalised by fs (reference torchproc.py:42-44)."""
    return _cascade(x, design.butter_sos(cutoff / fs, "lowpass", int(order)))


def highpass(x: torch.Tensor, fs: float, cutoff: float, order: int = 2) -> torch.Tensor:
    """Causal Butterworth high-pass (reference torchproc.py:47-49)."""
    return _cascade(x, design.butter_sos(cutoff / fs, "highpass", int(order)))


def bandpass_cascade(x: torch.Tensor, fs: float, low: float, high: float, order: int = 2) -> torch.Tensor:
    """Low-pass at ``high`` then high-pass at ``low`` in ONE kernel pass (reference torchproc.py:52-53)."""
    sos = np.concatenate([design.butter_sos(high / fs, "lowpass", int(order)),
                          design.butter_sos(low / fs, "highpass", int(order))])
    return _cascade(x, sos)


def _resample_plan(fs_in: float, fs_out: float, t_in: int, mode: str):
    up, down = design.reduce_ratio(fs_in, fs_out)
    if mode == "torch":
        g, off, depth = design.sinc_hann_frames(up, down)
        t_out = design.sinc_out_len(t_in, up, down)
    else:
        g, off, depth = design.kaiser_poly_frames(up, down, t_in)
        t_out = design.kaiser_out_len(t_in, up, down)
    return up, down, np.ascontiguousarray(g, dtype=np.float32), off, depth, t_out


def resample(x: torch.Tensor, fs_in: float, fs_out: float, *, mode: str = "torch") -> torch.Tensor:
    """Rational resampling of the last dim (reference torchproc.py:56-59; identity when rates match)."""
    _check_mode(mode)
    if fs_in == fs_out:
        return x
    rows, lead = _as_rows(x)
    up, down, taps, off, depth, t_out = _resample_plan(fs_in, fs_out, rows.shape[1], mode)
    if up == down:                                           # rounded rates coincide: nothing to do
        return x
    out = torch.empty((rows.shape[0], t_out), device=rows.device, dtype=torch.float32)
    _lib.check(_lib.lib().mpcg_resample_f32(rows.data_ptr(), out.data_ptr(), rows.shape[0], rows.shape[1], t_out,
                                            taps.ctypes.data, up, down, depth, off, _lib.stream_ptr(rows)),
               "resample")
    return out.reshape(*lead, t_out)


def abs_max_normalise(x: torch.Tensor, *, mode: str = "torch") -> torch.Tensor:
    """Zero-mean, divide by the peak magnitude, clamp to [-1, 1] per row (reference torchproc.py:62-66)."""
    _check_mode(mode)
    if mode == "numpy":
        x = fill_nans(x)                                     # normalize.abs_max_normalise interpolates NaNs first (:25)
    rows, lead = _as_rows(x)
    out = torch.empty_like(rows)
    flags = _lib.NORM_NAN_TO_NUM if mode == "torch" else _lib.NORM_PEAK_GT0
    _lib.check(_lib.lib().mpcg_absmax_norm_f32(rows.data_ptr(), out.data_ptr(), rows.shape[0], rows.shape[1], flags,
                                               _lib.stream_ptr(rows)), "abs-max normalise")
    return out.reshape(*lead, rows.shape[1])


def remove_spikes(x: torch.Tensor, fs: float, threshold: float = 3.0, max_iterations: int = 1000, *,
                  mode: str = "torch", return_trace: bool = False, trace_cap: int = 64):
    """Batched Schmidt spike removal over ``[B, T]`` with 500 ms frames (reference torchproc.py:69-98).

    The input is never modified.  ``return_trace=True`` also returns ``(edits[B], trace[B, cap, 4])``
    int32 tensors holding ``(frame, peak, lo, hi)`` of each flattening pass, for parity checks.
    """
    _check_mode(mode)
    x = _lib.require_cuda_f32(x)
    squeeze = x.dim() == 1
    rows = (x[None] if squeeze else x).clone()
    if rows.dim() != 2:
        raise ValueError("remove_spikes expects [T] or [B, T]")
    b, t = rows.shape
    win = round(float(fs) / 2.0)
    edits = trace = None
    if return_trace:
        edits = torch.zeros(b, dtype=torch.int32, device=rows.device)
        trace = torch.full((b, trace_cap, 4), -1, dtype=torch.int32, device=rows.device)
    _lib.check(_lib.lib().mpcg_despike_f32(rows.data_ptr(), b, t, int(win), float(threshold), int(max_iterations),
                                           _lib.MEDIAN_LOWER if mode == "torch" else _lib.MEDIAN_MEAN,
                                           _lib.ptr(edits), _lib.ptr(trace), trace_cap if return_trace else 0,
                                           _lib.stream_ptr(rows)), "despike")
    out = rows[0] if squeeze else rows
    return (out, edits, trace) if return_trace else out


class _WholeRow:
    """A 'window' that is the whole conditioned row: lets the fused preprocess+segment kernel serve preprocess_pcg/ecg."""
    start_pad_s = 0.0

    def __init__(self, t_out: int):
        self.t_out = int(t_out)

    def window_len(self, fs) -> int:
        return self.t_out

    def hop_len(self, fs) -> int:
        return self.t_out


def _fused_rows(x: torch.Tensor, fs_in: float, fs_out: float, kind: str, despike: bool, mode: str):
    """One fused launch (resample -> [despike] -> band -> normalise, the row written once) or None when the row does
    suit the fused kernel, e.g. a resampling ratio without baked taps (the caller then chains the stand-alone kernels:
    same arithmetic, more HBM traffic)."""
    from .pipeline import preprocess_segment           # late: pipeline imports this module
    lead, t_in = x.shape[:-1], x.shape[-1]
    if t_in == 0 or x.numel() == 0:
        return None
    if fs_in == fs_out:
        t_out = t_in
    else:
        up, down, _, _, _, t_out = _resample_plan(fs_in, fs_out, t_in, mode)
        if up == down:
            t_out = t_in
    rows = x.reshape(-1, t_in).contiguous()
    try:
        out = preprocess_segment(rows, fs_in, fs_out, _WholeRow(t_out), kinds=(kind,), despike=despike, mode=mode, fused=True)
    except ValueError as exc:
        if "does not fit the fused kernel" in str(exc):
            return None
        raise
    return out.reshape(*lead, t_out)


def preprocess_pcg(x: torch.Tensor, fs_in: float, fs_out: float, *, despike: bool = True,
                   mode: str = "torch", fused: bool | None = None) -> torch.Tensor:
    """resample -> Schmidt despike -> 25-450 Hz band (fs-normalised) -> abs-max normalise
    (reference torchproc.py:101-108).  ``fused``: ``None`` = one fused cluster-kernel launch when the row fits it,
    ``False`` = always the four stand-alone kernels, ``True`` = require the fused launch."""
    x = _lib.require_cuda_f32(x)
    if mode == "numpy":
        x = fill_nans(x)                                     # signalproc/preprocess.py:25
    if fused is not False:
        out = _fused_rows(x, fs_in, fs_out, "pcg", despike, mode)
        if out is not None:
            return out
        if fused is True:
            raise ValueError("this geometry does not fit the fused kernel")
    squeeze = x.dim() == 1
    v = resample(x[None] if squeeze else x, fs_in, fs_out, mode=mode)
    if despike:
        v = remove_spikes(v, fs_out, mode=mode)
    v = abs_max_normalise(bandpass_cascade(v, fs_out, *PCG_BAND, order=2), mode=mode)
    return v[0] if squeeze else v


def preprocess_ecg(x: torch.Tensor, fs_in: float, fs_out: float, *, mode: str = "torch",
                   fused: bool | None = None) -> torch.Tensor:
    """resample -> 2-40 Hz band (fs-normalised) -> abs-max normalise (reference torchproc.py:111-116).
    ``fused`` as in :func:`preprocess_pcg`."""
    x = _lib.require_cuda_f32(x)
    if mode == "numpy":
        x = fill_nans(x)                                     # signalproc/preprocess.py:34
    if fused is not False:
        out = _fused_rows(x, fs_in, fs_out, "ecg", False, mode)
        if out is not None:
            return out
        if fused is True:
            raise ValueError("this geometry does not fit the fused kernel")
    squeeze = x.dim() == 1
    v = resample(x[None] if squeeze else x, fs_in, fs_out, mode=mode)
    v = abs_max_normalise(bandpass_cascade(v, fs_out, *ECG_BAND, order=2), mode=mode)
    return v[0] if squeeze else v


def window_count(t: int, fs: float, spec) -> int:
    return int(_lib.lib().mpcg_window_count(int(t), start_index(fs, spec), spec.window_len(fs), spec.hop_len(fs)))


def segment(x: torch.Tensor, fs: float, spec: WindowSpec, *, channels_last: bool = False) -> torch.Tensor:
    """``[B, T]`` -> ``[B, N, win]`` (``[T]`` -> ``[N, win]``, ``[B, C, T]`` -> ``[B, C, N, win]``), reference
    torchproc.py:119-129.  ``channels_last=True`` turns ``[B, C, T]`` into the loaders' ``[B, N, win, C]``
    (reference signalproc/segment.py:40-52, datasets/cinc.py:90)."""
    x = _lib.require_cuda_f32(x)
    squeeze = x.dim() == 1
    v = x[None] if squeeze else x
    if v.dim() not in (2, 3):
        raise ValueError("segment expects [T], [B, T] or [B, C, T]")
    if channels_last and v.dim() != 3:
        raise ValueError("channels_last needs a [B, C, T] input")
    b = v.shape[0]
    c = v.shape[1] if v.dim() == 3 else 1
    t = v.shape[-1]
    win, hop, start = spec.window_len(fs), spec.hop_len(fs), start_index(fs, spec)
    n = int(_lib.lib().mpcg_window_count(t, start, win, hop))
    if n < 0:
        raise ValueError("invalid window geometry")
    if channels_last:
        out = torch.empty((b, n, win, c), device=v.device, dtype=torch.float32)
    elif v.dim() == 3:
        out = torch.empty((b, c, n, win), device=v.device, dtype=torch.float32)
    else:
        out = torch.empty((b, n, win), device=v.device, dtype=torch.float32)
    _lib.check(_lib.lib().mpcg_segment_f32(v.data_ptr(), out.data_ptr(), b, c, t, start, win, hop, n,
                                           1 if channels_last else 0, _lib.stream_ptr(v)), "segment")
    return out[0] if squeeze else out
