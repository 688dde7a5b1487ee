"""Recording -> windows in one call: the composition the reference's loaders perform per record
(``datasets/cinc.py:86-94,115``, ``datasets/vest.py:50-51,84``:  ``preprocess_pcg`` / ``preprocess_ecg``
then ``segment``), batched on the device.

``preprocess_segment`` drives the fused cluster kernel (``mpcg_preprocess_segment_f32``): samples are read
from HBM once and windows written once.  When the geometry does not fit the fused kernel (row too long for
an 8-CTA cluster, a resampling ratio without a baked instance, ...) the same result is produced by chaining
the stand-alone CUDA entry points; there is no CPU path either way.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, design, torchproc
from .segment import start_index

_BANDS = {"pcg": torchproc.PCG_BAND, "ecg": torchproc.ECG_BAND}


def _kind_struct(kind: str, fs_out: float, despike: bool) -> _lib.ChainKind:
    low, high = _BANDS[kind]
    sos = np.concatenate([design.butter_sos(high / fs_out, "lowpass", 2), design.butter_sos(low / fs_out, "highpass", 2)])
    k = _lib.ChainKind()
    k.despike = 1 if (kind == "pcg" and despike) else 0
    k.n_sections = 2
    for i in range(2):
        for j in range(6):
            k.sos[i][j] = float(sos[i, j])
    return k


def preprocess_segment(x: torch.Tensor, fs_in: float, fs_out: float, spec, *, kinds=None, despike: bool = True,
                       mode: str = "torch", channels_last: bool = False, return_trace: bool = False,
                       trace_cap: int = 64, fused: bool | None = None):
    """``[B, T]`` -> ``[B, N, win]``  or  ``[B, C, T]`` -> ``[B, C, N, win]`` (``[B, N, win, C]`` with
    ``channels_last``): resample, (PCG only) Schmidt despike, band-limit, abs-max normalise, segment.

    ``kinds``: one of ``"pcg"`` / ``"ecg"`` per channel (default: all ``"pcg"``), e.g. ``("pcg", "ecg")`` for the
    Training-A pair.  ``fused``: ``None`` = fused kernel when the geometry allows, ``True`` = require it,
    ``False`` = always chain the stand-alone kernels.
    """
    torchproc._check_mode(mode)
    x = _lib.require_cuda_f32(x)
    if x.dim() not in (2, 3):
        raise ValueError("preprocess_segment expects [B, T] or [B, C, T]")
    planar_in = x.dim() == 2
    v = x[:, None] if planar_in else x
    b, c, t_in = v.shape
    kinds = tuple(kinds) if kinds is not None else ("pcg",) * c
    if len(kinds) != c or any(k not in _BANDS for k in kinds):
        raise ValueError(f"kinds must name 'pcg' or 'ecg' for each of the {c} channels")
    if channels_last and planar_in:
        raise ValueError("channels_last needs a [B, C, T] input")
    win, hop, start = spec.window_len(fs_out), spec.hop_len(fs_out), start_index(fs_out, spec)

    same_rate = fs_in == fs_out
    if same_rate:
        up = down = 1
        taps, off, depth, t_out = None, 0, 1, t_in
    else:
        up, down, taps, off, depth, t_out = torchproc._resample_plan(fs_in, fs_out, t_in, mode)
        if up == down:
            same_rate, t_out = True, t_in
    n = int(_lib.lib().mpcg_window_count(t_out, start, win, hop))
    shape = (b, n, win, c) if channels_last else ((b, n, win) if planar_in else (b, c, n, win))

    edits = trace = None
    if return_trace:
        edits = torch.zeros(b * c, dtype=torch.int32, device=x.device)
        trace = torch.full((b * c, trace_cap, 4), -1, dtype=torch.int32, device=x.device)

    if fused is not False and c <= 8:
        uniq = sorted(set(kinds))
        if len(uniq) <= 2:
            d = _lib.ChainDesc()
            d.t_in, d.t_out = t_in, t_out
            d.up, d.down, d.taps_per_phase, d.offset = up, down, depth, off
            d.taps = None if same_rate else taps.ctypes.data
            d.despike_win = int(round(float(fs_out) / 2.0))
            d.despike_threshold, d.despike_max_iterations = 3.0, 1000
            d.median_mode = _lib.MEDIAN_LOWER if mode == "torch" else _lib.MEDIAN_MEAN
            d.norm_flags = _lib.NORM_NAN_TO_NUM if mode == "torch" else _lib.NORM_PEAK_GT0
            d.seg_start, d.seg_win, d.seg_hop, d.seg_n = start, win, hop, n
            d.channels_last = 1 if channels_last else 0
            d.n_kinds = len(uniq)
            for i, k in enumerate(uniq):
                d.kinds[i] = _kind_struct(k, fs_out, despike)
            for ch in range(c):
                d.kind_of_channel[ch] = uniq.index(kinds[ch])
            out = torch.empty(shape, device=x.device, dtype=torch.float32)
            rc = _lib.lib().mpcg_preprocess_segment_f32(v.data_ptr(), out.data_ptr(), b, c, ctypes.byref(d),
                                                        _lib.ptr(edits), _lib.ptr(trace),
                                                        trace_cap if return_trace else 0, _lib.stream_ptr(v))
            if rc != _lib.EUNSUPPORTED:
                _lib.check(rc, "fused preprocess+segment")
                return (out, edits, trace) if return_trace else out
    if fused is True:
        raise ValueError("this geometry does not fit the fused kernel")

    # ---- chained stand-alone kernels (same arithmetic, more HBM traffic)
    chans = []
    for ch, kind in enumerate(kinds):
        col = v[:, ch].contiguous()
        if kind == "pcg":
            col = torchproc.resample(col, fs_in, fs_out, mode=mode)
            if despike:
                if return_trace:
                    col, e, tr = torchproc.remove_spikes(col, fs_out, mode=mode, return_trace=True, trace_cap=trace_cap)
                    edits.view(b, c)[:, ch] = e
                    trace.view(b, c, trace_cap, 4)[:, ch] = tr
                else:
                    col = torchproc.remove_spikes(col, fs_out, mode=mode)
            col = torchproc.abs_max_normalise(torchproc.bandpass_cascade(col, fs_out, *torchproc.PCG_BAND), mode=mode)
        else:
            col = torchproc.preprocess_ecg(col, fs_in, fs_out, mode=mode)
        chans.append(col)
    stacked = chans[0] if planar_in else torch.stack(chans, dim=1)
    out = torchproc.segment(stacked, fs_out, spec, channels_last=channels_last)
    return (out, edits, trace) if return_trace else out
