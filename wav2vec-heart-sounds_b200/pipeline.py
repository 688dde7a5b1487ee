"""Recording -> windows in one call: the composition the reference's loaders perform per record
(``datasets/cinc.py:86-94,115``, ``datasets/vest.py:50-51,84``:  ``preprocess_pcg`` / ``preprocess_ecg``
then ``segment``), batched on the device.

``preprocess_segment`` drives the fused row-streaming kernel (``mpcg_preprocess_segment_f32``): one launch, rows of
any length (and, with ``lengths``, a different length per recording).  When the geometry does not fit the fused kernel
(a resampling ratio without a baked tap set, ...) the same result is produced by chaining the stand-alone CUDA entry
points; there is no CPU path either way.
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


def _out_lengths(lengths: np.ndarray, fs_in: float, fs_out: float, mode: str) -> np.ndarray:
    """Resampled length of every recording, with the oracle's own rounding."""
    if fs_in == fs_out:
        return lengths.copy()
    up, down = design.reduce_ratio(fs_in, fs_out)
    if up == down:
        return lengths.copy()
    f = design.sinc_out_len if mode == "torch" else design.kaiser_out_len
    return np.array([f(int(t), up, down) for t in lengths], dtype=np.int64)


def _preprocess_segment_ragged(v: torch.Tensor, lengths, fs_in, fs_out, spec, kinds, despike, mode, channels_last,
                               channel_major, planar_in, return_edits):
    """One launch over recordings of different lengths: ``v`` is ``[B, C, Tmax]`` (rows padded to a common pitch) and
    ``lengths[b]`` the valid samples of recording ``b``.  Returns the windows of all recordings back to back --
    ``[sum N, win]`` (mono), ``[sum N, win, C]`` (``channels_last``) or ``[C, sum N, win]`` (``channel_major``) -- and
    the window count of each recording (host int64 array)."""
    b, c, pitch = v.shape
    lengths = np.asarray(lengths.cpu() if torch.is_tensor(lengths) else lengths, dtype=np.int64).reshape(-1)
    if lengths.shape[0] != b or (lengths < 0).any() or (lengths > pitch).any():
        raise ValueError("lengths must hold one value in [0, T] per recording")
    if not (planar_in or channels_last or channel_major):
        raise ValueError("a ragged batch of multichannel recordings needs channels_last or channel_major")
    win, hop, start = spec.window_len(fs_out), spec.hop_len(fs_out), start_index(fs_out, spec)
    t_max = int(lengths.max()) if b else 0
    same_rate = fs_in == fs_out
    if same_rate:
        up = down = 1
        taps, off, depth = None, 0, 1
    else:
        up, down, taps, off, depth, _ = torchproc._resample_plan(fs_in, fs_out, max(t_max, 1), mode)
        same_rate = up == down
    t_out = _out_lengths(lengths, fs_in, fs_out, mode)
    counts = np.array([int(_lib.lib().mpcg_window_count(int(t), start, win, hop)) for t in t_out], dtype=np.int64)
    total = int(counts.sum())
    first = np.concatenate([[0], np.cumsum(counts)[:-1]]) if b else np.zeros(0, np.int64)
    if planar_in:
        shape, unit = (total, win), win
    elif channels_last:
        shape, unit = (total, win, c), win * c
    else:
        shape, unit = (c, total, win), win
    out = torch.empty(shape, device=v.device, dtype=torch.float32)
    edits = torch.zeros(b * c, dtype=torch.int32, device=v.device) if return_edits else None
    if b == 0 or total == 0:
        return (out, counts, edits) if return_edits else (out, counts)
    uniq = sorted(set(kinds))
    if len(uniq) > 2 or c > 8:
        raise ValueError("at most two channel kinds and eight channels per launch")
    d = _lib.ChainDesc()
    d.t_in, d.t_out = pitch, max(int(t_out.max()), 1)
    d.up, d.down, d.taps_per_phase, d.offset = up, down, depth, off
    d.taps = None if same_rate else taps.ctypes.data
    d.despike_win = int(round(float(fs_out) / 2.0))
    d.despike_threshold, d.despike_max_iterations = 3.0, 1000
    d.median_mode = _lib.MEDIAN_LOWER if mode == "torch" else _lib.MEDIAN_MEAN
    d.norm_flags = _lib.NORM_NAN_TO_NUM if mode == "torch" else _lib.NORM_PEAK_GT0
    d.seg_start, d.seg_win, d.seg_hop, d.seg_n = start, win, hop, 0
    d.channels_last = 2 if channel_major else (1 if channels_last else 0)
    d.n_kinds = len(uniq)
    for i, k in enumerate(uniq):
        d.kinds[i] = _kind_struct(k, fs_out, despike)
    for ch in range(c):
        d.kind_of_channel[ch] = uniq.index(kinds[ch])
    tabs = torch.from_numpy(np.stack([lengths, t_out]).astype(np.int32)).to(v.device)              # one small upload
    offs = torch.from_numpy((first * unit).astype(np.int64)).to(v.device)
    d.row_t_in, d.row_t_out, d.row_out_offset = tabs[0].data_ptr(), tabs[1].data_ptr(), offs.data_ptr()
    d.plane_elems = total * win
    nbytes = int(_lib.lib().mpcg_preprocess_segment_work_bytes(int(d.t_out)))
    work = _lib.workspace(v, nbytes)
    rc = _lib.lib().mpcg_preprocess_segment_f32(v.data_ptr(), out.data_ptr(), b, c, ctypes.byref(d), work.data_ptr(), work.numel(),
                                                _lib.ptr(edits), None, 0, _lib.stream_ptr(v))
    if rc == _lib.EUNSUPPORTED:
        raise ValueError("this geometry does not fit the fused kernel")
    _lib.check(rc, "fused preprocess+segment (ragged)")
    return (out, counts, edits) if return_edits else (out, counts)


def preprocess_segment(x: torch.Tensor, fs_in: float, fs_out: float, spec, *, kinds=None, despike: bool = True,
                       mode: str = "torch", channels_last: bool = False, return_trace: bool = False,
                       trace_cap: int = 64, fused: bool | None = None, out: torch.Tensor | None = None,
                       return_edits: bool = False, channel_major: bool = False, lengths=None):
    """``[B, T]`` -> ``[B, N, win]``  or  ``[B, C, T]`` -> ``[B, C, N, win]`` (``[B, N, win, C]`` with
    ``channels_last``): resample, (PCG only) Schmidt despike, band-limit, abs-max normalise, segment.

    ``kinds``: one of ``"pcg"`` / ``"ecg"`` per channel (default: all ``"pcg"``), e.g. ``("pcg", "ecg")`` for the
    Training-A pair.  ``fused``: ``None`` = fused kernel when the geometry allows, ``True`` = require it,
    ``False`` = always chain the stand-alone kernels.  ``out``: optional preallocated result tensor.
    ``return_trace``: also return ``(edits[B*C], trace[B*C, cap, 4])``, the despike passes in the reference's order
    (this takes the fused kernel's serial despike path); ``return_edits``: also return the pass counts alone
    (fused kernel only; the parallel despike rounds stay on).  ``channel_major``: ``[B, C, T]`` -> ``[C, B, N, win]`` (every
    channel's windows form one contiguous ``[B * N, win]`` batch, ready for ``augment_pcg_batch``; fused kernel only).
    ``lengths``: valid samples of each recording of a padded ``[B, (C,) Tmax]`` batch (a ragged batch, still ONE launch);
    the result is then ``(windows, counts)`` with all recordings' windows back to back (``[sum N, win]``,
    ``[sum N, win, C]`` with ``channels_last``, ``[C, sum N, win]`` with ``channel_major``) and ``counts[b]`` windows of
    recording ``b`` (the tensor rule: a recording shorter than the start pad still yields one zero window).
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
    if mode == "numpy":                                      # the NumPy chains bridge NaNs first (signalproc/preprocess.py:25,34)
        lens_dev = None
        if lengths is not None:
            lens_dev = torch.as_tensor(np.asarray(lengths.cpu() if torch.is_tensor(lengths) else lengths), dtype=torch.int32).to(v.device)
        v = torchproc.fill_nans(v.contiguous(), lens_dev)
    if lengths is not None:
        if return_trace or fused is False or out is not None:
            raise ValueError("lengths= (a ragged batch) is a fused-kernel call without trace / out")
        return _preprocess_segment_ragged(v.contiguous(), lengths, fs_in, fs_out, spec, kinds, despike, mode, channels_last,
                                          channel_major, planar_in, return_edits)
    if channels_last and planar_in:
        raise ValueError("channels_last needs a [B, C, T] input")
    if channel_major and (planar_in or channels_last or fused is False):
        raise ValueError("channel_major needs a [B, C, T] input, the fused kernel and excludes channels_last")
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
    if channel_major:
        shape = (c, b, n, win)

    edits = trace = None
    if return_trace:
        edits = torch.zeros(b * c, dtype=torch.int32, device=x.device)
        trace = torch.full((b * c, trace_cap, 4), -1, dtype=torch.int32, device=x.device)

    if return_edits and not return_trace:
        if fused is False:
            raise ValueError("return_edits without return_trace is a fused-kernel option")
        edits = torch.zeros(b * c, dtype=torch.int32, device=x.device)
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
            d.channels_last = 2 if channel_major else (1 if channels_last else 0)
            d.n_kinds = len(uniq)
            for i, k in enumerate(uniq):
                d.kinds[i] = _kind_struct(k, fs_out, despike)
            for ch in range(c):
                d.kind_of_channel[ch] = uniq.index(kinds[ch])
            if out is None:
                out = torch.empty(shape, device=x.device, dtype=torch.float32)
            elif tuple(out.shape) != tuple(shape) or not out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous():
                raise ValueError(f"out must be a contiguous CUDA float32 tensor of shape {tuple(shape)}")
            nbytes = int(_lib.lib().mpcg_preprocess_segment_work_bytes(max(int(t_out), 1)))
            work = _lib.workspace(v, nbytes)
            rc = _lib.lib().mpcg_preprocess_segment_f32(v.data_ptr(), out.data_ptr(), b, c, ctypes.byref(d),
                                                        work.data_ptr(), work.numel(), _lib.ptr(edits), _lib.ptr(trace),
                                                        trace_cap if return_trace else 0, _lib.stream_ptr(v))
            if rc != _lib.EUNSUPPORTED:
                _lib.check(rc, "fused preprocess+segment")
                if return_trace:
                    return out, edits, trace
                return (out, edits) if return_edits else out
    if fused is True or channel_major or (return_edits and not return_trace):
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
            col = torchproc.preprocess_ecg(col, fs_in, fs_out, mode=mode, fused=False)
        chans.append(col)
    stacked = chans[0] if planar_in else torch.stack(chans, dim=1)
    res = torchproc.segment(stacked, fs_out, spec, channels_last=channels_last)
    if out is not None:
        out.copy_(res)
        res = out
    return (res, edits, trace) if return_trace else res


class HostPipeline:
    """Host buffers in, host buffers out: the call a loader makes per batch of recordings.

    The batch is cut into chunks; chunk i's host->device copy, chunk i-1's kernel(s) and chunk i-2's device->host
    copy run concurrently on three CUDA streams (PCIe is full duplex), so a step costs about
    max(H2D, D2H) instead of their sum.  Buffers are allocated once and reused.

    ``augment`` (an ``AugmentConfig`` or ``True``): the windows of channel 0 (the PCG) also go through the fused
    ``augment_pcg_batch`` chain before they leave the device; the output is then channel-major ``[C, B, N, win]``.
    """

    def __init__(self, recordings: int, channels: int, t_in: int, fs_in: float, fs_out: float, spec, *, kinds=None,
                 mode: str = "torch", channels_last: bool = False, chunk: int = 128, device="cuda", augment=None):
        self.augment = augment
        if augment is not None and channels_last:
            raise ValueError("augment uses the channel-major layout; channels_last is not available with it")
        self.args = dict(fs_in=fs_in, fs_out=fs_out, spec=spec, kinds=kinds, mode=mode, channels_last=channels_last)
        if augment is not None:
            self.args["channel_major"] = True
        self.fs_out = fs_out
        self.device = torch.device(device)
        self.chunk = min(chunk, recordings)
        self.recordings, self.channels, self.t_in = recordings, channels, t_in
        probe = preprocess_segment(torch.zeros(1, channels, t_in, device=self.device), **self.args)
        nbuf = 3
        self.dev_in = [torch.empty(self.chunk, channels, t_in, device=self.device) for _ in range(nbuf)]
        if augment is None:
            self.out_shape = (recordings,) + tuple(probe.shape[1:])
            self.dev_out = [torch.empty((self.chunk,) + tuple(probe.shape[1:]), device=self.device) for _ in range(nbuf)]
        else:
            self.win_shape = tuple(probe.shape[2:])                                   # (N, win)
            self.out_shape = (channels, recordings) + self.win_shape
            self.dev_out = [torch.empty((channels, self.chunk) + self.win_shape, device=self.device) for _ in range(nbuf)]
            self.dev_aug = [torch.empty((self.chunk,) + self.win_shape, device=self.device) for _ in range(nbuf)]
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(self.device) for _ in range(3))
        self.h2d_bytes = recordings * channels * t_in * 4
        self.d2h_bytes = int(torch.tensor(self.out_shape).prod()) * 4

    def empty_output(self) -> torch.Tensor:
        return torch.empty(self.out_shape, dtype=torch.float32).pin_memory()

    def __call__(self, x_host: torch.Tensor, out_host: torch.Tensor) -> torch.Tensor:
        """x_host [recordings, channels, t_in] (pinned for speed) -> out_host (pinned), both on the host."""
        if x_host.is_cuda or out_host.is_cuda:
            raise ValueError("HostPipeline takes host tensors; use preprocess_segment for device tensors")
        from . import torchaug
        n = self.recordings
        nbuf = len(self.dev_in)
        copied = [torch.cuda.Event() for _ in range(nbuf)]
        ran = [torch.cuda.Event() for _ in range(nbuf)]
        drained = [torch.cuda.Event() for _ in range(nbuf)]
        entry = torch.cuda.current_stream(self.device)
        cfg = None if self.augment in (None, True) else self.augment
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(entry)
        for i, lo in enumerate(range(0, n, self.chunk)):
            hi = min(lo + self.chunk, n)
            m = hi - lo
            b = i % nbuf
            with torch.cuda.stream(self.s_in):
                if i >= nbuf:
                    self.s_in.wait_event(ran[b])              # the kernel that read this buffer is done
                self.dev_in[b][:m].copy_(x_host[lo:hi], non_blocking=True)
                copied[b].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(copied[b])
                if i >= nbuf:
                    self.s_run.wait_event(drained[b])         # the copy-out that read this buffer is done
                if self.augment is None:
                    preprocess_segment(self.dev_in[b][:m], out=self.dev_out[b][:m], **self.args)
                else:
                    dout = self.dev_out[b].view(-1)[: self.channels * m * self.win_shape[0] * self.win_shape[1]]
                    dout = dout.view((self.channels, m) + self.win_shape)
                    preprocess_segment(self.dev_in[b][:m], out=dout, **self.args)
                    daug = self.dev_aug[b][:m]
                    torchaug.augment_pcg_batch(dout[0].reshape(m * self.win_shape[0], self.win_shape[1]), self.fs_out, cfg,
                                               noise="philox", out=daug.view(m * self.win_shape[0], self.win_shape[1]))
                ran[b].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ran[b])
                if self.augment is None:
                    out_host[lo:hi].copy_(self.dev_out[b][:m], non_blocking=True)
                else:
                    out_host[0, lo:hi].copy_(daug, non_blocking=True)
                    for c in range(1, self.channels):
                        out_host[c, lo:hi].copy_(dout[c], non_blocking=True)
                drained[b].record(self.s_out)
        for s in (self.s_in, self.s_run, self.s_out):
            entry.wait_stream(s)
        done = torch.cuda.Event()
        done.record(self.s_out)
        done.synchronize()                                   # host buffers out: the last device->host copy has landed
        return out_host
