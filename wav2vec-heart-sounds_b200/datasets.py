"""Batched dataset building on the device: the caller side of the hot path.

The reference builds its in-memory fragment lists one record at a time on the CPU --
``datasets/cinc.py:54-125`` (``build_fragments``: read record -> ``preprocess_pcg`` [-> ``preprocess_ecg``] ->
``segment`` -> one ``Fragment`` per window) and ``datasets/vest.py:54-113`` (six PCG channels) -- and hands them to
``datasets/fragments.py:30-83`` (``FragmentDataset``).  Here the records of one call are grouped by input rate only
(records of any lengths ride in one ragged batch), each group goes through ONE fused preprocess + segment launch (``pipeline.preprocess_segment``, NumPy-path arithmetic
by default because that is what the loaders call), and the windows land in one tensor in the reference's order with
label / record index tensors beside it.  File reading (wfdb / WAV) stays with the caller: records arrive as arrays.

``FragmentTensorDataset`` serves the same item dictionaries as the reference's ``FragmentDataset`` (so ``pad_collate``
and the trainers work unchanged); ``device_batch_transform`` is the ``SupervisedTrainer(batch_transform=...)`` hook
(``classify/trainer.py:65-68``) that runs the fused augmentation chain on the PCG channel of a device batch.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Sequence

import numpy as np
import torch

from . import _lib, torchaug
from .pipeline import preprocess_segment
from .segment import start_index


@dataclass
class FragmentBatch:
    """All windows of a set of records.  ``windows``: ``[F, win]`` (mono) or ``[F, win, C]`` (the loader layout of
    ``datasets/cinc.py:90`` / ``datasets/vest.py:83-85``); ``labels`` / ``record``: ``[F]`` int64, ``record`` indexing
    ``patients``; windows of a record are contiguous and records keep the caller's order, like the reference's list."""
    windows: torch.Tensor
    labels: torch.Tensor
    record: torch.Tensor
    patients: list
    fs: int

    def __len__(self) -> int:
        return int(self.windows.shape[0])

    def class_counts(self) -> dict:
        vals, cnt = torch.unique(self.labels, return_counts=True)
        return {int(v): int(c) for v, c in zip(vals.tolist(), cnt.tolist())}


def build_fragments_batched(records: Iterable, *, fs_out: int, window, ecg: bool = False, channels: Sequence[str] | None = None,
                            mode: str = "numpy", device="cuda") -> FragmentBatch:
    """``records``: iterable of ``(signal, fs, label, patient)`` with ``signal`` a ``[T]`` or ``[T, C]`` array (time
    major, as ``wfdb`` / the WAV readers return it).

    * CinC (``datasets/cinc.py``): column 0 is the PCG; with ``ecg=True`` column 1 is the ECG and fragments are
      ``[win, 2]``; otherwise fragments are mono ``[win]``.
    * Vest (``datasets/vest.py``): pass ``channels=("pcg",) * 6`` -- every column is conditioned as PCG, fragments are
      ``[win, 6]``.

    Records shorter than the start pad yield no fragment in ``mode="numpy"`` (``signalproc/segment.py:34-35``) and one
    zero window in ``mode="torch"`` (``signalproc/torchproc.py:119-129``), as in the reference.
    """
    recs = [(np.asarray(s), float(fs), int(label), patient) for s, fs, label, patient in records]
    dev = torch.device(device)
    kinds = tuple(channels) if channels is not None else (("pcg", "ecg") if ecg else ("pcg",))
    c = len(kinds)
    mono = c == 1
    win = window.window_len(fs_out)
    groups: dict = {}
    for i, (sig, fs, _, _) in enumerate(recs):
        sig2 = sig[:, None] if sig.ndim == 1 else sig
        if sig2.shape[1] < c:
            raise ValueError(f"record {i} has {sig2.shape[1]} channels, {c} are needed")
        groups.setdefault(fs, []).append(i)                  # one launch per input rate, whatever the record lengths
    counts = [0] * len(recs)
    pieces: dict = {}
    for fs, idx in groups.items():
        if mode == "numpy":                                  # segment() returns no window for these (signalproc/segment.py:34-35)
            idx = [i for i in idx if _numpy_out_len(recs[i][0].shape[0], fs, fs_out) > start_index(fs_out, window)]
        if not idx:
            continue
        lens = np.array([recs[i][0].shape[0] for i in idx], dtype=np.int64)
        pitch = (int(lens.max()) + 3) & ~3                   # rows start on 16-byte boundaries: the vector resampler path
        host = np.zeros((len(idx), c, pitch), dtype=np.float32)
        for k, i in enumerate(idx):
            sig = recs[i][0]
            host[k, :, :sig.shape[0]] = (sig[:, None] if sig.ndim == 1 else sig)[:, :c].T
        x = torch.from_numpy(host).to(dev)                                            # [B, C, Tmax]
        w, n_each = preprocess_segment(x[:, 0] if mono else x, fs, fs_out, window, kinds=kinds, mode=mode,
                                       channels_last=not mono, lengths=lens)
        at = 0
        for k, i in enumerate(idx):
            n = int(n_each[k])
            pieces[i] = w[at:at + n]
            counts[i] = n
            at += n
    total = sum(counts)
    shape = (total, win) if mono else (total, win, c)
    windows = torch.empty(shape, device=dev, dtype=torch.float32)
    labels = torch.empty(total, dtype=torch.int64)
    record = torch.empty(total, dtype=torch.int64)
    at = 0
    for i, n in enumerate(counts):
        if n:
            windows[at:at + n] = pieces[i]
            labels[at:at + n] = recs[i][2]
            record[at:at + n] = i
            at += n
    return FragmentBatch(windows, labels.to(dev), record.to(dev), [r[3] for r in recs], int(fs_out))


def _numpy_out_len(t_in: int, fs_in: float, fs_out: float) -> int:
    from . import design
    if fs_in == fs_out:
        return t_in
    up, down = design.reduce_ratio(fs_in, fs_out)
    return t_in if up == down else design.kaiser_out_len(t_in, up, down)


class FragmentTensorDataset(torch.utils.data.Dataset):
    """The reference's ``FragmentDataset`` (``datasets/fragments.py:30-83``) served from a :class:`FragmentBatch`: the same
    item dictionaries (``waveform``, ``label``, ``patient``), the same list of items -- every fragment once, followed by
    its augmented copies, ``int(round(augment_num * max_count / count[label]))`` of them when ``balance`` is on
    (``fragments.py:47-56``) -- and augmentation applied lazily, so every epoch sees fresh draws.  ``channel`` selects
    one column of multichannel fragments (-1 keeps all).  Augmented items run the fused ``augment_pcg_batch`` chain on
    the PCG column; :meth:`get_batch` does that for a whole batch of indices in ONE launch."""

    def __init__(self, batch: FragmentBatch, channel: int = -1, *, augment_num: int = 0, cfg=None, balance: bool = True,
                 pcg_channel: int = 0, noise: str | None = "philox"):
        self.batch, self.channel = batch, channel
        self.cfg, self.pcg_channel, self.noise = cfg, pcg_channel, noise
        labels = batch.labels.tolist()
        counts: dict = {}
        for lab in labels:
            counts[lab] = counts.get(lab, 0) + 1
        top = max(counts.values()) if counts else 1
        frag, aug = [], []
        for i, lab in enumerate(labels):
            copies = 0
            if augment_num > 0:
                copies = int(round(augment_num * top / counts[lab])) if balance else augment_num
            frag.extend([i] * (1 + copies))
            aug.extend([False] + [True] * copies)
        self._frag = torch.tensor(frag, dtype=torch.int64)
        self._aug = torch.tensor(aug, dtype=torch.bool)

    @property
    def labels(self) -> list:
        return self.batch.labels[self._frag.to(self.batch.labels.device)].tolist()

    def __len__(self) -> int:
        return int(self._frag.numel())

    def get_batch(self, indices) -> dict:
        """Items ``indices`` as one batch: ``waveform`` ``[B, win]`` or ``[B, win, C]`` on the fragments' device, the
        augmented ones through one fused launch."""
        idx = torch.as_tensor(indices, dtype=torch.int64)
        frag = self._frag[idx]
        dev = self.batch.windows.device
        w = self.batch.windows[frag.to(dev)].clone()
        todo = torch.nonzero(self._aug[idx]).flatten().to(dev)
        if todo.numel():
            if w.dim() == 2:
                w[todo] = torchaug.augment_pcg_batch(w[todo].contiguous(), self.batch.fs, self.cfg, noise=self.noise)
            else:
                col = w[todo][:, :, self.pcg_channel].contiguous()
                w[todo, :, self.pcg_channel] = torchaug.augment_pcg_batch(col, self.batch.fs, self.cfg, noise=self.noise)
        if w.dim() == 3 and self.channel != -1:
            w = w[:, :, self.channel]
        rec = self.batch.record[frag.to(self.batch.record.device)].tolist()
        return {"waveform": w, "label": self.batch.labels[frag.to(self.batch.labels.device)],
                "patient": [self.batch.patients[r] for r in rec]}

    def __getitem__(self, idx: int) -> dict:
        one = self.get_batch([int(idx)])
        return {"waveform": one["waveform"][0], "label": int(one["label"][0]), "patient": one["patient"][0]}


def device_augment_fn(cfg=None, *, pcg_channel: int = 0, noise: str | None = "philox", device="cuda"):
    """An ``augment_fn`` for the reference's own ``FragmentDataset`` (``datasets/fragments.py:19,36,68-72``:
    ``Callable[[np.ndarray, int], np.ndarray]``): one window ``[T]`` or ``[T, C]`` goes to the device, its PCG column
    through the fused chain, and comes back as float32 -- the drop-in for callers that keep the per-item dataset."""

    def fn(wave: np.ndarray, fs: int) -> np.ndarray:
        x = torch.from_numpy(np.ascontiguousarray(wave, dtype=np.float32)).to(device)
        if x.dim() == 1:
            return torchaug.augment_pcg_batch(x[None], int(fs), cfg, noise=noise)[0].cpu().numpy()
        out = x.clone()
        out[:, pcg_channel] = torchaug.augment_pcg_batch(x[:, pcg_channel][None].contiguous(), int(fs), cfg, noise=noise)[0]
        return out.cpu().numpy()

    return fn


def device_batch_transform(fs: int, cfg=None, *, pcg_channel: int = 0, noise: str | None = "philox"):
    """A ``batch_transform`` for ``SupervisedTrainer`` (``classify/trainer.py:65-68``): ``x`` is the device batch,
    ``[B, T]`` or time-major ``[B, T, C]`` (``pad_collate``'s layout); the PCG column goes through the fused
    ``augment_pcg_batch`` chain, other columns pass through."""

    def transform(x: torch.Tensor) -> torch.Tensor:
        x = _lib.require_cuda_f32(x)
        if x.dim() == 2:
            return torchaug.augment_pcg_batch(x.contiguous(), fs, cfg, noise=noise)
        if x.dim() != 3:
            raise ValueError("batch_transform expects [B, T] or [B, T, C]")
        out = x.clone()
        out[:, :, pcg_channel] = torchaug.augment_pcg_batch(x[:, :, pcg_channel].contiguous(), fs, cfg, noise=noise)
        return out

    return transform


def condition_generator_batch(reference: torch.Tensor, conditioning: torch.Tensor, fs: int, mel_transform, crop_frames: int,
                              hop_length: int, *, fade: int = 128, chirp: bool = True, cycles=None,
                              fade_ms: float = 10.0) -> dict:
    """The per-item work of the generator datasets for a whole batch on the device (SURVEY.md section 8f, ranks 2 and
    4; reference ``datasets/generative.py:62-115``): ``abs_max_normalise`` -> [cardiac-cycle rebuild] -> fade ->
    ``fit_length(crop_frames * hop_length)`` for the reference and the conditioning waveforms ``[B, T]``, ``log_mel``
    of the conditioning cut or zero-padded to ``crop_frames`` frames, and the chirp reference plot signal.  Keys as in
    the reference's item dictionary.

    ``cycles``: optional, one entry per row -- the row's cardiac cycles as ``(start, end)`` pairs in the order they
    are to be joined (``heart_cycles.cycle_bounds`` reordered by ``heart_cycles.rearrange_order``), or ``None`` /
    fewer than two cycles for a row that keeps its waveform (``generative.py:67-68,86-87``).  Both waveforms of a row
    are cut at the same joins and crossfaded over ``round(fade_ms / 1000 * fs)`` samples (``generative.py:56``)."""
    ref = _lib.require_cuda_f32(reference, "reference")
    con = _lib.require_cuda_f32(conditioning, "conditioning")
    if ref.dim() != 2 or con.dim() != 2 or ref.shape[0] != con.shape[0]:
        raise ValueError("reference and conditioning must be [B, T] batches of the same size")
    b, crop = ref.shape[0], int(crop_frames) * int(hop_length)
    from .spectrogram import log_mel

    if cycles is not None:
        from . import heart_cycles, torchproc
        if len(cycles) != b:
            raise ValueError("cycles needs one entry per row")
        if ref.shape[1] != con.shape[1]:
            raise ValueError("reference and conditioning must have the same length to be cut at the same joins")
        plan = heart_cycles.CyclePlan([list(c) if c is not None and len(c) >= 2 else None for c in cycles], ref.shape[1],
                                      ref.device)
        fade_n = int(round(fade_ms / 1000.0 * fs))

    def run(x, want_chirp):
        x = x.contiguous()
        lengths, flags = None, _lib.NORM_PEAK_GT0
        if cycles is not None:
            x, lengths = heart_cycles.rebuild_batch(torchproc.abs_max_normalise(x, mode="numpy"), plan, crop, fade_n)
            flags |= _lib.GEN_NO_NORM
        y = torch.empty((b, crop), device=x.device, dtype=torch.float32)
        c = torch.empty((b, crop), device=x.device, dtype=torch.float32) if want_chirp else None
        _lib.check(_lib.lib().mpcg_gen_condition_rows_f32(x.data_ptr(), y.data_ptr(), _lib.ptr(c), _lib.ptr(lengths), b,
                                                          x.shape[1], crop, int(fade), float(fs), flags, _lib.stream_ptr(x)),
                   "generator conditioning")
        return y, c

    ref_y, chirp_y = run(ref, chirp)
    con_y, _ = run(con, False)
    spec = log_mel(con_y, mel_transform)
    if spec.shape[-1] >= crop_frames:
        spec = spec[..., :crop_frames].contiguous()
    else:
        spec = torch.nn.functional.pad(spec, (0, crop_frames - spec.shape[-1]))
    out = {"ref_audio": ref_y, "con_audio": con_y, "con_spec": spec, "seg_wave": ref_y.clone()}
    if chirp:
        out["chirp_wave"] = chirp_y
    return out
