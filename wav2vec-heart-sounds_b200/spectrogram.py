"""Drop-in for ``mpcg_wav2vec.signalproc.spectrogram`` (reference ``signalproc/spectrogram.py:13-45``):
``MelConfig(...).build()`` returns a callable ``transform(signal[..., T]) -> [..., n_mels, frames]`` with the
arithmetic of ``torchaudio.transforms.MelSpectrogram(power=1, normalized=True)``, and ``log_mel`` maps it to the
[0, 1] dB scale the diffusion conditioner uses.  Both run as one CUDA kernel (framing + windowed DFT of the
bins that carry mel weight + magnitude + mel projection [+ dB map]).

Design-time pieces (the Hann window, the HTK filterbank) are produced by the same torch / torchaudio helpers the
reference's transform uses, on the host, once per config.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


class MelTransform:
    """Callable mel-spectrogram with torchaudio ``MelSpectrogram(power=1.0, normalized=True)`` semantics."""

    def __init__(self, sample_rate, n_fft, hop_length, win_length, n_mels, f_min, f_max, fast=False):
        self.fast = bool(fast)
        import torchaudio.functional as AF
        self.sample_rate, self.n_fft, self.hop_length = int(sample_rate), int(n_fft), int(hop_length)
        self.win_length, self.n_mels = int(win_length), int(n_mels)
        self.f_min, self.f_max = float(f_min), float(f_max)
        if self.win_length > self.n_fft:
            raise ValueError("win_length must not exceed n_fft")
        n_freqs = self.n_fft // 2 + 1
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")                   # all-zero filters at 16 kHz are expected (and reproduced)
            fb = AF.melscale_fbanks(n_freqs, self.f_min, self.f_max, self.n_mels, self.sample_rate,
                                    norm=None, mel_scale="htk")                     # [n_freqs, n_mels] float32
        used = torch.nonzero(fb.abs().sum(dim=1) > 0).flatten()
        k0 = int(used[0]) if used.numel() else 0
        k1 = int(used[-1]) + 1 if used.numel() else 1
        self.k0, self.nbins = k0, k1 - k0
        self.kpad = (self.nbins + 31) // 32 * 32
        # periodic Hann of win_length, centred inside n_fft (torch.stft's padding rule)
        w = torch.hann_window(self.win_length, periodic=True).double().numpy()
        left = (self.n_fft - self.win_length) // 2
        self.n_lo, self.n_hi = left, left + self.win_length
        norm = math.sqrt(float(np.sum(w ** 2)))
        n = np.arange(self.n_lo, self.n_hi, dtype=np.float64)[:, None]
        k = np.arange(k0, k1, dtype=np.float64)[None, :]
        ang = 2.0 * np.pi * ((n * k) % self.n_fft) / self.n_fft
        basis = np.zeros((self.win_length, 2, self.kpad), dtype=np.float64)
        basis[:, 0, :self.nbins] = w[:, None] * np.cos(ang) / norm
        basis[:, 1, :self.nbins] = -w[:, None] * np.sin(ang) / norm
        self._basis_host = torch.from_numpy(basis.astype(np.float32) if self.fast else basis)
        self._fb_host = fb[k0:k1].contiguous()
        self._dev = {}

    def _tables(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (self._basis_host.to(device), self._fb_host.to(device))
        return self._dev[key]

    def num_frames(self, t: int) -> int:
        return 1 + t // self.hop_length

    def _run(self, signal: torch.Tensor, log_map: bool) -> torch.Tensor:
        x = _lib.require_cuda_f32(signal)
        lead, t = x.shape[:-1], x.shape[-1]
        rows = x.reshape(-1, t)
        if t <= self.n_fft // 2:
            raise ValueError(f"signal of {t} samples is too short for reflect padding of {self.n_fft // 2}")
        frames = self.num_frames(t)
        basis, fb = self._tables(x.device)
        out = torch.empty((rows.shape[0], self.n_mels, frames), device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().mpcg_mel_f32(rows.data_ptr(), out.data_ptr(), rows.shape[0], t, self.n_fft, self.hop_length,
                                           self.n_lo, self.n_hi, self.nbins, self.kpad, basis.data_ptr(),
                                           0 if self.fast else 1, fb.data_ptr(),
                                           self.n_mels, frames, 1 if log_map else 0, _lib.stream_ptr(x)), "mel")
        return out.reshape(*lead, self.n_mels, frames)

    def __call__(self, signal: torch.Tensor) -> torch.Tensor:
        return self._run(signal, False)

    def log_mel(self, signal: torch.Tensor) -> torch.Tensor:
        return self._run(signal, True)


@dataclass(frozen=True)
class MelConfig:
    """Parameters of a conditioning mel-spectrogram (mirror of reference spectrogram.py:13-38)."""
    sample_rate: int
    n_fft: int
    hop_length: int
    win_length: int | None = None
    n_mels: int = 80
    f_min: float = 0.125
    f_max: float = 500.0

    def build(self, fast: bool = False) -> MelTransform:
        """``fast=True`` runs the DFT in float32 (about twice the throughput; error ~1e-7 of a frame's largest
        bin, which only shows in leakage skirts near the dB map's 1e-5 floor).  The default keeps the contraction
        in float64 and stays within 1e-5 of the float64 reference on any input."""
        return MelTransform(self.sample_rate, self.n_fft, self.hop_length, self.win_length or self.n_fft, self.n_mels,
                            self.f_min, self.f_max, fast=fast)


def log_mel(signal: torch.Tensor, transform) -> torch.Tensor:
    """Mel-spectrogram in dB shifted / scaled into [0, 1] (reference spectrogram.py:41-45).  With a transform built
    by :class:`MelConfig` everything is one kernel; any other callable's output goes through the dB-map kernel."""
    if isinstance(transform, MelTransform):
        return transform.log_mel(signal)
    mel = _lib.require_cuda_f32(transform(signal))
    out = torch.empty_like(mel)
    _lib.check(_lib.lib().mpcg_logmap_f32(mel.data_ptr(), out.data_ptr(), mel.numel(), _lib.stream_ptr(mel)), "log map")
    return out
