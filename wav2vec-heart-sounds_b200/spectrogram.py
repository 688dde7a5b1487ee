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
        # fast: False = the exact fp64 FMA tier (within 1e-5 of the float64 reference on ANY input); True = tensor cores when
        # the shape is eligible, else the fp32 FMA tier
        self.want_tc = bool(fast)
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
        self._basis_host = torch.from_numpy(basis.astype(np.float32) if self.fast else basis)      # (FMA tiers; the fallback of the tensor tier)
        self._fb_host = fb[k0:k1].contiguous()
        self._dev = {}
        # float64 tier on the fp64 tensor path (csrc/mel_dm.cu): the same basis in the kernel's fragment order,
        # [bin block of 32][slice of 16 samples][64 columns = (cos, -sin) per bin][16 + 4 pad]
        self._dm_host = None
        if not self.fast:
            nn16 = -(-self.win_length // 16) * 16
            full = np.zeros((nn16, 2, self.kpad), dtype=np.float64)
            full[:self.win_length] = basis
            dm = np.zeros((self.kpad // 32, nn16 // 16, 64, 20), dtype=np.float64)
            blk = full.reshape(nn16 // 16, 16, 2, self.kpad // 32, 32)                 # [slice, kk, part, block, bin]
            dm[..., :16] = blk.transpose(3, 0, 4, 2, 1).reshape(self.kpad // 32, nn16 // 16, 64, 16)   # col = 2 * bin + part
            nz = (self._fb_host != 0).numpy()                                            # [nbins, n_mels]
            rng = np.zeros((self.n_mels, 2), dtype=np.int32)
            for m in range(self.n_mels):
                ks = np.nonzero(nz[:, m])[0]
                if ks.size:
                    rng[m] = (ks[0], ks[-1] + 1)
            self._dm_host = (torch.from_numpy(dm), torch.from_numpy(rng))
        # tensor-core tier: n_fft = Q * hop and a full-length window.  Window and frame position are folded into Q bases
        # (one per position of a hop row inside a frame).  One launch takes 256 / (2 Q) weighted bins; presets with more --
        # the reference's own 4 kHz generator preset has 127 -- run as several launches over consecutive bin ranges whose
        # partial mel sums accumulate in the output, the dB map applied by the last one.
        self._tc = None
        q = self.n_fft // self.hop_length if self.hop_length else 0
        if (self.want_tc and self.win_length == self.n_fft and q * self.hop_length == self.n_fft and 1 <= q <= 8
                and self.hop_length % 16 == 0):
            per = 256 // (2 * q)
            parts = max(1, -(-self.nbins // per))
            cuts = [k0 + (self.nbins * i) // parts for i in range(parts + 1)]
            plans = [self._tc_plan(a, b - a, q, fb, w) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
            if plans and all(pl is not None for pl in plans):
                self._tc = dict(q=q, passes=plans, inv_norm=float(1.0 / norm))

    def _tc_plan(self, k0, nbins, q, fb, w):
        """Tables of one tensor-core launch over bins [k0, k0 + nbins), or None when the tile does not fit."""
        hop = self.hop_length
        step = 16 // math.gcd(16, q)                                              # Q * NQ must be a multiple of 16
        nq = -(-2 * nbins // max(step, 2)) * max(step, 2)
        ncols = q * nq
        ring = 4 * 2 * ncols * 32
        fits = (ncols <= 256 and ncols % 16 == 0
                and (2 * 128 * hop * 2 + ring + 1024 + 256 + nbins * ((self.n_mels + 3) // 4 * 4) * 4 + 8 * self.n_mels) <= 227 * 1024
                and (128 * (ncols + 1) * 4 + 128 * ((nbins + 1) | 1) * 4 <= 2 * 128 * hop * 2 + ring))
        if not fits:
            return None
        n = (np.arange(q)[:, None] * hop + np.arange(hop)[None, :]).astype(np.float64)           # [q, j]: sample index in the frame
        kk = (k0 + np.arange(nbins, dtype=np.float64))
        ang = 2.0 * np.pi * ((n[:, None, :] * kk[None, :, None]) % self.n_fft) / self.n_fft        # [q, bin, j]
        e = np.zeros((q, nq, hop), dtype=np.float64)
        e[:, :nbins] = w[n.astype(int)][:, None, :] * np.cos(ang)
        e[:, nq // 2:nq // 2 + nbins] = -w[n.astype(int)][:, None, :] * np.sin(ang)
        e = e.reshape(ncols, hop)
        hi = e.astype(np.float16)
        lo = (e - hi.astype(np.float64)).astype(np.float16)
        ksteps = hop // 16
        packed = np.zeros((ksteps, 2, ncols * 16), dtype=np.float16)               # per K = 16 step: hi chunk, lo chunk
        n_idx, j_idx = np.meshgrid(np.arange(ncols), np.arange(16), indexing="ij")
        off = ((n_idx >> 3) * 256 + (j_idx >> 3) * 128 + (n_idx & 7) * 16 + (j_idx & 7) * 2) // 2
        for ks in range(ksteps):
            packed[ks, 0, off.ravel()] = hi[:, 16 * ks:16 * ks + 16].ravel()
            packed[ks, 1, off.ravel()] = lo[:, 16 * ks:16 * ks + 16].ravel()
        return dict(k0=int(k0), nbins=int(nbins), ncols_q=int(nq), basis=torch.from_numpy(packed),
                    fb=fb[k0:k0 + nbins].contiguous())

    def _tables(self, device):
        key = str(device)
        if key not in self._dev:
            tc = None if self._tc is None else ([(pl["basis"].to(device), pl["fb"].to(device)) for pl in self._tc["passes"]],)
            dm = None if self._dm_host is None else tuple(v.to(device) for v in self._dm_host)
            self._dev[key] = (self._basis_host.to(device), self._fb_host.to(device), tc, dm)
        return self._dev[key]

    @property
    def backend(self) -> str:
        return "tcgen05 split-fp16" if self._tc is not None else ("fma fp32" if self.fast else "dmma fp64")

    def num_frames(self, t: int) -> int:
        return 1 + t // self.hop_length

    _MAX_ROWS = 65535                                 # rows of one launch (a grid dimension of the FMA / DMMA kernels)

    def _run(self, signal: torch.Tensor, log_map: bool) -> torch.Tensor:
        x = _lib.require_cuda_f32(signal)
        lead, t = x.shape[:-1], x.shape[-1]
        rows = x.reshape(-1, t)
        if rows.shape[0] > self._MAX_ROWS:                # bigger batches: consecutive launches over row blocks
            parts = [self._run(rows[i:i + self._MAX_ROWS], log_map) for i in range(0, rows.shape[0], self._MAX_ROWS)]
            return torch.cat(parts, dim=0).reshape(*lead, self.n_mels, parts[0].shape[-1])
        if t <= self.n_fft // 2:
            raise ValueError(f"signal of {t} samples is too short for reflect padding of {self.n_fft // 2}")
        frames = self.num_frames(t)
        basis, fb, tc, dm = self._tables(x.device)
        out = torch.empty((rows.shape[0], self.n_mels, frames), device=x.device, dtype=torch.float32)
        if tc is not None:
            passes, rc = self._tc["passes"], 0
            for i, (pl, (basis_tc, fb_tc)) in enumerate(zip(passes, tc[0])):
                # flags: bit 0 = dB map (last launch only), bit 1 = add to what the earlier launches left in `out`
                flags = (1 if (log_map and i == len(passes) - 1) else 0) | (2 if i else 0)
                rc = _lib.lib().mpcg_mel_tc_f32(rows.data_ptr(), out.data_ptr(), rows.shape[0], t, self.n_fft, self.hop_length,
                                                pl["k0"], pl["nbins"], pl["ncols_q"], basis_tc.data_ptr(), fb_tc.data_ptr(),
                                                self._tc["inv_norm"], self.n_mels, frames, flags, _lib.stream_ptr(x))
                if rc != 0:
                    break
            if rc != _lib.EUNSUPPORTED:
                _lib.check(rc, "mel (tensor cores)")
                return out.reshape(*lead, self.n_mels, frames)
        if dm is not None:
            rc = _lib.lib().mpcg_mel_dm_f32(rows.data_ptr(), out.data_ptr(), rows.shape[0], t, self.n_fft, self.hop_length,
                                            self.n_lo, self.n_hi, self.nbins, self.kpad, dm[0].data_ptr(), fb.data_ptr(),
                                            dm[1].data_ptr(), self.n_mels, frames, 1 if log_map else 0, _lib.stream_ptr(x))
            if rc != _lib.EUNSUPPORTED:
                _lib.check(rc, "mel (fp64 tensor path)")
                return out.reshape(*lead, self.n_mels, frames)
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
        """Default: the contraction runs in float64 and stays within 1e-5 of the float64 reference on any input.

        ``fast=True``: when ``n_fft`` is a multiple of ``hop_length`` and the window spans ``n_fft`` the windowed DFT runs on
        the tensor cores (tcgen05, split-fp16 operands, fp32 accumulation, the window applied in the time domain through
        windowed bases; see csrc/mel_tc.cu), 13x faster; other shapes take the fp32 FMA tier.  fp32 accumulation resolves a
        frame's spectrum to ~1e-6 of its largest bin.  The dB map spans 80 dB, and a 1e-5 bound at its lower end needs every
        bin to ~4e-8 of the frame's peak -- float64 territory -- so this tier is within 1e-5 wherever a frame's mel values
        lie within ~40 dB of its peak (the reference's 4 kHz generator preset passes white noise at 2e-6 everywhere; at the 16 kHz
        preset, whose mel filters sit on single bins, ~1e-5 of white-noise elements exceed 1e-5, max 4e-5) and up to ~3e-4 off
        close to the -80 dB floor of strongly tonal frames."""
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
