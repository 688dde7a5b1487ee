"""Time-varying sinc delay-and-sum beamformer on the device (reference ``classify/beamformer.py``): the gather that
collapses a multichannel PCG batch ``[B, M, T]`` to one channel, with its backward pass, behind the reference's module
name.  The per-sample delays still come from the reference's small transformer (plain ``torch.nn``, as upstream); what
is replaced is ``_delay_channel`` + the sum of squares (``beamformer.py:41-55``): the reference materialises a
``[B, T, 41]`` kernel tensor and an unfolded copy of the signal per microphone, here one kernel reads ``x`` and the delays
and writes ``[B, T]`` (and one more kernel produces both gradients).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib


def _window(kernel_size: int) -> np.ndarray:
    return torch.hamming_window(kernel_size, periodic=False).numpy().astype(np.float32)      # beamformer.py:39


class _DelayAndSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, delays, kernel_size):
        x = _lib.require_cuda_f32(x)
        delays = _lib.require_cuda_f32(delays, "delays")
        if x.dim() != 3 or delays.shape != x.shape:
            raise ValueError("delay_and_sum takes x and delays of shape [B, M, T]")
        b, m, t = x.shape
        win = _window(kernel_size)
        out = torch.empty((b, t), device=x.device, dtype=torch.float32)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        aux = torch.empty((b, m, t, 4), device=x.device, dtype=torch.float32) if need else None
        _lib.check(_lib.lib().mpcg_beamform_fwd_f32(x.data_ptr(), delays.data_ptr(), out.data_ptr(), _lib.ptr(aux), b, m, t,
                                                    win.ctypes.data, int(kernel_size), _lib.stream_ptr(x)), "beamformer forward")
        ctx.kernel_size = int(kernel_size)
        ctx.save_for_backward(delays, aux) if need else None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        delays, aux = ctx.saved_tensors
        b, m, t = delays.shape
        g = grad_out.contiguous().float()
        gx = torch.empty_like(delays) if ctx.needs_input_grad[0] else None
        gd = torch.empty_like(delays) if ctx.needs_input_grad[1] else None
        win = _window(ctx.kernel_size)
        _lib.check(_lib.lib().mpcg_beamform_bwd_f32(delays.data_ptr(), aux.data_ptr(), g.data_ptr(), _lib.ptr(gx), _lib.ptr(gd), b, m,
                                                    t, win.ctypes.data, ctx.kernel_size, _lib.stream_ptr(delays)),
                   "beamformer backward")
        return gx, gd, None


def delay_and_sum(x: torch.Tensor, delays: torch.Tensor, kernel_size: int = 41) -> torch.Tensor:
    """``sum_m _delay_channel(x[:, m], delays[:, m]) ** 2`` (reference beamformer.py:41-55), differentiable in both."""
    return _DelayAndSum.apply(x, delays, kernel_size)


class _DelayPredictor(nn.Module):
    """The reference's delay predictor (beamformer.py:15-28), upstream ``torch.nn`` modules."""

    def __init__(self, num_mics: int, d_model: int = 32, nhead: int = 4, num_layers: int = 2):
        super().__init__()
        self.input_proj = nn.Conv1d(num_mics, d_model, kernel_size=1)
        layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=nhead, dim_feedforward=64, batch_first=True)
        self.encoder = nn.TransformerEncoder(layer, num_layers=num_layers)
        self.output_proj = nn.Linear(d_model, num_mics)

    def forward(self, x):
        h = self.input_proj(x).transpose(1, 2)
        return self.output_proj(self.encoder(h)).transpose(1, 2)


class TimeVaryingSincBeamformer(nn.Module):
    """Drop-in for the reference module (same constructor, same parameter names, so its state dicts load)."""

    def __init__(self, num_mics: int, fs: float, max_delay_s: float = 0.01, kernel_size: int = 41):
        super().__init__()
        self.num_mics = num_mics
        self.max_delay_samples = max_delay_s * fs
        self.kernel_size = kernel_size
        self.half_k = kernel_size // 2
        self.delay_predictor = _DelayPredictor(num_mics)
        self.register_buffer("t_idx", torch.arange(-self.half_k, self.half_k + 1).float())
        self.register_buffer("window", torch.hamming_window(kernel_size, periodic=False))

    def forward(self, x: torch.Tensor) -> torch.Tensor:      # (B, M, T) -> (B, T)
        delays = torch.clamp(self.delay_predictor(x), 0.0, self.max_delay_samples)
        return delay_and_sum(x.contiguous(), delays.contiguous(), self.kernel_size)
