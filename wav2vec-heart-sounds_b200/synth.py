"""Deterministic synthetic PCG / ECG recordings generated on any torch device.

PCG: S1/S2 Gaussian-windowed tone bursts at a per-recording heart rate, white noise, a per-recording gain and
0-3 short large spikes so the despiker does real work.  ECG: Gaussian QRS complexes at the same kind of beat
grid, 0.3 Hz drift, white noise.  No NaNs.  Used by bench.py, the tests and smoke(); not part of the hot path.
"""
from __future__ import annotations

import math

import torch


def _beat_phase(t, period, offset):
    """Signed time to the nearest beat centre, beats at offset + k*period."""
    tau = torch.remainder(t - offset + 0.5 * period, period) - 0.5 * period
    return tau


def synth_pcg(rows: int, n: int, fs: float, *, seed: int = 1234, device="cpu", spikes: bool = True) -> torch.Tensor:
    g = torch.Generator(device=device).manual_seed(seed)
    u = lambda *s: torch.rand(*s, device=device, generator=g)
    t = torch.arange(n, device=device, dtype=torch.float32)[None] / fs
    period = 60.0 / (60.0 + 40.0 * u(rows, 1))
    f1 = 40.0 + 80.0 * u(rows, 1)
    f2 = 40.0 + 80.0 * u(rows, 1)
    tau1 = _beat_phase(t, period, 0.1)
    tau2 = _beat_phase(t, period, 0.1 + 0.35 * period)
    x = torch.exp(-0.5 * (tau1 / 0.020) ** 2) * torch.sin(2 * math.pi * f1 * tau1)
    x = x + 0.6 * torch.exp(-0.5 * (tau2 / 0.015) ** 2) * torch.sin(2 * math.pi * f2 * tau2)
    x = x + 0.05 * torch.randn(rows, n, device=device, generator=g)
    if spikes and n > 64:
        count = torch.randint(0, 4, (rows,), device=device, generator=g)
        reach = torch.arange(10, device=device)[None]
        for k in range(3):
            at = torch.randint(16, n - 16, (rows, 1), device=device, generator=g)
            width = torch.randint(3, 11, (rows, 1), device=device, generator=g)
            amp = (5.0 + 15.0 * u(rows, 1)) * torch.where(u(rows, 1) < 0.5, -1.0, 1.0)
            on = ((reach < width) & (count[:, None] > k)).float()
            x.scatter_add_(1, at + reach, amp * on)
    return (x * (0.1 + 1.9 * u(rows, 1))).contiguous()


def synth_ecg(rows: int, n: int, fs: float, *, seed: int = 4321, device="cpu") -> torch.Tensor:
    g = torch.Generator(device=device).manual_seed(seed)
    t = torch.arange(n, device=device, dtype=torch.float32)[None] / fs
    period = 60.0 / (60.0 + 40.0 * torch.rand(rows, 1, device=device, generator=g))
    tau = _beat_phase(t, period, 0.1)
    x = torch.exp(-0.5 * (tau / 0.010) ** 2) + 0.2 * torch.sin(2 * math.pi * 0.3 * t)
    return (x + 0.02 * torch.randn(rows, n, device=device, generator=g)).contiguous()


def synth_pair(recordings: int, n: int, fs: float, *, seed: int = 1234, device="cpu") -> torch.Tensor:
    """[recordings, 2, n]: channel 0 PCG, channel 1 ECG (Training-A layout)."""
    return torch.stack([synth_pcg(recordings, n, fs, seed=seed, device=device),
                        synth_ecg(recordings, n, fs, seed=seed + 1, device=device)], dim=1).contiguous()
