"""TEST INFRASTRUCTURE -- CPU restatement (NumPy float64) of the reference's time-varying sinc delay-and-sum
(``classify/beamformer.py:41-55``: ``_delay_channel`` and the sum of squares over microphones).  Pinned: checked
against ``tests/golden/beamformer.npz``, which ``oracle/make_golden.py`` produced by running the reference's own module
(outputs and autograd gradients, float64).  Only ``tests/`` may import this module."""
from __future__ import annotations

import numpy as np


def hamming(k: int) -> np.ndarray:
    """The module's taper buffer: ``torch.hamming_window(k, periodic=False)`` is built in float32 and stays float32-valued
    when the module is cast (beamformer.py:39), so the float32 values are what every precision of the reference uses."""
    import torch
    return torch.hamming_window(k, periodic=False).double().numpy()


def delay_channel(x: np.ndarray, delays: np.ndarray, kernel_size: int = 41) -> np.ndarray:
    """``x``, ``delays``: ``[B, T]`` -> ``[B, T]``."""
    half = kernel_size // 2
    tau = np.arange(-half, half + 1, dtype=np.float64)
    kern = np.sinc(tau[None, None, :] - delays[..., None]) * hamming(kernel_size)[None, None, :]
    kern = kern / kern.sum(axis=-1, keepdims=True)
    xp = np.pad(x, ((0, 0), (half, half)), mode="reflect")
    idx = np.arange(x.shape[1])[:, None] + np.arange(kernel_size)[None, :]
    return np.einsum("btk,btk->bt", xp[:, idx], kern)


def delay_and_sum(x: np.ndarray, delays: np.ndarray, kernel_size: int = 41) -> np.ndarray:
    """``[B, M, T]`` -> ``[B, T]``."""
    return sum(delay_channel(x[:, m], delays[:, m], kernel_size) ** 2 for m in range(x.shape[1]))
