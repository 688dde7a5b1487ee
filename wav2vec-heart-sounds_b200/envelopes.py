"""Envelope extraction on device rows (reference ``signalproc/envelopes.py:11-23``).

``hilbert_envelope`` is the magnitude of SciPy's analytic signal -- an N-point DFT of the whole row, negative
frequencies masked, inverse DFT -- for any row length up to 2**19 samples; both transforms run as fp64 Bluestein chirp
convolutions through four-step shared-memory FFTs (``csrc/envelope.cu``).  ``homomorphic_envelope`` low-passes the log
envelope with the zero-phase Butterworth of ``filters.py`` and exponentiates in that kernel's store.  Inputs are float32
CUDA tensors ``[..., T]``; no CPU or PyTorch fallback.
"""
from __future__ import annotations

import torch

from . import _lib, filters

__all__ = ["hilbert_envelope", "homomorphic_envelope"]

_WORK_BUDGET = 4 << 30            # bytes of fp64 workspace per launch group; longer batches run in chunks of rows


def _envelope(x: torch.Tensor, flags: int) -> torch.Tensor:
    x = _lib.require_cuda_f32(x)
    if x.dim() == 0:
        raise ValueError("envelopes need at least one dimension")
    shape, t = x.shape, int(x.shape[-1])
    rows = x.reshape(-1, t).contiguous() if t else x.reshape(0, 0)
    out = torch.empty_like(rows)
    if rows.numel() == 0:
        return out.reshape(shape)
    lib = _lib.lib()
    one = lib.mpcg_hilbert_work_bytes(1, t)
    if one < 0:
        raise ValueError(f"rows of {t} samples are beyond the envelope kernel's range (2**19)")
    fixed = lib.mpcg_hilbert_work_bytes(0, t)
    chunk = int(max(1, min(rows.shape[0], 65535, (_WORK_BUDGET - fixed) // (one - fixed))))
    work = torch.empty(fixed + chunk * (one - fixed), device=x.device, dtype=torch.uint8)
    for r0 in range(0, rows.shape[0], chunk):
        part = rows[r0:r0 + chunk]
        _lib.check(lib.mpcg_hilbert_envelope_f32(part.data_ptr(), out[r0:r0 + chunk].data_ptr(), work.data_ptr(), work.numel(),
                                                 part.shape[0], t, int(flags), _lib.stream_ptr(x)), "hilbert envelope")
    return out.reshape(shape)


def hilbert_envelope(x: torch.Tensor) -> torch.Tensor:
    """Analytic-signal amplitude envelope (``envelopes.py:11-13``)."""
    return _envelope(x, 0)


def homomorphic_envelope(x: torch.Tensor, fs: float, cutoff: float = 8.0, order: int = 6) -> torch.Tensor:
    """``exp(zero-phase low-pass(log(max(envelope, eps))))`` (``envelopes.py:16-23``), the classic envelogram."""
    if cutoff >= 0.5 * fs:
        raise ValueError(f"cutoff {cutoff} Hz is above Nyquist for fs={fs}")
    log_env = _envelope(x, _lib.ENV_LOG)
    sos, zi, edge = filters._design("butter", (int(order), cutoff / (0.5 * fs), "lowpass"))
    return filters._filtfilt(log_env, sos, zi, edge, epilogue=_lib.EPI_EXP)
