"""The amplitude normalisers of ``signalproc/normalize.py`` other than abs-max, on device rows (SURVEY.md section 8f,
rank 3).  ``abs_max_normalise`` itself is ``torchproc.abs_max_normalise``.

Names, argument order and defaults follow the reference (``signalproc/normalize.py:33-78``).  Inputs are float32 CUDA
tensors ``[..., T]``; every function returns a new tensor of the same shape.

* The reference's NumPy functions take ONE signal, its tensor functions reduce over the whole tensor
  (``x.max() - x.min()``, ``topk(...).values.mean()``).  With the default ``per_row=False`` a batched input is treated
  exactly like that -- one range for everything, identical to the reference on the same tensor; ``per_row=True`` is the
  batched form (every row of the last dimension is its own signal), which is what a loader that holds many recordings
  wants.  For a 1-D input the two coincide.
* ``z_normalise_torch`` is per row in the reference already (``dim=-1``).

The selection of the k largest / smallest samples is an exact three-level radix select on the device
(``csrc/rownorm.cu``); there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["minmax_normalise", "minmax_normalise_torch", "z_normalise", "z_normalise_torch", "kpeak_normalise",
           "kpeak_normalise_torch", "row_statistics"]


def _run(x: torch.Tensor, mode: int, k: int, lo: float, hi: float, flags: int, want_stats: bool = False):
    x = _lib.require_cuda_f32(x)
    if x.dim() == 0:
        raise ValueError("normalisers need at least one dimension")
    shape = x.shape
    t = int(shape[-1])
    rows = x.numel() // t if t else 0
    xc = x.contiguous()
    y = torch.empty_like(xc)
    glob = bool(flags & _lib.RN_GLOBAL)
    stats = None
    if glob or want_stats:
        stats = torch.empty((rows + 1, 8), device=x.device, dtype=torch.float64)
    _lib.check(_lib.lib().mpcg_row_normalise_f32(xc.data_ptr(), y.data_ptr(), _lib.ptr(stats), rows, t, int(mode), int(k),
                                                 float(lo), float(hi), int(flags), _lib.stream_ptr(x)), "row normalise")
    return (y.view(shape), stats) if want_stats else y.view(shape)


def minmax_normalise(x: torch.Tensor, lo: float = -1.0, hi: float = 1.0, *, per_row: bool = False) -> torch.Tensor:
    """``normalize.py:33-38``: rescale into ``[lo, hi]``; a constant signal maps to ``(lo + hi) / 2``."""
    return _run(x, _lib.RN_MINMAX, 0, lo, hi, 0 if per_row else _lib.RN_GLOBAL)


def minmax_normalise_torch(x: torch.Tensor, lo: float = -1.0, hi: float = 1.0, *, per_row: bool = False) -> torch.Tensor:
    """``normalize.py:41-44``: the tensor rule, ``span + 1e-8`` in the denominator."""
    return _run(x, _lib.RN_MINMAX, 0, lo, hi, _lib.RN_EPS | (0 if per_row else _lib.RN_GLOBAL))


def z_normalise(x: torch.Tensor) -> torch.Tensor:
    """``normalize.py:47-49`` for signals along the last dimension: ``(x - mean) / (population std + 1e-8)``."""
    return _run(x, _lib.RN_ZSCORE, 0, 0.0, 0.0, 0)


def z_normalise_torch(x: torch.Tensor) -> torch.Tensor:
    """``normalize.py:52-56``: per-channel z-score over the time dimension of ``[B, C, T]`` (any leading dims)."""
    return _run(x, _lib.RN_ZSCORE, 0, 0.0, 0.0, 0)


def kpeak_normalise(x: torch.Tensor, k: int = 3, lo: float = -1.0, hi: float = 1.0, *, per_row: bool = False) -> torch.Tensor:
    """``normalize.py:59-72``: the mean of the ``k`` smallest / largest samples as the range (``k`` beyond the length
    uses every sample, as the NumPy slices do); a non-positive span maps to ``(lo + hi) / 2``.  A batched input without
    ``per_row`` is sorted as ONE signal by the reference (``np.sort`` of a 2-D array sorts rows, then ``[:k]`` takes
    rows) -- that form is not reproduced: pass ``per_row=True`` for batches."""
    if int(k) < 1:
        raise ValueError("k must be at least 1")
    if x.dim() > 1 and not per_row:
        raise ValueError("kpeak_normalise takes one signal; pass per_row=True for a batch")
    return _run(x, _lib.RN_KPEAK, k, lo, hi, 0)


def kpeak_normalise_torch(x: torch.Tensor, k: int = 26, lo: float = -1.0, hi: float = 1.0, dim: int = -1, *,
                          per_row: bool = False) -> torch.Tensor:
    """``normalize.py:75-78``: ``topk`` along the last dimension, then ONE mean over all selected values (every row's
    ``k`` largest), ``span + 1e-8`` in the denominator.  ``k`` larger than the length raises like ``torch.topk``."""
    if dim not in (-1, x.dim() - 1):
        raise ValueError("kpeak_normalise_torch: only the last dimension is supported")
    if int(k) < 1 or int(k) > int(x.shape[-1]):
        raise RuntimeError("selected index k out of range")
    return _run(x, _lib.RN_KPEAK, k, lo, hi, _lib.RN_EPS | (0 if per_row else _lib.RN_GLOBAL))


def row_statistics(x: torch.Tensor, k: int = 0) -> torch.Tensor:
    """Per-row ``(min, max, mean, population std, hi_ref, lo_ref)`` as float64 ``[rows, 6]`` (``hi_ref`` / ``lo_ref``:
    mean of the ``k`` largest / smallest samples when ``k > 0``, else max / min; std only with ``k == 0``)."""
    mode = _lib.RN_KPEAK if k > 0 else _lib.RN_ZSCORE
    _, stats = _run(x, mode, max(int(k), 1), 0.0, 0.0, 0, want_stats=True)
    return stats[:-1, :6]
