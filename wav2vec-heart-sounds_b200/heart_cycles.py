"""Heart-cycle rearrangement for generator training, rebuilt on the device (reference
``datasets/heart_cycles.py``).

Host side (small integer work, kept in Python like the reference): reading the join indices of a segmentation file,
turning them into cycle bounds, and drawing the new cycle order from a ``random.Random`` in the reference's call order,
so the same seeded generator yields the same order.  Device side: :func:`rebuild_batch` joins the cycles of every row
with the reference's correlation-aware crossfade (``csrc/cycles.cu``, one CTA per row, joins replayed in order).
"""
from __future__ import annotations

import json
import random
from pathlib import Path
from typing import Sequence

import numpy as np
import torch

from . import _lib

__all__ = ["load_join_indices", "cycle_bounds", "rearrange_order", "CyclePlan", "rebuild_batch", "rebuild"]


def load_join_indices(seg_path, fs_out: float) -> list[int]:
    """Sorted cycle cut points in ``fs_out`` samples (``heart_cycles.py:22-29``): the first index of every non-empty
    group of the segmentation file, zeros dropped, rescaled with Python's ``round`` when the rates differ."""
    doc = json.loads(Path(seg_path).read_text())
    fs_seg = doc["fs"]
    cuts = sorted({int(group[0]) for group in doc["segments"] if len(group) and group[0] > 0})
    if fs_out != fs_seg:
        cuts = [round(c * fs_out / fs_seg) for c in cuts]
    return cuts


def cycle_bounds(length: int, joins: Sequence[int]) -> list[tuple[int, int]]:
    """``split_cycles`` (``heart_cycles.py:32-35``) as index pairs: consecutive joins strictly inside the signal."""
    inside = [int(j) for j in joins if 0 < j < length]
    return [(a, b) for a, b in zip(inside[:-1], inside[1:]) if b > a]


def rearrange_order(num: int, *, prob_contiguous: float = 0.0, random_start: bool = True,
                    rng: random.Random | None = None) -> list[int]:
    """The permutation ``rearrange`` applies to every signal's cycle list (``heart_cycles.py:72-98``), drawn with the
    same sequence of generator calls: one ``random()`` for the mode; rotation start, or five group sizes, the choice
    between single cycles and those groups, and the shuffle of the groups."""
    rng = rng or random.Random()
    if num < 2:
        return list(range(num))
    if rng.random() <= prob_contiguous:
        first = rng.randint(0, num - 1) if random_start else 0
        return [(first + i) % num for i in range(num)]
    sizes = rng.choice([[1], [rng.randint(1, 4) for _ in range(5)]])
    blocks, at, turn = [], 0, 0
    while at < num:
        width = sizes[turn % len(sizes)]
        blocks.append(list(range(at, min(at + width, num))))
        at += width
        turn += 1
    rng.shuffle(blocks)
    return [i for block in blocks for i in block]


class CyclePlan:
    """Device tables of one batch's cycle lists (``starts`` / ``lens`` ``[B, kmax]`` int32, ``counts`` ``[B]``), built
    once and shared by every signal that is cut at the same joins (reference and conditioning waveforms,
    ``datasets/generative.py:65-66``)."""

    def __init__(self, cycles: Sequence[Sequence[tuple[int, int]] | None], t: int, device):
        b = len(cycles)
        self.kmax = max([len(c) for c in cycles if c] + [1])
        starts = np.zeros((b, self.kmax), dtype=np.int32)
        lens = np.zeros((b, self.kmax), dtype=np.int32)
        counts = np.zeros(b, dtype=np.int32)
        for r, row in enumerate(cycles):
            if not row:
                continue
            for c, (lo, hi) in enumerate(row):                      # plain loops: a dozen pairs per row, NumPy calls cost more
                if not (0 <= lo <= hi <= t):
                    raise ValueError(f"row {r}: cycle ({lo}, {hi}) is outside the signal of {t} samples")
                starts[r, c], lens[r, c] = lo, hi - lo
            counts[r] = len(row)
        self.rows, self.t = b, int(t)
        self.longest = int(lens.max()) if b else 0
        self.passthrough = bool((counts == 0).any())
        self.starts, self.lens, self.counts = (torch.from_numpy(a).to(device) for a in (starts, lens, counts))


def rebuild_batch(x: torch.Tensor, cycles, target_len: int, fade_samples: int) -> tuple[torch.Tensor, torch.Tensor]:
    """``rebuild`` (``heart_cycles.py:55-69``) for every row of ``x [B, T]``: ``cycles[r]`` lists row r's cycles as
    ``(start, end)`` index pairs in joining order (``None`` or empty: the row passes through unchanged); a prepared
    :class:`CyclePlan` is accepted in its place.  Returns ``(y [B, cap], length [B] int64)``; row r is valid up to
    ``length[r]`` (at least ``target_len`` unless the reference's loop guard stops first), the rest of the row is
    unspecified."""
    x = _lib.require_cuda_f32(x)
    if x.dim() != 2:
        raise ValueError("x must be [B, T]")
    b, t = int(x.shape[0]), int(x.shape[1])
    plan = cycles if isinstance(cycles, CyclePlan) else None
    if plan is None:
        if len(cycles) != b:
            raise ValueError("x must be [B, T] with one cycle list per row")
        plan = CyclePlan(cycles, t, x.device)
    if plan.rows != b or plan.t != t:
        raise ValueError("the cycle plan was built for another batch shape")
    cap = max(int(target_len) + plan.longest, t if plan.passthrough else 0, 1)
    y = torch.empty((b, cap), device=x.device, dtype=torch.float32)
    length = torch.empty(b, device=x.device, dtype=torch.int64)
    xc = x.contiguous()
    _lib.check(_lib.lib().mpcg_cycle_rebuild_f32(xc.data_ptr(), y.data_ptr(), length.data_ptr(), plan.starts.data_ptr(),
                                                 plan.lens.data_ptr(), plan.counts.data_ptr(), b, t, cap, plan.kmax,
                                                 int(target_len), int(fade_samples), _lib.stream_ptr(x)), "cycle rebuild")
    return y, length


def rebuild(x: torch.Tensor, cycles: Sequence[tuple[int, int]], target_len: int, fade_samples: int) -> torch.Tensor:
    """One signal ``[T]``: the rebuilt waveform cut to its own length (``zeros(target_len)`` without cycles, as the
    reference returns)."""
    if not cycles:
        return torch.zeros(int(target_len), device=x.device, dtype=torch.float32)
    y, n = rebuild_batch(x[None], [cycles], target_len, fade_samples)
    return y[0, : int(n[0])]
