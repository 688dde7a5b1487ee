"""Window geometry (mirror of the reference's ``signalproc/segment.py:17-27`` dataclass).

Any object with ``window_len(fs)``, ``hop_len(fs)`` and ``start_pad_s`` is accepted by
:func:`torchproc.segment`, so the reference's own ``WindowSpec`` instances work unchanged.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class WindowSpec:
    window_s: float
    overlap_s: float = 0.25
    start_pad_s: float = 0.3

    def window_len(self, fs: float) -> int:
        # Python round(): half-to-even, e.g. round(0.3 * 4125) == 1238
        return int(round(fs * self.window_s))

    def hop_len(self, fs: float) -> int:
        step = int(round((self.window_s - self.overlap_s) * fs))
        return step if step > 1 else 1


# reference config.py:17-25
WINDOWS = {"cinc": WindowSpec(4.0), "training-a": WindowSpec(4.0), "vest": WindowSpec(2.0)}


def default_window(dataset: str) -> WindowSpec:
    return WINDOWS.get(dataset, WindowSpec(4.0))


def start_index(fs: float, spec) -> int:
    return int(round(spec.start_pad_s * fs))
