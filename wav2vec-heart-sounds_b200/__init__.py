"""B200-native signal conditioning for PCG/ECG: a drop-in for the tensor path of
``mpcg_wav2vec`` (``signalproc.torchproc``, ``augment.torchaug``, ``signalproc.spectrogram``).

Import as ``wav2vec_heart_sounds_b200`` (the on-disk directory carries the reference's hyphenated
name; the importable alias package next to it redirects here).
"""
from . import design, segment as _segment_mod  # noqa: F401
from .segment import WINDOWS, WindowSpec, default_window
from . import torchproc
from .pipeline import preprocess_segment

__all__ = ["torchproc", "preprocess_segment", "WindowSpec", "WINDOWS", "default_window", "design"]
