"""B200-native signal conditioning for PCG/ECG: a drop-in for the tensor path of
``mpcg_wav2vec`` (``signalproc.torchproc``, ``augment.torchaug``, ``signalproc.spectrogram``).

Import as ``wav2vec_heart_sounds_b200`` (the on-disk directory carries the reference's hyphenated
name; the importable alias package next to it redirects here).
"""
from . import design, segment as _segment_mod  # noqa: F401
from .segment import WINDOWS, WindowSpec, default_window
from . import torchproc, torchaug
from .torchaug import AugmentConfig, augment_pcg_batch
from .pipeline import preprocess_segment
from .spectrogram import MelConfig, log_mel
from . import beamformer, datasets, envelopes, filters, heart_cycles, normalize, pipelines
from .datasets import (build_fragments_batched, FragmentBatch, FragmentTensorDataset, device_batch_transform, device_augment_fn,
                       condition_generator_batch)

__all__ = ["torchproc", "MelConfig", "log_mel", "torchaug", "AugmentConfig", "augment_pcg_batch", "preprocess_segment", "WindowSpec", "WINDOWS", "default_window", "design",
           "datasets", "pipelines", "beamformer", "filters", "envelopes", "normalize", "heart_cycles", "build_fragments_batched", "FragmentBatch", "FragmentTensorDataset", "device_batch_transform", "device_augment_fn", "condition_generator_batch"]
