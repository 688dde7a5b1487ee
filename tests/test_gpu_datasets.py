"""the batched dataset builder against the reference's per-record loop
(datasets/cinc.py:54-125, datasets/vest.py:54-113) restated with the NumPy oracle: same windows (1e-5 of the
reference's scale), same order, same label / patient bookkeeping."""
import numpy as np
import pytest
import torch

from oracle import numpy_path as onp
from helpers import rel_err
from test_gpu_preprocess import _spiky

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def ds(built_lib):
    from wav2vec_heart_sounds_b200 import datasets
    return datasets


def _records(ecg, chans=2):
    rng = np.random.default_rng(3)
    lens = [20000, 9000, 20000, 3000, 14001, 9000, 500]                 # repeated lengths share a launch; 500 < start pad
    recs = []
    for i, n in enumerate(lens):
        cols = [_spiky(1, n, seed=50 + i)[0]]
        for c in range(1, chans):
            cols.append((np.sin(np.arange(n) / (200.0 + 30 * c)) + 0.05 * rng.standard_normal(n)).astype(np.float32))
        sig = np.stack(cols, axis=1) if (ecg or chans > 2) else cols[0]
        recs.append((sig, 2000, i % 2, f"p{i:03d}"))
    return recs


def _reference_loop(recs, fs_out, spec, kinds):
    wins, labels, pats = [], [], []
    for sig, fs, label, patient in recs:
        sig2 = sig[:, None] if sig.ndim == 1 else sig
        cols = [(onp.preprocess_pcg if k == "pcg" else onp.preprocess_ecg)(sig2[:, c].astype(np.float64), fs, fs_out)
                for c, k in enumerate(kinds)]
        base = cols[0] if len(kinds) == 1 else np.stack(cols, axis=1)
        for w in onp.segment(base, fs_out, spec):
            wins.append(w); labels.append(label); pats.append(patient)
    return np.stack(wins), labels, pats


@pytest.mark.parametrize("ecg", [False, True])
def test_builder_matches_reference_loop(ds, ecg):
    import wav2vec_heart_sounds_b200 as pkg
    recs = _records(ecg)
    spec = pkg.WindowSpec(1.0)
    fb = ds.build_fragments_batched(recs, fs_out=4125, window=spec, ecg=ecg)
    kinds = ("pcg", "ecg") if ecg else ("pcg",)
    want, labels, pats = _reference_loop(recs, 4125, onp.WindowSpec(1.0), kinds)
    got = fb.windows.cpu().numpy()
    assert got.shape == want.shape
    assert rel_err(got, want) < TOL
    assert fb.labels.tolist() == labels
    assert [fb.patients[i] for i in fb.record.tolist()] == pats
    assert "p006" not in pats                                             # shorter than the start pad: no fragment
    d = ds.FragmentTensorDataset(fb, channel=0 if ecg else -1)
    item = d[len(d) - 1]
    assert item["patient"] == pats[-1] and item["label"] == labels[-1] and item["waveform"].shape == (4125,)
    assert sum(fb.class_counts().values()) == len(fb)


def test_vest_six_channels_and_batch_transform(ds):
    import wav2vec_heart_sounds_b200 as pkg
    recs = _records(True, chans=6)[:3]
    fb = ds.build_fragments_batched(recs, fs_out=4125, window=pkg.WindowSpec(2.0), channels=("pcg",) * 6)
    want, labels, pats = _reference_loop(recs, 4125, onp.WindowSpec(2.0), ("pcg",) * 6)
    assert fb.windows.shape == want.shape and fb.windows.shape[1:] == (8250, 6)
    assert rel_err(fb.windows.cpu().numpy(), want) < TOL
    tf = ds.device_batch_transform(4125)
    torch.manual_seed(0); np.random.seed(0)
    y = tf(fb.windows[:5])
    assert y.shape == fb.windows[:5].shape and torch.isfinite(y).all()
    assert torch.equal(y[:, :, 1:], fb.windows[:5][:, :, 1:])             # only the PCG column is augmented
    assert float(y[:, :, 0].abs().max()) <= 1.0 and not torch.equal(y[:, :, 0], fb.windows[:5][:, :, 0])
