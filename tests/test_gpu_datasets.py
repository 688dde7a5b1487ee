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


@pytest.mark.parametrize("t", [24576, 30000, 9000, 200])
def test_generator_conditioning_batch(ds, t):
    """normalise -> fade -> fit_length -> log_mel / add_chirp of GenerativeDataset.__getitem__
    (datasets/generative.py:88-115), batched, against the NumPy restatement and torchaudio's float64 mel."""
    import wav2vec_heart_sounds_b200 as pkg
    from oracle import torch_path as otp
    rng = np.random.default_rng(t)
    fs, crop_frames, hop = 4000, 96, 256
    crop = crop_frames * hop
    ref = (np.sin(np.arange(t) / 9.0)[None] * rng.uniform(0.2, 3, (4, 1)) + 0.1 * rng.standard_normal((4, t)) + 0.3).astype(np.float32)
    con = (np.sin(np.arange(t) / 23.0)[None] + 0.2 * rng.standard_normal((4, t))).astype(np.float32)
    tr = pkg.MelConfig(sample_rate=fs, n_fft=1024, hop_length=hop, n_mels=80, f_max=500).build()
    got = ds.condition_generator_batch(torch.from_numpy(ref).cuda(), torch.from_numpy(con).cuda(), fs, tr, crop_frames, hop)
    otr = otp.mel_transform(fs, 1024, hop, n_mels=80, f_max=500).double()
    for r in range(4):
        wref, wcon, wchirp = onp.generator_item(ref[r], con[r], fs, crop)
        assert got["ref_audio"].shape == (4, crop)
        assert np.abs(got["ref_audio"][r].cpu().numpy() - wref).max() < TOL
        assert np.abs(got["con_audio"][r].cpu().numpy() - wcon).max() < TOL
        assert np.abs(got["chirp_wave"][r].cpu().numpy() - wchirp).max() < 2e-5 * max(1.0, np.abs(wchirp).max())
        wspec = otp.log_mel(torch.from_numpy(wcon).double(), otr)
        wspec = wspec[..., :crop_frames] if wspec.shape[-1] >= crop_frames else torch.nn.functional.pad(wspec, (0, crop_frames - wspec.shape[-1]))
        assert got["con_spec"].shape == (4, 80, crop_frames)
        assert float((got["con_spec"][r].cpu().double() - wspec).abs().max()) < TOL


def test_generator_conditioning_vs_golden(ds, golden):
    """The same kernel against the vectors the reference itself produced (tests/golden/gen_condition.npz)."""
    import wav2vec_heart_sounds_b200 as pkg
    g = golden("gen_condition.npz")
    fs, crop = int(g["fs"]), int(g["crop"])
    tr = pkg.MelConfig(sample_rate=fs, n_fft=1024, hop_length=256, n_mels=80, f_max=500).build()
    for tag in ("long", "exact", "short", "tiny"):
        x = torch.from_numpy(g[f"{tag}_x"]).cuda()
        got = ds.condition_generator_batch(x, x, fs, tr, crop // 256, 256)
        assert np.abs(got["ref_audio"].cpu().numpy() - g[f"{tag}_y"]).max() < TOL
        assert np.abs(got["chirp_wave"].cpu().numpy() - g[f"{tag}_chirp"]).max() < 2e-5


def test_fragment_dataset_augmented_copies(ds):
    """FragmentTensorDataset with augmented copies (reference datasets/fragments.py:47-83): originals come back
    untouched, augmented items are re-normalised rows of the fused chain (PCG column only for multichannel fragments),
    a whole batch of indices costs one launch, and device_augment_fn plugs into the reference's per-item hook."""
    import wav2vec_heart_sounds_b200 as pkg
    recs = _records(True)[:3]
    fb = ds.build_fragments_batched(recs, fs_out=4125, window=pkg.WindowSpec(2.0), ecg=True)
    data = ds.FragmentTensorDataset(fb, augment_num=2)
    assert len(data) > len(fb) and data.labels.count(0) > 0
    torch.manual_seed(1); np.random.seed(1)
    idx = list(range(len(data)))
    got = data.get_batch(idx)
    assert got["waveform"].shape == (len(data), 8250, 2) and len(got["patient"]) == len(data)
    aug = data._aug
    src = fb.windows[data._frag.to(fb.windows.device)]
    assert torch.equal(got["waveform"][~aug.cuda()], src[~aug.cuda()])                    # originals
    a = got["waveform"][aug.cuda()]
    assert torch.equal(a[:, :, 1], src[aug.cuda()][:, :, 1])                               # ECG column passes through
    assert torch.isfinite(a).all() and float(a[:, :, 0].abs().max()) <= 1.0
    assert float((a[:, :, 0].amax(dim=1) - a[:, :, 0].amin(dim=1)).min()) > 0.5            # re-normalised rows
    assert not torch.equal(a[:, :, 0], src[aug.cuda()][:, :, 0])
    item = data[int(torch.nonzero(aug)[0])]
    assert item["waveform"].shape == (8250, 2) and isinstance(item["label"], int)
    mono = ds.FragmentTensorDataset(fb, channel=0, augment_num=1, balance=False)
    assert mono.get_batch([0, 1, 2])["waveform"].shape == (3, 8250)
    fn = ds.device_augment_fn()
    w = fb.windows[0].cpu().numpy()
    out = fn(w, 4125)
    assert out.shape == w.shape and out.dtype == np.float32 and np.array_equal(out[:, 1], w[:, 1])
    assert fn(w[:, 0], 4125).shape == (8250,)
