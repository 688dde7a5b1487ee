"""CPU checks of the host side: tap design for both resampler oracles, window arithmetic, the C-ABI
library's exported symbols.  No kernel is launched here."""
import ctypes
import pathlib
import re

import numpy as np
import pytest
import torch

from oracle import numpy_path as onp
from oracle import torch_path as otp
from helpers import dense_frames_apply, rel_err

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("fs_in,fs_out,n", [(2000, 16000, 1500), (2000, 4125, 3000), (4000, 4125, 3211), (4125, 2000, 2000),
                                            (1000, 1500, 777)])
def test_sinc_frames_equal_torchaudio(fs_in, fs_out, n):
    from wav2vec_heart_sounds_b200 import design
    up, down = design.reduce_ratio(fs_in, fs_out)
    g, off, depth = design.sinc_hann_frames(up, down)
    x = np.random.default_rng(0).standard_normal(n)
    want = otp.resample(torch.from_numpy(x), fs_in, fs_out).numpy()
    t_out = design.sinc_out_len(n, up, down)
    assert t_out == want.shape[-1]
    got = dense_frames_apply(x, g, up, down, off, t_out)
    assert rel_err(got, want) < 1e-13


@pytest.mark.parametrize("fs_in,fs_out,n", [(2000, 16000, 1500), (2000, 4125, 3000), (4000, 4125, 3211), (4125, 2000, 2000),
                                            (1000, 1500, 777), (2000, 4125, 17)])
def test_kaiser_frames_equal_scipy(fs_in, fs_out, n):
    from wav2vec_heart_sounds_b200 import design
    up, down = design.reduce_ratio(fs_in, fs_out)
    g, off, depth = design.kaiser_poly_frames(up, down, n)
    x = np.random.default_rng(1).standard_normal(n)
    want = onp.resample(x, fs_in, fs_out)
    t_out = design.kaiser_out_len(n, up, down)
    assert t_out == want.shape[-1]
    got = dense_frames_apply(x, g, up, down, off, t_out)
    assert rel_err(got, want) < 1e-13


def test_named_config_shapes():
    """The three ratios of BASELINE.json's configs land on the specialised kernel instances."""
    from wav2vec_heart_sounds_b200 import design
    assert design.sinc_hann_frames(8, 1)[1:] == (-7, 15)
    assert design.sinc_hann_frames(33, 16)[1:] == (-7, 30)
    assert design.sinc_hann_frames(33, 32)[1:] == (-7, 46)
    assert design.kaiser_poly_frames(8, 1, 60000)[1:] == (-10, 22)
    assert design.kaiser_poly_frames(33, 16, 60000)[1:] == (-10, 36)
    assert design.kaiser_poly_frames(33, 32, 32000)[1:] == (-10, 52)
    assert design.sinc_out_len(60000, 33, 16) == 123750 and design.sinc_out_len(60000, 8, 1) == 480000
    assert design.sinc_out_len(32000, 33, 32) == 33000


def test_butter_sections_are_the_reference_design():
    from wav2vec_heart_sounds_b200 import design
    b, a = otp.butter_ba(450.0, 4125.0, "lowpass", 2)
    s = design.butter_sos(450.0 / 4125.0, "lowpass", 2)
    np.testing.assert_array_equal(s[0], np.concatenate([b, a]))
    # SURVEY section 8a row C quotes these
    assert abs(s[0, 0] - 0.02349489526897043) < 1e-15
    s4 = design.butter_sos(0.05, "highpass", 4)
    assert s4.shape == (2, 6)


def test_window_spec_matches_oracle_and_survey():
    from wav2vec_heart_sounds_b200.segment import WindowSpec, start_index
    for fs in (1000, 2000, 4000, 4125, 16000, 22050):
        for ws in (1.0, 2.0, 4.0):
            a, b = WindowSpec(ws), onp.WindowSpec(ws)
            assert a.window_len(fs) == b.window_len(fs) and a.hop_len(fs) == b.hop_len(fs)
    s = WindowSpec(4.0)
    assert (s.window_len(4125), s.hop_len(4125), start_index(4125, s)) == (16500, 15469, 1238)
    assert (s.window_len(16000), s.hop_len(16000), start_index(16000, s)) == (64000, 60000, 4800)
    v = WindowSpec(2.0)
    assert (v.window_len(4125), v.hop_len(4125)) == (8250, 7219)


def test_window_count_matches_reference_table(golden, built_lib):
    """mpcg_window_count (host arithmetic in the library) against the reference's window_starts();
    where the NumPy path returns no window at all the tensor path returns one zero window."""
    lib = built_lib.lib()
    tab = golden("segment_index.npz")["table"]
    for fs, ws, n, win, hop, start, count, first, last in tab:
        got = lib.mpcg_window_count(int(n), int(start), int(win), int(hop))
        assert got == max(int(count), 1)
    assert lib.mpcg_window_count(123750, 1238, 16500, 15469) == 7
    assert lib.mpcg_window_count(480000, 4800, 64000, 60000) == 7
    assert lib.mpcg_window_count(33000, 1238, 8250, 7219) == 4


def test_library_exports_every_declared_symbol(built_lib):
    header = (ROOT / "include" / "mpcg_b200.h").read_text()
    names = set(re.findall(r"^(?:int|int64_t|void|const char\*)\s+(mpcg_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert names, "no prototypes found in the header"
    handle = ctypes.CDLL(str(built_lib.LIB_PATH))
    for n in sorted(names):
        assert hasattr(handle, n), f"{n} declared in include/mpcg_b200.h but not exported"
    assert names == set(built_lib.declared_symbols())
    assert handle.mpcg_abi_version() == 2


def test_argument_errors_are_reported_without_a_gpu(built_lib):
    lib = built_lib.lib()
    sos = np.zeros((7, 6))
    sos[:, 0] = sos[:, 3] = 1.0
    assert lib.mpcg_biquad_cascade_f32(0, 0, 1, 10, sos.ctypes.data, 7, None) == -2      # too many sections
    assert lib.mpcg_biquad_cascade_f32(0, 0, -1, 10, sos.ctypes.data, 1, None) == -1
    assert lib.mpcg_biquad_cascade_f32(0, 0, 0, 10, sos.ctypes.data, 1, None) == 0        # empty batch is fine
    assert lib.mpcg_segment_f32(0, 0, 1, 1, 100, 0, 10, 5, 3, 0, None) == -1             # wrong window count
    assert lib.mpcg_despike_f32(0, 1, 100, 10, 3.0, 5, 9, None, None, 0, None) == -1      # bad median mode
    assert b"invalid" in lib.mpcg_error_string(-1)


def test_cpu_tensors_are_refused():
    from wav2vec_heart_sounds_b200 import torchproc
    with pytest.raises(ValueError, match="no CPU fallback"):
        torchproc.abs_max_normalise(torch.zeros(2, 8))
    with pytest.raises(TypeError):
        torchproc.lowpass(np.zeros(8), 1000, 100)


def test_eq_band_design_closed_form_matches_scipy():
    """design.eq_band_sos writes SciPy's butter(1, band) out in closed form (it runs on the host once per augmentation
    batch); it must agree with scipy.signal.butter to the last place or two on every band the chain can draw."""
    import numpy as np
    from wav2vec_heart_sounds_b200 import design
    rng = np.random.default_rng(0)
    for fs in (4125.0, 16000.0, 2000.0):
        bands = []
        for _ in range(200):
            lo = float(rng.uniform(2, 0.95 * 500))
            hi = float(rng.uniform(lo + 0.05 * 498, 500))
            if hi < fs / 2:
                bands.append((lo, hi))
        bands += [(0.25, 100.0), (2.0, 26.9), (470.0, 499.9)] if fs > 1000 else []
        got, want = design.eq_band_sos(fs, bands), design.eq_band_sos_scipy(fs, bands)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 4.5e-16 * max(1.0, np.abs(want).max())


def test_fragment_dataset_item_list_matches_reference_rule():
    """FragmentTensorDataset lists every fragment once followed by its augmented copies, balanced like the reference's
    FragmentDataset (datasets/fragments.py:47-56); checked against the live reference when it is mounted."""
    import torch
    from wav2vec_heart_sounds_b200.datasets import FragmentBatch, FragmentTensorDataset
    labels = [0, 0, 0, 0, 0, 1, 1, 0, 1, 0]
    fb = FragmentBatch(torch.zeros(len(labels), 8), torch.tensor(labels), torch.arange(len(labels)), [f"p{i}" for i in range(len(labels))], 1000)
    ds = FragmentTensorDataset(fb, augment_num=3)
    # 7 of class 0, 3 of class 1: copies 3 and round(3 * 7 / 3) = 7
    assert len(ds) == 7 * (1 + 3) + 3 * (1 + 7)
    assert ds.labels[:4] == [0, 0, 0, 0] and ds._aug[:5].tolist() == [False, True, True, True, False]
    assert len(FragmentTensorDataset(fb, augment_num=3, balance=False)) == 10 * 4
    assert len(FragmentTensorDataset(fb)) == 10
    from conftest import reference_modules
    if reference_modules() is not None:
        from mpcg_wav2vec.datasets.fragments import Fragment, FragmentDataset
        ref = FragmentDataset([Fragment(np.zeros(8), lab, f"p{i}") for i, lab in enumerate(labels)], 1000, augment_num=3,
                              augment_fn=lambda w, fs: w)
        assert ref.labels == ds.labels and [a for _, a in ref._items] == ds._aug.tolist()


def test_mel_tables_of_the_float64_tensor_tier():
    """The host-side tables of mpcg_mel_dm_f32 (include/mpcg_b200.h): the basis in fragment order reproduces the plain
    windowed DFT basis entry for entry (column = 2 * bin + part, 16-sample slices, 4 pad entries), and the per-filter
    bin ranges cover exactly the non-zero weights of the HTK filterbank."""
    import warnings
    from wav2vec_heart_sounds_b200 import MelConfig
    for kw in (dict(sample_rate=16000, n_fft=1024, hop_length=256, n_mels=80, f_max=500.0),
               dict(sample_rate=4000, n_fft=1024, hop_length=256, n_mels=80, f_max=500.0),
               dict(sample_rate=4000, n_fft=2048, win_length=1200, hop_length=300, n_mels=128, f_max=500.0)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            tr = MelConfig(**kw).build()
        dm, rng = (t.numpy() for t in tr._dm_host)
        basis = tr._basis_host.numpy()                                   # [win, 2, kpad] float64
        win = basis.shape[0]
        assert dm.shape == (tr.kpad // 32, -(-win // 16), 64, 20) and dm.dtype == np.float64
        assert np.all(dm[..., 16:] == 0)
        r = np.random.default_rng(0)
        for _ in range(400):
            n, k, part = int(r.integers(win)), int(r.integers(tr.kpad)), int(r.integers(2))
            assert dm[k // 32, n // 16, 2 * (k % 32) + part, n % 16] == basis[n, part, k]
        pad_rows = dm.reshape(dm.shape[0], -1, 64, 20)[:, -1, :, :16]    # samples beyond the window are zero rows
        assert np.all(pad_rows[..., (win - 1) % 16 + 1:] == 0)
        fb = tr._fb_host.numpy()                                         # [nbins, n_mels]
        for m in range(tr.n_mels):
            nz = np.nonzero(fb[:, m])[0]
            lo, hi = int(rng[m, 0]), int(rng[m, 1])
            assert (lo, hi) == ((int(nz[0]), int(nz[-1]) + 1) if nz.size else (0, 0))
            assert np.all(fb[:lo, m] == 0) and np.all(fb[hi:, m] == 0)
