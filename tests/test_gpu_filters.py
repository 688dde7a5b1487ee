"""zero-phase filters (signalproc/filters.py:44-90) on the device against the golden vectors
the reference produced and against the NumPy/SciPy oracle on fresh inputs.  Tolerance 1e-5 of the reference's scale."""
import numpy as np
import pytest
import torch

from oracle import numpy_path as onp
from helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def fl(built_lib):
    from wav2vec_heart_sounds_b200 import filters
    return filters


def test_zero_phase_vs_golden(fl, golden):
    g = golden("zerophase.npz")
    fs, x = float(g["fs"]), torch.from_numpy(g["x"]).cuda()
    assert rel_err(fl.butter_bandpass(x, fs, 25.0, 400.0).cpu().numpy(), g["bandpass"]) < TOL
    assert rel_err(fl.butter_lowpass(x, fs, 150.0).cpu().numpy(), g["lowpass"]) < TOL
    assert rel_err(fl.butter_highpass(x, fs, 20.0).cpu().numpy(), g["highpass"]) < TOL
    assert rel_err(fl.band_stop(x, fs, 45.0, 55.0).cpu().numpy(), g["band_stop"]) < TOL
    assert rel_err(fl.notch(x, fs, 50.0).cpu().numpy(), g["notch"]) < TOL
    assert rel_err(fl.notch_chain(x, fs, (50.0, 100.0, 150.0, 3000.0)).cpu().numpy(), g["notch_chain"]) < TOL


@pytest.mark.parametrize("t,fs", [(40013, 16000.0), (15616, 4125.0), (15617, 4125.0), (64, 2000.0)])
def test_zero_phase_vs_oracle_sizes(fl, t, fs):
    """Several tiles, a row that ends exactly on a tile boundary, one sample past it, a row barely longer than the padding."""
    rng = np.random.default_rng(t)
    x = (rng.standard_normal((3, t)) + 0.4 + np.sin(np.arange(t) / 11.0)[None]).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    want = np.stack([onp.butter_bandpass_zp(r, fs, 25.0, 400.0) for r in x])
    assert rel_err(fl.butter_bandpass(xd, fs, 25.0, 400.0).cpu().numpy(), want) < TOL
    want = np.stack([onp.notch_zp(r, fs, 50.0, 30.0) for r in x])
    assert rel_err(fl.notch(xd, fs, 50.0, 30.0).cpu().numpy(), want) < TOL
    want = np.stack([onp.butter_highpass_zp(r, fs, 15.0) for r in x])
    assert rel_err(fl.butter_highpass(xd[None], fs, 15.0).cpu().numpy()[0], want) < TOL      # leading dims are kept


def test_zero_phase_too_short_raises(fl):
    with pytest.raises(ValueError):
        fl.butter_bandpass(torch.zeros(2, 20, device="cuda"), 4125.0, 25.0, 400.0)           # padlen 27 for 4 sections
    with pytest.raises(ValueError):
        fl.notch(torch.zeros(9, device="cuda"), 4125.0, 50.0)


def test_fir_band_split(fl, golden):
    """decompose_bands / preprocess_four_bands (reference filters.py:85-101, preprocess.py:40-42): the golden vectors
    of the reference and the SciPy oracle on a fresh multi-row input."""
    g = golden("zerophase.npz")
    fs, x = float(g["fs"]), torch.from_numpy(g["x"]).cuda()
    got = fl.decompose_bands(x[0], fs).cpu().numpy()
    assert got.shape == g["bands"][0].shape and rel_err(got, g["bands"][0]) < TOL
    four = fl.preprocess_four_bands(x[0], fs)
    assert four.shape == (x.shape[1], 4) and torch.equal(four.t().cpu(), torch.from_numpy(got))
    rng = np.random.default_rng(5)
    y = (rng.standard_normal((3, 20011)) + 0.3).astype(np.float32)
    got = fl.decompose_bands(torch.from_numpy(y).cuda(), 2000.0).cpu().numpy()
    want = np.stack([onp.decompose_bands(r, 2000.0) for r in y])
    assert got.shape == (3, 4, 20011) and rel_err(got, want) < TOL
    with pytest.raises(ValueError):
        fl.decompose_bands(torch.zeros(183, device="cuda"), 2000.0)
