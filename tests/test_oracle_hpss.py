"""The HPSS oracle has no reference outputs to pin it to (librosa absent, no reference test: parity unpinned).
These identities keep the restatement honest; they run on the CPU."""
import numpy as np

from oracle import hpss_path as oh


def test_stft_istft_round_trip_and_shapes():
    x = np.random.default_rng(0).standard_normal(4000)
    for n_fft, hop in ((512, 16), (1024, 128), (2048, 64)):
        s = oh.stft(x, n_fft, hop)
        assert s.shape == (n_fft // 2 + 1, 1 + len(x) // hop)
        y = oh.istft(s, n_fft, hop)
        assert len(y) == hop * (len(x) // hop)
        np.testing.assert_allclose(y, x[:len(y)], atol=1e-12)


def test_components_add_up_and_masks_bounded():
    x = np.random.default_rng(1).standard_normal(3000)
    s = oh.stft(x, 512, 32)
    h, p, r = oh.hpss_spectra(s, (1.5, 2.5), (9, 12))
    np.testing.assert_allclose(h + p + r, s, atol=1e-12)
    assert np.all(np.abs(h) <= np.abs(s) + 1e-12) and np.all(np.abs(p) <= np.abs(s) + 1e-12)
    hw, pw, rw = oh.hpss_split(x, 512, 32, (1.5, 2.5), (9, 12))
    np.testing.assert_allclose(hw + pw + rw, x[:len(hw)], atol=1e-10)


def test_median_rank_and_reflection_rules():
    """Window [i - k//2, i - k//2 + k - 1], half-sample-symmetric reflection, element k//2 of the sorted window
    (upper median for even k) -- SURVEY.md section 8a row H."""
    a = np.array([[5.0, 1.0, 4.0, 2.0, 3.0, 9.0]])
    k = 4
    want = []
    for i in range(6):
        idx = np.arange(i - k // 2, i - k // 2 + k)
        idx = np.where(idx < 0, -idx - 1, idx)
        idx = np.where(idx >= 6, 2 * 6 - 1 - idx, idx)
        want.append(np.sort(a[0, idx])[k // 2])
    np.testing.assert_array_equal(oh.median_time(a, k)[0], want)
    np.testing.assert_array_equal(oh.median_freq(a.T, k)[:, 0], want)


def test_softmask_limits():
    x = np.array([0.0, 1.0, 1.0, 3.0]); r = np.array([0.0, 0.0, 1.0, 1.0])
    np.testing.assert_allclose(oh.softmask(x, r), [0.0, 1.0, 0.5, 0.9])
    assert oh.softmask(x, r, split_zeros=True)[0] == 0.5


def test_recombine_is_normalised_and_deterministic():
    x = np.random.default_rng(2).standard_normal(2048)
    p = dict(n_fft1=512, hop1=64, n_fft2=512, hop2=32, margin1=(1.2, 1.7), margin2=(2.0, 3.0), kernel1=(7, 11),
             kernel2=(5, 30), w1=[1, 2, 3, 4, 5, 6, 7], w2=[7, 6, 5, 4, 3, 2, 1], w_mix=0.03)
    y, n = oh.hpss_recombine(x, p)
    assert n == 2048 and y.shape == (2048,) and np.abs(y).max() <= 1.0 and abs(np.abs(y).max() - 1.0) < 1e-12
    y4, n4 = oh.hpss_recombine(x, dict(p, w1=p["w1"][:4], w2=p["w2"][:4]), include_residual=False)
    assert n4 == 2048 and not np.allclose(y, y4)


def test_transforms_equal_the_torch_and_scipy_equivalents():
    """librosa's stft / istft (centre padding with zeros, periodic Hann of n_fft, istft normalised by the window's
    sum of squares and trimmed) are the same transforms as ``torch.stft`` / ``torch.istft`` with those arguments and
    as ``scipy.signal.ShortTimeFFT``-style framing; the restatement must agree with those independent implementations
    to float64 rounding.  (Not a substitute for reference outputs -- the module stays "parity unpinned" -- but it ties
    two of its three stages to library code that is installed here.)"""
    import torch
    from scipy import ndimage
    rng = np.random.default_rng(12)
    for n_fft, hop, t in ((512, 32, 3000), (1024, 64, 5003), (2048, 128, 9000), (256, 50, 1234)):
        y = rng.standard_normal(t)
        s = oh.stft(y, n_fft, hop)                                             # [bins, frames]
        w = torch.hann_window(n_fft, periodic=True, dtype=torch.float64)
        ts = torch.stft(torch.from_numpy(y), n_fft, hop_length=hop, win_length=n_fft, window=w, center=True,
                        pad_mode="constant", return_complex=True).numpy()
        assert s.shape == ts.shape == (n_fft // 2 + 1, 1 + t // hop)
        assert np.abs(s - ts).max() < 1e-10 * np.abs(ts).max()
        back = oh.istft(s, n_fft, hop)
        tb = torch.istft(torch.from_numpy(ts), n_fft, hop_length=hop, win_length=n_fft, window=w, center=True).numpy()
        assert back.shape == tb.shape == (hop * (t // hop),)
        assert np.abs(back - tb).max() < 1e-10
        mag = np.abs(s)
        for k in (5, 17, 30):
            np.testing.assert_array_equal(oh.median_time(mag, k), ndimage.median_filter(mag, size=(1, k), mode="reflect"))
            np.testing.assert_array_equal(oh.median_freq(mag, k), ndimage.median_filter(mag, size=(k, 1), mode="reflect"))
