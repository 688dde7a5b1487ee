"""HPSS kernels against the oracle restatement (oracle/hpss_path.py; parity unpinned, see its header).
Median selections are bit-exact on identical magnitudes; transforms / masks / waveforms to 1e-5 of the scale."""
import numpy as np
import pytest
import torch

from oracle import hpss_path as oh
from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hp(built_lib):
    from wav2vec_heart_sounds_b200 import hpss
    return hpss


@pytest.mark.parametrize("n_fft,hop,t", [(512, 16, 4000), (1024, 128, 8192), (2048, 64, 5001), (1024, 50, 5003), (1024, 64, 700)])
def test_stft_vs_oracle(hp, n_fft, hop, t):
    x = np.random.default_rng(0).standard_normal((2, t)).astype(np.float32)
    got = hp.stft(torch.from_numpy(x).cuda(), n_fft, hop).cpu().numpy()            # [B, frames, bins]
    for r in range(2):
        want = oh.stft(x[r], n_fft, hop).T
        assert got[r].shape == want.shape
        assert np.abs(got[r] - want).max() < 1e-5 * np.abs(want).max()


@pytest.mark.parametrize("k", [5, 6, 17, 30, 31])
def test_medians_bit_exact(hp, k):
    """Integer-valued magnitudes (exact in float32, squares exact too) so the kernel's |S| equals the oracle's input."""
    rng = np.random.default_rng(k)
    mag = rng.integers(0, 2000, size=(2, 37, 45)).astype(np.float32)               # [B, frames, bins]
    mag[0, :, 3] = 7.0                                                             # ties
    spec = torch.complex(torch.from_numpy(mag), torch.zeros(2, 37, 45)).cuda().contiguous()
    ht = hp.median_magnitude(spec, k, True).cpu().numpy()
    pf = hp.median_magnitude(spec, k, False).cpu().numpy()
    for r in range(2):
        m_ft = mag[r].T.astype(np.float64)                                         # librosa layout [bins, frames]
        np.testing.assert_array_equal(ht[r].T, oh.median_time(m_ft, k))
        np.testing.assert_array_equal(pf[r].T, oh.median_freq(m_ft, k))


@pytest.mark.parametrize("n_fft,hop,margin,kernel", [(512, 32, (1.5, 2.5), (9, 12)), (1024, 64, (1.0, 1.0), (30, 5)),
                                                     (2048, 128, (2.0, 4.0), (17, 17))])
def test_split_vs_oracle(hp, n_fft, hop, margin, kernel):
    rng = np.random.default_rng(4)
    t = np.arange(6000) / 4000.0
    x = (np.sin(2 * np.pi * 80 * t)[None] + 0.3 * rng.standard_normal((2, 6000))).astype(np.float32)
    x[:, 2000:2010] += 3.0
    h, p, r = hp.hpss_split(torch.from_numpy(x).cuda(), n_fft, hop, margin, kernel)
    for row in range(2):
        wh, wp, wr = oh.hpss_split(x[row], n_fft, hop, margin, kernel)
        scale = max(np.abs(wh).max(), np.abs(wp).max())
        assert h.shape[1] == len(wh)
        # 1e-5 of the scale (measured: 1e-7 .. 5e-7, tools/hpss_err.py; a selection that flips between two near-equal
        # magnitudes moves a mask by O(1e-6))
        assert np.abs(h[row].cpu().numpy() - wh).max() < 1e-5 * scale
        assert np.abs(p[row].cpu().numpy() - wp).max() < 1e-5 * scale
        assert np.abs(r[row].cpu().numpy() - wr).max() < 1e-5 * scale
    tot = (h + p + r).cpu().numpy()
    assert rel_err(tot, x[:, :tot.shape[1]]) < 1e-5                               # the three parts add up to the input


@pytest.mark.parametrize("residual", [True, False])
def test_recombine_vs_oracle(hp, residual):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 4096)).astype(np.float32)
    n = 7 if residual else 4
    p = dict(n_fft1=512, hop1=64, n_fft2=1024, hop2=32, margin1=(1.2, 1.7), margin2=(2.0, 3.0), kernel1=(7, 11),
             kernel2=(5, 30), w1=list(rng.uniform(0.01, 10, n)), w2=list(rng.uniform(0.01, 10, n)), w_mix=0.03)
    got, m = hp.hpss_recombine(torch.from_numpy(x).cuda(), residual, params=p)
    for row in range(2):
        want, mw = oh.hpss_recombine(x[row], p, residual)
        assert m == mw
        assert np.abs(got[row].cpu().numpy() - want).max() < 1e-5          # outputs are normalised to [-1, 1]; measured 2e-7 .. 4e-7
    assert float(got.abs().max()) <= 1.0
    out, m2 = hp.hpss_recombine(torch.from_numpy(x).cuda(), residual)            # random draws: shape and bounds only
    assert out.shape == (2, m2) and torch.isfinite(out).all() and float(out.abs().max()) <= 1.0
