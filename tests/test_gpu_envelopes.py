"""Hilbert and homomorphic envelopes (signalproc/envelopes.py:11-23) on the device against the
golden vectors the reference produced and the SciPy oracle on fresh inputs.  Tolerance: 2e-6 of the row's peak for the
Hilbert envelope (fp64 transforms, one float32 rounding), 2e-5 relative for the homomorphic envelope (its logarithm
passes through float32)."""
import numpy as np
import pytest
import torch

from oracle import numpy_path as onp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ev(built_lib):
    from wav2vec_heart_sounds_b200 import envelopes
    return envelopes


def peak_err(got, want):
    got = got.cpu().numpy().astype(np.float64)
    assert got.shape == want.shape
    return np.abs(got - want).max() / np.abs(want).max()


def test_envelopes_vs_golden(ev, golden):
    g = golden("normalisers.npz")
    x, fs = torch.from_numpy(g["env_x"]).cuda(), float(g["env_fs"])
    assert peak_err(ev.hilbert_envelope(x), g["hilbert"]) < 2e-6
    assert peak_err(ev.hilbert_envelope(x[:, :2499]), g["hilbert_odd"]) < 2e-6          # odd length: no Nyquist bin
    got = ev.homomorphic_envelope(x, fs).cpu().numpy().astype(np.float64)
    assert np.abs(got / g["homomorphic"] - 1).max() < 2e-5


@pytest.mark.parametrize("t", [1, 2, 3, 129, 1000, 4096, 16500, 65537, 123750])
def test_hilbert_envelope_lengths(ev, t):
    """Lengths on both sides of every plan boundary (16 x 16 up to 512 x 512 transforms), powers of two, primes'
    neighbours and the configs[1] row length 123750 = 2 * 3^2 * 5^4 * 11."""
    rng = np.random.default_rng(t)
    n = np.arange(t)
    x = (np.sin(n / 5.0)[None] * (1 + 0.5 * np.sin(n / 200.0))[None] + 0.1 * rng.standard_normal((3, t)) + 0.05).astype(np.float32)
    want = np.stack([onp.hilbert_envelope(r) for r in x])
    assert peak_err(ev.hilbert_envelope(torch.from_numpy(x).cuda()), want) < 2e-6


def test_hilbert_envelope_properties_full_size(ev):
    """A full-scale batch (512 rows x 123750): a pure tone with a whole number of cycles has a flat envelope equal to its
    amplitude, the envelope dominates |x| everywhere, and leading dimensions are kept."""
    t, rows = 123750, 512
    n = torch.arange(t, device="cuda", dtype=torch.float64)
    amp = torch.linspace(0.1, 3.0, rows, device="cuda", dtype=torch.float64)[:, None]
    x = (amp * torch.cos(2 * torch.pi * 625.0 * n / t + 0.3)).float()
    env = ev.hilbert_envelope(x.view(256, 2, t))
    assert env.shape == (256, 2, t)
    env = env.view(rows, t)
    assert (env.double() / amp - 1).abs().max() < 1e-5
    gen = torch.Generator(device="cuda").manual_seed(3)
    y = torch.randn(64, t, device="cuda", generator=gen)
    e = ev.hilbert_envelope(y)
    assert torch.all(e >= y.abs() * (1 - 1e-6))


def test_homomorphic_envelope_vs_oracle(ev):
    fs, t = 2000.0, 20000
    rng = np.random.default_rng(4)
    n = np.arange(t)
    x = (np.sin(2 * np.pi * 60 * n / fs) * np.exp(-((n % 1600) - 300.0) ** 2 / 2e4) + 0.02 * rng.standard_normal((2, t))).astype(np.float32)
    want = np.stack([onp.homomorphic_envelope(r, fs) for r in x])
    got = ev.homomorphic_envelope(torch.from_numpy(x).cuda(), fs).cpu().numpy().astype(np.float64)
    assert np.abs(got / want - 1).max() < 2e-5
    with pytest.raises(ValueError):
        ev.homomorphic_envelope(torch.zeros(4, 100, device="cuda"), 10.0)                 # cutoff above Nyquist
    with pytest.raises(ValueError):
        ev.hilbert_envelope(torch.zeros(1, (1 << 19) + 1, device="cuda"))
