"""min-max / z-score / k-peak normalisers (signalproc/normalize.py:33-78) on the device against
the golden vectors the reference produced and against the oracle on fresh inputs.  Tolerance: 2e-6 absolute on outputs
of order one (a float32 rounding of the float64 map); the selected k-th values themselves are checked exactly."""
import numpy as np
import pytest
import torch

from oracle import numpy_path as onp
from oracle import torch_path as otp

pytestmark = pytest.mark.gpu
TOL = 2e-6


@pytest.fixture(scope="module")
def nz(built_lib):
    from wav2vec_heart_sounds_b200 import normalize
    return normalize


def close(got, want, tol=TOL):
    got = got.cpu().numpy().astype(np.float64)
    assert got.shape == want.shape
    return np.abs(got - want).max() <= tol * max(1.0, np.abs(want).max())


def test_normalisers_vs_golden(nz, golden):
    g = golden("normalisers.npz")
    x = torch.from_numpy(g["x"]).cuda()
    assert close(nz.minmax_normalise(x, per_row=True), g["minmax"])
    assert close(nz.minmax_normalise(x, 0.0, 2.0, per_row=True), g["minmax_02"])
    assert close(nz.minmax_normalise(x[0]), g["minmax"][0])                       # one signal: both scopes coincide
    assert close(nz.z_normalise(x), g["z"])
    assert close(nz.z_normalise_torch(x[None])[0], g["z_torch"])
    assert close(nz.kpeak_normalise(x, per_row=True), g["kpeak3"])
    assert close(nz.kpeak_normalise(x, k=40, lo=0.0, hi=1.0, per_row=True), g["kpeak40"])
    assert close(nz.kpeak_normalise(x[2]), g["kpeak3"][2])
    assert close(nz.minmax_normalise_torch(x, per_row=True), g["minmax_torch_rows"])
    assert close(nz.minmax_normalise_torch(x), g["minmax_torch_all"])             # the reference's whole-tensor range
    assert close(nz.kpeak_normalise_torch(x, per_row=True), g["kpeak_torch_rows"])
    assert close(nz.kpeak_normalise_torch(x), g["kpeak_torch_all"])
    assert close(nz.kpeak_normalise_torch(x, k=5, lo=0.0, hi=3.0), g["kpeak_torch_all_k5"])
    flat = torch.from_numpy(g["flat"]).cuda()
    assert torch.equal(nz.minmax_normalise(flat[0]).cpu(), torch.from_numpy(g["minmax_flat"][0]).float())
    assert torch.equal(nz.kpeak_normalise(flat[0]).cpu(), torch.from_numpy(g["kpeak_flat"][0]).float())


@pytest.mark.parametrize("t,k", [(123750, 26), (16500, 3), (517, 40), (33, 33), (1, 1)])
def test_kpeak_selection_exact(nz, t, k):
    """The reference values of the k-peak range (mean of the k largest / smallest) from the exact radix select, on
    rows with heavy ties (values quantised to 1/64), negative-only rows, and k equal to the row length."""
    rng = np.random.default_rng(t + k)
    x = rng.standard_normal((5, t)).astype(np.float32)
    x[1] = np.round(x[1] * 64) / 64                                               # many equal keys around the k-th value
    x[2] = -np.abs(x[2]) - 1.0
    x[3] = np.abs(x[3]) * 1e-30                                                   # tiny magnitudes, all in few key bins
    x[4, : t // 2] = 0.0
    st = nz.row_statistics(torch.from_numpy(x).cuda(), k=k).cpu().numpy()
    srt = np.sort(x.astype(np.float64), axis=1)
    np.testing.assert_allclose(st[:, 4], srt[:, -k:].mean(axis=1), rtol=1e-14, atol=0)
    np.testing.assert_allclose(st[:, 5], srt[:, :k].mean(axis=1), rtol=1e-14, atol=0)
    np.testing.assert_array_equal(st[:, 0], x.min(axis=1).astype(np.float64))
    np.testing.assert_array_equal(st[:, 1], x.max(axis=1).astype(np.float64))
    np.testing.assert_allclose(st[:, 2], x.astype(np.float64).mean(axis=1), rtol=0, atol=1e-15 * max(1, t) ** 0.5)


def test_normalisers_full_size_properties(nz):
    """configs[1] row shape (2048 rows x 123750 samples): exact range ends, zero mean / unit deviation, idempotence."""
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(2048, 123750, device="cuda", generator=gen) * 0.3 + 0.1
    y = nz.minmax_normalise(x, per_row=True)
    assert torch.all(y.amin(dim=1) == -1.0) and torch.all((y.amax(dim=1) - 1.0).abs() <= 1.2e-7)
    z = nz.z_normalise_torch(x.view(1024, 2, -1))
    assert z.shape == (1024, 2, 123750)
    assert z.mean(dim=-1).abs().max() < 1e-6 and (z.double().std(dim=-1, unbiased=False) - 1).abs().max() < 1e-6
    kp = nz.kpeak_normalise_torch(x, per_row=True)
    again = nz.kpeak_normalise_torch(kp, per_row=True)
    assert (kp - again).abs().max() < 5e-7                                       # already in [lo, hi] by its own range
    want = otp.kpeak_normalise_torch(x[:4].cpu().double())
    assert close(nz.kpeak_normalise_torch(x[:4]), want.numpy())
    w2 = np.stack([onp.kpeak_normalise(r) for r in x[:3].cpu().numpy()])
    assert close(nz.kpeak_normalise(x[:3], per_row=True), w2)


def test_normaliser_argument_errors(nz):
    x = torch.zeros(4, 100, device="cuda")
    with pytest.raises(RuntimeError):
        nz.kpeak_normalise_torch(x, k=101)
    with pytest.raises(ValueError):
        nz.kpeak_normalise(x)                                                     # a batch needs per_row=True
    with pytest.raises(ValueError):
        nz.kpeak_normalise_torch(x, dim=0)
    with pytest.raises(TypeError):
        nz.minmax_normalise(np.zeros(8, dtype=np.float32))
    assert nz.minmax_normalise(torch.zeros(0, 16, device="cuda")).shape == (0, 16)
    assert close(nz.kpeak_normalise(torch.arange(5, device="cuda", dtype=torch.float32), k=9), onp.kpeak_normalise(np.arange(5.0), k=9))
