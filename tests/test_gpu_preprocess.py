"""GPU parity tests proper: every call goes through the C ABI (ctypes -> libmpcg_b200.so) and is compared
with the CPU oracle on the same seeded inputs, or with the golden vectors the reference produced.

Tolerance for floating-point outputs: max|got - want| <= 1e-5 * max|want| (BASELINE.json: "<= 1e-5 relative,
fp32"; `want` is the oracle run in float64).  Integer decisions (despike frames/peaks/spans, window counts and
starts) and pure copies are compared bit for bit."""
import numpy as np
import pytest
import torch

from oracle import numpy_path as onp
from oracle import torch_path as otp
from helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def tp(built_lib):
    from wav2vec_heart_sounds_b200 import torchproc
    return torchproc


def _dev(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float32).cuda()


def _spiky(rows, n, seed, spikes=3):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    x = (np.sin(2 * np.pi * t[None] / rng.uniform(20, 80, (rows, 1))) * rng.uniform(0.2, 2.0, (rows, 1))
         + 0.05 * rng.standard_normal((rows, n)))
    for r in range(rows):
        for _ in range(spikes):
            at = int(rng.integers(10, n - 20)); w = int(rng.integers(3, 11))
            x[r, at:at + w] += rng.choice([-1.0, 1.0]) * rng.uniform(5, 20)
    return x.astype(np.float32)


def _check_trace(kernel_rows, n_kernel, ref_rows):
    """Integer decisions, bit for bit.  The reference repeats a pass that can no longer change anything until
    max_iterations; the kernel stops after the first such pass, so any extra reference passes must all be
    copies of the kernel's last one."""
    assert 0 <= n_kernel <= len(ref_rows)
    assert [tuple(v) for v in kernel_rows[:n_kernel]] == ref_rows[:n_kernel]
    if len(ref_rows) > n_kernel:
        assert n_kernel > 0 and all(r == ref_rows[n_kernel - 1] for r in ref_rows[n_kernel:])


# ------------------------------------------------------------------ filters
@pytest.mark.parametrize("fs,band", [(4125.0, (25.0, 450.0)), (4125.0, (2.0, 40.0)), (16000.0, (25.0, 450.0)),
                                     (16000.0, (2.0, 40.0))])
def test_bandpass_cascade_vs_both_oracles(tp, fs, band):
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((5, 40013)) + 0.5).astype(np.float32)          # 3 tiles, ragged tail, DC offset
    got = tp.bandpass_cascade(_dev(x), fs, *band).cpu().numpy()
    want_t = otp.bandpass_cascade(torch.from_numpy(x).double(), fs, *band).numpy()
    want_n = np.stack([onp.bandpass_cascade(r, fs, *band) for r in x])
    assert rel_err(got, want_t) < TOL
    assert rel_err(got, want_n) < TOL


@pytest.mark.parametrize("order", [2, 4, 6])
def test_single_filters_and_orders(tp, order):
    rng = np.random.default_rng(4)
    x = rng.standard_normal((3, 20000)).astype(np.float32)
    got = tp.lowpass(_dev(x), 4125.0, 450.0, order=order).cpu().numpy()
    want = np.stack([onp.lowpass(r, 4125.0, 450.0, order) for r in x])
    assert rel_err(got, want) < TOL
    got = tp.highpass(_dev(x), 4125.0, 25.0, order=order).cpu().numpy()
    want = np.stack([onp.highpass(r, 4125.0, 25.0, order) for r in x])
    assert rel_err(got, want) < TOL


def test_filter_shapes_and_edges(tp):
    x1 = _dev(np.random.default_rng(5).standard_normal(1000))
    assert tp.lowpass(x1, 1000.0, 100.0).shape == (1000,)
    x3 = _dev(np.random.default_rng(5).standard_normal((2, 3, 500)))
    y3 = tp.highpass(x3, 1000.0, 20.0)
    assert y3.shape == (2, 3, 500)
    want = otp.highpass(x3.cpu().double(), 1000.0, 20.0).numpy()
    assert rel_err(y3.cpu().numpy(), want) < TOL
    assert tp.lowpass(torch.zeros(0, 16, device="cuda"), 1000.0, 100.0).shape == (0, 16)
    one = tp.lowpass(_dev([[1.0]]), 1000.0, 100.0)                           # single sample = b0
    assert abs(float(one) - otp.butter_ba(100.0, 1000.0, "lowpass", 2)[0][0]) < 1e-7


def test_filter_is_linear_at_full_size(tp):
    """Size-independent property at a BASELINE-sized row (123 750 samples): F(a x + b y) = a F(x) + b F(y)."""
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(8, 123750, device="cuda", generator=g)
    y = torch.randn(8, 123750, device="cuda", generator=g)
    f = lambda v: tp.bandpass_cascade(v, 4125.0, 25.0, 450.0)
    lhs = f(0.5 * x - 2.0 * y)
    rhs = 0.5 * f(x) - 2.0 * f(y)
    assert float((lhs - rhs).abs().max()) < 2e-5 * float(rhs.abs().max())


# ------------------------------------------------------------------ resample
@pytest.mark.parametrize("fs_in,fs_out,n", [(2000, 16000, 6000), (2000, 4125, 6011), (4000, 4125, 9000),
                                            (1000, 1500, 4001), (4125, 2000, 5000)])
def test_resample_both_modes(tp, fs_in, fs_out, n):
    rng = np.random.default_rng(6)
    x = rng.standard_normal((4, n)).astype(np.float32)
    got = tp.resample(_dev(x), fs_in, fs_out).cpu().numpy()
    want = otp.resample(torch.from_numpy(x).double(), fs_in, fs_out).numpy()
    assert got.shape == want.shape
    assert rel_err(got, want) < TOL
    got = tp.resample(_dev(x), fs_in, fs_out, mode="numpy").cpu().numpy()
    want = np.stack([onp.resample(r.astype(np.float64), fs_in, fs_out) for r in x])
    assert got.shape == want.shape
    assert rel_err(got, want) < TOL


def test_resample_identity_and_golden(tp, golden):
    x = _dev(np.ones((2, 10)))
    assert tp.resample(x, 4125, 4125) is x
    g = golden("resample_ratios.npz")
    assert rel_err(tp.resample(_dev(g["x_2k"]), 2000, 16000).cpu().numpy(), g["t64_2k_16k"]) < TOL
    assert rel_err(tp.resample(_dev(g["x_4k"]), 4000, 4125).cpu().numpy(), g["t64_4k_4125"]) < TOL
    assert rel_err(tp.resample(_dev(g["x_2k"]), 2000, 16000, mode="numpy").cpu().numpy(), g["np_2k_16k"]) < TOL
    assert rel_err(tp.resample(_dev(g["x_4k"]), 4000, 4125, mode="numpy").cpu().numpy(), g["np_4k_4125"]) < TOL


# ------------------------------------------------------------------ despike
@pytest.mark.parametrize("fs,n", [(4125.0, 12400), (1000.0, 6250), (16000.0, 40000)])
def test_despike_torch_mode_bit_exact(tp, fs, n):
    x = _spiky(4, n, seed=7)
    trace = []
    want = otp.remove_spikes(torch.from_numpy(x), fs, trace=trace).numpy()
    got, edits, tr = tp.remove_spikes(_dev(x), fs, return_trace=True)
    np.testing.assert_array_equal(got.cpu().numpy(), want)                   # values: bit for bit
    assert len(trace) > 0                                                    # the input really had spikes
    tr = tr.cpu().numpy(); edits = edits.cpu().numpy()
    for r in range(4):
        _check_trace(tr[r], int(edits[r]), [t[1:] for t in trace if t[0] == r])   # (frame, peak, lo, hi)


def test_despike_numpy_mode_bit_exact(tp):
    fs = 4125.0
    x = _spiky(4, 12400, seed=8)
    got, edits, tr = tp.remove_spikes(_dev(x), fs, mode="numpy", return_trace=True)
    got = got.cpu().numpy(); tr = tr.cpu().numpy(); edits = edits.cpu().numpy()
    for r in range(4):
        trace = []
        want = onp.remove_spikes(x[r], fs, trace=trace)
        np.testing.assert_array_equal(got[r], want.astype(np.float32))
        _check_trace(tr[r], int(edits[r]), trace)


def test_despike_edge_cases(tp):
    x = _dev(_spiky(2, 3000, seed=9))
    assert torch.equal(tp.remove_spikes(x, 8000.0), x)                       # shorter than one frame: untouched copy
    assert tp.remove_spikes(x, 8000.0).data_ptr() != x.data_ptr()
    z = torch.zeros(2, 5000, device="cuda")
    assert torch.equal(tp.remove_spikes(z, 1000.0), z)                       # all-zero: nothing exceeds 3*0
    assert torch.equal(tp.remove_spikes(z, 1000.0, mode="numpy"), z)
    one = tp.remove_spikes(x[0], 1000.0)
    assert one.shape == (3000,)
    clean = tp.remove_spikes(x, 1000.0)
    assert torch.equal(tp.remove_spikes(clean, 1000.0), clean)               # idempotent
    assert torch.equal(tp.remove_spikes(x, 1000.0, max_iterations=0), x)
    # tail beyond the last full frame is never touched
    y2 = _dev(_spiky(1, 2300, seed=10)); y2[:, 2200:] = 500.0                # 2 full frames of 1000 + tail of 300
    assert torch.equal(tp.remove_spikes(y2, 2000.0)[:, 2000:], y2[:, 2000:])


def test_despike_stuck_pass_stops_like_the_reference(tp):
    """A span that cannot shrink (flip right at the peak) makes the reference repeat one no-op pass until
    max_iterations; the kernel stops at the fixed point and must still return identical samples."""
    fs = 1000.0
    rng = np.random.default_rng(11)
    x = 0.01 * rng.standard_normal((1, 3000)).astype(np.float32)
    x[0, 700] = 9.0; x[0, 701] = -0.5                                        # flip immediately after the peak
    want = otp.remove_spikes(torch.from_numpy(x), fs, max_iterations=50).numpy()
    got = tp.remove_spikes(_dev(x), fs, max_iterations=50).cpu().numpy()
    np.testing.assert_array_equal(got, want)


# ------------------------------------------------------------------ normalise / segment
def test_abs_max_normalise(tp):
    rng = np.random.default_rng(12)
    x = (rng.standard_normal((6, 33001)) * 3 + 2).astype(np.float32)
    got = tp.abs_max_normalise(_dev(x)).cpu().numpy()
    want = otp.abs_max_normalise(torch.from_numpy(x).double()).numpy()
    assert rel_err(got, want) < TOL and np.abs(got).max() <= 1.0
    gotn = tp.abs_max_normalise(_dev(x), mode="numpy").cpu().numpy()
    wantn = np.stack([onp.abs_max_normalise(r) for r in x])
    assert rel_err(gotn, wantn) < TOL
    c = tp.abs_max_normalise(torch.full((2, 100), 3.0, device="cuda"))       # constant row -> zeros
    assert float(c.abs().max()) == 0.0
    bad = x.copy(); bad[0, 5] = np.nan
    gotb = tp.abs_max_normalise(_dev(bad)).cpu().numpy()
    wantb = otp.abs_max_normalise(torch.from_numpy(bad).double()).numpy()
    assert rel_err(gotb, wantb) < TOL
    v = tp.abs_max_normalise(_dev(x)[:, 1:])                                 # misaligned rows
    assert rel_err(v.cpu().numpy(), otp.abs_max_normalise(torch.from_numpy(x[:, 1:]).double()).numpy()) < TOL


@pytest.mark.parametrize("fs,ws,n", [(4125.0, 4.0, 123750), (4125.0, 2.0, 33000), (1000.0, 2.0, 10000), (4125.0, 4.0, 5000),
                                     (4125.0, 4.0, 1000), (1000.0, 2.0, 2300)])
def test_segment_bit_exact(tp, fs, ws, n):
    from wav2vec_heart_sounds_b200 import WindowSpec
    spec = WindowSpec(ws)
    x = np.random.default_rng(13).standard_normal((3, n)).astype(np.float32)
    got = tp.segment(_dev(x), fs, spec).cpu().numpy()
    want = otp.segment(torch.from_numpy(x), fs, onp.WindowSpec(ws)).contiguous().numpy()
    np.testing.assert_array_equal(got, want)
    # NumPy path agrees whenever it yields at least one window
    ref = onp.segment(x[0], fs, onp.WindowSpec(ws))
    if ref.shape[0]:
        np.testing.assert_array_equal(got[0], ref)
    assert tp.segment(_dev(x[0]), fs, spec).shape == got.shape[1:]


def test_segment_channel_layouts(tp):
    from wav2vec_heart_sounds_b200 import WindowSpec
    spec = WindowSpec(2.0)
    x = np.random.default_rng(14).standard_normal((4, 6, 9000)).astype(np.float32)
    planar = tp.segment(_dev(x), 1000.0, spec).cpu().numpy()
    want = otp.segment(torch.from_numpy(x), 1000.0, onp.WindowSpec(2.0)).contiguous().numpy()
    np.testing.assert_array_equal(planar, want)                              # [B, C, N, win]
    last = tp.segment(_dev(x), 1000.0, spec, channels_last=True).cpu().numpy()
    for b in range(4):
        np.testing.assert_array_equal(last[b], onp.segment(x[b].T, 1000.0, onp.WindowSpec(2.0)))   # [N, win, C]


# ------------------------------------------------------------------ whole chains
def test_chains_vs_golden(tp, golden):
    g = golden("preprocess_2k_4125.npz")
    fs_in, fs_out = float(g["fs_in"]), float(g["fs_out"])
    pcg, ecg = _dev(g["pcg"]), _dev(g["ecg"])
    assert rel_err(tp.preprocess_pcg(pcg, fs_in, fs_out).cpu().numpy(), g["t64_pcg"]) < TOL
    assert rel_err(tp.preprocess_ecg(ecg, fs_in, fs_out).cpu().numpy(), g["t64_ecg"]) < TOL
    assert rel_err(tp.preprocess_pcg(pcg, fs_in, fs_out, mode="numpy").cpu().numpy(), g["np_pcg"]) < TOL
    assert rel_err(tp.preprocess_ecg(ecg, fs_in, fs_out, mode="numpy").cpu().numpy(), g["np_ecg"]) < TOL
    from wav2vec_heart_sounds_b200 import WindowSpec
    w = tp.segment(tp.preprocess_pcg(pcg, fs_in, fs_out), fs_out, WindowSpec(1.0)).cpu().numpy()
    assert w.shape == g["t64_windows"].shape and rel_err(w, g["t64_windows"]) < TOL
    g2 = golden("resample_ratios.npz")
    assert rel_err(tp.preprocess_pcg(_dev(g2["x_2k"]), 2000, 16000).cpu().numpy(), g2["t64_pcg_2k_16k"]) < TOL
    assert rel_err(tp.preprocess_ecg(_dev(g2["x_2k"]), 2000, 16000).cpu().numpy(), g2["t64_ecg_16k"]) < TOL
    assert rel_err(tp.preprocess_pcg(_dev(g2["x_2k"]), 2000, 16000, mode="numpy").cpu().numpy(), g2["np_pcg_2k_16k"]) < TOL
    one = tp.preprocess_pcg(pcg[0], fs_in, fs_out)
    assert one.dim() == 1


def test_reference_property_tests_on_gpu(tp):
    """The reference's own property tests (tests/test_torchaug.py:18-23, tests/test_signalproc.py:44-58)."""
    from wav2vec_heart_sounds_b200 import WindowSpec
    x = torch.randn(4, 6000, device="cuda")
    out = tp.preprocess_pcg(x, 2000, 4125)
    assert out.shape[0] == 4 and torch.isfinite(out).all() and float(out.abs().max()) <= 1.0 + 1e-6
    w = tp.segment(out, 4125, WindowSpec(window_s=2.0))
    assert w.dim() == 3 and w.shape[0] == 4 and w.shape[2] == 8250
