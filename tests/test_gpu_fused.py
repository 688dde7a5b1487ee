"""Parity of the fused cluster kernel (mpcg_preprocess_segment_f32) against the float64 oracle, the golden
vectors and the stand-alone kernels.  Tolerance as in test_gpu_preprocess.py: 1e-5 of the reference's scale;
despike decisions and window geometry bit for bit."""
import numpy as np
import pytest
import torch

from oracle import numpy_path as onp
from oracle import torch_path as otp
from helpers import rel_err
from test_gpu_preprocess import _check_trace, _spiky

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def pkg(built_lib):
    import wav2vec_heart_sounds_b200 as m
    return m


def _dev(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float32).cuda()


def _oracle_torch(x, fs_in, fs_out, ws, kind="pcg", despike=True):
    xt = torch.from_numpy(x).double()
    trace = []
    v = otp.resample(xt, fs_in, fs_out)
    if kind == "pcg":
        if despike:
            v = otp.remove_spikes(v, fs_out, trace=trace)
        v = otp.abs_max_normalise(otp.bandpass_cascade(v, fs_out, *otp.PCG_BAND))
    else:
        v = otp.abs_max_normalise(otp.bandpass_cascade(v, fs_out, *otp.ECG_BAND))
    return otp.segment(v, fs_out, onp.WindowSpec(ws)).contiguous().numpy(), trace


@pytest.mark.parametrize("fs_in,fs_out,n,ws", [(2000, 4125, 60000, 4.0),     # config 2 row: 6-CTA clusters
                                               (4000, 4125, 32000, 2.0),     # config 5 row: 2-CTA clusters
                                               (2000, 4125, 9000, 1.0),      # short row: single CTA
                                               (2000, 16000, 8000, 1.0)])    # 8x ratio
def test_fused_pcg_vs_float64_oracle(pkg, fs_in, fs_out, n, ws):
    x = _spiky(3, n, seed=21)
    want, trace = _oracle_torch(x, fs_in, fs_out, ws)
    got, edits, tr = pkg.preprocess_segment(_dev(x), fs_in, fs_out, pkg.WindowSpec(ws), fused=True, return_trace=True)
    assert got.shape == want.shape
    assert rel_err(got.cpu().numpy(), want) < TOL
    assert len(trace) > 0
    edits, tr = edits.cpu().numpy(), tr.cpu().numpy()
    for r in range(3):
        _check_trace(tr[r], int(edits[r]), [t[1:] for t in trace if t[0] == r])


@pytest.mark.parametrize("mode", ["torch", "numpy"])
def test_long_rows_config0_full_length(pkg, mode):
    """configs[0] rows at full length: 30 s at 2 kHz -> 16 kHz = 480 000 samples per row (27 tiles streamed through one
    CTA, 8000-sample despike frames fetched into the frame cache), 4 s windows, against the float64 oracle of the mode;
    despike decisions bit for bit (signalproc/preprocess.py:24-37, segment.py:40-52, torchproc.py:101-129)."""
    x = _spiky(2, 60000, seed=61, spikes=4)
    x[1, 30000:30003] += 30.0
    got, edits, tr = pkg.preprocess_segment(_dev(x), 2000, 16000, pkg.WindowSpec(4.0), mode=mode, fused=True,
                                            return_trace=True)
    assert got.shape == (2, 7, 64000)
    got, edits, tr = got.cpu().numpy(), edits.cpu().numpy(), tr.cpu().numpy()
    if mode == "torch":
        want, trace = _oracle_torch(x, 2000, 16000, 4.0)
        assert rel_err(got, want) < TOL
        for r in range(2):
            _check_trace(tr[r], int(edits[r]), [t[1:] for t in trace if t[0] == r])
    else:
        spec = onp.WindowSpec(4.0)
        for r in range(2):
            trace = []
            v = onp.remove_spikes(onp.resample(x[r].astype(np.float64), 2000, 16000), 16000, trace=trace)
            want = onp.segment(onp.abs_max_normalise(onp.bandpass_cascade(v, 16000, *onp.PCG_BAND)), 16000, spec)
            assert rel_err(got[r], want) < TOL
            _check_trace(tr[r], int(edits[r]), trace)
    assert int(edits.sum()) >= 4


def test_long_ecg_rows_and_row_independence(pkg):
    """ECG rows of 480 000 samples (pole radius 0.9997 at 16 kHz carried across 27 tiles) and a batch larger than the
    persistent grid: every row must equal the same row processed alone."""
    rng = np.random.default_rng(62)
    t = np.arange(60000)
    ecg = (np.sin(t[None] / rng.uniform(200, 400, (3, 1))) + 0.3 + 0.05 * rng.standard_normal((3, 60000))).astype(np.float32)
    want, _ = _oracle_torch(ecg, 2000, 16000, 4.0, "ecg")
    got = pkg.preprocess_segment(_dev(ecg), 2000, 16000, pkg.WindowSpec(4.0), kinds=("ecg",), fused=True)
    assert rel_err(got.cpu().numpy(), want) < TOL
    big = _dev(_spiky(700, 9000, seed=63))                     # more rows than resident CTAs: tickets wrap around
    a = pkg.preprocess_segment(big, 2000, 4125, pkg.WindowSpec(1.0), fused=True)
    b = pkg.preprocess_segment(big[690:].contiguous(), 2000, 4125, pkg.WindowSpec(1.0), fused=True)
    assert torch.equal(a[690:], b)


def test_ragged_batch_equals_per_record_calls(pkg):
    """lengths=: recordings of different lengths in ONE launch (per-row tables in the descriptor).  Every recording's
    windows must equal, bit for bit, the call on that recording alone; counts follow mpcg_window_count."""
    rng = np.random.default_rng(71)
    lens = [60000, 9000, 31001, 2400, 17, 45678, 20000]
    pitch = (max(lens) + 3) & ~3
    x = np.zeros((len(lens), 2, pitch), np.float32)
    for b, n in enumerate(lens):
        x[b, 0, :n] = _spiky(1, n, seed=80 + b, spikes=2 if n > 100 else 0)[0] if n > 50 else 1.0
        x[b, 1, :n] = np.sin(np.arange(n) / 250.0) + 0.05 * rng.standard_normal(n)
        x[b, :, n:] = 7.0                                    # garbage past the valid part must not matter
    spec = pkg.WindowSpec(4.0)
    for layout in ("channels_last", "channel_major"):
        got, counts = pkg.preprocess_segment(_dev(x), 2000, 4125, spec, kinds=("pcg", "ecg"), lengths=lens, **{layout: True})
        at = 0
        for b, n in enumerate(lens):
            one = pkg.preprocess_segment(_dev(x[b:b + 1, :, :n]), 2000, 4125, spec, kinds=("pcg", "ecg"), fused=True,
                                         **{layout: True})
            k = int(counts[b])
            assert k == one.shape[1 if layout == "channels_last" else 2]
            mine = got[at:at + k] if layout == "channels_last" else got[:, at:at + k]
            want = one[0] if layout == "channels_last" else one[:, 0]
            assert torch.equal(mine, want), (layout, b)
            at += k
        assert at == got.shape[0 if layout == "channels_last" else 1]
    mono, counts = pkg.preprocess_segment(_dev(x[:, 0]), 2000, 16000, pkg.WindowSpec(1.0), lengths=lens, mode="numpy")
    at = 0
    for b, n in enumerate(lens):
        one = pkg.preprocess_segment(_dev(x[b:b + 1, 0, :n]), 2000, 16000, pkg.WindowSpec(1.0), mode="numpy", fused=True)
        k = int(counts[b])
        assert torch.equal(mono[at:at + k], one[0])
        at += k


def test_numpy_mode_bridges_nans_like_the_reference(pkg):
    """wfdb marks invalid samples with NaN; the NumPy chains interpolate them away first (signalproc/preprocess.py:25,34,
    normalize.py:11-17).  Leading / trailing / interior runs and runs across thread chunks, against the float64 oracle."""
    x = _spiky(4, 20000, seed=91, spikes=2)
    x[0, :37] = np.nan; x[0, 5000:5003] = np.nan; x[0, -11:] = np.nan
    x[1, 77] = np.nan; x[1, 9990:10400] = np.nan
    x[2, 1::2] = np.nan                                     # every other sample
    spec = onp.WindowSpec(1.0)
    got = pkg.preprocess_segment(_dev(x), 2000, 4125, pkg.WindowSpec(1.0), mode="numpy", fused=True).cpu().numpy()
    for r in range(4):
        want = onp.segment(onp.preprocess_pcg(x[r], 2000, 4125), 4125, spec)
        assert np.isfinite(got[r]).all()
        assert rel_err(got[r], want) < TOL, r
    from wav2vec_heart_sounds_b200 import torchproc
    filled = torchproc.fill_nans(_dev(x)).cpu().numpy()
    for r in range(4):
        assert rel_err(filled[r], onp.fill_nans(x[r])) < 1e-6
    e = torchproc.preprocess_ecg(_dev(x[:2]), 2000, 4125, mode="numpy").cpu().numpy()
    for r in range(2):
        assert rel_err(e[r], onp.preprocess_ecg(x[r], 2000, 4125)) < TOL
    clean = _dev(x[3:4])
    assert torchproc.fill_nans(clean) is clean               # nothing to repair: no copy


def test_despike_on_tiny_amplitudes_follows_the_reference_order(pkg):
    """Recordings in small physical units: frame maxima below the 1e-4 fill value, so a flattening pass RAISES its frame's
    maximum (the case the parallel rounds' invariant excludes -- they must hand such rows to the serial order, whose
    sorted-key update must cope with a rising key).  Decisions bit for bit against the float64 oracle, both paths."""
    x = (_spiky(3, 20000, seed=95, spikes=3) * 2e-6).astype(np.float32)
    want, trace = _oracle_torch(x, 2000, 4125, 4.0)
    got, edits, tr = pkg.preprocess_segment(_dev(x), 2000, 4125, pkg.WindowSpec(4.0), fused=True, return_trace=True)
    assert rel_err(got.cpu().numpy(), want) < TOL
    edits, tr = edits.cpu().numpy(), tr.cpu().numpy()
    for r in range(3):
        _check_trace(tr[r], int(edits[r]), [t[1:] for t in trace if t[0] == r])
    fast, e_fast = pkg.preprocess_segment(_dev(x), 2000, 4125, pkg.WindowSpec(4.0), fused=True, return_edits=True)
    assert torch.equal(fast, got) and np.array_equal(e_fast.cpu().numpy(), edits)


def test_fused_equals_chained_kernels(pkg):
    x = _dev(_spiky(5, 60000, seed=22))
    spec = pkg.WindowSpec(4.0)
    a = pkg.preprocess_segment(x, 2000, 4125, spec, fused=True)
    b = pkg.preprocess_segment(x, 2000, 4125, spec, fused=False)
    assert a.shape == b.shape == (5, 7, 16500)
    assert float((a - b).abs().max()) < 5e-6


def test_fused_pcg_ecg_pair_and_layouts(pkg):
    """Training-A shape: [B, 2, T] with a despiked PCG channel and an ECG channel, both output layouts."""
    rng = np.random.default_rng(23)
    pcg = _spiky(2, 20000, seed=24)
    ecg = (np.sin(np.arange(20000) / 300.0)[None] + 0.05 * rng.standard_normal((2, 20000))).astype(np.float32)
    x = np.stack([pcg, ecg], axis=1)
    spec = pkg.WindowSpec(4.0)
    want_p, _ = _oracle_torch(pcg, 2000, 4125, 4.0, "pcg")
    want_e, _ = _oracle_torch(ecg, 2000, 4125, 4.0, "ecg")
    planar = pkg.preprocess_segment(_dev(x), 2000, 4125, spec, kinds=("pcg", "ecg"), fused=True).cpu().numpy()
    assert planar.shape == (2, 2) + want_p.shape[1:]
    assert rel_err(planar[:, 0], want_p) < TOL and rel_err(planar[:, 1], want_e) < TOL
    last = pkg.preprocess_segment(_dev(x), 2000, 4125, spec, kinds=("pcg", "ecg"), channels_last=True,
                                  fused=True).cpu().numpy()
    assert last.shape == (2,) + want_p.shape[1:] + (2,)
    np.testing.assert_array_equal(last[..., 0], planar[:, 0])
    np.testing.assert_array_equal(last[..., 1], planar[:, 1])


def test_fused_numpy_mode(pkg):
    x = _spiky(2, 12000, seed=25)
    spec = onp.WindowSpec(1.0)
    got, edits, tr = pkg.preprocess_segment(_dev(x), 2000, 4125, pkg.WindowSpec(1.0), mode="numpy", fused=True,
                                            return_trace=True)
    got = got.cpu().numpy()
    for r in range(2):
        trace = []
        v = onp.remove_spikes(onp.resample(x[r].astype(np.float64), 2000, 4125), 4125, trace=trace)
        want = onp.segment(onp.abs_max_normalise(onp.bandpass_cascade(v, 4125, *onp.PCG_BAND)), 4125, spec)
        assert rel_err(got[r], want) < TOL
        assert rel_err(got[r], onp.segment(onp.preprocess_pcg(x[r], 2000, 4125), 4125, spec)) < TOL
        # the kernel despikes the float32-rounded resampler output; its decisions must match the oracle run on
        # the oracle's own float64 signal (ties between the two are vanishingly unlikely on this input)
        _check_trace(tr[r].cpu().numpy(), int(edits[r]), trace)


def test_fused_no_resampling_and_six_channels(pkg):
    """Vest shape: six PCG channels; also the fs_in == fs_out identity path."""
    x = np.stack([_spiky(2, 16000, seed=30 + c) for c in range(6)], axis=1)        # [2, 6, T]
    spec = pkg.WindowSpec(2.0)
    got = pkg.preprocess_segment(_dev(x), 4125, 4125, spec, channels_last=True, fused=True).cpu().numpy()
    for c in range(6):
        xt = torch.from_numpy(x[:, c]).double()
        v = otp.abs_max_normalise(otp.bandpass_cascade(otp.remove_spikes(xt, 4125), 4125, *otp.PCG_BAND))
        want = otp.segment(v, 4125, onp.WindowSpec(2.0)).contiguous().numpy()
        assert rel_err(got[..., c], want) < TOL
    got2 = pkg.preprocess_segment(_dev(x), 4000, 4125, spec, fused=True).cpu().numpy()
    want2, _ = _oracle_torch(x[:, 3], 4000, 4125, 2.0)
    assert got2.shape[:2] == (2, 6) and rel_err(got2[:, 3], want2) < TOL


def test_fused_edge_lengths(pkg):
    spec = pkg.WindowSpec(4.0)
    for n in (300, 1000, 4000, 8001):        # shorter than the start pad / than one window / odd tails
        x = _spiky(2, n, seed=40 + n % 7, spikes=1) if n > 50 else np.ones((2, n), np.float32)
        want, _ = _oracle_torch(x, 2000, 4125, 4.0)
        got = pkg.preprocess_segment(_dev(x), 2000, 4125, spec, fused=True).cpu().numpy()
        assert got.shape == want.shape, n
        assert rel_err(got, want) < TOL or np.abs(want).max() == 0
        np.testing.assert_array_equal(got == 0, want == 0) if n <= 1000 else None
    assert pkg.preprocess_segment(torch.zeros(0, 5000, device="cuda"), 2000, 4125, spec).shape[0] == 0


def test_fused_vs_golden(pkg, golden):
    g = golden("preprocess_2k_4125.npz")
    got = pkg.preprocess_segment(_dev(g["pcg"]), 2000, 4125, pkg.WindowSpec(1.0), fused=True).cpu().numpy()
    assert rel_err(got, g["t64_windows"]) < TOL
    gotn = pkg.preprocess_segment(_dev(g["pcg"]), 2000, 4125, pkg.WindowSpec(1.0), mode="numpy", fused=True).cpu().numpy()
    assert rel_err(gotn, g["np_windows"]) < TOL


def test_fused_full_size_properties(pkg):
    """Config-2 size (1024 recordings x {PCG, ECG} x 30 s): properties that need no oracle."""
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(1024, 2, 60000, device="cuda", generator=g)
    x[:, 0, 7000:7004] += 25.0
    spec = pkg.WindowSpec(4.0)
    w = pkg.preprocess_segment(x, 2000, 4125, spec, kinds=("pcg", "ecg"), fused=True)
    assert w.shape == (1024, 2, 7, 16500)
    assert torch.isfinite(w).all() and float(w.abs().max()) <= 1.0
    # consecutive windows overlap by win - hop samples: the overlap must be bit-identical
    ov = 16500 - 15469
    assert torch.equal(w[:, :, :-1, -ov:], w[:, :, 1:, :ov])
    # every row was scaled to hit +-1 somewhere in the full row (not necessarily inside the windows) and is
    # unaffected by the other rows: recompute 3 rows alone
    sub = pkg.preprocess_segment(x[5:8].contiguous(), 2000, 4125, spec, kinds=("pcg", "ecg"), fused=True)
    assert torch.equal(sub, w[5:8])


@pytest.mark.parametrize("mode", ["torch", "numpy"])
def test_fast_despike_equals_reference_order(pkg, mode, monkeypatch):
    """The fused kernel's fast despike path (parallel frames, logged passes, exact handling of stuck frames) must
    leave exactly the samples and pass counts of its serial path, which replays the reference's order pass by
    pass (and is the one the trace tests pin to the oracle).  Synthetic PCG rows end stuck about 40 % of the time."""
    from wav2vec_heart_sounds_b200.synth import synth_pcg
    spec = pkg.WindowSpec(4.0)
    for seed, rows, n, fs_in in ((1234, 192, 60000, 2000), (77, 64, 32000, 4000), (5, 32, 9000, 2000)):
        x = synth_pcg(rows, n, float(fs_in), seed=seed, device="cuda")
        x[1, 500] += 40.0                                    # one-sample spike: the classic stuck frame
        x[2, 700:703] -= 35.0
        x[2, n // 2] += 30.0
        monkeypatch.delenv("MPCG_FZ_DESPIKE_SERIAL", raising=False)
        fast, e_fast = pkg.preprocess_segment(x, fs_in, 4125, spec, mode=mode, fused=True, return_edits=True)
        monkeypatch.setenv("MPCG_FZ_DESPIKE_SERIAL", "1")
        ser, e_ser = pkg.preprocess_segment(x, fs_in, 4125, spec, mode=mode, fused=True, return_edits=True)
        monkeypatch.delenv("MPCG_FZ_DESPIKE_SERIAL", raising=False)
        assert int(e_ser.sum()) > rows                       # the despiker really worked
        assert torch.equal(e_fast, e_ser)
        assert torch.equal(fast, ser)


def test_fast_despike_hand_over_paths(pkg, monkeypatch):
    """Rows that the fast despike path cannot settle from its logs and hands to the serial path: frames that need more
    than 16 passes (a long ringing burst eaten half-cycle by half-cycle), spans that exhaust the undo pool (no sign
    flips: every pass covers a whole frame), more frames above the threshold than one round's log.  Whatever the path,
    samples and pass counts must equal the serial path's."""
    rng = np.random.default_rng(7)
    n, fs_in = 60000, 2000
    t = np.arange(n)
    rows = []
    base = 0.05 * rng.standard_normal((6, n))
    ring = base[0].copy()                                    # 1: 60 half-cycles of a slowly decaying 40 Hz burst
    ring[20000:21500] += 8.0 * np.exp(-np.arange(1500) / 900.0) * np.sin(2 * np.pi * 40.0 * np.arange(1500) / fs_in)
    rows.append(ring)
    dc = base[1] + 3.0                                       # 2: never crosses zero; two big bumps in one slice
    dc[9000:9040] += 40.0
    dc[13000:13030] += 30.0
    dc[30000:30020] += 35.0
    rows.append(dc)
    many = base[2].copy()                                    # 3: a spike in every third frame (20 frames above the threshold)
    many[500::3000] += 6.0
    rows.append(many)
    rows.append(base[3] * 0.0)                               # 4: silence (median 0)
    alt = base[4].copy()                                     # 5: one-sample alternating spikes (stuck frames)
    alt[7001] += 20.0; alt[7002] -= 20.0; alt[33333] -= 25.0
    rows.append(alt)
    rows.append(base[5] + np.where((t // 3000) % 2 == 0, 1.0, 0.02) * np.sin(t / 5.0))   # 6: loud/quiet halves
    x = torch.from_numpy(np.stack(rows).astype(np.float32)).cuda()
    spec = pkg.WindowSpec(4.0)
    for mode in ("torch", "numpy"):
        monkeypatch.delenv("MPCG_FZ_DESPIKE_SERIAL", raising=False)
        fast, e_fast = pkg.preprocess_segment(x, fs_in, 4125, spec, mode=mode, fused=True, return_edits=True)
        monkeypatch.setenv("MPCG_FZ_DESPIKE_SERIAL", "1")
        ser, e_ser = pkg.preprocess_segment(x, fs_in, 4125, spec, mode=mode, fused=True, return_edits=True)
        monkeypatch.delenv("MPCG_FZ_DESPIKE_SERIAL", raising=False)
        assert torch.equal(e_fast, e_ser), (mode, e_fast.tolist(), e_ser.tolist())
        assert torch.equal(fast, ser), mode
        assert int(e_ser[0]) > 16 and int(e_ser[2]) >= 15
    # and the serial path itself against the float64 oracle on these rows (torch mode)
    want, _ = _oracle_torch(x.cpu().numpy(), fs_in, 4125, 4.0)
    got = pkg.preprocess_segment(x, fs_in, 4125, spec, fused=True).cpu().numpy()
    for r in range(x.shape[0]):
        assert rel_err(got[r], want[r]) < TOL or np.abs(want[r]).max() == 0, r


def test_fused_config5_chunk_properties(pkg):
    """configs[4] per-GPU chunk shape, scaled to 1024 recordings x 6 channels x 8 s at 4 kHz -> 4125 Hz, 2 s windows:
    2-CTA clusters, the 33/32 resampler instance, both output layouts."""
    from wav2vec_heart_sounds_b200.synth import synth_pcg
    r, c, t = 1024, 6, 32000
    x = synth_pcg(r * c, t, 4000.0, seed=9, device="cuda").reshape(r, c, t)
    spec = pkg.WindowSpec(2.0)
    planar = pkg.preprocess_segment(x, 4000, 4125, spec, fused=True)
    last = pkg.preprocess_segment(x, 4000, 4125, spec, channels_last=True, fused=True)
    assert planar.shape == (r, c, 4, 8250) and last.shape == (r, 4, 8250, c)
    assert torch.isfinite(planar).all() and float(planar.abs().max()) <= 1.0
    assert torch.equal(last.permute(0, 3, 1, 2), planar)
    ov = 8250 - 7219
    assert torch.equal(planar[:, :, :-1, -ov:], planar[:, :, 1:, :ov])       # overlapping windows carry the same samples
    sub = pkg.preprocess_segment(x[100:103].contiguous(), 4000, 4125, spec, fused=True)
    assert torch.equal(sub, planar[100:103])                                 # rows are independent
    # three recordings against the float64 oracle
    want, _ = _oracle_torch(x[7].cpu().numpy(), 4000, 4125, 2.0)
    assert rel_err(planar[7].cpu().numpy(), want) < TOL


def test_channel_major_layout_and_host_pipeline_augment(pkg):
    """channel_major output ([C, B, N, win]: every channel one contiguous batch of windows) equals the planar one;
    HostPipeline(augment=...) returns the ECG windows untouched and normalised augmented PCG windows."""
    from wav2vec_heart_sounds_b200.pipeline import HostPipeline
    from wav2vec_heart_sounds_b200.synth import synth_pair
    from wav2vec_heart_sounds_b200 import AugmentConfig
    x = synth_pair(70, 20000, 2000, seed=3, device="cuda")
    spec = pkg.WindowSpec(4.0)
    planar = pkg.preprocess_segment(x, 2000, 4125, spec, kinds=("pcg", "ecg"), fused=True)
    cm = pkg.preprocess_segment(x, 2000, 4125, spec, kinds=("pcg", "ecg"), fused=True, channel_major=True)
    assert cm.shape == (2, 70) + tuple(planar.shape[2:])
    assert torch.equal(cm.permute(1, 0, 2, 3), planar)
    with pytest.raises(ValueError):
        pkg.preprocess_segment(x, 2000, 4125, spec, kinds=("pcg", "ecg"), fused=False, channel_major=True)
    hp = HostPipeline(70, 2, 20000, 2000, 4125, spec, kinds=("pcg", "ecg"), chunk=32, augment=AugmentConfig())
    oh = hp.empty_output()
    torch.manual_seed(0); np.random.seed(0)
    hp(x.cpu().pin_memory(), oh)                             # (no synchronize: the call returns complete host buffers)
    assert oh.shape == cm.shape
    assert torch.equal(oh[1], cm[1].cpu())                                   # ECG windows pass through
    a = oh[0].reshape(-1, oh.shape[-1])
    assert torch.isfinite(a).all() and float(a.abs().max()) <= 1.0
    assert float(a.mean(dim=1).abs().max()) < 1e-5 and float((a.abs().amax(dim=1) - 1).abs().max()) < 1e-5
    plain = HostPipeline(70, 2, 20000, 2000, 4125, spec, kinds=("pcg", "ecg"), chunk=32)
    op = plain.empty_output()
    plain(x.cpu().pin_memory(), op)
    assert torch.equal(op, planar.cpu())


def test_no_writes_outside_the_output_and_workspace(pkg):
    """compute-sanitizer is not available on the GPU pool, so the hot kernel's global stores are fenced in by hand: the
    output tensor and the workspace are carved out of larger buffers whose margins hold a sentinel, for single-tile,
    multi-tile (despiked, parked rows), 8/1, ragged and channels-last launches; the margins must come back untouched
    and every output element must have been written (no sentinel left inside)."""
    import ctypes
    from wav2vec_heart_sounds_b200 import _lib
    from wav2vec_heart_sounds_b200.synth import synth_pair
    SENT = -123456.0
    pad = 4096
    x = synth_pair(6, 60000, 2000, seed=17, device="cuda")
    x[1, 0, 500] += 40.0
    cases = [dict(fs_out=4125, ws=4.0, kw=dict(kinds=("pcg", "ecg"), channel_major=True), xin=x),
             dict(fs_out=4125, ws=4.0, kw=dict(kinds=("pcg", "ecg"), channels_last=True), xin=x),
             dict(fs_out=16000, ws=4.0, kw=dict(), xin=x[:2, 0].contiguous()),
             dict(fs_out=4125, ws=1.0, kw=dict(mode="numpy"), xin=x[:, 0, :9001].contiguous())]
    for c in cases:
        ref = pkg.preprocess_segment(c["xin"], 2000, c["fs_out"], pkg.WindowSpec(c["ws"]), fused=True, **c["kw"])
        big = torch.full((ref.numel() + 2 * pad,), SENT, device="cuda")
        out = big[pad:pad + ref.numel()].view(ref.shape)
        # a private, guarded workspace for this stream: swap it into the shim's cache
        key = ("pre", c["xin"].device.index, torch.cuda.current_stream().cuda_stream)
        need = _lib.workspace(c["xin"], 1).numel()
        wbig = torch.full((need + 2 * pad,), 0x5A, dtype=torch.uint8, device="cuda")
        lo = (-wbig.data_ptr() - pad) % 16 + pad                      # keep the carved workspace 16-byte aligned
        old = _lib._workspaces[key]
        _lib._workspaces[key] = wbig[lo:lo + need]
        try:
            pkg.preprocess_segment(c["xin"], 2000, c["fs_out"], pkg.WindowSpec(c["ws"]), fused=True, out=out, **c["kw"])
            torch.cuda.synchronize()
        finally:
            _lib._workspaces[key] = old
        assert bool((big[:pad] == SENT).all()) and bool((big[pad + ref.numel():] == SENT).all())
        assert bool((wbig[:lo] == 0x5A).all()) and bool((wbig[lo + need:] == 0x5A).all())
        assert not bool((out == SENT).any())
        assert torch.equal(out, ref)
