"""Parity of the augmentation kernels with the reference's torchaug (golden vectors produced by running the
reference with a replayed RNG, tests/golden/torchaug_replay.npz) and with the oracle in float64.
Tolerance 1e-5 of the output scale (outputs are normalised to [-1, 1])."""
import numpy as np
import pytest
import torch

from oracle import torch_path as otp
from helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def ta(built_lib):
    from wav2vec_heart_sounds_b200 import torchaug
    return torchaug


def _dev(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float32).cuda()


def test_normalise_matches_reference(ta, golden):
    g = golden("torchaug_replay.npz")
    x = torch.from_numpy(g["x"]) * 3.0 + 0.7
    got = ta._normalise(x.cuda()).cpu().numpy()
    assert rel_err(got, otp.renormalise(x.double()).numpy()) < TOL


def test_noise_sine_wander_warp_vs_reference_outputs(ta, golden):
    g = golden("torchaug_replay.npz")
    x = _dev(g["x"])
    fs = int(g["fs"])
    got = ta.add_white_noise(x, std=float(g["noise_std"]), scale=g["noise_scale"], noise=_dev(g["noise_noise"]))
    assert rel_err(got.cpu().numpy(), g["noise_out"]) < TOL
    got = ta.sinusoidal_envelope(x, fs, amp=g["sine_amp"], freq=g["sine_freq"], phase=g["sine_phase"])
    assert rel_err(got.cpu().numpy(), g["sine_out"]) < TOL
    got = ta.baseline_wander(x, fs, amp=g["wander_amp"], freq=g["wander_freq"], phase=g["wander_phase"])
    assert rel_err(got.cpu().numpy(), g["wander_out"]) < TOL
    got = ta.amplitude_warp(x, amps=g["warp_amps"])
    assert rel_err(got.cpu().numpy(), g["warp_out"]) < TOL


def test_parametric_eq_vs_float64_reference(ta, golden):
    g = golden("torchaug_replay.npz")
    x = _dev(g["x"])
    bands = [tuple(b) for b in g["eq_bands"]]
    got = ta.parametric_eq(x, int(g["fs"]), 2, 500, bands=bands).cpu().numpy()
    assert rel_err(got, g["eq_out64"]) < TOL                 # the reference run in float64 is the target
    assert rel_err(got, g["eq_out"]) < 5e-3                  # its own float32 run is only this close to it


def test_whole_chain_vs_reference_and_oracle(ta, golden):
    g = golden("torchaug_replay.npz")
    from wav2vec_heart_sounds_b200 import AugmentConfig
    x = torch.from_numpy(g["x"])
    draws = {}
    for k in g.files:
        if k.startswith("chain_") and k != "chain_out":
            v = g[k]
            key = k[len("chain_"):]
            draws[key] = float(v) if key.startswith("std") else ([tuple(b) for b in v] if key == "bands" else v)
    cfg = AugmentConfig(prob_noise=2.0, prob_wandering_volume=0.75, prob_banding=0.6)
    dev_draws = {k: (_dev(v) if isinstance(v, np.ndarray) and v.dtype != object and k not in ("bands",) else v)
                 for k, v in draws.items()}
    got = ta.augment_pcg_batch(x.cuda(), int(g["fs"]), cfg, draws=dev_draws).cpu().numpy()
    assert got.shape == x.shape and np.isfinite(got).all() and np.abs(got).max() <= 1.0 + 1e-6
    odraws = {k: (torch.from_numpy(np.asarray(v)) if isinstance(v, np.ndarray) else v) for k, v in draws.items()}
    want64 = otp.augment_pcg_batch(x.double(), int(g["fs"]), odraws).numpy()
    assert rel_err(got, want64) < TOL
    assert rel_err(got, g["chain_out"]) < 5e-3               # the reference's float32 EQ is the noisy party
    masks = [draws[f"mask{i}"].reshape(-1) for i in (1, 2, 3, 4)]
    assert any(m.any() for m in masks) and not all(m.all() for m in masks)   # both branches exercised


def test_apply_blend_and_masks(ta):
    x = torch.randn(6, 3001, device="cuda")
    tr = torch.randn(6, 3001, device="cuda")
    m = torch.tensor([1, 0, 1, 1, 0, 0.0]).reshape(6, 1)
    got = ta._apply(x, tr, 0.5, mask=m).cpu()
    want = otp.blend(x.cpu().double(), tr.cpu().double(), m.double())
    assert rel_err(got.numpy(), want.numpy()) < TOL


def test_reference_property_test_default_rng(ta):
    """The reference's own test (tests/test_torchaug.py:9-15) with the default random draws."""
    from wav2vec_heart_sounds_b200 import AugmentConfig, augment_pcg_batch
    x = torch.randn(8, 4125, device="cuda")
    torch.manual_seed(0); np.random.seed(0)
    out = augment_pcg_batch(x, fs=4125, cfg=AugmentConfig(prob_hpss=0.0, prob_real_noise=0.0))
    assert out.shape == x.shape and torch.isfinite(out).all() and float(out.abs().max()) <= 1.0 + 1e-5
    torch.manual_seed(0); np.random.seed(0)
    again = augment_pcg_batch(x, fs=4125, cfg=AugmentConfig(prob_hpss=0.0, prob_real_noise=0.0))
    assert torch.equal(out, again)                           # seeded runs repeat


def test_default_draw_order_equals_reference_stream(ta):
    """With the same seeds the default path consumes numpy's and torch's CUDA generators like the reference:
    replay the reference's call order on the device and inject it -- results must be identical."""
    x = ta._normalise(torch.randn(5, 2048, device="cuda"))
    np.random.seed(7); torch.manual_seed(7)
    a = ta.add_white_noise(x)
    np.random.seed(7); torch.manual_seed(7)
    std = float(np.random.choice((0.0001, 0.001, 0.01)))
    scale = torch.rand(5, 1, device="cuda") * 0.1
    noise = torch.randn_like(x)
    assert torch.equal(a, ta.add_white_noise(x, std=std, scale=scale, noise=noise))
    assert torch.equal(a, x + scale * std * noise)           # and that IS the reference's expression


def test_philox_noise_statistics_and_repeatability(ta):
    x = torch.zeros(64, 64000, device="cuda")
    torch.manual_seed(3)
    a = ta.add_white_noise(x, std=0.01, scale=torch.full((64, 1), 0.1), noise="philox")
    z = (a / (0.1 * 0.01)).double()
    assert abs(float(z.mean())) < 3e-3 and abs(float(z.std()) - 1.0) < 3e-3
    assert abs(float((z ** 4).mean()) - 3.0) < 0.05          # Gaussian kurtosis
    c = torch.corrcoef(torch.stack([z[0], z[1], z[0].roll(1)]))
    assert float(c[0, 1].abs()) < 0.02 and float(c[0, 2].abs()) < 0.02
    torch.manual_seed(3)
    b = ta.add_white_noise(x, std=0.01, scale=torch.full((64, 1), 0.1), noise="philox")
    assert torch.equal(a, b)


def test_edge_shapes(ta):
    x = torch.randn(3, 1001, device="cuda")                 # odd length: tail group of the 4-sample loop
    out = ta.sinusoidal_envelope(x, 1000, amp=np.full((3, 2), 0.1), freq=np.full((3, 2), 0.3), phase=np.zeros((3, 2)))
    want = otp.sinusoidal_envelope(x.cpu(), 1000, torch.full((3, 2), 0.1), torch.full((3, 2), 0.3), torch.zeros(3, 2))
    assert rel_err(out.cpu().numpy(), want.numpy()) < TOL
    assert ta._normalise(torch.zeros(0, 16, device="cuda")).shape == (0, 16)
    with pytest.raises(ValueError):
        ta._normalise(torch.zeros(16, device="cuda"))
    with pytest.raises(ValueError, match="no CPU fallback"):
        ta.add_white_noise(torch.zeros(2, 16))
    w = ta.amplitude_warp(torch.randn(2, 5000, device="cuda"), amps=np.ones((2, 12)))    # flat gains = 65-tap mean
    assert w.shape == (2, 5000)


def test_time_warp_definition(ta):
    """y[j] = x(j * rate) with Catmull-Rom interpolation (this build's definition; parity with rubberband unpinned)."""
    rng = np.random.default_rng(9)
    t = 5000
    x = np.cumsum(rng.standard_normal((2, t)), axis=1).astype(np.float32) * 0.05
    for rate in (1.005, 0.8, 1.3):
        got = ta.time_warp(torch.from_numpy(x).cuda(), 4125, rate).cpu().numpy()
        n = int(round(t / rate))
        assert got.shape == (2, n)
        pos = np.arange(n) * rate
        i = np.floor(pos).astype(int); u = (pos - i).astype(np.float32)
        at = lambda k: x[:, np.clip(k, 0, t - 1)]
        p0, p1, p2, p3 = at(i - 1), at(i), at(i + 1), at(i + 2)
        want = ((( -0.5 * p0 + 1.5 * p1 - 1.5 * p2 + 0.5 * p3) * u + (p0 - 2.5 * p1 + 2 * p2 - 0.5 * p3)) * u
                + (-0.5 * p0 + 0.5 * p2)) * u + p1
        assert rel_err(got, want) < 1e-5
    same = ta.time_warp(torch.from_numpy(x).cuda(), 4125, 1.0).cpu().numpy()
    np.testing.assert_array_equal(same, x)                                     # rate 1 is the identity
    assert ta.time_warp(torch.from_numpy(x).cuda(), 4125, 0.7, keep_length=True).shape == (2, t)


def test_mix_noise_arithmetic(ta):
    from oracle import numpy_path as onp
    rng = np.random.default_rng(10)
    x = rng.standard_normal((3, 4000)).astype(np.float32)
    bank = (rng.standard_normal((2, 9000)) * 3 + 1).astype(np.float32)
    rows, starts, scale = [1, 0, 1], [0, 5000, 1234], [0.3, 0.0, 0.11]
    got = ta.mix_noise(torch.from_numpy(x).cuda(), torch.from_numpy(bank).cuda(), rows=rows, starts=starts, scale=scale)
    for r in range(3):
        crop = onp.abs_max_normalise(bank[rows[r], starts[r]:starts[r] + 4000])
        want = onp.abs_max_normalise(x[r].astype(np.float64) + scale[r] * crop)
        assert rel_err(got[r].cpu().numpy(), want) < 1e-5
    out = ta.mix_noise(torch.from_numpy(x).cuda(), torch.from_numpy(bank).cuda())          # random draws
    assert out.shape == (3, 4000) and float(out.abs().max()) <= 1.0
    for bad in (dict(rows=[2, 0, 1], starts=starts), dict(rows=rows, starts=[0, 5001, 0]), dict(rows=[-1, 0, 0], starts=starts)):
        with pytest.raises(ValueError):                      # injected tables are checked before the kernel trusts them
            ta.mix_noise(torch.from_numpy(x).cuda(), torch.from_numpy(bank).cuda(), scale=scale, **bad)


@pytest.mark.parametrize("t,rows", [(64000, 24), (8250, 40), (12347, 16), (133000, 5), (600, 9), (38000, 6)])
@pytest.mark.parametrize("noise", [None, "philox"])
def test_fused_chain_equals_stage_kernels(ta, t, rows, noise):
    """The one-kernel chain (rows resident in cluster shared memory) against the kernel-per-stage path on the same
    random draws: the per-element arithmetic is the same, only the summation order of the row means differs, so the
    two agree to float32 rounding (1e-6 of the [-1, 1] range).  Covers 1-, 2-, 4- and 8-CTA clusters, all four compiled
    chunk lengths (600 -> 9, 8250 -> 17, 12347 -> 25, the others 33), rows that do not start on 16-byte boundaries
    (12347), a cluster whose last CTA owns nothing (38000: four slices of 12800), ragged tails, rows with every mask
    on and with every mask off."""
    from wav2vec_heart_sounds_b200 import AugmentConfig
    g = torch.Generator(device="cuda").manual_seed(t)
    x = torch.randn(rows, t, device="cuda", generator=g) * 0.3 + torch.sin(torch.arange(t, device="cuda") / 37.0)[None]
    for cfg in (AugmentConfig(), AugmentConfig(prob_noise=4.0, prob_wandering_volume=1.0, prob_banding=1.0),
                AugmentConfig(prob_noise=0.0, prob_wandering_volume=0.0, prob_banding=0.0)):
        outs = []
        for fused, collapse in ((True, True), (True, False), (False, False)):
            torch.manual_seed(3); np.random.seed(3)
            outs.append(ta.augment_pcg_batch(x, 4125, cfg, noise=noise, fused=fused, collapse=collapse,
                                             fast_draws=False if fused else None))
        assert torch.isfinite(outs[0]).all()
        assert float((outs[1] - outs[2]).abs().max()) < 2e-6       # every stage re-normalised: float32 rounding apart
        assert float((outs[0] - outs[2]).abs().max()) < 4e-6       # idempotent re-normalisations collapsed


def test_fused_chain_on_rows_with_a_large_offset(ta):
    """A raw row whose offset dwarfs its swing (mean / peak up to 1e4) takes the two-term form of the first map; later
    maps are the one-FFMA form.  Same result as the stage kernels (float64 maps) to float32 rounding, and as the
    chain on the centred rows (N removes the offset)."""
    from wav2vec_heart_sounds_b200 import AugmentConfig
    g = torch.Generator(device="cuda").manual_seed(9)
    base = torch.randn(12, 16500, device="cuda", generator=g) * 0.01
    off = torch.tensor([0.0, 1.0, -3.0, 50.0, -100.0, 0.5, 7.0, -0.2, 20.0, 100.0, -60.0, 2.0], device="cuda")[:, None]
    x = base + off
    cfg = AugmentConfig(prob_noise=4.0, prob_wandering_volume=1.0, prob_banding=1.0)
    outs = []
    for fused in (True, False):
        torch.manual_seed(3); np.random.seed(3)
        outs.append(ta.augment_pcg_batch(x, 4125, cfg, noise="philox", fused=fused, collapse=False,
                                         fast_draws=False if fused else None))
    assert torch.isfinite(outs[0]).all() and float(outs[0].abs().max()) <= 1.0
    assert float((outs[0] - outs[1]).abs().max()) < 2e-6
    # with every stage off and the re-normalisations collapsed the ONE normalisation that is left has to be exact on its
    # own: float32 partial sums of values around -300 +- 0.003 once put the row mean 5e-5 of the range off
    x2 = torch.randn(4, 130741, device="cuda", generator=g) * 1e-3 + torch.tensor([-300.0, 5.0, 4000.0, 0.0], device="cuda")[:, None]
    none = AugmentConfig(prob_noise=0.0, prob_wandering_volume=0.0, prob_banding=0.0)
    y1 = ta.augment_pcg_batch(x2, 16000, none, noise="philox", fused=True, collapse=True)
    want = x2.double() - x2.double().mean(dim=1, keepdim=True)
    want = (want / want.abs().amax(dim=1, keepdim=True)).clamp(-1, 1)
    assert float((y1.double() - want).abs().max()) < 1e-6
    assert float(y1.double().mean(dim=1).abs().max()) < 1e-6
    const = torch.full((3, 4000), 2.5, device="cuda")                       # degenerate rows: zero swing
    y = ta.augment_pcg_batch(const, 4125, AugmentConfig(prob_noise=0.0, prob_wandering_volume=0.0, prob_banding=0.0))
    assert torch.isfinite(y).all() and float(y.abs().max()) <= 1.0


def test_fused_chain_rejects_rows_beyond_a_cluster(ta):
    x = torch.randn(2, 140000, device="cuda")
    with pytest.raises(ValueError):
        ta.augment_pcg_batch(x, 4125, fused=True)
    torch.manual_seed(1); np.random.seed(1)
    a = ta.augment_pcg_batch(x, 4125)                      # falls back to the stage kernels
    torch.manual_seed(1); np.random.seed(1)
    b = ta.augment_pcg_batch(x, 4125, fused=False)
    assert torch.equal(a, b)


def test_fused_chain_full_size_properties(ta):
    """configs[2] size (4096 windows x 64000 samples @16 kHz): properties that need no oracle."""
    from wav2vec_heart_sounds_b200 import AugmentConfig
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(4096, 64000, device="cuda", generator=g)
    cfg = AugmentConfig()
    torch.manual_seed(5); np.random.seed(5)
    y = ta.augment_pcg_batch(x, 16000, cfg, noise="philox")
    assert y.shape == x.shape and torch.isfinite(y).all() and float(y.abs().max()) <= 1.0
    # every row ends normalised: zero mean, unit peak (up to float32 rounding of the last map)
    assert float(y.mean(dim=1).abs().max()) < 1e-5 and float((y.abs().amax(dim=1) - 1).abs().max()) < 1e-5
    # rows are independent: the same draws on a slice of the batch reproduce those rows bit for bit
    torch.manual_seed(5); np.random.seed(5)
    full_again = ta.augment_pcg_batch(x, 16000, cfg, noise="philox")
    assert torch.equal(y, full_again)                                       # seeded runs repeat
    # with every mask off the chain is one normalisation
    off = AugmentConfig(prob_noise=0.0, prob_wandering_volume=0.0, prob_banding=0.0)
    z = ta.augment_pcg_batch(x[:512], 16000, off, noise="philox")
    assert float((z - ta._normalise(x[:512])).abs().max()) < 1e-6


def test_fast_draws_distributions(ta):
    """Throughput mode (noise="philox", no injected draws): the per-row quantities come from two device draws; their
    distributions are the reference's (torchaug.py:39-54,103-111) and the run is reproducible under torch / numpy seeds."""
    from wav2vec_heart_sounds_b200 import AugmentConfig
    b = 20000
    scale, offset = ta._fast_draw_affine(0.01, 0.001, torch.device("cuda"))
    torch.manual_seed(1)
    tab = torch.addcmul(offset, torch.rand(3, b, 8, device="cuda"), scale)
    assert float(tab[0, :, 0].min()) >= 0 and float(tab[0, :, 0].max()) <= 0.1 * 0.01 and float(tab[0, :, 1:].abs().max()) == 0
    assert 0.0009 * 0.1 * 0.5 * 0.9 < float(tab[2, :, 0].mean()) * 0.9 and float(tab[2, :, 0].max()) <= 0.1 * 0.001 * 1.0000001
    amp, freq, phase = tab[1, :, 0], tab[1, :, 1], tab[1, :, 2]
    assert 0.01 <= float(amp.min()) and float(amp.max()) <= 0.25 and abs(float(amp.mean()) - 0.13) < 0.005
    assert 0.05 <= float(freq.min()) and float(freq.max()) <= 0.5 and 0.001 <= float(tab[1, :, 4].min()) and float(tab[1, :, 4].max()) <= 0.05
    assert 0 <= float(phase.min()) and float(phase.max()) < 1 and float(tab[1, :, 6:].abs().max()) == 0
    # the one-launch draw kernel augment_pcg_batch uses (mpcg_aug_draw_f32): the same table from its own Philox stream
    from wav2vec_heart_sounds_b200 import _lib
    sc, of = ta._fast_draw_affine_host(0.01, 0.001)
    probs = np.array([0.075, 0.75, 0.25, 0.075], np.float32)
    t2, mk = torch.empty(3, b, 8, device="cuda"), torch.empty(4, b, device="cuda")
    _lib.check(_lib.lib().mpcg_aug_draw_f32(t2.data_ptr(), mk.data_ptr(), b, sc.ctypes.data, of.ctypes.data, probs.ctypes.data,
                                            5, 6, torch.cuda.current_stream().cuda_stream), "draw")
    assert np.allclose(sc, scale.cpu().numpy()[:, 0]) and np.allclose(of, offset.cpu().numpy()[:, 0])
    assert float(t2[0, :, 0].min()) >= 0 and float(t2[0, :, 0].max()) <= 0.1 * 0.01 and float(t2[0, :, 1:].abs().max()) == 0
    assert abs(float(t2[0, :, 0].mean()) - 0.5e-3) < 2e-5 and abs(float(t2[2, :, 0].mean()) - 0.5e-4) < 2e-6
    assert 0.01 <= float(t2[1, :, 0].min()) and float(t2[1, :, 0].max()) <= 0.25 and abs(float(t2[1, :, 0].mean()) - 0.13) < 0.005
    assert 0.05 <= float(t2[1, :, 1].min()) and float(t2[1, :, 1].max()) <= 0.5 and abs(float(t2[1, :, 1].mean()) - 0.275) < 0.01
    assert 0.001 <= float(t2[1, :, 4].min()) and float(t2[1, :, 4].max()) <= 0.05 and float(t2[1, :, 6:].abs().max()) == 0
    assert 0 <= float(t2[1, :, 5].min()) and float(t2[1, :, 5].max()) < 1 and abs(float(t2[1, :, 5].mean()) - 0.5) < 0.01
    for i in range(4):
        assert set(mk[i].unique().tolist()) <= {0.0, 1.0} and abs(float(mk[i].mean()) - float(probs[i])) < 0.012
    cols = torch.stack([t2[0, :, 0], t2[1, :, 0], t2[1, :, 1], t2[1, :, 2], t2[1, :, 3], t2[2, :, 0], mk[1], mk[2]])
    cc = torch.corrcoef(cols)                                           # the draws of a row are independent of each other
    assert float((cc - torch.eye(8, device="cuda")).abs().max()) < 0.03
    t3, mk3 = torch.empty_like(t2), torch.empty_like(mk)
    _lib.check(_lib.lib().mpcg_aug_draw_f32(t3.data_ptr(), mk3.data_ptr(), b, sc.ctypes.data, of.ctypes.data, probs.ctypes.data,
                                            5, 7, torch.cuda.current_stream().cuda_stream), "draw")
    assert not torch.equal(t2, t3)                                       # another stream id: another table
    x = torch.randn(512, 4125, device="cuda")
    outs = []
    for _ in range(2):
        torch.manual_seed(9); np.random.seed(9)
        outs.append(ta.augment_pcg_batch(x, 4125, AugmentConfig(), noise="philox"))
    assert torch.equal(outs[0], outs[1])
    base = ta._normalise(x)
    changed = ((outs[0] - base).abs().amax(dim=1) > 1e-4).float().mean()
    assert 0.6 < float(changed) < 0.95                         # P(at least one stage on) = 1 - .925 * .25 * .75 * .925 = 0.84


def test_chain_degenerate_inputs(ta):
    """Empty batches, five-sample rows, constant rows and a strided (transposed) input all go through."""
    assert ta.augment_pcg_batch(torch.zeros(0, 100, device="cuda"), 4125).shape == (0, 100)
    assert ta.augment_pcg_batch(torch.zeros(0, 100, device="cuda"), 4125, fused=False).shape == (0, 100)
    tiny = ta.augment_pcg_batch(torch.randn(3, 5, device="cuda"), 4125, noise="philox")
    assert tiny.shape == (3, 5) and torch.isfinite(tiny).all() and float(tiny.abs().max()) <= 1.0
    const = ta.augment_pcg_batch(torch.ones(2, 64, device="cuda"), 4125, cfg=None, draws=None, noise="philox")
    assert torch.isfinite(const).all()
    xt = torch.randn(3001, 5, device="cuda").t()
    torch.manual_seed(2); np.random.seed(2)
    a = ta.augment_pcg_batch(xt, 4125, noise="philox")
    torch.manual_seed(2); np.random.seed(2)
    b = ta.augment_pcg_batch(xt.contiguous(), 4125, noise="philox")
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        ta.augment_pcg_batch(torch.randn(10, device="cuda"), 4125)


def test_more_rows_than_a_grid_dimension(ta):
    """amplitude_warp and time_warp put rows on gridDim.y (65 535 at most): bigger batches run as consecutive launches."""
    x = torch.randn(66000, 300, device="cuda")
    amps = 0.7 + torch.rand(66000, 12) * 0.6
    y = ta.amplitude_warp(x, amps=amps)
    for sl in (slice(0, 3), slice(65533, 65538), slice(65997, 66000)):
        assert torch.equal(y[sl], ta.amplitude_warp(x[sl], amps=amps[sl]))
    z = ta.time_warp(x, 4125, 1.02)
    assert torch.equal(z[65530:65540], ta.time_warp(x[65530:65540].contiguous(), 4125, 1.02))
