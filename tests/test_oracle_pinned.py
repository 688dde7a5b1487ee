"""Pin the oracle restatement (oracle/*.py) to the reference: against the committed golden vectors that
oracle/make_golden.py produced by RUNNING the reference, and -- when /root/reference is mounted (build
container) -- against the reference itself, bit for bit."""
import numpy as np
import pytest
import torch

from oracle import numpy_path as onp
from oracle import torch_path as otp
from conftest import reference_modules


def test_numpy_chain_matches_golden(golden):
    g = golden("preprocess_2k_4125.npz")
    fs_in, fs_out = float(g["fs_in"]), float(g["fs_out"])
    pcg, ecg = g["pcg"], g["ecg"]
    for r in range(pcg.shape[0]):
        rs = onp.resample(pcg[r].astype(np.float64), fs_in, fs_out)
        np.testing.assert_array_equal(rs, g["np_resample"][r])
        ds = onp.remove_spikes(rs, fs_out)
        np.testing.assert_array_equal(ds, g["np_despike"][r])
        bd = onp.bandpass_cascade(ds, fs_out, 25.0, 450.0)
        np.testing.assert_array_equal(bd, g["np_band"][r])
        np.testing.assert_array_equal(onp.abs_max_normalise(bd), g["np_norm"][r])
        np.testing.assert_array_equal(onp.preprocess_pcg(pcg[r], fs_in, fs_out), g["np_pcg"][r])
        np.testing.assert_array_equal(onp.preprocess_ecg(ecg[r], fs_in, fs_out), g["np_ecg"][r])
        np.testing.assert_array_equal(onp.segment(g["np_pcg"][r], fs_out, onp.WindowSpec(1.0)), g["np_windows"][r])
    pair = np.stack([g["np_pcg"][0], g["np_ecg"][0]], axis=1)
    np.testing.assert_array_equal(onp.segment(pair, fs_out, onp.WindowSpec(1.0)), g["np_windows_tc"])


def test_numpy_despike_does_real_work(golden):
    g = golden("preprocess_2k_4125.npz")
    assert np.any(g["np_despike"] != g["np_resample"])       # the fixture really contains spikes


@pytest.mark.parametrize("dtype,tag", [(torch.float64, "t64"), (torch.float32, "t32")])
def test_tensor_chain_matches_golden(golden, dtype, tag):
    g = golden("preprocess_2k_4125.npz")
    fs_in, fs_out = float(g["fs_in"]), float(g["fs_out"])
    xp = torch.from_numpy(g["pcg"]).to(dtype)
    xe = torch.from_numpy(g["ecg"]).to(dtype)
    rs = otp.resample(xp, fs_in, fs_out)
    ds = otp.remove_spikes(rs, fs_out)
    np.testing.assert_array_equal(ds.numpy(), g[f"{tag}_despike"])
    np.testing.assert_array_equal(otp.preprocess_pcg(xp, fs_in, fs_out).numpy(), g[f"{tag}_pcg"])
    np.testing.assert_array_equal(otp.preprocess_ecg(xe, fs_in, fs_out).numpy(), g[f"{tag}_ecg"])
    if tag == "t64":
        np.testing.assert_array_equal(rs.numpy(), g["t64_resample"])
        bd = otp.bandpass_cascade(ds, fs_out, 25.0, 450.0)
        np.testing.assert_array_equal(bd.numpy(), g["t64_band"])
        np.testing.assert_array_equal(otp.abs_max_normalise(bd).numpy(), g["t64_norm"])
        w = otp.segment(otp.preprocess_pcg(xp, fs_in, fs_out), fs_out, onp.WindowSpec(1.0))
        np.testing.assert_array_equal(w.contiguous().numpy(), g["t64_windows"])


def test_other_ratios_match_golden(golden):
    g = golden("resample_ratios.npz")
    x2, x4 = g["x_2k"], g["x_4k"]
    for r in range(2):
        np.testing.assert_array_equal(onp.resample(x2[r].astype(np.float64), 2000, 16000), g["np_2k_16k"][r])
        np.testing.assert_array_equal(onp.resample(x4[r].astype(np.float64), 4000, 4125), g["np_4k_4125"][r])
        np.testing.assert_array_equal(onp.preprocess_pcg(x2[r], 2000, 16000), g["np_pcg_2k_16k"][r])
    np.testing.assert_array_equal(otp.resample(torch.from_numpy(x2).double(), 2000, 16000).numpy(), g["t64_2k_16k"])
    np.testing.assert_array_equal(otp.resample(torch.from_numpy(x4).double(), 4000, 4125).numpy(), g["t64_4k_4125"])
    np.testing.assert_array_equal(otp.preprocess_pcg(torch.from_numpy(x2).double(), 2000, 16000).numpy(),
                                  g["t64_pcg_2k_16k"])
    np.testing.assert_array_equal(otp.preprocess_ecg(torch.from_numpy(x2).double(), 2000, 16000).numpy(),
                                  g["t64_ecg_16k"])


def test_window_index_table(golden):
    """Integer decisions of segmentation: bit-exact against the reference's window_starts()."""
    tab = golden("segment_index.npz")["table"]
    for fs, ws, n, win, hop, start, count, first, last in tab:
        spec = onp.WindowSpec(float(ws))
        assert spec.window_len(fs) == int(win) and spec.hop_len(fs) == int(hop)
        st = onp.window_starts(int(n), fs, spec)
        assert len(st) == int(count)
        if st:
            assert st[0] == int(first) and st[-1] == int(last)


def test_augment_replay_matches_golden(golden):
    g = golden("torchaug_replay.npz")
    x = torch.from_numpy(g["x"])
    fs = int(g["fs"])
    out = otp.add_white_noise(x, float(g["noise_std"]), torch.from_numpy(g["noise_scale"]),
                              torch.from_numpy(g["noise_noise"]))
    np.testing.assert_array_equal(out.numpy(), g["noise_out"])
    out = otp.sinusoidal_envelope(x, fs, *(torch.from_numpy(g[k]) for k in ("sine_amp", "sine_freq", "sine_phase")))
    np.testing.assert_array_equal(out.numpy(), g["sine_out"])
    out = otp.baseline_wander(x, fs, *(torch.from_numpy(g[k]) for k in ("wander_amp", "wander_freq", "wander_phase")))
    np.testing.assert_array_equal(out.numpy(), g["wander_out"])
    np.testing.assert_array_equal(otp.amplitude_warp(x, torch.from_numpy(g["warp_amps"])).numpy(), g["warp_out"])
    bands = [tuple(b) for b in g["eq_bands"]]
    np.testing.assert_array_equal(otp.parametric_eq(x, fs, bands).numpy(), g["eq_out"])
    np.testing.assert_array_equal(otp.parametric_eq(x.double(), fs, bands).numpy(), g["eq_out64"])
    draws = {k[len("chain_"):]: g[k] for k in g.files if k.startswith("chain_") and k != "chain_out"}
    for k, v in list(draws.items()):
        if k.startswith(("std",)):
            draws[k] = float(v)
        elif k == "bands":
            draws[k] = [tuple(b) for b in v]
        else:
            draws[k] = torch.from_numpy(np.asarray(v))
    np.testing.assert_array_equal(otp.augment_pcg_batch(x, fs, draws).numpy(), g["chain_out"])


def test_mel_presets_match_golden(golden):
    import warnings
    g = golden("mel_presets.npz")
    x = torch.from_numpy(g["x"])
    presets = {"dw4k": dict(sample_rate=4000, n_fft=1024, hop_length=256, n_mels=80, f_max=500.0),
               "c4_16k": dict(sample_rate=16000, n_fft=1024, hop_length=256, n_mels=80, f_max=500.0),
               "wg4k": dict(sample_rate=4000, n_fft=2048, win_length=1200, hop_length=300, n_mels=128, f_max=500.0)}
    for tag, kw in presets.items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            tr = otp.mel_transform(**kw)
            np.testing.assert_array_equal(otp.log_mel(x, tr).numpy(), g[f"{tag}_logmel"])


def test_against_live_reference_when_mounted():
    """Build container only: fresh random inputs through the reference and the oracle, bit for bit."""
    mods = reference_modules()
    if mods is None:
        pytest.skip("reference not mounted (GPU box)")
    sp, tp, ta = mods
    rng = np.random.default_rng(7)
    x = rng.standard_normal((2, 4000)).astype(np.float32)
    x[0, 1000:1005] += 30.0
    x[1, 2500:2503] -= 25.0
    for r in range(2):
        np.testing.assert_array_equal(onp.preprocess_pcg(x[r], 2000, 4125), sp.preprocess_pcg(x[r], 2000, 4125))
        np.testing.assert_array_equal(onp.preprocess_ecg(x[r], 2000, 4125), sp.preprocess_ecg(x[r], 2000, 4125))
    xt = torch.from_numpy(x).double()
    np.testing.assert_array_equal(otp.preprocess_pcg(xt, 2000, 4125).numpy(), tp.preprocess_pcg(xt, 2000, 4125).numpy())
    np.testing.assert_array_equal(otp.remove_spikes(xt.float(), 2000).numpy(), tp.remove_spikes(xt.float(), 2000).numpy())
    from mpcg_wav2vec.signalproc.segment import WindowSpec as RefSpec
    w1 = otp.segment(xt, 2000, onp.WindowSpec(1.0)).numpy()
    np.testing.assert_array_equal(w1, tp.segment(xt, 2000, RefSpec(1.0)).numpy())


def test_generator_conditioning_matches_golden(golden):
    """fade / fit_length / add_chirp of the oracle against the reference's own functions
    (datasets/generative.py:36-42, signalproc/preprocess.py:45-64), bit for bit."""
    g = golden("gen_condition.npz")
    fs, crop = int(g["fs"]), int(g["crop"])
    for tag in ("long", "exact", "short", "tiny"):
        x = g[f"{tag}_x"]
        for r in range(x.shape[0]):
            ref, _, chirp = onp.generator_item(x[r], x[r], fs, crop)
            np.testing.assert_array_equal(ref, g[f"{tag}_y"][r])
            np.testing.assert_array_equal(chirp, g[f"{tag}_chirp"][r])


def test_zero_phase_filters_match_golden(golden):
    """the oracle's zero-phase filters against the reference's signalproc/filters.py:44-90, bit for bit."""
    g = golden("zerophase.npz")
    fs, x = float(g["fs"]), g["x"]
    for r in range(x.shape[0]):
        np.testing.assert_array_equal(onp.butter_bandpass_zp(x[r], fs, 25.0, 400.0), g["bandpass"][r])
        np.testing.assert_array_equal(onp.butter_lowpass_zp(x[r], fs, 150.0), g["lowpass"][r])
        np.testing.assert_array_equal(onp.butter_highpass_zp(x[r], fs, 20.0), g["highpass"][r])
        np.testing.assert_array_equal(onp.band_stop_zp(x[r], fs, 45.0, 55.0), g["band_stop"][r])
        np.testing.assert_array_equal(onp.notch_zp(x[r], fs, 50.0), g["notch"][r])
        np.testing.assert_array_equal(onp.notch_chain_zp(x[r], fs, (50.0, 100.0, 150.0, 3000.0)), g["notch_chain"][r])
    np.testing.assert_array_equal(onp.decompose_bands(x[0], fs), g["bands"][0])


def test_other_normalisers_and_envelopes_match_golden(golden):
    """the oracle's min-max / z-score / k-peak normalisers and envelopes against the reference's
    signalproc/normalize.py:33-78 and signalproc/envelopes.py:11-23, bit for bit."""
    g = golden("normalisers.npz")
    x = g["x"]
    for r in range(x.shape[0]):
        np.testing.assert_array_equal(onp.minmax_normalise(x[r]), g["minmax"][r])
        np.testing.assert_array_equal(onp.minmax_normalise(x[r], 0.0, 2.0), g["minmax_02"][r])
        np.testing.assert_array_equal(onp.z_normalise(x[r]), g["z"][r])
        np.testing.assert_array_equal(onp.kpeak_normalise(x[r]), g["kpeak3"][r])
        np.testing.assert_array_equal(onp.kpeak_normalise(x[r], k=40, lo=0.0, hi=1.0), g["kpeak40"][r])
    np.testing.assert_array_equal(onp.minmax_normalise(g["flat"][0]), g["minmax_flat"][0])
    np.testing.assert_array_equal(onp.kpeak_normalise(g["flat"][0]), g["kpeak_flat"][0])
    xt = torch.from_numpy(x)
    np.testing.assert_array_equal(otp.minmax_normalise_torch(xt).numpy(), g["minmax_torch_all"])
    np.testing.assert_array_equal(otp.z_normalise_torch(xt.double()).numpy(), g["z_torch"])
    np.testing.assert_array_equal(otp.kpeak_normalise_torch(xt.double()).numpy(), g["kpeak_torch_all"])
    np.testing.assert_array_equal(otp.kpeak_normalise_torch(xt.double(), k=5, lo=0.0, hi=3.0).numpy(), g["kpeak_torch_all_k5"])
    e, fs = g["env_x"], float(g["env_fs"])
    for r in range(e.shape[0]):
        np.testing.assert_array_equal(onp.hilbert_envelope(e[r]), g["hilbert"][r])
        np.testing.assert_array_equal(onp.hilbert_envelope(e[r][:2499]), g["hilbert_odd"][r])
        np.testing.assert_array_equal(onp.homomorphic_envelope(e[r], fs), g["homomorphic"][r])


def test_heart_cycle_rebuild_matches_golden(golden):
    """split / crossfade / rebuild of the oracle and the host-side order draw of the package against
    the reference's datasets/heart_cycles.py run with the same seeded random.Random, bit for bit."""
    import random
    from wav2vec_heart_sounds_b200 import heart_cycles as H
    g = golden("heart_cycles.npz")
    joins, crop, fade_n = g["joins"].tolist(), int(g["crop"]), int(g["fade_n"])
    for (seed, pc), want in zip(g["orders_seed_pc"], g["orders6"]):
        assert H.rearrange_order(6, prob_contiguous=float(pc), rng=random.Random(int(seed))) == want.tolist()
    for r in range(3):
        sig = onp.abs_max_normalise(g["x"][r])
        cycles = onp.split_cycles(sig, joins)
        assert [(len(c)) for c in cycles] == [hi - lo for lo, hi in H.cycle_bounds(len(sig), joins)]
        order = H.rearrange_order(len(cycles), rng=random.Random(10 + r))
        assert order == g[f"order_{r}"].tolist()
        out = onp.rebuild([cycles[i] for i in order], crop, fade_n)
        np.testing.assert_array_equal(out, g[f"rebuilt_{r}"])
        np.testing.assert_array_equal(onp.fit_length(onp.fade(out), crop)[0], g[f"item_{r}"])
    cycles = onp.split_cycles(onp.abs_max_normalise(g["x"][0]), joins)
    np.testing.assert_array_equal(onp.rebuild(cycles, 100, fade_n), g["rebuilt_short_target"])
    np.testing.assert_array_equal(onp.rebuild(cycles[:2], 200000, fade_n), g["rebuilt_long_target"])     # the loop guard


def test_load_join_indices(tmp_path):
    """heart_cycles.py:22-29 (reference test tests/test_heart_cycles.py:17-20): rescaled, zero dropped, sorted, unique."""
    import json
    from wav2vec_heart_sounds_b200 import heart_cycles as H
    path = tmp_path / "p0.json"
    path.write_text(json.dumps({"segments": [[0], [1500, 1600], [500], [], [1000, 7], [500]], "last_index": 1500, "fs": 1000}))
    assert H.load_join_indices(path, fs_out=2000) == [1000, 2000, 3000]
    assert H.load_join_indices(path, fs_out=1000) == [500, 1000, 1500]


def test_beamformer_oracle_matches_reference_vectors(golden):
    """oracle/beamformer_path.py against the outputs of the reference's own module (classify/beamformer.py:41-55, float64)."""
    from oracle import beamformer_path as ob
    g = golden("beamformer.npz")
    got = ob.delay_and_sum(g["x"], g["delays"], int(g["kernel_size"]))
    well = np.ones(g["out"].shape[1], bool)
    well[400:420] = False                                    # delays beyond the window: ill-conditioned by construction
    assert np.abs(got - g["out"])[:, well].max() < 1e-11 * np.abs(g["out"][:, well]).max()
    assert np.abs(got - g["out"]).max() < 1e-7 * np.abs(g["out"]).max()
    assert np.abs(ob.delay_and_sum(g["x"], g["module_delays"], 41) - g["module_out"]).max() < 1e-11 * np.abs(g["module_out"]).max()
