"""Generate ``tests/golden/*.npz`` by RUNNING THE REFERENCE (import from /root/reference/src).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

The fixtures hold seeded synthetic inputs plus the outputs of the reference's own
functions, so the oracle restatement (``oracle/numpy_path.py``, ``oracle/torch_path.py``)
and the CUDA path can both be held to the real thing anywhere.  librosa / pyrubberband /
wfdb are not installed here; empty stand-ins let ``mpcg_wav2vec.augment`` import (none of
the functions exercised below touches them).
"""
from __future__ import annotations

import pathlib
import sys
import types

import numpy as np
import scipy
import torch
import torchaudio

REF_SRC = pathlib.Path("/root/reference/src")
OUT = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden"


def import_reference():
    if not REF_SRC.exists():
        raise SystemExit("reference not mounted; fixtures can only be regenerated in the build container")
    for name in ("librosa", "pyrubberband", "wfdb"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__spec__ = None
            if name == "wfdb":
                m.Record = object
            sys.modules[name] = m
    if str(REF_SRC) not in sys.path:
        sys.path.insert(0, str(REF_SRC))
    import mpcg_wav2vec.signalproc as sp            # noqa: F401
    from mpcg_wav2vec.signalproc import torchproc    # noqa: F401
    from mpcg_wav2vec.augment import torchaug        # noqa: F401
    return sp, torchproc, torchaug


# --------------------------------------------------------------------------- inputs
def synth_pcg(rows: int, n: int, fs: float, seed: int, spikes: int = 2) -> np.ndarray:
    """Heart-sound-like bursts + noise + a few large spikes (float32)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    out = np.zeros((rows, n))
    for r in range(rows):
        bpm = rng.uniform(60, 100)
        period = 60.0 / bpm
        beat = np.arange(0.1, t[-1], period)
        f1, f2 = rng.uniform(40, 120, 2)
        sig = np.zeros(n)
        for b0 in beat:
            sig += 1.0 * np.exp(-0.5 * ((t - b0) / 0.020) ** 2) * np.sin(2 * np.pi * f1 * (t - b0))
            sig += 0.6 * np.exp(-0.5 * ((t - b0 - 0.35 * period) / 0.015) ** 2) * np.sin(2 * np.pi * f2 * (t - b0))
        sig += 0.05 * rng.standard_normal(n)
        for _ in range(spikes):
            at = int(rng.integers(50, n - 50))
            width = int(rng.integers(3, 11))
            sig[at:at + width] += rng.choice([-1, 1]) * rng.uniform(5, 20)
        out[r] = sig * rng.uniform(0.1, 2.0)
    return out.astype(np.float32)


def synth_ecg(rows: int, n: int, fs: float, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    out = np.zeros((rows, n))
    for r in range(rows):
        period = 60.0 / rng.uniform(60, 100)
        sig = 0.2 * np.sin(2 * np.pi * 0.3 * t) + 0.02 * rng.standard_normal(n)
        for b0 in np.arange(0.1, t[-1], period):
            sig += np.exp(-0.5 * ((t - b0) / 0.010) ** 2)
        out[r] = sig
    return out.astype(np.float32)


# --------------------------------------------------------------------------- RNG replay
def replay_noise(x, seed):
    np.random.seed(seed); torch.manual_seed(seed)
    std = float(np.random.choice((0.0001, 0.001, 0.01)))
    scale = torch.rand(x.shape[0], 1) * 0.1
    noise = torch.randn_like(x)
    return std, scale, noise


def replay_sines(rows, seed, span):
    torch.manual_seed(seed)
    amp, freq, phase = [], [], []
    for lo, hi in ((0.05, 0.5), (0.001, 0.05)):
        amp.append(0.01 + torch.rand(rows, 1) * span)
        freq.append(lo + torch.rand(rows, 1) * (hi - lo))
        phase.append(torch.rand(rows, 1))
    return torch.cat(amp, 1), torch.cat(freq, 1), torch.cat(phase, 1)


def replay_bands(seed, low, high, n=5):
    np.random.seed(seed)
    bands = []
    for _ in range(n):
        lo = float(np.random.uniform(low, 0.95 * high))
        hi = float(np.random.uniform(lo + 0.05 * (high - low), high))
        bands.append((lo, hi))
    return bands


def replay_chain(x, seed, cfg):
    """Draw order of ``augment_pcg_batch``: (noise draws, mask) (sines, mask) (bands, mask) (noise, mask)."""
    rows = x.shape[0]
    np.random.seed(seed); torch.manual_seed(seed)
    d = {}
    d["std1"] = float(np.random.choice((0.0001, 0.001, 0.01)))
    d["scale1"] = torch.rand(rows, 1) * 0.1
    d["noise1"] = torch.randn_like(x)
    d["mask1"] = (torch.rand(rows, 1) < cfg.prob_noise / 4).float()
    amp, freq, phase = [], [], []
    for lo, hi in ((0.05, 0.5), (0.001, 0.05)):
        amp.append(0.01 + torch.rand(rows, 1) * 0.24)
        freq.append(lo + torch.rand(rows, 1) * (hi - lo))
        phase.append(torch.rand(rows, 1))
    d["amp"], d["freq"], d["phase"] = torch.cat(amp, 1), torch.cat(freq, 1), torch.cat(phase, 1)
    d["mask2"] = (torch.rand(rows, 1) < cfg.prob_wandering_volume).float()
    bands = []
    for _ in range(5):
        lo = float(np.random.uniform(2, 0.95 * 500))
        hi = float(np.random.uniform(lo + 0.05 * (500 - 2), 500))
        bands.append((lo, hi))
    d["bands"] = bands
    d["mask3"] = (torch.rand(rows, 1) < cfg.prob_banding).float()
    d["std2"] = float(np.random.choice((0.0001, 0.001, 0.01)))
    d["scale2"] = torch.rand(rows, 1) * 0.1
    d["noise2"] = torch.randn_like(x)
    d["mask4"] = (torch.rand(rows, 1) < cfg.prob_noise / 4).float()
    return d


def main():
    sp, tp, ta = import_reference()
    from mpcg_wav2vec.augment import AugmentConfig
    from mpcg_wav2vec.signalproc.despike import remove_spikes as np_despike
    from mpcg_wav2vec.signalproc import filters as np_filters
    from mpcg_wav2vec.signalproc.segment import WindowSpec, window_starts
    OUT.mkdir(parents=True, exist_ok=True)
    versions = dict(numpy=np.__version__, scipy=scipy.__version__, torch=torch.__version__,
                    torchaudio=torchaudio.__version__)

    # ---- G1: preprocessing, both reference paths, 2 kHz -> 4125 Hz (config-2 shape, shortened to 3 s)
    fs_in, fs_out = 2000, 4125
    pcg = synth_pcg(2, 5000, fs_in, seed=11)
    ecg = synth_ecg(2, 5000, fs_in, seed=12)
    g = dict(pcg=pcg, ecg=ecg, fs_in=fs_in, fs_out=fs_out)
    # NumPy path (float64), one recording at a time
    g["np_resample"] = np.stack([sp.resample(r.astype(np.float64), fs_in, fs_out) for r in pcg])
    g["np_despike"] = np.stack([np_despike(r, fs_out) for r in g["np_resample"]])
    g["np_band"] = np.stack([np_filters.bandpass_cascade(r, fs_out, 25.0, 450.0) for r in g["np_despike"]])
    g["np_norm"] = np.stack([sp.abs_max_normalise(r) for r in g["np_band"]])
    g["np_pcg"] = np.stack([sp.preprocess_pcg(r, fs_in, fs_out) for r in pcg])
    g["np_ecg"] = np.stack([sp.preprocess_ecg(r, fs_in, fs_out) for r in ecg])
    spec = WindowSpec(window_s=1.0)
    g["np_windows"] = np.stack([sp.segment(r, fs_out, spec) for r in g["np_pcg"]])
    pair = np.stack([g["np_pcg"][0], g["np_ecg"][0]], axis=1)              # [T, 2] loader layout
    g["np_windows_tc"] = sp.segment(pair, fs_out, spec)
    # tensor path, float64 (the target) and float32 (the reference's usual dtype)
    for tag, dt in (("t64", torch.float64), ("t32", torch.float32)):
        xp = torch.from_numpy(pcg).to(dt)
        xe = torch.from_numpy(ecg).to(dt)
        rs = tp.resample(xp, fs_in, fs_out)
        ds = tp.remove_spikes(rs, fs_out)
        bd = tp.bandpass_cascade(ds, fs_out, 25.0, 450.0)
        g[f"{tag}_despike"] = ds.numpy()
        if tag == "t64":
            g[f"{tag}_resample"] = rs.numpy()
            g[f"{tag}_band"] = bd.numpy()
            g[f"{tag}_norm"] = tp.abs_max_normalise(bd).numpy()
        g[f"{tag}_pcg"] = tp.preprocess_pcg(xp, fs_in, fs_out).numpy()
        g[f"{tag}_ecg"] = tp.preprocess_ecg(xe, fs_in, fs_out).numpy()
        if tag == "t64":
            g[f"{tag}_windows"] = tp.segment(tp.preprocess_pcg(xp, fs_in, fs_out), fs_out, spec).contiguous().numpy()
    np.savez_compressed(OUT / "preprocess_2k_4125.npz", versions=str(versions), **g)

    # ---- G2: the other two resampling ratios (2 k -> 16 k, 4 k -> 4125) and a 16 kHz despike/band case
    g = {}
    x16 = synth_pcg(2, 2000, 2000, seed=21)
    g["x_2k"] = x16
    g["np_2k_16k"] = np.stack([sp.resample(r.astype(np.float64), 2000, 16000) for r in x16])
    g["t64_2k_16k"] = tp.resample(torch.from_numpy(x16).double(), 2000, 16000).numpy()
    g["t64_pcg_2k_16k"] = tp.preprocess_pcg(torch.from_numpy(x16).double(), 2000, 16000).numpy()
    g["np_pcg_2k_16k"] = np.stack([sp.preprocess_pcg(r, 2000, 16000) for r in x16])
    x4 = synth_pcg(2, 4000, 4000, seed=22)
    g["x_4k"] = x4
    g["np_4k_4125"] = np.stack([sp.resample(r.astype(np.float64), 4000, 4125) for r in x4])
    g["t64_4k_4125"] = tp.resample(torch.from_numpy(x4).double(), 4000, 4125).numpy()
    g["t64_ecg_16k"] = tp.preprocess_ecg(torch.from_numpy(x16).double(), 2000, 16000).numpy()
    np.savez_compressed(OUT / "resample_ratios.npz", versions=str(versions), **g)

    # ---- G3: segmentation integer decisions over a sweep of lengths / rates / specs
    rows = []
    for fs in (1000, 4000, 4125, 16000):
        for ws in (2.0, 4.0):
            sp_ = WindowSpec(window_s=ws)
            for n in (0, 100, int(0.3 * fs), int(0.3 * fs) + 1, int(ws * fs), int(ws * fs) + int(0.3 * fs),
                      int(10.3 * fs), int(30 * fs), 33000, 123750, 480000):
                st = window_starts(n, fs, sp_)
                rows.append((fs, ws, n, sp_.window_len(fs), sp_.hop_len(fs), int(round(0.3 * fs)),
                             len(st), st[0] if st else -1, st[-1] if st else -1))
    np.savez_compressed(OUT / "segment_index.npz", table=np.array(rows, dtype=np.float64),
                        columns="fs,window_s,n,win,hop,start,count,first,last")

    # ---- G4: augmentation with replayed draws (float32, the dtype torchaug actually runs in)
    torch.manual_seed(5)
    xw = ta._normalise(torch.from_numpy(synth_pcg(4, 4125, 4125, seed=31, spikes=0)))
    g = dict(x=xw.numpy(), fs=4125)
    np.random.seed(101); torch.manual_seed(101)
    g["noise_out"] = ta.add_white_noise(xw).numpy()
    std, scale, noise = replay_noise(xw, 101)
    g.update(noise_std=std, noise_scale=scale.numpy(), noise_noise=noise.numpy())
    torch.manual_seed(102)
    g["sine_out"] = ta.sinusoidal_envelope(xw, 4125).numpy()
    a, f, p = replay_sines(4, 102, 0.24)
    g.update(sine_amp=a.numpy(), sine_freq=f.numpy(), sine_phase=p.numpy())
    torch.manual_seed(103)
    g["wander_out"] = ta.baseline_wander(xw, 4125).numpy()
    a, f, p = replay_sines(4, 103, 0.19)
    g.update(wander_amp=a.numpy(), wander_freq=f.numpy(), wander_phase=p.numpy())
    torch.manual_seed(104)
    g["warp_out"] = ta.amplitude_warp(xw).numpy()
    torch.manual_seed(104)
    g["warp_amps"] = (0.7 + torch.rand(4, 12) * 0.6).numpy()
    np.random.seed(105)
    g["eq_out"] = ta.parametric_eq(xw, 4125, 2, 500).numpy()
    g["eq_out64"] = None
    np.random.seed(105)
    g["eq_out64"] = ta.parametric_eq(xw.double(), 4125, 2, 500).numpy()
    g["eq_bands"] = np.array(replay_bands(105, 2, 500))
    cfg = AugmentConfig(prob_noise=2.0, prob_wandering_volume=0.75, prob_banding=0.6)
    np.random.seed(106); torch.manual_seed(106)
    g["chain_out"] = ta.augment_pcg_batch(xw, 4125, cfg).numpy()
    d = replay_chain(xw, 106, cfg)
    for k, v in d.items():
        g["chain_" + k] = np.array(v) if not torch.is_tensor(v) else v.numpy()
    np.savez_compressed(OUT / "torchaug_replay.npz", versions=str(versions), **g)

    # ---- G5: mel conditioning (DiffWave preset at 4 kHz, the 16 kHz bench preset, WaveGrad preset)
    g = {}
    sig = torch.from_numpy(synth_pcg(2, 8192, 4000, seed=41, spikes=0))
    sig = sig / sig.abs().amax(dim=-1, keepdim=True)
    g["x"] = sig.numpy()
    for tag, kw in (("dw4k", dict(sample_rate=4000, n_fft=1024, hop_length=256, n_mels=80, f_max=500.0)),
                    ("c4_16k", dict(sample_rate=16000, n_fft=1024, hop_length=256, n_mels=80, f_max=500.0)),
                    ("wg4k", dict(sample_rate=4000, n_fft=2048, win_length=1200, hop_length=300, n_mels=128,
                                  f_max=500.0))):
        mc = sp.MelConfig(**kw)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            tr = mc.build()
            g[f"{tag}_mel"] = tr(sig).numpy()
            g[f"{tag}_logmel"] = sp.log_mel(sig, tr).numpy()
            g[f"{tag}_logmel64"] = sp.log_mel(sig.double(), tr.double()).numpy()
    np.savez_compressed(OUT / "mel_presets.npz", versions=str(versions), **g)
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)

    gen_condition_fixture(sp, versions)
    zerophase_fixture(versions)
    normalisers_fixture(sp, versions)
    heart_cycles_fixture(sp, versions)


def normalisers_fixture(sp, versions):
    """G8: min-max / z-score / k-peak normalisers and the envelopes through the reference's
    signalproc/normalize.py:33-78 and signalproc/envelopes.py:11-23."""
    rng = np.random.default_rng(81)
    t = 3000
    x = (np.sin(np.arange(t) / 7.0)[None] * rng.uniform(0.2, 3, (3, 1)) + 0.2 * rng.standard_normal((3, t)) - 0.4)
    x = x.astype(np.float32)
    x[1, 100:104] = x[1].max()                                  # ties at the top
    x[2, 5] = 40.0                                              # an isolated spike
    flat = np.full((1, 64), 0.25, dtype=np.float32)
    xt = torch.from_numpy(x)
    g = dict(x=x, flat=flat)
    g["minmax"] = np.stack([sp.minmax_normalise(r) for r in x])
    g["minmax_02"] = np.stack([sp.minmax_normalise(r, 0.0, 2.0) for r in x])
    g["minmax_flat"] = sp.minmax_normalise(flat[0])[None]
    g["z"] = np.stack([sp.z_normalise(r) for r in x])
    g["kpeak3"] = np.stack([sp.kpeak_normalise(r) for r in x])
    g["kpeak40"] = np.stack([sp.kpeak_normalise(r, k=40, lo=0.0, hi=1.0) for r in x])
    g["kpeak_flat"] = sp.kpeak_normalise(flat[0])[None]
    g["minmax_torch_rows"] = np.stack([sp.minmax_normalise_torch(r).numpy() for r in xt])
    g["minmax_torch_all"] = sp.minmax_normalise_torch(xt).numpy()
    g["z_torch"] = sp.z_normalise_torch(xt.double()[None]).numpy()[0]
    g["kpeak_torch_rows"] = np.stack([sp.kpeak_normalise_torch(r.double()).numpy() for r in xt])
    g["kpeak_torch_all"] = sp.kpeak_normalise_torch(xt.double()).numpy()
    g["kpeak_torch_all_k5"] = sp.kpeak_normalise_torch(xt.double(), k=5, lo=0.0, hi=3.0).numpy()
    fs = 1000.0
    e = (np.sin(2 * np.pi * 40.0 * np.arange(2500) / fs) * (1 + 0.5 * np.sin(2 * np.pi * 1.5 * np.arange(2500) / fs)))[None]
    e = (np.concatenate([e, e[:, ::-1] * 0.3]) + 0.05 * rng.standard_normal((2, 2500))).astype(np.float32)
    g["env_x"], g["env_fs"] = e, fs
    g["hilbert"] = np.stack([sp.hilbert_envelope(r) for r in e])
    g["hilbert_odd"] = np.stack([sp.hilbert_envelope(r[:2499]) for r in e])
    g["homomorphic"] = np.stack([sp.homomorphic_envelope(r, fs) for r in e])
    np.savez_compressed(OUT / "normalisers.npz", versions=str(versions), **g)


def heart_cycles_fixture(sp, versions):
    """G9: cardiac-cycle rearrangement through the reference's datasets/heart_cycles.py and the
    rebuilt branch of GenerativeDataset.__getitem__ (datasets/generative.py:62-91)."""
    import random
    from mpcg_wav2vec.datasets import heart_cycles as hc
    from mpcg_wav2vec.datasets.generative import _fade
    from mpcg_wav2vec.signalproc.preprocess import fit_length
    rng = np.random.default_rng(91)
    fs, t, crop, fade_n = 4000, 16000, 24 * 256, 40
    n = np.arange(t)
    beat = np.exp(-(((n % 3100) - 400.0) / 90.0) ** 2) * np.sin(2 * np.pi * 55 * n / fs)
    x = (beat[None] * rng.uniform(0.5, 2.0, (3, 1)) + 0.02 * rng.standard_normal((3, t)) + 0.1).astype(np.float32)
    x[2, 6000:9400] = 0.1                                           # a flat stretch: the linear-ramp branch
    joins = [0, 350, 3420, 6555, 9640, 9660, 12800, 15900, 16000, 17000]   # one 20-sample cycle (shorter than the fade)
    g = dict(fs=fs, crop=crop, fade_n=fade_n, x=x, joins=np.array(joins))
    orders = []
    for seed in (0, 1, 2, 3):
        for pc in (0.0, 1.0):
            orders.append(hc.rearrange({"a": list(range(6))}, prob_contiguous=pc, rng=random.Random(seed))["a"])
    g["orders6"] = np.array(orders)
    g["orders_seed_pc"] = np.array([(s, pc) for s in (0, 1, 2, 3) for pc in (0.0, 1.0)])
    for r in range(3):
        sig = sp.abs_max_normalise(x[r])
        cyc = hc.split_cycles(sig, joins)
        arranged = hc.rearrange({"ref": cyc}, prob_contiguous=0.0, rng=random.Random(10 + r))["ref"]
        order = hc.rearrange({"ref": list(range(len(cyc)))}, prob_contiguous=0.0, rng=random.Random(10 + r))["ref"]
        out = hc.rebuild(arranged, crop, fade_n)
        g[f"order_{r}"] = np.array(order)
        g[f"rebuilt_{r}"] = out
        g[f"item_{r}"] = fit_length(_fade(out), crop)[0]
    g["rebuilt_short_target"] = hc.rebuild(hc.split_cycles(sp.abs_max_normalise(x[0]), joins), 100, fade_n)
    g["rebuilt_long_target"] = hc.rebuild(hc.split_cycles(sp.abs_max_normalise(x[0]), joins)[:2], 200000, fade_n)
    np.savez_compressed(OUT / "heart_cycles.npz", versions=str(versions), **g)


def zerophase_fixture(versions):
    """G7: zero-phase filters through the reference's signalproc/filters.py:44-90."""
    from mpcg_wav2vec.signalproc import filters as F
    rng = np.random.default_rng(71)
    fs, t = 4125.0, 4000
    x = (np.sin(2 * np.pi * 50.0 * np.arange(t) / fs)[None] * 0.5 + np.sin(2 * np.pi * 180.0 * np.arange(t) / fs)[None]
         + 0.3 * rng.standard_normal((2, t)) + 0.7).astype(np.float32)
    g = dict(fs=fs, x=x)
    g["bandpass"] = np.stack([F.butter_bandpass(r, fs, 25.0, 400.0) for r in x])
    g["lowpass"] = np.stack([F.butter_lowpass(r, fs, 150.0) for r in x])
    g["highpass"] = np.stack([F.butter_highpass(r, fs, 20.0) for r in x])
    g["band_stop"] = np.stack([F.band_stop(r, fs, 45.0, 55.0) for r in x])
    g["notch"] = np.stack([F.notch(r, fs, 50.0) for r in x])
    g["notch_chain"] = np.stack([F.notch_chain(r, fs, (50.0, 100.0, 150.0, 3000.0)) for r in x])
    g["bands"] = np.stack([F.decompose_bands(r, fs) for r in x[:1]])
    np.savez_compressed(OUT / "zerophase.npz", versions=str(versions), **g)


def gen_condition_fixture(sp, versions):
    """G6: generator conditioning: the waveform part of GenerativeDataset.__getitem__ run through
    the reference's own functions (datasets/generative.py:36-42,88-104,112; signalproc/preprocess.py:45-64)."""
    from mpcg_wav2vec.datasets.generative import _fade
    from mpcg_wav2vec.signalproc.preprocess import fit_length
    rng = np.random.default_rng(61)
    fs, crop = 4000, 24 * 256
    g = dict(fs=fs, crop=crop)
    for tag, t in (("long", 8000), ("exact", crop), ("short", 3000), ("tiny", 200)):
        x = (np.sin(np.arange(t) / 9.0)[None] * rng.uniform(0.2, 3, (2, 1)) + 0.1 * rng.standard_normal((2, t)) + 0.3)
        x = x.astype(np.float32)
        y = np.stack([fit_length(_fade(sp.abs_max_normalise(r)), crop)[0] for r in x])
        g[f"{tag}_x"] = x
        g[f"{tag}_y"] = y
        g[f"{tag}_chirp"] = np.stack([sp.add_chirp(r, fs) for r in y])
    np.savez_compressed(OUT / "gen_condition.npz", versions=str(versions), **g)


def beamformer_fixture():
    """G7: the sinc delay-and-sum gather with its gradients, from the reference's own module
    (classify/beamformer.py:41-55) in float64.  The file is loaded on its own: the package __init__ pulls in
    transformers, which does not import under the librosa stub."""
    import importlib.util
    import torch
    spec = importlib.util.spec_from_file_location("ref_beamformer", REF_SRC / "mpcg_wav2vec" / "classify" / "beamformer.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(7)
    bf = mod.TimeVaryingSincBeamformer(num_mics=3, fs=4125.0).double().eval()
    b, m, t = 2, 3, 600
    x = torch.randn(b, m, t, dtype=torch.float64, requires_grad=True)
    # delays up to 12 samples keep the sinc's main lobe inside the 41-tap window; beyond ~20 the taps that remain sum to
    # almost nothing and the reference's own normalisation (kernel / kernel.sum()) amplifies rounding by orders of
    # magnitude -- the module allows it (clamp at 41.25), so one stretch of such delays is kept for a looser check
    delays = torch.rand(b, m, t, dtype=torch.float64) * 12.0
    delays[1, 1, 400:420] = torch.rand(20, dtype=torch.float64) * 20.0 + 21.0
    delays[0, 0, :50] = 0.0                                  # the clamp's lower end
    delays[1, 2, 100:110] = 3.0                              # an integer delay: the sinc(0) tap
    delays.requires_grad_(True)
    out = torch.stack([bf._delay_channel(x[:, i, :], delays[:, i, :]) ** 2 for i in range(m)], dim=1).sum(dim=1)
    g = torch.randn(b, t, dtype=torch.float64)
    out.backward(g)
    with torch.no_grad():
        full = bf(x.detach())                                # the whole module: predictor + clamp + gather
        own_delays = torch.clamp(bf.delay_predictor(x.detach()), 0.0, bf.max_delay_samples)
    state = {k: v.numpy() for k, v in bf.state_dict().items()}
    np.savez_compressed(OUT / "beamformer.npz", x=x.detach().numpy(), delays=delays.detach().numpy(), out=out.detach().numpy(),
                        grad_out=g.numpy(), grad_x=x.grad.numpy(), grad_delays=delays.grad.numpy(), kernel_size=41,
                        module_out=full.numpy(), module_delays=own_delays.numpy(), state_keys=np.array(list(state)),
                        **{"state__" + k: v for k, v in state.items()}, versions=f"torch {torch.__version__}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "beamformer":
        beamformer_fixture()
    else:
        main()
        beamformer_fixture()
