"""TEST INFRASTRUCTURE -- CPU restatement (NumPy float64, one signal at a time) of the reference's composed augmentation
pipelines with every random quantity injected:

* ``augment/pipelines.py:43-61``   -> :func:`augment_pcg`
* ``augment/pipelines.py:64-80``   -> :func:`augment_ecg`
* ``augment/pipelines.py:82-125``  -> :func:`augment_pcg_ecg`
* ``augment/pipelines.py:127-148`` -> :func:`augment_multi_pcg`
* ``augment/primitives.py:44-83``  -> :func:`add_white_noise`, :func:`sinusoidal_envelope`, :func:`baseline_wander`,
  :func:`parametric_eq` (the NumPy versions: each normalises its own result)
* ``augment/noise_sources.py:33-64`` -> :func:`recorded_noise` (the arithmetic after the file reads)

Pinned stages: the normalisers, noise / sine / EQ arithmetic (SciPy ``iirfilter`` / ``sosfilt`` are called, as the
reference calls them).  PARITY UNPINNED for three stages, as in DESIGN.md: HPSS (``oracle/hpss_path.py``; librosa absent),
the time stretch (rubberband absent: :func:`time_warp` restates THIS build's defined Catmull-Rom warp) and the noise
records' file reads (wfdb absent).  Only ``tests/`` may import this module.
"""
from __future__ import annotations

import numpy as np
from scipy import signal as sp

from . import hpss_path as oh
from . import numpy_path as onp

N = onp.abs_max_normalise


def add_white_noise(x, sigma, noise):
    return N(x + sigma * noise[: len(x)])


def _sines(n, fs, p):
    t = np.arange(n) / fs
    return p[0] * np.sin(2 * np.pi * (p[1] * t + p[2])) + p[3] * np.sin(2 * np.pi * (p[4] * t + p[5]))


def sinusoidal_envelope(x, fs, p):
    return N(x * (1.0 + _sines(x.size, fs, p)))


def baseline_wander(x, fs, p):
    return N(x + _sines(x.size, fs, p))


def parametric_eq(x, fs, bands):
    nyq = fs / 2.0
    coloured = np.asarray(x, dtype=np.float64)
    for lo, hi in bands:
        sos = sp.iirfilter(1, [lo / nyq, hi / nyq], btype="band", ftype="butter", output="sos")
        coloured = sp.sosfilt(sos, coloured)
    return N(N(coloured) / 50.0 + N(x))


def time_warp(x, rate, keep_length=False):
    """This build's stand-in for ``primitives.time_stretch``: y[j] = x(j * rate), Catmull-Rom, edges clamped."""
    x = np.asarray(x, dtype=np.float64)
    t = x.shape[0]
    n = int(round(t / float(rate)))
    pos = np.arange(n) * float(rate)
    i = np.floor(pos).astype(int)
    u = (pos - i).astype(np.float32).astype(np.float64)
    at = lambda k: x[np.clip(k, 0, t - 1)]
    p0, p1, p2, p3 = at(i - 1), at(i), at(i + 1), at(i + 2)
    y = (((-0.5 * p0 + 1.5 * p1 - 1.5 * p2 + 0.5 * p3) * u + (p0 - 2.5 * p1 + 2 * p2 - 0.5 * p3)) * u + (-0.5 * p0 + 0.5 * p2)) * u + p1
    return y[:t] if keep_length else y


def recorded_noise(bank, rows, starts, scale, length, normalise_sum):
    total = np.zeros(length)
    for r, s, k in zip(rows, starts, scale):
        if k != 0:
            total = total + float(k) * N(np.asarray(bank[r][s:s + length], dtype=np.float64))
    if normalise_sum and np.max(np.abs(total)) > 0:
        total = N(total)
    return total


def augment_pcg(x, fs, d, bank=None):
    """``d``: this signal's draws (masks as bools, per-call parameters as in the device call)."""
    x = onp.minmax_normalise(np.asarray(x, dtype=np.float64))
    if d["mask_hpss"]:
        x, _ = oh.hpss_recombine(x, d["hpss"], False)
    if d["mask_noise1"]:
        x = add_white_noise(x, d["noise1"]["sigma"], d["noise1"]["noise"])
    if d["mask_warp"]:
        x = N(time_warp(x, d["rate"]))
    if d["mask_volume"]:
        x = sinusoidal_envelope(x, fs, d["volume"])
    if d["mask_noise2"]:
        x = add_white_noise(x, d["noise2"]["sigma"], d["noise2"]["noise"])
    if d["mask_eq"]:
        x = parametric_eq(x, fs, d["bands"])
    if bank is not None and d["mask_real"]:
        r = d["real"]
        x = x + recorded_noise(bank["records"], r["rows"], r["starts"], r["scale"], len(x), bank["normalise_sum"])
    return N(x)


def augment_ecg(x, fs, d, bank=None):
    x = onp.minmax_normalise(np.asarray(x, dtype=np.float64))
    if d["mask_noise1"]:
        x = add_white_noise(x, d["noise1"]["sigma"], d["noise1"]["noise"])
    if d["mask_wander"]:
        x = baseline_wander(x, fs, d["wander"])
    if d["mask_warp"]:
        x = N(time_warp(x, d["rate"]))
    if d["mask_noise2"]:
        x = add_white_noise(x, d["noise2"]["sigma"], d["noise2"]["noise"])
    if d["mask_eq"]:
        x = parametric_eq(x, fs, d["bands"])
    if bank is not None and d["mask_real"]:
        r = d["real"]
        x = x + recorded_noise(bank["records"], r["rows"], r["starts"], r["scale"], len(x), bank["normalise_sum"])
    return N(x)


def augment_pcg_ecg(e, p, fs, d):
    e = onp.minmax_normalise(np.asarray(e, dtype=np.float64))
    p = onp.minmax_normalise(np.asarray(p, dtype=np.float64))
    if d["mask_hpss"]:
        p, n = oh.hpss_recombine(p, d["hpss"], True)
        e = e[:n]
    if d["mask_noise1_p"]:
        p = add_white_noise(p, d["noise1_p"]["sigma"], d["noise1_p"]["noise"])
    if d["mask_noise1_e"]:
        e = add_white_noise(e, d["noise1_e"]["sigma"], d["noise1_e"]["noise"])
    if d["mask_wander"]:
        e = baseline_wander(e, fs, d["wander"])
    if d["mask_warp"]:
        e = N(time_warp(e, d["rate"]))
        p = N(time_warp(p, d["rate"]))
    if d["mask_volume"]:
        p = sinusoidal_envelope(p, fs, d["volume"])
    if d["mask_noise2_p"]:
        p = add_white_noise(p, d["noise2_p"]["sigma"], d["noise2_p"]["noise"])
    if d["mask_noise2_e"]:
        e = add_white_noise(e, d["noise2_e"]["sigma"], d["noise2_e"]["noise"])
    if d["mask_eq_p"]:
        p = parametric_eq(p, fs, d["bands_p"])
    if d["mask_eq_e"]:
        e = parametric_eq(e, fs, d["bands_e"])
    return N(e), N(p)


def augment_multi_pcg(chans, fs, d, bank=None):
    """``chans``: list of channel arrays of one recording; ``d["noise1"]`` / ``d["noise2"]``: per-channel lists."""
    chans = [N(np.asarray(c, dtype=np.float64)) for c in chans]
    if d["mask_noise1"]:
        chans = [add_white_noise(c, q["sigma"], q["noise"]) for c, q in zip(chans, d["noise1"])]
    if d["mask_warp"]:
        chans = [N(time_warp(c, d["rate"], keep_length=True)) for c in chans]
    if d["mask_volume"]:
        mod = _sines(chans[0].size, fs, d["volume"])
        chans = [N(c * (1.0 + mod)) for c in chans]
    if d["mask_noise2"]:
        chans = [add_white_noise(c, q["sigma"], q["noise"]) for c, q in zip(chans, d["noise2"])]
    if bank is not None and d["mask_real"]:
        r = d["real"]
        shared = recorded_noise(bank["records"], r["rows"], r["starts"], r["scale"], len(chans[0]), bank["normalise_sum"])
        chans = [N(c + shared) for c in chans]
    return chans
