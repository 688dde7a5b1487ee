"""CPU oracle, tensor leg -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates the reference's batched tensor chain (``signalproc/torchproc.py``), its on-device
augmentation subset (``augment/torchaug.py``) and its mel conditioning
(``signalproc/spectrogram.py``) on CPU tensors.  torchaudio (installed on the GPU box as
well) supplies the third-party arithmetic the reference itself calls: ``resample``,
``lfilter``, ``MelSpectrogram``.  Run it in float64 to obtain the target the CUDA path is
held to (the reference's own float32 run is the noisier party: SURVEY.md section 8c).

Stochastic functions take their random draws as arguments ("injected-parameter mode").
``oracle/make_golden.py`` replays the reference's RNG call order to prove that, given the
same draws, these functions reproduce the reference bit for bit.

reference file:line -> function here

* ``torchproc.py:32-53``   -> :func:`butter_ba`, :func:`lowpass`, :func:`highpass`, :func:`bandpass_cascade`
* ``torchproc.py:56-59``   -> :func:`resample`
* ``torchproc.py:62-66``   -> :func:`abs_max_normalise`;  ``torchaug.py:24-27`` -> :func:`renormalise`
* ``torchproc.py:69-98``   -> :func:`remove_spikes` (lower median, per-row worst frame)
* ``torchproc.py:101-116`` -> :func:`preprocess_pcg`, :func:`preprocess_ecg`
* ``torchproc.py:119-129`` -> :func:`segment`
* ``torchaug.py:39-100``   -> :func:`add_white_noise`, :func:`sinusoidal_envelope`,
  :func:`baseline_wander`, :func:`amplitude_warp`, :func:`parametric_eq`
* ``torchaug.py:30-36,103-111`` -> :func:`blend`, :func:`augment_pcg_batch`
* ``spectrogram.py:13-45`` -> :func:`mel_transform`, :func:`log_mel`
* ``signalproc/normalize.py:41-44,52-56,75-78`` -> :func:`minmax_normalise_torch`, :func:`z_normalise_torch`,
  :func:`kpeak_normalise_torch`
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
import torchaudio
import torchaudio.functional as AF
from scipy import signal as _sig

PCG_BAND = (25.0, 450.0)
ECG_BAND = (2.0, 40.0)
SPIKE_FILL = 1e-4


def _rows(x: torch.Tensor):
    return (x[None], True) if x.dim() == 1 else (x, False)


# --------------------------------------------------------------------------- filters
def butter_ba(cutoff: float, fs: float, kind: str, order: int):
    """SciPy design with the cut-off divided by fs (not Nyquist): the reference convention."""
    return _sig.butter(order, cutoff / fs, btype=kind)


def _lfilter(x, b, a):
    bt = torch.as_tensor(b, dtype=x.dtype, device=x.device)
    at = torch.as_tensor(a, dtype=x.dtype, device=x.device)
    return AF.lfilter(x, at, bt, clamp=False, batching=True)


def lowpass(x, fs, cutoff, order=2):
    return _lfilter(x, *butter_ba(cutoff, fs, "lowpass", order))


def highpass(x, fs, cutoff, order=2):
    return _lfilter(x, *butter_ba(cutoff, fs, "highpass", order))


def bandpass_cascade(x, fs, low, high, order=2):
    return highpass(lowpass(x, fs, high, order), fs, low, order)


# --------------------------------------------------------------------------- resample
def resample(x, fs_in, fs_out):
    if fs_in == fs_out:
        return x
    return AF.resample(x, int(round(fs_in)), int(round(fs_out)))


# --------------------------------------------------------------------------- normalise
def renormalise(x):
    """Row-wise: subtract mean, divide by max(|.|) floored at 1e-12, clamp to [-1, 1]."""
    x = x - x.mean(dim=-1, keepdim=True)
    top = x.abs().amax(dim=-1, keepdim=True).clamp_min(1e-12)
    return (x / top).clamp(-1.0, 1.0)


def abs_max_normalise(x):
    return renormalise(torch.nan_to_num(x))


# --------------------------------------------------------------------------- despike
def remove_spikes(x, fs, threshold=3.0, max_iterations=1000, trace=None):
    """Batched Schmidt despike.  Every row runs its own sequence of edits: in one sweep of
    the reference's outer loop, each still-active row flattens one span in its worst frame.
    Rows are independent, so this restatement simply loops rows on the outside.

    ``trace`` (optional list) receives ``(row, frame, peak, lo, hi)`` per edit.
    """
    x, squeeze = _rows(x)
    out = x.clone()
    rows, n = out.shape
    win = round(float(fs) / 2.0)
    if win < 1 or n < win:
        return out[0] if squeeze else out
    covered = n - n % win
    for r in range(rows):
        frames = out[r, :covered].view(-1, win)      # shares storage with ``out``
        for _ in range(max_iterations):
            tops = frames.abs().amax(dim=1)
            mid = tops.median()                      # lower middle value for an even count
            if not bool((tops > threshold * mid).any()):
                break
            w = int(tops.argmax())
            fr = frames[w]
            peak = int(fr.abs().argmax())
            sg = torch.sign(fr)
            flips = torch.nonzero((sg[1:] - sg[:-1]).abs() > 1).flatten()
            left = flips[flips < peak]
            right = flips[flips >= peak]
            lo = int(left[-1]) + 1 if left.numel() else 0
            hi = int(right[0]) if right.numel() else win - 1
            if trace is not None:
                trace.append((r, w, peak, lo, hi))
            fr[lo:hi] = SPIKE_FILL
    return out[0] if squeeze else out


# --------------------------------------------------------------------------- chains
def preprocess_pcg(x, fs_in, fs_out, *, despike=True):
    x, squeeze = _rows(x)
    x = resample(x, fs_in, fs_out)
    if despike:
        x = remove_spikes(x, fs_out)
    x = abs_max_normalise(bandpass_cascade(x, fs_out, *PCG_BAND, order=2))
    return x[0] if squeeze else x


def preprocess_ecg(x, fs_in, fs_out):
    x, squeeze = _rows(x)
    x = resample(x, fs_in, fs_out)
    x = abs_max_normalise(bandpass_cascade(x, fs_out, *ECG_BAND, order=2))
    return x[0] if squeeze else x


def segment(x, fs, spec):
    """``[..., T]`` -> ``[..., N, win]``: drop the start pad, zero-pad to one window if short."""
    x, squeeze = _rows(x)
    win, hop = spec.window_len(fs), spec.hop_len(fs)
    x = x[..., int(round(spec.start_pad_s * fs)):]
    if x.shape[-1] < win:
        x = F.pad(x, (0, win - x.shape[-1]))
    out = x.unfold(-1, win, hop)
    return out[0] if squeeze else out


# --------------------------------------------------------------------------- augmentation
NOISE_STDS = (0.0001, 0.001, 0.01)
SINE_BANDS = ((0.05, 0.5), (0.001, 0.05))


def add_white_noise(x, std: float, scale, noise):
    """``scale`` [B,1] already holds U(0,1)*0.1; ``noise`` [B,T] is standard normal."""
    return x + scale * std * noise


def _sine_sum(n, fs, amp, freq, phase):
    """amp/freq/phase: [B, 2] (one column per band).  float32 accumulator like the reference."""
    t = torch.arange(n) / fs
    acc = torch.zeros(amp.shape[0], n)
    for k in range(amp.shape[1]):
        acc = acc + amp[:, k:k + 1] * torch.sin(2 * np.pi * (freq[:, k:k + 1] * t + phase[:, k:k + 1]))
    return acc


def sinusoidal_envelope(x, fs, amp, freq, phase):
    return x * (1.0 + _sine_sum(x.shape[-1], fs, amp, freq, phase))


def baseline_wander(x, fs, amp, freq, phase):
    return x + _sine_sum(x.shape[-1], fs, amp, freq, phase)


def warp_curve(amps: torch.Tensor, kernel: int = 65) -> torch.Tensor:
    """[B, P] control gains -> [B, kernel] unit-sum smoothing taps (piecewise-linear)."""
    p = amps.shape[1]
    pos = torch.clamp(torch.arange(kernel).float() / (kernel - 1) * (p - 1), max=p - 1)
    lo, hi = pos.floor().long(), pos.ceil().long()
    curve = amps[:, lo] + (amps[:, hi] - amps[:, lo]) * (pos - lo)[None]
    return curve / curve.sum(dim=-1, keepdim=True)


def amplitude_warp(x, amps, kernel: int = 65):
    """Depthwise correlation of every row with its own ``warp_curve`` taps, reflect-padded."""
    b, t = x.shape
    taps = warp_curve(amps, kernel).to(x.dtype)[:, None]
    padded = F.pad(x[:, None], (kernel // 2, kernel // 2), mode="reflect")
    return F.conv1d(padded.reshape(1, b, -1), taps, groups=b).reshape(b, -1)[:, :t]


def eq_sections(fs, bands):
    """[(lo, hi)] Hz -> list of (b, a) first-order Butterworth band-pass designs (Nyquist-normalised)."""
    nyq = fs / 2.0
    return [_sig.butter(1, [lo / nyq, hi / nyq], btype="band") for lo, hi in bands]


def parametric_eq(x, fs, bands):
    col = x
    for b, a in eq_sections(fs, bands):
        col = _lfilter(col, b, a)
    return renormalise(renormalise(col) / 50.0 + renormalise(x))


def blend(x, transformed, mask):
    """mask: [B,1] of 0./1.; every row is renormalised whether or not it was picked."""
    return renormalise(mask * transformed + (1.0 - mask) * x)


def augment_pcg_batch(x, fs, draws: dict):
    """Noise -> wandering volume -> EQ -> noise, each behind a row mask.

    ``draws`` keys: std1, scale1, noise1, mask1, amp, freq, phase, mask2, bands, mask3,
    std2, scale2, noise2, mask4.
    """
    x = renormalise(x)
    x = blend(x, add_white_noise(x, draws["std1"], draws["scale1"], draws["noise1"]), draws["mask1"])
    x = blend(x, sinusoidal_envelope(x, fs, draws["amp"], draws["freq"], draws["phase"]), draws["mask2"])
    x = blend(x, parametric_eq(x, fs, draws["bands"]), draws["mask3"])
    x = blend(x, add_white_noise(x, draws["std2"], draws["scale2"], draws["noise2"]), draws["mask4"])
    return x


# --------------------------------------------------------------------------- mel
def mel_transform(sample_rate, n_fft, hop_length, win_length=None, n_mels=80, f_min=0.125, f_max=500.0):
    return torchaudio.transforms.MelSpectrogram(
        sample_rate=sample_rate, n_fft=n_fft, win_length=win_length or n_fft, hop_length=hop_length,
        f_min=f_min, f_max=f_max, n_mels=n_mels, power=1.0, normalized=True)


def log_mel(x, transform):
    m = transform(x)
    m = 20.0 * torch.log10(torch.clamp(m, min=1e-5)) - 20.0
    return torch.clamp((m + 100.0) / 100.0, 0.0, 1.0)


# --------------------------------------------------------------------------- other normalisers
RANGE_EPS = 1e-8


def minmax_normalise_torch(x, lo=-1.0, hi=1.0):
    """``signalproc/normalize.py:41-44``: one range over the whole tensor, epsilon in the denominator."""
    bottom = x.min()
    return (x - bottom) / (x.max() - bottom + RANGE_EPS) * (hi - lo) + lo


def z_normalise_torch(x):
    """``signalproc/normalize.py:52-56``: last-dimension z-score, population standard deviation."""
    centre = x.mean(dim=-1, keepdim=True)
    return (x - centre) / (x.std(dim=-1, unbiased=False, keepdim=True) + RANGE_EPS)


def kpeak_normalise_torch(x, k=26, lo=-1.0, hi=1.0, dim=-1):
    """``signalproc/normalize.py:75-78``: top-k along ``dim``, then one mean over everything selected."""
    top = torch.topk(x, k=k, dim=dim).values.mean()
    bottom = -torch.topk(-x, k=k, dim=dim).values.mean()
    return lo + (x - bottom) / (top - bottom + RANGE_EPS) * (hi - lo)
