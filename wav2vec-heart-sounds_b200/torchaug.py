"""Drop-in for ``mpcg_wav2vec.augment.torchaug`` on B200 (reference ``augment/torchaug.py:24-111``): same names,
argument order and ``[B, T]`` layout, every transform a CUDA kernel behind ``libmpcg_b200.so``.

Randomness.  By default each function draws its parameters exactly as the reference does -- same calls, same
order (``np.random.choice`` for the noise level, ``torch.rand(B, 1, device=...)`` per row, ``torch.randn_like``
for the noise, ``np.random.uniform`` for the EQ band edges) -- so a seeded run consumes the RNG streams like
the reference and can be compared with it.  Every function also accepts its draws as keyword arguments
("injected-parameter mode", what the parity tests use), and ``noise="philox"`` replaces the materialised
``randn_like`` tensor by an in-kernel counter-based Philox stream (one HBM read and write less per noise stage).
"""
from __future__ import annotations

from dataclasses import dataclass

import functools

import numpy as np
import torch

from . import _lib, design

_NOISE_STDS = (0.0001, 0.001, 0.01)
_SINE_BANDS = ((0.05, 0.5), (0.001, 0.05))


@dataclass
class AugmentConfig:
    """Mirror of the reference's ``augment/pipelines.py:26-36`` (any object with these attributes works)."""
    ephnogram_dir: str = ""
    mit_dir: str = ""
    prob_hpss: float = 0.75
    prob_noise: float = 0.30
    prob_time_warp: float = 0.25
    prob_wandering_volume: float = 0.75
    prob_banding: float = 0.25
    prob_baseline_wander: float = 0.30
    prob_real_noise: float = 0.5


def _rows2d(x: torch.Tensor) -> torch.Tensor:
    x = _lib.require_cuda_f32(x)
    if x.dim() != 2:
        raise ValueError("torchaug functions take a [B, T] batch")
    return x


def _dev_f32(v, device, shape=None) -> torch.Tensor:
    t = torch.as_tensor(v, dtype=torch.float32, device=device)
    return t.reshape(shape).contiguous() if shape is not None else t.contiguous()


def _stage(x, op, *, fs=1.0, rowp=None, noise=None, mask=None, normalise=False, seed=0, stream_id=0):
    out = torch.empty_like(x)
    _lib.check(_lib.lib().mpcg_aug_stage_f32(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], op, float(fs),
                                             _lib.ptr(rowp), _lib.ptr(noise), _lib.ptr(mask), int(normalise),
                                             int(seed), int(stream_id), _lib.stream_ptr(x)), "augmentation stage")
    return out


@functools.lru_cache(maxsize=8)
def _zero_tables(b: int, device):
    return torch.zeros(b, 8, device=device), torch.zeros(b, device=device)


def _normalise(x: torch.Tensor) -> torch.Tensor:
    """Row-wise zero-mean, peak-normalise, clamp (reference torchaug.py:24-27).  Rows that fit a cluster go through
    the resident-row kernel with every stage masked off (bulk load, one statistics sweep, one store: the row crosses
    HBM once each way without the second read of the two-sweep stage kernel)."""
    x = _rows2d(x)
    b, t = x.shape
    if b > 0 and t > 0:
        rowp, off = _zero_tables(b, x.device)
        out = torch.empty_like(x)
        rc = _lib.lib().mpcg_aug_chain_f32(x.data_ptr(), out.data_ptr(), b, t, 1.0, rowp.data_ptr(), None, off.data_ptr(), 0, 0,
                                           rowp.data_ptr(), off.data_ptr(), None, 0, off.data_ptr(), rowp.data_ptr(), None,
                                           off.data_ptr(), 0, 0, 1, None, 0, _lib.stream_ptr(x))
        if rc != _lib.EUNSUPPORTED:
            _lib.check(rc, "normalise")
            return out
    return _stage(x, _lib.AUG_IDENTITY, normalise=True)


def _mask(batch: int, prob: float, device) -> torch.Tensor:
    return (torch.rand(batch, 1, device=device) < prob).float()


def _apply(x: torch.Tensor, transformed: torch.Tensor, prob: float, *, mask=None) -> torch.Tensor:
    """``_normalise(m * transformed + (1 - m) * x)`` with ``m ~ Bernoulli(prob)`` per row (reference :34-36)."""
    x = _rows2d(x)
    m = _mask(x.shape[0], prob, x.device) if mask is None else _dev_f32(mask, x.device, (x.shape[0], 1))
    return _stage(x, _lib.AUG_SELECT, noise=_rows2d(transformed), mask=m.reshape(-1).contiguous(), normalise=True)


# ------------------------------------------------------------------------------------------------ draws
def _draw_noise(x, std=None, scale=None, noise=None):
    b = x.shape[0]
    std = float(np.random.choice(_NOISE_STDS)) if std is None else float(std)
    scale = torch.rand(b, 1, device=x.device) * 0.1 if scale is None else _dev_f32(scale, x.device, (b, 1))
    if noise is None:
        noise = torch.randn_like(x)
    elif isinstance(noise, str):
        if noise != "philox":
            raise ValueError("noise must be a tensor, None or 'philox'")
        noise = None
    else:
        noise = _rows2d(noise)
    rowp = torch.zeros(b, 8, device=x.device)
    rowp[:, 0:1] = scale * std                      # fp32 product, as the reference forms it
    return rowp, noise


def _draw_sines(x, span, amp=None, freq=None, phase=None):
    b, dev = x.shape[0], x.device
    cols = []
    for k, (lo, hi) in enumerate(_SINE_BANDS):
        a = 0.01 + torch.rand(b, 1, device=dev) * span if amp is None else _dev_f32(amp, dev)[:, k:k + 1]
        f = lo + torch.rand(b, 1, device=dev) * (hi - lo) if freq is None else _dev_f32(freq, dev)[:, k:k + 1]
        p = torch.rand(b, 1, device=dev) if phase is None else _dev_f32(phase, dev)[:, k:k + 1]
        cols += [a, f, p]
    rowp = torch.zeros(b, 8, device=dev)
    rowp[:, :6] = torch.cat(cols, dim=1)
    return rowp


def _draw_bands(low, high, num_bands, bands=None):
    if bands is not None:
        return [tuple(map(float, b)) for b in bands]
    out = []
    for _ in range(num_bands):
        lo = float(np.random.uniform(low, 0.95 * high))
        hi = float(np.random.uniform(lo + 0.05 * (high - low), high))
        out.append((lo, hi))
    return out


@functools.lru_cache(maxsize=64)
def _fast_draw_affine_host(std1: float, std4: float):
    """The same (scale, offset) tables as :func:`_fast_draw_affine`, contiguous float32 on the host ([3, 8] each)."""
    scale, offset = np.zeros((3, 8), np.float32), np.zeros((3, 8), np.float32)
    scale[0, 0], scale[2, 0] = np.float32(0.1) * np.float32(std1), np.float32(0.1) * np.float32(std4)
    for k, (lo, hi) in enumerate(_SINE_BANDS):
        scale[1, 3 * k:3 * k + 3] = (0.24, hi - lo, 1.0)
        offset[1, 3 * k:3 * k + 3] = (0.01, lo, 0.0)
    return scale, offset


@functools.lru_cache(maxsize=64)
def _fast_draw_affine(std1: float, std4: float, device):
    """(scale, offset) [3, 1, 8] turning U(0,1) into the rows of (noise 1 | wandering volume | noise 2) parameter tables:
    noise p[0] = U * 0.1 * std; volume (amp, freq, phase) per band = 0.01 + U * 0.24, lo + U * (hi - lo), U."""
    scale, offset = np.zeros((3, 1, 8), np.float32), np.zeros((3, 1, 8), np.float32)
    scale[0, 0, 0], scale[2, 0, 0] = np.float32(0.1) * np.float32(std1), np.float32(0.1) * np.float32(std4)
    for k, (lo, hi) in enumerate(_SINE_BANDS):
        scale[1, 0, 3 * k:3 * k + 3] = (0.24, hi - lo, 1.0)
        offset[1, 0, 3 * k:3 * k + 3] = (0.01, lo, 0.0)
    return torch.from_numpy(scale).to(device), torch.from_numpy(offset).to(device)


@functools.lru_cache(maxsize=64)
def _fast_draw_probs(p1: float, p2: float, p3: float, device):
    return torch.tensor([[p1], [p2], [p3], [p1]], dtype=torch.float32, device=device)


def _philox_key():
    """A fresh (seed, stream) pair drawn from torch's CPU generator so torch.manual_seed controls it."""
    v = torch.randint(0, 2 ** 62, (2,), dtype=torch.int64)
    return int(v[0]), int(v[1])


# ------------------------------------------------------------------------------------------------ transforms
def add_white_noise(x: torch.Tensor, *, std=None, scale=None, noise=None) -> torch.Tensor:
    """``x + scale * std * randn`` (reference torchaug.py:39-42): one ``std`` per call, one ``scale`` per row."""
    x = _rows2d(x)
    rowp, nz = _draw_noise(x, std, scale, noise)
    seed, sid = _philox_key() if nz is None else (0, 0)
    return _stage(x, _lib.AUG_NOISE, rowp=rowp, noise=nz, seed=seed, stream_id=sid)


def sinusoidal_envelope(x: torch.Tensor, fs: int, *, amp=None, freq=None, phase=None) -> torch.Tensor:
    """``x * (1 + fast + slow)`` sinusoidal volume modulation (reference torchaug.py:45-54)."""
    x = _rows2d(x)
    return _stage(x, _lib.AUG_SINE_MUL, fs=fs, rowp=_draw_sines(x, 0.24, amp, freq, phase))


def baseline_wander(x: torch.Tensor, fs: int, *, amp=None, freq=None, phase=None) -> torch.Tensor:
    """``x + fast + slow`` sinusoidal drift (reference torchaug.py:57-66)."""
    x = _rows2d(x)
    return _stage(x, _lib.AUG_SINE_ADD, fs=fs, rowp=_draw_sines(x, 0.19, amp, freq, phase))


def _warp_curves(amps: torch.Tensor, kernel: int) -> torch.Tensor:
    """[B, P] control gains -> [B, kernel] unit-sum taps by linear interpolation (reference :75-81, on the host
    exactly as the reference does it)."""
    p = amps.shape[1]
    grid = torch.arange(kernel).float()
    idx = torch.clamp(grid / (kernel - 1) * (p - 1), max=p - 1)
    lo, hi = idx.floor().long(), idx.ceil().long()
    curve = amps[:, lo] + (amps[:, hi] - amps[:, lo]) * (idx - lo).unsqueeze(0)
    return curve / curve.sum(dim=-1, keepdim=True)


def amplitude_warp(x: torch.Tensor, num_points: int = 12, kernel: int = 65, *, amps=None) -> torch.Tensor:
    """Per-row smooth gain curve applied as a depthwise FIR over the reflect-padded row (reference :69-85)."""
    x = _rows2d(x)
    b, t = x.shape
    amps = 0.7 + torch.rand(b, num_points) * 0.6 if amps is None else torch.as_tensor(amps, dtype=torch.float32).cpu()
    curves = _warp_curves(amps, kernel).to(x.device).contiguous()
    out = torch.empty_like(x)
    _lib.check(_lib.lib().mpcg_aug_warp_f32(x.data_ptr(), out.data_ptr(), b, t, curves.data_ptr(), kernel,
                                            _lib.stream_ptr(x)), "amplitude warp")
    return out


def _coloured(x: torch.Tensor, fs: float, bands, mask=None) -> torch.Tensor:
    """The five cascaded band-pass sections; with a row mask only the selected rows are computed (the others are
    never read by the mix kernel)."""
    sos = np.ascontiguousarray(design.eq_band_sos(fs, bands), dtype=np.float64)
    out = torch.empty_like(x)
    _lib.check(_lib.lib().mpcg_biquad_cascade_masked_f32(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1],
                                                         sos.ctypes.data, sos.shape[0], _lib.ptr(mask),
                                                         _lib.stream_ptr(x)), "EQ cascade")
    return out


def _eq_mix(x, coloured, mask, mix_only):
    out = torch.empty_like(x)
    _lib.check(_lib.lib().mpcg_aug_eq_mix_f32(x.data_ptr(), coloured.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1],
                                              _lib.ptr(mask), 1 if mix_only else 0, _lib.stream_ptr(x)), "EQ mix")
    return out


def parametric_eq(x: torch.Tensor, fs: float, low: float, high: float, num_bands: int = 5, *, bands=None) -> torch.Tensor:
    """Blend with a stack of random first-order band-pass sections, bands shared by the batch (reference :88-100)."""
    x = _rows2d(x)
    bands = _draw_bands(low, high, num_bands, bands)
    if len(bands) > 6:
        raise ValueError("at most 6 EQ bands per call")
    return _eq_mix(x, _coloured(x, fs, bands), None, True)


def augment_pcg_batch(x: torch.Tensor, fs: int, cfg: AugmentConfig | None = None, *, draws: dict | None = None,
                      noise: str | None = None, fused: bool | None = None, collapse: bool = True,
                      out: torch.Tensor | None = None, fast_draws: bool | None = None) -> torch.Tensor:
    """Noise -> wandering volume -> EQ -> noise, each behind a per-row Bernoulli mask, every row re-normalised
    after every stage (reference torchaug.py:103-111).  ``draws`` injects every random quantity (keys as in
    ``oracle.torch_path.augment_pcg_batch``); ``noise="philox"`` draws the white noise inside the kernel.

    ``fused``: ``None`` = the one-kernel chain (rows resident in cluster shared memory, read once and written once)
    when a row fits an 8-CTA cluster, ``True`` = require it, ``False`` = one kernel per stage (transform + blend +
    normalise each).  Both paths consume the random draws in the reference's order and perform the same arithmetic.
    ``collapse`` (fused path): a stage whose mask is off for a row does not normalise the already normalised row a
    second time -- ``N(N(x)) == N(x)`` up to float32 rounding (~1e-7 of the [-1, 1] range, far inside the 1e-5
    tolerance) -- which saves that stage's sweep and its cluster exchange; ``False`` re-normalises every time.
    ``fast_draws`` (default with ``noise="philox"`` and no injected ``draws``): the per-row random quantities come from
    ONE device launch (``mpcg_aug_draw_f32``: same distributions as the reference, its own Philox stream) instead of the reference's sequence
    of ~40 small launches; ``False`` keeps the reference's draw order."""
    cfg = cfg or AugmentConfig()
    x = _rows2d(x)
    b, dev = x.shape[0], x.device
    d = draws or {}
    if b == 0 or x.shape[1] == 0:                                # nothing to augment
        return x.clone() if out is None else out

    def mask_of(key, prob):
        m = d.get(key)
        return _mask(b, prob, dev).reshape(b) if m is None else _dev_f32(m, dev, (b,))

    def noise_draws(i, mask_key):
        rowp, nz = _draw_noise(x, d.get(f"std{i}"), d.get(f"scale{i}"), d.get(f"noise{i}", noise))
        seed, sid = _philox_key() if nz is None else (0, 0)
        return rowp, nz, seed, sid, mask_of(mask_key, cfg.prob_noise / 4)

    if fused is not False and x.shape[1] > 0 and b > 0:
        if fast_draws is None:
            fast_draws = noise == "philox" and not d
        if fast_draws and (noise != "philox" or d):
            raise ValueError("fast_draws goes with noise='philox' and no injected draws")
        if fast_draws:
            # throughput mode: every per-row quantity from one draw kernel (a [3, B, 8] parameter table and a [4, B]
            # mask table) instead of the reference's ~40 small launches; same distributions, its own random stream
            std1, std4 = float(np.random.choice(_NOISE_STDS)), float(np.random.choice(_NOISE_STDS))
            scale, offset = _fast_draw_affine_host(std1, std4)
            probs = np.array([cfg.prob_noise / 4, cfg.prob_wandering_volume, cfg.prob_banding, cfg.prob_noise / 4], np.float32)
            tab = torch.empty(3, b, 8, device=dev)
            masks = torch.empty(4, b, device=dev)
            dseed, dsid = _philox_key()
            _lib.check(_lib.lib().mpcg_aug_draw_f32(tab.data_ptr(), masks.data_ptr(), b, scale.ctypes.data, offset.ctypes.data,
                                                    probs.ctypes.data, dseed, dsid, _lib.stream_ptr(x)), "per-row draws")
            rowp1, rowp2, rowp4 = tab[0], tab[1], tab[2]
            m1, m2, m3, m4 = masks[0], masks[1], masks[2], masks[3]
            nz1 = nz4 = None
            (seed1, sid1), (seed4, sid4) = _philox_key(), _philox_key()
            bands = _draw_bands(2, 500, 5, None)
        else:
            # every draw first (reference order), then one launch
            rowp1, nz1, seed1, sid1, m1 = noise_draws(1, "mask1")
            rowp2 = _draw_sines(x, 0.24, d.get("amp"), d.get("freq"), d.get("phase"))
            m2 = mask_of("mask2", cfg.prob_wandering_volume)
            bands = _draw_bands(2, 500, 5, d.get("bands"))
            m3 = mask_of("mask3", cfg.prob_banding)
            rowp4, nz4, seed4, sid4, m4 = noise_draws(2, "mask4")
        sos = np.ascontiguousarray(design.eq_band_sos(fs, bands), dtype=np.float64)
        if out is None:
            out = torch.empty_like(x)
        elif out.shape != x.shape or not out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous() or \
                out.data_ptr() == x.data_ptr():
            raise ValueError("out must be a contiguous CUDA float32 tensor of x's shape that does not alias x")
        work = _lib.aug_workspace(x)
        rc = _lib.lib().mpcg_aug_chain_f32(x.data_ptr(), out.data_ptr(), b, x.shape[1], float(fs), rowp1.data_ptr(),
                                           _lib.ptr(nz1), m1.data_ptr(), seed1, sid1, rowp2.data_ptr(), m2.data_ptr(),
                                           sos.ctypes.data, sos.shape[0], m3.data_ptr(), rowp4.data_ptr(), _lib.ptr(nz4),
                                           m4.data_ptr(), seed4, sid4, 1 if collapse else 0, work.data_ptr(), work.numel(),
                                           _lib.stream_ptr(x))
        if rc != _lib.EUNSUPPORTED:
            _lib.check(rc, "fused augmentation chain")
            return out
        if fused is True:
            raise ValueError("rows of this length do not fit the fused augmentation kernel")
        # (the draws above are reused below so the random stream is consumed once)
        pre = dict(n1=(rowp1, nz1, seed1, sid1, m1), rowp2=rowp2, m2=m2, bands=bands, m3=m3, n4=(rowp4, nz4, seed4, sid4, m4))
    else:
        if fused is True:
            raise ValueError("nothing to fuse for an empty batch")
        pre = None

    x = _normalise(x)

    def noise_stage(x, i, mask_key, key):
        rowp, nz, seed, sid, m = pre[key] if pre else noise_draws(i, mask_key)
        return _stage(x, _lib.AUG_NOISE, rowp=rowp, noise=nz, mask=m, normalise=True, seed=seed, stream_id=sid)

    x = noise_stage(x, 1, "mask1", "n1")
    rowp = pre["rowp2"] if pre else _draw_sines(x, 0.24, d.get("amp"), d.get("freq"), d.get("phase"))
    m2 = pre["m2"] if pre else mask_of("mask2", cfg.prob_wandering_volume)
    x = _stage(x, _lib.AUG_SINE_MUL, fs=fs, rowp=rowp, mask=m2, normalise=True)
    bands = pre["bands"] if pre else _draw_bands(2, 500, 5, d.get("bands"))
    m3 = pre["m3"] if pre else mask_of("mask3", cfg.prob_banding)
    x = _eq_mix(x, _coloured(x, fs, bands, m3), m3, False)
    x = noise_stage(x, 2, "mask4", "n4")
    if out is not None:
        out.copy_(x)
        return out
    return x


# ------------------------------------------------------------------------------------------------ NumPy-only rows
def time_warp(x: torch.Tensor, fs: int, rate: float, keep_length: bool = False) -> torch.Tensor:
    """Device stand-in for ``primitives.time_stretch(x, fs, rate, keep_length)`` (reference primitives.py:30-34).
    The reference calls the rubberband phase vocoder through a subprocess, which has no reproducible arithmetic;
    this is a DEFINED resampling warp ``y[j] = x(j * rate)`` (Catmull-Rom), length ``round(T / rate)``, cut to ``T``
    with ``keep_length``.  Parity with the reference is unpinned."""
    x = _rows2d(x)
    b, t = x.shape
    n = int(round(t / float(rate)))
    out = torch.empty((b, n), device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().mpcg_time_warp_f32(x.data_ptr(), out.data_ptr(), b, t, n, float(rate), _lib.stream_ptr(x)),
               "time warp")
    return out[:, :t].contiguous() if keep_length and n > t else out


def mix_noise(x: torch.Tensor, bank: torch.Tensor, *, rows=None, starts=None, scale=None, hi: float = 0.5) -> torch.Tensor:
    """``abs_max_normalise(x + scale * abs_max_normalise(crop))`` with ``crop`` a random ``T``-sample span of a
    random record of the device-resident noise ``bank [K, Tn]`` (reference noise_sources.py:53-64 and
    pipelines.py:59-60; record loading stays on the host).  ``scale`` defaults to the reference's
    ``choice([0, U(0, hi)])`` per row."""
    x = _rows2d(x)
    bank = _lib.require_cuda_f32(bank, "bank")
    b, t = x.shape
    k, tn = bank.shape
    if tn < t:
        raise ValueError("noise records must be at least as long as the signals")
    dev = x.device
    injected = rows is not None or starts is not None
    rows = torch.randint(0, k, (b,), device=dev) if rows is None else torch.as_tensor(rows, device=dev)
    starts = (torch.rand(b, device=dev) * (tn - t + 1)).long().clamp_(0, tn - t) if starts is None else torch.as_tensor(starts, device=dev)
    if injected:                                             # the kernel trusts its tables: check what the caller brought
        if rows.numel() != b or starts.numel() != b:
            raise ValueError("rows and starts need one entry per signal")
        if bool(((rows < 0) | (rows >= k) | (starts < 0) | (starts > tn - t)).any()):
            raise ValueError("rows must lie in [0, K) and starts in [0, Tn - T]")
    if scale is None:
        scale = torch.where(torch.rand(b, device=dev) < 0.5, torch.zeros(b, device=dev), torch.rand(b, device=dev) * hi)
    rows, starts = rows.long().contiguous(), starts.long().contiguous()
    scale = _dev_f32(scale, dev, (b,))
    out = torch.empty_like(x)
    _lib.check(_lib.lib().mpcg_mix_noise_f32(x.data_ptr(), bank.data_ptr(), out.data_ptr(), b, t, k, tn, rows.data_ptr(),
                                             starts.data_ptr(), scale.data_ptr(), _lib.stream_ptr(x)), "noise mix")
    return out
