"""Batched, probabilistic augmentation pipelines on the device: the composition layer of the reference's
``augment/pipelines.py:43-148`` (``augment_pcg``, ``augment_ecg``, ``augment_pcg_ecg``, ``augment_multi_pcg``) over the
primitives of ``augment/primitives.py`` -- HPSS recombination, additive white noise, time stretch, wandering volume,
parametric-EQ banding, baseline wander, recorded clinical noise -- every stage behind a per-row Bernoulli mask with the
probabilities of ``AugmentConfig``.

What is the same as the reference, row by row: the stage order; which stages normalise (the NumPy primitives normalise
inside the transform, so a stage that is skipped for a row leaves that row untouched -- unlike ``torchaug``, which
re-normalises every row after every stage); ``minmax_normalise`` on entry (abs-max for the multichannel pipeline),
NumPy ``abs_max_normalise`` on exit; parameters shared where the reference shares them (one stretch rate and one HPSS
length for a PCG/ECG pair, every parameter for the channels of a vest recording).

What is batched: stage parameters that the reference draws per signal and that steer a whole kernel launch -- the
HPSS transform sizes / margins / median lengths, the stretch rate, the EQ bands -- are drawn once per call and shared
by the rows the stage selects; per-row quantities (masks, noise scales, sine amplitudes / rates / phases, noise crops)
stay per row.  ``draws`` injects every random quantity (the parity mode of ``tests/test_gpu_pipelines.py``); without
it they are drawn with ``random`` / ``numpy.random`` / ``torch`` (throughput mode, in-kernel Philox noise).

Stages that change a row's length (HPSS trims to whole hops, the stretch rescales) split the batch into groups of
equal length; results come back zero-padded to the input length with the valid lengths beside them.

Parity: stages N, O, P, R and the normalisers are pinned through the reference's own functions;
HPSS, the time stretch (a DEFINED Catmull-Rom resampling warp here, rubberband in the reference) and the recorded-noise
file reads are unpinned; the composition itself is held to ``oracle/pipelines_path.py``.
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib, design, hpss, normalize, torchaug, torchproc
from .torchaug import AugmentConfig

PCG_STRETCH = (1.004, 1.006)            # reference pipelines.py:22-23
PAIR_STRETCH = (0.8, 1.2)
_MULTI_PROB_NOISE, _MULTI_PROB_TIME_WARP, _MULTI_PROB_WANDER, _MULTI_PROB_REAL_NOISE = 0.30, 0.35, 0.75, 0.25
_MULTI_STRETCH = (0.7, 1.3)
_NOISE_STDS = (0.0001, 0.001, 0.01)


# ------------------------------------------------------------------------------------------------ recorded noise
@dataclass
class NoiseBank:
    """Clinical noise records resident on the device, already at the signal rate (the file reads of
    ``noise_sources.py:24-30`` stay with the caller).  ``records``: ``[K, Tn]``; ``groups``: for every component of one
    noise draw the rows of ``records`` it may come from and its scale range ``(lo, hi)`` -- EPHNOGRAM: two components
    (AUX1 / AUX2 of one random record, scale ``choice([0, U(0, 0.05)])``, sum normalised); MIT-BIH: three components
    (em / bw / ma, ``choice([0, U(0, .25)])`` / ``(0, .5)`` / ``(0, .25)``, plain sum)."""
    records: torch.Tensor
    groups: list = field(default_factory=list)          # [(row_indices, lo, hi)] per component
    normalise_sum: bool = True
    tied_record: bool = True                             # components come from the same record index (EPHNOGRAM AUX1 / AUX2)

    @staticmethod
    def from_arrays(arrays, fs_records, fs: int, groups, *, normalise_sum=True, tied_record=True, device="cuda") -> "NoiseBank":
        """``arrays``: host 1-D arrays; each is resampled to ``fs`` with the SciPy polyphase arithmetic of the reference
        (``sp.resample_poly(sig, int(fs), rec.fs)``, noise_sources.py:42-43,58) on the device, then cut to the shortest."""
        rows = []
        for a, fr in zip(arrays, fs_records):
            x = torch.as_tensor(np.asarray(a, dtype=np.float32), device=device)[None]
            rows.append(torchproc.resample(x, float(fr), float(int(fs)), mode="numpy")[0])
        n = min(r.shape[0] for r in rows)
        return NoiseBank(torch.stack([r[:n] for r in rows]).contiguous(), list(groups), normalise_sum, tied_record)

    def draw(self, b: int, t: int, device) -> dict:
        """Per-row crops and scales in the reference's order: record, then per component the crop start and the scale."""
        k, tn = self.records.shape
        nc = len(self.groups)
        rows = np.zeros((b, nc), np.int64); starts = np.zeros((b, nc), np.int64); scale = np.zeros((b, nc), np.float32)
        for r in range(b):
            pick = random.randrange(len(self.groups[0][0])) if self.tied_record else None
            for c, (idx, lo, hi) in enumerate(self.groups):
                rows[r, c] = idx[pick] if self.tied_record else idx[random.randrange(len(idx))]
                starts[r, c] = random.randint(0, tn - t) if tn > t else 0
                scale[r, c] = random.choice([0.0, lo + random.random() * (hi - lo)])
        return {"rows": rows, "starts": starts, "scale": scale}

    def combine(self, b: int, t: int, d: dict) -> torch.Tensor:
        dev = self.records.device
        k, tn = self.records.shape
        if tn < t:
            raise ValueError("noise records must be at least as long as the signals")
        rows = torch.as_tensor(np.ascontiguousarray(d["rows"], dtype=np.int64), device=dev)
        starts = torch.as_tensor(np.ascontiguousarray(d["starts"], dtype=np.int64), device=dev)
        scale = torch.as_tensor(np.ascontiguousarray(d["scale"], dtype=np.float32), device=dev)
        nc = rows.shape[1]
        out = torch.empty((b, t), device=dev, dtype=torch.float32)
        _lib.check(_lib.lib().mpcg_noise_combine_f32(self.records.data_ptr(), out.data_ptr(), b, t, k, tn, nc, rows.data_ptr(),
                                                     starts.data_ptr(), scale.data_ptr(), 1 if self.normalise_sum else 0,
                                                     _lib.stream_ptr(out)), "recorded noise")
        return out


# ------------------------------------------------------------------------------------------------ draws
def _rf(lo, hi):
    return lo + random.random() * (hi - lo)


def _row_mask(b, p, injected=None):
    if injected is not None:
        return np.asarray(injected, dtype=bool).reshape(b)
    return np.random.rand(b) < p


def _noise_draw(b, t, dev, inj):
    """add_white_noise (primitives.py:44-46): std ~ choice, scale ~ U(0, 0.1) per signal; noise injected or Philox."""
    if inj is not None:
        sig = np.asarray(inj["sigma"], dtype=np.float32).reshape(b)
        nz = inj.get("noise")
        if nz is not None:
            nz = np.asarray(nz, dtype=np.float32)
            if nz.shape[1] < t:
                raise ValueError("injected noise is shorter than the rows it is added to (a stretch below 1 lengthens rows)")
            nz = torch.as_tensor(np.ascontiguousarray(nz[:, :t]), device=dev)
    else:
        sig = np.array([random.choice(_NOISE_STDS) * _rf(0.0, 0.1) for _ in range(b)], dtype=np.float32)
        nz = None
    rowp = torch.zeros(b, 8, device=dev)
    rowp[:, 0] = torch.as_tensor(sig, device=dev)
    return rowp, nz


def _sine_draw(b, dev, a_lo, a_hi, inj):
    """(amp, freq, phase) of the fast and the slow sine (primitives.py:59-70), per signal."""
    if inj is not None:
        tab = np.asarray(inj, dtype=np.float32).reshape(b, 6)
    else:
        tab = np.array([[_rf(a_lo, a_hi), _rf(0.05, 0.5), _rf(0, 1), _rf(a_lo, a_hi), _rf(0.001, 0.05), _rf(0, 1)] for _ in range(b)],
                       dtype=np.float32)
    rowp = torch.zeros(b, 8, device=dev)
    rowp[:, :6] = torch.as_tensor(tab, device=dev)
    return rowp


def _band_draw(low, high, num_bands=5, inj=None):
    """parametric_eq's bands (primitives.py:77-80), one set per call."""
    if inj is not None:
        return [tuple(map(float, bnd)) for bnd in inj]
    bands = []
    for _ in range(num_bands):
        lo = float(np.random.uniform(low, 0.95 * high))
        hi = random.choice([float(np.random.uniform(lo + 0.05 * (high - low), high)), lo + (high - low) / num_bands])
        bands.append((lo, hi))
    return bands


# ------------------------------------------------------------------------------------------------ masked stages
def _mask_dev(m: np.ndarray, dev) -> torch.Tensor:
    return torch.as_tensor(m.astype(np.float32), device=dev)


def _philox():
    return torchaug._philox_key()


def _stage_noise(x, m, inj):
    if not m.any():
        return x
    rowp, nz = _noise_draw(x.shape[0], x.shape[1], x.device, inj)
    seed, sid = (0, 0) if nz is not None else _philox()
    return torchaug._stage(x, _lib.AUG_NOISE, rowp=rowp, noise=nz, mask=_mask_dev(m, x.device), normalise=2, seed=seed, stream_id=sid)


def _stage_sine(x, fs, m, op, a_lo, a_hi, inj):
    if not m.any():
        return x
    rowp = _sine_draw(x.shape[0], x.device, a_lo, a_hi, inj)
    return torchaug._stage(x, op, fs=fs, rowp=rowp, mask=_mask_dev(m, x.device), normalise=2)


def _stage_eq(x, fs, m, low, high, inj):
    if not m.any():
        return x
    bands = _band_draw(low, high, 5, inj)
    md = _mask_dev(m, x.device)
    e = torchaug._eq_mix(x, torchaug._coloured(x, fs, bands, md), md, True)          # N(N(c)/50 + N(x)) (selected rows)
    return torchaug._stage(x, _lib.AUG_SELECT, noise=e, mask=md, normalise=0)          # the others pass through untouched


def _stage_real_noise(x, m, bank, inj):
    if bank is None or not m.any():
        return x
    d = inj if inj is not None else bank.draw(x.shape[0], x.shape[1], x.device)
    nz = bank.combine(x.shape[0], x.shape[1], d)
    rowp = torch.zeros(x.shape[0], 8, device=x.device)
    rowp[:, 0] = 1.0
    return torchaug._stage(x, _lib.AUG_NOISE, rowp=rowp, noise=nz, mask=_mask_dev(m, x.device), normalise=0)   # x + noise (pipelines.py:59-60)


def _np_norm(x):
    return torchproc.abs_max_normalise(x, mode="numpy")


class _Groups:
    """The batch as groups of rows of equal length (length-changing stages split groups)."""

    def __init__(self, x):
        self.items = [(np.arange(x.shape[0]), x)]

    def map_rows(self, fn):
        """fn(idx, x) -> x' of the same row count (any length); keeps the grouping."""
        self.items = [(idx, fn(idx, x)) for idx, x in self.items]

    def split(self, mask: np.ndarray, fn):
        """Rows with mask on go through fn(idx, x_sel) (may change the length) and form their own group."""
        out = []
        for idx, x in self.items:
            sel = mask[idx]
            if sel.any():
                si = torch.as_tensor(np.flatnonzero(sel), device=x.device)
                out.append((idx[sel], fn(idx[sel], x.index_select(0, si).contiguous())))
            if (~sel).any():
                ki = torch.as_tensor(np.flatnonzero(~sel), device=x.device)
                out.append((idx[~sel], x.index_select(0, ki).contiguous()))
        self.items = out

    def gather(self, b, t, dev):
        out = torch.zeros((b, t), device=dev, dtype=torch.float32)
        lengths = np.zeros(b, np.int64)
        for idx, x in self.items:
            n = min(x.shape[1], t)
            out[torch.as_tensor(idx, device=dev), :n] = x[:, :n]
            lengths[idx] = n
        return out, lengths


def _prep(x):
    x = _lib.require_cuda_f32(x)
    if x.dim() == 1:
        x = x[None]
    if x.dim() != 2:
        raise ValueError("pipelines take [T] or [B, T]")
    return x.contiguous()


def _stretch(x, fs, rate, keep_length=False):
    return _np_norm(torchaug.time_warp(x, fs, rate, keep_length=keep_length))


# ------------------------------------------------------------------------------------------------ pipelines
def augment_pcg(pcg: torch.Tensor, fs: int, cfg: AugmentConfig | None = None, *, draws: dict | None = None,
                noise_bank: NoiseBank | None = None):
    """Single-channel PCG windows ``[B, T]`` (reference ``augment_pcg``, pipelines.py:43-61; HPSS keeps 4 components,
    micro-stretch).  Returns ``(y [B, T], lengths [B])``: rows shortened by HPSS / the stretch are zero-padded."""
    cfg = cfg or AugmentConfig()
    d = draws or {}
    x = _prep(pcg)
    b, t = x.shape
    g = _Groups(normalize.minmax_normalise(x, per_row=True))
    m = _row_mask(b, cfg.prob_hpss, d.get("mask_hpss"))
    if m.any():
        hp = d.get("hpss") or hpss.draw_recombine_params(False)
        g.split(m, lambda idx, xs: hpss.hpss_recombine(xs, False, params=hp)[0])
    m = _row_mask(b, cfg.prob_noise / 4, d.get("mask_noise1"))
    g.map_rows(lambda idx, xs: _stage_noise(xs, m[idx], _sub(d.get("noise1"), idx)))
    m = _row_mask(b, cfg.prob_time_warp, d.get("mask_warp"))
    if m.any():
        rate = d.get("rate") or _rf(*PCG_STRETCH)
        g.split(m, lambda idx, xs: _stretch(xs, fs, rate))
    m = _row_mask(b, cfg.prob_wandering_volume, d.get("mask_volume"))
    g.map_rows(lambda idx, xs: _stage_sine(xs, fs, m[idx], _lib.AUG_SINE_MUL, 0.01, 0.25, _sub(d.get("volume"), idx)))
    m = _row_mask(b, cfg.prob_noise / 4, d.get("mask_noise2"))
    g.map_rows(lambda idx, xs: _stage_noise(xs, m[idx], _sub(d.get("noise2"), idx)))
    m = _row_mask(b, cfg.prob_banding, d.get("mask_eq"))
    bands = _band_draw(2, 500, 5, d.get("bands")) if m.any() else None
    g.map_rows(lambda idx, xs: _stage_eq(xs, fs, m[idx], 2, 500, bands))
    if noise_bank is not None:
        m = _row_mask(b, cfg.prob_real_noise, d.get("mask_real"))
        g.map_rows(lambda idx, xs: _stage_real_noise(xs, m[idx], noise_bank, _sub(d.get("real"), idx, xs.shape[1])))
    g.map_rows(lambda idx, xs: _np_norm(xs))
    return g.gather(b, t, x.device)


def augment_ecg(ecg: torch.Tensor, fs: int, cfg: AugmentConfig | None = None, *, draws: dict | None = None,
                noise_bank: NoiseBank | None = None):
    """ECG windows ``[B, T]`` (reference ``augment_ecg``, pipelines.py:64-80).  Returns ``(y, lengths)``."""
    cfg = cfg or AugmentConfig()
    d = draws or {}
    x = _prep(ecg)
    b, t = x.shape
    g = _Groups(normalize.minmax_normalise(x, per_row=True))
    m = _row_mask(b, cfg.prob_noise / 4, d.get("mask_noise1"))
    g.map_rows(lambda idx, xs: _stage_noise(xs, m[idx], _sub(d.get("noise1"), idx)))
    m = _row_mask(b, cfg.prob_baseline_wander, d.get("mask_wander"))
    g.map_rows(lambda idx, xs: _stage_sine(xs, fs, m[idx], _lib.AUG_SINE_ADD, 0.01, 0.2, _sub(d.get("wander"), idx)))
    m = _row_mask(b, cfg.prob_time_warp, d.get("mask_warp"))
    if m.any():
        rate = d.get("rate") or _rf(*PAIR_STRETCH)
        g.split(m, lambda idx, xs: _stretch(xs, fs, rate))
    m = _row_mask(b, cfg.prob_noise / 4, d.get("mask_noise2"))
    g.map_rows(lambda idx, xs: _stage_noise(xs, m[idx], _sub(d.get("noise2"), idx)))
    m = _row_mask(b, cfg.prob_banding, d.get("mask_eq"))
    bands = _band_draw(0.25, 100, 5, d.get("bands")) if m.any() else None
    g.map_rows(lambda idx, xs: _stage_eq(xs, fs, m[idx], 0.25, 100, bands))
    if noise_bank is not None:
        m = _row_mask(b, cfg.prob_real_noise, d.get("mask_real"))
        g.map_rows(lambda idx, xs: _stage_real_noise(xs, m[idx], noise_bank, _sub(d.get("real"), idx, xs.shape[1])))
    g.map_rows(lambda idx, xs: _np_norm(xs))
    return g.gather(b, t, x.device)


def augment_pcg_ecg(ecg: torch.Tensor, pcg: torch.Tensor, fs: int, cfg: AugmentConfig | None = None, *,
                    draws: dict | None = None, pcg_bank: NoiseBank | None = None, ecg_bank: NoiseBank | None = None):
    """A synchronised ECG / PCG batch (reference ``augment_pcg_ecg``, pipelines.py:82-125): HPSS with 7 components on
    the PCG (the ECG is cut to the same length), ONE stretch rate for both.  Returns ``(e, p, lengths)``."""
    cfg = cfg or AugmentConfig()
    d = draws or {}
    e, p = _prep(ecg), _prep(pcg)
    if e.shape != p.shape:
        raise ValueError("ecg and pcg must have the same shape")
    b, t = p.shape
    # the pair travels as one [2b, T] batch per group: rows 0..n-1 ECG, n..2n-1 PCG, so lengths stay in step
    ge = _Groups(normalize.minmax_normalise(e, per_row=True))
    gp = _Groups(normalize.minmax_normalise(p, per_row=True))

    def both(fe, fp):
        ge.map_rows(fe)
        gp.map_rows(fp)

    m = _row_mask(b, cfg.prob_hpss, d.get("mask_hpss"))
    if m.any():
        hp = d.get("hpss") or hpss.draw_recombine_params(True)
        n_box = {}

        def run_hpss(idx, xs):
            y, n = hpss.hpss_recombine(xs, True, params=hp)
            n_box["n"] = n
            return y
        gp.split(m, run_hpss)
        ge.split(m, lambda idx, xs: xs[:, :n_box["n"]].contiguous())               # e = e[:n]  (pipelines.py:92)
    m1, m2 = _row_mask(b, cfg.prob_noise / 4, d.get("mask_noise1_p")), _row_mask(b, cfg.prob_noise / 4, d.get("mask_noise1_e"))
    both(lambda idx, xs: _stage_noise(xs, m2[idx], _sub(d.get("noise1_e"), idx)),
         lambda idx, xs: _stage_noise(xs, m1[idx], _sub(d.get("noise1_p"), idx)))
    m = _row_mask(b, cfg.prob_baseline_wander, d.get("mask_wander"))
    ge.map_rows(lambda idx, xs: _stage_sine(xs, fs, m[idx], _lib.AUG_SINE_ADD, 0.01, 0.2, _sub(d.get("wander"), idx)))
    m = _row_mask(b, cfg.prob_time_warp, d.get("mask_warp"))
    if m.any():
        rate = d.get("rate") or _rf(*PAIR_STRETCH)                                  # shared by the pair (pipelines.py:98-101)
        ge.split(m, lambda idx, xs: _stretch(xs, fs, rate))
        gp.split(m, lambda idx, xs: _stretch(xs, fs, rate))
    m = _row_mask(b, cfg.prob_wandering_volume, d.get("mask_volume"))
    gp.map_rows(lambda idx, xs: _stage_sine(xs, fs, m[idx], _lib.AUG_SINE_MUL, 0.01, 0.25, _sub(d.get("volume"), idx)))
    m1, m2 = _row_mask(b, cfg.prob_noise / 4, d.get("mask_noise2_p")), _row_mask(b, cfg.prob_noise / 4, d.get("mask_noise2_e"))
    both(lambda idx, xs: _stage_noise(xs, m2[idx], _sub(d.get("noise2_e"), idx)),
         lambda idx, xs: _stage_noise(xs, m1[idx], _sub(d.get("noise2_p"), idx)))
    m = _row_mask(b, cfg.prob_banding, d.get("mask_eq_p"))
    bands = _band_draw(2, 500, 5, d.get("bands_p")) if m.any() else None
    gp.map_rows(lambda idx, xs: _stage_eq(xs, fs, m[idx], 2, 500, bands))
    m = _row_mask(b, cfg.prob_banding, d.get("mask_eq_e"))
    bands_e = _band_draw(0.25, 100, 5, d.get("bands_e")) if m.any() else None
    ge.map_rows(lambda idx, xs: _stage_eq(xs, fs, m[idx], 0.25, 100, bands_e))
    if ecg_bank is not None:
        m = _row_mask(b, cfg.prob_real_noise, d.get("mask_real_e"))
        ge.map_rows(lambda idx, xs: _stage_real_noise(xs, m[idx], ecg_bank, _sub(d.get("real_e"), idx, xs.shape[1])))
    if pcg_bank is not None:
        m = _row_mask(b, cfg.prob_real_noise, d.get("mask_real_p"))
        gp.map_rows(lambda idx, xs: _stage_real_noise(xs, m[idx], pcg_bank, _sub(d.get("real_p"), idx, xs.shape[1])))
    both(lambda idx, xs: _np_norm(xs), lambda idx, xs: _np_norm(xs))
    eo, le = ge.gather(b, t, p.device)
    po, lp = gp.gather(b, t, p.device)
    return eo, po, np.minimum(le, lp)


def augment_multi_pcg(channels: torch.Tensor, fs: int, cfg: AugmentConfig | None = None, *, draws: dict | None = None,
                      noise_bank: NoiseBank | None = None):
    """Vest recordings ``[B, C, T]`` (reference ``augment_multi_pcg``, pipelines.py:127-148): every channel of a recording
    gets the SAME decision, stretch rate, modulation and recorded noise, so cross-channel timing is preserved; the white
    noise is drawn per channel, as in the reference's list comprehension.  Returns ``(y [B, C, T], lengths [B])`` (the
    stretch keeps the length for rates below 1 and shortens the recording for rates above, as ``y[:len(x)]`` does)."""
    cfg = cfg or AugmentConfig()
    d = draws or {}
    x = _lib.require_cuda_f32(channels)
    if x.dim() != 3:
        raise ValueError("augment_multi_pcg takes [B, C, T]")
    b, c, t = x.shape
    g = _Groups(_np_norm(x.reshape(b * c, t).contiguous()))
    rep = lambda m: np.repeat(m, c)                                                 # one decision per recording -> its rows
    m1 = rep(_row_mask(b, _MULTI_PROB_NOISE / 4, d.get("mask_noise1")))
    g.map_rows(lambda idx, xs: _stage_noise(xs, m1[idx], _sub(d.get("noise1"), idx)))
    mw = rep(_row_mask(b, _MULTI_PROB_TIME_WARP, d.get("mask_warp")))
    if mw.any():
        rate = d.get("rate") or _rf(*_MULTI_STRETCH)
        g.split(mw, lambda idx, xs: _stretch(xs, fs, rate, keep_length=True))
    mv = rep(_row_mask(b, _MULTI_PROB_WANDER, d.get("mask_volume")))
    if mv.any():
        vol = d.get("volume")
        tab = np.asarray(vol, np.float32).reshape(b, 6) if vol is not None else \
            np.array([[_rf(0.01, 0.25), _rf(0.05, 0.5), _rf(0, 1), _rf(0.01, 0.25), _rf(0.001, 0.05), _rf(0, 1)] for _ in range(b)], np.float32)
        tab = np.repeat(tab, c, axis=0)                                             # shared by the channels of a recording
        g.map_rows(lambda idx, xs: _stage_sine(xs, fs, mv[idx], _lib.AUG_SINE_MUL, 0.01, 0.25, tab[idx]))
    m2 = rep(_row_mask(b, _MULTI_PROB_NOISE / 4, d.get("mask_noise2")))
    g.map_rows(lambda idx, xs: _stage_noise(xs, m2[idx], _sub(d.get("noise2"), idx)))
    if noise_bank is not None:
        mb = _row_mask(b, _MULTI_PROB_REAL_NOISE, d.get("mask_real"))
        if mb.any():
            dr = d.get("real") or noise_bank.draw(b, t, x.device)                   # one draw per recording
            mr = rep(mb)

            def real(idx, xs):
                if not mr[idx].any():
                    return xs
                rec = idx // c                                                      # the recording of every row: shared noise
                nz = noise_bank.combine(len(idx), xs.shape[1], {k: np.asarray(v)[rec] for k, v in dr.items()})
                rowp = torch.zeros(len(idx), 8, device=xs.device)
                rowp[:, 0] = 1.0
                return torchaug._stage(xs, _lib.AUG_NOISE, rowp=rowp, noise=nz, mask=_mask_dev(mr[idx], xs.device), normalise=2)
            g.map_rows(real)
    out, lengths = g.gather(b * c, t, x.device)
    return out.reshape(b, c, t), lengths.reshape(b, c).min(axis=1)


def _sub(inj, idx, t=None):
    """Rows `idx` of an injected per-row draw (an array, or a dict of arrays with one leading entry per batch row)."""
    if inj is None:
        return None
    if isinstance(inj, dict):
        return {k: (np.asarray(v)[idx] if np.ndim(v) >= 1 else v) for k, v in inj.items()}
    return np.asarray(inj)[idx]
