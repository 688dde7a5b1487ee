"""cardiac-cycle rebuild (datasets/heart_cycles.py:38-69) and the rebuilt branch of the generator
item (datasets/generative.py:62-91) on the device against vectors the reference produced and the oracle on fresh
inputs.  Tolerance 2e-6 absolute on signals normalised to [-1, 1] (float32 storage of the running output)."""
import random

import numpy as np
import pytest
import torch

from oracle import numpy_path as onp

pytestmark = pytest.mark.gpu
TOL = 2e-6


@pytest.fixture(scope="module")
def pk(built_lib):
    import wav2vec_heart_sounds_b200 as pkg
    return pkg


def test_rebuild_vs_golden(pk, golden):
    g = golden("heart_cycles.npz")
    H = pk.heart_cycles
    joins, crop, fade_n = g["joins"].tolist(), int(g["crop"]), int(g["fade_n"])
    x = torch.from_numpy(g["x"]).cuda()
    xn = pk.torchproc.abs_max_normalise(x, mode="numpy")
    bounds = H.cycle_bounds(x.shape[1], joins)
    plans = [[bounds[i] for i in g[f"order_{r}"].tolist()] for r in range(3)]
    y, n = H.rebuild_batch(xn, plans, crop, fade_n)
    for r in range(3):
        want = g[f"rebuilt_{r}"]
        assert int(n[r]) == len(want)
        assert np.abs(y[r, : len(want)].cpu().numpy() - want).max() < TOL
    # the whole item: normalise -> rebuild -> fade -> fit, both waveforms cut at the same joins
    mel = pk.MelConfig(sample_rate=int(g["fs"]), n_fft=1024, hop_length=256, n_mels=80).build()
    item = pk.condition_generator_batch(x, x.flip(0), int(g["fs"]), mel, crop // 256, 256, cycles=plans, fade_ms=10.0)
    for r in range(3):
        assert np.abs(item["ref_audio"][r].cpu().numpy() - g[f"item_{r}"]).max() < TOL
    assert item["con_spec"].shape == (3, 80, crop // 256)
    # short target: the first cycle alone already reaches it; long target: the reference's loop guard ends the rebuild
    one = H.rebuild(xn[0], bounds, 100, fade_n)
    assert one.shape[0] == len(g["rebuilt_short_target"]) and np.abs(one.cpu().numpy() - g["rebuilt_short_target"]).max() < TOL
    two = H.rebuild(xn[0], bounds[:2], 200000, fade_n)
    assert two.shape[0] == len(g["rebuilt_long_target"]) and np.abs(two.cpu().numpy() - g["rebuilt_long_target"]).max() < TOL


def test_rebuild_batch_vs_oracle_mixed_rows(pk):
    """64 rows with their own joins and orders, rows without segmentation (pass-through), cycles shorter than the fade,
    fade lengths 0 / 1 (plain concatenation) and 160."""
    H = pk.heart_cycles
    rng = np.random.default_rng(12)
    t, target = 30000, 20000
    n = np.arange(t)
    x = (np.sin(2 * np.pi * n / 97.0)[None] * np.exp(-(((n % 2900) - 500.0) / 200.0) ** 2)[None] * rng.uniform(0.3, 1.0, (64, 1))
         + 0.03 * rng.standard_normal((64, t))).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    for fade_n in (0, 1, 40, 160):
        plans = []
        for r in range(64):
            if r % 9 == 0:
                plans.append(None)
                continue
            cuts = np.sort(rng.choice(np.arange(1, t), size=rng.integers(3, 14), replace=False)).tolist()
            if r % 5 == 0:
                cuts.append(cuts[-1] + 25)                                   # a cycle shorter than the longer fades
            bounds = H.cycle_bounds(t, cuts)
            order = H.rearrange_order(len(bounds), prob_contiguous=0.3, rng=random.Random(r))
            plans.append([bounds[i] for i in order])
        y, ln = H.rebuild_batch(xd, plans, target, fade_n)
        for r in range(64):
            if plans[r] is None:
                assert int(ln[r]) == t and torch.equal(y[r, :t], xd[r])
                continue
            want = onp.rebuild([x[r].astype(np.float64)[a:b] for a, b in plans[r]], target, fade_n)
            assert int(ln[r]) == len(want), (r, fade_n)
            assert np.abs(y[r, : len(want)].cpu().numpy() - want).max() < TOL, (r, fade_n)


def test_rebuild_argument_errors(pk):
    H = pk.heart_cycles
    x = torch.zeros(2, 100, device="cuda")
    with pytest.raises(ValueError):
        H.rebuild_batch(x, [[(0, 50)]], 10, 4)                               # one plan per row
    with pytest.raises(ValueError):
        H.rebuild_batch(x, [[(0, 50)], [(90, 120)]], 10, 4)                  # cycle outside the signal
    assert H.rebuild(x[0], [], 17, 4).shape == (17,)
    assert H.rearrange_order(1) == [0] and H.rearrange_order(0) == []
