"""The sinc delay-and-sum gather and its backward pass (reference classify/beamformer.py:41-55) against the vectors the
reference's own module produced in float64 (tests/golden/beamformer.npz: output and autograd gradients) and against the
NumPy oracle.  Tolerance 1e-5 of each quantity's scale."""
import numpy as np
import pytest
import torch

from oracle import beamformer_path as ob
from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bf(built_lib):
    from wav2vec_heart_sounds_b200 import beamformer
    return beamformer


def test_forward_and_gradients_vs_reference_vectors(bf, golden):
    g = golden("beamformer.npz")
    x = torch.tensor(g["x"], dtype=torch.float32, device="cuda", requires_grad=True)
    d = torch.tensor(g["delays"], dtype=torch.float32, device="cuda", requires_grad=True)
    out = bf.delay_and_sum(x, d, int(g["kernel_size"]))
    go = torch.tensor(g["grad_out"], dtype=torch.float32, device="cuda")
    out.backward(go)
    # samples 400..419 of one microphone carry delays beyond the 41-tap window: the reference's normalisation divides by
    # an almost empty tap sum there and amplifies rounding by orders of magnitude (its own float32 run is off by percents);
    # everything that does not read those samples is held to 1e-5, the stretch itself to 1e-3
    well_out = np.ones(g["out"].shape, bool); well_out[1, 400:420] = False
    well_x = np.ones(g["x"].shape, bool); well_x[1, 1, 375:445] = False
    well_d = np.ones(g["x"].shape, bool); well_d[1, 1, 400:420] = False
    o, gx, gd = out.detach().cpu().numpy(), x.grad.cpu().numpy(), d.grad.cpu().numpy()
    for got, want, well in ((o, g["out"], well_out), (gx, g["grad_x"], well_x), (gd, g["grad_delays"], well_d)):
        assert np.abs(got - want)[well].max() < 1e-5 * np.abs(want[well]).max()
        assert np.abs(got - want).max() < 1e-3 * np.abs(want).max()


def test_module_is_a_drop_in_for_the_reference_module(bf, golden):
    """Same constructor and parameter names: the reference's state dict loads, and the forward pass (predictor -> clamp
    -> gather) reproduces the reference's float64 output to float32 accuracy of the transformer stack."""
    g = golden("beamformer.npz")
    mod = bf.TimeVaryingSincBeamformer(num_mics=3, fs=4125.0).cuda().eval()
    state = {str(k): torch.tensor(g["state__" + str(k)]) for k in g["state_keys"]}
    mod.load_state_dict({k: v.float() for k, v in state.items()})
    x = torch.tensor(g["x"], dtype=torch.float32, device="cuda")
    with torch.no_grad():
        got = mod(x)
        delays = torch.clamp(mod.delay_predictor(x), 0.0, mod.max_delay_samples)
    assert rel_err(delays.cpu().numpy(), g["module_delays"]) < 5e-3         # upstream fp32 transformer (fused attention kernels) vs its float64 run
    # the gather itself, on the float64 module's own delays
    ours = bf.delay_and_sum(x, torch.tensor(g["module_delays"], dtype=torch.float32, device="cuda"))
    assert rel_err(ours.cpu().numpy(), g["module_out"]) < 1e-5
    assert rel_err(got.cpu().numpy(), g["module_out"]) < 5e-2               # (carries the predictor's float32 error)
    # trainable end to end
    mod.train()
    y = mod(x).sum()
    y.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in mod.parameters())


@pytest.mark.parametrize("b,m,t,k", [(3, 6, 8250, 41), (1, 1, 100, 41), (2, 4, 777, 9)])
def test_vs_oracle_shapes_and_edges(bf, b, m, t, k):
    rng = np.random.default_rng(b * 100 + t)
    x = rng.standard_normal((b, m, t)).astype(np.float32)
    d = rng.uniform(0, 12.0, (b, m, t)).astype(np.float32)              # (main lobe inside the window: see the golden test)
    d[:, :, :3] = 0.0
    d[:, :, -2:] = 3.0
    got = bf.delay_and_sum(torch.from_numpy(x).cuda(), torch.from_numpy(d).cuda(), k).cpu().numpy()
    want = ob.delay_and_sum(x.astype(np.float64), d.astype(np.float64), k)
    assert got.shape == (b, t) and rel_err(got, want) < 1e-5
    # gradient of x against a finite difference of the oracle along a random direction
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    go = rng.standard_normal((b, t)).astype(np.float32)
    bf.delay_and_sum(xt, torch.from_numpy(d).cuda(), k).backward(torch.from_numpy(go).cuda())
    v = rng.standard_normal(x.shape)
    eps = 1e-6
    fd = ((ob.delay_and_sum(x + eps * v, d.astype(np.float64), k) - ob.delay_and_sum(x - eps * v, d.astype(np.float64), k)) / (2 * eps) * go).sum()
    an = float((xt.grad.cpu().numpy().astype(np.float64) * v).sum())
    assert abs(an - fd) < 1e-4 * max(abs(fd), 1.0)
