import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import torch_path as otp
from wav2vec_heart_sounds_b200 import torchaug as ta, _lib
g = np.load("tests/golden/torchaug_replay.npz", allow_pickle=True)
x = torch.from_numpy(g["x"]); fs = int(g["fs"])
d = {k[6:]: g[k] for k in g.files if k.startswith("chain_") and k != "chain_out"}
T = lambda v: torch.from_numpy(np.asarray(v))
def err(a, b): return float((a.cpu().double() - b).abs().max())
xo = otp.renormalise(x.double()); xg = ta._normalise(x.cuda()); print("N", err(xg, xo))
# stage 1
o1 = otp.blend(xo, otp.add_white_noise(xo, float(d["std1"]), T(d["scale1"]), T(d["noise1"])), T(d["mask1"]))
rowp, nz = ta._draw_noise(xg, float(d["std1"]), d["scale1"], T(d["noise1"]).cuda())
g1 = ta._stage(xg, _lib.AUG_NOISE, rowp=rowp, noise=nz, mask=T(d["mask1"]).reshape(-1).cuda(), normalise=True)
print("stage1", err(g1, o1), d["mask1"].ravel())
o2 = otp.blend(o1, otp.sinusoidal_envelope(o1, fs, T(d["amp"]), T(d["freq"]), T(d["phase"])), T(d["mask2"]))
rowp = ta._draw_sines(g1, 0.24, d["amp"], d["freq"], d["phase"])
g2 = ta._stage(g1, _lib.AUG_SINE_MUL, fs=fs, rowp=rowp, mask=T(d["mask2"]).reshape(-1).cuda(), normalise=True)
print("stage2", err(g2, o2), d["mask2"].ravel())
bands = [tuple(b) for b in d["bands"]]
o3 = otp.blend(o2, otp.parametric_eq(o2, fs, bands), T(d["mask3"]))
c = ta._coloured(g2, fs, bands)
oc = o2
for b_, a_ in otp.eq_sections(fs, bands): oc = otp._lfilter(oc, b_, a_)
print("coloured", err(c, oc), float(oc.abs().max()))
g3 = ta._eq_mix(g2, c, T(d["mask3"]).reshape(-1).cuda(), False)
print("stage3", err(g3, o3), d["mask3"].ravel())
e_o = otp.parametric_eq(o2, fs, bands); e_g = ta._eq_mix(g2, c, None, True); print("eq only", err(e_g, e_o))
print("per-row stage3 err", (g3.cpu().double() - o3).abs().max(dim=1).values)
