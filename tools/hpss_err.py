import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from wav2vec_heart_sounds_b200 import hpss as hp
from oracle import hpss_path as oh
for n_fft, hop, margin, kernel in [(512, 32, (1.5, 2.5), (9, 12)), (1024, 64, (1.0, 1.0), (30, 5)), (2048, 128, (2.0, 4.0), (17, 17))]:
    rng = np.random.default_rng(4)
    t = np.arange(6000) / 4000.0
    x = (np.sin(2 * np.pi * 80 * t)[None] + 0.3 * rng.standard_normal((2, 6000))).astype(np.float32)
    x[:, 2000:2010] += 3.0
    h, p, r = hp.hpss_split(torch.from_numpy(x).cuda(), n_fft, hop, margin, kernel)
    for row in range(2):
        wh, wp, wr = oh.hpss_split(x[row], n_fft, hop, margin, kernel)
        scale = max(np.abs(wh).max(), np.abs(wp).max())
        e = [np.abs(a[row].cpu().numpy() - w).max() / scale for a, w in ((h, wh), (p, wp), (r, wr))]
        d = np.abs(h[row].cpu().numpy() - wh) / scale
        print(n_fft, hop, "row", row, "err/scale h p r: %.2e %.2e %.2e" % tuple(e), " frac > 1e-5: %.2e  > 3e-6: %.2e" % ((d > 1e-5).mean(), (d > 3e-6).mean()))
for residual in (True, False):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 4096)).astype(np.float32)
    n = 7 if residual else 4
    p = dict(n_fft1=512, hop1=64, n_fft2=1024, hop2=32, margin1=(1.2, 1.7), margin2=(2.0, 3.0), kernel1=(7, 11),
             kernel2=(5, 30), w1=list(rng.uniform(0.01, 10, n)), w2=list(rng.uniform(0.01, 10, n)), w_mix=0.03)
    got, m = hp.hpss_recombine(torch.from_numpy(x).cuda(), residual, params=p)
    for row in range(2):
        want, mw = oh.hpss_recombine(x[row], p, residual)
        d = np.abs(got[row].cpu().numpy() - want)
        print("recombine residual", residual, "row", row, "max abs %.2e  frac > 1e-5: %.2e" % (d.max(), (d > 1e-5).mean()))
