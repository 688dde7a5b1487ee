"""Small invocations of the section-8f kernels and the rewritten HPSS kernels for compute-sanitizer (memcheck)."""
import sys, random
import numpy as np, torch
sys.path.insert(0, ".")
import wav2vec_heart_sounds_b200 as pkg
from wav2vec_heart_sounds_b200 import normalize as nz, envelopes as ev, heart_cycles as H, hpss
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(5, 1237, device="cuda", generator=g)
for fn in (lambda: nz.minmax_normalise(x, per_row=True), lambda: nz.minmax_normalise_torch(x), lambda: nz.z_normalise_torch(x),
           lambda: nz.kpeak_normalise_torch(x, k=26), lambda: nz.kpeak_normalise(x[:, 1:], per_row=True)):
    fn()
for t in (1, 7, 129, 1000, 4097):
    ev.hilbert_envelope(torch.randn(3, t, device="cuda", generator=g))
ev.homomorphic_envelope(torch.randn(2, 3000, device="cuda", generator=g), 1000.0)
xs = torch.randn(6, 5000, device="cuda", generator=g)
plans = [None if r == 2 else [(100 + 7 * r, 900), (900, 1700 + r), (1700 + r, 1720 + r), (1720 + r, 2600)] for r in range(6)]
H.rebuild_batch(xs, plans, 4000, 40)
mel = pkg.MelConfig(sample_rate=4000, n_fft=1024, hop_length=256, n_mels=80).build(fast=True)
pkg.condition_generator_batch(xs, xs, 4000, mel, 12, 256, cycles=plans)
w = torch.randn(3, 6000, device="cuda", generator=g)
for n_fft, hop, ker in ((512, 16, (5, 30)), (1024, 64, (17, 17)), (2048, 128, (31, 8)), (256, 64, (40, 3))):
    hpss.hpss_split(w, n_fft, hop, (1.5, 2.0), ker)
torch.cuda.synchronize()
print("ok")
