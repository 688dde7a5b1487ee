"""Throughput of the augmentation chain at config 3 shape: 4096 windows x 64000 samples @16 kHz."""
import sys, json, torch, numpy as np
sys.path.insert(0, ".")
from wav2vec_heart_sounds_b200 import torchaug as ta, AugmentConfig
PEAK = 6532.2
B, T = 4096, 64000
x = ta._normalise(torch.randn(B, T, device="cuda"))
def timeit(name, fn, nbytes, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(json.dumps({"op": name, "ms": round(best, 3), "GB/s": round(nbytes / best / 1e6, 1), "frac_of_measured_peak": round(nbytes / best / 1e6 / PEAK, 3)}))
nb = 2 * B * T * 4
timeit("_normalise", lambda: ta._normalise(x), nb)
timeit("add_white_noise (randn_like materialised)", lambda: ta.add_white_noise(x), nb)
timeit("add_white_noise (philox)", lambda: ta.add_white_noise(x, noise="philox"), nb)
timeit("sinusoidal_envelope", lambda: ta.sinusoidal_envelope(x, 16000), nb)
timeit("amplitude_warp", lambda: ta.amplitude_warp(x), nb)
timeit("parametric_eq", lambda: ta.parametric_eq(x, 16000, 2, 500), nb)
cfg = AugmentConfig()
timeit("augment_pcg_batch (reference draw order)", lambda: ta.augment_pcg_batch(x, 16000, cfg), nb)
timeit("augment_pcg_batch (philox noise)", lambda: ta.augment_pcg_batch(x, 16000, cfg, noise="philox"), nb)
print("windows/s (philox):", )
import wav2vec_heart_sounds_b200 as pkg
xm = torch.randn(8192, 64000, device="cuda")
for fast in (False, True):
    tr = pkg.MelConfig(sample_rate=16000, n_fft=1024, hop_length=256, n_mels=80, f_max=500).build(fast=fast)
    timeit(f"log_mel config 4 (8192 x 64000 @16k), {tr.backend}", lambda: pkg.log_mel(xm, tr), 8192 * 64000 * 4 + 8192 * 80 * 251 * 4)
