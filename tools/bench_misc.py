"""Timings of the section-8f rank 3/4 kernels at configs[1]-sized rows (CUDA events, best of 5 after 2 warm-up calls)."""
import sys, random
import numpy as np, torch
sys.path.insert(0, ".")
import wav2vec_heart_sounds_b200 as pkg
from wav2vec_heart_sounds_b200 import normalize as nz, envelopes as ev, heart_cycles as H, filters


def best(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


rows, t = 2048, 123750
gen = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(rows, t, device="cuda", generator=gen) * 0.3
gb = 2 * rows * t * 4 / 1e9
for name, fn in (("minmax per row", lambda: nz.minmax_normalise(x, per_row=True)),
                 ("minmax whole tensor", lambda: nz.minmax_normalise_torch(x)),
                 ("z-score", lambda: nz.z_normalise_torch(x)),
                 ("k-peak k=26 per row", lambda: nz.kpeak_normalise_torch(x, per_row=True)),
                 ("k-peak k=26 whole tensor", lambda: nz.kpeak_normalise_torch(x)),
                 ("abs-max (for scale)", lambda: pkg.torchproc.abs_max_normalise(x))):
    ms = best(fn)
    print(f"{name:28s} {ms:8.3f} ms  {gb / ms * 1e3:8.1f} GB/s")
for r, tt in ((512, 123750), (2048, 16500), (256, 480000)):
    xs = torch.randn(r, tt, device="cuda", generator=gen)
    ms = best(lambda: ev.hilbert_envelope(xs), n=3, warm=1)
    print(f"hilbert_envelope {r} x {tt}: {ms:8.3f} ms  {r * tt / ms / 1e6:8.2f} G samples/s")
    ms = best(lambda: ev.homomorphic_envelope(xs, 4125.0), n=3, warm=1)
    print(f"homomorphic_envelope {r} x {tt}: {ms:8.3f} ms")
b, tt, crop = 4096, 32000, 24576
xs = torch.randn(b, tt, device="cuda", generator=gen)
plans = []
for r in range(b):
    cuts = list(range(200 + (r % 50), tt, 3100 + (r % 7) * 40))
    bounds = H.cycle_bounds(tt, cuts)
    order = H.rearrange_order(len(bounds), rng=random.Random(r))
    plans.append([bounds[i] for i in order])
ms = best(lambda: H.rebuild_batch(xs, plans, crop, 40), n=3, warm=1)
print(f"rebuild_batch {b} x {tt} -> {crop} (host plan upload included): {ms:8.3f} ms")
mel = pkg.MelConfig(sample_rate=4000, n_fft=1024, hop_length=256, n_mels=80).build()
ms = best(lambda: pkg.condition_generator_batch(xs, xs, 4000, mel, 96, 256, cycles=plans), n=3, warm=1)
print(f"condition_generator_batch with cycles {b} x {tt}: {ms:8.3f} ms")
ms = best(lambda: pkg.condition_generator_batch(xs, xs, 4000, mel, 96, 256), n=3, warm=1)
print(f"condition_generator_batch plain      {b} x {tt}: {ms:8.3f} ms")
