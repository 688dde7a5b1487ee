"""Timing variants of the fused kernel at config-2 row shape to see which phase dominates."""
import sys, json, torch
sys.path.insert(0, ".")
import wav2vec_heart_sounds_b200 as pkg
from wav2vec_heart_sounds_b200.synth import synth_pcg, synth_ecg

spec = pkg.WindowSpec(4.0)
def timeit(name, fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"{name:50s} {best:8.3f} ms")
    return r

R = 2048
pcg_spiky = synth_pcg(R, 60000, 2000, seed=1, device="cuda")
pcg_clean = synth_pcg(R, 60000, 2000, seed=1, device="cuda", spikes=False)
ecg = synth_ecg(R, 60000, 2000, seed=2, device="cuda")
noise = torch.randn(R, 60000, device="cuda")
out = None
for name, x, kinds, dsp in [("pcg spiky despike", pcg_spiky, ("pcg",), True), ("pcg clean despike", pcg_clean, ("pcg",), True),
                            ("pcg spiky no-despike", pcg_spiky, ("pcg",), False), ("ecg", ecg, ("ecg",), True),
                            ("noise as ecg", noise, ("ecg",), True), ("noise as pcg despike", noise, ("pcg",), True)]:
    xx = x[:, None].contiguous()
    r, e, tr = pkg.preprocess_segment(xx, 2000, 4125, spec, kinds=kinds, despike=dsp, fused=True, return_trace=True)
    print("   mean despike passes", float(e.float().mean()), "max", int(e.max()))
    timeit(name, lambda: pkg.preprocess_segment(xx, 2000, 4125, spec, kinds=kinds, despike=dsp, fused=True))
# identity resample (4125 -> 4125) isolates filter+output
x4 = torch.randn(R, 1, 123750, device="cuda")
timeit("identity rate, ecg (filter+norm+windows only)", lambda: pkg.preprocess_segment(x4, 4125, 4125, spec, kinds=("ecg",), fused=True))
