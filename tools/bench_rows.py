"""torchproc.preprocess_pcg / preprocess_ecg on 2048 rows x 60000 samples (2 kHz -> 4125 Hz): one fused launch vs the
four stand-alone kernels."""
import sys, json, torch
sys.path.insert(0, ".")
from wav2vec_heart_sounds_b200 import torchproc as tp
from wav2vec_heart_sounds_b200.synth import synth_pcg, synth_ecg
PEAK = 6532.2
def best(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize(); b = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); b = min(b, s.elapsed_time(e))
    return b
for name, x, fn in (("preprocess_pcg", synth_pcg(2048, 60000, 2000.0, seed=1, device="cuda"), tp.preprocess_pcg),
                    ("preprocess_ecg", synth_ecg(2048, 60000, 2000.0, seed=2, device="cuda"), tp.preprocess_ecg)):
    nb = x.numel() * 4 + 2048 * 123750 * 4
    for fused in (None, False):
        ms = best(lambda: fn(x, 2000, 4125, fused=fused))
        print(json.dumps({"op": name, "path": "fused launch" if fused is None else "stand-alone kernels", "ms": round(ms, 3),
                          "GB/s (in + out)": round(nb / ms / 1e6, 1), "frac": round(nb / ms / 1e6 / PEAK, 3)}))
    a, b = fn(x[:64], 2000, 4125), fn(x[:64], 2000, 4125, fused=False)
    print("   max |fused - chained|:", float((a - b).abs().max()))
