"""configs[4] shape on one GPU: Vest-shaped 6-channel PCG, 8 s at 4 kHz -> 4125 Hz, 2 s windows; preprocess + augment."""
import sys, json, torch, numpy as np
sys.path.insert(0, ".")
import wav2vec_heart_sounds_b200 as pkg
from wav2vec_heart_sounds_b200 import torchaug as ta
from wav2vec_heart_sounds_b200.synth import synth_pcg
PEAK = 6532.2
R, C, T = 8192, 6, 32000
x = synth_pcg(R * C, T, 4000.0, seed=5, device="cuda").reshape(R, C, T)
spec = pkg.WindowSpec(2.0)
def best(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize(); b = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); b = min(b, s.elapsed_time(e))
    return b
out = pkg.preprocess_segment(x, 4000, 4125, spec, channels_last=True, fused=True)
ms = best(lambda: pkg.preprocess_segment(x, 4000, 4125, spec, channels_last=True, fused=True, out=out))
nb = x.numel() * 4 + out.numel() * 4
print(json.dumps({"op": "preprocess_segment config 5 chunk (8192 rec x 6 ch x 8 s), channels_last", "out": list(out.shape), "ms": round(ms, 3),
                  "audio_s_per_s": round(R * 8.0 / ms * 1e3), "GB/s": round(nb / ms / 1e6, 1), "frac": round(nb / ms / 1e6 / PEAK, 3)}))
outp = pkg.preprocess_segment(x, 4000, 4125, spec, fused=True)
ms = best(lambda: pkg.preprocess_segment(x, 4000, 4125, spec, fused=True, out=outp))
print(json.dumps({"op": "same, planar output", "ms": round(ms, 3), "GB/s": round(nb / ms / 1e6, 1), "frac": round(nb / ms / 1e6 / PEAK, 3)}))
w = outp.reshape(-1, outp.shape[-1])[: 4096 * 6 * 4]                       # windows as rows [B*C*N, 8250]
ms = best(lambda: ta.augment_pcg_batch(w, 4125, noise="philox"))
nb = 2 * w.numel() * 4
print(json.dumps({"op": f"augment_pcg_batch on {w.shape[0]} x {w.shape[1]} windows (python API)", "ms": round(ms, 3), "GB/s": round(nb / ms / 1e6, 1), "frac": round(nb / ms / 1e6 / PEAK, 3)}))
