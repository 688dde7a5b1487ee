"""Kernel time of the row-streaming fused preprocess kernel at the BASELINE.json shapes (CUDA events, best of N,
inputs resident):  configs[1] (1024 x {PCG, ECG} x 30 s, 2 kHz -> 4125 Hz), configs[0] replicated (rows of
480 000 samples at 16 kHz), the configs[4] per-GPU chunk (6 channels, 4 kHz -> 4125 Hz, 2 s windows)."""
import json, sys, torch
sys.path.insert(0, ".")
import wav2vec_heart_sounds_b200 as pkg
from wav2vec_heart_sounds_b200.synth import synth_pair, synth_pcg

PEAK = 6532.2


def best_ms(fn, reps=8):
    fn(); fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def run(name, x, fs_in, fs_out, ws, **kw):
    spec = pkg.WindowSpec(ws)
    out = pkg.preprocess_segment(x, fs_in, fs_out, spec, fused=True, **kw)
    ms = best_ms(lambda: pkg.preprocess_segment(x, fs_in, fs_out, spec, fused=True, out=out, **kw))
    nbytes = 4 * (x.numel() + out.numel())
    secs = x.shape[0] * x.shape[-1] / fs_in
    print(json.dumps({"case": name, "ms": round(ms, 4), "GB/s": round(nbytes / ms / 1e6, 1), "frac": round(nbytes / ms / 1e6 / PEAK, 4),
                      "audio_s_per_s": round(secs / ms * 1e3), "in": list(x.shape), "out": list(out.shape)}), flush=True)


which = sys.argv[1:] or ["c2", "c1", "c5", "c5cl"]
if "c2" in which:
    x = synth_pair(1024, 60000, 2000, seed=1234, device="cuda")
    run("configs[1] 1024x2x60000 2k->4125 (pcg,ecg) channel-major", x, 2000, 4125, 4.0, kinds=("pcg", "ecg"), channel_major=True)
    run("configs[1] pcg rows only", x[:, 0].contiguous(), 2000, 4125, 4.0)
    run("configs[1] ecg rows only", x[:, 1].contiguous(), 2000, 4125, 4.0, kinds=("ecg",))
    del x
if "c1" in which:
    x = synth_pcg(1024, 60000, 2000.0, seed=5, device="cuda")
    run("configs[0] x16: 1024 rows of 30 s @2k -> 16 kHz (480000 samples), 4 s windows", x, 2000, 16000, 4.0)
    del x
if "c5" in which:
    x = synth_pcg(2048 * 6, 32000, 4000.0, seed=9, device="cuda").reshape(2048, 6, 32000)
    run("configs[4] chunk 2048x6x32000 4k->4125 planar", x, 4000, 4125, 2.0)
    if "c5cl" in which:
        run("configs[4] chunk channels-last", x, 4000, 4125, 2.0, channels_last=True)
