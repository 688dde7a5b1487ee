"""Per-kernel CUDA-event timings of the unfused entry points at a BASELINE config's shape.
Usage: python tools/bench_kernels.py [c1|c2|c5] [rows]   (prints one JSON object per kernel)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from wav2vec_heart_sounds_b200 import WindowSpec, torchproc as tp  # noqa: E402

PEAK = 6532.2
try:
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
shape = {"c1": (64, 60000, 2000, 16000, 4.0), "c2": (2048, 60000, 2000, 4125, 4.0), "c5": (8192 * 6, 32000, 4000, 4125, 2.0)}[cfg]
rows, t_in, fs_in, fs_out, ws = shape
if len(sys.argv) > 2:
    rows = int(sys.argv[2])
spec = WindowSpec(ws)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(rows, t_in, device="cuda", generator=g)
x[:, 5000:5004] += 30.0
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")


def timeit(name, fn, nbytes, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    gbs = nbytes / best / 1e6
    print(json.dumps({"kernel": name, "ms": round(best, 4), "GB/s": round(gbs, 1), "frac_of_measured_peak": round(gbs / PEAK, 3)}))
    return out


rs = timeit("resample", lambda: tp.resample(x, fs_in, fs_out), 4 * rows * t_in)
t = rs.shape[1]
nb = 4 * rows * t
timeit("resample(out bytes incl.)", lambda: tp.resample(x, fs_in, fs_out), 4 * rows * (t_in + t))
timeit("resample numpy-mode", lambda: tp.resample(x, fs_in, fs_out, mode="numpy"), 4 * rows * (t_in + t))
ds = timeit("despike (incl. clone)", lambda: tp.remove_spikes(rs, fs_out), 3 * nb)
bp = timeit("bandpass_cascade", lambda: tp.bandpass_cascade(ds, fs_out, 25.0, 450.0), 2 * nb)
timeit("lowpass", lambda: tp.lowpass(ds, fs_out, 450.0), 2 * nb)
nm = timeit("abs_max_normalise", lambda: tp.abs_max_normalise(bp), 2 * nb)
sg = timeit("segment", lambda: tp.segment(nm, fs_out, spec), nb + 4 * nm.shape[0] * tp.window_count(t, fs_out, spec) * spec.window_len(fs_out))
timeit("chain pcg+segment (unfused)", lambda: tp.segment(tp.preprocess_pcg(x, fs_in, fs_out), fs_out, spec),
       4 * rows * t_in + sg.numel() * 4)
