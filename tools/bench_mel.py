"""Default (float64) log-mel tier: the DMMA kernel against the scalar-DFMA kernel it replaces, and the tcgen05 tier."""
import sys, json, torch
sys.path.insert(0, ".")
import wav2vec_heart_sounds_b200 as pkg
from wav2vec_heart_sounds_b200 import _lib
def best(fn, reps=5):
    fn(); torch.cuda.synchronize(); b = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); b = min(b, s.elapsed_time(e))
    return b
for name, rows, t, kw in (("configs[3] 8192 x 64000 @16 kHz, n_fft 1024 / hop 256 / 80 mels", 8192, 64000, dict(sample_rate=16000, n_fft=1024, hop_length=256)),
                          ("generator preset 4096 x 24576 @4 kHz, n_fft 1024 / hop 256 / 80 mels (127 bins)", 4096, 24576, dict(sample_rate=4000, n_fft=1024, hop_length=256))):
    x = torch.randn(rows, t, device="cuda")
    tr = pkg.MelConfig(**kw).build()
    a = pkg.log_mel(x, tr)
    ms_dm = best(lambda: pkg.log_mel(x, tr))
    dm = tr._dm_host; tr._dm_host = None; tr._dev = {}
    b = pkg.log_mel(x, tr)
    ms_fma = best(lambda: pkg.log_mel(x, tr), reps=2)
    tr._dm_host = dm; tr._dev = {}
    tc = pkg.MelConfig(**kw).build(fast=True)
    ms_tc = best(lambda: pkg.log_mel(x, tc))
    print(json.dumps({"case": name, "bins": tr.nbins, "dmma_fp64_ms": round(ms_dm, 3), "dfma_fp64_ms": round(ms_fma, 3), "tcgen05_ms": round(ms_tc, 3),
                      "max_abs_diff_dmma_vs_dfma": float((a - b).abs().max())}))
