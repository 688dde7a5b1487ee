"""HPSS conditioning (row H) timing at configs[3] window shape: per-stage kernel times of hpss_split and the whole
hpss_recombine on B windows of 64000 samples @16 kHz (librosa parameter ranges: n_fft 512..2048, hop 16..128)."""
import sys, json, torch
sys.path.insert(0, ".")
from wav2vec_heart_sounds_b200 import hpss, _lib
B, T = 256, 64000
x = torch.randn(B, T, device="cuda")
def best(fn, reps=3):
    fn(); torch.cuda.synchronize(); b = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); b = min(b, s.elapsed_time(e))
    return b
for n_fft, hop, ker in ((1024, 64, (17, 17)), (2048, 128, (30, 30)), (512, 16, (5, 30))):
    frames = 1 + T // hop
    spec = hpss.stft(x, n_fft, hop)
    t_stft = best(lambda: hpss.stft(x, n_fft, hop))
    t_mh = best(lambda: hpss.median_magnitude(spec, ker[0], True))
    t_mp = best(lambda: hpss.median_magnitude(spec, ker[1], False))
    t_all = best(lambda: hpss.hpss_split(x, n_fft, hop, (1.5, 2.0), ker))
    print(json.dumps({"op": "hpss_split", "windows": B, "n_fft": n_fft, "hop": hop, "kernel": ker, "frames": frames,
                      "stft_ms": round(t_stft, 2), "median_time_ms": round(t_mh, 2), "median_freq_ms": round(t_mp, 2),
                      "split_total_ms": round(t_all, 2), "windows_per_s": round(B / t_all * 1e3),
                      "stft_GFLOP/s": round(B * frames * 5 * n_fft * 10 / t_stft / 1e6), "median_Gsel/s": round(B * frames * (n_fft // 2 + 1) / t_mh / 1e6, 1)}))
    del spec
import random
random.seed(0)
t_rec = best(lambda: hpss.hpss_recombine(x[:64].contiguous(), params=dict(n_fft1=1024, hop1=64, n_fft2=1024, hop2=64, margin1=(1.5, 1.5), margin2=(2.0, 3.0),
                                                            kernel1=(17, 17), kernel2=(11, 23), w1=[1.0] * 7, w2=[2.0] * 7, w_mix=0.03)))
print(json.dumps({"op": "hpss_recombine (3 splits + mix)", "windows": 64, "ms": round(t_rec, 2), "windows_per_s": round(64 / t_rec * 1e3)}))
