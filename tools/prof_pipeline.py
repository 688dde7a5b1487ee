"""Where the composed pipeline (pipelines.augment_pcg incl. HPSS) spends its time: kernel totals from torch.profiler."""
import sys, time, random, numpy as np, torch
sys.path.insert(0, ".")
from wav2vec_heart_sounds_b200 import AugmentConfig, pipelines as pl
from torch.profiler import profile, ProfilerActivity
torch.manual_seed(0); random.seed(7); np.random.seed(7)
x = torch.randn(1024, 64000, device="cuda")
cfg = AugmentConfig()
hp = dict(n_fft1=1024, hop1=64, n_fft2=1024, hop2=64, margin1=(1.5, 1.5), margin2=(2.5, 2.5), kernel1=(17, 17),
          kernel2=(17, 17), w1=[1.0, 2.0, 3.0, 4.0], w2=[4.0, 3.0, 2.0, 1.0], w_mix=0.03)
full = lambda: pl.augment_pcg(x, 16000, cfg, draws={"hpss": hp})
full(); torch.cuda.synchronize()
t0 = time.perf_counter(); full(); torch.cuda.synchronize(); print("wall ms", 1e3 * (time.perf_counter() - t0))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    full(); torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.device_time_total for e in ev)
print("total device ms", tot / 1e3)
for e in sorted(ev, key=lambda e: -e.device_time_total)[:18]:
    print(f"{e.device_time_total / 1e3:9.3f} ms  x{e.count:4d}  {e.key[:90]}")
