"""One warm-up and one measured hpss_split (256 windows, n_fft 1024, hop 64, kernels 17/17) for an ncu launch list."""
import sys, torch
sys.path.insert(0, ".")
from wav2vec_heart_sounds_b200 import hpss
x = torch.randn(256, 64000, device="cuda")
hpss.hpss_split(x, 1024, 64, (1.5, 2.0), (17, 17)); torch.cuda.synchronize()
hpss.hpss_split(x, 1024, 64, (1.5, 2.0), (17, 17)); torch.cuda.synchronize()
print("ok")
