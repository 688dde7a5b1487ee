"""One hpss_split at 256 x 64000, n_fft 1024 / hop 64 (for ncu on the STFT / ISTFT kernels)."""
import sys, torch
sys.path.insert(0, ".")
from wav2vec_heart_sounds_b200 import hpss
x = torch.randn(256, 64000, device="cuda")
for _ in range(2):
    h, p, r = hpss.hpss_split(x, 1024, 64, (1.5, 2.0), (17, 17))
torch.cuda.synchronize()
print(tuple(h.shape))
