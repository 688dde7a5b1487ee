"""Two HPSS median launches (time direction, then frequency direction) for an ncu capture: 256 windows, n_fft 1024, hop 64."""
import sys, torch
sys.path.insert(0, ".")
from wav2vec_heart_sounds_b200 import hpss
x = torch.randn(256, 64000, device="cuda")
spec = hpss.stft(x, 1024, 64)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 17
hpss.median_magnitude(spec, k, True); hpss.median_magnitude(spec, k, False)
torch.cuda.synchronize()
print("ok")
