"""One launch of the default (float64, DMMA) log-mel at a quarter of configs[3] for ncu."""
import sys, torch
sys.path.insert(0, ".")
import wav2vec_heart_sounds_b200 as pkg
x = torch.randn(2048, 64000, device="cuda")
tr = pkg.MelConfig(sample_rate=16000, n_fft=1024, hop_length=256, n_mels=80, f_max=500).build()
for _ in range(3):
    y = pkg.log_mel(x, tr)
torch.cuda.synchronize()
print(tr.backend, tuple(y.shape))
