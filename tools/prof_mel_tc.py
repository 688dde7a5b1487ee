import sys, torch
sys.path.insert(0, ".")
import wav2vec_heart_sounds_b200 as pkg
tr = pkg.MelConfig(sample_rate=16000, n_fft=1024, hop_length=256, n_mels=80, f_max=500).build(fast=True)
x = torch.randn(592, 64000, device="cuda")
for _ in range(3):
    m = pkg.log_mel(x, tr)
torch.cuda.synchronize()
print("ok", m.shape, tr.backend)
