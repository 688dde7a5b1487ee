import sys, torch
sys.path.insert(0, ".")
import wav2vec_heart_sounds_b200 as pkg
spec = pkg.WindowSpec(4.0)
x4 = torch.randn(2048, 1, 123750, device="cuda")
for _ in range(3):
    r = pkg.preprocess_segment(x4, 4125, 4125, spec, kinds=("ecg",), fused=True)
torch.cuda.synchronize()
print("ok", r.shape)
