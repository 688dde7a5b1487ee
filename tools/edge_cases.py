import sys, torch, numpy as np
sys.path.insert(0, ".")
import wav2vec_heart_sounds_b200 as pkg
from wav2vec_heart_sounds_b200 import torchaug as ta, filters, datasets
x = torch.randn(5, 3001, device="cuda")
# non-contiguous input
xt = torch.randn(3001, 5, device="cuda").t()
print("noncontig", ta.augment_pcg_batch(xt, 4125, noise="philox").shape)
# 1-D input
try:
    print("1d", ta.augment_pcg_batch(x[0], 4125, noise="philox").shape)
except Exception as e: print("1d raises", type(e).__name__, str(e)[:80])
# empty batch
print("empty", ta.augment_pcg_batch(torch.zeros(0, 100, device="cuda"), 4125).shape)
# tiny rows
print("tiny", ta.augment_pcg_batch(torch.randn(3, 5, device="cuda"), 4125, noise="philox"))
# constant rows
print("const", ta.augment_pcg_batch(torch.ones(2, 64, device="cuda"), 4125, noise="philox").abs().max().item())
fb = datasets.build_fragments_batched([], fs_out=4125, window=pkg.WindowSpec(1.0))
print("empty fb", fb.windows.shape, fb.class_counts())
# NaN row through fused preprocess (torch mode: nan_to_num)
y = torch.randn(2, 20000, device="cuda"); y[0, 100] = float("nan")
o = pkg.preprocess_segment(y, 2000, 4125, pkg.WindowSpec(1.0), fused=True)
print("nan row finite:", torch.isfinite(o).all().item(), "row1 max", o[1].abs().max().item())
oc = pkg.preprocess_segment(y, 2000, 4125, pkg.WindowSpec(1.0), fused=False)
print("nan fused==chained:", torch.equal(torch.nan_to_num(o), torch.nan_to_num(oc)), float((o[1]-oc[1]).abs().max()))
print("mel fast tiny", pkg.log_mel(torch.randn(2, 600, device="cuda"), pkg.MelConfig(sample_rate=16000, n_fft=1024, hop_length=256, n_mels=80, f_max=500).build(fast=True)).shape)
torch.cuda.synchronize(); print("done")
