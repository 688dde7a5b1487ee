"""Tiny driver for ncu: one call of each unfused kernel at config-2 row length (256 rows)."""
import sys
import torch
sys.path.insert(0, ".")
from wav2vec_heart_sounds_b200 import WindowSpec, torchproc as tp
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 592
x = torch.randn(rows, 60000, device="cuda")
x[:, 5000:5004] += 30.0
for _ in range(2):
    rs = tp.resample(x, 2000, 4125)
    ds = tp.remove_spikes(rs, 4125)
    bp = tp.bandpass_cascade(ds, 4125, 25.0, 450.0)
    nm = tp.abs_max_normalise(bp)
    sg = tp.segment(nm, 4125, WindowSpec(4.0))
torch.cuda.synchronize()
print("ok", sg.shape)
