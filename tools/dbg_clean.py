import sys; sys.path.insert(0,'.')
import numpy as np, torch
import wav2vec_heart_sounds_b200 as pkg
from wav2vec_heart_sounds_b200.synth import synth_pcg
from oracle import torch_path as otp
x = synth_pcg(32, 60000, 2000, seed=1, device="cuda", spikes=False)
out, e, tr = pkg.preprocess_segment(x[:, None].contiguous(), 2000, 4125, pkg.WindowSpec(4.0), kinds=("pcg",), fused=True, return_trace=True)
print("kernel edits", e.cpu().numpy())
rs = otp.resample(x.cpu().double(), 2000, 4125)
t=[]; otp.remove_spikes(rs.float(), 4125, trace=t)
cnt = np.bincount([r for r, *_ in t], minlength=32)
print("oracle edits (fp32 despike of fp64 resample)", cnt)
fr = rs[3,:60*2062].reshape(60,2062).abs().amax(1)
print(np.round(fr.numpy(),3), float(fr.median()), float(fr.max()))
print(tr[3, :3].cpu().numpy())
