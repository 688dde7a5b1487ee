"""Small invocations of the three hot kernels for compute-sanitizer (memcheck / racecheck / synccheck):
the row-streaming fused preprocess kernel (single-tile rows, multi-tile rows with parallel and serial despike rounds,
a ragged batch, the 8/1 and 33/32 resampler instances, channels-last output), the fused augmentation chain (1-, 2- and
4-CTA clusters) and the tensor-core log-mel.

    compute-sanitizer --tool memcheck  python tools/sanitize_hot.py
    compute-sanitizer --tool racecheck python tools/sanitize_hot.py
"""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import wav2vec_heart_sounds_b200 as pkg
from wav2vec_heart_sounds_b200 import AugmentConfig, torchaug
from wav2vec_heart_sounds_b200.synth import synth_pair, synth_pcg

which = sys.argv[1:] or ["fused", "aug", "mel"]
if "fused" in which:
    x = synth_pair(5, 60000, 2000, seed=1, device="cuda")
    x[1, 0, 500] += 40.0                                                   # a stuck frame: the serial hand-over
    x[2, 0, 20000:21500] += 8.0 * torch.sin(torch.arange(1500, device="cuda") * 0.125)   # many passes in one frame
    pkg.preprocess_segment(x, 2000, 4125, pkg.WindowSpec(4.0), kinds=("pcg", "ecg"), fused=True, channel_major=True)
    pkg.preprocess_segment(x, 2000, 4125, pkg.WindowSpec(4.0), kinds=("pcg", "ecg"), fused=True, channels_last=True, return_trace=True)
    pkg.preprocess_segment(x[:, 0, :9000].contiguous(), 2000, 4125, pkg.WindowSpec(1.0), fused=True, mode="numpy")     # single tile
    pkg.preprocess_segment(x[:2, 0].contiguous(), 2000, 16000, pkg.WindowSpec(4.0), fused=True)                          # 8/1, 27 tiles
    y = synth_pcg(6, 32000, 4000.0, seed=2, device="cuda").reshape(1, 6, 32000)
    pkg.preprocess_segment(y, 4000, 4125, pkg.WindowSpec(2.0), fused=True, channels_last=True)                            # 33/32
    pkg.preprocess_segment(x[:, :, :31001].contiguous(), 2000, 4125, pkg.WindowSpec(4.0), kinds=("pcg", "ecg"), channels_last=True,
                           lengths=[31001, 9000, 17, 20000, 2400])                                                          # ragged
    torch.cuda.synchronize()
if "aug" in which:
    for t, rows in ((600, 5), (12347, 4), (40000, 3), (64000, 2)):
        w = torch.randn(rows, t, device="cuda")
        torchaug.augment_pcg_batch(w, 4125, AugmentConfig(prob_noise=1.2, prob_wandering_volume=1.0, prob_banding=1.0), noise="philox", fused=True)
        torchaug.augment_pcg_batch(w, 4125, AugmentConfig(), noise="philox", fused=True, collapse=False)
    torch.cuda.synchronize()
if "mel" in which:
    xs = torch.randn(3, 24576, device="cuda")
    mel = pkg.MelConfig(sample_rate=4000, n_fft=1024, hop_length=256, n_mels=80).build(fast=True)
    pkg.log_mel(xs, mel)
    mel16 = pkg.MelConfig(sample_rate=16000, n_fft=1024, hop_length=256, n_mels=80, f_max=500).build(fast=True)
    pkg.log_mel(torch.randn(2, 64000, device="cuda"), mel16)
    torch.cuda.synchronize()
print("ok")
