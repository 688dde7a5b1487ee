"""Random shapes through the fused augmentation chain against the kernel-per-stage path (same draws): every cluster size,
every compiled chunk length, aligned and unaligned rows, empty cluster ranks, rows with large offsets."""
import sys, random, numpy as np, torch
sys.path.insert(0, ".")
from wav2vec_heart_sounds_b200 import torchaug as ta, AugmentConfig
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
worst = 0.0
for it in range(60):
    t = random.choice([random.randint(40, 700), random.randint(700, 17000), random.randint(17000, 135000)])
    rows = random.randint(1, 6)
    fs = random.choice([2000, 4125, 16000])
    g = torch.Generator(device="cuda").manual_seed(it)
    x = torch.randn(rows, t, device="cuda", generator=g) * random.choice([1e-3, 0.3, 20.0]) + random.choice([0.0, 0.0, 5.0, -300.0])
    cfg = random.choice([AugmentConfig(), AugmentConfig(prob_noise=4.0, prob_wandering_volume=1.0, prob_banding=1.0),
                         AugmentConfig(prob_noise=2.0, prob_wandering_volume=0.5, prob_banding=0.5)])
    noise = random.choice([None, "philox"])
    outs = []
    for fused, collapse in ((True, False), (False, False), (True, True)):
        torch.manual_seed(it); np.random.seed(it)
        outs.append(ta.augment_pcg_batch(x, fs, cfg, noise=noise, fused=fused, collapse=collapse, fast_draws=False if fused else None))
    assert all(torch.isfinite(o).all() for o in outs), (it, t, rows)
    d1 = float((outs[0] - outs[1]).abs().max()); d2 = float((outs[2] - outs[1]).abs().max())
    worst = max(worst, d1)
    flag = "" if d1 < 2e-6 and d2 < 6e-6 else "   <-- CHECK"
    print(f"{it:2d} t={t:6d} rows={rows} fs={fs:5d} noise={noise} fused-vs-stage {d1:.2e} collapse {d2:.2e}{flag}")
print("worst fused-vs-stage", worst)
