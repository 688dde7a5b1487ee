"""Random HPSS configurations against the restated librosa algorithm (oracle/hpss_path.py; tools only)."""
import sys, random, numpy as np, torch
sys.path.insert(0, ".")
from wav2vec_heart_sounds_b200 import hpss as hp
from oracle import hpss_path as oh
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
rng = np.random.default_rng(1)
worst = 0.0
for it in range(24):
    n_fft = random.choice([512, 1024, 1024, 2048, 256])
    hop = random.choice([16, 32, 64, 128, 50, 100])
    t = random.randint(max(n_fft, 2000), 12000)
    kernel = (random.choice([5, 9, 17, 24, 30, 31, 40]), random.choice([5, 12, 17, 20, 21, 30, 33]))
    margin = random.choice([(1.0, 1.0), (1.5, 2.5), (2.0, 4.0), (1.0, 3.0)])
    rows = random.randint(1, 3)
    tt = np.arange(t) / 4000.0
    x = (np.sin(2 * np.pi * random.uniform(40, 300) * tt)[None] + 0.3 * rng.standard_normal((rows, t))).astype(np.float32)
    x[:, t // 3:t // 3 + 8] += 3.0
    h, p, r = hp.hpss_split(torch.from_numpy(x).cuda(), n_fft, hop, margin, kernel)
    d = 0.0
    for row in range(rows):
        wh, wp, wr = oh.hpss_split(x[row], n_fft, hop, margin, kernel)
        scale = max(np.abs(wh).max(), np.abs(wp).max(), 1e-12)
        d = max(d, np.abs(h[row].cpu().numpy() - wh).max() / scale, np.abs(p[row].cpu().numpy() - wp).max() / scale,
                np.abs(r[row].cpu().numpy() - wr).max() / scale)
    worst = max(worst, d)
    print(f"{it:2d} n_fft={n_fft} hop={hop} t={t} kernel={kernel} margin={margin} rows={rows}: {d:.2e}" + ("" if d < 1e-5 else "   <-- CHECK"))
print("worst", worst)
