"""Random mel configurations: the DMMA float64 tier against the scalar-DFMA float64 kernel (same basis, same arithmetic type)."""
import sys, random, warnings, torch
sys.path.insert(0, ".")
import wav2vec_heart_sounds_b200 as pkg
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
warnings.simplefilter("ignore")
worst = 0.0
for it in range(40):
    n_fft = random.choice([128, 256, 512, 1024, 2048])
    hop = random.choice([random.randint(4, 64), random.randint(4, n_fft), n_fft // 4])
    win = random.choice([n_fft, n_fft, random.randint(max(hop, n_fft // 4), n_fft)])
    fs = random.choice([2000, 4000, 16000])
    n_mels = random.choice([16, 40, 80, 128])
    f_max = random.choice([200.0, 500.0, fs / 2.0])
    t = random.randint(n_fft // 2 + 2, 30000)
    rows = random.randint(1, 5)
    x = torch.randn(rows, t, device="cuda") * random.choice([1e-3, 1.0, 50.0])
    tr = pkg.MelConfig(sample_rate=fs, n_fft=n_fft, hop_length=hop, win_length=win, n_mels=n_mels, f_max=f_max).build()
    a_mel, a_log = tr(x), pkg.log_mel(x, tr)
    dm = tr._dm_host; tr._dm_host = None; tr._dev = {}
    b_mel, b_log = tr(x), pkg.log_mel(x, tr)
    tr._dm_host = dm; tr._dev = {}
    scale = float(b_mel.abs().max()) or 1.0
    d1 = float((a_mel - b_mel).abs().max()) / scale; d2 = float((a_log - b_log).abs().max())
    worst = max(worst, d1, d2)
    flag = "" if d1 < 1e-6 and d2 < 1e-6 and bool(torch.isfinite(a_mel).all()) else "   <-- CHECK"
    print(f"{it:2d} n_fft={n_fft} hop={hop} win={win} fs={fs} mels={n_mels} fmax={f_max} t={t} rows={rows} bins={tr.nbins}: mel {d1:.1e} logmel {d2:.1e}{flag}")
print("worst", worst)
