"""Random shapes through the fused row-streaming preprocess kernel against the chain of stand-alone kernels (same
arithmetic, different kernels) and, for ragged batches, against per-recording calls."""
import sys, random, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import wav2vec_heart_sounds_b200 as pkg
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
random.seed(seed); rng = np.random.default_rng(seed)
def spiky(rows, n, spikes):
    t = np.arange(n)
    x = (np.sin(2 * np.pi * t[None] / rng.uniform(20, 80, (rows, 1))) * rng.uniform(0.2, 2.0, (rows, 1)) + 0.05 * rng.standard_normal((rows, n)))
    for r in range(rows):
        for _ in range(spikes):
            if n < 60: break
            at = int(rng.integers(10, n - 20)); w = int(rng.integers(3, 11))
            x[r, at:at + w] += rng.choice([-1.0, 1.0]) * rng.uniform(5, 20)
    return x
worst = 0.0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    fs_in, fs_out = random.choice([(2000, 4125), (2000, 4125), (4000, 4125), (2000, 16000), (4125, 4125)])
    n = random.choice([random.randint(200, 3000), random.randint(3000, 70000), random.randint(70000, 250000)])
    rows, c = random.randint(1, 3), random.choice([1, 1, 2, 3])
    ws = random.choice([1.0, 2.0, 4.0])
    mode = random.choice(["torch", "numpy"])
    kinds = tuple(random.choice(["pcg", "ecg"]) for _ in range(c))
    x = spiky(rows * c, n, random.randint(0, 6)).reshape(rows, c, n) * random.choice([1e-5, 1.0, 1e3]) + random.choice([0.0, 0.0, 0.7, -50.0])
    xd = torch.as_tensor(x, dtype=torch.float32).cuda()
    spec = pkg.WindowSpec(ws)
    try:
        a = pkg.preprocess_segment(xd, fs_in, fs_out, spec, kinds=kinds, mode=mode, fused=True)
    except Exception as e:
        print(f"{it:2d} n={n} {fs_in}->{fs_out} c={c} {kinds} {mode}: fused raised {type(e).__name__}: {e}"); continue
    b = pkg.preprocess_segment(xd, fs_in, fs_out, spec, kinds=kinds, mode=mode, fused=False)
    assert a.shape == b.shape, (a.shape, b.shape)
    d = float((a - b).abs().max()) if a.numel() else 0.0
    fin = bool(torch.isfinite(a).all())
    worst = max(worst, d)
    flag = "" if d < 5e-6 and fin else "   <-- CHECK"
    print(f"{it:2d} n={n:6d} {fs_in}->{fs_out} rows={rows} c={c} {kinds} ws={ws} {mode}: fused-vs-chained {d:.2e}{flag}")
print("worst", worst)
# ragged batches: one launch against per-recording launches
for it in range(6):
    b = random.randint(2, 5)
    lens = [random.randint(900, 90000) for _ in range(b)]
    tm = max(lens)
    x = np.zeros((b, 2, tm), np.float32)
    for i, l in enumerate(lens): x[i, :, :l] = spiky(2, l, 3)
    xd = torch.as_tensor(x).cuda()
    w, counts = pkg.preprocess_segment(xd, 2000, 4125, pkg.WindowSpec(4.0), kinds=("pcg", "ecg"), lengths=lens, fused=True,
                                       channel_major=True)                       # [C, sum N, win]
    off = 0; dmax = 0.0
    for i, l in enumerate(lens):
        one = pkg.preprocess_segment(xd[i:i + 1, :, :l].contiguous(), 2000, 4125, pkg.WindowSpec(4.0), kinds=("pcg", "ecg"), fused=True)
        k = int(counts[i]); assert one.shape[2] == k, (one.shape, k)
        dmax = max(dmax, float((w[:, off:off + k] - one[0]).abs().max()))
        off += k
    assert off == w.shape[1]
    print("ragged", lens, "max diff vs per-recording", dmax, "" if dmax == 0.0 else "   <-- CHECK")
