"""Host study behind the EQ cascade's precision choice (csrc/aug_chain.cu): the chunked scan with the first pass (zero-state end state of a
33-sample chunk) in float32 and everything else in float64, against scipy.signal.lfilter in float64, on noise / offset / tonal rows and
on ordinary, narrow and very low bands.  Prints the error of the coloured signal relative to its peak and what is left of it after the
1 / 50 mix.  Run: python tools/eq_precision_study.py   (CPU only, ~1 min)"""
import numpy as np, scipy.signal as sig
rng = np.random.default_rng(0)
fs, T, L = 16000.0, 64000, 33
def group_mats(secs):
    # secs: list of 2 (b, a); DF-II-T state space of the cascade: s = [z0a, z1a, z0b, z1b]
    (b0, a0), (b1, a1) = secs
    A = np.zeros((4, 4)); B = np.zeros(4)
    # section a: y0 = b0[0] x + z0; z0' = b0[1] x - a0[1] y0 + z1; z1' = b0[2] x - a0[2] y0
    # y0 = b0[0] x + s0
    A[0] = [-a0[1], 1, 0, 0]; B[0] = b0[1] - a0[1] * b0[0]
    A[1] = [-a0[2], 0, 0, 0]; B[1] = b0[2] - a0[2] * b0[0]
    # section b driven by y0 = s0 + b0[0] x: y1 = b1[0] y0 + s2
    A[2] = [b1[1] - a1[1] * b1[0], 0, -a1[1], 1]; B[2] = (b1[1] - a1[1] * b1[0]) * b0[0]
    A[3] = [b1[2] - a1[2] * b1[0], 0, -a1[2], 0]; B[3] = (b1[2] - a1[2] * b1[0]) * b0[0]
    return A, B
def run_group(x, secs, p1_dtype):
    A, B = group_mats(secs)
    n = len(x); nch = -(-n // L); xp = np.zeros(nch * L); xp[:n] = x
    X = xp.reshape(nch, L)
    W = np.zeros((L, 4)); v = B.copy()
    for j in range(L - 1, -1, -1):
        W[j] = v; v = A @ v
    if p1_dtype == np.float32:
        P = np.zeros((nch, 4), np.float32); Wf = W.astype(np.float32); Xf = X.astype(np.float32)
        for j in range(L):
            P = (Wf[j][None, :] * Xf[:, j:j + 1] + P).astype(np.float32)   # fmaf-like (product exact in float64 then rounded)
        P = P.astype(np.float64)
    else:
        P = X @ W
    M = np.linalg.matrix_power(A, L)
    Z = np.zeros((nch, 4)); z = np.zeros(4)
    for k in range(nch):
        Z[k] = z; z = M @ z + P[k]
    # pass 2 exact in float64, vectorised over chunks
    (b0, a0), (b1, a1) = secs
    z0, z1, z2, z3 = Z[:, 0].copy(), Z[:, 1].copy(), Z[:, 2].copy(), Z[:, 3].copy()
    Y = np.zeros_like(X)
    for j in range(L):
        xv = X[:, j]
        y0 = b0[0] * xv + z0; z0 = b0[1] * xv - a0[1] * y0 + z1; z1 = b0[2] * xv - a0[2] * y0
        y1 = b1[0] * y0 + z2; z2 = b1[1] * y0 - a1[1] * y1 + z3; z3 = b1[2] * y0 - a1[2] * y1
        Y[:, j] = y1
    return Y.reshape(-1)[:n].astype(np.float32).astype(np.float64)
def cascade(x, bands, p1_dtype):
    secs = [sig.butter(1, [lo / (fs / 2), hi / (fs / 2)], btype="band") for lo, hi in bands]
    secs.append((np.array([1.0, 0, 0]), np.array([1.0, 0, 0])))
    y = x
    for g in range(3):
        y = run_group(y, secs[2 * g:2 * g + 2], p1_dtype)
    return y
def ref(x, bands):
    y = x
    for lo, hi in bands:
        b, a = sig.butter(1, [lo / (fs / 2), hi / (fs / 2)], btype="band"); y = sig.lfilter(b, a, y)
    return y
cases = {"default-like": [(2 + 498 * rng.random(), 0) for _ in range(5)]}
def mk(edges): return [tuple(sorted(e)) for e in edges]
tests = [("random", mk(rng.uniform(2, 500, (5, 2)))), ("random2", mk(rng.uniform(2, 500, (5, 2)))),
         ("narrow low", [(2, 2.5), (2.2, 3), (3, 4), (2, 6), (2.5, 5)]), ("very narrow", [(2, 2.05)] * 5),
         ("wide", [(2, 500)] * 5), ("mixed", [(2, 3), (100, 101), (400, 500), (2, 500), (50, 60)])]
for kind in ("noise", "noise+dc", "tone5"):
    x = rng.standard_normal(T)
    if kind == "noise+dc": x = x * 0.05 + 0.9
    if kind == "tone5": x = np.sin(2 * np.pi * 5 * np.arange(T) / fs) + 0.01 * x
    x = (x - x.mean()); x = (x / np.abs(x).max()).astype(np.float32).astype(np.float64)
    for name, bands in tests:
        r = ref(x, bands); pk = np.abs(r - r.mean()).max()
        e64 = np.abs(cascade(x, bands, np.float64) - r).max() / pk
        e32 = np.abs(cascade(x, bands, np.float32) - r).max() / pk
        print(f"{kind:9s} {name:12s} col err / peak: fp64 pass1 {e64:.2e}   fp32 pass1 {e32:.2e}   -> output {e32/50:.1e}")
