"""Host study (task 1c): the fused kernel's band-limiting scan with the FIRST pass (zero-state end state of a 36-sample
chunk) in float32 and everything else in float64, against scipy.signal.lfilter in float64, for the PCG / ECG bands at 4125 Hz and
16 kHz on noise, offset, tonal and slowly drifting rows.  Prints the error relative to the filtered signal's peak (the normaliser
divides by that peak, so this is the error of the output).  Run: python tools/filter_precision_study.py   (CPU, ~2 min)"""
import numpy as np, scipy.signal as sig
rng = np.random.default_rng(0)
L = 36
def group_mats(secs):
    (b0, a0), (b1, a1) = secs
    A = np.zeros((4, 4)); B = np.zeros(4)
    A[0] = [-a0[1], 1, 0, 0]; B[0] = b0[1] - a0[1] * b0[0]
    A[1] = [-a0[2], 0, 0, 0]; B[1] = b0[2] - a0[2] * b0[0]
    A[2] = [b1[1] - a1[1] * b1[0], 0, -a1[1], 1]; B[2] = (b1[1] - a1[1] * b1[0]) * b0[0]
    A[3] = [b1[2] - a1[2] * b1[0], 0, -a1[2], 0]; B[3] = (b1[2] - a1[2] * b1[0]) * b0[0]
    return A, B
def run(x, secs, p1_dtype):
    A, B = group_mats(secs)
    n = len(x); nch = -(-n // L); xp = np.zeros(nch * L); xp[:n] = x
    X = xp.reshape(nch, L)
    W = np.zeros((L, 4)); v = B.copy()
    for j in range(L - 1, -1, -1):
        W[j] = v; v = A @ v
    if p1_dtype == np.float32:
        P = np.zeros((nch, 4), np.float32); Wf = W.astype(np.float32); Xf = X.astype(np.float32)
        for j in range(L):
            P = (Wf[j][None, :].astype(np.float64) * Xf[:, j:j + 1].astype(np.float64) + P.astype(np.float64)).astype(np.float32)
        P = P.astype(np.float64)
    else:
        P = X @ W
    M = np.linalg.matrix_power(A, L)
    Z = np.zeros((nch, 4)); z = np.zeros(4)
    for k in range(nch):
        Z[k] = z; z = M @ z + P[k]
    (b0, a0), (b1, a1) = secs
    z0, z1, z2, z3 = Z[:, 0].copy(), Z[:, 1].copy(), Z[:, 2].copy(), Z[:, 3].copy()
    Y = np.zeros_like(X)
    for j in range(L):
        xv = X[:, j]
        y0 = b0[0] * xv + z0; z0 = b0[1] * xv - a0[1] * y0 + z1; z1 = b0[2] * xv - a0[2] * y0
        y1 = b1[0] * y0 + z2; z2 = b1[1] * y0 - a1[1] * y1 + z3; z3 = b1[2] * y0 - a1[2] * y1
        Y[:, j] = y1
    return Y.reshape(-1)[:n]
for fs, band, name in ((4125.0, (25.0, 450.0), "PCG 4125"), (4125.0, (2.0, 40.0), "ECG 4125"), (16000.0, (25.0, 450.0), "PCG 16k"), (16000.0, (2.0, 40.0), "ECG 16k")):
    lo, hi = band
    secs = [sig.butter(2, hi / fs, btype="lowpass"), sig.butter(2, lo / fs, btype="highpass")]   # the reference's Wn = cutoff / fs
    for kind in ("noise", "noise+dc", "tone", "lowtone+dc"):
        T = int(30 * fs)
        x = rng.standard_normal(T)
        t = np.arange(T) / fs
        if kind == "noise+dc": x = 0.05 * x + 3.0
        if kind == "tone": x = np.sin(2 * np.pi * 60 * t) + 0.01 * x
        if kind == "lowtone+dc": x = 5.0 + np.sin(2 * np.pi * 0.3 * t) + 0.001 * x
        x = x.astype(np.float32).astype(np.float64)
        r = sig.lfilter(*secs[1], sig.lfilter(*secs[0], x))
        pk = np.abs(r).max()
        e64 = np.abs(run(x, secs, np.float64) - r).max() / pk
        e32 = np.abs(run(x, secs, np.float32) - r).max() / pk
        print(f"{name:9s} {kind:11s} err / peak: fp64 pass 1 {e64:.1e}   fp32 pass 1 {e32:.1e}")
