"""Small CPU helpers shared by the tests (not product code)."""
import numpy as np


def dense_frames_apply(x, g, up, down, off, t_out):
    """NumPy statement of the kernel's contract: y[i*up+p] = sum_d x[i*down+off+d] * g[p,d]."""
    x = np.asarray(x, dtype=np.float64)
    depth = g.shape[1]
    frames = -(-t_out // up)
    lo = max(0, -off)
    hi = max(0, (frames - 1) * down + off + depth - x.shape[-1])
    xp = np.concatenate([np.zeros(lo), x, np.zeros(hi)])
    idx = (np.arange(frames) * down + off + lo)[:, None] + np.arange(depth)[None]
    y = xp[idx] @ g.T                                   # [frames, up]
    return y.reshape(-1)[:t_out]


def rel_err(a, b):
    """max |a-b| relative to the reference's own scale max|b| (outputs here are O(1) signals)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))
