"""Kernel-only timing of the fused augmentation chain (mpcg_aug_chain_f32) at config 3 shape, all draws prepared
up front; MPCG_AC_CLUSTER=1/2/4/8 forces the cluster size."""
import sys, json, os, torch, numpy as np
sys.path.insert(0, ".")
from wav2vec_heart_sounds_b200 import torchaug as ta, AugmentConfig, _lib, design
PEAK = 6532.2
B, T, FS = int(os.environ.get("ROWS", 4096)), int(os.environ.get("T", 64000)), 16000
torch.manual_seed(0); np.random.seed(0)
x = torch.randn(B, T, device="cuda")
cfg = AugmentConfig()
rowp1, _ = ta._draw_noise(x, None, None, "philox"); m1 = ta._mask(B, cfg.prob_noise / 4, "cuda").reshape(B)
rowp2 = ta._draw_sines(x, 0.24); m2 = ta._mask(B, cfg.prob_wandering_volume, "cuda").reshape(B)
bands = ta._draw_bands(2, 500, 5); m3 = ta._mask(B, cfg.prob_banding, "cuda").reshape(B)
rowp4, _ = ta._draw_noise(x, None, None, "philox"); m4 = ta._mask(B, cfg.prob_noise / 4, "cuda").reshape(B)
sos = np.ascontiguousarray(design.eq_band_sos(FS, bands), dtype=np.float64)
out = torch.empty_like(x)
work = _lib.aug_workspace(x)
def run(masks=(m1, m2, m3, m4)):
    rc = _lib.lib().mpcg_aug_chain_f32(x.data_ptr(), out.data_ptr(), B, T, float(FS), rowp1.data_ptr(), None, masks[0].data_ptr(),
                                       1, 2, rowp2.data_ptr(), masks[1].data_ptr(), sos.ctypes.data, sos.shape[0],
                                       masks[2].data_ptr(), rowp4.data_ptr(), None, masks[3].data_ptr(), 3, 4,
                                       int(os.environ.get("COLLAPSE", 1)), work.data_ptr(), work.numel(),
                                       torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
def timeit(name, fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    nb = 2 * B * T * 4
    print(json.dumps({"op": name, "cluster": os.environ.get("MPCG_AC_CLUSTER", "auto"), "collapse": int(os.environ.get("COLLAPSE", 1)), "ms": round(best, 3),
                      "GB/s": round(nb / best / 1e6, 1), "frac_of_measured_peak": round(nb / best / 1e6 / PEAK, 3)}))
timeit("fused chain, default masks (p = .075/.75/.25/.075)", run)
ones, zeros = torch.ones(B, device="cuda"), torch.zeros(B, device="cuda")
timeit("fused chain, all masks off", lambda: run((zeros, zeros, zeros, zeros)))
timeit("fused chain, all masks on", lambda: run((ones, ones, ones, ones)))
timeit("only noise 1 on", lambda: run((ones, zeros, zeros, zeros)))
timeit("only wandering volume on", lambda: run((zeros, ones, zeros, zeros)))
timeit("only EQ on", lambda: run((zeros, zeros, ones, zeros)))
timeit("only noise 2 on", lambda: run((zeros, zeros, zeros, ones)))
