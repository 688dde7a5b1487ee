"""Per-phase clock64 breakdown of the fused kernel (uses mpcg_debug_set_phase_clock_buffer).

The stamps are compiled in only with -DMPCG_FZ_PHASE_CLOCKS=1: build an instrumented library next to the product one
(tools/build_clocks.sh writes tools/libmpcg_b200_clocks.so) and run this script with MPCG_B200_LIB pointing at it."""
import sys, torch, numpy as np
sys.path.insert(0, ".")
import wav2vec_heart_sounds_b200 as pkg
from wav2vec_heart_sounds_b200 import _lib
from wav2vec_heart_sounds_b200.synth import synth_pair
spec = pkg.WindowSpec(4.0)
x = synth_pair(1024, 60000, 2000, seed=1234, device="cuda")
names = ["resample", "tables+zero", "despike", "pass1", "scan", "cluster carry", "pass2", "stats xchg", "store"]
for kinds in (("pcg", "ecg"),):
    out = pkg.preprocess_segment(x, 2000, 4125, spec, kinds=kinds, fused=True)
    ctas = 2048 * 8
    buf = torch.zeros(ctas, 16, dtype=torch.int64, device="cuda")
    _lib.lib().mpcg_debug_set_phase_clock_buffer(buf.data_ptr())
    pkg.preprocess_segment(x, 2000, 4125, spec, kinds=kinds, fused=True, out=out)
    torch.cuda.synchronize()
    _lib.lib().mpcg_debug_set_phase_clock_buffer(None)
    b = buf.cpu().numpy()
    used = b[:, 0] != 0
    b = b[used]
    ncl = b.shape[0] // 2048
    d = np.diff(b[:, :10], axis=1)
    rows = np.arange(b.shape[0]) // ncl
    for label, sel in (("PCG rows", rows % 2 == 0), ("ECG rows", rows % 2 == 1)):
        dd = d[sel]
        print(label, "ctas", dd.shape[0], "cluster", ncl, "total mean", dd.sum(1).mean().round(), "p95", np.percentile(dd.sum(1), 95).round())
        for i, nme in enumerate(names):
            print(f"   {nme:14s} mean {dd[:, i].mean():9.0f}  median {np.median(dd[:, i]):9.0f}  p95 {np.percentile(dd[:, i], 95):9.0f}  max {dd[:, i].max():9.0f}")
