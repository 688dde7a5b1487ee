"""Per-phase cycle totals of the row-streaming fused kernel (uses mpcg_debug_set_phase_clock_buffer).

The counters are compiled in only with -DMPCG_FZ_PHASE_CLOCKS=1: build an instrumented library next to the product
one (tools/build_clocks.sh writes tools/libmpcg_b200_clocks.so) and run this script with MPCG_B200_LIB pointing at it.
Each persistent CTA adds up, over all the rows it processed, the cycles between its phase boundaries."""
import sys, torch, numpy as np
sys.path.insert(0, ".")
import wav2vec_heart_sounds_b200 as pkg
from wav2vec_heart_sounds_b200 import _lib
from wav2vec_heart_sounds_b200.synth import synth_pair
spec = pkg.WindowSpec(4.0)
x = synth_pair(1024, 60000, 2000, seed=1234, device="cuda")
names = ["row setup", "A: resample+maxima+park", "A: despike passes", "B: tile load/resample", "B: filter", "B: park un-normalised",
         "stats + C: windows"]
cases = {"pcg+ecg": (x, ("pcg", "ecg")), "pcg only": (x[:, :1].contiguous(), ("pcg",)), "ecg only": (x[:, 1:].contiguous(), ("ecg",))}
for label, (xx, kinds) in cases.items():
    out, edits = pkg.preprocess_segment(xx, 2000, 4125, spec, kinds=kinds, fused=True, return_edits=True)
    buf = torch.zeros(4096, 16, dtype=torch.int64, device="cuda")
    _lib.lib().mpcg_debug_set_phase_clock_buffer(buf.data_ptr())
    pkg.preprocess_segment(xx, 2000, 4125, spec, kinds=kinds, fused=True, out=out)
    torch.cuda.synchronize()
    _lib.lib().mpcg_debug_set_phase_clock_buffer(None)
    b = buf.cpu().numpy()[:, :7]
    b = b[b.sum(1) != 0]
    tot = b.sum(1)
    print(f"{label}: ctas {b.shape[0]}, cycles per CTA mean {tot.mean():.0f} min {tot.min():.0f} max {tot.max():.0f}")
    for i, nme in enumerate(names):
        print(f"   {nme:26s} mean {b[:, i].mean():10.0f}  ({100 * b[:, i].sum() / tot.sum():5.1f} %)  max {b[:, i].max():10.0f}")
    e = edits.cpu().numpy()
    print("   despike passes per row: mean %.2f  p50 %d  p90 %d  p99 %d  max %d  rows>0: %d of %d" %
          (e.mean(), np.percentile(e, 50), np.percentile(e, 90), np.percentile(e, 99), e.max(), (e > 0).sum(), e.size))
