"""One warm-up and one measured Hilbert-envelope call (for an ncu launch list): ROWS x T from argv."""
import sys, torch
sys.path.insert(0, ".")
from wav2vec_heart_sounds_b200 import envelopes as ev
rows, t = int(sys.argv[1]), int(sys.argv[2])
x = torch.randn(rows, t, device="cuda")
ev.hilbert_envelope(x); torch.cuda.synchronize()
ev.hilbert_envelope(x); torch.cuda.synchronize()
print("ok")
