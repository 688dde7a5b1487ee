"""End-to-end (host buffers in/out) throughput of HostPipeline at configs[1] for several chunk sizes."""
import sys, time, json, torch
sys.path.insert(0, ".")
import wav2vec_heart_sounds_b200 as pkg
from wav2vec_heart_sounds_b200.pipeline import HostPipeline
from wav2vec_heart_sounds_b200.synth import synth_pair
x = synth_pair(1024, 60000, 2000, seed=1234, device="cuda")
xh = x.cpu().pin_memory()
spec = pkg.WindowSpec(4.0)
import os
from wav2vec_heart_sounds_b200 import AugmentConfig
aug = AugmentConfig() if os.environ.get("AUG") else None
for chunk in (32, 64, 128, 256):
    hp = HostPipeline(1024, 2, 60000, 2000, 4125, spec, kinds=("pcg", "ecg"), chunk=chunk, augment=aug)
    oh = hp.empty_output()
    for _ in range(2): hp(xh, oh)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        hp(xh, oh); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(json.dumps({"augment": aug is not None, "chunk": chunk, "ms": round(dt * 1e3, 2), "audio_s_per_s": round(1024 * 30 / dt), "GB/s_d2h": round(hp.d2h_bytes / dt / 1e9, 1)}))
    del hp, oh
