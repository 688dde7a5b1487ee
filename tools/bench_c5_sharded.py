"""configs[4]: 65 536 Vest-shaped recordings (6-channel PCG, 8 s at 4 kHz -> 4125 Hz, 2 s windows), preprocess + segment +
augment, sharded by recording over the ranks of one node (no collective on the data path).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
        tools/bench_c5_sharded.py [--total 65536] [--chunk 8192] [--reps 3]

A rank walks its shard in chunks of --chunk recordings (synthetic, generated on the device before the timed region; the
chunk buffers are reused).  Timed: CUDA events around the whole shard, barrier + synchronize on both sides, max over ranks.
Rank 0 prints one JSON line."""
import argparse, json, os, sys
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wav2vec_heart_sounds_b200 as pkg
from wav2vec_heart_sounds_b200 import torchaug as ta
from wav2vec_heart_sounds_b200.shard import shard_bounds
from wav2vec_heart_sounds_b200.synth import synth_pcg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total", type=int, default=65536)
    ap.add_argument("--chunk", type=int, default=8192)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = shard_bounds(a.total, rank, world)
    C, T, fs_in, fs = 6, 32000, 4000, 4125
    spec = pkg.WindowSpec(2.0)
    chunk = min(a.chunk, hi - lo)
    x = synth_pcg(chunk * C, T, float(fs_in), seed=5 + rank, device=dev).reshape(chunk, C, T)
    win = pkg.preprocess_segment(x, fs_in, fs, spec, fused=True)                 # [chunk, C, N, W]
    rows = win.reshape(-1, win.shape[-1])
    aug = torch.empty_like(rows)

    def shard_pass():
        done = lo
        while done < hi:
            n = min(chunk, hi - done)
            pkg.preprocess_segment(x[:n], fs_in, fs, spec, fused=True, out=win[:n])
            r = win[:n].reshape(-1, win.shape[-1])
            ta.augment_pcg_batch(r, fs, noise="philox", out=aug[: r.shape[0]])
            done += n

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    shard_pass(); shard_pass()
    best = None
    for _ in range(a.reps):
        sync()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); shard_pass(); e.record()
        sync()
        t = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = float(t) if best is None else min(best, float(t))
    if rank == 0:
        nwin = win.shape[2]
        bytes_rec = C * T * 4 + 3 * C * nwin * win.shape[-1] * 4               # raw in, windows out, windows in + out of the chain
        print(json.dumps({"workload": "configs[4]: %d recordings x 6 ch x 8 s @4 kHz -> 4125 Hz, 2 s windows, preprocess + segment + "
                                      "augment_pcg_batch, sharded by recording" % a.total,
                          "n_gpus": world, "recordings_per_gpu": hi - lo, "chunk": chunk, "windows_per_recording": nwin,
                          "ms": round(best, 3), "audio_s_per_s": round(a.total * 8.0 / best * 1e3),
                          "recordings_per_s": round(a.total / best * 1e3),
                          "algorithmic_GB/s_per_gpu": round((hi - lo) * bytes_rec / best / 1e6, 1), "scaling": "strong"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
