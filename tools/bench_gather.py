"""Optional exchange of SURVEY 8e: all-gather of each rank's window tensor into every rank (NCCL over NVLink), outside
the data path.  Run under torchrun; checks the gathered tensor against the ranks' own seeds and times the collective
(CUDA events, max over ranks).  Shape: configs[1] windows of 1024 recordings per rank, [1024, 2, 7, 16500] fp32 = 946 MB."""
import json, os, sys, torch, torch.distributed as dist
sys.path.insert(0, ".")
from wav2vec_heart_sounds_b200 import shard
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
per = 1024
g = torch.Generator(device=dev).manual_seed(100 + rank)
mine = torch.rand(per, 2, 7, 16500, device=dev, generator=g)
out = shard.gather_windows(mine, per * world)
assert out.shape[0] == per * world and torch.equal(out[rank * per:(rank + 1) * per], mine)
chk = torch.Generator(device=dev).manual_seed(100 + (rank + 1) % world)
other = torch.rand(per, 2, 7, 16500, device=dev, generator=chk)
o = (rank + 1) % world
assert torch.equal(out[o * per:(o + 1) * per], other), "gathered block differs from its owner's data"
del out, other
best = 1e9
for _ in range(5):
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); res = shard.gather_windows(mine, per * world); b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    best = min(best, float(t))
    del res
if rank == 0:
    nbytes = mine.numel() * 4
    print(json.dumps({"op": "gather_windows (all_gather_into_tensor, NCCL)", "ranks": world, "bytes_per_rank": nbytes, "ms": round(best, 3),
                      "recv_GB/s_per_rank": round(nbytes * (world - 1) / best / 1e6, 1),
                      "bus_GB/s": round(nbytes * (world - 1) / best / 1e6, 1), "nvlink_peer_copy_reference_GB/s": 770}))
dist.destroy_process_group()
