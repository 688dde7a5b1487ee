"""world_size-2 (and 3) gloo runs of the sharding plumbing on the CPU: block bounds, per-rank processing and the
optional all-gather reproduce the unsharded result.  The per-row function here is a CPU stand-in (the CUDA
kernels need a GPU); what is under test is the host-side partitioning and the collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _row_fn(block):                       # stand-in for preprocess_segment: per-row, shape-changing
    b = block - block.mean(dim=-1, keepdim=True)
    return b.unfold(-1, 8, 4).contiguous()


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from wav2vec_heart_sounds_b200 import shard
        g = torch.Generator().manual_seed(0)
        x = torch.randn(n, 64, generator=g)
        lo, hi = shard.shard_bounds(n, rank, world)
        local = shard.sharded_apply(_row_fn, x)
        assert local.shape[0] == hi - lo
        full = shard.sharded_apply(_row_fn, x, gather=True)
        ok = torch.equal(full, _row_fn(x)) and torch.equal(local, _row_fn(x)[lo:hi])
        q.put((rank, lo, hi, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 10), (2, 7), (3, 8)])
def test_sharded_apply_matches_unsharded(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    assert all(ok for *_, ok in got)
    assert got[0][1] == 0 and got[-1][2] == n
    for a, b in zip(got, got[1:]):
        assert a[2] == b[1]                               # blocks tile the batch without gaps or overlap


def test_shard_bounds_cover_everything():
    from wav2vec_heart_sounds_b200 import shard
    for n in (0, 1, 5, 1024, 65536):
        for world in (1, 2, 4, 8):
            spans = [shard.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= -(-n // world)
    assert shard.shard_bounds(65536, 3, 8) == (24576, 32768)          # config 5: 8192 recordings per GPU
    with pytest.raises(ValueError):
        shard.shard_bounds(10, 2, 2)
