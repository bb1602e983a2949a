"""CPU, world_size 2 over gloo: clip sharding and the final gather (the only multi-GPU step of the path)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, n_clips, port, q):
    sys.path.insert(0, ROOT)
    from diff_sal_b200.parallel import gather_maps, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_clips, rank, world)
    # stand-in "denoiser": clip i -> map filled with i + 0.5
    local = torch.stack([torch.full((1, 4, 6), i + 0.5) for i in range(lo, hi)]) if hi > lo else torch.zeros(0, 1, 4, 6)
    full = gather_maps(local, n_clips)
    q.put((rank, lo, hi, full[:, 0, 0, 0].tolist()))
    dist.destroy_process_group()


def _worker_overlapped(rank, world, n_clips, port, q):
    sys.path.insert(0, ROOT)
    from diff_sal_b200.parallel import MapGatherer, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_clips, rank, world)
    g = MapGatherer(n_clips, (1, 4, 6), torch.float32, "cpu")
    got = []
    # three consecutive "batches" through the start / finish form, two gathers in flight at a time
    for step in range(3):
        local = (torch.stack([torch.full((1, 4, 6), 100.0 * step + i + 0.5) for i in range(lo, hi)]) if hi > lo
                 else torch.zeros(0, 1, 4, 6))
        g.release(step)                                    # slot reuse: the gather of step - 2 must be consumed first
        g.start(local, step)
        if step >= 1:
            got.append(g.finish(step - 1)[:, 0, 0, 0].tolist())
    got.append(g.finish(2)[:, 0, 0, 0].tolist())
    q.put((rank, got))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 5])
def test_overlapped_gather_world2(n_clips):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() + n_clips) % 90
    procs = [ctx.Process(target=_worker_overlapped, args=(r, 2, n_clips, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, got in res:
        assert got == [[100.0 * s + i + 0.5 for i in range(n_clips)] for s in range(3)]


@pytest.mark.parametrize("n_clips", [8, 5, 1])
def test_shard_and_gather_world2(n_clips):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_clips) % 400
    procs = [ctx.Process(target=_worker, args=(r, 2, n_clips, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = []
    for rank, lo, hi, vals in res:
        assert vals == [i + 0.5 for i in range(n_clips)]       # every rank sees all maps, in clip order
        covered += list(range(lo, hi))
    assert sorted(covered) == list(range(n_clips))             # each clip owned by exactly one rank


def test_shard_range_partitions():
    from diff_sal_b200.parallel import shard_range
    for n in (1, 7, 8, 256, 257):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
