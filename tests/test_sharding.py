"""CPU, world_size 2, gloo: the N>1 host logic of bench.py -- stream assignment is a disjoint cover,
the timing reduction is a max over ranks, and rank 0 alone reports."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "x264-i386pic_b200"))


def _worker(rank, world, port, q):
    import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.assign_streams(8, world, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    t = torch.tensor([10.0 + rank, 20.0 - rank], dtype=torch.float64)   # per-rank device times (value, e2e)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        q.put((gathered, t.tolist(), [shard.stream_seed(s) for s in mine]))
    dist.destroy_process_group()


def test_stream_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, tmax, seeds = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(gathered[0] + gathered[1]) == list(range(8))
    assert not set(gathered[0]) & set(gathered[1])
    assert tmax == [11.0, 20.0]
    assert seeds == [1, 15, 29, 43]


def test_field_sharding_covers_everything():
    import shard
    fields = [(b, b - j, 0) for b in range(1, 18) for j in range(1, min(b, 17) + 1)] + \
             [(b, b + j, 1) for b in range(0, 17) for j in range(1, 17 - b)]
    for world in (2, 4, 8):
        parts = shard.assign_fields(fields, world)
        flat = [f for r in range(world) for f in parts[r]]
        assert sorted(flat) == sorted(fields)
        sizes = [len(parts[r]) for r in range(world)]
        assert max(sizes) - min(sizes) <= len(fields) // world // 2 + 2


def test_one_stream_step_units_are_a_disjoint_cover():
    """BASELINE.json configs[3]: the step that frame n opens, split over `world` GPUs, covers every search field and every
    evaluation of the single-GPU step exactly once."""
    import shard
    B = 16
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 5, 17, 40):
            lo = max(0, n - (B + 1))
            want_l0 = sorted(range(lo, n))
            want_l1 = sorted(range(lo + 1, n))
            want_ev = sorted([(n, n, n)] + [(p0, n, n) for p0 in range(lo, n)] + [(p0, n, b) for p0 in range(lo, n) for b in range(p0 + 1, n)])
            l0, l1, ev = [], [], []
            for rank in range(world):
                a, b_, c = shard.step_units(n, B, rank, world)
                l0 += a; l1 += b_; ev += c
                # a B evaluation lives where its frame's list-1 search lives, a P evaluation where its list-0 search does
                assert all((p1 == b and (p0 == b or p0 % world == rank)) or (b % world == rank) for (p0, p1, b) in c)
            assert sorted(l0) == want_l0 and sorted(l1) == want_l1 and sorted(ev) == want_ev, (world, n)
            if n >= B + 1 and world <= 8:
                sizes = [len(shard.step_units(n, B, r, world)[2]) for r in range(world)]
                assert max(sizes) - min(sizes) <= 2 * (B + 1) // world + 2 + (B + 1)
