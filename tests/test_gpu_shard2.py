"""BASELINE.json configs[3] / SURVEY.md section 8e: ONE encode stream's lookahead split over two GPUs (search fields and
evaluations by step_units(); the list-0 MV fields of the newest frame exchanged either through the peer mailboxes --
kernels storing into the other GPU's HBM, flags polled on the device -- or through an NCCL all-gather) must give exactly
the costs of the single-GPU window, step by step.  Needs two GPUs (skipped otherwise): gpurun --gpus 2."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "x264-i386pic_b200"))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, W, H, B, nframes, passes, pipelined, exchange, sync_every, q):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import shard
    import x264cu
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from checkers import synth_frames
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    frames = [np.ascontiguousarray(f) for f in synth_frames(W, H, nframes, seed=51, pan=(3, 1))]
    ctx = x264cu.Context(W, H, bframes=B, passes=passes, device=rank)
    peer = shard.PeerExchange(x264cu, ctx, B, rank, world, dist) if exchange == "p2p" else None
    win = shard.ShardedWindow(x264cu, ctx, B, rank, world, dist=dist, device=torch.device("cuda", rank), peer=peer)
    mine = []
    for i, f in enumerate(frames):
        nxt = frames[i + 1].ctypes.data if pipelined and i + 1 < len(frames) else None
        win.push(f.ctypes.data, W, False, next_luma=nxt)      # pipelined: frame i+1 is uploaded and searched ahead
        if sync_every or i == len(frames) - 1:
            win.sync()
            got = win.costs()
            if not sync_every:
                # the window's host-side stats are per (d0, d1): without a sync per step, results of older steps that
                # this rank does not own in the newest step are still there -- keep the newest step's own units
                own = {(b - p0, i - b) for (p0, _, b) in shard.step_units(i, B, rank, world)[2]}
                got = {k: v for k, v in got.items() if k in own}
            mine.append(got)
        else:
            mine.append(None)                                 # free-running: only the device-side flags order the ranks
    win.close()
    if peer is not None:
        peer.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        # the same stream on one GPU
        L = ctx.L
        ref = L.x264cu_stream_open(ctx.ctx, B, 0)
        single = []
        st = (C.c_int32 * 4)()
        for f in frames:
            assert L.x264cu_stream_push(ref, f.ctypes.data, W, None, 0) >= 0
            x264cu.check(L.x264cu_stream_sync(ref), "sync")
            d = {}
            for d0 in range(B + 2):
                for d1 in range(B + 2 - d0):
                    L.x264cu_stream_cost(ref, d0, d1, st)
                    if st[0] >= 0:
                        d[(d0, d1)] = list(st)
            single.append(d)
        L.x264cu_stream_close(ref)
        q.put((gathered, single))
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("passes,pipelined,exchange,sync_every,n,world", [
    (2, False, "nccl", True, 14, 2), (0, False, "nccl", True, 14, 2), (2, True, "nccl", True, 14, 2),
    (2, False, "p2p", True, 14, 2), (0, True, "p2p", True, 14, 2),
    (2, True, "p2p", False, 45, 2),      # free-running: the 4-deep mailboxes and the frame ring wrap several times
    (2, True, "p2p", True, 14, 3),       # uneven shares: 6 list-0 fields over 3 ranks, ranks that own nothing in early steps
    (2, True, "p2p", False, 45, 4),
    (0, True, "nccl", True, 10, 4),
])
def test_one_stream_on_two_gpus_equals_one_gpu(passes, pipelined, exchange, sync_every, n, world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    W, H, B = 320, 192, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300 + passes + 5 * pipelined + 11 * (exchange == "p2p") + n + 50 * world
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, H, B, n, passes, pipelined, exchange, sync_every, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        gathered, single = q.get(timeout=300)
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    for step in range(n):
        if gathered[0][step] is None:
            continue
        merged = {}
        for r in range(world):
            for k, v in gathered[r][step].items():
                assert k not in merged, (step, k)          # every evaluation on exactly one rank
                merged[k] = v
        assert merged == single[step], step
    # both ranks had work
    assert all(sum(len(gathered[r][s] or ()) for s in range(n)) > 0 for r in range(world))


@pytest.mark.parametrize("passes", [2, 0])
def test_one_stream_two_ranks_on_one_gpu(passes):
    """The same sharding with both ranks as contexts of this process on ONE device (x264cu_peer_connect_local), so that the
    peer push / import kernels and the partition run on a single-GPU box as well.  Kernels that wait on one another must
    not be run as separate launches on one GPU (B200_PROFILING.md), so the two phases are ordered through the host here:
    every rank's searches + mailbox stores are complete before any rank's import is launched -- the import kernels find
    their flags already published.  Costs must equal the one-context window's, step by step."""
    import ctypes as C
    import shard
    import x264cu
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from checkers import synth_frames
    W, H, B, n, world = 320, 192, 5, 12, 2
    frames = [np.ascontiguousarray(f) for f in synth_frames(W, H, n, seed=51, pan=(3, 1))]
    ctxs = [x264cu.Context(W, H, bframes=B, passes=passes) for _ in range(world)]
    peers = shard.PeerExchange.local(x264cu, ctxs, B)
    wins = [shard.ShardedWindow(x264cu, ctxs[r], B, r, world, peer=peers[r]) for r in range(world)]
    ref_ctx = x264cu.Context(W, H, bframes=B, passes=passes)
    L = ref_ctx.L
    ref = L.x264cu_stream_open(ref_ctx.ctx, B, 0)
    st = (C.c_int32 * 4)()
    try:
        launches0 = sum(c.launches() for c in ctxs)
        for i, f in enumerate(frames):
            for r in range(world):
                wins[r]._phase0(i, f.ctypes.data, W, False)      # upload, this rank's searches, stores into the peer's mailbox
                wins[r].ahead = 1
            for r in range(world):
                ctxs[r].flush()                                   # flags are published before any import kernel starts
            merged = {}
            for r in range(world):
                wins[r].push(f.ctypes.data, W, False)             # import (no waiting) + this rank's evaluations
                wins[r].sync()
                own = {(b - p0, i - b) for (p0, _, b) in shard.step_units(i, B, r, world)[2]}
                for k, v in wins[r].costs().items():
                    if k in own:
                        assert k not in merged
                        merged[k] = v
            assert L.x264cu_stream_push(ref, f.ctypes.data, W, None, 0) >= 0
            x264cu.check(L.x264cu_stream_sync(ref), "sync")
            single = {}
            for d0 in range(B + 2):
                for d1 in range(B + 2 - d0):
                    L.x264cu_stream_cost(ref, d0, d1, st)
                    if st[0] >= 0:
                        single[(d0, d1)] = list(st)
            assert merged == single, i
        assert sum(c.launches() for c in ctxs) - launches0 >= 4 * n      # incl. one push and one import launch per rank and step
    finally:
        L.x264cu_stream_close(ref)
        ref_ctx.close()
        for w in wins:
            w.close()
        for p in peers:
            p.close()
        for c in ctxs:
            c.close()
