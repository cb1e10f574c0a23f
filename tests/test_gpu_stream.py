"""GPU: the lookahead-window driver (x264cu_stream_*, host/stream.cpp), the per-field entry point
(x264cu_motionsearch), pipelined waits, empty batches and maximum-size (4K) frames, against the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from checkers import GET_LOWRES_COSTS, GET_MV_COSTS, GET_MVS, GET_ROW_SATDS, GET_STATS, Lookahead, OracleLib, synth_frames

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "x264-i386pic_b200"))
pytestmark = pytest.mark.gpu
vp = C.c_void_p


@pytest.mark.parametrize("bframes,want_arrays", [(3, 0), (2, 1), (5, 1)])
def test_stream_cost_matrix_matches_oracle(bframes, want_arrays):
    """every push must produce exactly slicetype_frame_cost(p0, n, b) for n-(bframes+1) <= p0 < b <= n, evaluated in the
    order P first, then the B frames, per p0 from near to far (the order bench.py's CPU arm uses as well)."""
    import x264cu
    W, H, n = 208, 160, 10
    frames = synth_frames(W, H, n, seed=31, pan=(2, 2))
    mvr = x264cu.default_mv_range(W, H)
    O = Lookahead(OracleLib(), "oracle", W, H, bframes=bframes, mv_range=mvr)
    ctx = x264cu.Context(W, H, bframes=bframes, mv_range=mvr)
    L = ctx.L
    win = L.x264cu_stream_open(ctx.ctx, bframes, want_arrays)
    assert win
    try:
        st = (C.c_int32 * 4)()
        for k in range(n):
            f = np.ascontiguousarray(frames[k])
            O.put_frame(k, f)
            nev = L.x264cu_stream_push(win, f.ctypes.data, W, None, 0)
            assert nev > 0, x264cu.last_error()
            assert L.x264cu_stream_sync(win) == 0
            # the oracle in the same order
            expect = {(0, 0): O.frame_cost(k, k, k)}
            for p0 in range(k - 1, max(k - (bframes + 1), 0) - 1, -1):
                expect[(k - p0, 0)] = O.frame_cost(p0, k, k)
                for b in range(p0 + 1, k):
                    expect[(b - p0, k - b)] = O.frame_cost(p0, k, b)
            assert nev == len(expect)
            for (d0, d1), sc in expect.items():
                assert L.x264cu_stream_cost(win, d0, d1, st) == 0
                assert st[0] == sc, (k, d0, d1, list(st), sc)
                ost = O.get(k - d1, GET_STATS, d0, d1)
                assert st[1] == ost[1]
    finally:
        L.x264cu_stream_close(win)
        ctx.close()
        O.close()


def test_motionsearch_entry_point_and_pipelined_wait():
    import x264cu
    W, H = 208, 160
    frames = synth_frames(W, H, 4, seed=8)
    mvr = x264cu.default_mv_range(W, H)
    O = Lookahead(OracleLib(), "oracle", W, H, bframes=3, mv_range=mvr)
    ctx = x264cu.Context(W, H, bframes=3, mv_range=mvr)
    L = ctx.L
    try:
        fr = [ctx.frame_new() for _ in range(4)]
        for i, f in enumerate(frames):
            O.put_frame(i, f)
            ctx.upload(fr[i], np.ascontiguousarray(f))
        # several batches in flight, waited for one by one (x264cu_wait), out of a single final flush
        mv = [np.zeros((ctx.nmb, 2), dtype=np.int16) for _ in range(3)]
        mc = [np.zeros(ctx.nmb, dtype=np.int32) for _ in range(3)]
        seqs = []
        for d in (1, 2, 3):
            x264cu.check(L.x264cu_motionsearch(ctx.ctx, fr[3], fr[3 - d], 0, d, None, mv[d - 1].ctypes.data, mc[d - 1].ctypes.data), "motionsearch")
            seqs.append(L.x264cu_last_batch(ctx.ctx))
        assert seqs == sorted(seqs) and len(set(seqs)) == 3
        for d in (1, 2, 3):
            x264cu.check(L.x264cu_wait(ctx.ctx, seqs[d - 1]), "wait")
            O.frame_cost(3 - d, 3, 3)
            np.testing.assert_array_equal(mv[d - 1], O.get(3, GET_MVS, 0, d - 1))
            np.testing.assert_array_equal(mc[d - 1], O.get(3, GET_MV_COSTS, 0, d - 1))
        # the finalisation then finds the searches cached
        l0 = ctx.launches()
        r = ctx.frame_cost(fr[3], fr[1], fr[3], 2, 0, want=("stats", "lowres_costs", "row_satds"))
        assert ctx.launches() - l0 <= 2          # decision + gather, no search
        assert int(r["stats"][0]) == O.frame_cost(1, 3, 3)
        np.testing.assert_array_equal(r["lowres_costs"], O.get(3, GET_LOWRES_COSTS, 2, 0))
        np.testing.assert_array_equal(r["row_satds"], O.get(3, GET_ROW_SATDS, 2, 0))
        # empty batch and empty primitive batches are no-ops
        assert L.x264cu_frame_cost_batch(ctx.ctx, None, 0) == 0
        assert ctx.pixel_cmp(0, 3, np.zeros(64, np.uint8), 8, np.zeros(64, np.uint8), 8, [], []).size == 0
        ctx.flush()
    finally:
        ctx.close()
        O.close()


def test_4k_maximum_size():
    """3840x2160 (configs[3]): plane pitch 2048, 32400 blocks; properties + a spot check of one evaluation's interior
    against the oracle run on a crop would not be exact (borders differ), so compare the full frame cost on a
    downsampled-in-time check: identical frames and a pure pan."""
    import x264cu
    W, H = 3840, 2160
    rng = np.random.default_rng(3)
    base = rng.integers(0, 256, (H // 8 + 4, W // 8 + 4)).astype(np.float32)
    tex = np.kron(base, np.ones((8, 8), dtype=np.float32))
    for ax in (0, 1):
        tex = (np.roll(tex, 1, ax) + 2 * tex + np.roll(tex, -1, ax)) / 4
    f0 = tex[8:8 + H, 8:8 + W].astype(np.uint8)
    f1 = tex[8:8 + H, 12:12 + W].astype(np.uint8)          # pan by 4 full-res px = 8 qpel in lowres
    G = x264cu.CudaLookahead(W, H, bframes=3)
    try:
        assert (G.mbw, G.mbh) == (240, 135)
        G.put_frame(0, f0); G.put_frame(1, f1); G.put_frame(2, f0)
        s = G.frame_cost(0, 2, 2)                           # identical content
        r = G.results[(2, 2, 0)]
        assert (r["mvs0"] == 0).all() and s == 4 * (G.mbw - 2) * (G.mbh - 2)
        G.frame_cost(0, 1, 1)
        r = G.results[(1, 1, 0)]
        mv = r["mvs0"].reshape(G.mbh, G.mbw, 2)[2:-2, 2:-2]
        assert np.mean((mv[..., 0] == 8) & (mv[..., 1] == 0)) > 0.9
        costs = (r["lowres_costs"] & 0x3FFF).astype(np.int64).reshape(G.mbh, G.mbw)
        assert costs.sum(axis=1).tolist() == r["row_satds"].tolist()
        sb = G.frame_cost(0, 2, 1)                          # B evaluation between two identical frames
        assert sb > 0 and ((G.results[(1, 1, 1)]["lowres_costs"] >> 14) != 0).all()
    finally:
        G.close()


def test_concurrent_contexts_from_host_threads():
    """64 concurrent encode streams = many contexts driven by many host threads (one lookahead thread per encoder,
    encoder/lookahead.c:223-252).  Four threads, one context each, both schedules: results must equal the oracle's."""
    import threading
    import x264cu
    W, H, n = 208, 160, 5
    mvr = x264cu.default_mv_range(W, H)
    triples = [(0, 0, 0), (0, 1, 1), (0, 4, 4), (0, 4, 2), (1, 4, 2), (2, 4, 3), (3, 4, 4)]
    jobs = [(seed, passes) for seed in (3, 4) for passes in (2, 0)]
    expect = {}
    for seed, passes in jobs:
        frames = synth_frames(W, H, n, seed=seed)
        cpu = Lookahead(OracleLib(), "oracle", W, H, bframes=3, mv_range=mvr, passes=passes)
        for i, f in enumerate(frames):
            cpu.put_frame(i, f)
        expect[(seed, passes)] = ([cpu.frame_cost(*t) for t in triples],
                                  [cpu.get(t[2], GET_LOWRES_COSTS, t[2] - t[0], t[1] - t[2]).copy() for t in triples if t[0] != t[1]])
        cpu.close()
    got, errors = {}, []

    def worker(seed, passes):
        try:
            frames = synth_frames(W, H, n, seed=seed)
            gpu = x264cu.CudaLookahead(W, H, bframes=3, mv_range=mvr, passes=passes)
            for rep in range(3):                       # keep the contexts busy at the same time
                for i, f in enumerate(frames):
                    gpu.put_frame(i, f)
                costs = gpu.frame_cost_many(triples)
            arrays = [gpu.results[(t[2], t[2] - t[0], t[1] - t[2])]["lowres_costs"].copy() for t in triples if t[0] != t[1]]
            got[(seed, passes)] = (costs, arrays)
            gpu.close()
        except Exception as e:                         # surfaced in the main thread
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=j) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for j in jobs:
        assert got[j][0] == expect[j][0], j
        for a, b in zip(got[j][1], expect[j][1]):
            assert (a == b).all(), j
