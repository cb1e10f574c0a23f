"""Slicetype decisions: the host decision logic of libx264cuda.so (x264cu_slicetype_analyse, a restatement of
slicetype.c:1288-1743) against the UNMODIFIED reference x264_slicetype_analyse.  Both are fed the same costs:
ours through the cost callback, the reference by pre-populating its i_cost_est caches (every
slicetype_frame_cost becomes a cache hit), so any difference is a difference in decision logic.
The GPU test then checks that costs from the device produce the same decisions as costs from the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from checkers import GET_STATS, Lookahead, OracleLib, RefLib, have_ref, synth_frames

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "x264-i386pic_b200"))


class StCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("bframes", "b_adapt", "bframe_pyramid", "scenecut_threshold", "keyint_max",
                                       "keyint_min", "open_gop", "intra_refresh", "psy_mbtree")]


class StFrame(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("frame_num", "type", "forced_type", "b_scenecut", "analysed_type")]


COST_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32))


def make_sequence(W, H, n, seed):
    """panning scene, a hard cut, a few repeated frames (cheap B frames) and a flash frame."""
    a = synth_frames(W, H, n, seed=seed, pan=(2, 1), scene_cut_at=n // 2)
    a[5] = a[4].copy()
    a[6] = a[4].copy()
    if n > 14:
        a[n // 2 + 4] = (255 - a[n // 2 + 4]).astype(np.uint8)   # a one-frame flash
    return a


def analyse_ours(lib, cfg, first, n, last_keyframe, cost_cb, intra_minigop=0, forced=None):
    frames = (StFrame * n)()
    for i in range(n):
        ft = 0 if forced is None else int(forced[i])
        frames[i] = StFrame(first + i, ft, ft, 1, 0)
    frames[0].type = 2
    reset = lib.x264cu_slicetype_analyse(C.byref(cfg), frames, n - 1, last_keyframe, intra_minigop, cost_cb, None)
    assert reset >= 0
    return [frames[i].type for i in range(n)], [frames[i].analysed_type for i in range(n)], reset


def all_triples(first, n, bframes):
    out = []
    for p0 in range(first, first + n):
        for p1 in range(p0, min(first + n, p0 + bframes + 2)):
            for b in range(p0, p1 + 1):
                if b == p1 or b > p0:
                    out.append((p0, p1, b))
    return out


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libx264ref.so not built")
@pytest.mark.parametrize("bframes,b_adapt,seed", [(3, 2, 1), (3, 1, 2), (3, 0, 3), (5, 2, 4), (2, 2, 5), (0, 0, 6), (4, 1, 7)])
def test_decisions_match_reference_logic(bframes, b_adapt, seed):
    import x264cu
    L = x264cu.load()
    L.x264cu_slicetype_analyse.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, COST_FN, C.c_void_p]
    W, H, n = 96, 64, 18
    frames = make_sequence(W, H, n, seed)
    ref, orc = RefLib(), OracleLib()
    ref.lib.ref_la_param.argtypes = [C.c_void_p, C.c_void_p]
    ref.lib.ref_la_inject_cost.argtypes = [C.c_void_p] + [C.c_int] * 6
    ref.lib.ref_la_analyse2.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]

    def new_ref():
        return Lookahead(ref, "ref", W, H, bframes=bframes, mbtree=0, b_adapt=b_adapt, rc_lookahead=20)
    R = new_ref()
    par = (C.c_int * 11)()
    ref.lib.ref_la_param(R.ctx, par)
    mv_range = R.mv_range
    R.close()
    cfg = StCfg(*list(par)[:9])
    assert cfg.bframes == bframes and cfg.b_adapt == b_adapt
    O = Lookahead(orc, "oracle", W, H, bframes=bframes, mbtree=0, mv_range=mv_range)
    try:
        for i, f in enumerate(frames):
            O.put_frame(i, f)
        # the cost matrix, from the oracle
        costs = {}
        for (p0, p1, b) in all_triples(0, n, bframes):
            sc = O.frame_cost(p0, p1, b)
            st = O.get(b, GET_STATS, b - p0, p1 - b)
            costs[(p0, p1, b)] = (sc, int(st[1]), int(st[2]))
        asked = []
        base = [0]

        def cb(user, p0, p1, b, st):                 # indices are relative to the window: frames[0] = last non-B
            k = (base[0] + p0, base[0] + p1, base[0] + b)
            asked.append(k)
            c = costs[k]
            st[0], st[1], st[2], st[3] = c[0], c[1], c[2], c[0]
            return c[0]
        cost_cb = COST_FN(cb)
        # slide the window like the lookahead does: analyse, take the decided frames, continue after the next non-B
        first, decided_ours, last_key = 0, [], 0
        while first < n - 2:
            cnt = min(n - first, 12, par[9] + 2)      # last non-B + at most i_slicetype_length+1 lookahead frames
            # the reference on a fresh encoder holding frames first.. at indices 0.., its caches pre-populated
            R = new_ref()
            for i in range(cnt):
                R.put_frame(i, frames[first + i])
            for (p0, p1, b) in all_triples(first, cnt, bframes):
                c = costs[(p0, p1, b)]
                ref.lib.ref_la_inject_cost(R.ctx, b - first, b - p0, p1 - b, c[0], c[1], c[2])
            tr = np.zeros(cnt, dtype=np.uint8)
            # i_frame of the shim's frames starts at 0, so the keyframe position is passed window-relative
            assert ref.lib.ref_la_analyse2(R.ctx, 0, cnt, 0, last_key - first, None, tr.ctypes.data) == 0
            R.close()
            base[0] = first
            ours, analysed, reset = analyse_ours(L, cfg, first, cnt, last_key, cost_cb)
            assert list(tr[1:]) == ours[1:], (first, list(tr), ours)
            j = 1
            while j < cnt - 1 and ours[j] in (4, 5):
                j += 1
            if ours[j] in (1, 2):
                last_key = first + j
            decided_ours += ours[1:j + 1]
            first += j
        assert len(asked) > 0
        if bframes and b_adapt == 0:
            assert any(t == 5 for t in decided_ours)
    finally:
        O.close()


def test_fixed_pattern_without_costs():
    """b_adapt 0 places B frames without consulting costs except for scenecut (slicetype.c:1623-1645)."""
    import x264cu
    L = x264cu.load()
    L.x264cu_slicetype_analyse.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, COST_FN, C.c_void_p]
    cfg = StCfg(2, 0, 0, 0, 250, 25, 0, 0, 0)
    cb = COST_FN(lambda u, p0, p1, b, st: 1000)
    types, analysed, reset = analyse_ours(L, cfg, 0, 8, 0, cb)
    assert analysed[1:] == [5, 5, 3, 5, 5, 3, 3]
    assert reset == 4 and types[1:4] == [5, 5, 3] and types[4:] == [0, 0, 0, 0]


@pytest.mark.gpu
@pytest.mark.parametrize("bframes,b_adapt", [(3, 2), (5, 1)])
def test_decisions_from_device_costs(bframes, b_adapt):
    """costs from the CUDA path drive the same decisions as costs from the oracle."""
    import x264cu
    L = x264cu.load()
    L.x264cu_slicetype_analyse.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, COST_FN, C.c_void_p]
    W, H, n = 176, 144, 16
    frames = make_sequence(W, H, n, 11)
    mvr = x264cu.default_mv_range(W, H)
    O = Lookahead(OracleLib(), "oracle", W, H, bframes=bframes, mv_range=mvr)
    G = x264cu.CudaLookahead(W, H, bframes=bframes, mv_range=mvr)
    try:
        for i, f in enumerate(frames):
            O.put_frame(i, f)
            G.put_frame(i, f)
        cfg = StCfg(bframes, b_adapt, 1, 40, 250, 25, 0, 0, 1)
        res = {}
        for name, la in (("oracle", O), ("cuda", G)):
            def cb(user, p0, p1, b, st, la=la):
                return la.frame_cost(p0, p1, b)
            res[name] = analyse_ours(L, cfg, 0, 13, 0, COST_FN(cb))
        assert res["oracle"] == res["cuda"]
        assert set(res["cuda"][1][1:]) <= {1, 2, 3, 5}
    finally:
        O.close()
        G.close()
