"""Slicetype decisions.  The decision logic (x264_slicetype_analyse, slicetype.c:1288-1743: scenecut, B-adapt
none/fast/trellis, keyframe limit) is NOT part of the product: it stays the reference's own host C and reaches the
device through the hook (tests/test_gpu_x264_cli.py runs it end to end).  These tests drive the UNMODIFIED reference
logic (oracle/_ref, ref_la_analyse2) with cost matrices injected into its i_cost_est caches, so that every
slicetype_frame_cost it asks for is a cache hit:
  CPU: the oracle's dependency-free costs (K = 2) against the reference's own raster-serial costs -- not
       bit-comparable (secondary sanity oracle), frame types must agree on >= 75 % of frames even at 6x4 macroblocks;
  GPU: costs from the CUDA path against costs from the oracle -- identical decisions, frame by frame."""
import ctypes as C

import numpy as np
import pytest

from checkers import GET_STATS, Lookahead, OracleLib, RefLib, have_ref, synth_frames


def make_sequence(W, H, n, seed):
    """panning scene, a hard cut, a few repeated frames (cheap B frames) and a flash frame."""
    a = synth_frames(W, H, n, seed=seed, pan=(2, 1), scene_cut_at=n // 2)
    a[5] = a[4].copy()
    a[6] = a[4].copy()
    if n > 14:
        a[n // 2 + 4] = (255 - a[n // 2 + 4]).astype(np.uint8)   # a one-frame flash
    return a


def all_triples(first, n, bframes):
    out = []
    for p0 in range(first, first + n):
        for p1 in range(p0, min(first + n, p0 + bframes + 2)):
            for b in range(p0, p1 + 1):
                if b == p1 or b > p0:
                    out.append((p0, p1, b))
    return out


def cost_matrix(la, n, bframes):
    """every (p0, p1, b) the decision logic can ask for -> (cost_est, cost_est_aq, intra_mbs)"""
    costs = {}
    for (p0, p1, b) in all_triples(0, n, bframes):
        sc = la.frame_cost(p0, p1, b)
        st = la.results[(b, b - p0, p1 - b)]["stats"] if hasattr(la, "results") else la.get(b, GET_STATS, b - p0, p1 - b)
        # i_intra_mbs exists per P distance only (frame.h: i_intra_mbs[X264_BFRAME_MAX+2], written when b == p1)
        costs[(p0, p1, b)] = (int(sc), int(st[1]), int(st[2]) if b == p1 else 0)
    return costs


def decide_with_reference_logic(ref, frames, costs, W, H, bframes, b_adapt):
    """slide the window like the lookahead does (lookahead.c:68-129): analyse, take the decided frames, continue after
    the next non-B; the reference's caches are pre-populated from `costs`.  Returns the decided types of frames 1.."""
    L = ref.lib
    L.ref_la_param.argtypes = [C.c_void_p, C.c_void_p]
    L.ref_la_inject_cost.argtypes = [C.c_void_p] + [C.c_int] * 6
    L.ref_la_analyse2.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    n = len(frames)
    first, decided, last_key = 0, [], 0
    while first < n - 2:
        R = Lookahead(ref, "ref", W, H, bframes=bframes, mbtree=0, b_adapt=b_adapt, rc_lookahead=20)
        par = (C.c_int * 11)()
        L.ref_la_param(R.ctx, par)
        cnt = min(n - first, 12, par[9] + 2)      # last non-B + at most i_slicetype_length+1 lookahead frames
        for i in range(cnt):
            R.put_frame(i, frames[first + i])
        for (p0, p1, b) in all_triples(first, cnt, bframes):
            c = costs[(p0, p1, b)]
            L.ref_la_inject_cost(R.ctx, b - first, b - p0, p1 - b, c[0], c[1], c[2])
        tr = np.zeros(cnt, dtype=np.uint8)
        # i_frame of the shim's frames starts at 0, so the keyframe position is passed window-relative
        assert L.ref_la_analyse2(R.ctx, 0, cnt, 0, last_key - first, None, tr.ctypes.data) == 0
        R.close()
        types = [int(t) for t in tr]
        j = 1
        while j < cnt - 1 and types[j] in (4, 5):
            j += 1
        if types[j] in (1, 2):
            last_key = first + j
        decided += types[1:j + 1]
        first += j
    return decided


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libx264ref.so not built")
@pytest.mark.parametrize("bframes,b_adapt,seed", [(3, 2, 1), (3, 1, 2), (5, 2, 4)])
def test_dependency_free_costs_give_the_serial_decisions(bframes, b_adapt, seed):
    W, H, n = 96, 64, 18
    frames = make_sequence(W, H, n, seed)
    ref = RefLib()
    R = Lookahead(ref, "ref", W, H, bframes=bframes, mbtree=0)
    mv_range = R.mv_range
    R.close()
    # mbtree=1: the shim's frame pool is sized by rc.i_lookahead, which x264 zeroes without MB-tree / VBV
    S = Lookahead(ref, "ref_serial", W, H, bframes=bframes, mbtree=1)
    O = Lookahead(OracleLib(), "oracle", W, H, bframes=bframes, mbtree=1, mv_range=mv_range)
    try:
        for i, f in enumerate(frames):
            S.put_frame(i, f)
            O.put_frame(i, f)
        cs, co = cost_matrix(S, n, bframes), cost_matrix(O, n, bframes)
    finally:
        S.close()
        O.close()
    serial = decide_with_reference_logic(ref, frames, cs, W, H, bframes, b_adapt)
    free = decide_with_reference_logic(ref, frames, co, W, H, bframes, b_adapt)
    m = min(len(serial), len(free))
    same = sum(a == b for a, b in zip(serial[:m], free[:m]))
    assert m >= n - 4 and same >= 0.75 * m, (serial, free)
    assert set(free) <= {1, 2, 3, 4, 5}


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libx264ref.so not built")
@pytest.mark.parametrize("bframes,b_adapt", [(3, 2), (5, 1)])
def test_decisions_from_device_costs(bframes, b_adapt):
    """costs from the CUDA path drive the unmodified reference logic to the same decisions as costs from the oracle."""
    import x264cu
    W, H, n = 176, 144, 16
    frames = make_sequence(W, H, n, 11)
    mvr = x264cu.default_mv_range(W, H)
    ref = RefLib()
    O = Lookahead(OracleLib(), "oracle", W, H, bframes=bframes, mv_range=mvr)
    G = x264cu.CudaLookahead(W, H, bframes=bframes, mv_range=mvr)
    try:
        for i, f in enumerate(frames):
            O.put_frame(i, f)
            G.put_frame(i, f)
        co, cg = cost_matrix(O, n, bframes), cost_matrix(G, n, bframes)
        assert co == cg
        do = decide_with_reference_logic(ref, frames, co, W, H, bframes, b_adapt)
        dg = decide_with_reference_logic(ref, frames, cg, W, H, bframes, b_adapt)
        assert do == dg and set(dg) <= {1, 2, 3, 5}
    finally:
        O.close()
        G.close()
