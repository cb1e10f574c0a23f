"""GPU parity: the CUDA lookahead (through the C ABI) against the CPU oracle on the same frames.

Bit-exact on every output array: lowres planes incl. border, intra costs, MVs, MV costs, packed
lowres_costs, row SATDs and the frame statistics.  Cases mirror what slicetype_frame_cost can be
asked (I / P / B, several distances, cached and uncached searches, edges on/off, AQ on/off, both
metric flavours, ragged sizes, passes 1..3).
"""
import importlib.util
import os
import sys

import numpy as np
import pytest

from checkers import (GET_INTRA, GET_LOWRES_COSTS, GET_MV_COSTS, GET_MVS, GET_ROW_SATDS, GET_STATS, Lookahead,
                      OracleLib, RefLib, have_ref, synth_frames)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "x264-i386pic_b200"))

pytestmark = pytest.mark.gpu


def _cu():
    import x264cu
    return x264cu


TRIPLES = [(0, 0, 0), (0, 1, 1), (0, 2, 2), (0, 2, 1), (0, 4, 4), (0, 4, 2), (0, 4, 1), (0, 4, 3),
           (1, 1, 1), (1, 3, 2), (2, 5, 3), (2, 5, 4), (2, 5, 5), (1, 2, 2), (3, 4, 4)]


def compare_eval(gpu, cpu, p0, p1, b, tag):
    r = gpu.results[(b, b - p0, p1 - b)]
    i, j = b - p0, p1 - b
    np.testing.assert_array_equal(r["intra_costs"], cpu.get(b, GET_INTRA), err_msg=f"{tag} intra")
    np.testing.assert_array_equal(r["lowres_costs"], cpu.get(b, GET_LOWRES_COSTS, i, j), err_msg=f"{tag} lowres_costs")
    if i > 0:
        np.testing.assert_array_equal(r["mvs0"], cpu.get(b, GET_MVS, 0, i - 1), err_msg=f"{tag} mvs0")
        np.testing.assert_array_equal(r["mv_costs0"], cpu.get(b, GET_MV_COSTS, 0, i - 1), err_msg=f"{tag} mv_costs0")
    if j > 0:
        np.testing.assert_array_equal(r["mvs1"], cpu.get(b, GET_MVS, 1, j - 1), err_msg=f"{tag} mvs1")
        np.testing.assert_array_equal(r["mv_costs1"], cpu.get(b, GET_MV_COSTS, 1, j - 1), err_msg=f"{tag} mv_costs1")
    np.testing.assert_array_equal(r["row_satds"], cpu.get(b, GET_ROW_SATDS, i, j), err_msg=f"{tag} row_satds")
    np.testing.assert_array_equal(r["intra_row_satds"], cpu.get(b, GET_ROW_SATDS, 0, 0), err_msg=f"{tag} intra rows")
    st = cpu.get(b, GET_STATS, i, j)
    assert r["stats"][0] == st[0] and r["stats"][1] == st[1], f"{tag} stats {r['stats']} vs {st}"
    if j == 0:
        assert r["stats"][2] == st[2], f"{tag} intra_mbs {r['stats']} vs {st}"
    st0 = cpu.get(b, GET_STATS, 0, 0)
    assert r["intra_stats"][0] == st0[0] and r["intra_stats"][1] == st0[1], f"{tag} intra stats"


@pytest.mark.parametrize("W,H,subme,mbtree,aq,passes", [
    (352, 288, 7, 1, 1, 2),
    (352, 288, 7, 0, 1, 2),
    (352, 288, 1, 1, 1, 2),
    (352, 288, 1, 0, 0, 1),
    (360, 270, 7, 1, 0, 3),      # not a multiple of 16: mod-16 replication path
    (64, 48, 7, 1, 1, 2),
    (32, 32, 7, 0, 1, 2),        # tiny: 2x2 blocks, every block scores
])
def test_frame_costs_match_oracle(W, H, subme, mbtree, aq, passes):
    x264cu = _cu()
    frames = synth_frames(W, H, 6, seed=3)
    rng = np.random.default_rng(5)
    mvr = x264cu.default_mv_range(W, H)
    cpu = Lookahead(OracleLib(), "oracle", W, H, bframes=3, subme=subme, mbtree=mbtree, aq=aq, passes=passes, mv_range=mvr)
    gpu = x264cu.CudaLookahead(W, H, bframes=3, subme=subme, mbtree=mbtree, aq=aq, passes=passes, mv_range=mvr)
    try:
        for i, f in enumerate(frames):
            q = rng.integers(128, 512, cpu.nmb).astype(np.uint16) if aq else None
            cpu.put_frame(i, f, q)
            gpu.put_frame(i, f, q)
        for p in range(4):
            np.testing.assert_array_equal(gpu.lowres(2, p), cpu.lowres(2, p), err_msg=f"lowres plane {p}")
        for (p0, p1, b) in TRIPLES:
            sc = cpu.frame_cost(p0, p1, b)
            sg = gpu.frame_cost(p0, p1, b)
            compare_eval(gpu, cpu, p0, p1, b, f"({p0},{p1},{b})")
            assert sc == sg
    finally:
        cpu.close()
        gpu.close()


@pytest.mark.parametrize("lanes", [1, 8])
@pytest.mark.parametrize("W,H,subme,mbtree,aq", [(352, 288, 7, 1, 1), (360, 270, 1, 0, 1), (200, 120, 7, 0, 0), (32, 32, 7, 1, 1)])
def test_both_lane_mappings(monkeypatch, lanes, W, H, subme, mbtree, aq):
    """the thread-per-block kernels (wide launches) and the 8-lanes-per-block kernels (single
    evaluations) are two mappings of the same search: both must match the oracle bit for bit."""
    monkeypatch.setenv("X264CU_LANES", str(lanes))
    x264cu = _cu()
    frames = synth_frames(W, H, 6, seed=17, pan=(4, -1))
    rng = np.random.default_rng(6)
    mvr = x264cu.default_mv_range(W, H)
    cpu = Lookahead(OracleLib(), "oracle", W, H, bframes=3, subme=subme, mbtree=mbtree, aq=aq, mv_range=mvr)
    gpu = x264cu.CudaLookahead(W, H, bframes=3, subme=subme, mbtree=mbtree, aq=aq, mv_range=mvr)
    try:
        for i, f in enumerate(frames):
            q = rng.integers(128, 512, cpu.nmb).astype(np.uint16) if aq else None
            cpu.put_frame(i, f, q)
            gpu.put_frame(i, f, q)
        triples = [(0, 0, 0), (0, 1, 1), (0, 4, 4), (0, 4, 2), (0, 4, 1), (0, 4, 3), (2, 5, 3), (2, 5, 5), (3, 4, 4)]
        sg = gpu.frame_cost_many(triples)
        for (p0, p1, b), s in zip(triples, sg):
            assert cpu.frame_cost(p0, p1, b) == s
            compare_eval(gpu, cpu, p0, p1, b, f"lanes={lanes} ({p0},{p1},{b})")
    finally:
        cpu.close()
        gpu.close()


@pytest.mark.parametrize("W,H,subme,passes", [(352, 288, 7, 2), (360, 270, 1, 2), (1920, 1080, 7, 2), (352, 288, 7, 0)])
def test_runtime_pitch_kernels(monkeypatch, W, H, subme, passes):
    """frames wider than 4K run the instantiations with a run-time plane pitch (the compile-time-pitch ones cover
    pitch 1024 / 2048); X264CU_RUNTIME_PITCH forces that path at small sizes.  Weighted search included."""
    monkeypatch.setenv("X264CU_RUNTIME_PITCH", "1")
    monkeypatch.setenv("X264CU_LANES", "1")
    x264cu = _cu()
    frames = synth_frames(W, H, 4, seed=23, pan=(2, 1))
    mvr = x264cu.default_mv_range(W, H)
    cpu = Lookahead(OracleLib(), "oracle", W, H, bframes=3, subme=subme, mv_range=mvr, passes=passes)
    gpu = x264cu.CudaLookahead(W, H, bframes=3, subme=subme, mv_range=mvr, passes=passes)
    try:
        assert gpu.ctx.stride not in (1024, 2048)
        for i, f in enumerate(frames):
            cpu.put_frame(i, f)
            gpu.put_frame(i, f)
        w = (1, 60, 6, -3)
        assert gpu.frame_cost(0, 1, 1, weight=w[1:]) == cpu.frame_cost(0, 1, 1, weight=w)
        compare_eval(gpu, cpu, 0, 1, 1, "runtime pitch weighted (0,1,1)")
        triples = [(0, 3, 3), (0, 3, 1), (0, 3, 2), (1, 3, 2)]
        sg = gpu.frame_cost_many(triples)
        for (p0, p1, b), sgpu in zip(triples, sg):
            assert cpu.frame_cost(p0, p1, b) == sgpu
            compare_eval(gpu, cpu, p0, p1, b, f"runtime pitch ({p0},{p1},{b})")
    finally:
        cpu.close()
        gpu.close()


@pytest.mark.parametrize("W,H,bframes,subme,passes,pan", [
    (208, 160, 16, 7, 2, (5, -2)),     # the bench's bframes with long distances: MVs far from zero, temporal direct at every distance
    (208, 160, 16, 7, 0, (5, -2)),     # the same in the exact raster-serial mode
    (176, 144, 0, 7, 2, (2, 1)),       # no B frames at all
    (176, 144, 1, 1, 2, (-3, 2)),      # one B frame, SAD flavour
    (528, 80, 7, 7, 2, (9, 0)),        # wide and flat: 66 x 10 blocks, warp tiles 32+32+2
])
def test_bframe_depths_and_shapes(W, H, bframes, subme, passes, pan):
    """every (p0, p1, b) with p1 - p0 <= bframes + 1 that contains the newest frame, as the lookahead window issues them"""
    x264cu = _cu()
    n = min(bframes + 2, 8) if bframes < 16 else 18
    frames = synth_frames(W, H, n, seed=41, pan=pan)
    mvr = x264cu.default_mv_range(W, H)
    cpu = Lookahead(OracleLib(), "oracle", W, H, bframes=bframes, subme=subme, passes=passes, mv_range=mvr)
    gpu = x264cu.CudaLookahead(W, H, bframes=bframes, subme=subme, passes=passes, mv_range=mvr)
    try:
        for i, f in enumerate(frames):
            cpu.put_frame(i, f)
            gpu.put_frame(i, f)
        last = n - 1
        triples = [(last, last, last)]
        for p0 in range(last - 1, max(last - (bframes + 1), 0) - 1, -1):
            triples.append((p0, last, last))
            triples += [(p0, last, b) for b in range(p0 + 1, last)]
        if bframes == 16:
            triples = triples[:1] + triples[1::3]          # keep the run short: every third evaluation, all distances
        sg = gpu.frame_cost_many(triples)
        for (p0, p1, b), s_gpu in zip(triples, sg):
            assert cpu.frame_cost(p0, p1, b) == s_gpu, (p0, p1, b)
            compare_eval(gpu, cpu, p0, p1, b, f"bframes={bframes} ({p0},{p1},{b})")
    finally:
        cpu.close()
        gpu.close()


def test_batch_equals_sequential():
    """x264cu_frame_cost_batch == the same evaluations issued one by one (array order semantics)."""
    x264cu = _cu()
    W, H = 352, 288
    frames = synth_frames(W, H, 7, seed=11, pan=(-2, 3))
    mvr = x264cu.default_mv_range(W, H)
    cpu = Lookahead(OracleLib(), "oracle", W, H, bframes=4, mv_range=mvr)
    gpu = x264cu.CudaLookahead(W, H, bframes=4, mv_range=mvr)
    try:
        for i, f in enumerate(frames):
            cpu.put_frame(i, f)
            gpu.put_frame(i, f)
        triples = [(0, 5, 5), (0, 5, 1), (0, 5, 2), (0, 5, 3), (0, 5, 4), (1, 6, 3), (1, 6, 6), (1, 6, 2), (2, 2, 2), (2, 3, 3)]
        sg = gpu.frame_cost_many(triples)
        for (p0, p1, b), s in zip(triples, sg):
            assert cpu.frame_cost(p0, p1, b) == s
            compare_eval(gpu, cpu, p0, p1, b, f"batch ({p0},{p1},{b})")
    finally:
        cpu.close()
        gpu.close()


@pytest.mark.parametrize("lanes", [0, 1])
def test_weighted_p_search(monkeypatch, lanes):
    """explicit weighted prediction of list 0 (weights supplied by the host, slicetype.c:859-864); lanes = 1 forces the
    thread-per-block kernels with the compile-time pitch (weighted half-pel diamond with pairwise shared fetches)."""
    if lanes:
        monkeypatch.setenv("X264CU_LANES", str(lanes))
    x264cu = _cu()
    import ctypes as C
    W, H = 352, 288
    frames = synth_frames(W, H, 3, seed=5, fade_at=0)
    mvr = x264cu.default_mv_range(W, H)
    orc = OracleLib()
    cpu = Lookahead(orc, "oracle", W, H, bframes=3, mv_range=mvr)
    gpu = x264cu.CudaLookahead(W, H, bframes=3, mv_range=mvr)
    try:
        for i, f in enumerate(frames):
            cpu.put_frame(i, f)
            gpu.put_frame(i, f)
        for (p0, p1, wt) in [(0, 1, (58, 6, 3)), (0, 2, (120, 7, -2)), (1, 2, (70, 6, 0))]:
            sc = orc.lib.xo_frame_cost(cpu.ctx, p0, p1, p1, wt[0], wt[1], wt[2])
            sg = gpu.frame_cost(p0, p1, p1, weight=wt)
            assert sc == sg
            compare_eval(gpu, cpu, p0, p1, p1, f"weighted ({p0},{p1})")
    finally:
        cpu.close()
        gpu.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libx264ref.so not built")
def test_frame_costs_match_reference_build():
    """the CUDA path against the UNMODIFIED reference slicetype_mb_cost driven in the same schedule."""
    x264cu = _cu()
    W, H = 352, 288
    frames = synth_frames(W, H, 5, seed=9, pan=(5, -2))
    ref = Lookahead(RefLib(), "ref", W, H, bframes=3, passes=2)
    gpu = x264cu.CudaLookahead(W, H, bframes=3, passes=2, mv_range=ref.mv_range)
    try:
        for i, f in enumerate(frames):
            ref.put_frame(i, f)
            gpu.put_frame(i, f)
        for (p0, p1, b) in [(0, 0, 0), (0, 1, 1), (0, 3, 3), (0, 3, 1), (0, 3, 2), (1, 4, 4), (1, 4, 2)]:
            sr = ref.frame_cost(p0, p1, b)
            sg = gpu.frame_cost(p0, p1, b)
            assert sr == sg
            compare_eval(gpu, ref, p0, p1, b, f"ref ({p0},{p1},{b})")
    finally:
        ref.close()
        gpu.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libx264ref.so not built")
@pytest.mark.parametrize("W,H,subme,mbtree", [(352, 288, 7, 1), (208, 160, 7, 0), (352, 288, 1, 1)])
def test_exact_raster_mode_matches_unmodified_reference(W, H, subme, mbtree):
    """search_passes = 0 (one launch per anti-diagonal wave): bit-exact with the reference's OWN raster-serial
    slicetype_frame_cost (slicetype.c:836) -- no schedule change, nothing modified, one lookahead thread."""
    x264cu = _cu()
    frames = synth_frames(W, H, 5, seed=12, pan=(-4, 3), noise=5)
    rng = np.random.default_rng(2)
    ref = Lookahead(RefLib(), "ref_serial", W, H, bframes=3, subme=subme, mbtree=mbtree, aq=1)
    gpu = x264cu.CudaLookahead(W, H, bframes=3, subme=subme, mbtree=mbtree, aq=1, passes=0, mv_range=ref.mv_range)
    try:
        for i, f in enumerate(frames):
            q = rng.integers(64, 1024, ref.nmb).astype(np.uint16)
            ref.put_frame(i, f, q)
            gpu.put_frame(i, f, q)
        for (p0, p1, b) in [(0, 0, 0), (0, 1, 1), (0, 3, 3), (0, 3, 1), (0, 3, 2), (1, 4, 4), (1, 4, 2), (1, 4, 3), (2, 4, 4), (2, 4, 3)]:
            sr = ref.frame_cost(p0, p1, b)
            sg = gpu.frame_cost(p0, p1, b)
            r = gpu.results[(b, b - p0, p1 - b)]
            i, j = b - p0, p1 - b
            np.testing.assert_array_equal(r["lowres_costs"], ref.get(b, GET_LOWRES_COSTS, i, j), err_msg=f"({p0},{p1},{b})")
            if i:
                np.testing.assert_array_equal(r["mvs0"], ref.get(b, GET_MVS, 0, i - 1))
                np.testing.assert_array_equal(r["mv_costs0"], ref.get(b, GET_MV_COSTS, 0, i - 1))
            if j:
                np.testing.assert_array_equal(r["mvs1"], ref.get(b, GET_MVS, 1, j - 1))
            assert sr == sg
    finally:
        ref.close()
        gpu.close()


def test_1080p_properties():
    """full-size run: properties that do not need the (slow) CPU oracle at 1080p."""
    x264cu = _cu()
    W, H = 1920, 1080
    frames = synth_frames(W, H, 3, seed=7)
    gpu = x264cu.CudaLookahead(W, H, bframes=3)
    try:
        for i, f in enumerate(frames):
            gpu.put_frame(i, f)
        # identical frames -> zero motion, cost = lowres_penalty-dominated, every block inter
        gpu.put_frame(3, frames[2])
        s = gpu.frame_cost(2, 3, 3)
        r = gpu.results[(3, 1, 0)]
        assert (r["mvs0"] == 0).all()
        assert (r["mv_costs0"] == 0).all()
        assert ((r["lowres_costs"] & 0x3FFF) == 4).all() and ((r["lowres_costs"] >> 14) == 1).all()
        assert s == 4 * (gpu.mbw - 2) * (gpu.mbh - 2)
        # row sums add up to the frame sums plus the border blocks
        s = gpu.frame_cost(0, 1, 1)
        r = gpu.results[(1, 1, 0)]
        costs = (r["lowres_costs"] & 0x3FFF).astype(np.int64).reshape(gpu.mbh, gpu.mbw)
        assert costs.sum(axis=1).tolist() == r["row_satds"].tolist()
        assert costs[1:-1, 1:-1].sum() == r["stats"][0] == s
        # pure global pan (3,1) full-res px/frame = (6,2) qpel in lowres: most interior MVs find it
        mv = r["mvs0"].reshape(gpu.mbh, gpu.mbw, 2)[2:-2, 2:-2]
        frac = np.mean((mv[..., 0] == 6) & (mv[..., 1] == 2))
        assert frac > 0.6, frac
    finally:
        gpu.close()
