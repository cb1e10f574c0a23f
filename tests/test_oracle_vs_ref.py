"""CPU: the oracle restatement against the live reference build (oracle/_ref/libx264ref.so) on
randomised inputs.  Skipped where the reference build is absent (it is git-ignored; the golden
tests cover that case)."""
import numpy as np
import pytest

from checkers import (GET_INTRA, GET_LOWRES_COSTS, GET_MV_COSTS, GET_MVS, GET_ROW_SATDS, GET_STATS, SIZE_WH, Lookahead,
                      OracleLib, RefLib, have_ref, synth_frames)

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libx264ref.so not built")


@pytest.fixture(scope="module")
def libs():
    return RefLib(), OracleLib()


def _planes(seed, h=270, w=480, extreme=False):
    rng = np.random.default_rng(seed)
    if extreme:
        a = np.where(rng.integers(0, 2, (h, w)) == 1, 255, 0).astype(np.uint8)
        return a, (255 - a).astype(np.uint8)
    return rng.integers(0, 256, (h, w), dtype=np.uint8), rng.integers(0, 256, (h, w), dtype=np.uint8)


@pytest.mark.parametrize("extreme", [False, True])
def test_pixel_metrics_random_offsets(libs, extreme):
    ref, orc = libs
    a, b = _planes(1, extreme=extreme)
    rng = np.random.default_rng(2)
    h, w = a.shape
    for kind in ("sad", "satd", "sa8d"):
        for size, (bw, bh) in enumerate(SIZE_WH):
            if kind == "sa8d" and size not in (0, 3):
                continue
            n = 300
            o1 = rng.integers(0, h - bh, n) * w + rng.integers(0, w - bw, n)
            o2 = rng.integers(0, h - bh, n) * w + rng.integers(0, w - bw, n)
            np.testing.assert_array_equal(orc.pixel_cmp(kind, size, a.reshape(-1), w, b.reshape(-1), w, o1, o2),
                                          ref.pixel_cmp(kind, size, a.reshape(-1), w, b.reshape(-1), w, o1, o2))


@pytest.mark.parametrize("extreme", [False, True])
def test_hadamard_dct_quant(libs, extreme):
    ref, orc = libs
    a, b = _planes(3, extreme=extreme)
    rng = np.random.default_rng(4)
    h, w = a.shape
    for size in range(4):
        bw, bh = SIZE_WH[size]
        o = rng.integers(0, h - bh, 200) * w + rng.integers(0, w - bw, 200)
        np.testing.assert_array_equal(orc.hadamard_ac(size, a.reshape(-1), w, o), ref.hadamard_ac(size, a.reshape(-1), w, o))
    for bs in (4, 8, 16):
        o1 = rng.integers(0, h - bs, 200) * w + rng.integers(0, w - bs, 200)
        o2 = rng.integers(0, h - bs, 200) * w + rng.integers(0, w - bs, 200)
        co = ref.sub_dct(bs, a.reshape(-1), w, b.reshape(-1), w, o1, o2)
        np.testing.assert_array_equal(orc.sub_dct(bs, a.reshape(-1), w, b.reshape(-1), w, o1, o2), co)
        for qp in (0, 17, 30, 51):
            for b8 in (0, 1):
                mf, bias = ref.cqm_tables(b8, qp & 1, qp)
                q1, n1 = orc.quant(b8, co, mf, bias)
                q2, n2 = ref.quant(b8, co, mf, bias)
                np.testing.assert_array_equal(q1, q2)
                np.testing.assert_array_equal(n1, n2)


@pytest.mark.parametrize("W,H,subme,mbtree,aq,passes,bframes", [
    (176, 144, 7, 1, 1, 2, 3),
    (176, 144, 1, 0, 1, 2, 2),
    (200, 120, 7, 0, 0, 1, 3),      # ragged size
    (48, 32, 7, 1, 1, 3, 1),
])
def test_lookahead_random(libs, W, H, subme, mbtree, aq, passes, bframes):
    ref, orc = libs
    frames = synth_frames(W, H, 5, seed=W + subme, pan=(-3, 2), noise=6)
    rng = np.random.default_rng(8)
    A = Lookahead(ref, "ref", W, H, bframes=bframes, subme=subme, mbtree=mbtree, aq=aq, passes=passes)
    B = Lookahead(orc, "oracle", W, H, bframes=bframes, subme=subme, mbtree=mbtree, aq=aq, passes=passes, mv_range=A.mv_range)
    try:
        for i, f in enumerate(frames):
            q = rng.integers(64, 1024, A.nmb).astype(np.uint16) if aq else None
            A.put_frame(i, f, q)
            B.put_frame(i, f, q)
        triples = [(p0, p1, b) for p0 in range(3) for p1 in range(p0, min(5, p0 + bframes + 2)) for b in range(p0, p1 + 1)
                   if (b == p1 or b > p0)]
        for (p0, p1, b) in triples:
            assert A.frame_cost(p0, p1, b) == B.frame_cost(p0, p1, b), (p0, p1, b)
            i, j = b - p0, p1 - b
            for what in (GET_LOWRES_COSTS, GET_ROW_SATDS, GET_STATS):
                np.testing.assert_array_equal(A.get(b, what, i, j), B.get(b, what, i, j), err_msg=str((p0, p1, b, what)))
            np.testing.assert_array_equal(A.get(b, GET_INTRA), B.get(b, GET_INTRA))
            if i:
                np.testing.assert_array_equal(A.get(b, GET_MVS, 0, i - 1), B.get(b, GET_MVS, 0, i - 1))
                np.testing.assert_array_equal(A.get(b, GET_MV_COSTS, 0, i - 1), B.get(b, GET_MV_COSTS, 0, i - 1))
            if j:
                np.testing.assert_array_equal(A.get(b, GET_MVS, 1, j - 1), B.get(b, GET_MVS, 1, j - 1))
                np.testing.assert_array_equal(A.get(b, GET_MV_COSTS, 1, j - 1), B.get(b, GET_MV_COSTS, 1, j - 1))
    finally:
        A.close()
        B.close()


@pytest.mark.parametrize("subme,mbtree", [(7, 1), (7, 0), (1, 1)])
def test_serial_mode_equals_unmodified_reference(libs, subme, mbtree):
    """search_passes = 0: the oracle in the reference's raster-serial schedule against the reference's OWN
    slicetype_frame_cost (slicetype.c:836), nothing modified, one lookahead thread."""
    ref, orc = libs
    W, H = 208, 160
    frames = synth_frames(W, H, 5, seed=12, pan=(-4, 3), noise=5)
    rng = np.random.default_rng(2)
    A = Lookahead(ref, "ref_serial", W, H, bframes=3, subme=subme, mbtree=mbtree, aq=1)
    B = Lookahead(orc, "oracle", W, H, bframes=3, subme=subme, mbtree=mbtree, aq=1, passes=0, mv_range=A.mv_range)
    try:
        for i, f in enumerate(frames):
            q = rng.integers(64, 1024, A.nmb).astype(np.uint16)
            A.put_frame(i, f, q)
            B.put_frame(i, f, q)
        for (p0, p1, b) in [(0, 0, 0), (0, 1, 1), (0, 3, 3), (0, 3, 1), (0, 3, 2), (1, 4, 4), (1, 4, 2), (1, 4, 3), (2, 4, 4), (2, 4, 3)]:
            assert A.frame_cost(p0, p1, b) == B.frame_cost(p0, p1, b), (p0, p1, b)
            i, j = b - p0, p1 - b
            np.testing.assert_array_equal(A.get(b, GET_LOWRES_COSTS, i, j), B.get(b, GET_LOWRES_COSTS, i, j))
            if i:
                np.testing.assert_array_equal(A.get(b, GET_MVS, 0, i - 1), B.get(b, GET_MVS, 0, i - 1))
                np.testing.assert_array_equal(A.get(b, GET_MV_COSTS, 0, i - 1), B.get(b, GET_MV_COSTS, 0, i - 1))
            if j:
                np.testing.assert_array_equal(A.get(b, GET_MVS, 1, j - 1), B.get(b, GET_MVS, 1, j - 1))
    finally:
        A.close()
        B.close()


def test_jacobi_tracks_raster_serial_quality(libs):
    """sanity (not parity): the dependency-free schedule with K=2 stays close to the reference's
    own raster-serial search on panning content (secondary oracle)."""
    ref, _ = libs
    W, H = 352, 288
    frames = synth_frames(W, H, 4, seed=2)
    A = Lookahead(ref, "ref", W, H, passes=2)
    S = Lookahead(ref, "ref_serial", W, H)
    try:
        for i, f in enumerate(frames):
            A.put_frame(i, f)
            S.put_frame(i, f)
        for (p0, p1, b) in [(0, 1, 1), (0, 3, 3), (0, 3, 1)]:
            a, s = A.frame_cost(p0, p1, b), S.frame_cost(p0, p1, b)
            assert abs(a - s) <= 0.05 * s, (a, s)
    finally:
        A.close()
        S.close()
