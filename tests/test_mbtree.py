"""MB-tree (first "next" row): mbtree_propagate_cost / mbtree_propagate_list (common/mc.c:511-598),
macroblock_tree_propagate / _finish and the macroblock_tree frame walk (encoder/slicetype.c:1029-1188).

CPU: the oracle restatement against the reference's own mc table entries on checkasm-style random data
(tools/checkasm.c:1760-1844) and against the reference's own macroblock_tree run on the same frames, costs and MVs.
GPU: the device kernels (driven by the same test-side frame walk) against the oracle, bit-exact including the float qp offsets."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from checkers import (GET_INVQ, Lookahead, OracleLib, RefLib, have_ref, synth_frames)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "x264-i386pic_b200"))
vp = C.c_void_p
f32 = np.float32


def clipdur(x):
    return f32(min(max(f32(x), f32(0.01)), f32(1.0)))


def walk(be, types, durations, num_frames, b_intra, rc_lookahead=40, pyramid=1, vbv=0, qcompress=f32(0.6)):
    """macroblock_tree (slicetype.c:1091-1188) over an abstract backend."""
    is_b = lambda t: t in (4, 5)   # noqa: E731
    idx, cur_nonb, bframes = int(not b_intra), 1, 0
    total = f32(0.0)
    for j in range(num_frames + 1):
        total = f32(total + f32(durations[j]))
    avg = f32(total / f32(num_frames + 1))

    def finish(i, ref0_distance):
        fps = int(round(float(f32(f32(f32(clipdur(avg) / clipdur(durations[i])) * f32(256)) / f32(0.5)))))
        be.finish(i, fps, f32(0.0), f32(f32(5.0) * f32(f32(1.0) - qcompress)))

    def propagate(p0, p1, b, referenced):
        fps = f32(f32(clipdur(durations[b]) / f32(clipdur(avg) * f32(256.0))) * f32(0.5))
        be.propagate(p0, p1, b, fps, referenced)

    i = num_frames
    if b_intra:
        be.cost(0, 0, 0)
    while i > 0 and is_b(types[i]):
        i -= 1
    last_nonb = i
    assert rc_lookahead
    if last_nonb < idx:
        return
    be.reset(last_nonb)
    while True:
        i -= 1
        if not (i + 1 > idx):
            break
        cur_nonb = i
        while is_b(types[cur_nonb]) and cur_nonb > 0:
            cur_nonb -= 1
        if cur_nonb < idx:
            break
        be.cost(cur_nonb, last_nonb, last_nonb)
        be.reset(cur_nonb)
        bframes = last_nonb - cur_nonb - 1
        if pyramid and bframes > 1:
            middle = (bframes + 1) // 2 + cur_nonb
            be.cost(cur_nonb, last_nonb, middle)
            be.reset(middle)
            while i > cur_nonb:
                p0 = middle if i > middle else cur_nonb
                p1 = middle if i < middle else last_nonb
                if i != middle:
                    be.cost(p0, p1, i)
                    propagate(p0, p1, i, 0)
                i -= 1
            propagate(cur_nonb, last_nonb, middle, 1)
        else:
            while i > cur_nonb:
                be.cost(cur_nonb, last_nonb, i)
                propagate(cur_nonb, last_nonb, i, 0)
                i -= 1
        propagate(cur_nonb, last_nonb, last_nonb, 1)
        last_nonb = cur_nonb
    finish(last_nonb, last_nonb)
    if pyramid and bframes > 1 and not vbv:
        finish(last_nonb + (bframes + 1) // 2, 0)


class OracleBackend:
    def __init__(self, la, qp_aq):
        self.la, self.L = la, la.lib.lib
        self.qp_aq = qp_aq
        self.qp = [a.copy() for a in qp_aq]
        L = self.L
        L.xo_mbtree_reset.argtypes = [vp, C.c_int]
        L.xo_mbtree_get.argtypes = [vp, C.c_int, vp]
        L.xo_mbtree_propagate.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int]
        L.xo_mbtree_finish.argtypes = [vp, C.c_int, C.c_int, C.c_float, C.c_float, vp, vp]

    def cost(self, p0, p1, b):
        self.la.frame_cost(p0, p1, b)

    def reset(self, i):
        self.L.xo_mbtree_reset(self.la.ctx, i)

    def propagate(self, p0, p1, b, fps, referenced):
        assert self.L.xo_mbtree_propagate(self.la.ctx, p0, p1, b, float(fps), referenced) == 0

    def finish(self, i, fps, wd, strength):
        self.L.xo_mbtree_finish(self.la.ctx, i, fps, float(wd), float(strength), self.qp_aq[i].ctypes.data, self.qp[i].ctypes.data)

    def prop(self, i):
        out = np.zeros(self.la.nmb, dtype=np.uint16)
        self.L.xo_mbtree_get(self.la.ctx, i, out.ctypes.data)
        return out


TYPES = [2, 5, 5, 3, 5, 3, 3, 5, 5, 5, 3]         # I B B P B P P B B B P


def setup_ref(W, H, frames):
    ref = RefLib()
    L = ref.lib
    L.ref_la_put_frame_aq.argtypes = [vp, C.c_int, vp, C.c_int]
    L.ref_la_get_mbtree.argtypes = [vp, C.c_int, C.c_int, vp]
    L.ref_la_macroblock_tree.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp]
    R = Lookahead(ref, "ref", W, H, bframes=3, mbtree=1, aq=1, rc_lookahead=20)
    invq, qp_aq = [], []
    for i, f in enumerate(frames):
        f = np.ascontiguousarray(f)
        assert L.ref_la_put_frame_aq(R.ctx, i, f.ctypes.data, W) == 0
        invq.append(R.get(i, GET_INVQ))
        a = np.zeros(R.nmb, dtype=np.float32)
        L.ref_la_get_mbtree(R.ctx, i, 0, a.ctypes.data)
        qp_aq.append(a)
    return ref, R, invq, qp_aq


def ref_mbtree(ref, R, n, durations):
    L = ref.lib
    # the triples the walk asks for, evaluated in the dependency-free schedule so that the walk hits the caches
    class B:
        def cost(self, p0, p1, b): R.frame_cost(p0, p1, b)
        def reset(self, i): pass
        def propagate(self, *a): pass
        def finish(self, *a): pass
    walk(B(), TYPES[:n], durations, n - 1, 1)
    t = np.array(TYPES[:n], dtype=np.uint8)
    d = np.array(durations[:n], dtype=np.float32)
    assert L.ref_la_macroblock_tree(R.ctx, 0, n - 1, 1, t.ctypes.data, d.ctypes.data) == 0
    props, qps = [], []
    for i in range(n):
        p = np.zeros(R.nmb, dtype=np.uint16); L.ref_la_get_mbtree(R.ctx, i, 2, p.ctypes.data); props.append(p)
        q = np.zeros(R.nmb, dtype=np.float32); L.ref_la_get_mbtree(R.ctx, i, 1, q.ctypes.data); qps.append(q)
    return props, qps


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libx264ref.so not built")
def test_primitives_vs_reference():
    ref, orc = RefLib(), OracleLib()
    rng = np.random.default_rng(3)
    lut = np.zeros(128, dtype=np.float32)
    ref.lib.ref_log2_lut(lut.ctypes.data)
    np.testing.assert_array_equal(lut, np.array([np.float32(round(np.log2(1 + i / 128.0) * 1e5) / 1e5) for i in range(128)], dtype=np.float32))
    n = 4000
    intra = rng.integers(9, 0x3FFF, n).astype(np.uint16)
    inter = (rng.integers(0, 0x3FFF, n) | (rng.integers(0, 4, n) << 14)).astype(np.uint16)
    prop_in = rng.integers(0, 32768, n).astype(np.uint16)
    invq = rng.integers(1, 2000, n).astype(np.uint16)
    ref.lib.ref_mbtree_propagate_cost.argtypes = [vp] * 5 + [C.c_float, C.c_int]
    orc.lib.xo_mbtree_propagate_cost.argtypes = [vp] * 5 + [C.c_float, C.c_int]
    for fps in (0.001953125, 0.0007, 0.02):
        a, b = np.zeros(n, dtype=np.int16), np.zeros(n, dtype=np.int16)
        ref.lib.ref_mbtree_propagate_cost(a.ctypes.data, prop_in.ctypes.data, intra.ctypes.data, inter.ctypes.data, invq.ctypes.data, fps, n)
        orc.lib.xo_mbtree_propagate_cost(b.ctypes.data, prop_in.ctypes.data, intra.ctypes.data, inter.ctypes.data, invq.ctypes.data, fps, n)
        np.testing.assert_array_equal(a, b)
    # propagate_list over a whole frame with random MVs reaching outside (checkasm.c:1797-1844)
    W, H = 320, 192
    R = Lookahead(ref, "ref", W, H, bframes=3)
    mbw, mbh = R.mbw, R.mbh
    ref.lib.ref_mbtree_propagate_list.argtypes = [vp] * 5 + [C.c_int] * 4
    orc.lib.xo_mbtree_propagate_list.argtypes = [C.c_int, C.c_int] + [vp] * 4 + [C.c_int] * 4
    for list_, bw in ((0, 37), (1, 27), (0, 32)):
        ca = rng.integers(0, 30000, mbw * mbh).astype(np.uint16)
        cb = ca.copy()
        for y in range(mbh):
            mvs = rng.integers(-700, 700, (mbw, 2)).astype(np.int16)
            mvs[::5] = 0
            amt = rng.integers(0, 32768, mbw).astype(np.int16)
            lc = (rng.integers(0, 0x3FFF, mbw) | (rng.integers(0, 4, mbw) << 14)).astype(np.uint16)
            ref.lib.ref_mbtree_propagate_list(R.ctx, ca.ctypes.data, mvs.ctypes.data, amt.ctypes.data, lc.ctypes.data, bw, y, mbw, list_)
            orc.lib.xo_mbtree_propagate_list(mbw, mbh, cb.ctypes.data, mvs.ctypes.data, amt.ctypes.data, lc.ctypes.data, bw, y, mbw, list_)
        np.testing.assert_array_equal(ca, cb)
    R.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libx264ref.so not built")
def test_walk_oracle_vs_reference():
    W, H, n = 176, 144, len(TYPES)
    frames = synth_frames(W, H, n, seed=4, pan=(3, -2))
    durations = [f32(1 / 30.0)] * n
    ref, R, invq, qp_aq = setup_ref(W, H, frames)
    O = Lookahead(OracleLib(), "oracle", W, H, bframes=3, mbtree=1, aq=1, mv_range=R.mv_range)
    try:
        for i, f in enumerate(frames):
            O.put_frame(i, f, invq[i])
        props_r, qps_r = ref_mbtree(ref, R, n, durations)
        be = OracleBackend(O, qp_aq)
        walk(be, TYPES, durations, n - 1, 1, rc_lookahead=20)
        for i in range(n):
            np.testing.assert_array_equal(be.qp[i], qps_r[i], err_msg=f"f_qp_offset of frame {i}")
        for i in (0, 3, 5, 6):      # frames whose accumulators the walk owns at the end (the non-B frames)
            np.testing.assert_array_equal(be.prop(i), props_r[i], err_msg=f"i_propagate_cost of frame {i}")
        assert any((be.qp[i] != qp_aq[i]).any() for i in range(n))
    finally:
        R.close()
        O.close()


class CudaBackend:
    """the same abstract backend over the C ABI: x264cu_frame_cost / x264cu_mbtree_reset / _propagate / _finish.  The frame
    walk itself is host control logic and stays outside the product (the reference's own macroblock_tree runs it)."""

    def __init__(self, G, qp_aq):
        self.G, self.L = G, G.ctx.L
        self.qp_aq = qp_aq
        self.qp = [a.copy() for a in qp_aq]
        L = self.L
        L.x264cu_mbtree_reset.argtypes = [vp, vp]
        L.x264cu_mbtree_propagate.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_float, C.c_int]
        L.x264cu_mbtree_finish.argtypes = [vp, vp, C.c_int, C.c_float, C.c_float, vp, vp]
        L.x264cu_mbtree_get_propagate.argtypes = [vp, vp, vp]

    def cost(self, p0, p1, b):
        self.G.frame_cost(p0, p1, b)

    def reset(self, i):
        assert self.L.x264cu_mbtree_reset(self.G.ctx.ctx, self.G.frames[i]) == 0

    def propagate(self, p0, p1, b, fps, referenced):
        F = self.G.frames
        assert self.L.x264cu_mbtree_propagate(self.G.ctx.ctx, F[b], F[p0], F[p1], b - p0, p1 - b, float(fps), referenced) == 0

    def finish(self, i, fps, wd, strength):
        assert self.L.x264cu_mbtree_finish(self.G.ctx.ctx, self.G.frames[i], fps, float(wd), float(strength),
                                           self.qp_aq[i].ctypes.data, self.qp[i].ctypes.data) == 0

    def prop(self, i):
        out = np.zeros(self.G.nmb, dtype=np.uint16)
        assert self.L.x264cu_mbtree_get_propagate(self.G.ctx.ctx, self.G.frames[i], out.ctypes.data) == 0
        return out


@pytest.mark.gpu
@pytest.mark.parametrize("W,H", [(176, 144), (640, 360)])
def test_device_walk_vs_oracle(W, H):
    import x264cu
    n = len(TYPES)
    frames = synth_frames(W, H, n, seed=4, pan=(3, -2))
    durations = [f32(1 / 30.0)] * n
    rng = np.random.default_rng(1)
    mvr = x264cu.default_mv_range(W, H)
    O = Lookahead(OracleLib(), "oracle", W, H, bframes=3, mbtree=1, aq=1, mv_range=mvr)
    G = x264cu.CudaLookahead(W, H, bframes=3, mbtree=1, aq=1, mv_range=mvr)
    try:
        invq = [rng.integers(100, 700, O.nmb).astype(np.uint16) for _ in range(n)]
        qp_aq = [(rng.standard_normal(O.nmb) * 1.5).astype(np.float32) for _ in range(n)]
        for i, f in enumerate(frames):
            O.put_frame(i, f, invq[i])
            G.put_frame(i, f, invq[i])
        be = OracleBackend(O, qp_aq)
        walk(be, TYPES, durations, n - 1, 1, rc_lookahead=20)
        gb = CudaBackend(G, qp_aq)
        walk(gb, TYPES, durations, n - 1, 1, rc_lookahead=20)
        for i in range(n):
            np.testing.assert_array_equal(gb.qp[i], be.qp[i], err_msg=f"f_qp_offset of frame {i}")
        for i in (0, 3, 5, 6):
            np.testing.assert_array_equal(gb.prop(i), be.prop(i), err_msg=f"i_propagate_cost of frame {i}")
    finally:
        O.close()
        G.close()
