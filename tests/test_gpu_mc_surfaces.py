"""GPU: the batched entry points for the checkasm mc / intra / quant surfaces (csrc/primitives_mc.cu: x264cu_mc_luma,
x264cu_pixel_avg, x264cu_mc_weight, x264cu_intra_x3_8x8c, x264cu_quant_4x4x4, and x264cu_quant under six CQM configurations
x QP 0..51) against the CPU oracle and the reference's golden vectors, bit for bit.  Mirrors tools/checkasm.c:754,
1234-1429, 2098-2227."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import mc_cases as M
from checkers import OracleLib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "x264-i386pic_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
pytestmark = pytest.mark.gpu
vp = C.c_void_p


@pytest.fixture(scope="module")
def env():
    import make_golden_mc as G
    import x264cu
    ctx = x264cu.Context(320, 192, bframes=3)
    yield x264cu, ctx, OracleLib(), G, G.golden_inputs(), np.load(os.path.join(ROOT, "tests", "golden", "mc_surfaces.npz"))
    ctx.close()


def gpu_mc_luma(x264cu, ctx, buf, po, stride, org, mv, w, h, wt):
    out = np.zeros((len(org), h, w), dtype=np.uint8)
    poa = (C.c_int64 * 4)(*po)
    weight = None if wt is None else x264cu.Weight(1, *wt)
    x264cu.check(ctx.L.x264cu_mc_luma(ctx.ctx, buf.ctypes.data, buf.nbytes, poa, stride, len(org), org.ctypes.data, np.ascontiguousarray(mv).ctypes.data,
                                      w, h, None if weight is None else C.addressof(weight), out.ctypes.data, 0), "mc_luma")
    return out


def test_mc_luma_phases_sizes_weights(env):
    x264cu, ctx, O, G, I, gold = env
    buf, po, stride = I["buf"], I["po"], I["stride"]
    rng = np.random.default_rng(3)
    for (w, h) in M.MC_SIZES:
        org, mv = M.mc_cases(stride, (len(buf) // 4) // stride, w, h, rng)
        assert len({(int(x) & 3, int(y) & 3) for x, y in mv}) == 16
        for wt in M.WEIGHTS:                                  # incl. denom == 0 (mc.c:118) and x264_weight_none
            np.testing.assert_array_equal(gpu_mc_luma(x264cu, ctx, buf, po, stride, org, mv, w, h, wt),
                                          M.oracle_mc_luma(O, buf, po, stride, org, mv, w, h, wt), err_msg=str((w, h, wt)))
    for (w, h), (org, mv) in I["mc"].items():
        for k, wt in enumerate(G.GOLD_WEIGHTS):
            np.testing.assert_array_equal(gpu_mc_luma(x264cu, ctx, buf, po, stride, org, mv, w, h, wt), gold["mc_%dx%d_%d" % (w, h, k)])


def test_avg_sizes_and_weights(env):
    x264cu, ctx, O, G, I, gold = env
    a, b, off1, off2 = I["a"], I["b"], I["off1"], I["off2"]
    for size in range(12):
        w, h = M.AVG_WH[size]
        for weight in list(range(-63, 128, 7)) + G.AVG_WEIGHTS:
            out = np.zeros((len(off1), h, w), dtype=np.uint8)
            x264cu.check(ctx.L.x264cu_pixel_avg(ctx.ctx, size, a.ctypes.data, a.nbytes, 64, b.ctypes.data, b.nbytes, 48, len(off1),
                                                off1.ctypes.data, off2.ctypes.data, weight, out.ctypes.data, 0), "avg")
            np.testing.assert_array_equal(out, M.oracle_avg(O, a, 64, b, 48, off1, off2, size, weight), err_msg=str((size, weight)))
            if weight in G.AVG_WEIGHTS:
                np.testing.assert_array_equal(out, gold["avg_%d_%d" % (size, weight)])


def test_mc_weight_incl_denom_zero(env):
    x264cu, ctx, O, G, I, gold = env
    src = I["wsrc"]
    for k, wt in enumerate(G.W_CASES + [(64, 6, 0), (91, 6, 12), (1, 0, -40)]):
        for width in (4, 20, 17):
            out = np.zeros((16, width), dtype=np.uint8)
            weight = x264cu.Weight(1, *wt)
            x264cu.check(ctx.L.x264cu_mc_weight(ctx.ctx, src.ctypes.data, 32, out.ctypes.data, width, width, 16, C.addressof(weight), 0), "mc_weight")
            np.testing.assert_array_equal(out, M.oracle_weight(O, src, 32, width, 16, wt), err_msg=str((wt, width)))
            if width == 20 and k < len(G.W_CASES):
                np.testing.assert_array_equal(out, gold["weight_%d" % k])


@pytest.mark.parametrize("b_satd", [0, 1])
def test_intra_x3_8x8c(env, b_satd):
    x264cu, ctx, O, G, I, gold = env
    p1, p2, s1, s2, o1, o2 = I["p1"], I["p2"], I["s1"], I["s2"], I["io1"], I["io2"]
    res = np.zeros((len(o1), 3), dtype=np.int32)
    x264cu.check(ctx.L.x264cu_intra_x3_8x8c(ctx.ctx, b_satd, p1.ctypes.data, p1.nbytes, s1, p2.ctypes.data, p2.nbytes, s2, len(o1),
                                            o1.ctypes.data, o2.ctypes.data, res.ctypes.data, 0), "intra_x3")
    np.testing.assert_array_equal(res, gold["intra_x3_%d" % b_satd])
    # every 8x8 of a plane against the oracle (neighbours from the same plane, as the lookahead reads them)
    rng = np.random.default_rng(9)
    s = 136
    pl = rng.integers(0, 256, s * 72, dtype=np.uint8)
    off = np.array([s * (8 * y + 1) + 8 * x + 1 for y in range(8) for x in range(16)], dtype=np.int64)
    res = np.zeros((len(off), 3), dtype=np.int32)
    x264cu.check(ctx.L.x264cu_intra_x3_8x8c(ctx.ctx, b_satd, pl.ctypes.data, pl.nbytes, s, pl.ctypes.data, pl.nbytes, s, len(off),
                                            off.ctypes.data, off.ctypes.data, res.ctypes.data, 0), "intra_x3")
    np.testing.assert_array_equal(res, M.oracle_intra_x3(O, b_satd, pl, s, pl, s, off, off))


def test_quant_under_six_cqm_configurations(env):
    """quant_4x4, quant_8x8 and quant_4x4x4 with the reference's x264_cqm_init tables (flat, JVT, custom 4x4, custom 8x8,
    random, all-ones) at every QP 0..51 (checkasm.c:2098-2227)"""
    x264cu, ctx, O, G, I, gold = env
    rng = np.random.default_rng(7)
    k = 0
    for i_cqm in range(6):
        for qp in range(52):
            mf4, b4 = gold["cqm_mf4"][i_cqm, qp], gold["cqm_bias4"][i_cqm, qp]
            mf8, b8 = gold["cqm_mf8"][i_cqm, qp], gold["cqm_bias8"][i_cqm, qp]
            c = M.quant_coefs(rng, 8, [0xF, 0x5, 0x0, 0x8][qp & 3])
            exp_q, exp_mask = M.quant_4x4x4_expect(O.quant, c, mf4, b4)
            got = c.copy()
            mask = np.zeros(8, dtype=np.int32)
            x264cu.check(ctx.L.x264cu_quant_4x4x4(ctx.ctx, got.ctypes.data, 8, np.ascontiguousarray(mf4).ctypes.data, np.ascontiguousarray(b4).ctypes.data,
                                                  mask.ctypes.data, 0), "quant_4x4x4")
            np.testing.assert_array_equal(got, exp_q, err_msg=str((i_cqm, qp))); np.testing.assert_array_equal(mask, exp_mask)
            q4, nz4 = ctx.quant(0, c.reshape(-1, 16), mf4, b4)
            eq4, en4 = O.quant(0, c.reshape(-1, 16), mf4, b4)
            np.testing.assert_array_equal(q4, eq4); np.testing.assert_array_equal(nz4, en4)
            c8 = rng.integers(-4000, 4001, (5, 64)).astype(np.int16)
            q8, nz8 = ctx.quant(1, c8, mf8, b8)
            eq8, en8 = O.quant(1, c8, mf8, b8)
            np.testing.assert_array_equal(q8, eq8); np.testing.assert_array_equal(nz8, en8)
            if qp % 13 == 0:
                got = gold["q444_in"].copy()
                mask = np.zeros(4, dtype=np.int32)
                x264cu.check(ctx.L.x264cu_quant_4x4x4(ctx.ctx, got.ctypes.data, 4, np.ascontiguousarray(mf4).ctypes.data,
                                                      np.ascontiguousarray(b4).ctypes.data, mask.ctypes.data, 0), "quant_4x4x4")
                np.testing.assert_array_equal(got, gold["q444_out"][k]); np.testing.assert_array_equal(mask, gold["q444_mask"][k])
                k += 1
