"""CPU: the oracle restatement (oracle/x264_oracle.c) against the golden vectors that
tests/golden/make_golden.py produced from the unmodified reference build."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from checkasm_data import PBUF2_OFF, SIZE_WH, pixel_cases
from checkers import (GET_INTRA, GET_LOWRES_COSTS, GET_MV_COSTS, GET_MVS, GET_ROW_SATDS, GET_STATS, Lookahead, OracleLib)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def prim():
    return np.load(os.path.join(GOLD, "primitives.npz"))


@pytest.fixture(scope="module")
def look():
    return np.load(os.path.join(GOLD, "lookahead.npz"))


@pytest.fixture(scope="module")
def orc():
    return OracleLib()


@pytest.mark.parametrize("kind", ["sad", "satd", "sa8d"])
def test_pixel_metrics(prim, orc, kind):
    for size in range(8):
        if kind == "sa8d" and size not in (0, 3):
            continue
        res = []
        for (which, s1, s2, o1, o2) in pixel_cases(size):
            a, b = (prim["pbuf1"], prim["pbuf1"]) if which == "rand" else (prim["pbuf3"], prim["pbuf4"])
            res.append(orc.pixel_cmp(kind, size, a, s1, b, s2, o1, o2))
        np.testing.assert_array_equal(np.concatenate(res), prim[f"{kind}_{size}"], err_msg=f"{kind} {size}")


def test_hadamard_ac(prim, orc):
    offs = np.arange(32, dtype=np.int64)
    for size in range(4):
        got = np.concatenate([orc.hadamard_ac(size, prim["pbuf1"], 16, offs * 30),
                              orc.hadamard_ac(size, prim["pbuf3"], 16, offs * 100),
                              orc.hadamard_ac(size, prim["pbuf4"], 32, offs * 64)])
        np.testing.assert_array_equal(got, prim[f"hadamard_ac_{size}"])


def test_sad_x4(prim, orc):
    p = prim["pbuf1"]
    for size in range(7):
        offs = np.array([PBUF2_OFF, PBUF2_OFF + 6, PBUF2_OFF + 1, PBUF2_OFF + 10], dtype=np.int64)
        got = orc.pixel_cmp("sad", size, p, 16, p, 64, np.zeros(4, dtype=np.int64), offs)
        np.testing.assert_array_equal(got, prim[f"sad_x4_{size}"])


def test_sub_dct(prim, orc):
    a, b, offs = prim["dct_a"].reshape(-1), prim["dct_b"].reshape(-1), prim["dct_offs"]
    for bs in (4, 8, 16):
        np.testing.assert_array_equal(orc.sub_dct(bs, a, 64, b, 64, offs, offs), prim[f"dct_{bs}"])


def test_quant(prim, orc):
    for b8 in (0, 1):
        for cat in (0, 1):
            tag = f"quant{'8' if b8 else '4'}_{cat}"
            for k in range(len(prim["quant_qps"])):
                q, nz = orc.quant(b8, prim[tag + "_in"][k], prim[tag + "_mf"][k], prim[tag + "_bias"][k])
                np.testing.assert_array_equal(q.reshape(prim[tag + "_out"][k].shape), prim[tag + "_out"][k])
                np.testing.assert_array_equal(nz, prim[tag + "_nz"][k])


def test_lowres_core(prim, orc):
    src = prim["lowres_src"].reshape(-1)
    for w in (96, 104, 112, 120):
        got = np.stack(orc.lowres_core(src, 256, w, 8, 128))
        np.testing.assert_array_equal(got, prim[f"lowres_{w}"])


def test_mc_luma_and_avg(prim, orc):
    planes = [np.ascontiguousarray(p) for p in prim["mc_planes"]]
    mc = []
    for (w, h) in ((8, 8), (8, 9), (12, 8)):
        for mvy in range(-5, 7):
            for mvx in range(-5, 7):
                for wt in (None, (58, 6, 3), (127, 7, -10), (3, 0, 1)):
                    mc.append(orc.mc_luma(planes, 64, 24 * 64 + 24, mvx, mvy, w, h, wt).reshape(-1))
    np.testing.assert_array_equal(np.concatenate(mc), prim["mc_out"])
    src = prim["lowres_src"]
    avg = [orc.avg(src[:8, 10:18].copy(), src[8:16, 33:41].copy(), 8, 8, wgt).reshape(-1) for wgt in (32, 16, 48, 64 - 43, 0, 64)]
    np.testing.assert_array_equal(np.concatenate(avg), prim["avg_out"])


def test_intra_cost(prim, orc):
    res = []
    for plane in prim["intra_planes"]:
        plane = np.ascontiguousarray(plane)
        for (x, y) in ((8, 8), (16, 8), (24, 40), (9, 13), (40, 17)):
            for b_satd in (0, 1):
                res.append(orc.lib.xo_intra_cost_8x8(C.c_void_p(plane.ctypes.data + y * 64 + x), 64, b_satd))
    np.testing.assert_array_equal(np.array(res, dtype=np.int32), prim["intra_out"])


def test_cost_mv_table(prim, orc):
    n = 8 * int(prim["cost_mv_range"][0])
    np.testing.assert_array_equal(orc.cost_mv_table(n), prim["cost_mv"])


@pytest.mark.parametrize("name", ["satd_edges", "sad_noedges", "satd_3pass"])
def test_lookahead(look, orc, name):
    W, H, subme, mbtree, aq, passes, mvr = [int(v) for v in look[name + "_cfg"]]
    la = Lookahead(orc, "oracle", W, H, bframes=3, subme=subme, mbtree=mbtree, aq=aq, passes=passes, mv_range=mvr)
    try:
        for i, f in enumerate(look[name + "_frames"]):
            la.put_frame(i, f, look[name + "_invq"][i] if aq else None)
        for p in range(4):
            np.testing.assert_array_equal(la.lowres(1, p), look[name + "_lowres"][p])
        for t, (p0, p1, b) in enumerate(look["triples"]):
            p0, p1, b = int(p0), int(p1), int(b)
            i, j = b - p0, p1 - b
            assert la.frame_cost(p0, p1, b) == int(look[f"{name}_{t}_score"][0])
            np.testing.assert_array_equal(la.get(b, GET_LOWRES_COSTS, i, j), look[f"{name}_{t}_costs"])
            np.testing.assert_array_equal(la.get(b, GET_ROW_SATDS, i, j), look[f"{name}_{t}_rows"])
            np.testing.assert_array_equal(la.get(b, GET_STATS, i, j), look[f"{name}_{t}_stats"])
            np.testing.assert_array_equal(la.get(b, GET_INTRA), look[f"{name}_{t}_intra"])
            if i:
                np.testing.assert_array_equal(la.get(b, GET_MVS, 0, i - 1), look[f"{name}_{t}_mv0"])
                np.testing.assert_array_equal(la.get(b, GET_MV_COSTS, 0, i - 1), look[f"{name}_{t}_mc0"])
            if j:
                np.testing.assert_array_equal(la.get(b, GET_MVS, 1, j - 1), look[f"{name}_{t}_mv1"])
                np.testing.assert_array_equal(la.get(b, GET_MV_COSTS, 1, j - 1), look[f"{name}_{t}_mc1"])
    finally:
        la.close()


# ---- the checkasm mc / intra / quant surfaces (tests/golden/mc_surfaces.npz, make_golden_mc.py) ----
def _mc_golden():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_mc as G
    return G, np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mc_surfaces.npz"))


def test_oracle_mc_surfaces_match_golden():
    import mc_cases as M
    G, gold = _mc_golden()
    O = OracleLib()
    I = G.golden_inputs()
    for (w, h), (org, mv) in I["mc"].items():
        for k, wt in enumerate(G.GOLD_WEIGHTS):
            np.testing.assert_array_equal(M.oracle_mc_luma(O, I["buf"], I["po"], I["stride"], org, mv, w, h, wt), gold["mc_%dx%d_%d" % (w, h, k)])
    for size in range(12):
        for weight in G.AVG_WEIGHTS:
            np.testing.assert_array_equal(M.oracle_avg(O, I["a"], 64, I["b"], 48, I["off1"], I["off2"], size, weight), gold["avg_%d_%d" % (size, weight)])
    for k, wt in enumerate(G.W_CASES):
        np.testing.assert_array_equal(M.oracle_weight(O, I["wsrc"], 32, 20, 16, wt), gold["weight_%d" % k])
    for b_satd in (0, 1):
        np.testing.assert_array_equal(M.oracle_intra_x3(O, b_satd, I["p1"], I["s1"], I["p2"], I["s2"], I["io1"], I["io2"]), gold["intra_x3_%d" % b_satd])
    k = 0
    for i_cqm in range(6):
        for qp in range(0, 52, 13):
            q, mask = M.quant_4x4x4_expect(O.quant, gold["q444_in"], gold["cqm_mf4"][i_cqm, qp], gold["cqm_bias4"][i_cqm, qp])
            np.testing.assert_array_equal(q, gold["q444_out"][k]); np.testing.assert_array_equal(mask, gold["q444_mask"][k])
            k += 1
