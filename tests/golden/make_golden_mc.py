"""Golden vectors of the checkasm mc / intra / quant surfaces from the UNMODIFIED reference build (oracle/_ref):
tests/golden/mc_surfaces.npz.  Inputs are regenerated from seeds by tests/mc_cases.py; only the reference's outputs
and its CQM tables (x264_cqm_init under six configurations x QP 0..51) are stored.
Run in the container that has /root/reference:   python tests/golden/make_golden_mc.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import mc_cases as M  # noqa: E402
from checkers import RefLib  # noqa: E402

GOLD_WEIGHTS = [None, (37, 5, -7), (1, 0, 40), (5, 0, 3)]


def golden_inputs():
    """the deterministic inputs shared by the generator and the tests"""
    buf, po, stride, psz = M.planes_buffer()
    rng = np.random.default_rng(3)
    mc = {}
    for (w, h) in M.MC_SIZES:
        org, mv = M.mc_cases(stride, psz // stride, w, h, rng)
        mc[(w, h)] = (org[:16], mv[:16])
    rng = np.random.default_rng(4)
    a = rng.integers(0, 256, 64 * 40, dtype=np.uint8)
    b = rng.integers(0, 256, 48 * 40, dtype=np.uint8)
    a[:300] = 255; b[:300] = 0
    off1 = np.array([0, 5, 64 * 3 + 17, 64 * 20 + 1], dtype=np.int64)
    off2 = np.array([1, 48 * 2, 48 * 7 + 9, 48 * 19 + 30], dtype=np.int64)
    rng = np.random.default_rng(6)
    s1, s2 = 80, 96
    p1 = rng.integers(0, 256, s1 * 48, dtype=np.uint8)
    p2 = rng.integers(0, 256, s2 * 48, dtype=np.uint8)
    p2[s2 * 10:s2 * 14] = 255; p1[s1 * 10:s1 * 20] = 0
    io1 = np.array([s1 * y + x for y in (1, 9, 17, 30) for x in (1, 8, 33, 70)], dtype=np.int64)
    io2 = np.array([s2 * y + x for y in (1, 10, 13, 31) for x in (1, 9, 40, 85)], dtype=np.int64)
    rng = np.random.default_rng(5)
    wsrc = rng.integers(0, 256, 32 * 16 + 32, dtype=np.uint8)
    wsrc[:40] = 255; wsrc[40:80] = 0
    return dict(buf=buf, po=po, stride=stride, mc=mc, a=a, b=b, off1=off1, off2=off2, p1=p1, p2=p2, s1=s1, s2=s2, io1=io1, io2=io2, wsrc=wsrc)


AVG_WEIGHTS = [-63, -20, 0, 7, 32, 33, 64, 100, 127]
W_CASES = [(s, d, o) for s in (0, 1, 31, 127) for d in (0, 3, 7) for o in (-128, 0, 77)]


def main():
    R = RefLib()
    G = golden_inputs()
    out = {}
    for (w, h), (org, mv) in G["mc"].items():
        for k, wt in enumerate(GOLD_WEIGHTS):
            out["mc_%dx%d_%d" % (w, h, k)] = M.ref_mc_luma(R, G["buf"], G["po"], G["stride"], org, mv, w, h, wt)
    for size in range(12):
        for weight in AVG_WEIGHTS:
            out["avg_%d_%d" % (size, weight)] = M.ref_avg(R, G["a"], 64, G["b"], 48, G["off1"], G["off2"], size, weight)
    for k, wt in enumerate(W_CASES):
        out["weight_%d" % k] = M.ref_weight(R, G["wsrc"], 32, 20, 16, wt)
    for b_satd in (0, 1):
        out["intra_x3_%d" % b_satd] = M.ref_intra_x3(R, b_satd, G["p1"], G["s1"], G["p2"], G["s2"], G["io1"], G["io2"])
    L = R.lib
    L.ref_cqm_select.argtypes = [C.c_int, C.c_uint]
    L.ref_quant_4x4x4_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    mf4 = np.zeros((6, 52, 16), np.uint16); b4 = np.zeros_like(mf4)
    mf8 = np.zeros((6, 52, 64), np.uint16); b8 = np.zeros_like(mf8)
    qg, qm = [], []
    rng = np.random.default_rng(7)
    coefs = M.quant_coefs(rng, 4, 0xD)
    for i_cqm in range(6):
        L.ref_cqm_select(i_cqm, 1234)
        for qp in range(52):
            mf4[i_cqm, qp], b4[i_cqm, qp] = R.cqm_tables(0, 0, qp)
            mf8[i_cqm, qp], b8[i_cqm, qp] = R.cqm_tables(1, 0, qp)
            if qp % 13 == 0:
                c = coefs.copy()
                nza = np.zeros(4, dtype=np.int32)
                L.ref_quant_4x4x4_batch(c.ctypes.data, 4, mf4[i_cqm, qp].ctypes.data, b4[i_cqm, qp].ctypes.data, nza.ctypes.data)
                qg.append(c); qm.append(nza)
    L.ref_cqm_select(-1, 0)
    out.update(cqm_mf4=mf4, cqm_bias4=b4, cqm_mf8=mf8, cqm_bias8=b8, q444_in=coefs, q444_out=np.stack(qg), q444_mask=np.stack(qm))
    path = os.path.join(HERE, "mc_surfaces.npz")
    np.savez_compressed(path, **out)
    print("mc_surfaces.npz", os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
