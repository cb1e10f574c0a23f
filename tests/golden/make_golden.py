"""Generates the golden vectors under tests/golden/ from the UNMODIFIED reference build
(oracle/_ref/libx264ref.so, compiled from /root/reference by oracle/Makefile).

Run here (the container that has /root/reference):   python tests/golden/make_golden.py
The fixtures are small .npz files; the oracle and the CUDA kernels are both tested against them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ctypes as C  # noqa: E402

from checkasm_data import PBUF2_OFF, SIZE_WH, dct_overflow_blocks, make_buffers, pixel_cases  # noqa: E402
from checkers import (GET_INTRA, GET_LOWRES_COSTS, GET_MV_COSTS, GET_MVS, GET_ROW_SATDS, GET_STATS, Lookahead, RefLib,  # noqa: E402
                      synth_frames)


def main():
    ref = RefLib()
    pbuf1, pbuf3, pbuf4 = make_buffers(1234)
    out = {"pbuf1": pbuf1, "pbuf3": pbuf3, "pbuf4": pbuf4}

    # ---- pixel metrics: checkasm TEST_PIXEL (checkasm.c:384-423) ----
    for kind in ("sad", "satd", "sa8d"):
        for size in range(8):
            if kind == "sa8d" and size not in (0, 3):
                continue
            res = []
            for (which, s1, s2, o1, o2) in pixel_cases(size):
                a, b = (pbuf1, pbuf1) if which == "rand" else (pbuf3, pbuf4)
                res.append(ref.pixel_cmp(kind, size, a, s1, b, s2, o1, o2))
            out[f"{kind}_{size}"] = np.concatenate(res)

    # ---- hadamard_ac: 32 positions over random + overflow buffers (checkasm.c:551-574) ----
    for size in range(4):
        offs = np.arange(32, dtype=np.int64)
        r1 = ref.hadamard_ac(size, pbuf1, 16, offs * 30)
        r2 = ref.hadamard_ac(size, pbuf3, 16, offs * 100)
        r3 = ref.hadamard_ac(size, pbuf4, 32, offs * 64)
        out[f"hadamard_ac_{size}"] = np.concatenate([r1, r2, r3])

    # ---- sad_x3 / sad_x4 (checkasm.c:462-502): refs at pix2, +6, +1, +10, stride 64 ----
    for size in range(7):
        w, h = SIZE_WH[size]
        fenc = np.ascontiguousarray(pbuf1[:16 * 16])
        sc = (C.c_int * 4)()
        base = pbuf1.ctypes.data + PBUF2_OFF
        ref.lib.ref_sad_x4(size, fenc.ctypes.data, base, base + 6, base + 1, base + 10, 64, sc)
        out[f"sad_x4_{size}"] = np.array(list(sc), dtype=np.int32)

    # ---- dct: random + overflow blocks (checkasm.c:908-976) ----
    rng = np.random.default_rng(77)
    a = rng.integers(0, 256, (16 * 12, 64), dtype=np.uint8)
    b = rng.integers(0, 256, (16 * 12, 64), dtype=np.uint8)
    for k, (pa, pb) in enumerate(dct_overflow_blocks()):
        a[16 * (6 + k):16 * (7 + k), :16] = pa
        b[16 * (6 + k):16 * (7 + k), :16] = pb
    out["dct_a"], out["dct_b"] = a, b
    offs = np.array([y * 16 * 64 + x for y in range(12) for x in (0, 16, 32)], dtype=np.int64)
    out["dct_offs"] = offs
    for bs in (4, 8, 16):
        out[f"dct_{bs}"] = ref.sub_dct(bs, a.reshape(-1), 64, b.reshape(-1), 64, offs, offs)

    # ---- quant: flat CQM tables from x264_cqm_init at several QPs (checkasm.c:2078-2227) ----
    qps = [0, 12, 20, 26, 38, 51]
    out["quant_qps"] = np.array(qps)
    for b8 in (0, 1):
        n = 64 if b8 else 16
        for cat in (0, 1):
            mfs, biases, coefs_in, coefs_out, nzs = [], [], [], [], []
            for qp in qps:
                mf, bias = ref.cqm_tables(b8, cat, qp)
                # coefficient magnitudes spanning the dead zone up to the DCT's dynamic range
                c = rng.integers(-32768, 32768, (24, n)).astype(np.int32)
                scale = np.array([1, 4, 16, 64, 256, 2048] * 4)[:, None]
                c = (c // scale).astype(np.int16)
                c[0] = 0
                c[1] = 32767
                c[2] = -32768
                q, nz = ref.quant(b8, c, mf, bias)
                mfs.append(mf); biases.append(bias); coefs_in.append(c); coefs_out.append(q.reshape(c.shape)); nzs.append(nz)
            tag = f"quant{'8' if b8 else '4'}_{cat}"
            out[tag + "_mf"] = np.array(mfs); out[tag + "_bias"] = np.array(biases)
            out[tag + "_in"] = np.array(coefs_in); out[tag + "_out"] = np.array(coefs_out); out[tag + "_nz"] = np.array(nzs)

    # ---- lowres core: w = 96..120 step 8, h = 8 (checkasm.c:1703-1733) ----
    src = rng.integers(0, 256, (17, 256), dtype=np.uint8)
    out["lowres_src"] = src
    for w in (96, 104, 112, 120):
        planes = ref.lowres_core(src.reshape(-1), 256, w, 8, 128)
        out[f"lowres_{w}"] = np.stack(planes)

    # ---- mc_luma / get_ref over all 16 qpel phases, with and without weights (checkasm.c:1234-1429) ----
    planes = [rng.integers(0, 256, 64 * 64, dtype=np.uint8) for _ in range(4)]
    out["mc_planes"] = np.stack(planes)
    mc = []
    for (w, h) in ((8, 8), (8, 9), (12, 8)):
        for mvy in range(-5, 7):
            for mvx in range(-5, 7):
                for wt in (None, (58, 6, 3), (127, 7, -10), (3, 0, 1)):
                    mc.append(ref.mc_luma(planes, 64, 24 * 64 + 24, mvx, mvy, w, h, wt).reshape(-1))
    out["mc_out"] = np.concatenate(mc)
    avg = []
    for wgt in (32, 16, 48, 64 - 43, 0, 64):
        avg.append(ref.avg(3, src[:8, 10:18].copy(), src[8:16, 33:41].copy(), wgt).reshape(-1))
    out["avg_out"] = np.concatenate(avg)

    # ---- intra cost of single blocks, both flavours, recomposed from the reference's own tables ----
    pl = rng.integers(0, 256, 64 * 64, dtype=np.uint8)
    smooth = (np.add.outer(np.arange(64) * 3, np.arange(64) * 2) % 256).astype(np.uint8).reshape(-1)
    out["intra_planes"] = np.stack([pl, smooth])
    res = []
    for plane in (pl, smooth):
        for (x, y) in ((8, 8), (16, 8), (24, 40), (9, 13), (40, 17)):
            for b_satd in (0, 1):
                res.append(ref.intra_cost_8x8(plane, 64, y * 64 + x, b_satd))
    out["intra_out"] = np.array(res, dtype=np.int32)

    # ---- mv cost table (analyse.c:143-157) ----
    la = Lookahead(ref, "ref", 96, 64, bframes=3)
    n = 8 * la.mv_range
    tab = np.zeros(2 * n + 1, dtype=np.uint16)
    ref.lib.ref_la_cost_mv(la.ctx, tab.ctypes.data, n)
    out["cost_mv"] = tab
    out["cost_mv_range"] = np.array([la.mv_range])
    la.close()
    np.savez_compressed(os.path.join(HERE, "primitives.npz"), **out)

    # ---- lookahead: the unmodified slicetype_mb_cost in the dependency-free schedule ----
    look = {}
    cases = [("satd_edges", 96, 64, 7, 1, 1, 2), ("sad_noedges", 96, 64, 1, 0, 1, 2), ("satd_3pass", 112, 80, 7, 1, 0, 3)]
    triples = [(0, 0, 0), (0, 1, 1), (0, 3, 3), (0, 3, 1), (0, 3, 2), (1, 4, 4), (1, 4, 2), (1, 4, 3), (2, 2, 2)]
    look["triples"] = np.array(triples)
    for (name, W, H, subme, mbtree, aq, passes) in cases:
        frames = synth_frames(W, H, 5, seed=21, pan=(2, 1))
        r2 = np.random.default_rng(3)
        la = Lookahead(ref, "ref", W, H, bframes=3, subme=subme, mbtree=mbtree, aq=aq, passes=passes)
        look[name + "_cfg"] = np.array([W, H, subme, mbtree, aq, passes, la.mv_range])
        look[name + "_frames"] = np.stack(frames)
        qs = []
        for i, f in enumerate(frames):
            q = r2.integers(128, 512, la.nmb).astype(np.uint16)
            qs.append(q)
            la.put_frame(i, f, q if aq else None)
        look[name + "_invq"] = np.stack(qs)
        look[name + "_lowres"] = np.stack([la.lowres(1, p) for p in range(4)])
        for t, (p0, p1, b) in enumerate(triples):
            sc = la.frame_cost(p0, p1, b)
            i, j = b - p0, p1 - b
            look[f"{name}_{t}_score"] = np.array([sc])
            look[f"{name}_{t}_costs"] = la.get(b, GET_LOWRES_COSTS, i, j)
            look[f"{name}_{t}_rows"] = la.get(b, GET_ROW_SATDS, i, j)
            look[f"{name}_{t}_stats"] = la.get(b, GET_STATS, i, j)
            look[f"{name}_{t}_intra"] = la.get(b, GET_INTRA)
            if i:
                look[f"{name}_{t}_mv0"] = la.get(b, GET_MVS, 0, i - 1)
                look[f"{name}_{t}_mc0"] = la.get(b, GET_MV_COSTS, 0, i - 1)
            if j:
                look[f"{name}_{t}_mv1"] = la.get(b, GET_MVS, 1, j - 1)
                look[f"{name}_{t}_mc1"] = la.get(b, GET_MV_COSTS, 1, j - 1)
        la.close()
    np.savez_compressed(os.path.join(HERE, "lookahead.npz"), **look)
    for f in ("primitives.npz", "lookahead.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
