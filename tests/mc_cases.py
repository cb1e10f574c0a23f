"""Shared case generators + CPU evaluators for the checkasm surfaces of the fetch / prediction code
(tools/checkasm.c:754, 1234-1429, 2098-2227): mc_luma == get_ref over the 16 quarter-pel phases x sizes with and without
weights (incl. denom == 0), avg over 12 sizes x weights -63..127, mc_weight, intra_x3_8x8c, quant_4x4x4, quant over six
CQM configurations x QP 0..51.  TEST INFRASTRUCTURE."""
import ctypes as C

import numpy as np

vp = C.c_void_p
AVG_WH = [(16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4), (4, 16), (4, 2), (2, 8), (2, 4), (2, 2)]
MC_SIZES = [(20, 18), (16, 16), (16, 8), (12, 10), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4)]      # checkasm.c:1299-1307
WEIGHTS = [None, (64, 6, 0), (37, 5, -7), (91, 6, 12), (3, 1, -128), (127, 7, 127), (1, 0, 40), (1, 0, -40), (5, 0, 3), (0, 3, 9)]


def planes_buffer(seed=11, stride=96, rows=64):
    """four 'phase planes' inside one random buffer, like src2[4] in check_mc"""
    rng = np.random.default_rng(seed)
    psz = stride * rows
    buf = rng.integers(0, 256, 4 * psz + 64, dtype=np.uint8)
    buf[:psz // 7] = 255
    buf[psz:psz + psz // 9] = 0
    return buf, [k * psz for k in range(4)], stride, psz


def mc_cases(stride, rows, w, h, rng, n=48):
    """block origins + quarter-pel vectors covering all 16 phases, negative and positive displacements"""
    org, mv = [], []
    for k in range(n):
        ox, oy = int(rng.integers(12, stride - w - 14)), int(rng.integers(10, rows - h - 12))
        mvx = int(rng.integers(-32, 33)) if k >= 16 else (k & 3) + 4 * int(rng.integers(-6, 7))
        mvy = int(rng.integers(-24, 25)) if k >= 16 else (k >> 2) + 4 * int(rng.integers(-5, 6))
        org.append(oy * stride + ox)
        mv.append((mvx, mvy))
    return np.array(org, dtype=np.int64), np.array(mv, dtype=np.int16)


def oracle_mc_luma(O, buf, po, stride, org, mv, w, h, weight):
    out = np.zeros((len(org), h, w), dtype=np.uint8)
    planes = [buf[p:] for p in po]
    for i in range(len(org)):
        out[i] = O.mc_luma(planes, stride, int(org[i]), int(mv[i][0]), int(mv[i][1]), w, h, weight)
    return out


def ref_mc_luma(R, buf, po, stride, org, mv, w, h, weight):
    out = np.zeros((len(org), h, w), dtype=np.uint8)
    planes = [buf[p:] for p in po]
    for i in range(len(org)):
        out[i] = R.mc_luma(planes, stride, int(org[i]), int(mv[i][0]), int(mv[i][1]), w, h, weight)
    return out


def oracle_avg(O, a, sa, b, sb, off1, off2, size, weight):
    w, h = AVG_WH[size]
    out = np.zeros((len(off1), h, w), dtype=np.uint8)
    for i in range(len(off1)):
        d = np.zeros((h, w), dtype=np.uint8)
        O.lib.xo_avg(d.ctypes.data_as(vp), w, C.c_void_p(a.ctypes.data + int(off1[i])), sa, C.c_void_p(b.ctypes.data + int(off2[i])), sb, w, h, weight)
        out[i] = d
    return out


def ref_avg(R, a, sa, b, sb, off1, off2, size, weight):
    w, h = AVG_WH[size]
    out = np.zeros((len(off1), h, w), dtype=np.uint8)
    for i in range(len(off1)):
        d = np.zeros((h, w), dtype=np.uint8)
        R.lib.ref_avg(size, d.ctypes.data_as(vp), w, C.c_void_p(a.ctypes.data + int(off1[i])), sa, C.c_void_p(b.ctypes.data + int(off2[i])), sb, weight)
        out[i] = d
    return out


def oracle_weight(O, src, ss, width, height, wt):
    O.lib.xo_mc_weight.argtypes = [vp, C.c_ssize_t, vp, C.c_ssize_t] + [C.c_int] * 5
    O.lib.xo_mc_weight.restype = None
    d = np.zeros((height, width), dtype=np.uint8)
    O.lib.xo_mc_weight(d.ctypes.data_as(vp), width, src.ctypes.data_as(vp), ss, width, height, *wt)
    return d


def ref_weight(R, src, ss, width, height, wt):
    """width must be one of 4, 8, 12, 16, 20 (x264_mc_functions_t.weight[1..5])"""
    R.lib.ref_mc_weight.argtypes = [vp, C.c_ssize_t, vp, C.c_ssize_t] + [C.c_int] * 5
    R.lib.ref_mc_weight.restype = None
    d = np.zeros((height, 32), dtype=np.uint8)
    R.lib.ref_mc_weight(d.ctypes.data_as(vp), 32, src.ctypes.data_as(vp), ss, width // 4, height, *wt)
    return d[:, :width].copy()


def oracle_intra_x3(O, b_satd, p1, s1, p2, s2, off1, off2):
    O.lib.xo_intra_x3_8x8c_batch.argtypes = [C.c_int, vp, C.c_ssize_t, vp, C.c_ssize_t, C.c_int64, vp, vp, vp]
    res = np.zeros((len(off1), 3), dtype=np.int32)
    assert O.lib.xo_intra_x3_8x8c_batch(int(b_satd), p1.ctypes.data_as(vp), s1, p2.ctypes.data_as(vp), s2, len(off1),
                                        off1.ctypes.data_as(vp), off2.ctypes.data_as(vp), res.ctypes.data_as(vp)) == 0
    return res


def ref_intra_x3(R, b_satd, p1, s1, p2, s2, off1, off2):
    """the reference's intra_{sad,satd}_x3_8x8c on FENC_STRIDE / FDEC_STRIDE scratch copies of each block + neighbours"""
    FDEC = 32
    res = np.zeros((len(off1), 3), dtype=np.int32)
    for i in range(len(off1)):
        fenc = np.zeros(16 * 8, dtype=np.uint8)
        fdec = np.zeros(10 * FDEC + 64, dtype=np.uint8)
        o1, o2, po = int(off1[i]), int(off2[i]), 8 + FDEC
        for y in range(8):
            fenc[y * 16:y * 16 + 8] = p1[o1 + y * s1:o1 + y * s1 + 8]
        fdec[po - FDEC - 1:po - FDEC + 8] = p2[o2 - s2 - 1:o2 - s2 + 8]
        for y in range(8):
            fdec[po + y * FDEC - 1] = p2[o2 + y * s2 - 1]
        r = (C.c_int * 3)()
        R.lib.ref_intra_x3_8x8c(int(b_satd), fenc.ctypes.data_as(vp), C.c_void_p(fdec.ctypes.data + po), r)
        res[i] = list(r)
    return res


def quant_coefs(rng, ngroups, j_mask=0xF):
    """INIT_QUANT4 (checkasm.c:2153-2161) for groups of four 4x4 blocks; block k of a group is zero unless bit k of j_mask"""
    scale1d = np.array([4, 6, 4, 6])
    scale = (255 * scale1d[:, None] * scale1d[None, :]).reshape(16)
    c = np.zeros((ngroups, 4, 16), dtype=np.int16)
    for k in range(4):
        if (j_mask >> k) & 1:
            c[:, k, :] = (rng.integers(0, 1 << 30, (ngroups, 16)) % (2 * scale + 1)) - scale
    return c


def quant_4x4x4_expect(lib_quant, coefs, mf, bias):
    """quant_4x4x4 == four quant_4x4 + the mask (quant.c:75-86), through a checker's quant_4x4"""
    flat = coefs.reshape(-1, 16)
    q, nz = lib_quant(0, flat, mf, bias)
    mask = (nz.reshape(-1, 4) != 0).astype(np.int32) @ np.array([1, 2, 4, 8], dtype=np.int32)
    return q.reshape(coefs.shape), mask.astype(np.int32)
