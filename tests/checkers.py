"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE):

  RefLib     oracle/_ref/libx264ref.so  -- the unmodified reference C sources + oracle/ref_shim.c
  OracleLib  oracle/libx264oracle.so    -- our plain-C restatement (oracle/x264_oracle.c)

Both expose the same lookahead driver interface (`Lookahead`) so a test can run the same
evaluation sequence through either and compare arrays.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libx264oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libx264ref.so")

SIZES = ["16x16", "16x8", "8x16", "8x8", "8x4", "4x8", "4x4", "4x16"]
SIZE_WH = [(16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4), (4, 16)]
KIND = {"sad": 0, "satd": 1, "sa8d": 2}

vp = C.c_void_p


def _ptr(a):
    return a.ctypes.data_as(vp)


def build_oracle():
    src = os.path.join(ROOT, "oracle", "x264_oracle.c")
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return ORACLE_SO


def have_ref():
    return os.path.exists(REF_SO)


class XoCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "width", "height", "bframes", "subme", "me_method", "me_range", "mv_range", "do_edges",
        "aq", "weighted_bipred", "bframe_bias", "search_passes")]


class _Base:
    """primitive calls common to both libraries (prefix differs)."""

    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.p = prefix
        L = self.lib
        f = getattr(L, prefix + "pixel_cmp_batch")
        f.argtypes = [C.c_int, C.c_int, vp, C.c_ssize_t, vp, C.c_ssize_t, C.c_int64, vp, vp, vp]
        f.restype = C.c_int
        f = getattr(L, prefix + "hadamard_ac_batch")
        f.argtypes = [C.c_int, vp, C.c_ssize_t, C.c_int64, vp, vp]
        f = getattr(L, prefix + "sub_dct_batch")
        f.argtypes = [C.c_int, vp, C.c_ssize_t, vp, C.c_ssize_t, C.c_int64, vp, vp, vp]
        f = getattr(L, prefix + "quant_batch")
        f.argtypes = [C.c_int, vp, C.c_int64, vp, vp, vp]
        f = getattr(L, prefix + "lowres_core")
        f.argtypes = [vp, C.c_ssize_t, vp, vp, vp, vp, C.c_ssize_t, C.c_int, C.c_int]
        f.restype = None

    def pixel_cmp(self, kind, size, p1, s1, p2, s2, off1, off2):
        """p1/p2: flat uint8 arrays; off*: int64 byte offsets of each block's top-left."""
        off1 = np.ascontiguousarray(off1, dtype=np.int64)
        off2 = np.ascontiguousarray(off2, dtype=np.int64)
        out = np.empty(len(off1), dtype=np.int32)
        r = getattr(self.lib, self.p + "pixel_cmp_batch")(KIND[kind], size, _ptr(p1), s1, _ptr(p2), s2,
                                                        len(off1), _ptr(off1), _ptr(off2), _ptr(out))
        assert r == 0
        return out

    def hadamard_ac(self, size, p, stride, off):
        off = np.ascontiguousarray(off, dtype=np.int64)
        out = np.empty(len(off), dtype=np.uint64)
        r = getattr(self.lib, self.p + "hadamard_ac_batch")(size, _ptr(p), stride, len(off), _ptr(off), _ptr(out))
        assert r == 0
        return out

    def sub_dct(self, bsize, p1, s1, p2, s2, off1, off2):
        off1 = np.ascontiguousarray(off1, dtype=np.int64)
        off2 = np.ascontiguousarray(off2, dtype=np.int64)
        out = np.empty((len(off1), bsize * bsize), dtype=np.int16)
        r = getattr(self.lib, self.p + "sub_dct_batch")(bsize, _ptr(p1), s1, _ptr(p2), s2, len(off1),
                                                      _ptr(off1), _ptr(off2), _ptr(out))
        assert r == 0
        return out

    def quant(self, b8x8, coefs, mf, bias):
        n = 64 if b8x8 else 16
        coefs = np.ascontiguousarray(coefs, dtype=np.int16).copy()
        nb = coefs.size // n
        mf = np.ascontiguousarray(mf, dtype=np.uint16)
        bias = np.ascontiguousarray(bias, dtype=np.uint16)
        nz = np.empty(nb, dtype=np.int32)
        getattr(self.lib, self.p + "quant_batch")(int(b8x8), _ptr(coefs), nb, _ptr(mf), _ptr(bias), _ptr(nz))
        return coefs, nz

    def lowres_core(self, src, sstride, width, height, dstride):
        """src: flat uint8 with >= (2*height+1) rows of >= 2*width+1 valid bytes. returns 4 planes."""
        planes = [np.zeros(dstride * height, dtype=np.uint8) for _ in range(4)]
        getattr(self.lib, self.p + "lowres_core")(_ptr(src), sstride, *[_ptr(p) for p in planes], dstride, width, height)
        return [p.reshape(height, dstride)[:, :width].copy() for p in planes]


class OracleLib(_Base):
    def __init__(self):
        super().__init__(build_oracle(), "xo_")
        L = self.lib
        L.xo_open.restype = vp
        L.xo_open.argtypes = [C.POINTER(XoCfg)]
        L.xo_close.argtypes = [vp]
        L.xo_info.argtypes = [vp, vp]
        L.xo_put_frame.argtypes = [vp, C.c_int, vp, C.c_int, vp]
        L.xo_get_lowres.argtypes = [vp, C.c_int, C.c_int, vp]
        L.xo_frame_cost.argtypes = [vp] + [C.c_int] * 6
        L.xo_get.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.xo_counters.argtypes = [vp, vp]
        L.xo_mc_luma.argtypes = [vp, C.c_ssize_t, vp, C.c_ssize_t] + [C.c_int] * 7
        L.xo_mc_luma.restype = None
        L.xo_avg.argtypes = [vp, C.c_ssize_t, vp, C.c_ssize_t, vp, C.c_ssize_t, C.c_int, C.c_int, C.c_int]
        L.xo_avg.restype = None
        L.xo_intra_cost_8x8.argtypes = [vp, C.c_ssize_t, C.c_int]
        L.xo_cost_mv_table.argtypes = [vp, C.c_int]
        L.xo_cost_mv_table.restype = None

    def cost_mv_table(self, n):
        buf = np.zeros(2 * n + 1, dtype=np.uint16)
        self.lib.xo_cost_mv_table(C.c_void_p(buf.ctypes.data + 2 * n), n)
        return buf

    def mc_luma(self, planes, stride, org, mvx, mvy, w, h, weight=None):
        """planes: 4 flat uint8 arrays; org: offset of the block origin inside each plane."""
        arr = (vp * 4)(*[p.ctypes.data + org for p in planes])
        dst = np.zeros((h, w), dtype=np.uint8)
        sc, dn, of = weight if weight else (0, -1, 0)
        self.lib.xo_mc_luma(_ptr(dst), w, arr, stride, mvx, mvy, w, h, sc, dn, of)
        return dst

    def avg(self, a, b, w, h, weight):
        dst = np.zeros((h, w), dtype=np.uint8)
        self.lib.xo_avg(_ptr(dst), w, _ptr(a), a.shape[1], _ptr(b), b.shape[1], w, h, weight)
        return dst


def ref_flavours():
    """the builds of the unmodified reference that oracle/Makefile produced: '' = parity (-O2), '_fast' = the
    reference's release flags (-O3 -ffast-math -fno-tree-vectorize), '_simd' = the same auto-vectorised for AVX2."""
    out = {}
    for suf in ("", "_fast", "_simd"):
        p = REF_SO.replace(".so", suf + ".so")
        if os.path.exists(p):
            out[suf] = p
    if "_simd" in out:
        try:
            flags = open("/proc/cpuinfo").read()
            if not all(f in flags for f in (" avx2", " fma", " bmi2")):
                del out["_simd"]
        except OSError:
            del out["_simd"]
    return out


class RefLib(_Base):
    def __init__(self, path=None):
        super().__init__(path or REF_SO, "ref_")
        L = self.lib
        L.ref_la_open.restype = vp
        L.ref_la_open.argtypes = [C.c_int] * 11
        L.ref_la_close.argtypes = [vp]
        L.ref_la_info.argtypes = [vp, vp]
        L.ref_la_cost_mv.argtypes = [vp, vp, C.c_int]
        L.ref_la_put_frame.argtypes = [vp, C.c_int, vp, C.c_int, vp]
        L.ref_la_release.argtypes = [vp, C.c_int]
        L.ref_la_get_lowres.argtypes = [vp, C.c_int, C.c_int, vp]
        L.ref_la_lowres_geometry.argtypes = [vp, C.c_int, vp]
        L.ref_la_frame_cost_serial.argtypes = [vp] + [C.c_int] * 3
        L.ref_la_frame_cost_jacobi.argtypes = [vp] + [C.c_int] * 4
        L.ref_la_get.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.ref_la_slicetype_analyse.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
        L.ref_mc_luma.argtypes = [vp, C.c_ssize_t, vp, vp, vp, vp, C.c_ssize_t] + [C.c_int] * 7
        L.ref_mc_luma.restype = None
        L.ref_avg.argtypes = [C.c_int, vp, C.c_ssize_t, vp, C.c_ssize_t, vp, C.c_ssize_t, C.c_int]
        L.ref_avg.restype = None
        L.ref_cqm_tables.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp]
        L.ref_intra_x3_8x8c.argtypes = [C.c_int, vp, vp, vp]
        L.ref_predict_8x8c.argtypes = [C.c_int, vp]
        L.ref_predict_8x8_filter.argtypes = [vp, vp, C.c_int, C.c_int]
        L.ref_predict_8x8.argtypes = [C.c_int, vp, vp]
        L.ref_pixel_cmp.argtypes = [C.c_int, C.c_int, vp, C.c_ssize_t, vp, C.c_ssize_t]
        L.ref_sad_x3.argtypes = [C.c_int, vp, vp, vp, vp, C.c_ssize_t, vp]
        L.ref_sad_x4.argtypes = [C.c_int, vp, vp, vp, vp, vp, C.c_ssize_t, vp]
        L.ref_now.restype = C.c_double

    def mc_luma(self, planes, stride, org, mvx, mvy, w, h, weight=None):
        dst = np.zeros((h, w), dtype=np.uint8)
        sc, dn, of = weight if weight else (0, -1, 0)
        self.lib.ref_mc_luma(_ptr(dst), w, *[C.c_void_p(p.ctypes.data + org) for p in planes], stride,
                             mvx, mvy, w, h, sc, dn, of)
        return dst

    def avg(self, size, a, b, weight):
        w, h = SIZE_WH[size]
        dst = np.zeros((h, w), dtype=np.uint8)
        self.lib.ref_avg(size, _ptr(dst), w, _ptr(a), a.shape[1], _ptr(b), b.shape[1], weight)
        return dst

    def cqm_tables(self, b8x8, cat, qp):
        n = 64 if b8x8 else 16
        mf = np.zeros(n, dtype=np.uint16)
        bias = np.zeros(n, dtype=np.uint16)
        assert self.lib.ref_cqm_tables(int(b8x8), cat, qp, _ptr(mf), _ptr(bias)) == 0
        return mf, bias

    def intra_cost_8x8(self, plane, stride, org, b_satd):
        """the reference's intra section (slicetype.c:714-746) recomposed from its own predictor
        and pixel tables, for one block at byte offset `org` of a padded plane."""
        FDEC = 32
        L = self.lib
        buf = np.zeros(9 * FDEC + 64, dtype=np.uint8)
        pixo = 8 + FDEC          # pix = &pix1[8+FDEC_STRIDE]
        fenc = np.zeros(16 * 8, dtype=np.uint8)
        for y in range(8):
            fenc[y * 16:y * 16 + 8] = plane[org + y * stride: org + y * stride + 8]
        buf[pixo - FDEC: pixo - FDEC + 16] = plane[org - stride: org - stride + 16]
        for i in range(-1, 8):
            buf[pixo + i * FDEC - 4: pixo + i * FDEC] = plane[org + i * stride - 4: org + i * stride]
        pix = C.c_void_p(buf.ctypes.data + pixo)
        res = (C.c_int * 3)()
        L.ref_intra_x3_8x8c(int(b_satd), _ptr(fenc), pix, res)
        best = min(res)
        if b_satd:
            kind = 1
            L.ref_predict_8x8c(3, pix)   # I_PRED_CHROMA_P
            best = min(best, L.ref_pixel_cmp(kind, 3, _ptr(fenc), 16, pix, FDEC))
            edge = np.zeros(40, dtype=np.uint8)
            L.ref_predict_8x8_filter(pix, _ptr(edge), 0xF, 0xF)   # ALL_NEIGHBORS
            for m in range(3, 9):
                L.ref_predict_8x8(m, pix, _ptr(edge))
                best = min(best, L.ref_pixel_cmp(kind, 3, _ptr(fenc), 16, pix, FDEC))
        return best + 5 + 4


GET_LOWRES_COSTS, GET_MVS, GET_MV_COSTS, GET_ROW_SATDS, GET_STATS, GET_INTRA, GET_INVQ = range(7)


class Lookahead:
    """uniform driver over RefLib ('ref': Jacobi schedule on the unmodified slicetype_mb_cost;
    'ref_serial': the reference's own raster-serial slicetype_frame_cost) and OracleLib ('oracle')."""

    def __init__(self, lib, kind, width, height, bframes=3, subme=7, mbtree=1, aq=1, passes=2,
                 b_adapt=2, rc_lookahead=40, la_threads=1, mv_range=None, weightp=0):
        self.lib, self.kind = lib, kind
        self.W, self.H, self.passes = width, height, passes
        if kind.startswith("ref"):
            self.ctx = lib.lib.ref_la_open(width, height, bframes, subme, mbtree, aq, weightp, b_adapt, la_threads, rc_lookahead, 0)
            assert self.ctx
            info = (C.c_int * 10)()
            lib.lib.ref_la_info(self.ctx, info)
            self.info = list(info)
            self.mbw, self.mbh, self.mv_range = info[0], info[1], info[2]
        else:
            cfg = XoCfg(width, height, bframes, subme, 1, 16, mv_range, mbtree, aq, 1, 0, passes)
            self.ctx = lib.lib.xo_open(C.byref(cfg))
            assert self.ctx
            info = (C.c_int * 4)()
            lib.lib.xo_info(self.ctx, info)
            self.mbw, self.mbh = info[0], info[1]
            self.mv_range = mv_range
        self.nmb = self.mbw * self.mbh

    def close(self):
        if self.ctx:
            (self.lib.lib.ref_la_close if self.kind.startswith("ref") else self.lib.lib.xo_close)(self.ctx)
            self.ctx = None

    def put_frame(self, idx, luma, inv_qscale=None):
        luma = np.ascontiguousarray(luma, dtype=np.uint8)
        assert luma.shape == (self.H, self.W)
        q = None if inv_qscale is None else _ptr(np.ascontiguousarray(inv_qscale, dtype=np.uint16))
        f = self.lib.lib.ref_la_put_frame if self.kind.startswith("ref") else self.lib.lib.xo_put_frame
        assert f(self.ctx, idx, _ptr(luma), self.W, q) == 0

    # ---- adaptive quantisation + weighted-prediction analysis ----
    def put_frame_yuv(self, idx, luma, chroma, aq_mode=1, aq_strength=1.0, quant_offsets=None):
        """x264_adaptive_quant_frame + x264_frame_init_lowres; chroma: interleaved NV12, (H/2, W) uint8"""
        luma = np.ascontiguousarray(luma, dtype=np.uint8)
        chroma = np.ascontiguousarray(chroma, dtype=np.uint8)
        assert luma.shape == (self.H, self.W) and chroma.shape == (self.H // 2, self.W)
        qo = None if quant_offsets is None else _ptr(np.ascontiguousarray(quant_offsets, dtype=np.float32))
        if self.kind.startswith("ref"):      # aq mode / strength are the encoder's parameters (ref_la_open)
            L = self.lib.lib
            L.ref_la_put_frame_yuv.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp]
            assert L.ref_la_put_frame_yuv(self.ctx, idx, _ptr(luma), self.W, _ptr(chroma), self.W, qo) == 0
        else:
            L = self.lib.lib
            L.xo_put_frame_aq.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, C.c_float, vp]
            assert L.xo_put_frame_aq(self.ctx, idx, _ptr(luma), self.W, _ptr(chroma), self.W, aq_mode, aq_strength, qo) == 0

    def get_aq(self, idx):
        qp = np.zeros(self.nmb, dtype=np.float32)
        inv = np.zeros(self.nmb, dtype=np.uint16)
        sums = np.zeros(6, dtype=np.uint64)
        f = self.lib.lib.ref_la_get_aq if self.kind.startswith("ref") else self.lib.lib.xo_get_aq
        f.argtypes = [vp, C.c_int, vp, vp, vp]
        assert f(self.ctx, idx, _ptr(qp), _ptr(inv), _ptr(sums)) == 0
        return qp, inv, sums

    def weights_analyse(self, b, ref, numslices=1):
        """x264_weights_analyse(h, frames[b], frames[ref], 1) -> (enabled, scale, denom, offset)"""
        out = (C.c_int * 6)()
        if self.kind.startswith("ref"):
            self.lib.lib.ref_la_weights_analyse.argtypes = [vp, C.c_int, C.c_int, vp]
            assert self.lib.lib.ref_la_weights_analyse(self.ctx, b, ref, out) == 0
        else:
            self.lib.lib.xo_weights_analyse.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
            assert self.lib.lib.xo_weights_analyse(self.ctx, b, ref, numslices, out) == 0
        return tuple(out[:4]) if out[0] else (0, 1, 0, 0)

    def lowres(self, idx, plane):
        w, h = self.mbw * 8 + 64, self.mbh * 8 + 64
        out = np.zeros((h, w), dtype=np.uint8)
        f = self.lib.lib.ref_la_get_lowres if self.kind.startswith("ref") else self.lib.lib.xo_get_lowres
        assert f(self.ctx, idx, plane, _ptr(out)) == 0
        return out

    def frame_cost(self, p0, p1, b, weight=None):
        """weight (oracle only): (enabled, scale, denom, offset) for list 0 of a P evaluation; the reference analyses
        its own weights inside slicetype_frame_cost when it was opened with weightp"""
        if self.kind == "oracle" and weight is not None and weight[0]:
            return self.lib.lib.xo_frame_cost(self.ctx, p0, p1, b, weight[1], weight[2], weight[3])
        if self.kind == "ref":
            return self.lib.lib.ref_la_frame_cost_jacobi(self.ctx, p0, p1, b, self.passes)
        if self.kind == "ref_serial":
            return self.lib.lib.ref_la_frame_cost_serial(self.ctx, p0, p1, b)
        return self.lib.lib.xo_frame_cost(self.ctx, p0, p1, b, 0, -1, 0)

    def get(self, idx, what, i=0, j=0):
        shape, dt = {
            GET_LOWRES_COSTS: ((self.nmb,), np.uint16), GET_MVS: ((self.nmb, 2), np.int16),
            GET_MV_COSTS: ((self.nmb,), np.int32), GET_ROW_SATDS: ((self.mbh,), np.int32),
            GET_STATS: ((3,), np.int32), GET_INTRA: ((self.nmb,), np.uint16), GET_INVQ: ((self.nmb,), np.uint16)}[what]
        out = np.zeros(shape, dtype=dt)
        f = self.lib.lib.ref_la_get if self.kind.startswith("ref") else self.lib.lib.xo_get
        assert f(self.ctx, idx, what, i, j, _ptr(out)) == 0, (what, i, j)
        return out


# ------------------------------------------------------------------------------------------------
# synthetic content: panning texture + moving square + noise
# ------------------------------------------------------------------------------------------------
def synth_frames(width, height, n, seed=7, pan=(3, 1), noise=3, scene_cut_at=None, fade_at=None):
    rng = np.random.default_rng(seed)
    margin = 64 + max(abs(pan[0]), abs(pan[1])) * n
    tw, th = width + 2 * margin, height + 2 * margin

    def texture(r):
        t = r.integers(0, 256, (th // 8 + 2, tw // 8 + 2)).astype(np.float32)
        t = np.kron(t, np.ones((8, 8), dtype=np.float32))[:th, :tw]
        # cheap separable blur so that sub-pel interpolation matters
        for ax in (0, 1):
            t = (np.roll(t, 1, ax) + 2 * t + np.roll(t, -1, ax)) / 4
            t = (np.roll(t, 2, ax) + 2 * t + np.roll(t, -2, ax)) / 4
        fine = r.integers(-24, 25, (th, tw)).astype(np.float32)
        return np.clip(t * 0.8 + 25 + fine, 0, 255)

    tex = texture(rng)
    tex2 = None
    frames = []
    for i in range(n):
        if scene_cut_at is not None and i == scene_cut_at:
            tex2 = texture(np.random.default_rng(seed + 1000))
        src = tex2 if (tex2 is not None and i >= scene_cut_at) else tex
        ox, oy = margin + pan[0] * i, margin + pan[1] * i
        f = src[oy:oy + height, ox:ox + width].copy()
        sq = min(160, width // 4, height // 4)
        sx, sy = (40 + 11 * i) % max(1, width - sq), (30 + 5 * i) % max(1, height - sq)
        f[sy:sy + sq, sx:sx + sq] = 255 - f[sy:sy + sq, sx:sx + sq]
        if fade_at is not None and fade_at <= i < fade_at + 12:
            f = f * (1.0 - (i - fade_at + 1) / 16.0) + 8
        f = f + rng.integers(-noise, noise + 1, f.shape)
        frames.append(np.clip(f, 0, 255).astype(np.uint8))
    return frames


def synth_chroma(width, height, n, seed=11):
    """interleaved NV12 chroma planes (H/2, W) that follow the luma generator's pan: smooth field + noise"""
    rng = np.random.default_rng(seed)
    base = rng.integers(64, 192, (height // 16 + 4, width // 16 + 4, 2)).astype(np.float32)
    base = np.kron(base, np.ones((8, 8, 1), dtype=np.float32))
    out = []
    for i in range(n):
        c = base[i:i + height // 2, 2 * i:2 * i + width // 2, :]
        c = c + rng.integers(-6, 7, c.shape)
        out.append(np.clip(c, 0, 255).astype(np.uint8).reshape(height // 2, width))
    return out


# ------------------------------------------------------------------------------------------------
# one steady-state step of a lookahead window against the oracle (tests/test_gpu_bench_config.py, bench.py parity_check)
# ------------------------------------------------------------------------------------------------
def stream_step_evaluations(n, bframes, b_list=None):
    """(p0, n, b) of one x264cu_stream_push of frame n in the window's own order (host/stream.cpp): the I cost, then per p0
    from near to far the P evaluation and the B evaluations between p0 and n.  b_list restricts the B evaluations."""
    out = [(n, n, n)]
    for p0 in range(n - 1, max(n - (bframes + 1), 0) - 1, -1):
        out.append((p0, n, n))
        for b in range(p0 + 1, n):
            if b_list is None or b in b_list:
                out.append((p0, n, b))
    return out


def check_step_against_oracle(window_frames, get_stats, get_arrays=None, bframes=16, subme=7, passes=2, mv_range=512,
                              b_list=None, aq=1):
    """window_frames: the frames of the window, oldest first, newest last (index n = len - 1).
    get_stats(d0, d1) -> 4 ints of the CUDA path (cost_est, cost_est_aq, intra_mbs, raw cost);
    get_arrays(d0, d1) -> dict with any of lowres_costs / row_satds / mvs0 / mvs1 (None entries are skipped).
    Runs the same evaluations through the CPU oracle (K-pass dependency-free schedule) and compares bit for bit.
    Returns {"evaluations": ..., "arrays": ..., "mismatches": [...]}."""
    H, W = window_frames[0].shape
    n = len(window_frames) - 1
    O = Lookahead(OracleLib(), "oracle", W, H, bframes=bframes, subme=subme, mbtree=1, aq=aq, passes=passes, mv_range=mv_range)
    bad, narr = [], 0
    try:
        for i, f in enumerate(window_frames):
            O.put_frame(i, f)
        evs = stream_step_evaluations(n, bframes, b_list)
        for (p0, p1, b) in evs:
            d0, d1 = b - p0, p1 - b
            sc = O.frame_cost(p0, p1, b)
            st = O.get(b, GET_STATS, d0, d1)
            g = get_stats(d0, d1)
            if int(g[0]) != int(sc) or int(g[1]) != int(st[1]) or (d1 == 0 and d0 > 0 and int(g[2]) != int(st[2])):
                bad.append(("stats", d0, d1, [int(x) for x in g], int(sc), [int(x) for x in st]))
            if get_arrays is None:
                continue
            a = get_arrays(d0, d1) or {}
            want = {"lowres_costs": lambda: O.get(b, GET_LOWRES_COSTS, d0, d1) if d0 + d1 else O.get(b, GET_INTRA),
                    "intra_costs": lambda: O.get(b, GET_INTRA),
                    "row_satds": lambda: O.get(b, GET_ROW_SATDS, d0, d1),
                    "mvs0": lambda: O.get(b, GET_MVS, 0, d0 - 1), "mvs1": lambda: O.get(b, GET_MVS, 1, d1 - 1)}
            for k, v in a.items():
                if v is None:
                    continue
                narr += 1
                if not np.array_equal(np.asarray(v).reshape(-1), want[k]().reshape(-1)):
                    bad.append((k, d0, d1))
    finally:
        O.close()
    return {"evaluations": len(evs), "arrays": narr, "mismatches": bad}
