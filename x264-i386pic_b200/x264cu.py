"""ctypes binding of libx264cuda.so (include/x264cu.h) for the tests, smoke() and bench.py.

This is plumbing only: every call goes straight through the C ABI.  Loading fails loudly when the
library has not been built -- there is no CPU fallback in the product.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("X264CU_LIB", os.path.join(HERE, "libx264cuda.so"))     # X264CU_LIB: tuning builds only

vp = C.c_void_p
BFRAME_MAX = 16
SAD, SATD, SA8D = 0, 1, 2
SIZES = ["16x16", "16x8", "8x16", "8x8", "8x4", "4x8", "4x4", "4x16"]
SIZE_WH = [(16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4), (4, 16)]


class Cfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "width", "height", "bframes", "subme", "me_method", "me_range", "mv_range", "do_edges",
        "aq", "weighted_bipred", "bframe_bias", "search_passes", "max_frames")] + [("cost_mv", vp)]


class Weight(C.Structure):
    _fields_ = [("enabled", C.c_int), ("scale", C.c_int), ("denom", C.c_int), ("offset", C.c_int)]


class Out(C.Structure):
    _fields_ = [("lowres_costs", vp), ("mvs", vp * 2), ("mv_costs", vp * 2), ("row_satds", vp),
                ("stats", vp), ("intra_costs", vp), ("intra_row_satds", vp), ("intra_stats", vp)]


class Job(C.Structure):
    _fields_ = [("fenc", vp), ("fref0", vp), ("fref1", vp), ("d0", C.c_int), ("d1", C.c_int),
                ("weight", Weight), ("out", Out)]


class SearchJob(C.Structure):
    _fields_ = [("fenc", vp), ("fref", vp), ("list", C.c_int), ("dist", C.c_int), ("weight", Weight)]


class AqCfg(C.Structure):
    _fields_ = [("mode", C.c_int), ("strength", C.c_float)]


class AqOut(C.Structure):
    _fields_ = [("qp_offset", vp), ("inv_qscale", vp), ("pixel_stats", vp)]


WEIGHT_DEVICE = 2


def _ptr(a):
    return a.ctypes.data_as(vp) if a is not None else None


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libx264cuda.so is not built (run __graft_entry__.build()); no CPU fallback exists")
    L = C.CDLL(LIB_PATH)
    L.x264cu_last_error.restype = C.c_char_p
    L.x264cu_open.restype = vp
    L.x264cu_open.argtypes = [C.c_int, C.POINTER(Cfg)]
    L.x264cu_close.argtypes = [vp]
    L.x264cu_info.argtypes = [vp, vp]
    L.x264cu_frame_new.restype = vp
    L.x264cu_frame_new.argtypes = [vp]
    L.x264cu_frame_free.argtypes = [vp, vp]
    L.x264cu_frame_upload.argtypes = [vp, vp, vp, C.c_int, vp]
    L.x264cu_frame_upload_device.argtypes = [vp, vp, vp, C.c_int, vp]
    L.x264cu_frame_upload_aq.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int, C.POINTER(AqCfg), vp, C.POINTER(AqOut)]
    L.x264cu_frame_upload_aq_device.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int, C.POINTER(AqCfg), vp]
    L.x264cu_frame_set_pixel_stats.argtypes = [vp, vp, C.c_uint64, C.c_uint64]
    L.x264cu_weights_analyse.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.POINTER(Weight)]
    L.x264cu_frame_cost.argtypes = [vp, C.POINTER(Job)]
    L.x264cu_frame_cost_batch.argtypes = [vp, C.POINTER(Job), C.c_int]
    L.x264cu_motionsearch.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp]
    L.x264cu_flush.argtypes = [vp]
    L.x264cu_last_batch.argtypes = [vp]
    L.x264cu_last_batch.restype = C.c_uint64
    L.x264cu_wait.argtypes = [vp, C.c_uint64]
    L.x264cu_launch_count.argtypes = [vp]
    L.x264cu_launch_count.restype = C.c_int64
    L.x264cu_kernel_times.argtypes = [vp, vp, vp]
    L.x264cu_set_profiling.argtypes = [vp, C.c_int]
    L.x264cu_frame_get_lowres.argtypes = [vp, vp, C.c_int, vp]
    L.x264cu_pixel_cmp.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int64, C.c_int, vp, C.c_int64, C.c_int,
                                   C.c_int64, vp, vp, vp, C.c_int]
    L.x264cu_pixel_cmp_grid.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, vp, C.c_int]
    L.x264cu_sad_x.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int64, C.c_int, vp, vp]
    L.x264cu_hadamard_ac.argtypes = [vp, C.c_int, vp, C.c_int64, C.c_int, C.c_int64, vp, vp, C.c_int]
    L.x264cu_hadamard_ac_grid.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]
    L.x264cu_sub_dct.argtypes = [vp, C.c_int, vp, C.c_int64, C.c_int, vp, C.c_int64, C.c_int, C.c_int64,
                                 vp, vp, vp, C.c_int]
    L.x264cu_sub_dct_grid.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]
    L.x264cu_quant.argtypes = [vp, C.c_int, vp, C.c_int64, vp, vp, vp, C.c_int]
    L.x264cu_mc_luma.argtypes = [vp, vp, C.c_int64, vp, C.c_int, C.c_int64, vp, vp, C.c_int, C.c_int, vp, vp, C.c_int]
    L.x264cu_pixel_avg.argtypes = [vp, C.c_int, vp, C.c_int64, C.c_int, vp, C.c_int64, C.c_int, C.c_int64, vp, vp, C.c_int, vp, C.c_int]
    L.x264cu_mc_weight.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]
    L.x264cu_intra_x3_8x8c.argtypes = [vp, C.c_int, vp, C.c_int64, C.c_int, vp, C.c_int64, C.c_int, C.c_int64, vp, vp, vp, C.c_int]
    L.x264cu_quant_4x4x4.argtypes = [vp, vp, C.c_int64, vp, vp, vp, C.c_int]
    L.x264cu_lowres_core.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.x264cu_stream_open.restype = vp
    L.x264cu_stream_open.argtypes = [vp, C.c_int, C.c_int]
    L.x264cu_stream_close.argtypes = [vp]
    L.x264cu_stream_push.argtypes = [vp, vp, C.c_int, vp, C.c_int]
    L.x264cu_stream_push_shard.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.x264cu_stream_frame.restype = vp
    L.x264cu_stream_frame.argtypes = [vp, C.c_int64]
    L.x264cu_fence.argtypes = [vp]
    L.x264cu_fence.restype = C.c_uint64
    L.x264cu_peer_create.restype = vp
    L.x264cu_peer_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.POINTER(C.c_uint64)]
    L.x264cu_peer_connect.argtypes = [vp, vp, C.POINTER(C.c_uint64)]
    L.x264cu_peer_connect_local.argtypes = [vp, C.POINTER(vp), C.c_int]
    L.x264cu_peer_push.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
    L.x264cu_peer_import.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
    L.x264cu_peer_error.argtypes = [vp]
    L.x264cu_peer_destroy.argtypes = [vp]
    L.x264cu_search_batch.argtypes = [vp, C.POINTER(SearchJob), C.c_int]
    L.x264cu_field_device.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp)]
    L.x264cu_field_set_searched.argtypes = [vp, vp, C.c_int, C.c_int]
    L.x264cu_memcpy_d2d.argtypes = [vp, vp, vp, C.c_size_t]
    L.x264cu_stream_sync.argtypes = [vp]
    L.x264cu_stream_cost.argtypes = [vp, C.c_int, C.c_int, vp]
    L.x264cu_stream_arrays.argtypes = [vp, C.c_int, C.c_int, C.POINTER(Out)]
    L.x264cu_stream_fetch.argtypes = [vp, C.c_int64, C.c_int, C.c_int, C.POINTER(Out)]
    L.x264cu_stream_frames.argtypes = [vp]
    L.x264cu_stream_frames.restype = C.c_int64
    L.x264cu_stream_result_bytes.argtypes = [vp]
    L.x264cu_stream_result_bytes.restype = C.c_int64
    L.x264cu_malloc.restype = vp
    L.x264cu_malloc.argtypes = [vp, C.c_size_t]
    L.x264cu_free.argtypes = [vp, vp]
    L.x264cu_host_alloc.restype = vp
    L.x264cu_host_alloc.argtypes = [vp, C.c_size_t]
    L.x264cu_host_free.argtypes = [vp, vp]
    L.x264cu_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    L.x264cu_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    L.x264cu_timer_begin.argtypes = [vp]
    L.x264cu_timer_end.argtypes = [vp]
    L.x264cu_timer_end.restype = C.c_float
    _lib = L
    return L


def last_error():
    return load().x264cu_last_error().decode()


def check(r, what=""):
    if r != 0:
        raise RuntimeError("%s failed: %s" % (what, last_error()))


def default_mv_range(width, height):
    """x264's level-derived analyse.i_mv_range (encoder/set.c x264_levels + encoder.c validation) for
    progressive content: 512 from level 3.1 upward, 256 at 2.1-3.0, 128 at 1.1-2.0, else 64."""
    mbs = ((width + 15) // 16) * ((height + 15) // 16)
    if mbs > 1620:
        return 512
    if mbs > 396:
        return 256
    if mbs > 99:
        return 128
    return 64


class Context:
    """x264cu_ctx + convenience.  Mirrors the reference hook protocol: upload -> frame_cost* -> flush."""

    def __init__(self, width, height, bframes=3, subme=7, me_method=1, me_range=16, mv_range=None,
                 do_edges=1, aq=1, weighted_bipred=1, bframe_bias=0, passes=2, max_frames=64, device=0,
                 cost_mv=None):
        self.L = load()
        self.W, self.H = width, height
        if mv_range is None:
            mv_range = default_mv_range(width, height)
        self._cost_mv = None
        cm = None
        if cost_mv is not None:
            self._cost_mv = np.ascontiguousarray(cost_mv, dtype=np.uint16)
            cm = self._cost_mv.ctypes.data
        self.cfg = Cfg(width, height, bframes, subme, me_method, me_range, mv_range, do_edges, aq,
                       weighted_bipred, bframe_bias, passes, max_frames, cm)
        self.ctx = self.L.x264cu_open(device, C.byref(self.cfg))
        if not self.ctx:
            raise RuntimeError("x264cu_open failed: " + last_error())
        info = (C.c_int * 6)()
        check(self.L.x264cu_info(self.ctx, info), "x264cu_info")
        self.mbw, self.mbh, self.lw, self.lh, self.stride, self.sm_count = list(info)
        self.nmb = self.mbw * self.mbh
        self.bframes = bframes
        self._keep = []          # arrays referenced by queued jobs until the next flush

    def close(self):
        if self.ctx:
            self.L.x264cu_close(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- frames ----
    def frame_new(self):
        f = self.L.x264cu_frame_new(self.ctx)
        if not f:
            raise RuntimeError("x264cu_frame_new failed: " + last_error())
        return f

    def frame_free(self, f):
        self.L.x264cu_frame_free(self.ctx, f)

    def upload(self, f, luma, inv_qscale=None, stride=None):
        """luma: numpy uint8 (H, W) or an integer host address (e.g. pinned memory)"""
        if isinstance(luma, np.ndarray):
            assert luma.dtype == np.uint8 and luma.shape == (self.H, self.W) and luma.flags.c_contiguous
            self._keep.append(luma)
            lp, st = luma.ctypes.data, self.W
        else:
            lp, st = luma, stride or self.W
        qp = None
        if inv_qscale is not None:
            if isinstance(inv_qscale, np.ndarray):
                inv_qscale = np.ascontiguousarray(inv_qscale, dtype=np.uint16)
                self._keep.append(inv_qscale)
                qp = inv_qscale.ctypes.data
            else:
                qp = inv_qscale
        check(self.L.x264cu_frame_upload(self.ctx, f, lp, st, qp), "x264cu_frame_upload")

    def upload_aq(self, f, luma, chroma, aq_mode=1, aq_strength=1.0, quant_offsets=None):
        """x264cu_frame_upload_aq: luma (H, W), chroma interleaved NV12 (H/2, W).  Returns the arrays filled at flush:
        qp_offset (f32), inv_qscale (u16), pixel_stats (u64 x 6: sum/ssd of Y, U, V)."""
        assert luma.dtype == np.uint8 and luma.shape == (self.H, self.W) and luma.flags.c_contiguous
        assert chroma.dtype == np.uint8 and chroma.shape == (self.H // 2, self.W) and chroma.flags.c_contiguous
        res = {"qp_offset": np.zeros(self.nmb, np.float32), "inv_qscale": np.zeros(self.nmb, np.uint16),
               "pixel_stats": np.zeros(6, np.uint64)}
        out = AqOut(res["qp_offset"].ctypes.data, res["inv_qscale"].ctypes.data, res["pixel_stats"].ctypes.data)
        cfg = AqCfg(aq_mode, aq_strength)
        qo = None
        if quant_offsets is not None:
            qo = np.ascontiguousarray(quant_offsets, dtype=np.float32)
            self._keep.append(qo)
        self._keep.extend([luma, chroma, res, out, cfg])
        check(self.L.x264cu_frame_upload_aq(self.ctx, f, luma.ctypes.data, self.W, chroma.ctypes.data, self.W, C.byref(cfg),
                                            _ptr(qo), C.byref(out)), "x264cu_frame_upload_aq")
        return res

    def set_pixel_stats(self, f, pixel_sum, pixel_ssd):
        check(self.L.x264cu_frame_set_pixel_stats(self.ctx, f, int(pixel_sum), int(pixel_ssd)), "x264cu_frame_set_pixel_stats")

    def weights_analyse(self, fenc, fref, dist, numslices=1):
        """x264cu_weights_analyse; returns a Weight that is filled at flush (and stays on the device for
        make_job(..., weight='device'))"""
        w = Weight(0, 0, 0, 0)
        self._keep.append(w)
        check(self.L.x264cu_weights_analyse(self.ctx, fenc, fref, dist, numslices, C.byref(w)), "x264cu_weights_analyse")
        return w

    def upload_device(self, f, d_luma, stride, d_inv_qscale=None):
        check(self.L.x264cu_frame_upload_device(self.ctx, f, d_luma, stride, d_inv_qscale), "x264cu_frame_upload_device")

    def get_lowres(self, f, plane):
        out = np.zeros((self.lh + 64, self.lw + 64), dtype=np.uint8)
        check(self.L.x264cu_frame_get_lowres(self.ctx, f, plane, _ptr(out)), "x264cu_frame_get_lowres")
        return out

    # ---- evaluations ----
    def make_job(self, fenc, fref0, fref1, d0, d1, want=("stats",), weight=None):
        """returns (Job, dict of numpy arrays that will be filled at flush)."""
        res = {}
        o = Out()
        if "lowres_costs" in want:
            res["lowres_costs"] = np.zeros(self.nmb, dtype=np.uint16); o.lowres_costs = res["lowres_costs"].ctypes.data
        if "mvs" in want:
            for l, d in ((0, d0), (1, d1)):
                if d > 0:
                    res["mvs%d" % l] = np.zeros((self.nmb, 2), dtype=np.int16); o.mvs[l] = res["mvs%d" % l].ctypes.data
                    res["mv_costs%d" % l] = np.zeros(self.nmb, dtype=np.int32); o.mv_costs[l] = res["mv_costs%d" % l].ctypes.data
        if "row_satds" in want:
            res["row_satds"] = np.zeros(self.mbh, dtype=np.int32); o.row_satds = res["row_satds"].ctypes.data
        if "stats" in want:
            res["stats"] = np.zeros(4, dtype=np.int32); o.stats = res["stats"].ctypes.data
        if "intra" in want:
            res["intra_costs"] = np.zeros(self.nmb, dtype=np.uint16); o.intra_costs = res["intra_costs"].ctypes.data
            res["intra_row_satds"] = np.zeros(self.mbh, dtype=np.int32); o.intra_row_satds = res["intra_row_satds"].ctypes.data
            res["intra_stats"] = np.zeros(2, dtype=np.int32); o.intra_stats = res["intra_stats"].ctypes.data
        w = Weight(0, 0, 0, 0)
        if isinstance(weight, str) and weight == "device":
            w = Weight(WEIGHT_DEVICE, 0, 0, 0)
        elif weight is not None:
            w = Weight(1, weight[0], weight[1], weight[2])
        return Job(fenc, fref0, fref1, d0, d1, w, o), res

    def frame_cost_batch(self, jobs_and_results):
        n = len(jobs_and_results)
        arr = (Job * n)(*[j for j, _ in jobs_and_results])
        self._keep.append(arr)
        self._keep.extend(r for _, r in jobs_and_results)
        check(self.L.x264cu_frame_cost_batch(self.ctx, arr, n), "x264cu_frame_cost_batch")

    def frame_cost(self, fenc, fref0, fref1, d0, d1, want=("stats",), weight=None, flush=True):
        job, res = self.make_job(fenc, fref0, fref1, d0, d1, want, weight)
        self._keep.append((job, res))
        check(self.L.x264cu_frame_cost(self.ctx, C.byref(job)), "x264cu_frame_cost")
        if flush:
            self.flush()
        return res

    def flush(self):
        check(self.L.x264cu_flush(self.ctx), "x264cu_flush")
        self._keep = []

    def launches(self):
        return int(self.L.x264cu_launch_count(self.ctx))

    def set_profiling(self, on):
        check(self.L.x264cu_set_profiling(self.ctx, int(on)))

    def kernel_times(self):
        ms = (C.c_float * 4)()
        n = (C.c_int * 4)()
        check(self.L.x264cu_kernel_times(self.ctx, ms, n))
        names = ("lowres", "intra", "search", "final")
        return {k: (ms[i], n[i]) for i, k in enumerate(names)}

    # ---- device memory ----
    def malloc(self, nbytes):
        p = self.L.x264cu_malloc(self.ctx, nbytes)
        if not p:
            raise RuntimeError("x264cu_malloc failed: " + last_error())
        return p

    def free(self, p):
        self.L.x264cu_free(self.ctx, p)

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        p = self.malloc(arr.nbytes + 64)
        check(self.L.x264cu_memcpy_h2d(self.ctx, p, _ptr(arr), arr.nbytes), "h2d")
        return p

    def from_device(self, p, shape, dtype):
        out = np.zeros(shape, dtype=dtype)
        check(self.L.x264cu_memcpy_d2h(self.ctx, _ptr(out), p, out.nbytes), "d2h")
        return out

    def host_alloc(self, nbytes):
        p = self.L.x264cu_host_alloc(self.ctx, nbytes)
        if not p:
            raise RuntimeError("x264cu_host_alloc failed: " + last_error())
        return p

    def host_free(self, p):
        self.L.x264cu_host_free(self.ctx, p)

    def timer_begin(self):
        check(self.L.x264cu_timer_begin(self.ctx))

    def timer_end(self):
        return float(self.L.x264cu_timer_end(self.ctx))

    # ---- primitives (host-array convenience) ----
    def pixel_cmp(self, metric, size, p1, s1, p2, s2, off1, off2):
        off1 = np.ascontiguousarray(off1, dtype=np.int64); off2 = np.ascontiguousarray(off2, dtype=np.int64)
        out = np.zeros(len(off1), dtype=np.int32)
        check(self.L.x264cu_pixel_cmp(self.ctx, metric, size, _ptr(p1), p1.nbytes, s1, _ptr(p2), p2.nbytes, s2,
                                      len(off1), _ptr(off1), _ptr(off2), _ptr(out), 0), "x264cu_pixel_cmp")
        return out

    def pixel_cmp_grid(self, metric, size, img1, img2, dx=0, dy=0):
        h, w = img1.shape
        bw, bh = SIZE_WH[size]
        assert img2.shape[0] >= h + dy and img2.shape[1] >= w + dx
        out = np.zeros((h // bh, w // bw), dtype=np.int32)
        check(self.L.x264cu_pixel_cmp_grid(self.ctx, metric, size, _ptr(img1), img1.strides[0], _ptr(img2), img2.strides[0],
                                           w, h, dx, dy, _ptr(out), 0), "x264cu_pixel_cmp_grid")
        return out

    def sad_x(self, size, fenc16, refs, stride, offs):
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        out = np.zeros(len(offs), dtype=np.int32)
        check(self.L.x264cu_sad_x(self.ctx, size, len(offs), _ptr(fenc16), _ptr(refs), refs.nbytes, stride, _ptr(offs), _ptr(out)), "x264cu_sad_x")
        return out

    def hadamard_ac(self, size, p, stride, off):
        off = np.ascontiguousarray(off, dtype=np.int64)
        out = np.zeros(len(off), dtype=np.uint64)
        check(self.L.x264cu_hadamard_ac(self.ctx, size, _ptr(p), p.nbytes, stride, len(off), _ptr(off), _ptr(out), 0), "x264cu_hadamard_ac")
        return out

    def hadamard_ac_grid(self, size, img):
        h, w = img.shape
        bw, bh = SIZE_WH[size]
        out = np.zeros((h // bh, w // bw), dtype=np.uint64)
        check(self.L.x264cu_hadamard_ac_grid(self.ctx, size, _ptr(img), img.strides[0], w, h, _ptr(out), 0), "x264cu_hadamard_ac_grid")
        return out

    def sub_dct(self, bsize, p1, s1, p2, s2, off1, off2):
        off1 = np.ascontiguousarray(off1, dtype=np.int64); off2 = np.ascontiguousarray(off2, dtype=np.int64)
        out = np.zeros((len(off1), bsize * bsize), dtype=np.int16)
        check(self.L.x264cu_sub_dct(self.ctx, bsize, _ptr(p1), p1.nbytes, s1, _ptr(p2), p2.nbytes, s2, len(off1),
                                    _ptr(off1), _ptr(off2), _ptr(out), 0), "x264cu_sub_dct")
        return out

    def sub_dct_grid(self, bsize, img1, img2):
        h, w = img1.shape
        out = np.zeros(((h // bsize) * (w // bsize), bsize * bsize), dtype=np.int16)
        check(self.L.x264cu_sub_dct_grid(self.ctx, bsize, _ptr(img1), img1.strides[0], _ptr(img2), img2.strides[0], w, h, _ptr(out), 0), "x264cu_sub_dct_grid")
        return out

    def quant(self, b8x8, coefs, mf, bias):
        n = 64 if b8x8 else 16
        coefs = np.ascontiguousarray(coefs, dtype=np.int16).copy()
        nb = coefs.size // n
        mf = np.ascontiguousarray(mf, dtype=np.uint16); bias = np.ascontiguousarray(bias, dtype=np.uint16)
        nz = np.zeros(nb, dtype=np.int32)
        check(self.L.x264cu_quant(self.ctx, int(b8x8), _ptr(coefs), nb, _ptr(mf), _ptr(bias), _ptr(nz), 0), "x264cu_quant")
        return coefs, nz

    def lowres_core(self, src, sstride, width, height, dstride):
        planes = [np.zeros(dstride * height, dtype=np.uint8) for _ in range(4)]
        check(self.L.x264cu_lowres_core(self.ctx, _ptr(src), sstride, *[_ptr(p) for p in planes], dstride, width, height, 0), "x264cu_lowres_core")
        return [p.reshape(height, dstride)[:, :width].copy() for p in planes]


class CudaLookahead:
    """same driver interface as tests/checkers.Lookahead, on the CUDA product through the C ABI."""

    def __init__(self, width, height, bframes=3, subme=7, mbtree=1, aq=1, passes=2, mv_range=None, **kw):
        self.ctx = Context(width, height, bframes=bframes, subme=subme, do_edges=mbtree, aq=aq, passes=passes,
                           mv_range=mv_range, **kw)
        self.frames = {}
        self.results = {}
        self.mbw, self.mbh, self.nmb = self.ctx.mbw, self.ctx.mbh, self.ctx.nmb

    def close(self):
        for f in self.frames.values():
            self.ctx.frame_free(f)
        self.frames = {}
        self.ctx.close()

    def put_frame(self, idx, luma, inv_qscale=None):
        if idx not in self.frames:
            self.frames[idx] = self.ctx.frame_new()
        self.ctx.upload(self.frames[idx], np.ascontiguousarray(luma), inv_qscale)
        self.ctx.flush()
        for k in [k for k in self.results if k[0] == idx]:
            del self.results[k]

    def put_frame_yuv(self, idx, luma, chroma, aq_mode=1, aq_strength=1.0, quant_offsets=None):
        if idx not in self.frames:
            self.frames[idx] = self.ctx.frame_new()
        r = self.ctx.upload_aq(self.frames[idx], np.ascontiguousarray(luma), np.ascontiguousarray(chroma), aq_mode, aq_strength, quant_offsets)
        self.ctx.flush()
        self.aq = getattr(self, "aq", {})
        self.aq[idx] = r
        for k in [k for k in self.results if k[0] == idx]:
            del self.results[k]

    def get_aq(self, idx):
        r = self.aq[idx]
        return r["qp_offset"], r["inv_qscale"], r["pixel_stats"]

    def weights_analyse(self, b, ref, numslices=1, flush=True):
        w = self.ctx.weights_analyse(self.frames[b], self.frames[ref], b - ref, numslices)
        if not flush:
            return w
        self.ctx.flush()
        return (w.enabled, w.scale, w.denom, w.offset) if w.enabled else (0, 1, 0, 0)

    def lowres(self, idx, plane):
        return self.ctx.get_lowres(self.frames[idx], plane)

    ALL = ("lowres_costs", "mvs", "row_satds", "stats", "intra")

    def frame_cost(self, p0, p1, b, weight=None):
        """weight: None, (scale, denom, offset), or 'device' (the result of the last weights_analyse(b, p0))"""
        r = self.ctx.frame_cost(self.frames[b], self.frames[p0], self.frames[p1], b - p0, p1 - b, want=self.ALL, weight=weight)
        self.results[(b, b - p0, p1 - b)] = r
        return int(r["stats"][0])

    def frame_cost_many(self, triples):
        jr = [self.ctx.make_job(self.frames[b], self.frames[p0], self.frames[p1], b - p0, p1 - b, want=self.ALL) for (p0, p1, b) in triples]
        self.ctx.frame_cost_batch(jr)
        self.ctx.flush()
        out = []
        for (p0, p1, b), (_, r) in zip(triples, jr):
            self.results[(b, b - p0, p1 - b)] = r
            out.append(int(r["stats"][0]))
        return out
