"""GPU parity of the batched pixel / dct / quant primitives through the C ABI:
  * the checkasm cases (tools/checkasm.c:368-423,551-574,908-976,2078-2227) against the golden
    vectors generated from the reference build, overflow vectors included;
  * large random and extreme planes against the CPU oracle, for the offsets and the grid shapes;
  * size-independent properties at full 1080p size (symmetry, zero on identical inputs,
    additivity of SAD over sub-blocks, DC checks).
Everything is bit-exact."""
import os
import sys

import numpy as np
import pytest

from checkasm_data import PBUF2_OFF, SIZE_WH, pixel_cases
from checkers import OracleLib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "x264-i386pic_b200"))
GOLD = os.path.join(ROOT, "tests", "golden")
pytestmark = pytest.mark.gpu
KIND = {"sad": 0, "satd": 1, "sa8d": 2}


@pytest.fixture(scope="module")
def cu():
    import x264cu
    c = x264cu.Context(1920, 1080)
    yield c
    c.close()


@pytest.fixture(scope="module")
def prim():
    return np.load(os.path.join(GOLD, "primitives.npz"))


@pytest.fixture(scope="module")
def orc():
    return OracleLib()


@pytest.mark.parametrize("kind", ["sad", "satd", "sa8d"])
def test_checkasm_pixel_cases(cu, prim, kind):
    for size in range(8):
        if kind == "sa8d" and size not in (0, 3):
            continue
        res = []
        for (which, s1, s2, o1, o2) in pixel_cases(size):
            a, b = (prim["pbuf1"], prim["pbuf1"]) if which == "rand" else (prim["pbuf3"], prim["pbuf4"])
            res.append(cu.pixel_cmp(KIND[kind], size, a, s1, b, s2, o1, o2))
        np.testing.assert_array_equal(np.concatenate(res), prim[f"{kind}_{size}"], err_msg=f"{kind} {size}")


def test_checkasm_hadamard_ac(cu, prim):
    offs = np.arange(32, dtype=np.int64)
    for size in range(4):
        got = np.concatenate([cu.hadamard_ac(size, prim["pbuf1"], 16, offs * 30),
                              cu.hadamard_ac(size, prim["pbuf3"], 16, offs * 100),
                              cu.hadamard_ac(size, prim["pbuf4"], 32, offs * 64)])
        np.testing.assert_array_equal(got, prim[f"hadamard_ac_{size}"])


def test_checkasm_sad_x4(cu, prim):
    p = prim["pbuf1"]
    for size in range(7):
        fenc = np.ascontiguousarray(p[:16 * 16])
        refs = np.ascontiguousarray(p[PBUF2_OFF:])
        got = cu.sad_x(size, fenc, refs, 64, [0, 6, 1, 10])
        np.testing.assert_array_equal(got, prim[f"sad_x4_{size}"])
        np.testing.assert_array_equal(cu.sad_x(size, fenc, refs, 64, [0, 6, 1]), prim[f"sad_x4_{size}"][:3])


def test_checkasm_dct_quant_lowres(cu, prim):
    a, b, offs = prim["dct_a"].reshape(-1), prim["dct_b"].reshape(-1), prim["dct_offs"]
    for bs in (4, 8, 16):
        np.testing.assert_array_equal(cu.sub_dct(bs, a, 64, b, 64, offs, offs), prim[f"dct_{bs}"])
    for b8 in (0, 1):
        for cat in (0, 1):
            tag = f"quant{'8' if b8 else '4'}_{cat}"
            for k in range(len(prim["quant_qps"])):
                q, nz = cu.quant(b8, prim[tag + "_in"][k], prim[tag + "_mf"][k], prim[tag + "_bias"][k])
                np.testing.assert_array_equal(q.reshape(prim[tag + "_out"][k].shape), prim[tag + "_out"][k])
                np.testing.assert_array_equal(nz, prim[tag + "_nz"][k])
    src = prim["lowres_src"].reshape(-1)
    for w in (96, 104, 112, 120):
        np.testing.assert_array_equal(np.stack(cu.lowres_core(src, 256, w, 8, 128)), prim[f"lowres_{w}"])


def _planes(seed, h, w, extreme=False):
    rng = np.random.default_rng(seed)
    if extreme:
        a = np.where(rng.integers(0, 2, (h, w)) == 1, 255, 0).astype(np.uint8)
        return a, (255 - a).astype(np.uint8)
    return rng.integers(0, 256, (h, w), dtype=np.uint8), rng.integers(0, 256, (h, w), dtype=np.uint8)


@pytest.mark.parametrize("extreme", [False, True])
def test_offsets_vs_oracle(cu, orc, extreme):
    a, b = _planes(11, 272, 480, extreme)
    h, w = a.shape
    rng = np.random.default_rng(12)
    for kind in ("sad", "satd", "sa8d"):
        for size, (bw, bh) in enumerate(SIZE_WH):
            if kind == "sa8d" and size not in (0, 3):
                continue
            n = 2000
            o1 = rng.integers(0, h - bh, n) * w + rng.integers(0, w - bw, n)
            o2 = rng.integers(0, h - bh, n) * w + rng.integers(0, w - bw, n)
            np.testing.assert_array_equal(cu.pixel_cmp(KIND[kind], size, a.reshape(-1), w, b.reshape(-1), w, o1, o2),
                                          orc.pixel_cmp(kind, size, a.reshape(-1), w, b.reshape(-1), w, o1, o2),
                                          err_msg=f"{kind} {size}")
    for size in range(4):
        bw, bh = SIZE_WH[size]
        o = rng.integers(0, h - bh, 1000) * w + rng.integers(0, w - bw, 1000)
        np.testing.assert_array_equal(cu.hadamard_ac(size, a.reshape(-1), w, o), orc.hadamard_ac(size, a.reshape(-1), w, o))
    for bs in (4, 8, 16):
        o1 = rng.integers(0, h - bs, 500) * w + rng.integers(0, w - bs, 500)
        o2 = rng.integers(0, h - bs, 500) * w + rng.integers(0, w - bs, 500)
        np.testing.assert_array_equal(cu.sub_dct(bs, a.reshape(-1), w, b.reshape(-1), w, o1, o2),
                                      orc.sub_dct(bs, a.reshape(-1), w, b.reshape(-1), w, o1, o2))


def _grid_offsets(h, w, bw, bh, pitch, dx=0, dy=0):
    ys, xs = np.mgrid[0:h // bh, 0:w // bw]
    return ((ys * bh + dy) * pitch + xs * bw + dx).reshape(-1).astype(np.int64)


@pytest.mark.parametrize("dx,dy", [(0, 0), (1, 0), (5, 3), (8, 1), (63, 2)])
def test_grid_vs_oracle(cu, orc, dx, dy):
    """every non-overlapping block of plane 1 vs plane 2 displaced by the 64 sub-offsets idea of
    checkasm (pbuf2+j), regular-grid kernels."""
    h, w = 144, 256
    a, _ = _planes(21, h, w)
    _, b = _planes(22, h + 8, w + 64)
    for kind in ("sad", "satd", "sa8d"):
        for size, (bw, bh) in enumerate(SIZE_WH):
            if kind == "sa8d" and size not in (0, 3):
                continue
            got = cu.pixel_cmp_grid(KIND[kind], size, a, b, dx, dy)
            exp = orc.pixel_cmp(kind, size, a.reshape(-1), w, b.reshape(-1), w + 64, _grid_offsets(h, w, bw, bh, w),
                                _grid_offsets(h, w, bw, bh, w + 64, dx, dy))
            np.testing.assert_array_equal(got.reshape(-1), exp, err_msg=f"{kind} {size} dx={dx} dy={dy}")


def test_grid_hadamard_dct_vs_oracle(cu, orc):
    h, w = 144, 256
    a, b = _planes(31, h, w)
    for size in range(4):
        bw, bh = SIZE_WH[size]
        np.testing.assert_array_equal(cu.hadamard_ac_grid(size, a).reshape(-1),
                                      orc.hadamard_ac(size, a.reshape(-1), w, _grid_offsets(h, w, bw, bh, w)))
    for bs in (4, 8, 16):
        o = _grid_offsets(h, w, bs, bs, w)
        np.testing.assert_array_equal(cu.sub_dct_grid(bs, a, b), orc.sub_dct(bs, a.reshape(-1), w, b.reshape(-1), w, o, o))


def test_full_size_properties(cu):
    """1920x1080 planes (config 2 size): properties that need no CPU reference."""
    a, b = _planes(41, 1080, 1920)
    for size in (0, 3, 6):
        for kind in (0, 1):
            ab = cu.pixel_cmp_grid(kind, size, a, b)
            ba = cu.pixel_cmp_grid(kind, size, b, a)
            np.testing.assert_array_equal(ab, ba)                       # symmetry
            assert (cu.pixel_cmp_grid(kind, size, a, a) == 0).all()     # identity
    s16 = cu.pixel_cmp_grid(0, 0, a[:1072], b[:1072])
    s8 = cu.pixel_cmp_grid(0, 3, a[:1072], b[:1072])
    s4 = cu.pixel_cmp_grid(0, 6, a[:1072], b[:1072])
    np.testing.assert_array_equal(s16, s8.reshape(67, 2, 120, 2).sum(axis=(1, 3)))      # SAD is additive
    np.testing.assert_array_equal(s8, s4.reshape(134, 2, 240, 2).sum(axis=(1, 3)))
    exp = np.abs(a[:1072].astype(np.int64) - b[:1072].astype(np.int64)).reshape(67, 16, 120, 16).sum(axis=(1, 3))
    np.testing.assert_array_equal(s16, exp)
    # sub16x16_dct: coefficient 0 of each 4x4 is the sum of its residual
    d = cu.sub_dct_grid(4, a[:1072], b[:1072]).reshape(268, 480, 16)
    dc = (a[:1072].astype(np.int64) - b[:1072].astype(np.int64)).reshape(268, 4, 480, 4).sum(axis=(1, 3))
    np.testing.assert_array_equal(d[:, :, 0], dc)
    # a flat plane has no AC energy
    flat = np.full((64, 256), 200, dtype=np.uint8)
    assert (cu.hadamard_ac_grid(0, flat) == 0).all()


@pytest.mark.parametrize("sstride,dstride", [(208, 104), (232, 116), (211, 101), (216, 102)])   # 16-byte, 8-byte, unaligned paths
@pytest.mark.parametrize("width,height", [(101, 9), (100, 4), (7, 3), (4, 1), (3, 5)])
def test_lowres_core_ragged_and_unaligned(cu, orc, sstride, dstride, width, height):
    """frame_init_lowres_core (mc.c:484-507) at widths that are not a multiple of the 4-pixel vector, row counts that are
    not a multiple of the rows per thread, and strides that keep the kernel off its aligned path"""
    rng = np.random.default_rng(width * 131 + height * 7 + sstride)
    src = rng.integers(0, 256, sstride * (2 * height + 2), dtype=np.uint8)
    want = orc.lowres_core(src, sstride, width, height, dstride)
    got = cu.lowres_core(src, sstride, width, height, dstride)
    for k in range(4):
        np.testing.assert_array_equal(got[k], want[k], err_msg="plane %d" % k)
