"""CPU checks of the packed-integer identities the CUDA kernels rely on (x264-i386pic_b200/csrc/common.cuh, la_block.cuh).

The device code replaces a few table lookups and multi-instruction sequences by immediates and packed forms; each of them is
an exact identity, restated here in plain Python against the definition it replaces.  The constants are parsed out of the
sources, so the test follows the code.  (The GPU parity tests cover the same code end to end; this file pins the arithmetic
without a GPU.)"""
import os
import random
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "x264-i386pic_b200", "csrc", "la_block.cuh")).read()
M32 = 0xFFFFFFFF


def prmt(a, b, sel):
    """PTX prmt.b32 in its default mode: result byte i = byte (sel >> 4i) & 7 of the 8-byte pool {a, b}"""
    pool = [(a >> (8 * i)) & 255 for i in range(4)] + [(b >> (8 * i)) & 255 for i in range(4)]
    return sum(pool[(sel >> (4 * i)) & 7] << (8 * i) for i in range(4))


def imm(name):
    m = re.search(r"int %s\( int \w+ \) \{ return \(int\)\( \( (0x[0-9a-fA-F]+)u >>" % name, SRC)
    assert m, name
    return int(m.group(1), 16)


def test_hpel_plane_tables_as_immediates():
    # x264_hpel_ref0 / x264_hpel_ref1 (reference common/tables.c:183-184)
    ref0 = [0, 1, 1, 1, 0, 1, 1, 1, 2, 3, 3, 3, 0, 1, 1, 1]
    ref1 = [0, 0, 1, 0, 2, 2, 3, 2, 2, 2, 3, 2, 2, 2, 3, 2]
    c0, c1 = imm("hpel_ref0"), imm("hpel_ref1")
    assert [(c0 >> (2 * i)) & 3 for i in range(16)] == ref0
    assert [(c1 >> (2 * i)) & 3 for i in range(16)] == ref1


def test_square_and_hexagon_offsets_as_immediates():
    # square1 (reference encoder/me.c:56) and hex2 (me.c:55)
    sq = [(0, 0), (0, -1), (0, 1), (-1, 0), (1, 0), (-1, -1), (-1, 1), (1, -1), (1, 1)]
    cx = int(re.search(r"int sq1x\( int i \) \{ return \(int\)\( \( (0x[0-9a-f]+)u", SRC).group(1), 16)
    cy = int(re.search(r"int sq1y\( int i \) \{ return \(int\)\( \( (0x[0-9a-f]+)u", SRC).group(1), 16)
    assert [(((cx >> (2 * i)) & 3) - 1, ((cy >> (2 * i)) & 3) - 1) for i in range(9)] == sq
    hexp = [(-1, -2), (-2, 0), (-1, 2), (1, 2), (2, 0), (1, -2), (-1, -2), (-2, 0)]
    hx = int(re.search(r"int hex2x\( int i \) \{ return \(int\)\( \( (0x[0-9a-f]+)u", SRC).group(1), 16)
    hy = int(re.search(r"int hex2y\( int i \) \{ return \(int\)\( \( (0x[0-9a-f]+)u", SRC).group(1), 16)
    assert [(((hx >> (4 * i)) & 7) - 2, ((hy >> (4 * i)) & 7) - 2) for i in range(8)] == hexp


def test_weighted_biprediction_average_scaled_by_four():
    """biavg4: (a*w + b*(64-w) + 32) >> 6 on 4 packed pixels (reference common/mc.c:63-99 without the clip, which a convex
    combination never needs) as two 16:16 multiply-adds scaled by 4 and one byte gather"""
    assert "0x00800080u" in SRC and "0x7351" in SRC
    rng = random.Random(7)
    edge = [(0, 0), (M32, M32), (M32, 0), (0, M32), (0x00FF00FF, 0xFF00FF00)]
    for w in range(1, 64):
        for n in range(120):
            a, b = edge[n] if n < len(edge) else (rng.getrandbits(32), rng.getrandbits(32))
            w4, v4 = 4 * w, 256 - 4 * w
            ae, ao, be, bo = prmt(a, 0, 0x4240), prmt(a, 0, 0x4341), prmt(b, 0, 0x4240), prmt(b, 0, 0x4341)
            re_ = (ae * w4 + be * v4 + 0x00800080) & M32
            ro = (ao * w4 + bo * v4 + 0x00800080) & M32
            # no carry may cross from the low 16-bit field into the high one
            assert (ae & 0xFFFF) * w4 + (be & 0xFFFF) * v4 + 0x80 < 1 << 16
            got = prmt(re_, ro, 0x7351)
            want = sum((((((a >> (8 * i)) & 255) * w + ((b >> (8 * i)) & 255) * (64 - w) + 32) >> 6) & 255) << (8 * i) for i in range(4))
            assert got == want, (w, hex(a), hex(b))


def test_rounding_byte_average():
    """avg4: (a + b + 1) >> 1 per byte == (a | b) - (((a ^ b) & 0xfefefefe) >> 1); in particular avg4(x, x) == x, which the
    even-centre SATD stage uses for the centre itself"""
    rng = random.Random(3)
    for _ in range(4000):
        a, b = rng.getrandbits(32), rng.getrandbits(32)
        got = ((a | b) - (((a ^ b) & 0xFEFEFEFE) >> 1)) & M32
        want = sum(((((a >> (8 * i)) & 255) + ((b >> (8 * i)) & 255) + 1) >> 1) << (8 * i) for i in range(4))
        assert got == want
        assert ((a | a) - (((a ^ a) & 0xFEFEFEFE) >> 1)) & M32 == a


def s16(v):
    return v - 65536 if v & 0x8000 else v


def pack(lo, hi):
    """x264's "true sum" packing of two signed 16-bit values (common/pixel.c:242-259): lo + (hi << 16) as an integer"""
    return (lo + (hi << 16)) & M32


def abs2(a):
    """common.cuh abs2(): PRMT in sign-replicate mode builds 0xffff per 16-bit field whose bit 15 is set"""
    s = (0xFFFF if a & 0x8000 else 0) | (0xFFFF0000 if a & 0x80000000 else 0)
    return ((a + s) & M32) ^ s


def test_abs2_and_the_folded_last_butterfly():
    """the last Hadamard stage is folded into the absolute sum: |u + v| + |u - v| == 2 * max(|u|, |v|) per field, on the
    reference's packed words (fields stay far below 2^15 for 8-bit pixels)"""
    rng = random.Random(11)
    for _ in range(4000):
        u = [rng.randint(-4000, 4000) for _ in range(2)]
        v = [rng.randint(-4000, 4000) for _ in range(2)]
        pu, pv = pack(*u), pack(*v)
        au, av = abs2(pu), abs2(pv)
        assert (au & 0xFFFF, au >> 16) == (abs(u[0]), abs(u[1]))
        s, d = (pu + pv) & M32, (pu - pv) & M32
        want = abs2(s) + abs2(d)                       # what pixel.c:300-306 adds up
        got = 2 * (max(au & 0xFFFF, av & 0xFFFF) + (max(au >> 16, av >> 16) << 16))
        assert got == want
