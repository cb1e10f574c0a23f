"""CPU: the oracle's restatement of the fetch / prediction functions behind the checkasm mc / intra / quant tests against
the UNMODIFIED reference tables (oracle/_ref): mc_luma over every quarter-pel phase and size with weights incl. denom 0,
avg over 12 sizes x weights -63..127, mc_weight, intra_x3_8x8c, quant_4x4x4 and quant under six CQM configurations."""
import ctypes as C

import numpy as np
import pytest

import mc_cases as M
from checkers import OracleLib, RefLib, have_ref

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libx264ref.so not built")


def test_mc_luma_all_phases_sizes_weights():
    O, R = OracleLib(), RefLib()
    buf, po, stride, psz = M.planes_buffer()
    rng = np.random.default_rng(3)
    for (w, h) in M.MC_SIZES:
        org, mv = M.mc_cases(stride, psz // stride, w, h, rng)
        assert len({(int(x) & 3, int(y) & 3) for x, y in mv}) == 16
        for wt in M.WEIGHTS:
            np.testing.assert_array_equal(M.oracle_mc_luma(O, buf, po, stride, org, mv, w, h, wt),
                                          M.ref_mc_luma(R, buf, po, stride, org, mv, w, h, wt), err_msg=str((w, h, wt)))


def test_avg_all_sizes_and_weights():
    O, R = OracleLib(), RefLib()
    rng = np.random.default_rng(4)
    a = rng.integers(0, 256, 64 * 40, dtype=np.uint8)
    b = rng.integers(0, 256, 48 * 40, dtype=np.uint8)
    a[:300] = 255; b[:300] = 0
    off1 = np.array([0, 5, 64 * 3 + 17, 64 * 20 + 1], dtype=np.int64)
    off2 = np.array([1, 48 * 2, 48 * 7 + 9, 48 * 19 + 30], dtype=np.int64)
    for size in range(12):
        for weight in list(range(-63, 128, 5)) + [32, 0, 64, 127, -63]:
            np.testing.assert_array_equal(M.oracle_avg(O, a, 64, b, 48, off1, off2, size, weight),
                                          M.ref_avg(R, a, 64, b, 48, off1, off2, size, weight), err_msg=str((size, weight)))


def test_mc_weight_incl_denom_zero():
    O, R = OracleLib(), RefLib()
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, 32 * 16 + 32, dtype=np.uint8)
    src[:40] = 255; src[40:80] = 0
    for wt in [(s, d, o) for s in (0, 1, 31, 64, 127) for d in range(8) for o in (-128, -5, 0, 77, 127)]:
        for width in (4, 8, 12, 16, 20):
            np.testing.assert_array_equal(M.oracle_weight(O, src, 32, width, 16, wt), M.ref_weight(R, src, 32, width, 16, wt), err_msg=str((wt, width)))


@pytest.mark.parametrize("b_satd", [0, 1])
def test_intra_x3_8x8c(b_satd):
    O, R = OracleLib(), RefLib()
    rng = np.random.default_rng(6)
    s1, s2 = 80, 96
    p1 = rng.integers(0, 256, s1 * 48, dtype=np.uint8)
    p2 = rng.integers(0, 256, s2 * 48, dtype=np.uint8)
    p2[s2 * 10:s2 * 14] = 255; p1[s1 * 10:s1 * 20] = 0
    off1 = np.array([s1 * y + x for y in (1, 9, 17, 30) for x in (1, 8, 33, 70)], dtype=np.int64)
    off2 = np.array([s2 * y + x for y in (1, 10, 13, 31) for x in (1, 9, 40, 85)], dtype=np.int64)
    np.testing.assert_array_equal(M.oracle_intra_x3(O, b_satd, p1, s1, p2, s2, off1, off2), M.ref_intra_x3(R, b_satd, p1, s1, p2, s2, off1, off2))


def test_quant_4x4x4_and_cqm_configurations():
    O, R = OracleLib(), RefLib()
    L = R.lib
    L.ref_cqm_select.argtypes = [C.c_int, C.c_uint]
    L.ref_quant_4x4x4_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(7)
    try:
        for i_cqm in range(6):
            assert L.ref_cqm_select(i_cqm, 1234) == 0
            for qp in range(52):
                mf, bias = R.cqm_tables(0, 0, qp)
                for j in (0xF, 0x5, 0x0, 0x8):
                    c = M.quant_coefs(rng, 6, j)
                    ref = c.copy()
                    nza = np.zeros(6, dtype=np.int32)
                    assert L.ref_quant_4x4x4_batch(ref.ctypes.data, 6, mf.ctypes.data, bias.ctypes.data, nza.ctypes.data) == 0
                    q, mask = M.quant_4x4x4_expect(O.quant, c, mf, bias)
                    np.testing.assert_array_equal(q, ref, err_msg=str((i_cqm, qp, j)))
                    np.testing.assert_array_equal(mask, nza, err_msg=str((i_cqm, qp, j)))
                mf8, bias8 = R.cqm_tables(1, 0, qp)          # quant_8x8 under the same configuration
                c8 = (rng.integers(-4000, 4001, (5, 64))).astype(np.int16)
                qo, nzo = O.quant(1, c8, mf8, bias8)
                qr, nzr = R.quant(1, c8, mf8, bias8)
                np.testing.assert_array_equal(qo, qr); np.testing.assert_array_equal(nzo, nzr)
    finally:
        L.ref_cqm_select(-1, 0)
