"""SURVEY.md section 8f rows 2-3 on the CPU: the oracle's restatement of x264_adaptive_quant_frame (aq-mode 1,
ratecontrol.c:226-415) and of the lookahead flavour of x264_weights_analyse (slicetype.c:62-501) against the
UNMODIFIED reference functions (oracle/_ref), bit for bit -- floats included."""
import ctypes as C
import numpy as np
import pytest

from checkers import Lookahead, OracleLib, RefLib, have_ref, synth_chroma, synth_frames

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")


def _pair(W, H, bframes=3, subme=7, **kw):
    ref = Lookahead(RefLib(), "ref", W, H, bframes=bframes, subme=subme, weightp=2, **kw)
    orc = Lookahead(OracleLib(), "oracle", W, H, bframes=bframes, subme=subme, mv_range=ref.mv_range, **kw)
    return ref, orc


def test_luts_match_reference_tables():
    ref, orc = RefLib(), OracleLib()
    a, b = np.zeros(64, np.uint8), np.zeros(64, np.uint8)
    ref.lib.ref_exp2_lut(a.ctypes.data_as(C.c_void_p))
    orc.lib.xo_exp2_lut(b.ctypes.data_as(C.c_void_p))
    assert (a == b).all()


@pytest.mark.parametrize("W,H", [(320, 192), (200, 120), (1920, 1080)])
def test_adaptive_quant_matches_reference(W, H):
    n = 2
    luma = synth_frames(W, H, n, seed=3)
    chroma = synth_chroma(W, H, n)
    luma[1][: H // 3] = 17            # flat area: energy 0 -> log2(1) path; and a saturated one
    luma[1][H // 3: H // 2, : W // 2] = 255
    ref, orc = _pair(W, H)
    qo = np.linspace(-2.5, 2.5, ref.nmb).astype(np.float32)
    for i in range(n):
        offs = qo if i else None
        ref.put_frame_yuv(i, luma[i], chroma[i], quant_offsets=offs)
        orc.put_frame_yuv(i, luma[i], chroma[i], quant_offsets=offs)
        (qr, ir, sr), (qo_, io, so) = ref.get_aq(i), orc.get_aq(i)
        assert (qr.view(np.uint32) == qo_.view(np.uint32)).all()      # f_qp_offset, bit for bit
        assert (ir == io).all()                                        # i_inv_qscale_factor
        assert (sr == so).all()                                        # i_pixel_sum / i_pixel_ssd of Y, U, V
    ref.close(); orc.close()


@pytest.mark.parametrize("mode", [2, 3])
@pytest.mark.parametrize("W,H", [(320, 192), (200, 120), (1920, 1080)])
def test_auto_variance_modes_match_reference(W, H, mode):
    """aq-mode 2 / 3 (ratecontrol.c:355-392): powf(energy + 1, 0.125), the frame averages accumulated in raster order, then
    the offsets -- floats bit for bit"""
    n = 2
    luma = synth_frames(W, H, n, seed=13)
    chroma = synth_chroma(W, H, n)
    luma[1][: H // 3] = 17
    luma[1][H // 3: H // 2, : W // 2] = 255
    ref = Lookahead(RefLib(), "ref", W, H, bframes=3, weightp=2, aq=mode)
    orc = Lookahead(OracleLib(), "oracle", W, H, bframes=3, mv_range=ref.mv_range, aq=mode)
    qo = np.linspace(-2.5, 2.5, ref.nmb).astype(np.float32)
    for i in range(n):
        offs = qo if i else None
        ref.put_frame_yuv(i, luma[i], chroma[i], quant_offsets=offs)
        orc.put_frame_yuv(i, luma[i], chroma[i], aq_mode=mode, quant_offsets=offs)
        (qr, ir, sr), (qo_, io, so) = ref.get_aq(i), orc.get_aq(i)
        assert (qr.view(np.uint32) == qo_.view(np.uint32)).all()
        assert (ir == io).all() and (sr == so).all()
    assert len(np.unique(ir)) > 8
    ref.close(); orc.close()


def _fade(frames, k):
    out = []
    for i, f in enumerate(frames):
        g = f.astype(np.float32) * (1.0 - k * i) + 6 * i
        out.append(np.clip(g, 0, 255).astype(np.uint8))
    return out


@pytest.mark.parametrize("subme", [7, 1])
def test_weights_analyse_matches_reference(subme):
    W, H, n = 320, 192, 5
    seqs = {
        "plain": synth_frames(W, H, n, seed=5),                             # early termination (no weight)
        "fade": _fade(synth_frames(W, H, n, seed=6, pan=(0, 0), noise=1), 0.09),
        "fade_pan": _fade(synth_frames(W, H, n, seed=7), 0.06),
        "brighten": _fade(synth_frames(W, H, n, seed=8, pan=(1, 0)), -0.05),
    }
    chroma = synth_chroma(W, H, n)
    got_weighted = 0
    for name, frames in seqs.items():
        ref, orc = _pair(W, H, subme=subme)
        for i in range(n):
            ref.put_frame_yuv(i, frames[i], chroma[i])
            orc.put_frame_yuv(i, frames[i], chroma[i])
        for b, r in [(1, 0), (2, 0), (3, 1), (4, 1), (4, 3)]:
            wr, wo = ref.weights_analyse(b, r), orc.weights_analyse(b, r)
            assert wr == wo, (name, b, r, wr, wo)
            got_weighted += wr[0]
        ref.close(); orc.close()
    assert got_weighted >= 4          # the fades must actually exercise the weighted branch


def test_weighted_search_and_mc_reanalysis_match_reference():
    """The reference analyses the weight inside slicetype_frame_cost (slicetype.c:859-864), searches list 0 on the
    weighted plane, and a later x264_weights_analyse finds the MV field and motion-compensates the reference
    (weight_cost_init_luma, slicetype.c:76-103).  Same sequence through the oracle, serial schedule."""
    W, H, n = 320, 192, 3
    frames = _fade(synth_frames(W, H, n, seed=9, pan=(1, 0), noise=1), 0.15)
    chroma = synth_chroma(W, H, n)
    ref = Lookahead(RefLib(), "ref_serial", W, H, bframes=3, weightp=2)
    orc = Lookahead(OracleLib(), "oracle", W, H, bframes=3, mv_range=ref.mv_range, passes=0)
    for i in range(n):
        ref.put_frame_yuv(i, frames[i], chroma[i])
        orc.put_frame_yuv(i, frames[i], chroma[i])
    w = orc.weights_analyse(2, 0)
    assert w[0] == 1
    assert ref.frame_cost(0, 2, 2) == orc.frame_cost(0, 2, 2, weight=w)
    from checkers import GET_MVS, GET_LOWRES_COSTS
    assert (ref.get(2, GET_MVS, 0, 1) == orc.get(2, GET_MVS, 0, 1)).all()
    assert (ref.get(2, GET_LOWRES_COSTS, 2, 0) == orc.get(2, GET_LOWRES_COSTS, 2, 0)).all()
    assert ref.weights_analyse(2, 0) == orc.weights_analyse(2, 0)          # second analysis: MC path on both sides
    ref.close(); orc.close()
