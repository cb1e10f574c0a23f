"""SURVEY.md section 8f rows 2-3 on the device: x264cu_frame_upload_aq (x264_adaptive_quant_frame, aq-mode 0/1) and
x264cu_weights_analyse (the lookahead x264_weights_analyse) through the C ABI against the oracle, bit for bit,
including the float outputs; the oracle itself is pinned against the unmodified reference in tests/test_aq_weights.py."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "x264-i386pic_b200"))
from checkers import GET_LOWRES_COSTS, GET_MVS, GET_INTRA, Lookahead, OracleLib, synth_chroma, synth_frames

pytestmark = pytest.mark.gpu


def _fade(frames, k):
    return [np.clip(f.astype(np.float32) * (1.0 - k * i) + 6 * i, 0, 255).astype(np.uint8) for i, f in enumerate(frames)]


def _pair(W, H, **kw):
    import x264cu
    mvr = x264cu.default_mv_range(W, H)
    gpu = x264cu.CudaLookahead(W, H, mv_range=mvr, **kw)
    orc = Lookahead(OracleLib(), "oracle", W, H, mv_range=mvr, **kw)
    return gpu, orc


# 5120x2880: luma + chroma of one pageable upload exceed a quarter of the 64 MiB staging buffer the context had in round 1
# (the second staging allocation of a batch recycled the region under the first)
@pytest.mark.parametrize("W,H", [(320, 192), (200, 120), (1920, 1080), (3840, 2160), (5120, 2880)])
def test_device_aq_matches_oracle(W, H):
    n = 2
    luma = synth_frames(W, H, n, seed=3)
    chroma = synth_chroma(W, H, n)
    luma[1][: H // 3] = 17
    luma[1][H // 3: H // 2, : W // 2] = 255
    gpu, orc = _pair(W, H, bframes=3)
    qo = np.linspace(-2.5, 2.5, gpu.nmb).astype(np.float32)
    for i in range(n):
        offs = qo if i else None
        gpu.put_frame_yuv(i, luma[i], chroma[i], quant_offsets=offs)
        orc.put_frame_yuv(i, luma[i], chroma[i], quant_offsets=offs)
        (qg, ig, sg), (qo_, io, so) = gpu.get_aq(i), orc.get_aq(i)
        assert (qg.view(np.uint32) == qo_.view(np.uint32)).all()
        assert (ig == io).all()
        assert (sg == so).all()
    for pl in range(4):
        assert (gpu.lowres(1, pl) == orc.lowres(1, pl)).all(), pl
    # the AQ output feeds the path: intra costs / frame costs with the device-computed inv_qscale
    assert gpu.frame_cost(0, 1, 1) == orc.frame_cost(0, 1, 1)
    assert (gpu.results[(1, 1, 0)]["lowres_costs"] == orc.get(1, GET_LOWRES_COSTS, 1, 0)).all()
    gpu.close(); orc.close()


def test_pageable_upload_whose_luma_fills_a_staging_region():
    """4096x4096 from pageable memory with a host inv_qscale table: W*H equals a quarter of a 64 MiB staging buffer, so the
    table's staging allocation used to land on the first bytes of the staged luma"""
    W, H = 4096, 4096
    rng = np.random.default_rng(12)
    luma = rng.integers(0, 256, (H, W), dtype=np.uint8)
    gpu, orc = _pair(W, H, bframes=3)
    q = rng.integers(128, 512, gpu.nmb).astype(np.uint16)
    gpu.put_frame(0, luma, q)
    orc.put_frame(0, luma, q)
    for pl in range(4):
        assert (gpu.lowres(0, pl) == orc.lowres(0, pl)).all(), pl
    assert gpu.frame_cost(0, 0, 0) == orc.frame_cost(0, 0, 0)
    assert (gpu.results[(0, 0, 0)]["intra_costs"] == orc.get(0, GET_INTRA)).all()
    gpu.close(); orc.close()


def test_device_aq_mode0_and_zero_strength():
    W, H = 320, 192
    luma, chroma = synth_frames(W, H, 1, seed=4), synth_chroma(W, H, 1)
    gpu, orc = _pair(W, H, bframes=3)
    for mode, strength in [(0, 1.0), (1, 0.0)]:
        gpu.put_frame_yuv(0, luma[0], chroma[0], aq_mode=mode, aq_strength=strength)
        orc.put_frame_yuv(0, luma[0], chroma[0], aq_mode=mode, aq_strength=strength)
        (qg, ig, sg), (qo_, io, so) = gpu.get_aq(0), orc.get_aq(0)
        assert (qg == qo_).all() and (ig == io).all() and (sg == so).all()
    gpu.close(); orc.close()


@pytest.mark.parametrize("mode", [2, 3])
@pytest.mark.parametrize("W,H", [(320, 192), (200, 120), (1920, 1080)])
def test_device_auto_variance_modes_match_oracle(W, H, mode):
    """aq-mode 2 / 3 (ratecontrol.c:355-392) on the device: powf(energy + 1, 0.125f) reproduced bit for bit, the frame
    averages accumulated in the reference's raster order, then the offsets; with and without quant_offsets, zero strength"""
    n = 3
    luma = synth_frames(W, H, n, seed=13)
    chroma = synth_chroma(W, H, n)
    luma[1][: H // 3] = 17
    luma[1][H // 3: H // 2, : W // 2] = 255
    gpu, orc = _pair(W, H, bframes=3, aq=mode)
    qo = np.linspace(-2.5, 2.5, gpu.nmb).astype(np.float32)
    for i in range(n):
        offs = qo if i == 1 else None
        strength = 0.0 if i == 2 else 1.3
        gpu.put_frame_yuv(i, luma[i], chroma[i], aq_mode=mode, aq_strength=strength, quant_offsets=offs)
        orc.put_frame_yuv(i, luma[i], chroma[i], aq_mode=mode, aq_strength=strength, quant_offsets=offs)
        (qg, ig, sg), (qo_, io, so) = gpu.get_aq(i), orc.get_aq(i)
        assert (qg.view(np.uint32) == qo_.view(np.uint32)).all(), i
        assert (ig == io).all() and (sg == so).all(), i
    assert gpu.frame_cost(0, 1, 1) == orc.frame_cost(0, 1, 1)
    gpu.close(); orc.close()


def test_device_powf_eighth_root_over_a_wide_energy_sweep():
    """8160 macroblocks whose AC energies sweep 0 .. ~4e6 (random two-level blocks of random amplitude and density): thousands
    of distinct arguments of powf(energy + 1, 0.125f), against the oracle, which calls libm's powf like the reference.
    (The exception table itself is built from libm over the WHOLE domain 1 .. 2^24, capi.cu ensure_pow_table.)"""
    W, H = 1920, 1088
    rng = np.random.default_rng(21)
    luma = np.zeros((H, W), dtype=np.uint8)
    mbh, mbw = H // 16, W // 16
    for my in range(mbh):                  # per macroblock: amplitude / pattern chosen so that energies cover 0 .. ~4e6 densely
        for mx in range(mbw):
            amp = int(rng.integers(0, 256))
            frac = rng.random()
            blk = (rng.random((16, 16)) < frac).astype(np.uint8) * amp + int(rng.integers(0, 256 - amp))
            luma[16 * my:16 * my + 16, 16 * mx:16 * mx + 16] = blk
    chroma = np.full((H // 2, W), 128, dtype=np.uint8)
    gpu, orc = _pair(W, H, bframes=3, aq=2)
    gpu.put_frame_yuv(0, luma, chroma, aq_mode=2, aq_strength=1.0)
    orc.put_frame_yuv(0, luma, chroma, aq_mode=2, aq_strength=1.0)
    (qg, ig, sg), (qo_, io, so) = gpu.get_aq(0), orc.get_aq(0)
    assert (qg.view(np.uint32) == qo_.view(np.uint32)).all() and (ig == io).all()
    assert len(np.unique(qg)) > 5000
    gpu.close(); orc.close()


@pytest.mark.parametrize("subme", [7, 1])
def test_device_weights_analyse_matches_oracle(subme):
    W, H, n = 320, 192, 5
    seqs = {
        "plain": synth_frames(W, H, n, seed=5),
        "fade": _fade(synth_frames(W, H, n, seed=6, pan=(0, 0), noise=1), 0.09),
        "fade_pan": _fade(synth_frames(W, H, n, seed=7), 0.06),
        "brighten": _fade(synth_frames(W, H, n, seed=8, pan=(1, 0)), -0.05),
    }
    chroma = synth_chroma(W, H, n)
    weighted = 0
    for name, frames in seqs.items():
        gpu, orc = _pair(W, H, bframes=3, subme=subme)
        for i in range(n):
            gpu.put_frame_yuv(i, frames[i], chroma[i])
            orc.put_frame_yuv(i, frames[i], chroma[i])
        for b, r in [(1, 0), (2, 0), (3, 1), (4, 1), (4, 3)]:
            wg, wo = gpu.weights_analyse(b, r), orc.weights_analyse(b, r)
            assert wg == wo, (name, b, r, wg, wo)
            weighted += wg[0]
        gpu.close(); orc.close()
    assert weighted >= 4


@pytest.mark.parametrize("passes", [2, 0])
def test_device_weight_feeds_the_search_without_host_round_trip(passes):
    """analysis -> weighted plane -> weighted list-0 search -> decision, all enqueued before the first flush
    (X264CU_WEIGHT_DEVICE); then the re-analysis on the motion-compensated reference (slicetype.c:76-103)."""
    W, H, n = 320, 192, 3
    frames = _fade(synth_frames(W, H, n, seed=9, pan=(1, 0), noise=1), 0.15)
    chroma = synth_chroma(W, H, n)
    gpu, orc = _pair(W, H, bframes=3, passes=passes)
    for i in range(n):
        gpu.put_frame_yuv(i, frames[i], chroma[i])
        orc.put_frame_yuv(i, frames[i], chroma[i])
    wo = orc.weights_analyse(2, 0)
    assert wo[0] == 1
    w = gpu.weights_analyse(2, 0, flush=False)                    # stays on the device
    assert gpu.frame_cost(0, 2, 2, weight="device") == orc.frame_cost(0, 2, 2, weight=wo)
    assert (w.enabled, w.scale, w.denom, w.offset) == wo         # the host copy arrived with the same flush
    r = gpu.results[(2, 2, 0)]
    assert (r["mvs0"] == orc.get(2, GET_MVS, 0, 1)).all()
    assert (r["lowres_costs"] == orc.get(2, GET_LOWRES_COSTS, 2, 0)).all()
    assert gpu.weights_analyse(2, 0) == orc.weights_analyse(2, 0)          # MC path on both sides
    # a pair whose analysis terminates early: the device weight is "none" and the search must run unweighted
    plain = synth_frames(W, H, n, seed=5)
    gpu2, orc2 = _pair(W, H, bframes=3, passes=passes)
    for i in range(n):
        gpu2.put_frame_yuv(i, plain[i], chroma[i])
        orc2.put_frame_yuv(i, plain[i], chroma[i])
    assert orc2.weights_analyse(1, 0)[0] == 0
    gpu2.weights_analyse(1, 0, flush=False)
    assert gpu2.frame_cost(0, 1, 1, weight="device") == orc2.frame_cost(0, 1, 1)
    assert (gpu2.results[(1, 1, 0)]["mvs0"] == orc2.get(1, GET_MVS, 0, 0)).all()
    for g in (gpu, gpu2):
        g.close()
    orc.close(); orc2.close()


def test_device_weight_requires_analysis():
    W, H = 320, 192
    frames, chroma = synth_frames(W, H, 2, seed=5), synth_chroma(W, H, 2)
    gpu, _ = _pair(W, H, bframes=3)
    for i in range(2):
        gpu.put_frame_yuv(i, frames[i], chroma[i])
    with pytest.raises(RuntimeError, match="weights_analyse"):
        gpu.frame_cost(0, 1, 1, weight="device")
    gpu.close()


def test_weights_analyse_with_host_side_aq_statistics():
    """the unmodified encoder keeps x264_adaptive_quant_frame on the host: plain upload + x264cu_frame_set_pixel_stats"""
    W, H, n = 320, 192, 3
    frames = _fade(synth_frames(W, H, n, seed=6, pan=(0, 0), noise=1), 0.09)
    chroma = synth_chroma(W, H, n)
    gpu, orc = _pair(W, H, bframes=3)
    for i in range(n):
        orc.put_frame_yuv(i, frames[i], chroma[i])
        _, inv, sums = orc.get_aq(i)
        gpu.put_frame(i, frames[i], inv)
        gpu.ctx.set_pixel_stats(gpu.frames[i], sums[0], sums[1])
    for b, r in [(1, 0), (2, 0), (2, 1)]:
        assert gpu.weights_analyse(b, r) == orc.weights_analyse(b, r)
    gpu.close(); orc.close()
