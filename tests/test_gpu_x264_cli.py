"""The drop-in boundary end to end: the reference's UNMODIFIED encoder and CLI sources, linked with
x264-i386pic_b200/host/x264_glue.c in place of common/opencl.c + encoder/slicetype-cl.c (oracle/Makefile `cli`),
encode a clip with `--opencl`, i.e. with every lookahead cost coming from libx264cuda.so on the B200.

* exact search mode (X264CU_SEARCH_PASSES=0): costs, MVs, slicetype decisions, MB-tree and therefore every slice NAL
  must be byte-identical to the CPU-only run of the same binary (`--lookahead-threads 1`);
* default dependency-free schedule: a valid stream of the same length whose size stays within a few percent.
The binary is test infrastructure (it contains the reference); the product is the .so it links."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from checkers import ROOT, synth_frames

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "oracle", "_ref", "x264cu_cli")


def _y4m(path, W, H, n, seed, **kw):
    frames = synth_frames(W, H, n, seed=seed, **kw)
    with open(path, "wb") as f:
        f.write(b"YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C420jpeg\n" % (W, H))
        for i, x in enumerate(frames):
            f.write(b"FRAME\n")
            f.write(x.tobytes())
            c = np.full((H // 2, W // 2), 128, np.uint8)
            c[:, : W // 4] = 96 + 2 * i          # some chroma structure for AQ
            f.write(c.tobytes())
            f.write((255 - c).tobytes())


def _nals(path):
    """Annex-B NAL units as (type, payload bytes)"""
    data = open(path, "rb").read()
    out, i, starts = [], 0, []
    while True:
        j = data.find(b"\x00\x00\x01", i)
        if j < 0:
            break
        starts.append(j + 3)
        i = j + 3
    for k, s in enumerate(starts):
        e = (starts[k + 1] - 3) if k + 1 < len(starts) else len(data)
        nal = data[s:e].rstrip(b"\x00") if k + 1 < len(starts) else data[s:e]
        out.append((nal[0] & 0x1F, nal))
    return out


def _encode(tmp, name, src, args, env=None):
    out = os.path.join(tmp, name)
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([CLI] + args + ["--threads", "1", "-o", out, src], capture_output=True, text=True, env=e, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return out, r.stderr


@pytest.mark.skipif(not os.path.exists(CLI), reason="oracle/_ref/x264cu_cli not built (make -C oracle cli)")
@pytest.mark.parametrize("W,H,n,extra", [
    (352, 288, 40, ["--preset", "medium"]),                                       # b-adapt 1, weightp 2, mbtree, aq 1
    (320, 192, 36, ["--preset", "medium", "--b-adapt", "2", "--bframes", "5"]),   # trellis: slicetype_prep batches the searches
    (200, 120, 30, ["--preset", "fast", "--no-mbtree"]),                           # no edges, not a multiple of 16
])
def test_x264_cli_exact_mode_is_bit_identical_to_cpu(tmp_path, W, H, n, extra):
    src = str(tmp_path / "in.y4m")
    _y4m(src, W, H, n, seed=31, scene_cut_at=n // 2, fade_at=4)
    cpu, _ = _encode(str(tmp_path), "cpu.264", src, extra + ["--lookahead-threads", "1"])
    # the glue selects the exact mode from the encoder's own parameters (one lookahead thread + deterministic encode);
    # the environment override is exercised by the other two cases
    env = {} if extra == ["--preset", "medium"] else {"X264CU_SEARCH_PASSES": "0"}
    gpu, log = _encode(str(tmp_path), "gpu.264", src, extra + ["--opencl"], env)
    assert "failed to load OpenCL" not in log and "x264cu:" not in log, log[-1500:]
    a = [(t, hashlib.md5(p).hexdigest()) for t, p in _nals(cpu) if t in (1, 5, 7, 8)]      # slices, SPS, PPS
    b = [(t, hashlib.md5(p).hexdigest()) for t, p in _nals(gpu) if t in (1, 5, 7, 8)]
    assert len(a) >= n and a == b          # (the SEI differs: it records opencl=1 in the option string, base.c:1446)


@pytest.mark.skipif(not os.path.exists(CLI), reason="oracle/_ref/x264cu_cli not built (make -C oracle cli)")
def test_x264_cli_default_schedule_encodes(tmp_path):
    W, H, n = 352, 288, 40
    src = str(tmp_path / "in.y4m")
    _y4m(src, W, H, n, seed=32, scene_cut_at=20)
    cpu, _ = _encode(str(tmp_path), "cpu.264", src, ["--preset", "medium", "--lookahead-threads", "1"])
    gpu, log = _encode(str(tmp_path), "gpu.264", src, ["--preset", "medium", "--opencl"], {"X264CU_SEARCH_PASSES": "2"})
    assert "failed to load OpenCL" not in log and "x264cu:" not in log, log[-1500:]
    a, b = _nals(cpu), _nals(gpu)
    assert sum(t in (1, 5) for t, _ in a) == sum(t in (1, 5) for t, _ in b) == n
    sa, sb = os.path.getsize(cpu), os.path.getsize(gpu)
    assert abs(sa - sb) <= 0.05 * sa, (sa, sb)
