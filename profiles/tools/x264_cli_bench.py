#!/usr/bin/env python
"""BASELINE.json configs[0] on the box: the reference's own CLI (oracle/_ref/x264cu_cli, C-only build: no nasm in the
image) encoding a synthetic 1080p y4m with --preset medium --rc-lookahead 40 --b-adapt 2, once with the CPU lookahead
and once with --opencl, i.e. with the lookahead costs from libx264cuda.so through the per-call hook.
Prints one JSON line per run.  usage: x264_cli_bench.py [frames] [threads]"""
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from bench import make_stream_frames

# the fastest prebuilt flavour of the unmodified sources (oracle/Makefile: _simd = -O3 -ffast-math + AVX2 auto-vectorisation,
# _fast = the reference's release flags, "" = the -O2 parity build)
CLI = next(p for p in (os.path.join(ROOT, "oracle", "_ref", "x264cu_cli" + suf) for suf in ("_simd", "_fast", "")) if os.path.exists(p))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
preset = sys.argv[3] if len(sys.argv) > 3 else "medium"        # a faster preset makes the lookahead the bottleneck
W, H = 1920, 1080
src = "/tmp/x264cu_bench_1080p.y4m"
frames = make_stream_frames(7, n)
with open(src, "wb") as f:
    f.write(b"YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C420jpeg\n" % (W, H))
    c = np.full(W * H // 2, 128, np.uint8).tobytes()
    for x in frames:
        f.write(b"FRAME\n"); f.write(x.tobytes()); f.write(c)
base = [CLI, "--preset", preset, "--rc-lookahead", "40", "--b-adapt", "2", "--bframes", "3", "--threads", str(threads), "-o", "/tmp/x264cu_bench.264", src]
for name, extra, env in (("cpu lookahead (%d lookahead threads auto)" % threads, [], {}),
                         ("b200 lookahead, default schedule", ["--opencl"], {"X264CU_SEARCH_PASSES": "2"}),
                         ("b200 lookahead, exact mode", ["--opencl"], {"X264CU_SEARCH_PASSES": "0"})):
    e = dict(os.environ); e.update(env)
    t0 = time.time()
    r = subprocess.run(base + extra, capture_output=True, text=True, env=e)
    dt = time.time() - t0
    m = re.search(r"encoded (\d+) frames, ([0-9.]+) fps, ([0-9.]+) kb/s", r.stderr)
    ok = r.returncode == 0 and m and "failed to load OpenCL" not in r.stderr and "x264cu:" not in r.stderr
    print(json.dumps({"cli": os.path.basename(CLI), "preset": preset, "run": name, "ok": bool(ok), "frames": int(m.group(1)) if m else None, "fps": float(m.group(2)) if m else None,
                      "kbps": float(m.group(3)) if m else None, "wall_s": round(dt, 2), "threads": threads,
                      "bytes": os.path.getsize("/tmp/x264cu_bench.264") if os.path.exists("/tmp/x264cu_bench.264") else None}))
