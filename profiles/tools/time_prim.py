#!/usr/bin/env python3
"""time single streaming primitives (frame_init_lowres_core, hadamard_ac) on device-resident stacks > L2;
X264CU_LIB selects a tuning build.  usage: time_prim.py [tag]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "x264-i386pic_b200"))
import x264cu  # noqa: E402

W, nfr = 1920, 64
hh = 1080 * nfr
ctx = x264cu.Context(W, 1080, bframes=3)
L = ctx.L
d1 = ctx.malloc(W * hh + 256)
chunk = np.random.default_rng(5).integers(0, 256, W * 1080 * 8, dtype=np.uint8)
for k in range(nfr // 8):
    x264cu.check(L.x264cu_memcpy_h2d(ctx.ctx, d1 + k * chunk.nbytes, chunk.ctypes.data, chunk.nbytes))


def timed(call, reps=20):
    for _ in range(3):
        x264cu.check(call(), "call")
    ctx.timer_begin()
    for _ in range(reps):
        call()
    return ctx.timer_end() / reps


lw, lh = W // 2 - 8, hh // 2 - 1
dl = [ctx.malloc((W // 2) * (hh // 2) + 256) for _ in range(4)]
ms = timed(lambda: L.x264cu_lowres_core(ctx.ctx, d1, W, dl[0], dl[1], dl[2], dl[3], W // 2, lw, lh, 1))
alg = (2 * lw + 1) * (2 * lh + 1) + 4 * lw * lh
dac = ctx.malloc((W // 16) * (hh // 16) * 8)
ms2 = timed(lambda: L.x264cu_hadamard_ac_grid(ctx.ctx, 0, d1, W, W, hh, dac, 1))
print(sys.argv[1] if len(sys.argv) > 1 else "default", "lowres_core %.4f ms %.0f GB/s" % (ms, alg / ms / 1e6),
      "hadamard_ac_16x16 %.4f ms %.0f GB/s" % (ms2, (W * hh + (W // 16) * (hh // 16) * 8) / ms2 / 1e6))
ctx.close()
