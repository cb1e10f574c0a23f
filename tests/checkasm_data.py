"""checkasm-style inputs (tools/checkasm.c:368-397,3038-3058): one random buffer pbuf1 with
pbuf2 = pbuf1 + 0xf00, and the two overflow buffers pbuf3/pbuf4 made of complementary
0/PIXEL_MAX patterns.  The C program seeds rand(); here the bytes come from numpy PCG64 with a
fixed seed and are stored in the golden files, so every consumer sees the same data."""
import numpy as np

SIZE_WH = [(16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4), (4, 16)]
BUF_LEN = 0x1e00
PBUF2_OFF = 0xf00


def make_buffers(seed=1234):
    rng = np.random.default_rng(seed)
    pbuf1 = rng.integers(0, 256, BUF_LEN + 64, dtype=np.uint8)
    pbuf3 = np.zeros(0x1000 + 512, dtype=np.uint8)
    pbuf4 = np.zeros(0x1000 + 512, dtype=np.uint8)
    for i in range(256):                      # "maximize sum" (checkasm.c:369-376)
        z = i | (i >> 4)
        z ^= z >> 2
        z ^= z >> 1
        pbuf4[i] = 255 if (z & 1) else 0
        pbuf3[i] = 255 - pbuf4[i]
    for i in range(256, 0x1000):              # random maxed differences (:378-382)
        pbuf4[i] = 255 if (pbuf1[i & ~0x88] & 1) else 0
        pbuf3[i] = 255 - pbuf4[i]
    return pbuf1, pbuf3, pbuf4


def pixel_cases(size):
    """[(buffer-pair id, stride1, stride2, off1[], off2[])] as TEST_PIXEL issues them."""
    groups = []
    j = np.arange(64)
    wide = (j & 31) == 31
    groups.append(("rand", 16, 64, np.zeros((~wide).sum(), dtype=np.int64), (PBUF2_OFF + j[~wide]).astype(np.int64)))
    groups.append(("rand", 32, 64, np.zeros(wide.sum(), dtype=np.int64), (PBUF2_OFF + j[wide]).astype(np.int64)))
    o = np.arange(0, 0x1000, 256, dtype=np.int64)
    w, h = SIZE_WH[size]
    o = o[o + 15 * 16 + 16 <= 0x1000 + 512]
    groups.append(("ovf", 16, 16, o, o.copy()))
    return groups


def dct_overflow_blocks():
    """crafted extreme residual blocks in the spirit of checkasm.c:908-929: every fenc/fdec pair at
    the ends of the pixel range in a few spatial patterns."""
    pats = []
    yy, xx = np.mgrid[0:16, 0:16]
    for k in range(5):
        if k == 0:
            m = np.zeros((16, 16), bool)
        elif k == 1:
            m = np.ones((16, 16), bool)
        elif k == 2:
            m = (xx + yy) % 2 == 0
        elif k == 3:
            m = xx % 2 == 0
        else:
            m = yy % 2 == 0
        a = np.where(m, 255, 0).astype(np.uint8)
        pats.append((a, (255 - a).astype(np.uint8)))
    return pats
