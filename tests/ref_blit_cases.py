"""Seeded vkCmdBlitImage cases for tests/test_reference_blit.py and tests/golden/make_ref_golden.py: R32G32B32A32_SFLOAT images,
one region each — enlargements, reductions, 1:1, sub-rectangles at offsets, flipped extents on either side, 1-texel sources and
destinations, both filters. header = (srcW, srcH, dstW, dstH, filter, srcX0, srcY0, srcX1, srcY1, dstX0, dstY0, dstX1, dstY1)."""
import numpy as np


def cases():
    rng = np.random.default_rng(20261018)
    out = []

    def add(sw, sh, dw, dh, filt, s, d):
        src = rng.uniform(-2.0, 2.0, size=(sh, sw, 4)).astype(np.float32)
        if len(out) % 5 == 0:  # large and tiny magnitudes now and then: the lerp chain runs in double and rounds once per step
            src *= np.float32(3.0e7)
        if len(out) % 7 == 0:
            src *= np.float32(1.0e-30)
        dst = rng.uniform(-1.0, 1.0, size=(dh, dw, 4)).astype(np.float32)
        out.append((np.array([sw, sh, dw, dh, filt, s[0], s[1], s[2], s[3], d[0], d[1], d[2], d[3]], dtype=np.int32), src, dst))

    for filt in (0, 1):
        add(8, 8, 8, 8, filt, (0, 0, 8, 8), (0, 0, 8, 8))                # 1:1
        add(7, 5, 19, 13, filt, (0, 0, 7, 5), (0, 0, 19, 13))            # enlarge, odd sizes
        add(23, 17, 6, 9, filt, (0, 0, 23, 17), (0, 0, 6, 9))            # reduce
        add(16, 12, 20, 20, filt, (3, 2, 11, 9), (5, 4, 18, 15))         # sub-rectangles at offsets, the rest of the destination untouched
        add(9, 9, 14, 10, filt, (9, 0, 0, 9), (0, 0, 14, 10))            # source flipped in x
        add(9, 9, 14, 10, filt, (0, 9, 9, 0), (2, 1, 12, 9))             # source flipped in y
        add(10, 6, 12, 12, filt, (0, 0, 10, 6), (12, 0, 0, 12))          # destination flipped in x
        add(10, 6, 12, 12, filt, (1, 1, 9, 5), (2, 11, 10, 3))           # destination flipped in y, sub-rectangle
        add(11, 7, 13, 9, filt, (10, 6, 1, 0), (12, 8, 0, 1))            # both flipped in both axes
        add(1, 1, 6, 4, filt, (0, 0, 1, 1), (0, 0, 6, 4))                # one source texel
        add(12, 10, 1, 1, filt, (0, 0, 12, 10), (0, 0, 1, 1))            # one destination texel
        add(5, 4, 31, 3, filt, (0, 0, 5, 4), (0, 0, 31, 3))              # enlarge x, reduce y
        for _ in range(6):                                               # random regions, either orientation
            sw, sh, dw, dh = (int(v) for v in rng.integers(2, 24, 4))
            sx = sorted(int(v) for v in rng.choice(sw + 1, 2, replace=False)); sy = sorted(int(v) for v in rng.choice(sh + 1, 2, replace=False))
            dx = sorted(int(v) for v in rng.choice(dw + 1, 2, replace=False)); dy = sorted(int(v) for v in rng.choice(dh + 1, 2, replace=False))
            if rng.random() < 0.3: sx.reverse()
            if rng.random() < 0.3: sy.reverse()
            if rng.random() < 0.3: dx.reverse()
            if rng.random() < 0.3: dy.reverse()
            add(sw, sh, dw, dh, filt, (sx[0], sy[0], sx[1], sy[1]), (dx[0], dy[0], dx[1], dy[1]))
    return out


def payload(cs):
    parts = [np.array([len(cs)], dtype=np.uint32).tobytes()]
    for h, src, dst in cs:
        parts += [h.tobytes(), src.tobytes(), dst.tobytes()]
    return b"".join(parts)
