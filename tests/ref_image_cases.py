"""Seeded cases for tests/test_reference_image.py and tests/golden/make_ref_golden.py: the reference's shader-side image functions
(GlslFunctions.cpp:324-737) on R32G32B32A32_SFLOAT data.
  kind 0  ImageSampleExplicitLod on a 2-D image view: explicit lods, LOD bias (also beyond +-MAX_SAMPLER_LOD_BIAS), min / max LOD
          clamps, views that start at a higher mip level and hold fewer levels, component swizzles, filters, mipmap modes, address modes
  kind 1  ImageFetch on an image view (level 0 of the VIEW, coordinates in and out of range, swizzles)
  kind 2  ImageFetch on a uniform texel buffer view (offset and range in bytes, indices in and out of range)
A case = (header u32[19], floats f32[3] = bias, minLod, maxLod, bytes, coords u32-view[n, 3])."""
import numpy as np

REMAINING = 0xFFFFFFFF


def mip_sizes(w, h, mips):
    out = []
    for _ in range(mips):
        out.append((w, h))
        w, h = max(w // 2, 1), max(h // 2, 1)
    return out


def cases():
    rng = np.random.default_rng(324737)
    out = []

    def image_bytes(w, h, mips):
        return b"".join(rng.uniform(-3.0, 3.0, size=(lh, lw, 4)).astype(np.float32).tobytes() for lw, lh in mip_sizes(w, h, mips))

    def add(kind, w, h, mips, base, count, swz, sampler, lods3, data, coords, view=(0, 0)):
        hdr = np.array([kind, w, h, mips, base, count, swz[0], swz[1], swz[2], swz[3], sampler[0], sampler[1], sampler[2], sampler[3], sampler[4],
                        sampler[5], len(coords), view[0], view[1]], dtype=np.uint32)
        out.append((hdr, np.array(lods3, dtype=np.float32), data, np.ascontiguousarray(coords)))

    swizzles = [(0, 0, 0, 0), (3, 4, 5, 6), (6, 5, 4, 3), (1, 2, 3, 3), (0, 3, 0, 2), (5, 0, 1, 0), (4, 4, 4, 4)]
    w, h, mips = 16, 8, 4
    data = image_bytes(w, h, mips)
    n = 24
    for i in range(60):  # kind 0
        base = int(rng.integers(0, 3))
        count = REMAINING if i % 3 == 0 else int(rng.integers(1, mips - base + 1))
        sampler = (int(rng.integers(0, 2)), int(rng.integers(0, 2)), int(rng.integers(0, 2)), int(rng.choice([0, 1, 2, 3])), int(rng.choice([0, 1, 2, 3])), int(rng.choice([0, 2, 4])))
        bias = float(rng.choice([0.0, 0.7, -1.3, 0.25, 40.0, -40.0, 1.0]))
        lo = float(rng.choice([0.0, 0.0, 0.5, 1.0]))
        hi = float(rng.choice([1000.0, 1000.0, 0.25, 1.5, 2.0]))
        if hi < lo:
            lo, hi = hi, lo
        uv = rng.uniform(-0.6, 1.6, size=(n, 2)).astype(np.float32)
        lod = rng.choice(np.array([0.0, 0.5, 1.0, 1.49, 1.5, 2.5, -1.0, 7.0, 0.999], dtype=np.float32), size=(n, 1))
        coords = np.concatenate([uv, lod], axis=1).astype(np.float32).view(np.uint32)
        add(0, w, h, mips, base, count, swizzles[i % len(swizzles)], sampler, (bias, lo, hi), data, coords)
    for i in range(16):  # kind 1
        base = int(rng.integers(0, 3))
        lw, lh = mip_sizes(w, h, mips)[base]
        xy = np.stack([rng.integers(-2, lw + 2, size=n), rng.integers(-2, lh + 2, size=n), np.zeros(n, dtype=np.int64)], axis=1).astype(np.int32)
        add(1, w, h, mips, base, REMAINING, swizzles[i % len(swizzles)], (0, 0, 0, 0, 0, 0), (0.0, 0.0, 1000.0), data, xy.view(np.uint32))
    buf = rng.uniform(-3.0, 3.0, size=(40, 4)).astype(np.float32).tobytes()
    for off, rng_bytes in ((0, 640), (16, 160), (160, 480), (48, 16)):  # kind 2
        texels = rng_bytes // 16
        idx = np.stack([rng.integers(-2, texels + 3, size=n), np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)], axis=1).astype(np.int32)
        add(2, 1, 1, 1, 0, 1, (0, 0, 0, 0), (0, 0, 0, 0, 0, 0), (0.0, 0.0, 1000.0), buf, idx.view(np.uint32), view=(off, rng_bytes))
    return out


def payload(cs):
    parts = [np.array([len(cs)], dtype=np.uint32).tobytes()]
    for hdr, f3, data, coords in cs:
        parts += [hdr.tobytes(), f3.tobytes(), np.array([len(data)], dtype=np.uint32).tobytes(), data, coords.tobytes()]
    return b"".join(parts)
