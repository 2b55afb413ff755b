"""Known-answer tests that pin the CPU oracle. The reference ships no golden vectors for this path, so
each KAT checks the oracle against an INDEPENDENT restatement of the cited reference arithmetic written here in
numpy (float32, one rounding per operator), plus hand-derived corner cases (ties, shared edges, zero area, NaN)."""
import ctypes as C

import numpy as np
import pytest

from cpvulkan_b200 import capi, scenes

F = scenes


def pack(oracle, fmt, values):
    v = np.ascontiguousarray(values, dtype=np.float32).reshape(-1, 4)
    out = np.zeros(len(v) * scenes.TEXEL_SIZE[fmt], dtype=np.uint8)
    oracle.cpvk_oracle_pack_f32(fmt, v.ctypes.data_as(C.c_void_p), len(v), out.ctypes.data_as(C.c_void_p))
    return out


def unpack(oracle, fmt, raw):
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    n = len(raw) // scenes.TEXEL_SIZE[fmt]
    out = np.zeros((n, 4), dtype=np.float32)
    oracle.cpvk_oracle_unpack_f32(fmt, raw.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p))
    return out


def round_half_away(x):  # llvm.round (ImageCompiler.cpp:20-32)
    x = np.asarray(x, dtype=np.float32).astype(np.float64)  # exact in double; float32 'x + 0.5' would round for large x
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def test_unorm8_unpack_all_codes(oracle):
    codes = np.arange(256, dtype=np.uint8)
    raw = np.stack([codes, codes, codes, codes], axis=1).reshape(-1)
    got = unpack(oracle, F.R8G8B8A8_UNORM, raw)
    want = (codes.astype(np.float32) / np.float32(255.0)).astype(np.float32)  # uitofp / 255.0f (ImageCompiler.cpp:49-53)
    assert np.array_equal(got[:, 0].view(np.uint32), want.view(np.uint32)) and np.array_equal(got[:, 3], want)


def test_unorm8_pack_boundaries_ties_clamp_nan(oracle):
    k = np.arange(256, dtype=np.float32)
    exact = (k / np.float32(255.0)).astype(np.float32)
    vals = np.concatenate([exact, np.nextafter(exact, np.float32(2)), np.nextafter(exact, np.float32(-1)),
                           ((k + np.float32(0.5)) / np.float32(255.0)).astype(np.float32),
                           np.array([-1.0, 2.0, np.nan, np.inf, -np.inf, 0.5, -0.0], dtype=np.float32)])
    quad = np.stack([vals] * 4, axis=1)
    got = pack(oracle, F.R8G8B8A8_UNORM, quad).reshape(-1, 4)[:, 0]
    c = vals.copy()
    c[np.isnan(c)] = 0            # maxnum(NaN, 0) = 0
    c = np.minimum(np.maximum(c, np.float32(0)), np.float32(1))
    want = round_half_away((c * np.float32(255.0)).astype(np.float32)).astype(np.uint8)
    assert np.array_equal(got, want)
    # BGRA places red at byte 2 (Formats.cpp:255)
    one = pack(oracle, F.B8G8R8A8_UNORM, [[1.0, 0.5, 0.0, 0.25]])
    assert list(one) == [0, 128, 255, 64]


def test_unorm8_roundtrip_is_identity(oracle):
    codes = np.arange(256, dtype=np.uint8)
    raw = np.stack([codes] * 4, axis=1).reshape(-1)
    assert np.array_equal(pack(oracle, F.R8G8B8A8_UNORM, unpack(oracle, F.R8G8B8A8_UNORM, raw)), raw)


def test_half_all_codes_against_numpy(oracle):
    codes = np.arange(65536, dtype=np.uint16)
    got = np.array([oracle.cpvk_oracle_half_to_float(int(c)) for c in codes[::7]], dtype=np.float32)
    want = codes[::7].view(np.float16).astype(np.float32)
    ok = np.isnan(want) | (got.view(np.uint32) == want.view(np.uint32))
    assert ok.all()
    # round trip of every non-NaN half code is the identity
    finite = codes[~np.isnan(codes.view(np.float16))][::5]
    back = np.array([oracle.cpvk_oracle_float_to_half(float(np.float32(np.array([c], dtype=np.uint16).view(np.float16)[0]))) for c in finite], dtype=np.uint16)
    assert np.array_equal(back, finite)


def test_float_to_half_rtne_against_numpy(oracle):
    rng = np.random.RandomState(3)
    bits = rng.randint(0, 2 ** 32, size=20000, dtype=np.uint64).astype(np.uint32)
    vals = bits.view(np.float32)
    vals = vals[~np.isnan(vals)]
    extra = np.array([65504.0, 65519.9, 65520.0, 1e-8, 5.96e-8, 2.98e-8, 2.99e-8, 6.1e-5, -0.0, 0.0, np.inf, -np.inf, 1.0009765625, 1.00048828125], dtype=np.float32)
    vals = np.concatenate([vals, extra])
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).view(np.uint16)  # numpy converts RTNE with denormals, like FloatFormat.h:138-255
    got = np.array([oracle.cpvk_oracle_float_to_half(float(v)) for v in vals], dtype=np.uint16)
    assert np.array_equal(got, want)


def test_rgba16f_pack_unpack(oracle):
    v = np.array([[0.1, -2.5, 1000.0, 1.0 / 3.0]], dtype=np.float32)
    raw = pack(oracle, F.R16G16B16A16_SFLOAT, v)
    assert np.array_equal(raw.view(np.uint16), v.astype(np.float16).view(np.uint16).reshape(-1))
    assert np.array_equal(unpack(oracle, F.R16G16B16A16_SFLOAT, raw), v.astype(np.float16).astype(np.float32))


def test_depth_codecs(oracle):
    d = np.array([0.0, 1.0, 0.5, 0.25, 1.5, -1.0, 0.3333333, 0.9999999], dtype=np.float32)
    for fmt, size in ((F.D16_UNORM, 2), (F.D32_SFLOAT, 4), (F.D24_UNORM_S8_UINT, 4)):
        out = np.zeros(len(d) * size, dtype=np.uint8)
        st = np.full(len(d), 0xAB, dtype=np.uint8)
        oracle.cpvk_oracle_pack_depth(fmt, d.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p), len(d), out.ctypes.data_as(C.c_void_p))
        back = np.zeros(len(d), dtype=np.float32)
        oracle.cpvk_oracle_unpack_depth(fmt, out.ctypes.data_as(C.c_void_p), len(d), back.ctypes.data_as(C.c_void_p))
        c = np.minimum(np.maximum(d, np.float32(0)), np.float32(1))
        if fmt == F.D16_UNORM:
            code = round_half_away((c * np.float32(65535.0)).astype(np.float32))
            assert np.array_equal(out.view(np.uint16), code.astype(np.uint16))
            assert np.array_equal(back, (code.astype(np.float32) / np.float32(65535.0)).astype(np.float32))
        elif fmt == F.D32_SFLOAT:
            assert np.array_equal(back, d)  # raw float, no clamp (ImageCompiler.cpp:926-929)
        else:
            code = round_half_away((c * np.float32(16777215.0)).astype(np.float32)).astype(np.uint32)
            assert np.array_equal(out.view(np.uint32), code | np.uint32(0xAB << 24))


def make_descriptor(tex, filt, address, fmt=F.R8G8B8A8_UNORM, border=0):
    d = capi.Descriptor()
    d.type, d.format, d.dimensions, d.levelCount = capi.DESC_IMAGE, fmt, 2, 1
    d.levels[0] = capi.MipLevel(tex.ctypes.data, tex.shape[1], tex.shape[0], 1, 0)
    d.sampler.magFilter = d.sampler.minFilter = filt
    d.sampler.addressModeU = d.sampler.addressModeV = d.sampler.addressModeW = address
    d.sampler.borderColor = border
    return d


def np_wrap(v, size, mode):  # ImageSampler.cpp:12-38 with C (truncating) % semantics
    cmod = lambda a, b: int(np.fmod(a, b))
    if mode == F.REPEAT:
        return cmod(cmod(v, size) + size, size)
    if mode == F.MIRRORED_REPEAT:
        n = cmod(cmod(v, 2 * size) + 2 * size, 2 * size) - size
        return size - 1 - (n if n >= 0 else -(1 + n))
    if mode == F.CLAMP_TO_EDGE:
        return min(max(v, 0), size - 1)
    if mode == F.CLAMP_TO_BORDER:
        return min(max(v, -1), size)
    return min(max(v if v >= 0 else -(1 + v), 0), size - 1)


@pytest.mark.parametrize("mode", [F.REPEAT, F.MIRRORED_REPEAT, F.CLAMP_TO_EDGE, F.CLAMP_TO_BORDER, F.MIRROR_CLAMP_TO_EDGE])
@pytest.mark.parametrize("filt", [F.NEAREST, F.LINEAR])
def test_sampler_against_numpy(oracle, mode, filt):
    rng = np.random.RandomState(5)
    size = 8
    tex = rng.randint(0, 256, size=(size, size, 4), dtype=np.uint8)
    d = make_descriptor(tex, filt, mode, border=2)  # FLOAT_OPAQUE_BLACK
    us = np.array([-1.25, -0.5 / size, 0.0, 0.5 / size, 0.37, 1 - 1e-7, 1.0, 2.3], dtype=np.float32)
    coords = np.array([[u, v, 0] for u in us for v in us], dtype=np.float32)
    got = np.zeros((len(coords), 4), dtype=np.float32)
    oracle.cpvk_oracle_sample(C.byref(d), coords.ctypes.data_as(C.c_void_p), len(coords), C.c_float(0.0), got.ctypes.data_as(C.c_void_p))
    texf = (tex.astype(np.float32) / np.float32(255.0)).astype(np.float32)
    border = np.array([0, 0, 0, 1], dtype=np.float32)

    def texel(i, j):
        return border if (i < 0 or i >= size or j < 0 or j >= size) else texf[j, i]

    def lerp(a, b, t):  # ImageSampler.cpp:51-55
        return (a.astype(np.float64) + (b - a).astype(np.float32).astype(np.float64) * np.float64(t)).astype(np.float32)

    for k, (u, v, _) in enumerate(coords):
        if filt == F.NEAREST:
            i = np_wrap(int(np.floor(np.float32(u * np.float32(size)))), size, mode)
            j = np_wrap(int(np.floor(np.float32(v * np.float32(size)))), size, mode)
            want = texel(i, j)
        else:
            su, sv = np.float32(u * np.float32(size)) - np.float32(0.5), np.float32(v * np.float32(size)) - np.float32(0.5)
            i0, j0 = int(np.floor(su)), int(np.floor(sv))
            i1, j1 = np_wrap(i0 + 1, size, mode), np_wrap(j0 + 1, size, mode)
            i0, j0 = np_wrap(i0, size, mode), np_wrap(j0, size, mode)
            fx, fy = np.float32(su - np.floor(su)), np.float32(sv - np.floor(sv))
            want = lerp(lerp(texel(i0, j0), texel(i1, j0), fx), lerp(texel(i0, j1), texel(i1, j1), fx), fy)
        assert np.array_equal(got[k].view(np.uint32), np.asarray(want, dtype=np.float32).view(np.uint32)), (mode, filt, u, v, got[k], want)


def numpy_coverage(scene):
    """Independent float32 restatement of Draw.cpp:1541-1592 + 879-903: returns the number of covered (pixel, triangle)
    pairs for identity-MVP triangle lists."""
    W, H = np.float32(scene.color.width), np.float32(scene.color.height)
    vb = scene.buffers["vb"].view(np.float32).reshape(-1, 8)
    n_cov = 0
    xs = np.arange(scene.color.width, dtype=np.float32)
    ys = np.arange(scene.color.height, dtype=np.float32)
    xf = ((xs / W + (np.float32(1.0) / W) * np.float32(0.5)) * np.float32(2) - np.float32(1)).astype(np.float32)
    yf = ((ys / H + (np.float32(1.0) / H) * np.float32(0.5)) * np.float32(2) - np.float32(1)).astype(np.float32)
    X, Y = np.meshgrid(xf, yf)
    for t in range(scene.count // 3):
        idx = [3 * t, 3 * t + 1, 3 * t + 2]
        if scene.front_face == F.FRONT_CW:
            idx[0], idx[2] = idx[2], idx[0]
        P = [(vb[i, :4] / vb[i, 3]).astype(np.float32) for i in idx]
        sx = [int(np.float32((p[0] + np.float32(1)) * np.float32(0.5) * W)) for p in P]
        sy = [int(np.float32((p[1] + np.float32(1)) * np.float32(0.5) * H)) for p in P]
        x0, x1 = max(0, min(sx)), min(int(W), max(sx) + 1)
        y0, y1 = max(0, min(sy)), min(int(H), max(sy) + 1)
        if x1 <= x0 or y1 <= y0:
            continue
        E = lambda a, b, cx, cy: ((cx - a[0]) * (b[1] - a[1])).astype(np.float32) - ((cy - a[1]) * (b[0] - a[0])).astype(np.float32)
        area = E(P[0], P[1], np.float32(P[2][0]), np.float32(P[2][1]))
        front = not (area < 0)
        if (scene.cull & 2 and not front) or (scene.cull & 1 and front):
            continue
        cx, cy = X[y0:y1, x0:x1], Y[y0:y1, x0:x1]
        if front:
            w = (E(P[1], P[2], cx, cy), E(P[2], P[0], cx, cy), E(P[0], P[1], cx, cy))
        else:
            w = (E(P[2], P[1], cx, cy), E(P[0], P[2], cx, cy), E(P[1], P[0], cx, cy))
        n_cov += int(np.count_nonzero(~((w[0] < 0) | (w[1] < 0) | (w[2] < 0))))
    return n_cov


@pytest.mark.parametrize("kw", [dict(seed=1), dict(seed=2, cull=F.CULL_BACK, front_face=F.FRONT_CW), dict(seed=4, snap=True, perspective=False, width=64, height=48),
                                dict(seed=6, cull=F.CULL_FRONT)], ids=["plain", "cullback_cw", "snapped", "cullfront"])
def test_coverage_count_against_numpy(oracle, kw):
    scene = scenes.random_triangles(tris=120, depth_fmt=None, **kw)
    _, _, st = scenes.run_oracle(scene)
    assert st.fragmentsCovered == numpy_coverage(scene)
    assert st.fragmentsWritten == st.fragmentsCovered  # no depth test: everything that is covered is written


def test_shared_edge_pixels_are_hit_twice(oracle):
    """No top-left rule (F1): two triangles sharing the diagonal of a pixel-centre-aligned square both cover the
    pixel centres on that diagonal."""
    s = scenes.random_triangles(width=8, height=8, tris=2, depth_fmt=None, perspective=False, seed=0)
    W = 8.0
    c = lambda p: (np.float32(p) / np.float32(W) + np.float32(0.5 / W)) * np.float32(2) - np.float32(1)
    a, b = c(1), c(5)
    quad = [(a, a), (a, b), (b, b), (a, a), (b, b), (b, a)]
    vb = s.buffers["vb"].view(np.float32).reshape(-1, 8)
    for i, (x, y) in enumerate(quad):
        vb[i, :4] = (x, y, 0.5, 1.0)
    _, _, st = scenes.run_oracle(s)
    # 5x5 pixel centres inside the square, the 5 on the diagonal counted by both triangles
    assert st.fragmentsCovered == 25 + 5


def test_zero_area_triangle_yields_nan_weights_but_is_drawn(oracle):
    s = scenes.random_triangles(width=8, height=8, tris=1, depth_fmt=None, perspective=False, seed=0)
    vb = s.buffers["vb"].view(np.float32).reshape(-1, 8)
    W = 8.0
    cx = (np.float32(3) / np.float32(W) + np.float32(0.5 / W)) * np.float32(2) - np.float32(1)
    for i in range(3):
        vb[i, :4] = (cx, cx, 0.5, 1.0)  # all three vertices on the centre of pixel (3,3): area == 0, w = 0/0
    color, _, st = scenes.run_oracle(s)
    assert st.fragmentsCovered == 1  # NaN compares false, so the fragment is accepted (Draw.cpp:900)
    px = color.reshape(8, 8, 4)[3, 3]
    assert list(px) == [0, 0, 0, 0]  # NaN colour packs to 0 through maxnum(NaN, 0)


def test_cube_scene_fragment_count_and_faces(oracle):
    color, depth, st = scenes.run_oracle(scenes.draw_cube())
    assert st.primitives == 12 and st.fragmentsCovered == st.fragmentsWritten
    img = color.reshape(500, 500, 4)
    faces = {tuple(c) for c in np.unique(img.reshape(-1, 4), axis=0)}
    assert (51, 51, 51, 51) in faces and len(faces) == 4  # clear colour 0.2 -> 51, three visible faces
    d16 = depth.view(np.uint16)
    assert d16.max() == 65535 and d16.min() < 65535
