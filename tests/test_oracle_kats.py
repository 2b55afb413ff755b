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


# ---- late depth / stencil and blend: the oracle against an independent per-pixel model -------------------------------------
# Flat layers: every layer is one quad (two triangles) over the whole 48x36 viewport with one depth and one colour, so each
# pixel sees exactly one fragment per layer (checked through fragmentsCovered) and the whole frame must equal what a scalar
# model of the reference's fragment epilogue computes for that fragment sequence.

def flat_layers(layers, color_fmt, depth_fmt, width=48, height=36, front_ccw=True):
    s = scenes.random_triangles(width=width, height=height, tris=2 * len(layers), depth_fmt=depth_fmt, color_fmt=color_fmt, perspective=False, seed=0)
    vb = s.buffers["vb"].view(np.float32).reshape(-1, 8)
    corners = [(-1, -1), (-1, 1), (1, 1), (-1, -1), (1, 1), (1, -1)]
    for li, (z, rgba) in enumerate(layers):
        for k, (x, y) in enumerate(corners):
            vb[6 * li + k] = (x, y, z, 1.0) + tuple(rgba)
    return s


def with_edit(scene, fn):
    scene.mutate = fn
    return scene


def stencil_op(op, cur, ref):  # CompileGetStencilResult, PipelineCompiler.cpp:1382-1413 (INC/DEC_CLAMP saturate as signed i8)
    s8 = cur - 256 if cur > 127 else cur
    return {0: cur, 1: 0, 2: ref, 3: min(s8 + 1, 127) & 0xFF, 4: max(s8 - 1, -128) & 0xFF, 5: ~cur & 0xFF, 6: (cur + 1) & 0xFF, 7: (cur - 1) & 0xFF}[op]


def compare_op(op, a, b):  # reference OP stored
    return [False, a < b, a == b, a <= b, a > b, a != b, a >= b, True][op]


@pytest.mark.parametrize("ops", [(0, 2, 0, 7, 0xFF, 0xFF, 0x40), (0, 3, 4, 1, 0xFF, 0xFF, 0x01), (5, 6, 7, 3, 0x0F, 0xF0, 0x05), (1, 0, 2, 5, 0xFF, 0x3C, 0x80),
                                 (2, 2, 2, 0, 0xFF, 0xFF, 0x7F), (3, 3, 3, 6, 0xFF, 0xFF, 0x7E), (4, 4, 4, 4, 0xFF, 0xFF, 0x90)])
@pytest.mark.parametrize("depth_op", [F.LESS, 3, F.GREATER, F.ALWAYS])
def test_depth_stencil_epilogue_against_scalar_model(oracle, ops, depth_op):
    """PipelineCompiler.cpp:1061-1080, :1168-1223, :1246-1413 on D24_UNORM_S8_UINT: stencil compare on masked values, depth
    compare of the fragment against the STORED (quantised) depth, the three stencil ops with the write mask, depth written
    whenever the DEPTH test passes (a reference quirk: the stencil verdict does not gate it), colour only when both pass."""
    fail_op, pass_op, dfail_op, cmp_op, cmask, wmask, ref = ops
    rng = np.random.RandomState(5)
    # distinct, well separated depths: the interpolated depth z*w0 + z*w1 + z*w2 may differ from z in the last bit per pixel,
    # which must not decide a comparison here (the stored code is therefore checked to +-2 codes, everything else exactly)
    layers = [(float(z), tuple(rng.random_sample(4).astype(np.float32))) for z in (0.7, 0.3, 0.5, 0.2, 0.9, 0.1, 0.6)]

    def edit(m):
        m.desc.stencilTestEnable = 1
        m.desc.depthCompareOp = depth_op
        for st in (m.desc.front, m.desc.back):
            st.failOp, st.passOp, st.depthFailOp, st.compareOp, st.compareMask, st.writeMask, st.reference = ops
    sc = flat_layers(layers, F.R8G8B8A8_UNORM, F.D24_UNORM_S8_UINT)
    sc.depth.clear = ("depth", (0.5, 0x7E))
    color, depth, st = scenes.run_oracle(with_edit(sc, edit))
    assert st.fragmentsCovered == 48 * 36 * len(layers)
    # scalar model of one pixel
    d24 = int(np.float32(round(float(np.float32(0.5) * np.float32(16777215.0)))))  # stored depth code of the clear
    sten = 0x7E
    px = tuple(int(v) for v in np.floor((np.array((0.1, 0.2, 0.3, 1.0), dtype=np.float32) * np.float32(255)).astype(np.float64) + 0.5))  # llvm.round: half away
    written = 0
    for z, rgba in layers:
        stored = np.float32(d24) / np.float32(16777215.0)
        s_pass = compare_op(cmp_op, ref & cmask, sten & cmask)
        d_pass = bool(compare_op(depth_op, np.float32(z), stored))
        new = stencil_op(pass_op if d_pass else dfail_op, sten, ref) if s_pass else stencil_op(fail_op, sten, ref)
        sten = (new & wmask) | (sten & ~wmask & 0xFF)
        if d_pass:  # depthWrite = depthResult && shouldAttemptDepthWrite: the stencil verdict is NOT part of it (PipelineCompiler.cpp:1318-1324)
            c = np.minimum(np.maximum(np.float32(z), np.float32(0)), np.float32(1)) * np.float32(16777215.0)
            d24 = int(np.floor(np.float32(c) + np.float32(0.5)))  # llvm.round on a non-negative value with no tie here
        if s_pass and d_pass:  # CompileWriteFragment runs only when both passed
            col = np.asarray(rgba, dtype=np.float32) * np.float32(255.0)
            px = tuple(int(v) for v in np.floor(col.astype(np.float64) + 0.5))
            written += 1
        # a stencil-passing, depth-failing fragment rewrites the stored depth with itself (GlslFunctions.cpp:898-914): no change
    assert st.fragmentsWritten == written * 48 * 36
    got = depth.view(np.uint32)
    assert np.all((got >> 24) == sten), "stencil: oracle %#x model %#x" % (int(got[0] >> 24), sten)
    assert np.all(np.abs((got & 0xFFFFFF).astype(np.int64) - d24) <= 2), "depth code: oracle %d model %d" % (int(got[0] & 0xFFFFFF), d24)
    assert np.all(color.reshape(-1, 4) == np.array(px, dtype=np.uint8)), (color.reshape(-1, 4)[0], px)


def blend_factor(f, s, d, c):  # ApplyBlendFactor, Draw.cpp:956-1103 (colour factor for rgb and, unless overridden, for alpha)
    one = np.float32(1)
    return {0: np.zeros(4, np.float32), 1: np.ones(4, np.float32), 2: s, 3: one - s, 4: d, 5: one - d, 6: np.full(4, s[3]), 7: np.full(4, one - s[3]),
            8: np.full(4, d[3]), 9: np.full(4, one - d[3]), 10: c, 11: one - c, 12: np.full(4, c[3]), 13: np.full(4, one - c[3]),
            14: np.array([min(s[3], one - d[3])] * 3 + [one], np.float32)}[f].astype(np.float32)


def blend_op(op, s, sf, d, df):  # ApplyBlend, Draw.cpp:1105-1262
    if op == 0: return (s * sf).astype(np.float32) + (d * df).astype(np.float32)
    if op == 1: return (s * sf).astype(np.float32) - (d * df).astype(np.float32)
    if op == 2: return (d * df).astype(np.float32) - (s * sf).astype(np.float32)
    return np.minimum(s, d) if op == 3 else np.maximum(s, d)


@pytest.mark.parametrize("blend", [dict(src=1, dst=1, op=0), dict(src=6, dst=7, op=0), dict(src=6, dst=7, op=1), dict(src=4, dst=2, op=2, srcA=1, dstA=0, opA=0),
                                   dict(src=14, dst=1, op=0), dict(src=10, dst=11, op=0, srcA=12, dstA=13, opA=0), dict(src=1, dst=1, op=3),
                                   dict(src=1, dst=1, op=4, srcA=6, dstA=7, opA=0), dict(src=8, dst=9, op=0), dict(src=3, dst=5, op=0)],
                         ids=lambda b: "-".join(str(v) for v in b.values()))
def test_blend_against_scalar_model(oracle, blend):
    """ApplyBlendFactor / ApplyBlend (Draw.cpp:956-1262, dead code in the reference: SURVEY F3) on an RGBA32F target against a
    float32 numpy model: factors per VkBlendFactor, the alpha factor / op overriding component 3 when they differ from the
    colour ones, five layers deep so that a wrong factor or operand order compounds."""
    rng = np.random.RandomState(11)
    layers = [(0.5, tuple(rng.random_sample(4).astype(np.float32))) for _ in range(5)]
    consts = np.array((0.25, 0.5, 0.75, 0.6), dtype=np.float32)

    def edit(m):
        for i, c in enumerate(consts):
            m.desc.blendConstants[i] = float(c)
    sc = flat_layers(layers, F.R32G32B32A32_SFLOAT, None)
    sc.blend = blend
    color, _, st = scenes.run_oracle(with_edit(sc, edit))
    assert st.fragmentsCovered == 48 * 36 * len(layers)
    d = np.array((0.1, 0.2, 0.3, 1.0), dtype=np.float32)  # random_triangles' clear colour
    srcA, dstA, opA = blend.get("srcA", blend["src"]), blend.get("dstA", blend["dst"]), blend.get("opA", blend["op"])
    for _, rgba in layers:
        s = np.asarray(rgba, dtype=np.float32)
        sf, df = blend_factor(blend["src"], s, d, consts), blend_factor(blend["dst"], s, d, consts)
        if srcA != blend["src"]: sf[3] = blend_factor(srcA, s, d, consts)[3]
        if dstA != blend["dst"]: df[3] = blend_factor(dstA, s, d, consts)[3]
        out = blend_op(blend["op"], s, sf, d, df).astype(np.float32)
        if opA != blend["op"]: out[3] = blend_op(opA, s, sf, d, df)[3]
        d = out.astype(np.float32)
    got = color.view(np.float32).reshape(-1, 4)
    # the source colour reaches the blender through perspective interpolation, (w0*c + w1*c + w2*c) / (w0 + w1 + w2), which is c only
    # to the last bit or two and differs per pixel — so this KAT pins factor / op selection and operand order to 1e-5, not the bits
    assert np.allclose(got, np.broadcast_to(d, got.shape), rtol=1e-5, atol=1e-6), (got[0], d)


# ---- points and lines: the oracle's coverage against independent numpy restatements of Draw.cpp:1315-1508 ----

def cvtt(v):  # static_cast<int32_t>(float) on x86: truncation, INT_MIN when out of range or NaN
    v = np.float32(v)
    return -2**31 if (np.isnan(v) or v >= np.float32(2**31) or v < np.float32(-2**31)) else int(np.trunc(v))


def numpy_point_coverage(scene):
    vb = scene.buffers["vb"].view(np.float32).reshape(-1, 9)
    W, H = np.float32(scene.color.width), np.float32(scene.color.height)
    n = 0
    for v in vb[:scene.count]:
        w = v[3]
        X, Y = np.float32(v[0] / w), np.float32(v[1] / w)
        sx = cvtt(np.float32(np.float32(np.float32(X + np.float32(1)) * np.float32(0.5)) * np.float32(W - np.float32(1))))
        sy = cvtt(np.float32(np.float32(np.float32(Y + np.float32(1)) * np.float32(0.5)) * np.float32(H - np.float32(1))))
        size = np.float32(v[8])
        half = cvtt(np.ceil(np.float32(size / np.float32(2))))
        x0, y0 = max(0, sx - half), max(0, sy - half)
        x1, y1 = min(int(W), sx + half + 1), min(int(H), sy + half + 1)
        if x1 <= x0 or y1 <= y0:
            continue
        xs = np.arange(x0, x1, dtype=np.int64); ys = np.arange(y0, y1, dtype=np.int64)
        s = np.float32(0.5) + ((xs - sx).astype(np.float32) / size).astype(np.float32)
        t = np.float32(0.5) + ((ys - sy).astype(np.float32) / size).astype(np.float32)
        n += int(np.count_nonzero((s >= 0) & (s <= 1))) * int(np.count_nonzero((t >= 0) & (t <= 1)))
    return n


def numpy_line_coverage(scene):
    vb = scene.buffers["vb"].view(np.float32).reshape(-1, 9)[:scene.count]
    W, H = np.float32(scene.color.width), np.float32(scene.color.height)
    f = np.float32
    xs = ((np.arange(int(W), dtype=np.float32) / W + (f(1) / W) * f(0.5)) * f(2) - f(1)).astype(np.float32)
    ys = ((np.arange(int(H), dtype=np.float32) / H + (f(1) / H) * f(0.5)) * f(2) - f(1)).astype(np.float32)
    X, Y = np.meshgrid(xs, ys)
    pairs = [(2 * i, 2 * i + 1) for i in range(len(vb) // 2)] if scene.topology == F.LINE_LIST else [(i, i + 1) for i in range(len(vb) - 1)]
    lw = np.array([f(scene.line_width) / W, f(scene.line_width) / H], dtype=np.float32)

    def E(a, b):  # EdgeFunction(a, b, p) = (p.x - a.x) * (b.y - a.y) - (p.y - a.y) * (b.x - a.x), one rounding per operator
        return ((X - a[0]) * f(b[1] - a[1])).astype(np.float32) - ((Y - a[1]) * f(b[0] - a[0])).astype(np.float32)

    n = 0
    for i0, i1 in pairs:
        P0 = np.array([vb[i0][0] / vb[i0][3], vb[i0][1] / vb[i0][3], vb[i0][2] / vb[i0][3], vb[i0][3]], dtype=np.float32)
        P1 = np.array([vb[i1][0] / vb[i1][3], vb[i1][1] / vb[i1][3], vb[i1][2] / vb[i1][3], vb[i1][3]], dtype=np.float32)
        d = (P1 - P0).astype(np.float32)                        # glm::normalize(vec4): x * inversesqrt(dot(x, x)), all four components
        sq = f(f(f(d[0] * d[0]) + f(d[1] * d[1])) + f(d[2] * d[2])) + f(d[3] * d[3])
        inv = f(1) / np.sqrt(f(sq), dtype=np.float32)
        dirx, diry = f(d[0] * inv), f(d[1] * inv)
        perp = np.array([diry, -dirx], dtype=np.float32) * lw
        p00, p01 = (P0[:2] + perp).astype(np.float32), (P0[:2] - perp).astype(np.float32)
        p10, p11 = (P1[:2] + perp).astype(np.float32), (P1[:2] - perp).astype(np.float32)
        inside = (E(p00, p01) >= 0) & (E(p11, p10) >= 0) & (E(p10, p00) >= 0) & (E(p01, p11) >= 0)
        n += int(np.count_nonzero(inside))
    return n


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("perspective", [True, False])
def test_point_coverage_against_numpy(oracle, seed, perspective):
    sc = scenes.random_points_lines(count=80, seed=seed, topology=F.POINT_LIST, depth_fmt=None, perspective=perspective)
    _, _, st = scenes.run_oracle(sc)
    assert st.primitives == 80 and st.fragmentsCovered == numpy_point_coverage(sc)


@pytest.mark.parametrize("topology", [F.LINE_LIST, F.LINE_STRIP])
@pytest.mark.parametrize("width", [1.0, 4.5])
def test_line_coverage_against_numpy(oracle, topology, width):
    sc = scenes.random_points_lines(count=30, seed=5, topology=topology, line_width=width, depth_fmt=None, perspective=True)
    _, _, st = scenes.run_oracle(sc)
    assert st.fragmentsCovered == numpy_line_coverage(sc)


# ---- interpolation and depth: GetFragmentInput / SetDatum (Draw.cpp:816-954) on one triangle, value by value ----

@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("perspective", [True, False])
def test_interpolated_colour_and_depth_against_numpy(oracle, seed, perspective):
    """Every covered pixel of a single triangle: weights w_i = E_i / area, depth = z0*w0 + z1*w1 + z2*w2, and each colour
    component numerator / denominator with numerator = ((0 + w0*v0/pw0) + w1*v1/pw1) + w2*v2/pw2 and the matching denominator —
    float32, one rounding per operator, in the reference's order; RGBA32F colour and D32 depth compared bit for bit."""
    f = np.float32
    sc = scenes.random_triangles(width=40, height=28, tris=1, seed=seed, perspective=perspective, color_fmt=F.R32G32B32A32_SFLOAT,
                                 depth_fmt=F.D32_SFLOAT, depth_op=F.ALWAYS)
    color, depth, st = scenes.run_oracle(sc)
    W, H = 40, 28
    vb = sc.buffers["vb"].view(np.float32).reshape(-1, 8)[:3]
    P = [np.array([v[0] / v[3], v[1] / v[3], v[2] / v[3], v[3]], dtype=np.float32) for v in vb]
    xs = ((np.arange(W, dtype=np.float32) / f(W) + (f(1) / f(W)) * f(0.5)) * f(2) - f(1)).astype(np.float32)
    ys = ((np.arange(H, dtype=np.float32) / f(H) + (f(1) / f(H)) * f(0.5)) * f(2) - f(1)).astype(np.float32)
    X, Y = np.meshgrid(xs, ys)

    def E(a, b, cx, cy):
        return ((cx - a[0]) * f(b[1] - a[1])).astype(np.float32) - ((cy - a[1]) * f(b[0] - a[0])).astype(np.float32)

    area = f(f((P[2][0] - P[0][0]) * f(P[1][1] - P[0][1])) - f((P[2][1] - P[0][1]) * f(P[1][0] - P[0][0])))
    if area < 0:
        area = -area
        w = [E(P[2], P[1], X, Y), E(P[0], P[2], X, Y), E(P[1], P[0], X, Y)]
    else:
        w = [E(P[1], P[2], X, Y), E(P[2], P[0], X, Y), E(P[0], P[1], X, Y)]
    inside = ~((w[0] < 0) | (w[1] < 0) | (w[2] < 0))
    # the integer bounding box of the reference's pixel loop (Draw.cpp:1548-1569) also limits coverage
    sx = [int(np.trunc(f(f(f(p[0] + f(1)) * f(0.5)) * f(W)))) for p in P]; sy = [int(np.trunc(f(f(f(p[1] + f(1)) * f(0.5)) * f(H)))) for p in P]
    box = np.zeros((H, W), dtype=bool)
    box[max(0, min(sy)):min(H, max(sy) + 1), max(0, min(sx)):min(W, max(sx) + 1)] = True
    inside &= box
    assert st.fragmentsCovered == int(np.count_nonzero(inside))
    with np.errstate(all="ignore"):
        w = [(wi / area).astype(np.float32) for wi in w]
        z = ((P[0][2] * w[0]).astype(np.float32) + (P[1][2] * w[1]).astype(np.float32)).astype(np.float32) + (P[2][2] * w[2]).astype(np.float32)
        den = np.zeros_like(X)
        for i in range(3):
            den = (den + (w[i] / P[i][3]).astype(np.float32)).astype(np.float32)
        want = np.zeros((H, W, 4), dtype=np.float32)
        for c in range(4):
            num = np.zeros_like(X)
            for i in range(3):
                num = (num + ((w[i] * vb[i][4 + c]).astype(np.float32) / P[i][3]).astype(np.float32)).astype(np.float32)
            want[:, :, c] = (num / den).astype(np.float32)
    got_c = color.view(np.float32).reshape(H, W, 4); got_z = depth.view(np.float32).reshape(H, W)
    assert np.array_equal(got_c[inside].view(np.uint32), want[inside].view(np.uint32))
    assert np.array_equal(got_z[inside].view(np.uint32), z.astype(np.float32)[inside].view(np.uint32))
    assert np.all(got_z[~inside] == f(1.0)) and np.all(got_c[~inside] == np.array((0.1, 0.2, 0.3, 1.0), dtype=np.float32))


# ---- more codecs: SNORM, 16-bit UNORM / SNORM, BGRA ordering, A2B10G10R10 (ImageCompiler.cpp:15-53, :160-301, :1077-1348) ----

def pack_raw(oracle, fmt, texel, values):
    v = np.ascontiguousarray(values, dtype=np.float32).reshape(-1, 4)
    out = np.zeros(len(v) * texel, dtype=np.uint8)
    oracle.cpvk_oracle_pack_f32(fmt, v.ctypes.data_as(C.c_void_p), len(v), out.ctypes.data_as(C.c_void_p))
    return out


def unpack_raw(oracle, fmt, texel, raw):
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    n = len(raw) // texel
    out = np.zeros((n, 4), dtype=np.float32)
    oracle.cpvk_oracle_unpack_f32(fmt, raw.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p))
    return out


def probe_values(scale):
    """Inputs around every kind of boundary for a normalised format with `scale` = 2^n - 1 or 2^(n-1) - 1."""
    rng = np.random.RandomState(int(scale) & 0xFFFF)
    k = rng.randint(0, int(scale) + 1, size=400).astype(np.float64)
    core = np.concatenate([k / scale, (k + 0.5) / scale, (k + 0.5) / scale * (1 + 2e-7), (k + 0.5) / scale * (1 - 2e-7)])
    v = np.concatenate([core, -core, [0.0, -0.0, 1.0, -1.0, 1.5, -1.5, 1e30, -1e30, np.inf, -np.inf, np.nan]]).astype(np.float32)
    pad = (-len(v)) % 4
    return np.concatenate([v, np.zeros(pad, dtype=np.float32)]).reshape(-1, 4)


def model_norm(v, lo, scale):  # minnum(maxnum(v, lo), 1) * scale, llvm.round, fptoui / fptosi
    c = np.where(np.isnan(v), np.float32(lo), v)
    c = np.minimum(np.maximum(c, np.float32(lo)), np.float32(1)).astype(np.float32)
    return round_half_away((c * np.float32(scale)).astype(np.float32)).astype(np.int64)


@pytest.mark.parametrize("fmt,dtype,texel,lo,scale", [(38, np.int8, 4, -1.0, 127.0), (91, np.uint16, 8, 0.0, 65535.0), (92, np.int16, 8, -1.0, 32767.0)],
                         ids=["R8G8B8A8_SNORM", "R16G16B16A16_UNORM", "R16G16B16A16_SNORM"])
def test_normalised_pack_and_unpack_against_numpy(oracle, fmt, dtype, texel, lo, scale):
    v = probe_values(scale)
    raw = pack_raw(oracle, fmt, texel, v)
    want = model_norm(v.reshape(-1), lo, scale).astype(dtype)
    assert np.array_equal(raw.view(dtype), want)
    codes = np.arange(np.iinfo(dtype).min, np.iinfo(dtype).max + 1, dtype=np.int64)
    codes = np.concatenate([codes, np.zeros((-len(codes)) % 4, dtype=np.int64)]).astype(dtype)
    back = unpack_raw(oracle, fmt, texel, codes.view(np.uint8))
    # sitofp / uitofp, then a true divide by the same constant (ImageCompiler.cpp:41-53): -128 / 127 < -1 is NOT clamped
    assert np.array_equal(back.reshape(-1), (codes.astype(np.float32) / np.float32(scale)).astype(np.float32))


def test_bgra8_component_order(oracle):
    v = np.array([[0.0, 1.0 / 255, 2.0 / 255, 3.0 / 255], [1.0, 0.5, 0.25, 0.125]], dtype=np.float32)
    rgba, bgra = pack_raw(oracle, 37, 4, v).reshape(-1, 4), pack_raw(oracle, 44, 4, v).reshape(-1, 4)
    assert np.array_equal(bgra, rgba[:, [2, 1, 0, 3]])  # B8G8R8A8: blue in byte 0 (Formats.cpp table: RedOffset 2, BlueOffset 0)
    assert np.array_equal(unpack_raw(oracle, 44, 4, bgra.reshape(-1)), unpack_raw(oracle, 37, 4, rgba.reshape(-1)))


def test_a2b10g10r10_pack_and_unpack_against_numpy(oracle):
    rng = np.random.RandomState(9)
    v = np.concatenate([rng.uniform(-0.2, 1.2, size=(500, 4)), [[0, 0, 0, 0], [1, 1, 1, 1], [0.5 / 1023, 1.5 / 1023, 1022.5 / 1023, 0.5 / 3]]]).astype(np.float32)
    raw = pack_raw(oracle, 64, 4, v).view(np.uint32)
    r, g, b = (model_norm(v[:, i], 0.0, 1023.0) for i in range(3))
    a = model_norm(v[:, 3], 0.0, 3.0)
    assert np.array_equal(raw, (r | (g << 10) | (b << 20) | (a << 30)).astype(np.uint32))  # A2B10G10R10: red in the low bits
    back = unpack_raw(oracle, 64, 4, raw.view(np.uint8))
    want = np.stack([r.astype(np.float32) / np.float32(1023), g.astype(np.float32) / np.float32(1023), b.astype(np.float32) / np.float32(1023),
                     a.astype(np.float32) / np.float32(3)], axis=1).astype(np.float32)
    assert np.array_equal(back, want)


# ---- vkCmdBlitImage coordinate arithmetic (CommandBuffer.cpp:186-226) with NEAREST taps, RGBA32F to RGBA32F ----

@pytest.mark.parametrize("case", [((12, 9), (12, 9), (0, 0, 12, 9), (0, 0, 12, 9)),          # 1:1
                                  ((12, 9), (30, 20), (0, 0, 12, 9), (0, 0, 30, 20)),        # magnify
                                  ((31, 17), (10, 6), (0, 0, 31, 17), (0, 0, 10, 6)),        # minify
                                  ((16, 16), (20, 20), (3, 2, 13, 11), (4, 5, 17, 19)),      # sub-rectangles
                                  ((16, 16), (16, 16), (0, 0, 16, 16), (16, 16, 0, 0)),      # destination mirrored in x and y
                                  ((16, 16), (16, 16), (16, 0, 0, 16), (0, 0, 16, 16))],     # source mirrored in x
                         ids=["same", "magnify", "minify", "subrect", "dst-mirror", "src-mirror"])
def test_blit_nearest_against_numpy(oracle, case):
    """u = (dstX + 0.5f - dst0.x) * (float(src1.x - src0.x) / (dst1.x - dst0.x)) + src0.x, the coordinate handed to the sampler
    is u / width, and NEAREST takes floor(coordinate * width) clamped to the edge — float32, in that order. A mirrored
    destination walks x + dst1.x (the reference's negativeWidth branch). Texels are raw floats, so the copy must be exact."""
    f = np.float32
    (sw, sh), (dw, dh), (sx0, sy0, sx1, sy1), (dx0, dy0, dx1, dy1) = case
    rng = np.random.RandomState(sw * 31 + dw)
    src = rng.uniform(-4, 4, size=(sh, sw, 4)).astype(np.float32)
    dst = rng.uniform(-4, 4, size=(dh, dw, 4)).astype(np.float32)
    want = dst.copy()

    def axis(d0, d1, s0, s1, size_src):
        n = abs(d1 - d0)
        out = []
        for i in range(n):
            d = i + (d1 if d1 - d0 < 0 else d0)
            scale = f(f(s1 - s0) / f(d1 - d0))
            u = f(f(f(f(d) + f(0.5)) - f(d0)) * scale) + f(s0)
            t = int(np.floor(f(f(u / f(size_src)) * f(size_src))))
            out.append((d, min(max(t, 0), size_src - 1)))
        return out

    for dy, ty in axis(dy0, dy1, sy0, sy1, sh):
        for dx, tx in axis(dx0, dx1, sx0, sx1, sw):
            if 0 <= dx < dw and 0 <= dy < dh:
                want[dy, dx] = src[ty, tx]
    b = capi.Blit(capi.Attachment(src.ctypes.data, sw, sh, sw * 16, 109), capi.Attachment(dst.ctypes.data, dw, dh, dw * 16, 109),
                  sx0, sy0, sx1, sy1, dx0, dy0, dx1, dy1, 0)
    assert oracle.cpvk_oracle_blit(C.byref(b)) == 0
    assert np.array_equal(dst.view(np.uint32), want.view(np.uint32))


# ---- the sampling wrapper: LOD bias / clamp, mip choice and view swizzle (GlslFunctions.cpp:539-555, :598-654) ----

MAX_BIAS = np.float32(32.0)


def level_chain():
    """Four mip levels (8x8 .. 1x1), RGBA32F, every texel of level L = (L, 10 + L, 20 + L, 30 + L): the sampled value names the level."""
    return [np.ascontiguousarray(np.broadcast_to(np.array([l, 10 + l, 20 + l, 30 + l], dtype=np.float32), (8 >> l, 8 >> l, 4))) for l in range(4)]


def chain_descriptor(levels, mipmap, bias, min_lod, max_lod, swizzle=(0, 0, 0, 0)):
    d = capi.Descriptor()
    d.type, d.format, d.dimensions, d.levelCount = capi.DESC_IMAGE, 109, 2, len(levels)
    for i, t in enumerate(levels):
        d.levels[i] = capi.MipLevel(t.ctypes.data, t.shape[1], t.shape[0], 1, 0)
    s = d.sampler
    s.magFilter = s.minFilter = F.NEAREST
    s.mipmapMode, s.mipLodBias, s.minLod, s.maxLod = mipmap, bias, min_lod, max_lod
    for i, c in enumerate(swizzle):
        d.swizzle[i] = c
    return d


@pytest.mark.parametrize("mipmap", [0, 1], ids=["mip-nearest", "mip-linear"])
def test_lod_bias_clamp_and_mip_choice_against_numpy(oracle, mipmap):
    f = np.float32
    levels = level_chain()
    coords = np.array([[0.3, 0.6, 0.0]], dtype=np.float32)
    for lod in (-1.0, 0.0, 0.2, 0.5, 0.75, 1.0, 1.49, 1.5, 2.2, 3.0, 7.0):
        for bias, lo, hi in ((0.0, 0.0, 1000.0), (0.6, 0.0, 1000.0), (-0.75, 0.0, 1000.0), (50.0, 0.0, 1000.0), (0.0, 1.25, 2.5), (1.0, 0.5, 0.5), (0.0, 0.0, 0.0)):
            d = chain_descriptor(levels, mipmap, bias, lo, hi)
            got = np.zeros((1, 4), dtype=np.float32)
            oracle.cpvk_oracle_sample(C.byref(d), coords.ctypes.data_as(C.c_void_p), 1, C.c_float(lod), got.ctypes.data_as(C.c_void_p))
            lam = f(f(lod) + min(max(f(f(bias) + f(0)), -MAX_BIAS), MAX_BIAS))  # MAX_SAMPLER_LOD_BIAS = 32 (Config.h:156)
            lam = min(max(lam, f(lo)), f(hi))                      # std::clamp(lambdaPrime, minLod, maxLod)
            if lam <= 0:
                want = f(0)
            else:
                m = min(max(lam, f(0)), f(len(levels) - 1))
                if mipmap == 0:
                    want = f(int(np.ceil(f(m + f(0.5)))) - 1)      # ceil(l + 0.5) - 1 (ImageSampler.cpp:632-637)
                else:
                    l1 = int(np.floor(m)); delta = f(m - f(l1))
                    want = f(l1) if delta == 0 else f(np.float64(l1) + np.float64(f(f(l1 + 1) - f(l1))) * np.float64(delta))  # double lerp of the two levels
            assert got[0, 0] == want and got[0, 3] == f(30) + want, (lod, bias, lo, hi, got[0], want)


def test_view_swizzle_against_numpy(oracle):
    levels = level_chain()[:1]
    coords = np.array([[0.5, 0.5, 0.0]], dtype=np.float32)
    base = np.array([0, 10, 20, 30], dtype=np.float32)
    pick = {0: None, 1: 0.0, 2: 1.0, 3: base[0], 4: base[1], 5: base[2], 6: base[3]}  # IDENTITY, ZERO, ONE, R, G, B, A
    for swz in ((0, 0, 0, 0), (3, 4, 5, 6), (6, 5, 4, 3), (1, 2, 0, 3), (4, 4, 4, 2), (0, 3, 0, 1)):
        d = chain_descriptor(levels, 0, 0.0, 0.0, 0.0, swz)
        got = np.zeros((1, 4), dtype=np.float32)
        oracle.cpvk_oracle_sample(C.byref(d), coords.ctypes.data_as(C.c_void_p), 1, C.c_float(0.0), got.ctypes.data_as(C.c_void_p))
        want = [base[i] if pick[s] is None else pick[s] for i, s in enumerate(swz)]
        assert list(got[0]) == want, (swz, got[0], want)


# ---- shader runtime math (a14): the reference's GLSL.std.450 subset and OpDot, value by value ----

def test_glsl_std_450_subset_against_numpy(oracle):
    """glslmath.frag evaluated by the oracle's SPIR-V interpreter against a float32 numpy model that spells each function the
    way the reference does: std::min / std::max / std::clamp comparison forms, Mix = x*(1-a) + y*a, NMin / NMax / NClamp,
    glm::normalize = v * (1 / sqrt(dot)), glm::reflect = I - N * dot(N, I) * 2, glm::dot = (x + y) + z resp. (x + y) + (z + w)
    (GlslFunctions.cpp:19-321, SpirvFunctions.cpp:6-60). The per-pixel input colour is taken from a run of cube.frag on the
    same scene, which writes the interpolated input unchanged."""
    f = np.float32
    sc = scenes.random_triangles(width=48, height=36, tris=20, seed=70, color_fmt=F.R32G32B32A32_SFLOAT, depth_fmt=None)
    cin, _, st0 = scenes.run_oracle(sc)
    sc.fs = "glslmath.frag"
    out, _, st = scenes.run_oracle(sc)
    assert st.fragmentsCovered == st0.fragmentsCovered > 1000
    c = cin.view(np.float32).reshape(-1, 4); got = out.view(np.float32).reshape(-1, 4)
    clear = np.array((0.1, 0.2, 0.3, 1.0), dtype=np.float32)
    drawn = np.any(c != clear, axis=1)  # last writer wins in both runs (no depth, no blend), so rows correspond pixel by pixel
    c = c[drawn]; got = got[drawn]
    mn = lambda x, y: np.where(y < x, y, x); mx = lambda x, y: np.where(x < y, y, x)
    clampf = lambda v, lo, hi: np.where(v < lo, lo, np.where(hi < v, hi, v))
    a = ((c * f(4)).astype(np.float32) - f(2)).astype(np.float32)
    c3q = (c[:, :3] + f(0.25)).astype(np.float32)
    dd = (((c3q[:, 0] * c3q[:, 0]).astype(np.float32) + (c3q[:, 1] * c3q[:, 1]).astype(np.float32)).astype(np.float32) + (c3q[:, 2] * c3q[:, 2]).astype(np.float32)).astype(np.float32)
    inv = (f(1) / np.sqrt(dd, dtype=np.float32)).astype(np.float32)
    n = (c3q * inv[:, None]).astype(np.float32)
    d = (((n[:, 0] * a[:, 0]).astype(np.float32) + (n[:, 1] * a[:, 1]).astype(np.float32)).astype(np.float32) + (n[:, 2] * a[:, 2]).astype(np.float32)).astype(np.float32)
    r = (a[:, :3] - ((n * d[:, None]).astype(np.float32) * f(2)).astype(np.float32)).astype(np.float32)
    f0 = (mn(a[:, 0], a[:, 1]) + mx(a[:, 2], a[:, 3])).astype(np.float32)
    nclamp = mn(mx(a[:, 0], f(-0.5)), f(0.75))  # no NaN in this scene: NMin / NMax reduce to min / max
    f1 = ((mn(a[:, 0], a[:, 3]) + mx(a[:, 1], a[:, 2])).astype(np.float32) + nclamp).astype(np.float32)
    dot4 = (((a[:, 0] * c[:, 0]).astype(np.float32) + (a[:, 1] * c[:, 1]).astype(np.float32)).astype(np.float32) +
            ((a[:, 2] * c[:, 2]).astype(np.float32) + (a[:, 3] * c[:, 3]).astype(np.float32)).astype(np.float32)).astype(np.float32)
    f2 = ((clampf(a[:, 1], f(-1), f(0.5)) + np.abs(a[:, 2])).astype(np.float32) + dot4).astype(np.float32)
    cw = c[:, ::-1]
    m = ((a * (f(1) - cw).astype(np.float32)).astype(np.float32) + (c * cw).astype(np.float32)).astype(np.float32)
    i = np.trunc((a * f(100)).astype(np.float32)).astype(np.int64); u = np.trunc((c * f(1000)).astype(np.float32)).astype(np.int64)
    s = np.abs(i[:, 0]) + np.sign(i[:, 1]) + np.minimum(i[:, 2], i[:, 3]) + np.maximum(i[:, 0], i[:, 2]) + np.clip(i[:, 3], -50, 60)
    q = np.minimum(u[:, 0], u[:, 1]) + np.maximum(u[:, 2], u[:, 3]) + np.clip(u[:, 0], 100, 700)
    want = np.stack([((f0 + f1).astype(np.float32) + r[:, 0]).astype(np.float32),
                     (((f2 + r[:, 1]).astype(np.float32) + m[:, 0]).astype(np.float32) + m[:, 1]).astype(np.float32),
                     (((r[:, 2] + m[:, 2]).astype(np.float32) + m[:, 3]).astype(np.float32) + s.astype(np.float32)).astype(np.float32),
                     q.astype(np.float32)], axis=1)
    bad = np.nonzero(np.any(got.view(np.uint32) != want.view(np.uint32), axis=1))[0]
    assert len(bad) == 0, "%d of %d pixels differ; first: c=%s oracle=%s model=%s" % (len(bad), len(got), c[bad[0]], got[bad[0]], want[bad[0]])


# ---- matrix arithmetic in shaders (a14, @Matrix.Mult.*, SpirvFunctions.cpp:6-60 -> glm) ----

def matmath_scene(W=40, H=30, seed=12):
    """A point list, one vertex per pixel of a sparse grid, size-1 points (exactly one pixel each: Draw.cpp:1345-1361), colours
    produced by matmath.vert: (a * b * k) * inColor + pos * b."""
    rng = np.random.RandomState(seed)
    sc = scenes.random_points_lines(width=W, height=H, count=1, seed=seed, topology=F.POINT_LIST, depth_fmt=None, color_fmt=F.R32G32B32A32_SFLOAT, perspective=False)
    sc.vs = "matmath.vert"
    pix = [(i, j) for j in range(1, H - 1, 3) for i in range(1, W - 1, 3)]
    n = len(pix)
    f = np.float32
    x = np.array([(i + 0.25) * 2.0 / (W - 1) - 1.0 for i, _ in pix], dtype=np.float32)
    y = np.array([(j + 0.25) * 2.0 / (H - 1) - 1.0 for _, j in pix], dtype=np.float32)
    pos = np.stack([x, y, rng.uniform(0, 1, n).astype(np.float32), np.ones(n, dtype=np.float32)], axis=1)
    col = rng.uniform(-1, 1, size=(n, 4)).astype(np.float32)
    vb = np.concatenate([pos, col, np.ones((n, 1), dtype=np.float32)], axis=1).astype(np.float32)
    sc.buffers["vb"] = vb.view(np.uint8).reshape(-1)
    sc.count = n
    a = rng.uniform(-1.5, 1.5, size=(4, 4)).astype(np.float32)  # [column][row], as std140 column-major stores it
    b = rng.uniform(-1.5, 1.5, size=(4, 4)).astype(np.float32)
    k = f(0.7)
    ubo = np.zeros(36, dtype=np.float32)
    ubo[0:16] = a.reshape(-1); ubo[16:32] = b.reshape(-1); ubo[32] = k
    sc.buffers["ubo"] = ubo.view(np.uint8)
    sc.uniforms = [(0, 0, "ubo")]
    return sc, pix, pos, col, a, b, k


def test_matrix_products_against_numpy(oracle):
    """glm 0.9.5's operand order (the copy vendored with the reference's samples, detail/type_mat4x4.inl:620-780):
    mat*mat column j = ((A0*B[j][0] + A1*B[j][1]) + A2*B[j][2]) + A3*B[j][3]; mat*scalar per element; mat*vec =
    (m0*v0 + m1*v1) + (m2*v2 + m3*v3); vec*mat component j = ((m[j][0]*v0 + m[j][1]*v1) + m[j][2]*v2) + m[j][3]*v3."""
    f = np.float32
    sc, pix, pos, col, a, b, k = matmath_scene()
    out, _, st = scenes.run_oracle(sc)
    assert st.fragmentsCovered == len(pix)
    img = out.view(np.float32).reshape(sc.color.height, sc.color.width, 4)
    mm = np.zeros((4, 4), dtype=np.float32)
    for j in range(4):
        acc = (a[0] * b[j][0]).astype(np.float32)
        for r in range(1, 4):
            acc = (acc + (a[r] * b[j][r]).astype(np.float32)).astype(np.float32)
        mm[j] = acc
    ms = (mm * k).astype(np.float32)
    for (i, j), p, c in zip(pix, pos, col):
        mv = (((ms[0] * c[0]).astype(np.float32) + (ms[1] * c[1]).astype(np.float32)).astype(np.float32) +
              ((ms[2] * c[2]).astype(np.float32) + (ms[3] * c[3]).astype(np.float32)).astype(np.float32)).astype(np.float32)
        vm = np.zeros(4, dtype=np.float32)
        for q in range(4):
            acc = f(b[q][0] * p[0])
            for r in range(1, 4):
                acc = f(acc + f(b[q][r] * p[r]))
            vm[q] = acc
        want = (mv + vm).astype(np.float32)
        assert np.array_equal(img[j, i].view(np.uint32), want.view(np.uint32)), ((i, j), img[j, i], want)


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("perspective", [True, False])
def test_line_interpolation_and_depth_against_numpy(oracle, seed, perspective):
    """One wide line: t = dot(p - p0, p1 - p0) / (length(p1 - p0) * length(p1 - p0)), colour = SetDatum<true, 2> with weights
    (1 - t, t), depth = p0.z * t + p1.z * (1 - t) — the weights swapped, as the reference has it (Draw.cpp:1440-1494)."""
    f = np.float32
    sc = scenes.random_points_lines(width=40, height=28, count=1, seed=seed, topology=F.LINE_LIST, line_width=5.0, depth_fmt=F.D32_SFLOAT,
                                    color_fmt=F.R32G32B32A32_SFLOAT, perspective=perspective)
    sc.depth_op = F.ALWAYS
    color, depth, st = scenes.run_oracle(sc)
    W, H = 40, 28
    vb = sc.buffers["vb"].view(np.float32).reshape(-1, 9)[:2]
    P = [np.array([v[0] / v[3], v[1] / v[3], v[2] / v[3], v[3]], dtype=np.float32) for v in vb]
    xs = ((np.arange(W, dtype=np.float32) / f(W) + (f(1) / f(W)) * f(0.5)) * f(2) - f(1)).astype(np.float32)
    ys = ((np.arange(H, dtype=np.float32) / f(H) + (f(1) / f(H)) * f(0.5)) * f(2) - f(1)).astype(np.float32)
    X, Y = np.meshgrid(xs, ys)
    d = (P[1] - P[0]).astype(np.float32)
    sq = f(f(f(d[0] * d[0]) + f(d[1] * d[1])) + f(d[2] * d[2])) + f(d[3] * d[3])
    inv = f(1) / np.sqrt(f(sq), dtype=np.float32)
    perp = np.array([f(d[1] * inv), -f(d[0] * inv)], dtype=np.float32) * np.array([f(5.0) / f(W), f(5.0) / f(H)], dtype=np.float32)
    p00, p01 = (P[0][:2] + perp).astype(np.float32), (P[0][:2] - perp).astype(np.float32)
    p10, p11 = (P[1][:2] + perp).astype(np.float32), (P[1][:2] - perp).astype(np.float32)
    E = lambda a, b: ((X - a[0]) * f(b[1] - a[1])).astype(np.float32) - ((Y - a[1]) * f(b[0] - a[0])).astype(np.float32)
    inside = (E(p00, p01) >= 0) & (E(p11, p10) >= 0) & (E(p10, p00) >= 0) & (E(p01, p11) >= 0)
    assert st.fragmentsCovered == int(np.count_nonzero(inside)) > 20
    dx, dy = f(P[1][0] - P[0][0]), f(P[1][1] - P[0][1])
    length = np.sqrt(f(f(dx * dx) + f(dy * dy)), dtype=np.float32)
    with np.errstate(all="ignore"):
        t = ((((X - P[0][0]).astype(np.float32) * dx).astype(np.float32) + ((Y - P[0][1]).astype(np.float32) * dy).astype(np.float32)).astype(np.float32)
             / f(length * length)).astype(np.float32)
        w = [(f(1) - t).astype(np.float32), t]
        z = ((P[0][2] * t).astype(np.float32) + (P[1][2] * w[0]).astype(np.float32)).astype(np.float32)
        den = np.zeros_like(X)
        for i in range(2):
            den = (den + (w[i] / P[i][3]).astype(np.float32)).astype(np.float32)
        want = np.zeros((H, W, 4), dtype=np.float32)
        for c in range(4):
            num = np.zeros_like(X)
            for i in range(2):
                num = (num + ((w[i] * vb[i][4 + c]).astype(np.float32) / P[i][3]).astype(np.float32)).astype(np.float32)
            want[:, :, c] = (num / den).astype(np.float32)
    got_c = color.view(np.float32).reshape(H, W, 4); got_z = depth.view(np.float32).reshape(H, W)
    assert np.array_equal(got_c[inside].view(np.uint32), want[inside].view(np.uint32))
    assert np.array_equal(got_z[inside].view(np.uint32), z[inside].view(np.uint32))


# ---- sRGB (ImageCompiler.cpp:103-158, :1350-1383): piecewise 0.04045 / 12.92 / 2.4 with llvm.pow — libm, so 1e-5, not bits ----

def test_srgb_unpack_and_pack_against_numpy(oracle):
    f = np.float32
    codes = np.arange(256, dtype=np.uint8)
    raw = np.stack([codes, codes[::-1], (codes * 7).astype(np.uint8), codes], axis=1).reshape(-1)
    got = unpack_raw(oracle, 43, 4, raw)  # R8G8B8A8_SRGB
    v = (raw.reshape(-1, 4).astype(np.float32) / f(255.0)).astype(np.float32)
    lin = np.where(v > f(0.04045), np.power(((v + f(0.055)).astype(np.float32) / f(1.055)).astype(np.float64), 2.4), (v / f(12.92)).astype(np.float64))
    assert np.allclose(got[:, :3], lin[:, :3], rtol=1e-5, atol=1e-7)
    assert np.array_equal(got[:, 3], v[:, 3])  # alpha is linear (the conversion skips channel 3)
    # pack: clamp, linear -> sRGB on rgb, * 255, llvm.round; compared through the code it produces (+-1 only where pow's last bits decide a tie)
    x = np.linspace(0, 1, 1001, dtype=np.float32)
    vals = np.stack([x, x[::-1], (x * x).astype(np.float32), x], axis=1)
    packed = pack_raw(oracle, 43, 4, vals).reshape(-1, 4)
    s = np.where(vals > f(0.0031308), np.power(vals.astype(np.float64), 1.0 / 2.4) * 1.055 - 0.055, vals.astype(np.float64) * 12.92)
    want_rgb = np.floor(s[:, :3] * 255.0 + 0.5)
    assert np.max(np.abs(packed[:, :3].astype(np.int64) - want_rgb.astype(np.int64))) <= 1
    assert np.mean(packed[:, :3] == want_rgb) > 0.995
    assert np.array_equal(packed[:, 3], np.floor((vals[:, 3] * f(255.0)).astype(np.float64) + 0.5).astype(np.uint8))
    # round trip of every code is the identity
    back = pack_raw(oracle, 43, 4, got).reshape(-1, 4)
    assert np.array_equal(back, raw.reshape(-1, 4))


def test_uniform_arrays_with_std140_stride_against_premultiplied_colours(oracle):
    """uboarray.vert computes outColor = inColor * p.tint[int(inColor.x * 3.99)] + vec4(p.scale[2]) from a uniform block in
    descriptor set 1 whose float array has ArrayStride 16. The same frame must come out of the plain cube shaders when the host
    applies that formula to the vertex colours beforehand (float32, one rounding per operation) — which pins the array stride,
    the dynamic index and the second descriptor set of the interpreter without reference to its own addressing code."""
    f = np.float32
    sc = scenes.ubo_arrays(160, 120)
    got, gd, st = scenes.run_oracle(sc)
    params = sc.buffers["params"].view(np.float32)
    tint = params[0:16].reshape(4, 4); scale2 = params[24]
    pre = scenes.draw_cube(160, 120)
    vb = pre.buffers["vb"].view(np.float32).reshape(-1, 8).copy()
    col = vb[:, 4:8]
    idx = np.trunc((col[:, 0] * f(3.99)).astype(np.float32)).astype(np.int64)
    vb[:, 4:8] = ((col * tint[idx]).astype(np.float32) + scale2).astype(np.float32)
    pre.buffers["vb"] = vb.view(np.uint8).reshape(-1)
    want, wd, st2 = scenes.run_oracle(pre)
    assert st.fragmentsCovered == st2.fragmentsCovered > 500
    assert np.array_equal(got, want) and np.array_equal(gd, wd)
    assert len(set(idx.tolist())) >= 2  # the dynamic index really varies over the vertices
