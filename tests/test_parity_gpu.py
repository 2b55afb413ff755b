"""GPU parity: the sm_100a draw path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): coverage, depth results and UNORM / integer attachments bit-exact; float attachments
within 2 ULP (RGBA16F here is produced by one IEEE float->half rounding of identical float values, so it is
compared bit-exact as well).
"""
import numpy as np
import pytest

from cpvulkan_b200 import scenes
from cpvulkan_b200.device import run_cuda

pytestmark = pytest.mark.gpu


def compare(dev, scene, check_depth=True):
    oc, od, ost = scenes.run_oracle(scene)
    gc, gd, gst = run_cuda(dev, scene)
    assert gst.primitives == ost.primitives
    assert gst.fragmentsCovered == ost.fragmentsCovered, "coverage differs"
    assert gst.fragmentsWritten == ost.fragmentsWritten, "depth/stencil pass count differs"
    texel = scenes.TEXEL_SIZE[scene.color.format]
    if not np.array_equal(oc, gc):
        a = oc.reshape(scene.color.height, scene.color.width, texel)
        b = gc.reshape(scene.color.height, scene.color.width, texel)
        bad = np.argwhere((a != b).any(axis=2))
        y, x = bad[0]
        raise AssertionError("colour differs at %d pixels, first (x=%d,y=%d): oracle %s gpu %s" % (len(bad), x, y, a[y, x], b[y, x]))
    if check_depth and scene.depth is not None:
        assert np.array_equal(od, gd), "depth attachment differs"
    return gst  # primitives / fragment counts are the oracle's (asserted above); binEntries exists on the GPU side only


def test_draw_cube(dev):
    st = compare(dev, scenes.draw_cube())
    assert st.primitives == 12 and st.fragmentsCovered > 10000


@pytest.mark.parametrize("filt", [scenes.NEAREST, scenes.LINEAR])
def test_draw_textured_cube(dev, filt):
    compare(dev, scenes.draw_textured_cube(filt=filt))


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("cull,front", [(scenes.CULL_NONE, scenes.FRONT_CCW), (scenes.CULL_BACK, scenes.FRONT_CW), (scenes.CULL_FRONT, scenes.FRONT_CCW)])
def test_random_triangles(dev, seed, cull, front):
    compare(dev, scenes.random_triangles(tris=300, seed=seed, cull=cull, front_face=front))


def test_random_triangles_snapped_shared_edges(dev):
    # vertices exactly on pixel centres: edge-on-centre, double hits on shared edges, zero-area triangles
    compare(dev, scenes.random_triangles(width=64, height=48, tris=400, seed=11, snap=True, perspective=False))


@pytest.mark.parametrize("depth_fmt", [scenes.D16_UNORM, scenes.D32_SFLOAT, scenes.D24_UNORM_S8_UINT, None])
def test_depth_formats(dev, depth_fmt):
    compare(dev, scenes.random_triangles(tris=200, seed=5, depth_fmt=depth_fmt))


@pytest.mark.parametrize("op", [scenes.LESS, scenes.GREATER, scenes.EQUAL, scenes.ALWAYS, scenes.NOT_EQUAL, scenes.NEVER])
def test_depth_ops(dev, op):
    compare(dev, scenes.random_triangles(tris=150, seed=6, depth_op=op))


@pytest.mark.parametrize("topology", [scenes.TRIANGLE_STRIP, scenes.TRIANGLE_FAN])
def test_strip_and_fan(dev, topology):
    compare(dev, scenes.random_triangles(tris=60, seed=7, topology=topology))


@pytest.mark.parametrize("stride", [1, 2, 4])
def test_indexed(dev, stride):
    compare(dev, scenes.random_triangles(tris=80 if stride > 1 else 60, seed=8, indexed=stride))


@pytest.mark.parametrize("size", [(33, 17), (500, 500), (1, 1), (257, 3)])
def test_odd_sizes(dev, size):
    compare(dev, scenes.random_triangles(width=size[0], height=size[1], tris=50, seed=9))


def test_mesh_small(dev):
    compare(dev, scenes.mesh_indexed(width=640, height=360, nx=160, ny=90))


def test_mesh_layers(dev):
    compare(dev, scenes.mesh_indexed(width=320, height=200, nx=40, ny=25, layers=4))


@pytest.mark.parametrize("fmt", [scenes.R16G16B16A16_SFLOAT, scenes.R8G8B8A8_UNORM])
def test_overdraw_blend(dev, fmt):
    compare(dev, scenes.overdraw_quads(width=96, height=64, quads=12, tex_size=32, color_fmt=fmt))


def test_overdraw_opaque(dev):
    compare(dev, scenes.overdraw_quads(width=96, height=64, quads=3, tex_size=32, blend=False))


def test_long_tile_lists_multi_chunk(dev):
    # > 256 primitives over one tile: k_bin_sort path + several staged chunks per tile in k_raster
    compare(dev, scenes.random_triangles(width=64, height=64, tris=1500, seed=12))


def test_very_long_tile_lists_sort_fallback(dev):
    # > 16384 primitives over one tile: the in-HBM stable split-sort fallback of k_bin_sort
    compare(dev, scenes.random_triangles(width=40, height=40, tris=17000, seed=13, perspective=False))


@pytest.mark.parametrize("make", [lambda: scenes.draw_cube(), lambda: scenes.random_triangles(tris=300, seed=21),
                                  lambda: scenes.overdraw_quads(96, 64, 6, 16)], ids=["cube", "triangles", "blended"])
def test_immediate_clears_take_the_tile_load_path(dev, make):
    """With deferred clears off, k_clear runs first and k_raster reads its tiles from HBM (the path every draw after
    the first one of a render pass takes)."""
    dev.set_lazy_clear(False)
    try:
        compare(dev, make())
    finally:
        dev.set_lazy_clear(True)


def test_two_draws_into_one_target(dev):
    """Second draw without a clear in between: tiles come back from HBM with the first draw's colour and depth."""
    from cpvulkan_b200.device import SceneOnDevice
    scene = scenes.random_triangles(tris=150, seed=31)
    s = SceneOnDevice(dev, scene)
    try:
        s.render(); s.draw()
        gc, gd = s.read_color(), s.read_depth()
    finally:
        s.close()
    # oracle: same two draws
    import ctypes as C
    from cpvulkan_b200 import capi
    lib = capi.load_oracle(); mem = scenes.HostMemory(); m = scenes.materialize(scene, mem.alloc)
    for img, att in ((scene.color, m.color_attachment), (scene.depth, m.depth_attachment)):
        if img is not None and img.clear is not None:
            cv, is_ds = scenes.clear_value(img); assert lib.cpvk_oracle_clear(C.byref(att), C.byref(cv), is_ds) == 0
    st = capi.DrawStats()
    for _ in range(2):
        assert lib.cpvk_oracle_draw(C.byref(m.desc), C.byref(m.state), C.byref(st)) == 0
    assert np.array_equal(gc, mem.arrays["color"][:scene.color.nbytes])
    if scene.depth is not None:
        assert np.array_equal(gd, mem.arrays["depth"][:scene.depth.nbytes])


@pytest.mark.parametrize("band", [(0, 40), (37, 70), (64, 96)])
def test_band_equals_oracle_window(dev, band):
    """Sort-first band: rows [y0, y1) of the target are rendered, the rest keeps the clear value."""
    scene = scenes.random_triangles(width=96, height=96, tris=300, seed=41)
    oc, od, ost = scenes.run_oracle(scene, window=(0, band[0], 96, band[1]))
    gc, gd, gst = run_cuda(dev, scene, band=band)
    assert gst.fragmentsCovered == ost.fragmentsCovered and gst.fragmentsWritten == ost.fragmentsWritten
    assert np.array_equal(oc, gc) and np.array_equal(od, gd)


def _indexed_variant(kind):
    s = scenes.random_triangles(tris=120, seed=17, indexed=4)
    ib = s.buffers["ib"].view(np.uint32).copy()
    vb = s.buffers["vb"].view(np.float32).reshape(-1, 8).copy()
    if kind == "reuse":  # 120 triangles over 40 distinct vertices
        ib = (ib % 40).astype(np.uint32)
    elif kind == "sparse":  # index range 7x the index count: the vertex stage falls back to one invocation per index
        big = np.zeros((len(vb) * 7, 8), dtype=np.float32); big[::7] = vb; vb = big
        ib = (ib * 7).astype(np.uint32)
    elif kind == "offset":  # firstIndex and a negative vertexOffset on top of reuse
        ib = np.concatenate([np.array([9, 9, 9, 9, 9], dtype=np.uint32), (ib % 50) + 3]).astype(np.uint32)
        s.first, s.vertex_offset = 5, -3
    elif kind == "strip":
        s = scenes.random_triangles(tris=90, seed=18, topology=scenes.TRIANGLE_STRIP, indexed=2)
        ib16 = s.buffers["ib"].view(np.uint16).copy()
        s.buffers["ib"] = (ib16 % 30).astype(np.uint16).view(np.uint8).reshape(-1)
        return s
    s.buffers["ib"] = ib.view(np.uint8).reshape(-1)
    s.buffers["vb"] = vb.reshape(-1).view(np.uint8)
    return s


@pytest.mark.parametrize("kind", ["reuse", "sparse", "offset", "strip"])
def test_indexed_vertex_reuse(dev, kind):
    """Indexed draws shade each vertex of a compact index range once (the reference shades every index); the records
    are the same bits either way, and a sparse range falls back to per-index shading."""
    compare(dev, _indexed_variant(kind))


def test_speculative_plan_replays(built):
    """Draw tails are enqueued on the previous draw's binning answers; a wrong guess must be replayed invisibly.
    A fresh device starts with no sort and a small list capacity, so this sequence hits every kind of mismatch:
    lists longer than a raster chunk, more list entries than the default capacity, large primitives after none,
    and back to small ones."""
    from cpvulkan_b200.device import Device
    d = Device(0, stats=True)
    try:
        seq = [scenes.mesh_indexed(width=320, height=200, nx=40, ny=25),                  # small primitives only
               scenes.random_triangles(width=64, height=64, tris=900, seed=51),            # lists > 256: needs k_bin_sort
               scenes.overdraw_quads(256, 256, 40, 16),                                     # 80 full-screen triangles x 64 tiles > default capacity, large primitives
               scenes.mesh_indexed(width=320, height=200, nx=40, ny=25),
               scenes.random_triangles(width=64, height=64, tris=900, seed=52)]
        for sc in seq + seq[::-1]:
            compare(d, sc)
        d.set_speculation(False)
        compare(d, seq[1])
    finally:
        d.close()


@pytest.mark.parametrize("quads,size", [(255, 32), (256, 32), (257, 32), (256, 64), (300, 96)])
def test_single_pass_binning_slot_boundary(built, quads, size):
    """Single-pass binning gives every tile CPVK_ORDER_MAX = 512 list slots (fixed_kernels.h, CpvkSetupArgs::directLists):
    a tile with exactly 512 primitives fits, one with 514 overflows and the draw is replayed through count -> scan -> fill
    -> sort. Either way the bytes are the oracle's, the (primitive, tile) pairs are all there, and a draw that fits
    again afterwards goes back to the short path."""
    from cpvulkan_b200.device import Device
    d = Device(0, stats=True)
    try:
        tiles = ((size + 31) // 32) ** 2
        for sc in (scenes.overdraw_quads(size, size, quads, 16), scenes.mesh_indexed(width=160, height=96, nx=20, ny=12),
                   scenes.overdraw_quads(size, size, quads, 16)):
            st = compare(d, sc, check_depth=sc.depth is not None)
            if sc.name.startswith("overdraw"):
                assert st.binEntries == 2 * quads * tiles
    finally:
        d.close()


def test_texel_buffer_sample(dev):
    """BASELINE C5, third item (Samples/texel_buffer): texelFetch from a uniform texel buffer in the vertex shader, a
    private array indexed by gl_VertexIndex % 3, no vertex buffers."""
    st = compare(dev, scenes.texel_buffer(500, 500), check_depth=False)
    assert st.primitives == 1 and st.fragmentsCovered > 50000
    compare(dev, scenes.texel_buffer(96, 64, texels=(0.25, 0.5, 0.75), triangles=3), check_depth=False)


def test_texel_buffer_8k_property(dev):
    """The same draw at 7680x4320 (too large for the oracle): every pixel is either the clear colour or the fetched
    colour, the covered count equals the device's own statistic, and the triangle is left-right symmetric."""
    scene = scenes.texel_buffer(7680, 4320)
    gc, _, st = run_cuda(dev, scene)
    img = gc.view(np.uint32).reshape(4320, 7680)
    clear, tri = np.uint32(0x33333333), np.uint32(0xFFFF00FF)  # BGRA8: (0.2,0.2,0.2,0.2) and (1,0,1,1)
    inside = img == tri
    assert np.all(inside | (img == clear))
    assert int(inside.sum()) == st.fragmentsCovered == st.fragmentsWritten
    assert abs(int(inside[:, :3840].sum()) - int(inside[:, 3840:].sum())) <= 4320


def test_mirror_stores_copy_the_band(dev):
    """CpvkDrawState.mirrorColor0 (the multi-GPU gather fused into k_raster): every tile row of the band is also stored
    at the same offset of each mirror allocation — here two more buffers on the same GPU stand in for the peers."""
    from cpvulkan_b200.device import SceneOnDevice
    scene = scenes.random_triangles(width=96, height=96, tris=200, seed=61)
    band = (37, 70)
    s = SceneOnDevice(dev, scene, band)
    mirrors = [dev.alloc(scene.color.nbytes) for _ in range(2)]
    try:
        for m in mirrors:
            dev.upload(m, np.full(scene.color.nbytes, 0xAB, dtype=np.uint8))
        s.m.state.mirrorCount = 2
        for i, m in enumerate(mirrors):
            s.m.state.mirrorColor0[i] = m
        s.render()
        own = s.read_color().reshape(96, -1)
        oc, _, _ = scenes.run_oracle(scene, window=(0, band[0], 96, band[1]))
        assert np.array_equal(own.reshape(-1), oc)
        for m in mirrors:
            got = dev.download(m, scene.color.nbytes).reshape(96, -1)
            assert np.array_equal(got[band[0]:band[1]], own[band[0]:band[1]]), "band rows must arrive in the mirror"
            assert np.all(got[:band[0]] == 0xAB) and np.all(got[band[1]:] == 0xAB), "rows outside the band belong to other GPUs"
    finally:
        for m in mirrors:
            dev.free(m)
        s.close()


# ---- fixed-function state outside C1-C5 (§8(f) f4): same bar, bit-exact ----

def _with(scene, fn):
    scene.mutate = fn
    return scene


STENCIL_OPS = [  # (fail, pass, depthFail, compare, compareMask, writeMask, reference)
    (0, 2, 0, 7, 0xFF, 0xFF, 0x40),   # ALWAYS / REPLACE
    (0, 3, 4, 1, 0xFF, 0xFF, 0x01),   # LESS, INCREMENT_AND_CLAMP on pass, DECREMENT_AND_CLAMP on depth fail (signed-saturate quirk)
    (5, 6, 7, 3, 0x0F, 0xF0, 0x05),   # LESS_OR_EQUAL with masks, INVERT / INC_WRAP / DEC_WRAP
    (1, 0, 2, 5, 0xFF, 0x3C, 0x80),   # NOT_EQUAL, ZERO on fail, REPLACE on depth fail, partial write mask
    (2, 2, 2, 0, 0xFF, 0xFF, 0x7F),   # NEVER: only the fail op runs
]


@pytest.mark.parametrize("ops", STENCIL_OPS)
@pytest.mark.parametrize("fmt", [scenes.D24_UNORM_S8_UINT, 130, 128])
def test_stencil(dev, ops, fmt):
    def edit(m):
        m.desc.stencilTestEnable = 1
        for st in (m.desc.front, m.desc.back):
            st.failOp, st.passOp, st.depthFailOp, st.compareOp, st.compareMask, st.writeMask, st.reference = ops
        m.desc.back.reference = (ops[6] + 3) & 0xFF
        m.desc.back.compareOp = (ops[3] + 1) % 8
    sc = scenes.random_triangles(width=96, height=64, tris=250, seed=71, depth_fmt=fmt)
    sc.depth.clear = ("depth", (1.0, 0x10))
    compare(dev, _with(sc, edit))


def test_stencil_only_attachment(dev):
    def edit(m):
        m.desc.stencilTestEnable = 1
        m.desc.depthTestEnable = m.desc.depthWriteEnable = 0
        for st in (m.desc.front, m.desc.back):
            st.failOp, st.passOp, st.depthFailOp, st.compareOp, st.compareMask, st.writeMask, st.reference = (0, 6, 0, 4, 0xFF, 0xFF, 2)
    sc = scenes.random_triangles(width=64, height=64, tris=200, seed=72, depth_fmt=127)
    sc.depth.clear = ("depth", (0.0, 0))
    compare(dev, _with(sc, edit))


@pytest.mark.parametrize("bounds", [(0.2, 0.7), (0.0, 0.0), (0.9, 0.1)])
def test_depth_bounds(dev, bounds):
    def edit(m):
        m.desc.depthBoundsTestEnable = 1
        m.desc.minDepthBounds, m.desc.maxDepthBounds = bounds
    sc = scenes.random_triangles(width=96, height=64, tris=250, seed=73)
    sc.depth.clear = ("depth", (0.5, 0))
    compare(dev, _with(sc, edit))


@pytest.mark.parametrize("depth_range", [(0.0, 1.0), (0.25, 0.75), (1.0, 0.0)])
def test_viewport_depth_range(dev, depth_range):
    sc = scenes.random_triangles(width=80, height=60, tris=200, seed=74)
    sc.viewport = (0.0, 0.0, 80.0, 60.0, depth_range[0], depth_range[1])
    compare(dev, sc)


def test_viewport_larger_than_the_attachment(dev):
    sc = scenes.random_triangles(width=80, height=60, tris=200, seed=75)
    sc.viewport = (0.0, 0.0, 128.0, 100.0, 0.0, 1.0)
    # The reference has no clamp: fragments beyond 80x60 are written out of bounds (undefined behaviour).
    # The CUDA path drops them, which is the oracle restricted to the attachment's window.
    oc, od, ost = scenes.run_oracle(sc, window=(0, 0, 80, 60))
    gc, gd, gst = run_cuda(dev, sc)
    assert gst.fragmentsCovered == ost.fragmentsCovered and gst.fragmentsWritten == ost.fragmentsWritten
    assert np.array_equal(oc, gc) and np.array_equal(od, gd)


BLENDS = [dict(src=1, dst=1, op=0), dict(src=6, dst=7, op=1), dict(src=4, dst=2, op=2, srcA=1, dstA=0, opA=0), dict(src=14, dst=1, op=0),
          dict(src=10, dst=11, op=0, srcA=12, dstA=13, opA=0), dict(src=1, dst=1, op=3), dict(src=1, dst=1, op=4, srcA=6, dstA=7, opA=0),
          dict(src=8, dst=9, op=0), dict(src=3, dst=5, op=0)]


@pytest.mark.parametrize("blend", BLENDS, ids=lambda b: "-".join(str(v) for v in b.values()))
@pytest.mark.parametrize("fmt", [scenes.R8G8B8A8_UNORM, scenes.R16G16B16A16_SFLOAT, scenes.R32G32B32A32_SFLOAT])
def test_blend_factors_and_ops(dev, blend, fmt):
    def edit(m):
        for i, c in enumerate((0.25, 0.5, 0.75, 0.6)):
            m.desc.blendConstants[i] = c
    sc = scenes.random_triangles(width=64, height=48, tris=150, seed=76, color_fmt=fmt, depth_fmt=None)
    sc.blend = blend
    compare(dev, _with(sc, edit))


@pytest.mark.parametrize("mask", [0x1, 0x6, 0x8, 0x0])
@pytest.mark.parametrize("blend", [None, dict(src=6, dst=7, op=0)])
def test_colour_write_mask(dev, mask, blend):
    sc = scenes.random_triangles(width=64, height=48, tris=150, seed=77)
    sc.write_mask, sc.blend = mask, blend
    compare(dev, sc)


def test_instanced_draw_with_instance_rate_attribute(dev):
    """instanceCount > 1, firstInstance, and a per-instance colour fetched with VK_VERTEX_INPUT_RATE_INSTANCE."""
    sc = scenes.random_triangles(width=96, height=64, tris=40, seed=78)
    n_inst, first = 3, 2
    inst_colors = np.random.RandomState(5).random_sample((n_inst + first, 4)).astype(np.float32)
    sc.buffers["inst"] = inst_colors.view(np.uint8).reshape(-1)
    sc.bindings = [(0, 32, 0), (1, 16, 1)]
    sc.attributes = [(0, 0, scenes.R32G32B32A32_SFLOAT, 0), (1, 1, scenes.R32G32B32A32_SFLOAT, 0)]
    sc.vertex_buffers = {0: "vb", 1: "inst"}
    sc.instances, sc.first_instance = n_inst, first
    sc.depth_op = scenes.ALWAYS  # every instance overwrites the previous one: the last instance's colour must win
    st = compare(dev, sc)
    assert st.primitives == 40 * n_inst


@pytest.mark.parametrize("iterations,threshold,scale", [(4, 0.9, None), (0, 0.5, None), (9, 100.0, 1.25), (3, -1.0, 0.0)])
def test_shader_front_end_breadth(dev, iterations, threshold, scale):
    """a fragment shader with a phi-based loop with break, a called function, OpKill, a push-constant
    block, a specialisation constant, bit operations and GLSL.std.450 calls — translator (CUDA) vs interpreter (oracle)."""
    import struct

    def edit(m):
        if scale is not None:
            m.desc.fragment.specCount = 1
            m.desc.fragment.spec[0].constantId = 3
            m.desc.fragment.spec[0].value = struct.unpack("<I", struct.pack("<f", scale))[0]
    sc = scenes.random_triangles(width=64, height=48, tris=120, seed=81)
    sc.fs = "complex.frag"
    sc.push_constants = struct.pack("<4fif", 0.3, 0.1, 0.2, 0.05, iterations, threshold)
    st = compare(dev, _with(sc, edit))
    assert st.fragmentsCovered > 1000


# ---- points and lines (ProcessPoints / ProcessLines, Draw.cpp:1315-1508) ----

@pytest.mark.parametrize("seed", [1, 2])
@pytest.mark.parametrize("perspective", [True, False])
def test_points(dev, seed, perspective):
    st = compare(dev, scenes.random_points_lines(topology=scenes.POINT_LIST, count=150, seed=seed, perspective=perspective))
    assert st.primitives == 150 and st.fragmentsCovered > 1000


@pytest.mark.parametrize("topology", [scenes.LINE_LIST, scenes.LINE_STRIP])
@pytest.mark.parametrize("width", [1.0, 3.0, 7.5])
def test_lines(dev, topology, width):
    st = compare(dev, scenes.random_points_lines(topology=topology, count=40, seed=3, line_width=width))
    assert st.fragmentsCovered > 500


def test_lines_unit_w_and_blend(dev):
    sc = scenes.random_points_lines(topology=scenes.LINE_STRIP, count=50, seed=4, line_width=4.0, perspective=False, depth_fmt=None,
                                    color_fmt=scenes.R16G16B16A16_SFLOAT)
    sc.blend = dict(src=6, dst=7, op=0)
    compare(dev, sc)


def test_points_many_tiles_and_chunks(dev):
    # more than one 256-primitive chunk, several tiles, indexed with vertex reuse
    sc = scenes.random_points_lines(width=200, height=150, topology=scenes.POINT_LIST, count=700, seed=5)
    sc.buffers["ib"] = (np.random.RandomState(6).permutation(700) % 350).astype(np.uint16).view(np.uint8).reshape(-1)
    sc.index_buffer, sc.index_stride = "ib", 2
    compare(dev, sc)


def test_degenerate_line_covers_like_the_reference(dev):
    """p0 == p1 in x and y: the four quad edges collapse, every edge function is 0 >= 0 and the reference shades the whole
    viewport (with t = NaN). The conservative pixel box must not clip that."""
    sc = scenes.random_points_lines(width=40, height=30, topology=scenes.LINE_LIST, count=2, seed=7, perspective=False)
    vb = sc.buffers["vb"].view(np.float32).reshape(-1, 9).copy()
    vb[1, 0:2] = vb[0, 0:2]  # same x, y; different z
    sc.buffers["vb"] = vb.reshape(-1).view(np.uint8)
    compare(dev, sc)


# Sampler state matrix (ImageSampler.cpp:12-38, :461-673): the oracle's sampler is pinned bit for bit to the reference's own
# compiled ImageSampler.cpp (tests/test_reference_sampler.py); here the CUDA sampler is held to the oracle over the same state
# space — every address mode (mixed per axis), both filters on the magnification and minification paths, both mipmap modes
# with fractional LODs over a 3-level chain, border colours — for an 8-bit UNORM and a raw float texture.
SAMPLER_STATES = [dict(address=(a, a), mag=f, min_=f) for a in range(5) for f in (scenes.NEAREST, scenes.LINEAR)]
SAMPLER_STATES += [dict(address=(3, 3), mag=scenes.LINEAR, min_=scenes.LINEAR, border=b) for b in (2, 4)]
SAMPLER_STATES += [dict(address=(0, 2), mag=scenes.LINEAR, min_=scenes.NEAREST), dict(address=(1, 3), mag=scenes.NEAREST, min_=scenes.LINEAR, border=4),
                   dict(address=(4, 0), mag=scenes.LINEAR, min_=scenes.LINEAR)]
SAMPLER_STATES += [dict(address=(a, a), mag=1 - f, min_=f, mipmap=m, min_lod=lod)
                   for a, f, m, lod in ((0, 1, 0, 0.25), (0, 1, 0, 0.5), (2, 1, 0, 1.5), (1, 0, 0, 2.0), (0, 1, 1, 0.25), (2, 1, 1, 1.0),
                                        (1, 1, 1, 1.5), (0, 0, 1, 0.75), (3, 1, 1, 3.7), (4, 1, 1, 1.25))]


@pytest.mark.parametrize("state", SAMPLER_STATES, ids=lambda s: "-".join("%s" % (v,) for v in s.values()).replace(" ", ""))
@pytest.mark.parametrize("tex_fmt", [scenes.R8G8B8A8_UNORM, scenes.R32G32B32A32_SFLOAT])
def test_sampler_state_matrix(dev, state, tex_fmt):
    compare(dev, scenes.sampler_matrix(tex_fmt=tex_fmt, **state))


# Whole-region rejection of large triangles (k_raster): must never drop a fragment the reference's per-pixel test accepts.
@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("scale", ["mixed", "extreme"])
def test_large_triangle_region_rejection_is_exact(dev, seed, scale):
    compare(dev, scenes.large_triangles(seed=seed, scale=scale))


def test_large_triangles_many_tiles(dev):
    compare(dev, scenes.large_triangles(width=517, height=389, tris=40, seed=7, scale="extreme"))


@pytest.mark.parametrize("filt", [scenes.NEAREST, scenes.LINEAR])
def test_separate_image_and_sampler(dev, filt):
    # OpSampledImage: image data from one descriptor, sampler state from another (Samples/separate_image_sampler)
    compare(dev, scenes.separate_image_sampler(300, 220, filt))


def test_input_attachment_read(dev):
    compare(dev, scenes.input_attachment(300, 220))


# ---- empty and ragged inputs (CalculatePrimitives, Draw.cpp:567-673: count / 3 triangles, the remainder is dropped) ----

@pytest.mark.parametrize("count", [0, 1, 2, 3, 4, 5, 7])
@pytest.mark.parametrize("topology", [scenes.TRIANGLE_LIST, scenes.TRIANGLE_STRIP, scenes.TRIANGLE_FAN])
def test_empty_and_ragged_vertex_counts(dev, count, topology):
    sc = scenes.random_triangles(width=64, height=48, tris=8, seed=91, topology=topology)  # at least 10 vertices in the buffer
    sc.count = count
    st = compare(dev, sc)
    assert st.primitives == (count // 3 if topology == scenes.TRIANGLE_LIST else max(count - 2, 0))


@pytest.mark.parametrize("count", [0, 2, 4, 11])
@pytest.mark.parametrize("stride", [1, 2, 4])
def test_ragged_index_counts(dev, count, stride):
    sc = scenes.random_triangles(width=64, height=48, tris=6, seed=92, indexed=stride)
    sc.count = count
    compare(dev, sc)


def test_zero_instances_draw_nothing(dev):
    sc = scenes.random_triangles(width=64, height=48, tris=20, seed=93)
    sc.instances = 0
    st = compare(dev, sc)
    assert st.fragmentsCovered == 0


def test_everything_culled_or_off_screen(dev):
    sc = scenes.random_triangles(width=64, height=48, tris=50, seed=94, cull=3)  # FRONT_AND_BACK
    assert compare(dev, sc).fragmentsCovered == 0
    sc = scenes.random_triangles(width=64, height=48, tris=50, seed=95)
    vb = sc.buffers["vb"].view(np.float32).reshape(-1, 8).copy()
    vb[:, 0] += 5.0  # all of it to the right of the viewport (w == 1 .. 3: still outside after the divide)
    vb[:, 0] *= 4.0
    sc.buffers["vb"] = vb.view(np.uint8).reshape(-1)
    assert compare(dev, sc).fragmentsCovered == 0


def test_back_to_back_empty_and_full_draws_share_the_device(dev):
    # an empty draw must leave the device's speculative launch plan and scratch in a usable state for the next draw
    empty = scenes.random_triangles(width=64, height=48, tris=4, seed=96); empty.count = 0
    compare(dev, empty)
    compare(dev, scenes.random_triangles(width=200, height=150, tris=400, seed=97))
    compare(dev, empty)
    compare(dev, scenes.draw_cube(128, 128))


# ---- maximum sizes ----

def test_widest_and_tallest_render_targets(dev):
    """The bbox records hold 16-bit pixel coordinates: 32767 is the largest extent the path takes (larger is refused, below)."""
    compare(dev, scenes.random_triangles(width=32767, height=3, tris=6, seed=98, depth_fmt=scenes.D16_UNORM))
    compare(dev, scenes.random_triangles(width=5, height=32767, tris=6, seed=99, depth_fmt=None))


def test_render_targets_beyond_the_limit_are_refused(dev):
    from cpvulkan_b200.device import CpvkError
    with pytest.raises(CpvkError) as e:
        run_cuda(dev, scenes.random_triangles(width=32768, height=2, tris=2, seed=1, depth_fmt=None))
    assert "32767" in str(e.value)
    compare(dev, scenes.draw_cube(64, 64))  # the device is still usable


# LOD bias / clamp and the view swizzle of the sampling wrapper (GlslFunctions.cpp:539-555, :598-654); the oracle side of these is
# pinned by tests/test_oracle_kats.py::test_lod_bias_clamp_and_mip_choice_against_numpy and ::test_view_swizzle_against_numpy.
@pytest.mark.parametrize("state", [dict(bias=0.6, mipmap=1), dict(bias=-0.75, min_lod=1.5, mipmap=1), dict(bias=50.0, mipmap=0), dict(min_lod=0.25, max_lod=0.25, mipmap=1),
                                   dict(bias=1.0, min_lod=0.5, max_lod=1.25, mipmap=0), dict(swizzle=(6, 5, 4, 3)), dict(swizzle=(1, 2, 0, 3), bias=0.4, mipmap=1),
                                   dict(swizzle=(4, 4, 4, 2), address=(1, 3), border=4)],
                         ids=lambda s: "-".join("%s%s" % (k, v) for k, v in s.items()).replace(" ", ""))
@pytest.mark.parametrize("tex_fmt", [scenes.R8G8B8A8_UNORM, scenes.R32G32B32A32_SFLOAT])
def test_sampler_lod_bias_clamp_and_swizzle(dev, state, tex_fmt):
    compare(dev, scenes.sampler_matrix(tex_fmt=tex_fmt, **state))


# Integer colour attachments: the fragment output is uvec4 and the attachment write goes through @setPixelU32 with the format's
# clamp (PipelineCompiler.cpp:1528-1653, ImageCompiler.cpp:1130-1160); partial write masks read the destination back as integers.
@pytest.mark.parametrize("fmt", [41, 95, 107], ids=["R8G8B8A8_UINT", "R16G16B16A16_UINT", "R32G32B32A32_UINT"])
@pytest.mark.parametrize("mask", [0xF, 0x5])
def test_unsigned_integer_colour_attachment(dev, fmt, mask):
    sc = scenes.random_triangles(width=64, height=48, tris=40, seed=61, color_fmt=fmt)
    sc.fs = "uintout.frag"
    sc.color.clear = ("color_uint", (1, 2, 3, 4))
    sc.write_mask = mask
    compare(dev, sc)


@pytest.mark.parametrize("fmt", [42, 96, 108], ids=["R8G8B8A8_SINT", "R16G16B16A16_SINT", "R32G32B32A32_SINT"])
@pytest.mark.parametrize("mask", [0xF, 0xA])
def test_signed_integer_colour_attachment(dev, fmt, mask):
    sc = scenes.random_triangles(width=64, height=48, tris=40, seed=62, color_fmt=fmt)
    sc.fs = "sintout.frag"                                        # ivec4(color * 600 - 300): clamps at both ends of an 8-bit target
    sc.color.clear = ("color_uint", (0xFFFFFFFF, 2, 3, 4))        # -1, 2, 3, 4 as VkClearColorValue.int32
    sc.write_mask = mask
    compare(dev, sc)


# ---- multiple colour attachments (a10: the output with Location == attachment index feeds that attachment) ----

def two_attachments(fmt1, blend1=None, mask1=0xF, seed=63):
    """random triangles into two colour attachments: #0 RGBA8 (the scene's own), #1 `fmt1` placed right behind it in the same
    allocation, so both come back with the colour read-back."""
    sc = scenes.random_triangles(width=72, height=40, tris=60, seed=seed)
    sc.fs = "mrt.frag"
    w, h = sc.color.width, sc.color.height
    texel1 = scenes.TEXEL_SIZE[fmt1]
    sc.color.chain_bytes = w * h * 4 + w * h * texel1
    sc.color.data = np.zeros(sc.color.chain_bytes, dtype=np.uint8)
    sc.color.data[w * h * 4:] = 0x3C  # attachment 1 is not cleared by the scene: give it a known, non-trivial start

    def edit(m):
        a0 = m.color_attachment
        m.desc.colorAttachmentCount = 2
        m.desc.colorFormats[1] = fmt1
        b = m.desc.blend[1]
        b.colorWriteMask = mask1
        if blend1:
            b.blendEnable = 1
            b.srcColorBlendFactor, b.dstColorBlendFactor, b.colorBlendOp = blend1["src"], blend1["dst"], blend1["op"]
            b.srcAlphaBlendFactor, b.dstAlphaBlendFactor, b.alphaBlendOp = blend1["src"], blend1["dst"], blend1["op"]
        m.state.color[1] = type(a0)(a0.address + a0.rowPitch * a0.height, w, h, w * texel1, fmt1)
    sc.mutate = edit
    return sc


@pytest.mark.parametrize("fmt1,blend1,mask1", [(scenes.B8G8R8A8_UNORM, None, 0xF), (scenes.R16G16B16A16_SFLOAT, None, 0xF),
                                               (scenes.R8G8B8A8_UNORM, dict(src=6, dst=7, op=0), 0xF), (scenes.R32G32B32A32_SFLOAT, None, 0x6)],
                         ids=["bgra8", "rgba16f", "rgba8-blended", "rgba32f-masked"])
def test_two_colour_attachments(dev, fmt1, blend1, mask1):
    compare(dev, two_attachments(fmt1, blend1, mask1))


# ---- interpolation qualifiers and gl_FragCoord (a6, a7) ----

@pytest.mark.parametrize("topology", [scenes.TRIANGLE_LIST, scenes.TRIANGLE_STRIP, scenes.TRIANGLE_FAN])
def test_flat_inputs_copy_the_provoking_vertex(dev, topology):
    sc = scenes.random_triangles(width=64, height=48, tris=30, seed=64, color_fmt=scenes.R32G32B32A32_SFLOAT, topology=topology)
    sc.fs = "flat.frag"  # Draw.cpp:940-943; provoking vertex per CalculatePrimitives (:614-661: list 3i, strip i, fan i + 1)
    compare(dev, sc)


@pytest.mark.parametrize("perspective", [True, False])
def test_noperspective_inputs(dev, perspective):
    sc = scenes.random_triangles(width=64, height=48, tris=30, seed=65, color_fmt=scenes.R32G32B32A32_SFLOAT, perspective=perspective)
    sc.fs = "nopersp.frag"  # SetDatum<false, 3>: w0*v0 + w1*v1 + w2*v2 (Draw.cpp:833-839)
    compare(dev, sc)


@pytest.mark.parametrize("depth_range", [(0.0, 1.0), (0.2, 0.6)])
def test_frag_coord_is_the_integer_pixel_and_the_viewport_depth(dev, depth_range):
    sc = scenes.random_triangles(width=64, height=48, tris=30, seed=66, color_fmt=scenes.R32G32B32A32_SFLOAT)
    sc.fs = "fragcoord.frag"  # fragCoord = (x, y, depth', 1) with integer x, y (Draw.cpp:1300-1313, :1574-1579)
    sc.viewport = (0.0, 0.0, 64.0, 48.0, depth_range[0], depth_range[1])
    compare(dev, sc)


def test_flat_and_noperspective_on_lines(dev):
    for fs in ("flat.frag", "nopersp.frag"):
        sc = scenes.random_points_lines(count=24, seed=8, topology=scenes.LINE_STRIP, line_width=3.0, color_fmt=scenes.R32G32B32A32_SFLOAT, fs=fs)
        compare(dev, sc)


# ---- depth-only passes: no colour attachment in the subpass (a shadow-map style draw) ----

@pytest.mark.parametrize("depth_fmt", [scenes.D32_SFLOAT, scenes.D16_UNORM, scenes.D24_UNORM_S8_UINT])
def test_depth_only_pass(dev, depth_fmt):
    from cpvulkan_b200 import capi

    def edit(m):
        m.desc.colorAttachmentCount = 0
        m.desc.colorFormats[0] = 0
        m.state.color[0] = capi.Attachment()
    sc = scenes.random_triangles(width=64, height=48, tris=40, seed=67, depth_fmt=depth_fmt)
    st = compare(dev, _with(sc, edit))  # the colour image keeps its clear value on both sides, the depth attachment is compared
    assert 0 < st.fragmentsWritten < st.fragmentsCovered


# ---- vertex fetch conversions (a3, EmitCopyInput, PipelineCompiler.cpp:821-896) and a non-zero first vertex ----

@pytest.mark.parametrize("fmt,texel", [(scenes.R8G8B8A8_UNORM, 4), (scenes.B8G8R8A8_UNORM, 4), (scenes.R16G16B16A16_SFLOAT, 8), (64, 4), (38, 4)],
                         ids=["rgba8_unorm", "bgra8_unorm", "rgba16f", "a2b10g10r10_unorm", "rgba8_snorm"])
def test_vertex_attribute_formats(dev, fmt, texel):
    """The colour attribute arrives in a format other than the shader's vec4: half floats are extended (the "simple format"
    path), packed / normalised formats go through the pixel decoder (EmitGetPixel)."""
    sc = scenes.random_triangles(width=64, height=48, tris=30, seed=68, color_fmt=scenes.R32G32B32A32_SFLOAT)
    vb = sc.buffers["vb"].view(np.float32).reshape(-1, 8)
    n = len(vb)
    rng = np.random.RandomState(3)
    if fmt == scenes.R16G16B16A16_SFLOAT:
        packed = rng.uniform(-2, 2, size=(n, 4)).astype(np.float16).view(np.uint8).reshape(n, 8)
    else:
        packed = rng.randint(0, 256, size=(n, texel), dtype=np.uint8)
    stride = 16 + texel
    out = np.zeros((n, stride), dtype=np.uint8)
    out[:, :16] = np.ascontiguousarray(vb[:, :4]).view(np.uint8).reshape(n, 16)
    out[:, 16:] = packed
    sc.buffers["vb"] = out.reshape(-1)
    sc.bindings = [(0, stride, 0)]
    sc.attributes = [(0, 0, scenes.R32G32B32A32_SFLOAT, 0), (1, 0, fmt, 16)]
    compare(dev, sc)


def test_first_vertex_of_a_non_indexed_draw(dev):
    sc = scenes.random_triangles(width=64, height=48, tris=30, seed=69)
    sc.first, sc.count = 21, 45  # vertexId = firstVertex + i (Draw.cpp:675-688)
    st = compare(dev, sc)
    assert st.primitives == 15


def test_glsl_std_450_subset(dev):
    """The reference's whole GLSL.std.450 subset + OpDot + integer conversions in one shader (glslmath.frag): translator vs
    interpreter, bit for bit; the interpreter side is pinned by tests/test_oracle_kats.py::test_glsl_std_450_subset_against_numpy."""
    for seed, persp in ((70, True), (71, False)):
        sc = scenes.random_triangles(width=48, height=36, tris=20, seed=seed, color_fmt=scenes.R32G32B32A32_SFLOAT, depth_fmt=None, perspective=persp)
        sc.fs = "glslmath.frag"
        compare(dev, sc)


def test_matrix_products_in_a_vertex_shader(dev):
    """mat*mat, mat*scalar, mat*vec and vec*mat (matmath.vert) on a point list: translator vs interpreter; the interpreter's
    operand order is pinned against glm's by tests/test_oracle_kats.py::test_matrix_products_against_numpy."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("kats", os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_oracle_kats.py"))
    kats = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(kats)
    compare(dev, kats.matmath_scene()[0])
    compare(dev, kats.matmath_scene(64, 48, seed=13)[0])


@pytest.mark.parametrize("kw", [dict(), dict(first=5, count=20), dict(instances=3, first_instance=2), dict(indexed=True)],
                         ids=["plain", "first-vertex", "instances", "indexed"])
def test_vertex_and_instance_index_builtins(dev, kw):
    """gl_VertexIndex (firstVertex + i, or vertexOffset + index) and gl_InstanceIndex (firstInstance + instance) feed the colour and
    the position of a point list (builtins.vert)."""
    sc = scenes.random_points_lines(width=96, height=64, count=40, seed=14, topology=scenes.POINT_LIST, perspective=False)
    sc.vs = "builtins.vert"
    if "first" in kw:
        sc.first, sc.count = kw["first"], kw["count"]
    if "instances" in kw:
        sc.instances, sc.first_instance = kw["instances"], kw["first_instance"]
    if kw.get("indexed"):
        idx = np.random.RandomState(2).permutation(40).astype(np.uint16)
        sc.buffers["ib"] = idx.view(np.uint8).reshape(-1)
        sc.index_buffer, sc.index_stride, sc.vertex_offset = "ib", 2, 0
    compare(dev, sc)


def test_shared_reciprocal_division_is_ieee_division(dev):
    """cpvk_div_shared (edge weights / area, interpolants / denominator, position / w share one reciprocal per denominator) must
    return the bits of the `/` operator = IEEE-754 round-to-nearest-even division, which numpy computes on the host: random bit
    patterns over every exponent (subnormals, zeros of both signs, infinities, NaNs), operands at the edges of the fast path's
    exponent window, quotients that round to the subnormal / overflow boundaries, and mantissa patterns that are hard to round."""
    rng = np.random.default_rng(2026)
    n = 1 << 20
    a = rng.integers(0, 1 << 32, size=(n, 3), dtype=np.uint64).astype(np.uint32)
    b = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    # a second million with exponents packed around the fast-path window [2^-62, 2^63) and around 1
    e = rng.choice(np.array([1, 2, 63, 64, 65, 66, 100, 126, 127, 128, 150, 188, 189, 190, 191, 253, 254], dtype=np.uint32), size=(n, 4))
    m = rng.integers(0, 1 << 23, size=(n, 4), dtype=np.uint64).astype(np.uint32)
    m[rng.random((n, 4)) < 0.15] = 0x7FFFFF
    m[rng.random((n, 4)) < 0.15] = 0
    sgn = rng.integers(0, 2, size=(n, 4), dtype=np.uint64).astype(np.uint32) << 31
    packed = sgn | (e << 23) | m
    a = np.concatenate([a, packed[:, :3]]); b = np.concatenate([b, packed[:, 3]])
    special = np.array([0x00000000, 0x80000000, 0x7F800000, 0xFF800000, 0x7FC00000, 0x00000001, 0x807FFFFF, 0x00800000, 0x3F800000, 0xBF800000], dtype=np.uint32)
    sa = np.array([[x, y, z] for x in special for y in special[:3] for z in special[-3:]], dtype=np.uint32)
    a = np.concatenate([a, np.tile(sa, (len(special), 1))]); b = np.concatenate([b, np.repeat(special, len(sa))])
    total = len(b)
    with np.errstate(all="ignore"):
        want = (a.view(np.float32) / b.view(np.float32)[:, None]).astype(np.float32).view(np.uint32)
    da, db, ds, dp = dev.alloc(a.nbytes), dev.alloc(b.nbytes), dev.alloc(a.nbytes), dev.alloc(a.nbytes)
    dev.upload(da, a); dev.upload(db, b)
    from cpvulkan_b200.device import _check
    _check(dev.lib, dev.lib.cpvk_cuda_selftest_div(dev.handle, da, db, total, ds, dp))
    shared = dev.download(ds, a.nbytes).view(np.uint32).reshape(-1, 3)
    plain = dev.download(dp, a.nbytes).view(np.uint32).reshape(-1, 3)
    for x in (da, db, ds, dp):
        dev.free(x)

    def canon(v):
        v = v.copy(); v[(v & 0x7FFFFFFF) > 0x7F800000] = 0x7FC00000
        return v
    assert np.array_equal(canon(plain), canon(want)), "the GPU's `/` is not IEEE division?"
    bad = np.argwhere(canon(shared) != canon(want))
    assert len(bad) == 0, "cpvk_div_shared differs from IEEE division in %d quotients, first: %08x / %08x -> %08x, want %08x" % (
        len(bad), a[bad[0][0], bad[0][1]], b[bad[0][0]], shared[bad[0][0], bad[0][1]], want[bad[0][0], bad[0][1]])


def test_index_range_is_recomputed_when_the_indices_change(dev):
    """The lowest / highest index of an indexed draw is remembered across draws (an application redraws the same range every
    frame); rewriting the index buffer through the library must drop it. Same buffer address, same count, another range."""
    from cpvulkan_b200.device import SceneOnDevice
    a = scenes.mesh_indexed(width=320, height=200, nx=40, ny=25)
    s = SceneOnDevice(dev, a)
    try:
        for round_ in range(2):
            s.render()
            assert np.array_equal(s.read_color(), scenes.run_oracle(a)[0])
        # the same triangles drawn from the upper half of the vertex range only: indices clamped from below change lowest
        idx = a.buffers["ib"].view(np.uint32).copy()
        idx[:] = np.maximum(idx, 400)
        b = scenes.mesh_indexed(width=320, height=200, nx=40, ny=25)
        b.buffers["ib"] = idx.view(np.uint8).reshape(-1)
        dev.upload(s.m.addr["ib"], b.buffers["ib"])
        s.render()
        oc, od, _ = scenes.run_oracle(b)
        assert np.array_equal(s.read_color(), oc) and np.array_equal(s.read_depth(), od)
        # ... and back, through a device-side copy this time
        tmp = dev.alloc(a.buffers["ib"].nbytes)
        dev.upload(tmp, a.buffers["ib"])
        dev.copy_rows(s.m.addr["ib"], a.buffers["ib"].nbytes, tmp, a.buffers["ib"].nbytes, a.buffers["ib"].nbytes, 1)
        s.render()
        assert np.array_equal(s.read_color(), scenes.run_oracle(a)[0])
        dev.free(tmp)
    finally:
        s.close()


def test_unorm8_decode_all_codes(dev):
    """(float)k / 255.0f without the divide (cpvk_unorm8): all 256 codes through a NEAREST blit R8G8B8A8_UNORM -> R32G32B32A32_SFLOAT
    against the host's IEEE division."""
    import ctypes as C
    from cpvulkan_b200 import capi
    src = np.arange(256, dtype=np.uint8).repeat(4).reshape(256, 4).copy().reshape(-1)
    sa, da = dev.alloc(src.nbytes), dev.alloc(256 * 16)
    dev.upload(sa, src)
    dev.blit(capi.Blit(capi.Attachment(sa, 256, 1, 1024, 37), capi.Attachment(da, 256, 1, 4096, 109), 0, 0, 256, 1, 0, 0, 256, 1, 0))
    got = dev.download(da, 256 * 16).view(np.float32).reshape(256, 4)
    want = (np.arange(256, dtype=np.float32) / np.float32(255.0)).astype(np.float32)
    for c in range(4):
        assert np.array_equal(got[:, c].view(np.uint32), want.view(np.uint32)), "channel %d" % c
    dev.free(sa); dev.free(da)


def test_multiple_descriptor_sets_and_uniform_arrays(dev):
    """f3: resources spread over descriptor sets 0 and 1 (Samples/multiple_sets), and a uniform block with std140 arrays whose
    stride exceeds the element size, one of them indexed dynamically."""
    compare(dev, scenes.multiple_sets(200, 150, filt=scenes.LINEAR))
    compare(dev, scenes.ubo_arrays(200, 150))


# ---- front-end overlap (cpvk_cuda_device_set_overlap): a draw's vertex / setup / binning work runs on a second stream while the
# previous draw's raster kernel is busy; results must not depend on it ----

def test_front_end_overlap_alternating_draws(built):
    """Back-to-back frames of two different scenes on one device (its own stream: overlap is on by default), no other command
    in between: draw k+1's front end runs while draw k rasterises, on the other scratch set. Every frame equals the oracle's,
    and equals the same sequence with overlap switched off."""
    from cpvulkan_b200.device import Device, SceneOnDevice
    a = scenes.mesh_indexed(width=640, height=400, nx=160, ny=100)
    b = scenes.random_triangles(width=320, height=240, tris=700, seed=77, indexed=4)
    want = {id(a): scenes.run_oracle(a), id(b): scenes.run_oracle(b)}
    for overlap in (True, False):
        d = Device(0, stats=False)
        d.set_overlap(overlap)
        sa, sb = SceneOnDevice(d, a), SceneOnDevice(d, b)
        try:
            for _ in range(6):
                sa.clear(); sa.draw()
                sb.clear(); sb.draw()
            sa.draw()  # the same draw again on top of its own result, no clear: depth test LESS_OR_EQUAL passes again, same bytes
            for sod, sc in ((sa, a), (sb, b)):
                oc, od, _ = want[id(sc)]
                assert np.array_equal(sod.read_color(), oc), "overlap=%s: colour of %s differs" % (overlap, sc.name)
                assert np.array_equal(sod.read_depth(), od), "overlap=%s: depth of %s differs" % (overlap, sc.name)
        finally:
            sa.close(); sb.close(); d.close()


def test_front_end_overlap_waits_for_a_vertex_buffer_being_rendered(built):
    """Render to vertex buffer: draw 2 reads its vertices from the RGBA32F colour attachment draw 1 is still rendering. The
    front end of draw 2 must not start before draw 1's raster kernel is done (the attachment lies in an allocation draw 2's
    vertex buffer points into: no overlap)."""
    from cpvulkan_b200.device import Device, SceneOnDevice
    s1 = scenes.overdraw_quads(256, 256, quads=40, tex_size=64, blend=False, color_fmt=scenes.R32G32B32A32_SFLOAT)
    s1.textures[0].image.data.reshape(-1, 4)[:, 3] = 255  # alpha 1.0: the pixels become positions with w = 1
    o1, _, _ = scenes.run_oracle(s1)
    tris = 300
    s2 = scenes.random_triangles(width=200, height=160, tris=tris, seed=5, perspective=False)
    s2.buffers["vb"] = np.ascontiguousarray(o1[:tris * 3 * 32])  # vertex i = pixels 2i (position) and 2i + 1 (colour) of draw 1's frame
    o2c, o2d, _ = scenes.run_oracle(s2)
    d = Device(0, stats=False)
    a1, a2 = SceneOnDevice(d, s1), SceneOnDevice(d, s2)
    try:
        a2.m.state.vertexBuffers[0] = a1.m.addr["color"]
        for _ in range(3):
            a1.clear(); a1.draw()
            a2.clear(); a2.draw()
        assert np.array_equal(a1.read_color(), o1)
        assert np.array_equal(a2.read_color(), o2c) and np.array_equal(a2.read_depth(), o2d)
    finally:
        a1.close(); a2.close(); d.close()


def test_front_end_overlap_on_a_caller_stream(built):
    """On a caller-supplied stream overlap is opt-in; with it on, uploads between frames (which change what the vertex stage
    reads) still order the next front end behind them."""
    import torch
    from cpvulkan_b200.device import Device, SceneOnDevice
    stream = torch.cuda.Stream()
    d = Device(0, stream=stream.cuda_stream, stats=False)
    d.set_overlap(True)
    sc = scenes.mesh_indexed(width=480, height=320, nx=120, ny=80)
    sod = SceneOnDevice(d, sc)
    try:
        for k in range(5):
            sod.clear(); sod.draw()
            sod.clear(); sod.draw()
        oc, od, _ = scenes.run_oracle(sc)
        assert np.array_equal(sod.read_color(), oc) and np.array_equal(sod.read_depth(), od)
        # move the mesh: a new uniform matrix between two frames
        ubo = sc.buffers["ubo"].view(np.float32).copy().reshape(4, 4)
        ubo[0, 0] *= 0.5
        sc.buffers["ubo"] = ubo.reshape(-1).view(np.uint8)
        d.upload(sod.m.addr["ubo"], sc.buffers["ubo"])
        sod.clear(); sod.draw()
        oc, od, _ = scenes.run_oracle(sc)
        assert np.array_equal(sod.read_color(), oc) and np.array_equal(sod.read_depth(), od)
    finally:
        sod.close(); d.close()
