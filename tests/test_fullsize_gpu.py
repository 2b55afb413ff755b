"""GPU parity at the sizes BASELINE.json states (north_star: "output must match the reference ICD on the same command
streams"), byte for byte against the CPU oracle through the C ABI:

  C3  the whole 3840x2160 frame of the 1,000,000-triangle indexed mesh (colour RGBA8 and depth D32, every byte);
  C4  all 2,000 alpha-blended LINEAR-textured quads at 7680x4320 RGBA16F — the oracle renders windows of that frame
      (cpvk_oracle_draw_window: pixels are independent in the reference, Draw.cpp:1526-1593): the four corners, the centre,
      a tile corner, the last columns of the frame, and every boundary between the bands eight GPUs would own; the same
      frame rendered as eight bands one after the other into one image must be the same bytes as the unbanded frame;
  C5  vkCmdBlitImage 8K RGBA8 -> RGBA16F NEAREST and 4K -> 8K RGBA16F LINEAR on windows (cpvk_oracle_blit_window),
      vkCmdCopyImage 8K whole (a copy is the source's bytes), the Samples/texel_buffer draw at 7680x4320 on windows.
"""
import ctypes as C

import numpy as np
import pytest

from cpvulkan_b200 import capi, scenes
from cpvulkan_b200.device import SceneOnDevice, run_cuda

pytestmark = pytest.mark.gpu


def oracle_windows(scene, windows):
    """Clear once, then render only the pixels inside each window with the oracle. Returns the colour attachment bytes
    (valid inside the windows) and the fragments counted over all windows."""
    lib = capi.load_oracle()
    mem = scenes.HostMemory()
    m = scenes.materialize(scene, mem.alloc)
    color = mem.arrays["color"]
    texel = scenes.TEXEL_SIZE[scene.color.format]
    cv, is_ds = scenes.clear_value(scene.color)
    # the clear is a per-texel SetPixel of one value: only the windows need it (clearing 33 M texels one by one takes longer
    # than everything else here), through the oracle's own clear on a sub-rectangle view of the attachment
    for (x0, y0, x1, y1) in windows:
        sub = capi.Attachment(m.color_attachment.address + y0 * m.color_attachment.rowPitch + x0 * texel, x1 - x0, y1 - y0,
                              m.color_attachment.rowPitch, m.color_attachment.format)
        assert lib.cpvk_oracle_clear(C.byref(sub), C.byref(cv), is_ds) == 0
    covered = 0
    for w in windows:
        st = capi.DrawStats()
        rc = lib.cpvk_oracle_draw_window(C.byref(m.desc), C.byref(m.state), *w, C.byref(st))
        assert rc == 0, lib.cpvk_oracle_last_error().decode()
        covered += int(st.fragmentsCovered)
    return color[:scene.color.nbytes], covered


def assert_windows_equal(scene, got, want, windows):
    texel = scenes.TEXEL_SIZE[scene.color.format]
    a = np.asarray(got).reshape(scene.color.height, scene.color.width * texel)
    b = np.asarray(want).reshape(scene.color.height, scene.color.width * texel)
    for (x0, y0, x1, y1) in windows:
        ga, gb = a[y0:y1, x0 * texel:x1 * texel], b[y0:y1, x0 * texel:x1 * texel]
        if not np.array_equal(ga, gb):
            bad = np.argwhere(ga != gb)
            raise AssertionError("window (%d,%d)-(%d,%d): %d bytes differ, first at x=%d y=%d: gpu %d oracle %d"
                                 % (x0, y0, x1, y1, len(bad), x0 + bad[0][1] // texel, y0 + bad[0][0], ga[tuple(bad[0])], gb[tuple(bad[0])]))


def test_c3_full_frame(dev):
    scene = scenes.mesh_indexed()  # 3840x2160, 1000x500 quads
    assert (scene.color.width, scene.color.height, scene.count) == (3840, 2160, 3000000)
    oc, od, ost = scenes.run_oracle(scene)
    gc, gd, gst = run_cuda(dev, scene)
    assert (gst.primitives, gst.fragmentsCovered, gst.fragmentsWritten) == (ost.primitives, ost.fragmentsCovered, ost.fragmentsWritten)
    assert ost.primitives == 1000000 and ost.fragmentsCovered > 8000000
    assert np.array_equal(gc, oc), "C3 colour attachment differs from the oracle (%d bytes)" % int((gc != oc).sum())
    assert np.array_equal(gd, od), "C3 depth attachment differs from the oracle (%d bytes)" % int((gd != od).sum())


def c4_windows(width, height, bands=8):
    w = [(0, 0, 16, 16), (width - 16, 0, width, 16), (0, height - 16, 16, height), (width - 16, height - 16, width, height),
         (width // 2 - 8, height // 2 - 8, width // 2 + 8, height // 2 + 8),
         (24, 24, 40, 40),                           # the corner four 32x32 tiles share
         (width - 8, 1000, width, 1016)]             # the last columns of the frame
    rows = height // bands
    for k in range(1, bands):                        # every boundary between the bands of an 8-GPU split
        x0 = (k * 911) % (width - 8)
        w.append((x0, k * rows - 4, x0 + 8, k * rows + 4))
    return w


def test_c4_full_size_against_oracle_windows(dev):
    scene = scenes.overdraw_quads()  # 7680x4320 RGBA16F, 2,000 quads, 1024^2 RGBA8 texture, LINEAR / REPEAT
    assert (scene.color.width, scene.color.height, scene.count) == (7680, 4320, 12000)
    windows = c4_windows(7680, 4320)
    want, covered = oracle_windows(scene, windows)
    assert covered >= 2000 * sum((x1 - x0) * (y1 - y0) for x0, y0, x1, y1 in windows)  # every quad covers every pixel (those on the shared diagonal twice)
    s = SceneOnDevice(dev, scene)
    try:
        s.render()
        st = dev.stats()
        got = s.read_color()
        assert st.primitives == 4000
        # every pixel is covered by each quad once; pixels whose centre the shared diagonal hits exactly are drawn by both triangles
        assert 2000 * 7680 * 4320 <= st.fragmentsCovered <= 2000 * (7680 * 4320 + 7680)
        assert_windows_equal(scene, got, want, windows)
        # sort-first: eight bands rendered one after the other into one frame == the unbanded frame, byte for byte
        rows = 4320 // 8
        for k in range(8):
            s.m.state.bandY0, s.m.state.bandY1 = k * rows, (k + 1) * rows
            s.clear(band_only=True)
            s.draw()
        banded = s.read_color()
        assert np.array_equal(banded, got), "the frame assembled from 8 bands differs from the unbanded frame"
    finally:
        s.close()


class DevImage:
    def __init__(self, dev, fmt, width, height, texel, seed):
        self.dev, self.fmt, self.width, self.height, self.pitch = dev, fmt, width, height, width * texel
        self.nbytes = self.pitch * height
        self.host = np.random.default_rng(seed).integers(0, 256, self.nbytes, dtype=np.uint8)
        self.addr = dev.alloc(self.nbytes)
        dev.upload(self.addr, self.host)

    def att(self, side):
        return capi.Attachment(self.host.ctypes.data if side == "host" else self.addr, self.width, self.height, self.pitch, self.fmt)

    def free(self):
        self.dev.free(self.addr)


@pytest.mark.parametrize("case", ["rgba8_to_rgba16f_nearest_1to1", "rgba16f_4k_to_8k_linear", "rgba8_4k_to_8k_rgba16f_linear"])
def test_c5_blit_8k_windows(dev, case):
    lib = capi.load_oracle()
    W, H = 7680, 4320
    if case == "rgba8_to_rgba16f_nearest_1to1":
        src = DevImage(dev, 37, W, H, 4, seed=31); filt = 0
    elif case == "rgba16f_4k_to_8k_linear":
        src = DevImage(dev, 97, W // 2, H // 2, 8, seed=32); filt = 1
        # random bytes as half floats include inf / NaN codes; a NaN's sign and payload differ between x86 and the GPU, so the
        # source holds finite halves only (exponent 31 -> 15); every finite code, denormals included, stays
        h16 = src.host.view(np.uint16)
        h16[(h16 & 0x7C00) == 0x7C00] &= 0xBFFF
        dev.upload(src.addr, src.host)
    else:
        src = DevImage(dev, 37, W // 2, H // 2, 4, seed=33); filt = 1
    dst = DevImage(dev, 97, W, H, 8, seed=34)
    hb = capi.Blit(src.att("host"), dst.att("host"), 0, 0, src.width, src.height, 0, 0, W, H, filt)
    db = capi.Blit(src.att("dev"), dst.att("dev"), 0, 0, src.width, src.height, 0, 0, W, H, filt)
    windows = [(0, 0, 24, 24), (W - 24, 0, W, 24), (0, H - 24, 24, H), (W - 24, H - 24, W, H), (W // 2 - 12, H // 2 - 12, W // 2 + 12, H // 2 + 12),
               (250, 20, 262, 44), (5000, 3000, 5040, 3008), (W - 3, 2000, W, 2100)]
    for w in windows:
        assert lib.cpvk_oracle_blit_window(C.byref(hb), *w) == 0, lib.cpvk_oracle_last_error()
    dev.blit(db)
    got = dev.download(dst.addr, dst.nbytes).reshape(H, W * 8)
    want = dst.host.reshape(H, W * 8)
    for (x0, y0, x1, y1) in windows:
        a, b = got[y0:y1, x0 * 8:x1 * 8], want[y0:y1, x0 * 8:x1 * 8]
        # NaN halves: the reference's float -> half keeps a truncated payload; both sides are the same code path, compare bytes
        assert np.array_equal(a, b), "%s window (%d,%d)-(%d,%d): %d bytes differ" % (case, x0, y0, x1, y1, int((a != b).sum()))
    src.free(); dst.free()


def test_c5_copy_image_8k(dev):
    W, H = 7680, 4320
    src = DevImage(dev, 37, W, H, 4, seed=41)
    dst = DevImage(dev, 37, W, H, 4, seed=42)
    dev.copy_rows(dst.addr, dst.pitch, src.addr, src.pitch, W * 4, H)
    lib = capi.load_oracle()
    assert lib.cpvk_oracle_copy_rows(dst.host.ctypes.data, dst.pitch, src.host.ctypes.data, src.pitch, W * 4, H) == 0
    assert np.array_equal(dev.download(dst.addr, dst.nbytes), dst.host)
    src.free(); dst.free()


def test_c5_texel_buffer_8k_windows(dev):
    scene = scenes.texel_buffer(7680, 4320)
    windows = [(0, 0, 32, 32), (7680 - 32, 4320 - 32, 7680, 4320), (3800, 2100, 3880, 2180), (1000, 4000, 1064, 4064), (7000, 100, 7064, 164)]
    want, _ = oracle_windows(scene, windows)
    got, _, st = run_cuda(dev, scene)
    assert st.primitives == 1
    assert_windows_equal(scene, got, want, windows)
