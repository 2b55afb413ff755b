"""GPU parity of the transfer commands around the draw path (BASELINE config C5): ClearImage,
vkCmdCopyImage row copies and vkCmdBlitImage, through the C ABI against the CPU oracle. Bit-exact: every byte of the
destination allocation is compared, including the row padding the commands must not touch."""
import ctypes as C

import numpy as np
import pytest

from cpvulkan_b200 import capi, scenes

pytestmark = pytest.mark.gpu

R8_UNORM, R8G8_UNORM, R8G8B8_UNORM, R8G8B8A8_UNORM, R8G8B8A8_SNORM, B8G8R8A8_UNORM = 9, 16, 23, 37, 38, 44
R8G8B8A8_UINT, R8G8B8A8_SINT, R8G8B8A8_SRGB = 41, 42, 43
A2B10G10R10_UNORM = 64
R16_UNORM, R16G16_SFLOAT, R16G16B16_UNORM, R16G16B16A16_UNORM, R16G16B16A16_SFLOAT, R16G16B16A16_UINT = 70, 83, 84, 91, 97, 95
R32_UINT, R32_SFLOAT, R32G32_SFLOAT, R32G32B32_SFLOAT, R32G32B32A32_SFLOAT, R32G32B32A32_SINT = 98, 100, 103, 106, 109, 108
D16_UNORM, X8_D24, D32_SFLOAT, S8_UINT, D16_S8, D24_S8, D32_S8 = 124, 125, 126, 127, 128, 129, 130


def texel_size(fmt):
    out = (C.c_uint32 * 4)()
    assert capi.load_oracle().cpvk_oracle_format_info(fmt, C.byref(out)) == 0
    return int(out[2])


class Image:
    """The same bytes on both sides: a numpy array for the oracle, an HBM allocation for the CUDA path."""

    def __init__(self, dev, fmt, width, height, pad=0, seed=0):
        self.dev, self.fmt, self.width, self.height = dev, fmt, width, height
        self.pitch = texel_size(fmt) * width + pad
        self.nbytes = self.pitch * height
        self.host = np.random.default_rng(seed).integers(0, 256, self.nbytes, dtype=np.uint8)
        self.addr = dev.alloc(self.nbytes)
        dev.upload(self.addr, self.host)

    def att(self, side):
        return capi.Attachment(self.host.ctypes.data if side == "host" else self.addr, self.width, self.height, self.pitch, self.fmt)

    def check(self):
        got = self.dev.download(self.addr, self.nbytes)
        if not np.array_equal(got, self.host):
            bad = np.flatnonzero(got != self.host)
            raise AssertionError("format %d: %d bytes differ, first at offset %d (row %d): oracle %d gpu %d"
                                 % (self.fmt, len(bad), bad[0], bad[0] // self.pitch, self.host[bad[0]], got[bad[0]]))

    def free(self):
        self.dev.free(self.addr)


COLOR_CLEARS = [
    (R8_UNORM, (0.3, 0, 0, 0)), (R8G8_UNORM, (0.2, 0.9, 0, 0)), (R8G8B8_UNORM, (0.1, 0.5, 1.0, 0)), (R8G8B8A8_UNORM, (0.2, 0.4, 0.6, 1.0)),
    (R8G8B8A8_UNORM, (-1.0, 2.0, float("nan"), 0.5)), (R8G8B8A8_SNORM, (-0.7, 0.7, -1.5, 1.0)), (B8G8R8A8_UNORM, (0.25, 0.5, 0.75, 1.0)),
    (R8G8B8A8_SRGB, (0.2, 0.4, 0.6, 0.5)), (A2B10G10R10_UNORM, (0.1, 0.2, 0.3, 0.67)), (R16_UNORM, (0.123, 0, 0, 0)),
    (R16G16_SFLOAT, (1.5, -3.25e-5, 0, 0)), (R16G16B16_UNORM, (0.1, 0.2, 0.3, 0)), (R16G16B16A16_UNORM, (0.9, 0.8, 0.7, 0.6)),
    (R16G16B16A16_SFLOAT, (65504.0, 1e-8, -2.0, 0.333)), (R32_SFLOAT, (3.14159, 0, 0, 0)), (R32G32_SFLOAT, (1e30, -1e-30, 0, 0)),
    (R32G32B32_SFLOAT, (1.0, 2.0, 3.0, 0)), (R32G32B32A32_SFLOAT, (0.1, 0.2, 0.3, 0.4)),
]
INT_CLEARS = [(R8G8B8A8_UINT, (1, 2, 300, 255)), (R8G8B8A8_SINT, (-1, 127, -128, 5)), (R16G16B16A16_UINT, (65535, 70000, 3, 4)),
              (R32_UINT, (0xDEADBEEF, 0, 0, 0)), (R32G32B32A32_SINT, (-5, 6, -7, 8))]
DEPTH_CLEARS = [(D16_UNORM, 0.5, 0), (X8_D24, 0.25, 0), (D32_SFLOAT, 1.0, 0), (S8_UINT, 0.0, 0x5A), (D16_S8, 0.75, 3), (D24_S8, 0.123, 200), (D32_S8, 0.999, 255)]


def clear_both(dev, img, cv, is_ds):
    lib = capi.load_oracle()
    host_att, dev_att = img.att("host"), img.att("dev")
    assert lib.cpvk_oracle_clear(C.byref(host_att), C.byref(cv), is_ds) == 0, lib.cpvk_oracle_last_error()
    dev.clear(dev_att, cv, is_ds)
    img.check()


@pytest.mark.parametrize("lazy", [True, False], ids=["deferred", "immediate"])
@pytest.mark.parametrize("fmt,val", COLOR_CLEARS, ids=lambda v: str(v) if isinstance(v, int) else None)
def test_clear_color(dev, fmt, val, lazy):
    dev.set_lazy_clear(lazy)
    try:
        for (w, h, pad) in ((64, 32, 0), (37, 19, 0), (33, 7, 48)):
            img = Image(dev, fmt, w, h, pad, seed=fmt)
            cv = capi.ClearValue()
            for i in range(4):
                cv.float32[i] = val[i]
            clear_both(dev, img, cv, 0)
            img.free()
    finally:
        dev.set_lazy_clear(True)


@pytest.mark.parametrize("fmt,val", INT_CLEARS, ids=lambda v: str(v) if isinstance(v, int) else None)
def test_clear_integer(dev, fmt, val):
    img = Image(dev, fmt, 41, 13, 16, seed=fmt)
    cv = capi.ClearValue()
    for i in range(4):
        cv.uint32[i] = val[i] & 0xFFFFFFFF
    clear_both(dev, img, cv, 0)
    img.free()


@pytest.mark.parametrize("fmt,depth,stencil", DEPTH_CLEARS)
def test_clear_depth_stencil(dev, fmt, depth, stencil):
    img = Image(dev, fmt, 50, 21, 0, seed=fmt)
    cv = capi.ClearValue()
    cv.depthStencil.depth, cv.depthStencil.stencil = depth, stencil
    clear_both(dev, img, cv, 1)
    img.free()


def test_clear_then_clear_then_upload_order(dev):
    """Deferred clears keep command order: a second clear of the same image wins, a later upload wins over both, and an
    overlapping clear of a sub-range is applied after the first one."""
    img = Image(dev, R8G8B8A8_UNORM, 32, 16, 0, seed=1)
    cv1, cv2 = capi.ClearValue(), capi.ClearValue()
    for i in range(4):
        cv1.float32[i], cv2.float32[i] = 0.25, 0.75
    lib = capi.load_oracle()
    dev_att, host_att = img.att("dev"), img.att("host")
    dev.clear(dev_att, cv1, 0); dev.clear(dev_att, cv2, 0)
    lib.cpvk_oracle_clear(C.byref(host_att), C.byref(cv1), 0); lib.cpvk_oracle_clear(C.byref(host_att), C.byref(cv2), 0)
    # overlapping sub-image (rows 4..8) cleared with the first value again
    sub_dev = capi.Attachment(img.addr + 4 * img.pitch, 32, 4, img.pitch, img.fmt)
    sub_host = capi.Attachment(img.host.ctypes.data + 4 * img.pitch, 32, 4, img.pitch, img.fmt)
    dev.clear(sub_dev, cv1, 0); lib.cpvk_oracle_clear(C.byref(sub_host), C.byref(cv1), 0)
    # upload one row on top
    row = np.arange(img.pitch, dtype=np.uint8)
    dev.upload(img.addr + 5 * img.pitch, row); img.host[5 * img.pitch:6 * img.pitch] = row
    img.check()
    img.free()


@pytest.mark.parametrize("row_bytes,rows,dst_pad,src_pad,offset", [(256, 16, 0, 0, 0), (100, 9, 28, 12, 0), (33, 5, 7, 3, 1), (4096, 64, 0, 0, 0), (48, 3, 16, 16, 16)])
def test_copy_rows(dev, row_bytes, rows, dst_pad, src_pad, offset):
    lib = capi.load_oracle()
    src_pitch, dst_pitch = row_bytes + src_pad, row_bytes + dst_pad
    rng = np.random.default_rng(row_bytes)
    src = rng.integers(0, 256, src_pitch * rows + offset, dtype=np.uint8)
    dst = rng.integers(0, 256, dst_pitch * rows + offset, dtype=np.uint8)
    d_src, d_dst = dev.alloc(src.nbytes), dev.alloc(dst.nbytes)
    dev.upload(d_src, src); dev.upload(d_dst, dst)
    lib.cpvk_oracle_copy_rows(dst.ctypes.data + offset, dst_pitch, src.ctypes.data + offset, src_pitch, row_bytes, rows)
    dev.copy_rows(d_dst + offset, dst_pitch, d_src + offset, src_pitch, row_bytes, rows)
    assert np.array_equal(dev.download(d_dst, dst.nbytes), dst)
    dev.free(d_src); dev.free(d_dst)


BLITS = [
    # src fmt, dst fmt, src size, dst size, src rect, dst rect, filter
    (R8G8B8A8_UNORM, R8G8B8A8_UNORM, (32, 32), (32, 32), (0, 0, 32, 32), (0, 0, 32, 32), 0),
    (R8G8B8A8_UNORM, R8G8B8A8_UNORM, (32, 32), (64, 48), (0, 0, 32, 32), (0, 0, 64, 48), 1),       # magnify, linear
    (R8G8B8A8_UNORM, B8G8R8A8_UNORM, (64, 64), (20, 20), (0, 0, 64, 64), (0, 0, 20, 20), 1),       # minify + swizzled format
    (R8G8B8A8_UNORM, R16G16B16A16_SFLOAT, (40, 30), (50, 35), (5, 5, 35, 25), (10, 3, 45, 30), 1), # sub-rects + conversion
    (R16G16B16A16_SFLOAT, R8G8B8A8_UNORM, (16, 16), (33, 17), (0, 0, 16, 16), (0, 0, 33, 17), 0),
    (R32G32B32A32_SFLOAT, R16G16B16A16_SFLOAT, (24, 24), (24, 24), (0, 0, 24, 24), (24, 24, 0, 0), 0),  # mirrored in x and y
    (R8G8B8A8_UNORM, R8G8B8A8_UNORM, (32, 32), (32, 32), (32, 0, 0, 32), (0, 0, 32, 32), 1),       # mirrored source
    (R8_UNORM, R8G8B8A8_UNORM, (19, 11), (38, 22), (0, 0, 19, 11), (0, 0, 38, 22), 1),
    (R8G8B8A8_SRGB, R8G8B8A8_UNORM, (16, 16), (16, 16), (0, 0, 16, 16), (0, 0, 16, 16), 0),
    (A2B10G10R10_UNORM, R16G16B16A16_UNORM, (16, 8), (32, 16), (0, 0, 16, 8), (0, 0, 32, 16), 1),
]


@pytest.mark.parametrize("case", BLITS, ids=lambda c: "%d-%d-%dx%d-f%d" % (c[0], c[1], c[3][0], c[3][1], c[6]))
def test_blit(dev, case):
    sfmt, dfmt, ssize, dsize, srect, drect, filt = case
    lib = capi.load_oracle()
    src = Image(dev, sfmt, ssize[0], ssize[1], 0, seed=3)
    if sfmt in (R16G16B16A16_SFLOAT, R32G32B32A32_SFLOAT):  # finite, moderate values instead of random bit patterns
        vals = np.random.default_rng(4).uniform(-2.0, 2.0, ssize[0] * ssize[1] * 4)
        src.host[:] = vals.astype(np.float16 if sfmt == R16G16B16A16_SFLOAT else np.float32).view(np.uint8)
        dev.upload(src.addr, src.host)
    dst = Image(dev, dfmt, dsize[0], dsize[1], 48, seed=5)

    def make(side):
        return capi.Blit(src.att(side), dst.att(side), srect[0], srect[1], srect[2], srect[3], drect[0], drect[1], drect[2], drect[3], filt)

    hb, db = make("host"), make("dev")
    assert lib.cpvk_oracle_blit(C.byref(hb)) == 0, lib.cpvk_oracle_last_error()
    dev.blit(db)
    dst.check()
    src.free(); dst.free()


def test_blit_8k_round_trip_property(dev):
    """BASELINE C5 at full size (8K): far too large for the oracle, so check size-independent properties —
    a 1:1 NEAREST blit is the identity, and copy_rows of the result reproduces the source bytes exactly."""
    w, h = 7680, 4320
    src = Image(dev, R8G8B8A8_UNORM, w, h, 0, seed=9)
    dst = Image(dev, R8G8B8A8_UNORM, w, h, 0, seed=10)
    b = capi.Blit(src.att("dev"), dst.att("dev"), 0, 0, w, h, 0, 0, w, h, 0)
    dev.blit(b)
    assert np.array_equal(dev.download(dst.addr, dst.nbytes), src.host)
    third = Image(dev, R8G8B8A8_UNORM, w, h, 0, seed=11)
    dev.copy_rows(third.addr, third.pitch, dst.addr, dst.pitch, w * 4, h)
    assert np.array_equal(dev.download(third.addr, third.nbytes), src.host)
    src.free(); dst.free(); third.free()


def test_misaligned_images_are_refused(dev):
    """Rows of a linear image are texel-aligned in the reference (Stride = texel size x width); the ABI refuses anything
    else instead of faulting in a vector store."""
    from cpvulkan_b200.device import CpvkError
    img = Image(dev, R16G16B16A16_SFLOAT, 8, 8, 0, seed=1)
    bad = capi.Attachment(img.addr, 8, 8, img.pitch + 4, img.fmt)
    with pytest.raises(CpvkError):
        dev.clear(bad, capi.ClearValue(), 0)
    with pytest.raises(CpvkError):
        dev.blit(capi.Blit(bad, img.att("dev"), 0, 0, 8, 8, 0, 0, 8, 8, 0))
    img.free()


# ---- empty and degenerate regions ----

def test_empty_copies_and_blits_change_nothing(dev):
    """Zero rows, zero row bytes, and blit rectangles with no area (the reference's loops simply do not execute,
    CommandBuffer.cpp:75-226, CommandBuffer.Copy.cpp): the destination keeps every byte."""
    lib = capi.load_oracle()
    src = Image(dev, R8G8B8A8_UNORM, 16, 16, 0, seed=21)
    dst = Image(dev, R8G8B8A8_UNORM, 16, 16, 0, seed=22)
    dev.copy_rows(dst.addr, dst.pitch, src.addr, src.pitch, 0, 16)
    dev.copy_rows(dst.addr, dst.pitch, src.addr, src.pitch, 64, 0)
    for rect in ((4, 4, 4, 12), (4, 4, 12, 4), (0, 0, 0, 0)):
        hb = capi.Blit(src.att("host"), dst.att("host"), 0, 0, 16, 16, rect[0], rect[1], rect[2], rect[3], 1)
        db = capi.Blit(src.att("dev"), dst.att("dev"), 0, 0, 16, 16, rect[0], rect[1], rect[2], rect[3], 1)
        assert lib.cpvk_oracle_blit(C.byref(hb)) == 0
        dev.blit(db)
    dst.check()
    src.free(); dst.free()


@pytest.mark.parametrize("filt", [0, 1])
def test_one_texel_images(dev, filt):
    lib = capi.load_oracle()
    src = Image(dev, R8G8B8A8_UNORM, 1, 1, 0, seed=23)
    dst = Image(dev, R16G16B16A16_SFLOAT, 7, 5, 0, seed=24)
    hb = capi.Blit(src.att("host"), dst.att("host"), 0, 0, 1, 1, 0, 0, 7, 5, filt)
    db = capi.Blit(src.att("dev"), dst.att("dev"), 0, 0, 1, 1, 0, 0, 7, 5, filt)
    assert lib.cpvk_oracle_blit(C.byref(hb)) == 0
    dev.blit(db)
    dst.check()
    back = Image(dev, R8G8B8A8_UNORM, 1, 1, 0, seed=25)
    hb = capi.Blit(dst.att("host"), back.att("host"), 0, 0, 7, 5, 0, 0, 1, 1, filt)
    db = capi.Blit(dst.att("dev"), back.att("dev"), 0, 0, 7, 5, 0, 0, 1, 1, filt)
    assert lib.cpvk_oracle_blit(C.byref(hb)) == 0
    dev.blit(db)
    back.check()
    src.free(); dst.free(); back.free()
