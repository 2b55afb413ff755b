"""Pins the oracle's shader-side image functions against the REFERENCE's own (CPVulkan/GlslFunctions.cpp:324-737), lifted out of
that file and compiled in place into oracle/_ref/image_check against the reference's real ImageSampler.cpp and its Image, ImageView,
Buffer, BufferView, Sampler and ImageDescriptor types (oracle/ref_image_check.cpp). tests/golden/ref_image.npz holds what
ImageSampleExplicitLod<fvec4, fvec2>, ImageFetch<fvec4, ivec2> and ImageFetch<fvec4, int32_t> returned for the seeded cases of
tests/ref_image_cases.py. The oracle's ImageSampleExplicitLod / ImageFetch (oracle_sampler.h) must reproduce every result bit for
bit: LOD bias clamped to +-MAX_SAMPLER_LOD_BIAS and added to the explicit lod, the min / max LOD clamp, the mip levels a view
selects (GetImageData: the descriptor the C ABI takes carries the VIEW's levels, computed here from the reference-pinned image
layout as the ICD's FillDescriptor does), the component swizzle, texelFetch on level 0 of the view with out-of-range coordinates,
and texel-buffer views (offset, range / texel size). SURVEY §8(a) a13; BASELINE C2 / C4 (texture()) and C5 (texel buffer)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from cpvulkan_b200 import capi

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_image_cases as rc  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "ref_image.npz")
CHECK = os.path.join(ROOT, "oracle", "_ref", "image_check")
RGBA32F = 109


def descriptor(oracle, hdr, f3, buf):
    kind, w, h, mips, base, count = (int(v) for v in hdr[:6])
    d = capi.Descriptor()
    d.format = RGBA32F
    for i in range(4):
        d.swizzle[i] = int(hdr[6 + i])
    s = d.sampler
    s.magFilter, s.minFilter, s.mipmapMode, s.addressModeU, s.addressModeV, s.borderColor = (int(v) for v in hdr[10:16])
    s.mipLodBias, s.minLod, s.maxLod = float(f3[0]), float(f3[1]), float(f3[2])
    if kind == 2:
        d.type, d.dimensions = capi.DESC_TEXEL_BUFFER, 1
        d.address, d.range = buf.ctypes.data + int(hdr[17]), int(hdr[18])
        return d
    layout = (C.c_uint64 * (3 + 6 * mips))()
    assert oracle.cpvk_oracle_image_layout(RGBA32F, w, h, 1, 1, mips, layout) == 0
    levels = mips - base if count == rc.REMAINING else count  # GetFormatOffset, GlslFunctions.cpp:339-345
    d.type, d.dimensions, d.levelCount = capi.DESC_IMAGE, 2, levels
    for i in range(levels):
        l = layout[3 + 6 * (i + base):3 + 6 * (i + base) + 6]
        d.levels[i] = capi.MipLevel(buf.ctypes.data + int(l[0]), int(l[3]), int(l[4]), 1, 0)
    return d


def test_oracle_image_functions_match_the_reference(oracle):
    g = np.load(GOLD)
    cs = rc.cases()
    assert len(cs) == int(g["count"])
    for i, (hdr, f3, data, coords) in enumerate(cs):
        want = g["result_%d" % i].reshape(-1, 4)
        buf = np.frombuffer(data, dtype=np.uint8).copy()
        d = descriptor(oracle, hdr, f3, buf)
        got = np.zeros((len(coords), 4), dtype=np.float32)
        if int(hdr[0]) == 0:
            for k, row in enumerate(coords):
                uvl = row.view(np.float32)
                c3 = np.array([uvl[0], uvl[1], 0.0], dtype=np.float32)
                oracle.cpvk_oracle_sample(C.byref(d), c3.ctypes.data_as(C.c_void_p), 1, C.c_float(float(uvl[2])), got[k].ctypes.data_as(C.c_void_p))
        else:
            xyz = np.ascontiguousarray(coords.view(np.int32))
            oracle.cpvk_oracle_fetch(C.byref(d), xyz.ctypes.data_as(C.c_void_p), len(xyz), got.ctypes.data_as(C.c_void_p))
        bad = np.nonzero(np.any(got.view(np.uint32) != want, axis=1))[0]
        assert len(bad) == 0, "case %d %s %s coordinate %d: oracle %s reference %s" % (
            i, hdr.tolist(), f3.tolist(), bad[0], got[bad[0]], want[bad[0]].view(np.float32))


def test_golden_is_what_the_reference_binary_produces(tmp_path):
    if not os.path.exists(CHECK):
        pytest.skip("oracle/_ref/image_check not built (no reference checkout): the committed fixture stands")
    g = np.load(GOLD)
    cs = rc.cases()
    src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
    src.write_bytes(rc.payload(cs))
    subprocess.check_call([CHECK, str(src), str(dst)])
    raw = np.frombuffer(dst.read_bytes(), dtype="<u4")
    off = 0
    for i, (hdr, f3, data, coords) in enumerate(cs):
        n = 4 * len(coords)
        assert np.array_equal(raw[off:off + n], g["result_%d" % i].reshape(-1)), "case %d: the fixture is stale" % i
        off += n
    assert off == len(raw)
