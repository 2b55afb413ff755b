"""Pins the oracle's texture sampler against the REFERENCE's own CPVulkan/ImageSampler.cpp.

oracle/_ref/sampler_check is that file compiled in place (oracle/Makefile, oracle/ref_sampler_check.cpp: shim headers for
Vulkan / GSL / glm, a raw-copy texel function for R32G32B32A32_SFLOAT instead of the LLVM JIT). tests/golden/ref_sampler.npz
holds what it returned for a 3-level RGBA32F mip chain under 90 sampler configurations x 128 coordinates: all five
address modes, NEAREST / LINEAR on the magnification and minification paths, both mipmap modes, fractional LODs, border
colours, mixed U/V modes. The oracle (oracle_sampler.h) must reproduce every result bit for bit — including the double
lerp chain and the ceil(lod + 0.5) - 1 mip choice. The CUDA sampler is then held to the oracle by the -m gpu tests."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from cpvulkan_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "ref_sampler.npz")
CHECK = os.path.join(ROOT, "oracle", "_ref", "sampler_check")
RGBA32F = 109


def descriptor(levels, cfg):
    d = capi.Descriptor()
    d.type, d.format, d.dimensions, d.levelCount = capi.DESC_IMAGE, RGBA32F, 2, len(levels)
    for i, t in enumerate(levels):
        d.levels[i] = capi.MipLevel(t.ctypes.data, t.shape[1], t.shape[0], 1, 0)
    s = d.sampler
    s.magFilter, s.minFilter, s.mipmapMode = int(cfg["mag"]), int(cfg["min"]), int(cfg["mipmap"])
    s.addressModeU, s.addressModeV, s.addressModeW = int(cfg["au"]), int(cfg["av"]), 0
    s.borderColor = int(cfg["border"])
    s.mipLodBias, s.minLod, s.maxLod = 0.0, 0.0, 1000.0  # VK_LOD_CLAMP_NONE: the lod reaches SampleImage unchanged
    return d


def test_oracle_sampler_matches_reference(oracle):
    g = np.load(GOLD)
    levels = [np.ascontiguousarray(g["level%d" % i]) for i in range(3)]
    coords = np.concatenate([g["coords"], np.zeros((len(g["coords"]), 1), dtype=np.float32)], axis=1).astype(np.float32)
    coords = np.ascontiguousarray(coords)
    for k, cfg in enumerate(g["configs"]):
        d = descriptor(levels, cfg)
        got = np.zeros((len(coords), 4), dtype=np.float32)
        oracle.cpvk_oracle_sample(C.byref(d), coords.ctypes.data_as(C.c_void_p), len(coords), C.c_float(float(cfg["lod"])), got.ctypes.data_as(C.c_void_p))
        bad = np.nonzero(np.any(got.view(np.uint32) != g["result_bits"][k], axis=1))[0]
        assert len(bad) == 0, "config %d %s coordinate %s: oracle %s reference %s" % (
            k, cfg, coords[bad[0]], got[bad[0]], g["result_bits"][k][bad[0]].view(np.float32))


def test_golden_file_is_what_the_reference_returns(tmp_path):
    if not os.path.exists(CHECK):
        if not os.path.isdir("/root/reference/CPVulkan"):
            pytest.skip("neither the prebuilt oracle/_ref/sampler_check nor the reference checkout is available")
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    g = np.load(GOLD)
    path = tmp_path / "in.bin"
    with open(path, "wb") as f:
        f.write(np.array([3, len(g["configs"]), len(g["coords"])], dtype="<u4").tobytes())
        for i in range(3):
            t = g["level%d" % i]
            f.write(np.array([t.shape[1], t.shape[0]], dtype="<u4").tobytes())
            f.write(np.ascontiguousarray(t).tobytes())
        f.write(g["configs"].tobytes())
        f.write(np.ascontiguousarray(g["coords"]).tobytes())
    out = subprocess.run([CHECK, str(path)], stdout=subprocess.PIPE, text=True, check=True).stdout
    bits = np.array([[int(x) for x in l.split()] for l in out.splitlines()], dtype=np.uint32).reshape(g["result_bits"].shape)
    assert np.array_equal(bits, g["result_bits"])
    g3 = np.load(os.path.join(ROOT, "tests", "golden", "ref_sampler3d.npz"))
    for tag in ("a", "b"):
        tex, coords = g3["tex_" + tag], g3["coords_" + tag]
        path3 = tmp_path / ("in3d_%s.bin" % tag)
        with open(path3, "wb") as f:
            f.write(np.array([tex.shape[2], tex.shape[1], tex.shape[0], len(coords)], dtype="<u4").tobytes())
            f.write(np.ascontiguousarray(tex).tobytes()); f.write(np.ascontiguousarray(coords).tobytes())
        out = subprocess.run([CHECK, "3d", str(path3)], stdout=subprocess.PIPE, text=True, check=True).stdout
        bits = np.array([[int(x) for x in l.split()] for l in out.splitlines()], dtype=np.uint32).reshape(g3["bits_" + tag].shape)
        assert np.array_equal(bits, g3["bits_" + tag])


def test_oracle_3d_sampling_as_used_by_blit_matches_reference(oracle):
    """SampleImage(state, format, data, uvec3 range, fvec3 coordinates, filter) — the overload vkCmdBlitImage calls
    (CommandBuffer.cpp:75-226): lod 1, CLAMP_TO_EDGE on all axes, the minification filter on the only level, eight taps and
    seven double lerps when LINEAR. tests/golden/ref_sampler3d.npz holds the reference binary's results."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_sampler3d.npz"))
    for tag in ("a", "b"):
        tex = np.ascontiguousarray(g["tex_" + tag]); coords = np.ascontiguousarray(g["coords_" + tag])
        depth, h, w, _ = tex.shape
        for filt in (0, 1):
            d = capi.Descriptor()
            d.type, d.format, d.dimensions, d.levelCount = capi.DESC_IMAGE, RGBA32F, 3, 1
            d.levels[0] = capi.MipLevel(tex.ctypes.data, w, h, depth, 0)
            s = d.sampler
            s.magFilter = s.minFilter = filt
            s.addressModeU = s.addressModeV = s.addressModeW = 2  # CLAMP_TO_EDGE
            s.mipLodBias, s.minLod, s.maxLod = 0.0, 0.0, 1000.0
            got = np.zeros((len(coords), 4), dtype=np.float32)
            oracle.cpvk_oracle_sample(C.byref(d), coords.ctypes.data_as(C.c_void_p), len(coords), C.c_float(1.0), got.ctypes.data_as(C.c_void_p))
            bad = np.nonzero(np.any(got.view(np.uint32) != g["bits_" + tag][filt], axis=1))[0]
            assert len(bad) == 0, "image %s filter %d coordinate %s: oracle %s reference %s" % (
                tag, filt, coords[bad[0]], got[bad[0]], g["bits_" + tag][filt][bad[0]].view(np.float32))
