"""Pins the oracle's vkCmdBlitImage against the REFERENCE's own BlitImageCommand::Process (CPVulkan/CommandBuffer.cpp:57-232).

oracle/_ref/blit_check is that method compiled in place (oracle/Makefile: the method body is lifted out of the file by
oracle/ref_slice.py and runs against the reference's real ImageSampler.cpp, Image.h and Formats.cpp; oracle/ref_blit_check.cpp).
tests/golden/ref_blit.npz holds the destination images it produced for the seeded cases of tests/ref_blit_cases.py: enlargements,
reductions, 1:1, sub-rectangles, flipped extents on either side and on both, one-texel images, NEAREST and LINEAR, values from
1e-30 to 1e7. cpvk_oracle_blit must reproduce every destination bit for bit — the region loops, which texels a flipped region
writes, the u / v / w arithmetic, the eight-tap 3-D sample with its double lerps, untouched texels outside the region. The CUDA
blit is then held to the oracle by tests/test_transfer_gpu.py and tests/test_fullsize_gpu.py (-m gpu)."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

from cpvulkan_b200 import capi

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_blit_cases  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "ref_blit.npz")
CHECK = os.path.join(ROOT, "oracle", "_ref", "blit_check")
RGBA32F = 109


def oracle_blit(lib, h, src, dst):
    src, dst = np.ascontiguousarray(src), np.ascontiguousarray(dst.copy())
    sw, sh, dw, dh, filt = (int(v) for v in h[:5])
    b = capi.Blit(capi.Attachment(src.ctypes.data, sw, sh, sw * 16, RGBA32F), capi.Attachment(dst.ctypes.data, dw, dh, dw * 16, RGBA32F),
                  int(h[5]), int(h[6]), int(h[7]), int(h[8]), int(h[9]), int(h[10]), int(h[11]), int(h[12]), filt)
    assert lib.cpvk_oracle_blit(C.byref(b)) == 0, lib.cpvk_oracle_last_error()
    return dst


def test_oracle_blit_matches_reference(oracle):
    g = np.load(GOLD)
    cs = ref_blit_cases.cases()
    assert len(cs) == int(g["count"])
    for i, (h, src, dst) in enumerate(cs):
        want = g["dst_%d" % i]
        got = oracle_blit(oracle, h, src, dst).view(np.uint32).reshape(-1)
        bad = np.nonzero(got != want)[0]
        assert len(bad) == 0, "case %d %s: %d words differ, first at word %d: oracle %08x reference %08x" % (
            i, h.tolist(), len(bad), bad[0], got[bad[0]], want[bad[0]])
        # the blit wrote something, and (sub-rectangle cases) left the rest of the destination alone
        assert np.any(got != dst.view(np.uint32).reshape(-1))


def test_golden_is_what_the_reference_binary_produces():
    if not os.path.exists(CHECK):
        import pytest
        pytest.skip("oracle/_ref/blit_check not built (no reference checkout): the committed fixture stands")
    g = np.load(GOLD)
    cs = ref_blit_cases.cases()
    with tempfile.TemporaryDirectory() as tmp:
        inp, outp = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
        open(inp, "wb").write(ref_blit_cases.payload(cs))
        subprocess.run([CHECK, inp, outp], check=True)
        raw = np.fromfile(outp, dtype=np.uint32)
    off = 0
    for i, (h, src, dst) in enumerate(cs):
        n = dst.size
        assert np.array_equal(raw[off:off + n], g["dst_%d" % i]), "case %d: the fixture is stale" % i
        off += n
    assert off == len(raw)
