"""Generates tests/golden/ref_*.{txt,npz}: outputs of the REFERENCE's own code for the pieces of the path that execute
in this image — CPVulkanBase/Formats.cpp (format table, image layout, GetImagePixelOffset) and CPVulkanBase/FloatFormat.h
(half <-> float, the codec behind R16G16B16A16_SFLOAT) — compiled in place into oracle/_ref/formats_check
(oracle/Makefile, oracle/ref_formats_check.cpp). Run in the build container (needs /root/reference):

    make -C oracle ref && python tests/golden/make_ref_golden.py

tests/test_reference_formats.py checks the oracle against these files everywhere, and — where oracle/_ref exists — that
the files still equal what the reference binary prints."""
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CHECK = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "formats_check")


def run(mode, stdin=None):
    return subprocess.run([CHECK, mode], input=stdin, stdout=subprocess.PIPE, text=True, check=True).stdout


def tohalf_inputs():
    """float32 bit patterns around every half rounding boundary: for each pair of adjacent non-negative half codes the
    exact midpoint (a tie), one ulp either side, and the end points; both signs; plus specials."""
    h = np.arange(0, 0x7C00, dtype=np.uint32)
    lo = np.array([np.float32(np.array([c], dtype=np.uint16).view(np.float16)[0]) for c in (0, 1)], dtype=np.float32)  # warm-up for numpy
    f = h.astype(np.uint16).view(np.float16).astype(np.float32)
    nxt = (h + 1).astype(np.uint16).view(np.float16).astype(np.float32)  # 0x7C00 -> inf
    fb = f.view(np.uint32).astype(np.uint64)
    nb = np.where(np.isinf(nxt), np.float32(65536.0).view(np.uint32), nxt.view(np.uint32)).astype(np.uint64)
    mid = ((fb + nb) // 2).astype(np.uint32)  # same-binade neighbours: the bit average is the arithmetic midpoint
    mid = np.where(f == 0, (np.float32(2.0 ** -25)).view(np.uint32), mid)
    sel = (h % 3 == 0) | (h < 1100) | (h > 0x7B00)
    pats = np.concatenate([fb[sel].astype(np.uint32), mid[sel], mid[sel] - 1, mid[sel] + 1])
    specials = np.array([0x00000000, 0x80000000, 0x7F800000, 0xFF800000, 0x7FC00000, 0x7F800001, 0x7FA12345, 0xFFC00001, 0x7FFFFFFF,
                         0x00000001, 0x007FFFFF, 0x00800000, 0x33000000, 0x33000001, 0x32FFFFFF, 0x33800000, 0x477FE000, 0x477FEFFF,
                         0x477FF000, 0x477FF001, 0x47800000, 0x7F7FFFFF, 0x38800000, 0x387FFFFF, 0x387FE000, 0x387FF000], dtype=np.uint32)
    pats = np.concatenate([pats, pats | 0x80000000, specials])
    return np.unique(pats)


def main():
    open(os.path.join(HERE, "ref_formats.txt"), "w").write(run("formats"))
    open(os.path.join(HERE, "ref_layout.txt"), "w").write(run("layout"))
    half = np.array([int(l.split()[2]) for l in run("half").splitlines()], dtype=np.uint32)
    assert half.shape == (65536,)
    pats = tohalf_inputs()
    out = run("tohalf", "\n".join(str(int(p)) for p in pats) + "\n").splitlines()
    codes = np.array([int(l.split()[2]) for l in out], dtype=np.uint16)
    assert codes.shape == pats.shape
    np.savez_compressed(os.path.join(HERE, "ref_half.npz"), half_to_float_bits=half, float_bits=pats, half_codes=codes)
    print("formats %d lines, half table %d, float->half vectors %d" % (130, len(half), len(pats)))


if __name__ == "__main__":
    main()
