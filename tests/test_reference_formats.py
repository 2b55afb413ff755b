"""Pins the oracle's format table, image layout and half codec against the REFERENCE's own code.

CPVulkanBase/Formats.cpp and CPVulkanBase/FloatFormat.h are the arithmetic of the path that the
reference can execute in this image: oracle/Makefile compiles them in place into oracle/_ref/formats_check, and
tests/golden/make_ref_golden.py stored what that binary prints under tests/golden/ref_*. Here
  * the oracle must equal those golden files (runs everywhere, no reference needed), and
  * where oracle/_ref/formats_check exists (it travels to the GPU box), its output must still equal the files.
The CUDA path is then compared with the oracle by the -m gpu parity tests (clears, blits and draws over these formats)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CHECK = os.path.join(ROOT, "oracle", "_ref", "formats_check")

# formats the reference lists but the path does not build (64-bit channels, 4/5/6-bit packs, shared-exponent / 10-11-11 floats)
NOT_BUILT = set(range(1, 9)) | set(range(110, 124))


def lines(name):
    return open(os.path.join(GOLD, name)).read().splitlines()


@pytest.fixture(scope="module")
def checker():
    if not os.path.exists(CHECK):
        if not os.path.isdir("/root/reference/CPVulkanBase"):
            pytest.skip("neither the prebuilt oracle/_ref/formats_check nor the reference checkout is available")
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    return CHECK


def test_golden_files_are_what_the_reference_prints(checker):
    for mode, name in (("formats", "ref_formats.txt"), ("layout", "ref_layout.txt")):
        out = subprocess.run([checker, mode], stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()
        assert out == lines(name)
    g = np.load(os.path.join(GOLD, "ref_half.npz"))
    out = subprocess.run([checker, "half"], stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()
    assert np.array_equal(np.array([int(l.split()[2]) for l in out], dtype=np.uint32), g["half_to_float_bits"])
    pats = g["float_bits"][::37]
    out = subprocess.run([checker, "tohalf"], input="\n".join(str(int(p)) for p in pats) + "\n", stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()
    assert np.array_equal(np.array([int(l.split()[2]) for l in out], dtype=np.uint16), g["half_codes"][::37])


def test_format_table_matches_reference(oracle):
    seen = 0
    for l in lines("ref_formats.txt"):
        t = l.split()
        f, ref = int(t[1]), [int(x) for x in t[2:]]
        row = (C.c_uint32 * 12)()
        rc = oracle.cpvk_oracle_format_row(f, row)
        if f in NOT_BUILT:
            assert rc != 0, "format %d is documented as not built" % f
            continue
        assert rc == 0, "format %d missing from the oracle" % f
        got = list(row)
        if ref[0] == 1:  # Packed: the reference leaves ElementSize at whatever the union holds; only TotalSize is used
            got[2] = ref[2]
        assert got == ref, "format %d: oracle %s reference %s" % (f, got, ref)
        seen += 1
    assert seen == 130 - len(NOT_BUILT)


def test_image_layout_and_pixel_offset_match_reference(oracle):
    n = 0
    for l in lines("ref_layout.txt"):
        head, sizes, levels, probes = [p.split() for p in l.split("|")]
        f, w, h, d, layers, mips = [int(x) for x in head[1:]]
        out = (C.c_uint64 * (3 + 6 * mips))()
        assert oracle.cpvk_oracle_image_layout(f, w, h, d, layers, mips, out) == 0
        assert list(out)[:3] == [int(x) for x in sizes]
        assert list(out)[3:] == [int(x) for x in levels]
        for level, want in enumerate(int(x) for x in probes):  # probe texel per level: (w-1, h/2, d-1) of the last layer
            lw, lh, ld = out[3 + 6 * level + 3], out[3 + 6 * level + 4], out[3 + 6 * level + 5]
            assert oracle.cpvk_oracle_pixel_offset(out, lw - 1, lh // 2, ld - 1, level, layers - 1) == want
        n += 1
    assert n == 70


def test_half_to_float_all_codes_match_reference(oracle):
    want = np.load(os.path.join(GOLD, "ref_half.npz"))["half_to_float_bits"]
    got = np.array([oracle.cpvk_oracle_half_to_float(c) for c in range(65536)], dtype=np.float32).view(np.uint32)
    # ctypes returns floats by value: a signalling NaN may be quieted on the way through the x87/SSE return path, so
    # NaN codes are compared as "both NaN with the same sign"; everything else bit for bit
    nan = (want & 0x7F800000 == 0x7F800000) & (want & 0x007FFFFF != 0)
    assert np.array_equal(got[~nan], want[~nan])
    assert np.all((got[nan] & 0x7F800000 == 0x7F800000) & (got[nan] & 0x007FFFFF != 0) & ((got[nan] >> 31) == (want[nan] >> 31)))


def test_float_to_half_boundaries_match_reference(oracle):
    g = np.load(os.path.join(GOLD, "ref_half.npz"))
    pats, want = g["float_bits"], g["half_codes"]
    vals = pats.view(np.float32)
    got = np.array([oracle.cpvk_oracle_float_to_half(C.c_float(float(v))) if not np.isnan(v) else 0 for v in vals], dtype=np.uint16)
    ok = ~np.isnan(vals)  # NaN payloads cannot cross ctypes' float argument unchanged; they are pinned by test_oracle_kats
    bad = np.nonzero(got[ok] != want[ok])[0]
    assert len(bad) == 0, "first mismatch: float bits %#x oracle %#x reference %#x" % (pats[ok][bad[0]], got[ok][bad[0]], want[ok][bad[0]])
    assert ok.sum() > 90000
