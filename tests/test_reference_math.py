"""Pins the oracle's shader runtime math against the REFERENCE's own code: oracle/_ref/math_check is
LLVMRuntime/SpirvFunctions.cpp compiled as a whole translation unit (OpDot and the OpMatrixTimes* family, reached through the
name table JIT-compiled shaders resolve them from) plus the GLSL.std.450 templates of CPVulkan/GlslFunctions.cpp:19-321 lifted
out of that file at build time (oracle/ref_slice.py), against the glm copy vendored with the reference's samples.
tests/golden/ref_math.npz holds its results for the seeded operands of tests/ref_math_cases.py — FAbs, SAbs, SSign, Sin, Cos, Pow,
F/S/U Min, Max, Clamp, FMix, NMin, NMax, NClamp, Normalise, Reflect on 1..4 lanes, dot on 2..4, matrix * scalar / vector / matrix,
vector * matrix — with in-range, wide-range, tied, and non-finite operands. The oracle's interpreter must return the same bits
(NaN sign / payload canonicalised). Two places follow glm >= 0.9.9 — which the reference needs to compile — rather than the
vendored 0.9.5.3 and are held to the 0.9.9 formula instead: normalize of a vec4 (squares summed pairwise, not left to right)
and nothing else in this file. The CUDA translator is held to the oracle by the -m gpu tests (test_glsl_std_450_subset,
test_matrix_products_in_a_vertex_shader)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import ref_math_cases as mc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "ref_math.npz")
CHECK = os.path.join(ROOT, "oracle", "_ref", "math_check")


def test_oracle_math_matches_the_reference(oracle):
    hdr, A, B, Cc = mc.cases()
    ref = np.load(GOLD)["result_bits"]
    assert len(ref) == len(hdr)
    fn = oracle.cpvk_oracle_math
    fn.argtypes = [C.c_void_p] * 5
    got = np.zeros((len(hdr), 16), dtype=np.uint32)
    for i in range(len(hdr)):
        assert fn(hdr[i].ctypes.data, A[i].ctypes.data, B[i].ctypes.data, Cc[i].ctypes.data, got[i].ctypes.data) == 0, oracle.cpvk_oracle_last_error().decode()
    isf = hdr[:, 2] == 0
    got[isf] = mc.canonical(got[isf])
    version_dependent = (hdr[:, 0] == 0) & (hdr[:, 1] == 69) & (hdr[:, 3] == 4)  # glm::normalize(vec4): 0.9.5.3 vs >= 0.9.9 summation order
    f = np.float32
    for i in np.nonzero(version_dependent)[0]:
        v = A[i][:4].view(np.float32)
        with np.errstate(all="ignore"):
            sq = (v * v).astype(np.float32)
            d = f(f(sq[0] + sq[1]) + f(sq[2] + sq[3]))
            want = (v * f(f(1) / np.sqrt(d, dtype=np.float32))).astype(np.float32)
        assert np.array_equal(mc.canonical(want.view(np.uint32)), got[i][:4]), "normalize(vec4) case %d" % i
    bad = np.nonzero(np.any(got != ref, axis=1) & ~version_dependent)[0]
    assert len(bad) == 0, "%d cases differ; first: header %s a %s b %s c %s oracle %s reference %s" % (
        len(bad), hdr[bad[0]], A[bad[0]][:4].view(np.float32), B[bad[0]][:4].view(np.float32), Cc[bad[0]][:4].view(np.float32),
        got[bad[0]][:4], ref[bad[0]][:4])
    assert len(hdr) > 3000


def test_golden_file_is_what_the_reference_computes(tmp_path):
    if not os.path.exists(CHECK):
        if not os.path.isdir("/root/reference/CPVulkan"):
            pytest.skip("neither the prebuilt oracle/_ref/math_check nor the reference checkout is available")
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    hdr, A, B, Cc = mc.cases()
    src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
    src.write_bytes(mc.file_bytes(hdr, A, B, Cc))
    subprocess.check_call([CHECK, str(src), str(dst)])
    out = np.fromfile(str(dst), dtype="<u4").reshape(len(hdr), 16)
    isf = hdr[:, 2] == 0
    out[isf] = mc.canonical(out[isf])
    assert np.array_equal(out, np.load(GOLD)["result_bits"])
