"""Pins the oracle's rasteriser, interpolator and blend code against the REFERENCE's own CPVulkan/CommandBuffer.Draw.cpp.

oracle/_ref/draw_check is built by oracle/Makefile from the reference's source text: oracle/ref_slice.py lifts
EdgeFunction (:410-418), CalculatePrimitives (:567-673), SetDatum / GetFragmentInput (:816-954), ApplyBlendFactor /
ApplyBlend (:956-1262) and DrawPixel / ProcessPoints / ProcessLines / ProcessTriangles (:1300-1594) out of the file where it
lies, and oracle/ref_draw_check.cpp drives them with a recording function in place of the JIT-compiled fragment shader.
tests/golden/ref_draw.npz holds what that binary produced for the seeded draws of tests/ref_draw_cases.py: per emitted
fragment, in emission order, (x, y, front, depth as passed to the shader, fragCoord, interpolated inputs) — 137 612
fragments over 33 draws (random, snapped to pixel centres / corners, shared edges, zero area, NaN / inf / w = 0, both
windings x every cull mode, strips, fans, lines, points) — and ApplyBlend over 5 050 states (every factor pair x every op,
separate alpha factors and ops) x in-range / HDR / non-finite / tied operands.

The oracle (oracle_draw.cpp: cpvk_oracle_raster_records, cpvk_oracle_apply_blend) must reproduce all of it bit for bit:
fragment order, coverage, facing, depth, fragCoord and every interpolated word; a NaN must be a NaN on both sides, its
sign / payload bits are canonicalised (ref_draw_cases.canonical says why). The CUDA path is then held to the oracle by the -m gpu
parity tests."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

import ref_draw_cases as rc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "ref_draw.npz")
CHECK = os.path.join(ROOT, "oracle", "_ref", "draw_check")


def oracle_stream(oracle, c):
    fn = oracle.cpvk_oracle_raster_records
    fn.restype = C.c_int64
    fn.argtypes = [C.c_float] * 5 + [C.c_uint32] * 6 + [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64]
    cap = 1 << 16
    while True:
        buf = np.zeros(cap, dtype=np.uint32)
        n = fn(c.width, c.height, c.min_depth, c.max_depth, c.line_width, c.topology, c.front_face, c.cull, c.origin_upper,
               c.vertex_count, rc.STRIDE, rc.INPUTS.ctypes.data, len(rc.INPUTS), c.records.ctypes.data, buf.ctypes.data, cap)
        assert n >= 0, oracle.cpvk_oracle_last_error().decode()
        if n * rc.WORDS <= cap:
            return rc.canonical(buf[:n * rc.WORDS])
        cap = int(n * rc.WORDS)


FIELDS = ["x", "y", "front", "depth", "fragCoord.x", "fragCoord.y", "fragCoord.z", "fragCoord.w", "in0.x", "in0.y", "in0.z", "in0.w",
          "in1.x", "in1.y", "flat", "in3"]


def describe(c, got, ref):
    n = min(len(got), len(ref))
    bad = np.nonzero(np.any(got[:n] != ref[:n], axis=1))[0]
    if len(bad) == 0:
        return "%s: oracle emitted %d fragments, reference %d" % (c.name, len(got), len(ref))
    i = int(bad[0])
    cols = np.nonzero(got[i] != ref[i])[0]
    return "%s: fragment %d differs in %s: oracle %s reference %s" % (
        c.name, i, [FIELDS[k] for k in cols], [hex(int(got[i][k])) for k in cols], [hex(int(ref[i][k])) for k in cols])


def test_oracle_fragment_streams_match_the_reference(oracle):
    g = np.load(GOLD)
    cases = rc.raster_cases()
    assert [c.name for c in cases] == list(g["names"])
    total = 0
    for i, c in enumerate(cases):
        got = oracle_stream(oracle, c)
        assert len(got) == int(g["counts"][i]), "%s: oracle emitted %d fragments, reference %d" % (c.name, len(got), int(g["counts"][i]))
        if c.full:
            ref = g["stream_%d" % i]
            assert np.array_equal(got, ref.reshape(got.shape)), describe(c, got, ref)
        assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == str(g["sha256"][i]), \
            "%s: the oracle's fragment stream differs from the reference's (SHA-256 over %d fragments)" % (c.name, len(got))
        total += len(got)
    assert total >= 100000


def test_oracle_blend_matches_the_reference(oracle):
    g = np.load(GOLD)
    states, operands = rc.blend_cases()
    fn = oracle.cpvk_oracle_apply_blend
    fn.argtypes = [C.c_void_p] * 5
    got = np.zeros((len(states), 4), dtype=np.float32)
    for i in range(len(states)):
        rcode = fn(states[i].ctypes.data, operands[i, 0:4].ctypes.data, operands[i, 4:8].ctypes.data, operands[i, 8:12].ctypes.data, got[i].ctypes.data)
        assert rcode == 0, oracle.cpvk_oracle_last_error().decode()
    ref = g["blend_bits"]
    got = rc.canonical_floats(got.view(np.uint32)).view(np.float32)
    # MIN / MAX through glm::min / glm::max is the one place where the glm VERSION shows: the copy vendored with the
    # reference's samples (0.9.5.3, what draw_check is built with) spells min(x, y) = x < y ? x : y, while glm >= 0.9.9 — which
    # the reference needs to compile at all (glm::vec<L, T>, gtx/vec_swizzle.hpp) — spells it (y < x) ? y : x. The two differ
    # only when an operand is NaN or for min / max(+0, -0); those lanes are held to the 0.9.9 form here, every other lane
    # and case to the reference binary.
    src, dst = operands[:, 0:4], operands[:, 4:8]
    viaGlm = np.zeros((len(states), 4), dtype=bool)
    viaGlm[:, 0:3] = (states[:, 3] >= 3)[:, None]
    viaGlm[:, 3] = (states[:, 3] >= 3) & (states[:, 3] == states[:, 6])
    versionDependent = viaGlm & (np.isnan(src) | np.isnan(dst) | ((src == 0) & (dst == 0) & (np.signbit(src) != np.signbit(dst))))
    with np.errstate(invalid="ignore"):
        glm099 = np.where((states[:, 3] == 3)[:, None], np.where(dst < src, dst, src), np.where(src < dst, dst, src))
    assert np.array_equal(rc.canonical_floats(got.view(np.uint32))[versionDependent], rc.canonical_floats(glm099.view(np.uint32))[versionDependent])
    assert versionDependent.sum() < 0.02 * versionDependent.size
    bad = np.nonzero(np.any((got.view(np.uint32) != ref) & ~versionDependent, axis=1))[0]
    assert len(bad) == 0, "blend state %s on %s: oracle %s reference %s (%d cases differ)" % (
        states[bad[0]], operands[bad[0]], got[bad[0]], ref[bad[0]].view(np.float32), len(bad))
    # the state space the fixture covers
    assert {(int(s[1]), int(s[2]), int(s[3])) for s in states} >= {(a, b, o) for a in range(15) for b in range(15) for o in range(5)}
    assert {(int(s[4]), int(s[5]), int(s[6])) for s in states if s[3] != s[6]} >= {(a, b, o) for a in range(15) for b in range(15) for o in range(5)}


def test_golden_file_is_what_the_reference_produces(tmp_path):
    if not os.path.exists(CHECK):
        if not os.path.isdir("/root/reference/CPVulkan"):
            pytest.skip("neither the prebuilt oracle/_ref/draw_check nor the reference checkout is available")
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    g = np.load(GOLD)
    cases = rc.raster_cases()
    src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
    src.write_bytes(rc.raster_file(cases))
    subprocess.check_call([CHECK, "raster", str(src), str(dst)])
    streams = [rc.canonical(t) for t in rc.parse_raster_output(dst.read_bytes(), len(cases))]
    for i, (c, s) in enumerate(zip(cases, streams)):
        assert len(s) == int(g["counts"][i]), c.name
        assert hashlib.sha256(np.ascontiguousarray(s).tobytes()).hexdigest() == str(g["sha256"][i]), c.name
        if c.full:
            assert np.array_equal(s, g["stream_%d" % i].reshape(s.shape)), c.name
    states, operands = rc.blend_cases()
    src.write_bytes(rc.blend_file(states, operands))
    subprocess.check_call([CHECK, "blend", str(src), str(dst)])
    assert np.array_equal(rc.canonical_floats(np.frombuffer(dst.read_bytes(), dtype="<u4").reshape(len(states), 4)), g["blend_bits"])
