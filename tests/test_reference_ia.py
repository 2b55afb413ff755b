"""Pins the oracle's input assembly against the REFERENCE's own ProcessInputAssembler / ProcessIndexedVertices /
ProcessInputAssemblerIndexed (CPVulkan/CommandBuffer.Draw.cpp:675-760), lifted out of the file and compiled in place into
oracle/_ref/draw_check (`ia` mode, oracle/ref_draw_check.cpp) with a real CPVulkan/Buffer.h object bound as the index buffer.
tests/golden/ref_ia.npz holds the (rawId, vertexId) pairs it produced for the seeded cases of tests/ref_ia_cases.py: first
vertex, 8 / 16 / 32-bit indices, first index, binding offset, vertex offsets of both signs, all six topologies, ragged and empty
counts. The oracle's AssembledVertexId (oracle_draw.cpp) must give the same vertex id for every raw vertex."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from cpvulkan_b200 import capi

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_ia_cases  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "ref_ia.npz")
CHECK = os.path.join(ROOT, "oracle", "_ref", "draw_check")


def test_oracle_input_assembly_matches_the_reference(oracle):
    g = np.load(GOLD)
    cs = ref_ia_cases.cases()
    assert len(cs) == int(g["count"])
    seen = 0
    for i, (h, data) in enumerate(cs):
        want = g["pairs_%d" % i].reshape(-1, 2)
        indexed, first, count, vertex_offset, stride, _, binding_offset, _ = (int(v) for v in h)
        assert len(want) == count and np.array_equal(want[:, 0], np.arange(count, dtype=np.uint32)), "rawId is the position in the draw"
        buf = np.frombuffer(data, dtype=np.uint8).copy() if data else np.zeros(16, dtype=np.uint8)
        st = capi.DrawState()
        st.count, st.first = count, first
        if indexed:
            st.indexBuffer, st.indexStride = buf.ctypes.data + binding_offset, stride  # vkCmdBindIndexBuffer's offset is part of the address
            st.vertexOffset = vertex_offset - (1 << 32) if vertex_offset >= (1 << 31) else vertex_offset
        got = np.zeros(max(count, 1), dtype=np.uint32)
        assert oracle.cpvk_oracle_input_assembly(C.byref(st), got.ctypes.data_as(C.c_void_p)) == 0
        assert np.array_equal(got[:count], want[:, 1]), "case %d %s" % (i, h.tolist())
        seen += count
    assert seen > 2000


def test_golden_is_what_the_reference_binary_produces(tmp_path):
    if not os.path.exists(CHECK):
        pytest.skip("oracle/_ref/draw_check not built (no reference checkout): the committed fixture stands")
    g = np.load(GOLD)
    cs = ref_ia_cases.cases()
    src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
    src.write_bytes(ref_ia_cases.payload(cs))
    subprocess.check_call([CHECK, "ia", str(src), str(dst)])
    for i, pairs in enumerate(ref_ia_cases.parse(dst.read_bytes(), len(cs))):
        assert np.array_equal(pairs.reshape(-1), g["pairs_%d" % i].reshape(-1)), "case %d: the fixture is stale" % i
