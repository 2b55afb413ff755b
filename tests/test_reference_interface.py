"""Pins the oracle's VS -> FS linkage arithmetic against the REFERENCE's own GetVariableFormat / GetVariableSize /
GetVariablePointers (CPVulkan/CommandBuffer.Draw.cpp:151-354, :420-565), lifted out of that file and compiled in place into
oracle/_ref/interface_check, running on modules loaded by the reference's own SPIR-V front end (SPIRVParser/). For every fragment
shader in the tree the oracle's Reflect (oracle_draw.cpp) must list the same inputs in the same order with the same Location,
interpolation format (what SetDatum walks), interpolation kind, size and byte offset inside the vertex stage's record, and end
at the same input size. The expected lines are committed below; where oracle/_ref exists the test
also checks that they are still what the reference binary prints."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from cpvulkan_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK = os.path.join(ROOT, "oracle", "_ref", "interface_check")

# name -> ([(location, format, interpolation, size, offset)], input size at the end, [(set, binding)], push-constant bytes or -1)
# as printed by `interface_check fragment <file>`; formats are VkFormat values (100 / 103 / 106 / 109 = 1 .. 4 floats, 98 = uint)
EXPECTED = {
    "complex.frag": ([(0, 109, 0, 16, 24)], 40, [], 24),
    "cube.frag": ([(0, 109, 0, 16, 24)], 40, [], -1),
    "flat.frag": ([(0, 109, 2, 16, 24)], 40, [], -1),
    "fragcoord.frag": ([(0, 109, 0, 16, 24)], 40, [], -1),
    "glslmath.frag": ([(0, 109, 0, 16, 24)], 40, [], -1),
    "mrt.frag": ([(0, 109, 0, 16, 24)], 40, [], -1),
    "multisets.frag": ([(0, 103, 0, 8, 24)], 32, [(1, 0)], -1),
    "nopersp.frag": ([(0, 109, 1, 16, 24)], 40, [], -1),
    "sepsampler.frag": ([(0, 103, 0, 8, 24)], 32, [(0, 1), (0, 2)], -1),
    "sintout.frag": ([(0, 109, 0, 16, 24)], 40, [], -1),
    "subpass.frag": ([], 24, [(0, 1)], -1),
    "texcube.frag": ([(0, 103, 0, 8, 24)], 32, [(0, 1)], -1),
    "uintout.frag": ([(0, 109, 0, 16, 24)], 40, [], -1),
    # test-only shader: six inputs declared out of Location order — vec3 (12 bytes on this side), vec4, flat uint, noperspective vec2, float, flat ivec2
    "varyings.frag": ([(4, 106, 0, 12, 24), (0, 109, 0, 16, 36), (2, 98, 2, 4, 52), (1, 103, 1, 8, 56), (3, 100, 0, 4, 64), (5, 102, 2, 8, 68)], 76, [], -1),
}


def reference_lines(name, tmp_path):
    path = tmp_path / (name + ".spv")
    scenes.shader(name).tofile(str(path))
    return subprocess.run([CHECK, "fragment", str(path)], stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_oracle_fragment_inputs_match_the_reference(oracle, name):
    inputs, end, _, _ = EXPECTED[name]
    words = np.ascontiguousarray(scenes.shader(name))
    out = np.zeros(256, dtype=np.uint32)
    n = oracle.cpvk_oracle_fragment_inputs(words.ctypes.data_as(C.c_void_p), len(words), out.ctypes.data_as(C.c_void_p), len(out))
    assert n > 0, oracle.cpvk_oracle_last_error()
    count = int(out[0])
    got = [tuple(int(v) for v in out[1 + 5 * i:6 + 5 * i]) for i in range(count)]
    assert got == [tuple(t) for t in inputs]
    assert int(out[1 + 5 * count]) == end


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_expected_lines_are_what_the_reference_prints(tmp_path, name):
    if not os.path.exists(CHECK):
        pytest.skip("oracle/_ref/interface_check not built (no reference checkout): the committed table stands")
    lines = reference_lines(name, tmp_path)
    inputs, end, uniforms, push = EXPECTED[name]
    assert [tuple(int(x) for x in l.split()[1:]) for l in lines if l.startswith("input")] == [tuple(t) for t in inputs]
    assert [tuple(int(x) for x in l.split()[1:]) for l in lines if l.startswith("uniform")] == [tuple(t) for t in uniforms]
    assert [int(l.split()[1]) for l in lines if l.startswith("push")] == [push]
    assert [int(l.split()[1]) for l in lines if l.startswith("sizes")] == [end]


def test_every_fragment_shader_in_the_tree_is_checked():
    have = {f[:-len(".spvasm")] for f in os.listdir(os.path.join(ROOT, "cpvulkan_b200", "shaders")) if f.endswith(".frag.spvasm")}
    assert have == set(EXPECTED), have ^ set(EXPECTED)
