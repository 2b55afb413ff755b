"""Pins the hand-assembled SPIR-V (no glslang / spirv-as in the image) with the REFERENCE's own SPIR-V front end:
oracle/_ref/spirv_check is SPIRVParser/ from /root/reference compiled in place (oracle/Makefile) and loads each module
exactly like CPVulkan/ShaderModule.cpp:62-73. What it reports — validity, Logical/GLSL450 model, entry point, and the
module-order variable list that VS->FS linkage is built on — must match what our shaders intend."""
import os
import subprocess

import pytest

from cpvulkan_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK = os.path.join(ROOT, "oracle", "_ref", "spirv_check")

EXPECTED = {  # (storage class, Location) in module order; storage: 0 UniformConstant, 1 Input, 2 Uniform, 3 Output, 6 Private, 9 PushConstant
    "cube.vert": (0, [(3, 0), (1, 1), (3, -1), (2, -1), (1, 0)]),
    "cube.frag": (4, [(3, 0), (1, 0)]),
    "builtins.vert": (0, [(3, 0), (1, 1), (3, -1), (1, -1), (1, -1), (1, 0), (1, 2)]),               # VertexIndex, InstanceIndex: Inputs without Location
    "points.vert": (0, [(3, 0), (1, 1), (3, -1), (1, 0), (1, 2)]),                                   # + gl_PointSize from an attribute
    "matmath.vert": (0, [(3, 0), (1, 1), (3, -1), (2, -1), (1, 0), (1, 2)]),                          # two mat4 + a float in one UBO
    "texcube.vert": (0, [(3, 0), (1, 1), (3, -1), (2, -1), (1, 0)]),
    "texcube.frag": (4, [(3, 0), (0, -1), (1, 0)]),
    "texelbuf.vert": (0, [(6, -1), (0, -1), (6, -1), (6, -1), (3, 0), (6, -1), (3, -1), (1, -1)]),  # 6 = Private
    "complex.frag": (4, [(9, -1), (1, 0), (3, 0)]),                                                  # 9 = PushConstant
    "sepsampler.frag": (4, [(3, 0), (0, -1), (0, -1), (1, 0)]),                                      # texture2D + sampler
    "subpass.frag": (4, [(3, 0), (0, -1)]),                                                          # subpassInput
    "uintout.frag": (4, [(3, 0), (1, 0)]),                                                           # uvec4 output
    "sintout.frag": (4, [(3, 0), (1, 0)]),                                                           # ivec4 output
    "mrt.frag": (4, [(3, 1), (3, 0), (1, 0)]),                                                       # two outputs, declared 1 then 0
    "glslmath.frag": (4, [(3, 0), (1, 0)]),
    "flat.frag": (4, [(3, 0), (1, 0)]),
    "nopersp.frag": (4, [(3, 0), (1, 0)]),
    "varyings.frag": (4, [(3, 0), (1, 4), (1, 0), (1, 2), (1, 1), (1, 3), (1, 5)]),  # test-only (tests/test_reference_interface.py)
    "fragcoord.frag": (4, [(3, 0), (1, 0), (1, -1)]),                                                # gl_FragCoord: Input, no Location
    "multisets.frag": (4, [(3, 0), (0, -1), (1, 0)]),                                                # the sampler in descriptor set 1
    "uboarray.vert": (0, [(3, 0), (1, 1), (3, -1), (2, -1), (2, -1), (1, 0)]),                      # two uniform blocks (sets 0 and 1), arrays with ArrayStride
}


@pytest.fixture(scope="module")
def checker():
    if not os.path.exists(CHECK):
        if not os.path.isdir("/root/reference/SPIRVParser"):
            pytest.skip("neither the prebuilt oracle/_ref/spirv_check nor the reference checkout is available")
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    return CHECK


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_reference_front_end_accepts_our_shaders(checker, tmp_path, name):
    path = tmp_path / (name + ".spv")
    scenes.shader(name).tofile(str(path))
    out = subprocess.run([checker, str(path)], stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()
    assert out[0] == "valid 1"
    assert out[1] == "memory_model 1 addressing 0"  # GLSL450 / Logical (required: SPIRVCompiler.cpp:732-735)
    model, variables = EXPECTED[name]
    assert any(l.startswith("entry %d main" % model) for l in out)
    got = [(int(l.split()[3]), int(l.split()[5])) for l in out if l.startswith("variable")]
    assert got == variables


def test_every_shader_in_the_tree_is_checked():
    have = {f[:-len(".spvasm")] for f in os.listdir(os.path.join(ROOT, "cpvulkan_b200", "shaders")) if f.endswith(".spvasm")}
    assert have == set(EXPECTED), have ^ set(EXPECTED)
