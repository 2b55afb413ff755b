"""vkCreateGraphicsPipelines' compile path on the CPU: SPIR-V -> CUDA C++ -> NVRTC (LTO-IR) -> nvJitLink -> sm_100a cubin.
nvcc/NVRTC/nvJitLink cross-compile without a GPU, so this is the 'does every pipeline build' check."""
import ctypes as C

import numpy as np
import pytest

from cpvulkan_b200 import capi, scenes


def compile_only(scene, mutate=None):
    lib = capi.load_cuda()
    m = scenes.materialize(scene, scenes.HostMemory().alloc)
    if mutate:
        mutate(m.desc)
    p = C.c_void_p()
    rc = lib.cpvk_cuda_pipeline_compile_only(C.byref(m.desc), C.byref(p))
    return lib, rc, p, m


@pytest.mark.parametrize("scene", [scenes.draw_cube(64, 64), scenes.draw_textured_cube(64, 64, scenes.LINEAR), scenes.overdraw_quads(64, 64, 2, 16)], ids=lambda s: s.name)
def test_pipelines_link_to_sm100a_cubin(built, scene):
    lib, rc, p, _ = compile_only(scene)
    assert rc == 0, lib.cpvk_cuda_last_error()
    src = lib.cpvk_cuda_pipeline_source(p).decode()
    assert 'extern "C" __device__ void cpvk_vs_main' in src and 'extern "C" __device__ bool cpvk_fs_main' in src
    # glm mat4*vec4 operand order (m0*v0 + m1*v1) + (m2*v2 + m3*v3), SURVEY H1
    assert ") + (" in src
    n = C.c_size_t()
    cubin = lib.cpvk_cuda_pipeline_cubin(p, C.byref(n))
    blob = C.string_at(cubin, n.value)
    assert blob[:4] == b"\x7fELF" and b"cpvk_k_raster" in blob and b"cpvk_k_vertex" in blob
    lib.cpvk_cuda_pipeline_destroy(None, p)


def test_reference_aborts_are_refused(built):
    def logic(d):
        d.logicOpEnable = 1
    lib, rc, _, _ = compile_only(scenes.draw_cube(32, 32), logic)
    assert rc == capi.E_UNSUPPORTED and b"logic op" in lib.cpvk_cuda_last_error()

    def wire(d):
        d.polygonMode = 1
    lib, rc, _, _ = compile_only(scenes.draw_cube(32, 32), wire)
    assert rc == capi.E_UNSUPPORTED

    def adjacency(d):
        d.topology = 6
    lib, rc, _, _ = compile_only(scenes.draw_cube(32, 32), adjacency)
    assert rc == capi.E_UNSUPPORTED and b"adjacency" in lib.cpvk_cuda_last_error()

    def points(d):  # points and lines are built (ProcessPoints / ProcessLines)
        d.topology = 0
    lib, rc, _, _ = compile_only(scenes.draw_cube(32, 32), points)
    assert rc == 0


def test_malformed_spirv_is_an_error(built):
    lib = capi.load_cuda()
    m = scenes.materialize(scenes.draw_cube(32, 32), scenes.HostMemory().alloc)
    junk = np.array([0x07230203, 0x10000, 0, 8, 0, 0xFFFF0013], dtype=np.uint32)
    m.desc.vertex.spirv = junk.ctypes.data_as(C.POINTER(C.c_uint32))
    m.desc.vertex.wordCount = len(junk)
    p = C.c_void_p()
    assert lib.cpvk_cuda_pipeline_compile_only(C.byref(m.desc), C.byref(p)) == capi.E_SPIRV


def test_missing_vertex_attribute_is_refused(built):
    def drop(d):
        d.attributeCount = 1
    lib, rc, _, _ = compile_only(scenes.draw_cube(32, 32), drop)
    assert rc == capi.E_UNSUPPORTED and b"location" in lib.cpvk_cuda_last_error()


def _with_fs(scene, fs):
    scene.fs = fs
    return scene


def _with_vs(scene, vs):
    scene.vs = vs
    return scene


FRONT_END_SCENES = {
    "separate_image_sampler": lambda: scenes.separate_image_sampler(64, 64),
    "input_attachment": lambda: scenes.input_attachment(64, 64),
    "uint_output": lambda: _with_fs(scenes.random_triangles(width=32, height=32, tris=2, color_fmt=41), "uintout.frag"),
    "sint_output": lambda: _with_fs(scenes.random_triangles(width=32, height=32, tris=2, color_fmt=42), "sintout.frag"),
    "flat": lambda: _with_fs(scenes.random_triangles(width=32, height=32, tris=2), "flat.frag"),
    "noperspective": lambda: _with_fs(scenes.random_triangles(width=32, height=32, tris=2), "nopersp.frag"),
    "frag_coord": lambda: _with_fs(scenes.random_triangles(width=32, height=32, tris=2), "fragcoord.frag"),
    "glsl_math": lambda: _with_fs(scenes.random_triangles(width=32, height=32, tris=2, color_fmt=scenes.R32G32B32A32_SFLOAT), "glslmath.frag"),
    "push_spec_loop": lambda: _with_fs(scenes.random_triangles(width=32, height=32, tris=2), "complex.frag"),
    "builtins": lambda: _with_vs(scenes.random_points_lines(width=32, height=32, count=4, topology=scenes.POINT_LIST), "builtins.vert"),
    "lines": lambda: scenes.random_points_lines(width=32, height=32, count=4, topology=scenes.LINE_STRIP),
    "texel_buffer": lambda: scenes.texel_buffer(32, 32),
    "multiple_sets": lambda: scenes.multiple_sets(64, 64),
    "ubo_arrays": lambda: scenes.ubo_arrays(64, 64),
}


@pytest.mark.parametrize("name", sorted(FRONT_END_SCENES))
def test_every_shader_family_links(built, name):
    """Each shader in the tree, in a pipeline of the kind the GPU tests use it in, goes all the way to an sm_100a cubin here."""
    lib, rc, p, _ = compile_only(FRONT_END_SCENES[name]())
    assert rc == 0, lib.cpvk_cuda_last_error()
    n = C.c_size_t()
    cubin = lib.cpvk_cuda_pipeline_cubin(p, C.byref(n))
    blob = C.string_at(cubin, n.value)
    assert blob[:4] == b"\x7fELF" and b"cpvk_k_raster" in blob and b"cpvk_k_vertex" in blob
    lib.cpvk_cuda_pipeline_destroy(None, p)


def test_truncated_and_out_of_range_spirv_is_refused_not_read_out_of_bounds(built):
    """Every prefix of a valid module cut inside an instruction, and modules whose type operands name ids beyond the bound,
    must come back as an error (round-1 ADVICE: operands were indexed before their presence was checked)."""
    import numpy as np
    words = scenes.shader("texcube.frag").copy()
    lib = capi.load_cuda()

    def compile_words(ws):
        m = scenes.materialize(scenes.draw_textured_cube(32, 32), scenes.HostMemory().alloc)
        ws = np.ascontiguousarray(ws, dtype=np.uint32)
        m.desc.fragment.spirv = ws.ctypes.data_as(C.POINTER(C.c_uint32)); m.desc.fragment.wordCount = len(ws)
        p = C.c_void_p()
        rc = lib.cpvk_cuda_pipeline_compile_only(C.byref(m.desc), C.byref(p))
        if rc == 0:
            lib.cpvk_cuda_pipeline_destroy(None, p)
        return rc

    # shorten single instructions in place (word count field reduced, following words dropped)
    i, tried = 5, 0
    while i < len(words):
        wc, op = int(words[i]) >> 16, int(words[i]) & 0xFFFF
        if wc >= 3 and op in (19, 21, 22, 23, 24, 25, 27, 28, 30, 32, 33, 43, 59, 54, 71, 72, 15):  # types, constants, variables, decorations, entry point, function
            cut = np.concatenate([words[:i], [np.uint32(((wc - 1) << 16) | op)], words[i + 1:i + wc - 1], words[i + wc:]])
            assert compile_words(cut) != 0 or op in (71, 15), "op %d shortened to %d words was accepted" % (op, wc - 1)
            tried += 1
        i += wc
    assert tried > 10
    # a type operand beyond the id bound
    bad = words.copy()
    i = 5
    while i < len(bad):
        wc, op = int(bad[i]) >> 16, int(bad[i]) & 0xFFFF
        if op == 23:  # OpTypeVector: component type id
            bad[i + 2] = 0x00FFFFFF
            break
        i += wc
    assert compile_words(bad) != 0
