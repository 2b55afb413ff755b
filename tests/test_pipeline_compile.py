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
