"""CPU-side checks of the drop-in boundary: the C ABI library loads, exports every symbol include/cpvk_cuda.h declares,
agrees with the ctypes mirror on every struct size, and fails loudly (no CPU fallback) without a CUDA device."""
import ctypes as C
import os
import re

import pytest

from cpvulkan_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_all_exported(built):
    header = open(os.path.join(ROOT, "include", "cpvk_cuda.h")).read()
    declared = set(re.findall(r"\b(cpvk_cuda_[a-z_]+)\s*\(", header))
    assert declared == set(capi.ABI_SYMBOLS), declared ^ set(capi.ABI_SYMBOLS)
    lib = capi.load_cuda()
    for name in declared:
        assert getattr(lib, name) is not None


def test_struct_sizes_match_c(built):
    lib = capi.load_cuda()
    for name, st in capi.ABI_STRUCTS.items():
        assert lib.cpvk_cuda_abi_sizeof(name.encode()) == C.sizeof(st), name
    assert lib.cpvk_cuda_abi_version() == 1


def test_icd_exports_loader_entry_points(built):
    icd = C.CDLL(os.path.join(ROOT, "cpvulkan_b200", "icd", "build", "libCPVulkan_b200.so"))
    v = C.c_uint32(7)
    assert icd.vk_icdNegotiateLoaderICDInterfaceVersion(C.byref(v)) == 0 and v.value == 5  # caps at 5 (CPVulkan.cpp:95-104)
    icd.vk_icdGetInstanceProcAddr.restype = C.c_void_p
    icd.vk_icdGetInstanceProcAddr.argtypes = [C.c_void_p, C.c_char_p]
    for fn in ("vkCreateInstance", "vkCreateGraphicsPipelines", "vkCmdBindVertexBuffers", "vkCmdBindDescriptorSets", "vkCmdDraw", "vkCmdDrawIndexed",
               "vkQueueSubmit", "vkMapMemory", "vkCmdBeginRenderPass", "vkCmdBlitImage", "vkCmdCopyImage", "vkGetDeviceProcAddr"):
        assert icd.vk_icdGetInstanceProcAddr(None, fn.encode()), fn
    assert not icd.vk_icdGetInstanceProcAddr(None, b"vkCmdDispatch")  # outside the draw path
    icd.vk_icdGetPhysicalDeviceProcAddr.restype = C.c_void_p
    assert not icd.vk_icdGetPhysicalDeviceProcAddr(None, b"vkCreateInstance")
    manifest = open(os.path.join(ROOT, "cpvulkan_b200", "icd", "build", "CPVulkan_b200.json")).read()
    assert "library_path" in manifest and "1.1.121" in manifest


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less container")
    lib = capi.load_cuda()
    dev = C.c_void_p()
    rc = lib.cpvk_cuda_device_create(0, C.byref(dev))
    assert rc == capi.E_NO_DEVICE and not dev.value
    assert b"no CUDA device" in lib.cpvk_cuda_last_error()
