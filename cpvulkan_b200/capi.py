"""ctypes mirror of include/cpvk_cuda.h (the C ABI of the sm_100a draw path) and library loaders.

The structures below must stay field-for-field identical to the header; tests/test_abi.py checks sizes
against the values the C compiler reports (cpvk_cuda_abi_sizeof).
"""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC_BUILD = os.path.join(ROOT, "cpvulkan_b200", "csrc", "build")
ORACLE_BUILD = os.path.join(ROOT, "oracle", "build")

MAX_VERTEX_BINDINGS = 16
MAX_VERTEX_ATTRIBUTES = 16
MAX_COLOR_ATTACHMENTS = 8
MAX_DESCRIPTORS = 16
MAX_MIP_LEVELS = 13
MAX_PUSH_CONSTANT_BYTES = 128
MAX_SPEC_ENTRIES = 16

E_UNSUPPORTED, E_SPIRV, E_COMPILE, E_ARGUMENT, E_NO_DEVICE = -1, -2, -3, -4, -5
DESC_NONE, DESC_BUFFER, DESC_IMAGE, DESC_TEXEL_BUFFER, DESC_SAMPLER = 0, 1, 2, 3, 4

u32, i32, u64, f32 = C.c_uint32, C.c_int32, C.c_uint64, C.c_float


class VertexBinding(C.Structure):
    _fields_ = [("binding", u32), ("stride", u32), ("inputRate", u32)]


class VertexAttribute(C.Structure):
    _fields_ = [("location", u32), ("binding", u32), ("format", u32), ("offset", u32)]


class StencilOpState(C.Structure):
    _fields_ = [("failOp", u32), ("passOp", u32), ("depthFailOp", u32), ("compareOp", u32),
                ("compareMask", u32), ("writeMask", u32), ("reference", u32)]


class BlendAttachment(C.Structure):
    _fields_ = [("blendEnable", u32), ("srcColorBlendFactor", u32), ("dstColorBlendFactor", u32),
                ("colorBlendOp", u32), ("srcAlphaBlendFactor", u32), ("dstAlphaBlendFactor", u32),
                ("alphaBlendOp", u32), ("colorWriteMask", u32)]


class SpecEntry(C.Structure):
    _fields_ = [("constantId", u32), ("value", u32)]


class ShaderStage(C.Structure):
    _fields_ = [("spirv", C.POINTER(u32)), ("wordCount", C.c_size_t), ("entryPoint", C.c_char_p),
                ("specCount", u32), ("spec", SpecEntry * MAX_SPEC_ENTRIES)]


class PipelineDesc(C.Structure):
    _fields_ = [
        ("vertex", ShaderStage), ("fragment", ShaderStage),
        ("bindingCount", u32), ("bindings", VertexBinding * MAX_VERTEX_BINDINGS),
        ("attributeCount", u32), ("attributes", VertexAttribute * MAX_VERTEX_ATTRIBUTES),
        ("topology", u32), ("primitiveRestartEnable", u32),
        ("depthClampEnable", u32), ("rasterizerDiscardEnable", u32),
        ("polygonMode", u32), ("cullMode", u32), ("frontFace", u32),
        ("depthBiasEnable", u32), ("lineWidth", f32),
        ("rasterizationSamples", u32),
        ("depthTestEnable", u32), ("depthWriteEnable", u32), ("depthCompareOp", u32),
        ("depthBoundsTestEnable", u32), ("stencilTestEnable", u32),
        ("front", StencilOpState), ("back", StencilOpState),
        ("minDepthBounds", f32), ("maxDepthBounds", f32),
        ("logicOpEnable", u32), ("colorAttachmentCount", u32),
        ("colorFormats", u32 * MAX_COLOR_ATTACHMENTS),
        ("blend", BlendAttachment * MAX_COLOR_ATTACHMENTS),
        ("blendConstants", f32 * 4),
        ("depthStencilFormat", u32),
        ("dynamicViewport", u32),
    ]


class Viewport(C.Structure):
    _fields_ = [("x", f32), ("y", f32), ("width", f32), ("height", f32), ("minDepth", f32), ("maxDepth", f32)]


class Attachment(C.Structure):
    _fields_ = [("address", u64), ("width", u32), ("height", u32), ("rowPitch", u32), ("format", u32)]


class MipLevel(C.Structure):
    _fields_ = [("address", u64), ("width", u32), ("height", u32), ("depth", u32), ("pad", u32)]


class Sampler(C.Structure):
    _fields_ = [("magFilter", u32), ("minFilter", u32), ("mipmapMode", u32),
                ("addressModeU", u32), ("addressModeV", u32), ("addressModeW", u32),
                ("mipLodBias", f32), ("anisotropyEnable", u32), ("compareEnable", u32), ("compareOp", u32),
                ("minLod", f32), ("maxLod", f32), ("borderColor", u32), ("unnormalizedCoordinates", u32),
                ("flags", u32), ("reductionMode", u32)]


class Descriptor(C.Structure):
    _fields_ = [("set", u32), ("binding", u32), ("arrayElement", u32), ("type", u32),
                ("address", u64), ("range", u64),
                ("format", u32), ("dimensions", u32), ("levelCount", u32), ("swizzle", u32 * 4),
                ("levels", MipLevel * MAX_MIP_LEVELS), ("sampler", Sampler)]


MAX_MIRRORS = 7


class DrawState(C.Structure):
    _fields_ = [
        ("pipeline", C.c_void_p), ("viewport", Viewport),
        ("vertexBuffers", u64 * MAX_VERTEX_BINDINGS),
        ("indexBuffer", u64), ("indexStride", u32),
        ("count", u32), ("instanceCount", u32), ("first", u32), ("vertexOffset", i32), ("firstInstance", u32),
        ("descriptorCount", u32), ("descriptors", Descriptor * MAX_DESCRIPTORS),
        ("pushConstantSize", u32), ("pushConstants", C.c_uint8 * MAX_PUSH_CONSTANT_BYTES),
        ("color", Attachment * MAX_COLOR_ATTACHMENTS), ("depthStencil", Attachment),
        ("bandY0", u32), ("bandY1", u32),
        ("mirrorCount", u32), ("mirrorPad", u32), ("mirrorColor0", u64 * MAX_MIRRORS),
    ]


class DrawStats(C.Structure):
    _fields_ = [("primitives", u64), ("fragmentsCovered", u64), ("fragmentsWritten", u64), ("binEntries", u64),
                ("msVertex", f32), ("msSetup", f32), ("msBin", f32), ("msRaster", f32), ("msTotal", f32)]


class ClearDepthStencil(C.Structure):
    _fields_ = [("depth", f32), ("stencil", u32)]


class ClearValue(C.Union):
    _fields_ = [("float32", f32 * 4), ("int32", i32 * 4), ("uint32", u32 * 4), ("depthStencil", ClearDepthStencil)]


class Blit(C.Structure):
    _fields_ = [("src", Attachment), ("dst", Attachment),
                ("srcX0", i32), ("srcY0", i32), ("srcX1", i32), ("srcY1", i32),
                ("dstX0", i32), ("dstY0", i32), ("dstX1", i32), ("dstY1", i32), ("filter", u32)]


ABI_STRUCTS = {
    "CpvkVertexBinding": VertexBinding, "CpvkVertexAttribute": VertexAttribute, "CpvkStencilOpState": StencilOpState,
    "CpvkBlendAttachment": BlendAttachment, "CpvkSpecEntry": SpecEntry, "CpvkShaderStage": ShaderStage,
    "CpvkPipelineDesc": PipelineDesc, "CpvkViewport": Viewport, "CpvkAttachment": Attachment,
    "CpvkMipLevel": MipLevel, "CpvkSampler": Sampler, "CpvkDescriptor": Descriptor, "CpvkDrawState": DrawState,
    "CpvkDrawStats": DrawStats, "CpvkClearValue": ClearValue, "CpvkBlit": Blit,
}

# Every symbol include/cpvk_cuda.h declares (tests check the built library exports each one).
ABI_SYMBOLS = [
    "cpvk_cuda_abi_version", "cpvk_cuda_last_error", "cpvk_cuda_device_create", "cpvk_cuda_device_destroy",
    "cpvk_cuda_device_set_stream", "cpvk_cuda_device_set_timing", "cpvk_cuda_device_set_stats", "cpvk_cuda_sync",
    "cpvk_cuda_mem_alloc", "cpvk_cuda_mem_free", "cpvk_cuda_mem_upload", "cpvk_cuda_mem_download",
    "cpvk_cuda_pipeline_create", "cpvk_cuda_pipeline_destroy", "cpvk_cuda_pipeline_source",
    "cpvk_cuda_pipeline_cubin", "cpvk_cuda_pipeline_compile_only", "cpvk_cuda_draw", "cpvk_cuda_last_draw_stats",
    "cpvk_cuda_launch_count", "cpvk_cuda_clear", "cpvk_cuda_copy_rows", "cpvk_cuda_blit",
    "cpvk_cuda_flush", "cpvk_cuda_device_set_lazy_clear", "cpvk_cuda_device_set_speculation", "cpvk_cuda_device_set_overlap", "cpvk_cuda_mem_download_async",
    "cpvk_cuda_abi_sizeof",
    "cpvk_cuda_device_create_group", "cpvk_cuda_group_size", "cpvk_cuda_gather", "cpvk_cuda_mem_export", "cpvk_cuda_mem_import", "cpvk_cuda_mem_unimport", "cpvk_cuda_peer_barrier", "cpvk_cuda_selftest_div",
]


class LibraryMissing(RuntimeError):
    pass


_cuda_lib = None
_oracle_lib = None


def cuda_library_path():
    return os.path.join(CSRC_BUILD, "libcpvk_cuda.so")


def load_cuda():
    """Load libcpvk_cuda.so (the product). Fails loudly when it has not been built: there is no fallback."""
    global _cuda_lib
    if _cuda_lib is not None:
        return _cuda_lib
    path = cuda_library_path()
    if not os.path.exists(path):
        raise LibraryMissing("%s not built; run `python -m cpvulkan_b200.build`" % path)
    lib = C.CDLL(path)
    vp, pvp = C.c_void_p, C.POINTER(C.c_void_p)
    lib.cpvk_cuda_abi_version.restype = C.c_int
    lib.cpvk_cuda_last_error.restype = C.c_char_p
    lib.cpvk_cuda_device_create.argtypes = [C.c_int, pvp]
    lib.cpvk_cuda_device_destroy.argtypes = [vp]
    lib.cpvk_cuda_device_destroy.restype = None
    lib.cpvk_cuda_device_set_stream.argtypes = [vp, vp]
    lib.cpvk_cuda_device_set_timing.argtypes = [vp, C.c_int]
    lib.cpvk_cuda_device_set_stats.argtypes = [vp, C.c_int]
    lib.cpvk_cuda_sync.argtypes = [vp]
    lib.cpvk_cuda_mem_alloc.argtypes = [vp, C.c_size_t, C.POINTER(u64), pvp]
    lib.cpvk_cuda_mem_free.argtypes = [vp, u64]
    lib.cpvk_cuda_mem_upload.argtypes = [vp, u64, vp, C.c_size_t]
    lib.cpvk_cuda_mem_download.argtypes = [vp, vp, u64, C.c_size_t]
    lib.cpvk_cuda_pipeline_create.argtypes = [vp, C.POINTER(PipelineDesc), pvp]
    lib.cpvk_cuda_pipeline_destroy.argtypes = [vp, vp]
    lib.cpvk_cuda_pipeline_destroy.restype = None
    lib.cpvk_cuda_pipeline_source.argtypes = [vp]
    lib.cpvk_cuda_pipeline_source.restype = C.c_char_p
    lib.cpvk_cuda_pipeline_cubin.argtypes = [vp, C.POINTER(C.c_size_t)]
    lib.cpvk_cuda_pipeline_cubin.restype = vp
    lib.cpvk_cuda_pipeline_compile_only.argtypes = [C.POINTER(PipelineDesc), pvp]
    lib.cpvk_cuda_draw.argtypes = [vp, C.POINTER(DrawState)]
    lib.cpvk_cuda_last_draw_stats.argtypes = [vp, C.POINTER(DrawStats)]
    lib.cpvk_cuda_launch_count.argtypes = [vp]
    lib.cpvk_cuda_launch_count.restype = u64
    lib.cpvk_cuda_clear.argtypes = [vp, C.POINTER(Attachment), C.POINTER(ClearValue), C.c_int]
    lib.cpvk_cuda_flush.argtypes = [vp]
    lib.cpvk_cuda_mem_download_async.argtypes = [vp, vp, u64, C.c_size_t]
    lib.cpvk_cuda_device_set_lazy_clear.argtypes = [vp, C.c_int]
    lib.cpvk_cuda_device_set_speculation.argtypes = [vp, C.c_int]
    lib.cpvk_cuda_device_set_overlap.argtypes = [vp, C.c_int]
    lib.cpvk_cuda_copy_rows.argtypes = [vp, u64, u32, u64, u32, u32, u32]
    lib.cpvk_cuda_blit.argtypes = [vp, C.POINTER(Blit)]
    lib.cpvk_cuda_abi_sizeof.argtypes = [C.c_char_p]
    lib.cpvk_cuda_abi_sizeof.restype = C.c_size_t
    lib.cpvk_cuda_device_create_group.argtypes = [C.POINTER(C.c_int), u32, pvp]
    lib.cpvk_cuda_group_size.argtypes = [vp]
    lib.cpvk_cuda_group_size.restype = u32
    lib.cpvk_cuda_gather.argtypes = [vp, C.POINTER(Attachment), u32]
    lib.cpvk_cuda_mem_export.argtypes = [vp, u64, vp]
    lib.cpvk_cuda_mem_import.argtypes = [vp, vp, C.POINTER(u64)]
    lib.cpvk_cuda_mem_unimport.argtypes = [vp, u64]
    lib.cpvk_cuda_peer_barrier.argtypes = [vp, C.POINTER(u64), u32, u32, u32]
    lib.cpvk_cuda_selftest_div.argtypes = [vp, u64, u64, u32, u64, u64]
    _cuda_lib = lib
    return lib


def load_oracle():
    """Load the CPU oracle. Test infrastructure only: tests/, smoke() and bench.py's CPU legs."""
    global _oracle_lib
    if _oracle_lib is not None:
        return _oracle_lib
    path = os.path.join(ORACLE_BUILD, "libcpvk_oracle.so")
    if not os.path.exists(path):
        raise LibraryMissing("%s not built; run `python -m cpvulkan_b200.build --oracle`" % path)
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.cpvk_oracle_last_error.restype = C.c_char_p
    lib.cpvk_oracle_draw.argtypes = [C.POINTER(PipelineDesc), C.POINTER(DrawState), C.POINTER(DrawStats)]
    lib.cpvk_oracle_draw_window.argtypes = [C.POINTER(PipelineDesc), C.POINTER(DrawState), i32, i32, i32, i32, C.POINTER(DrawStats)]
    lib.cpvk_oracle_clear.argtypes = [C.POINTER(Attachment), C.POINTER(ClearValue), C.c_int]
    lib.cpvk_oracle_copy_rows.argtypes = [u64, u32, u64, u32, u32, u32]
    lib.cpvk_oracle_blit.argtypes = [C.POINTER(Blit)]
    lib.cpvk_oracle_input_assembly.argtypes = [C.POINTER(DrawState), vp]
    lib.cpvk_oracle_fragment_inputs.argtypes = [vp, u32, vp, u32]
    lib.cpvk_oracle_blit_window.argtypes = [C.POINTER(Blit), i32, i32, i32, i32]
    lib.cpvk_oracle_format_info.argtypes = [u32, C.POINTER(u32 * 4)]
    lib.cpvk_oracle_pack_f32.argtypes = [u32, vp, u32, vp]
    lib.cpvk_oracle_pack_f32.restype = None
    lib.cpvk_oracle_unpack_f32.argtypes = [u32, vp, u32, vp]
    lib.cpvk_oracle_unpack_f32.restype = None
    lib.cpvk_oracle_pack_depth.argtypes = [u32, vp, vp, u32, vp]
    lib.cpvk_oracle_pack_depth.restype = None
    lib.cpvk_oracle_unpack_depth.argtypes = [u32, vp, u32, vp]
    lib.cpvk_oracle_unpack_depth.restype = None
    lib.cpvk_oracle_sample.argtypes = [C.POINTER(Descriptor), vp, u32, f32, vp]
    lib.cpvk_oracle_sample.restype = None
    lib.cpvk_oracle_fetch.argtypes = [C.POINTER(Descriptor), vp, u32, vp]
    lib.cpvk_oracle_fetch.restype = None
    lib.cpvk_oracle_float_to_half.argtypes = [f32]
    lib.cpvk_oracle_float_to_half.restype = C.c_uint16
    lib.cpvk_oracle_half_to_float.argtypes = [C.c_uint16]
    lib.cpvk_oracle_half_to_float.restype = f32
    lib.cpvk_oracle_format_row.argtypes = [C.c_uint32, C.c_void_p]
    lib.cpvk_oracle_format_row.restype = C.c_int
    lib.cpvk_oracle_image_layout.argtypes = [C.c_uint32] * 6 + [C.c_void_p]
    lib.cpvk_oracle_image_layout.restype = C.c_int
    lib.cpvk_oracle_pixel_offset.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.c_uint32]
    lib.cpvk_oracle_pixel_offset.restype = C.c_uint64
    _oracle_lib = lib
    return lib
