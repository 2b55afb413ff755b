"""Writes oracle/shim/vulkan/vulkan_core.h: the slice of the public Vulkan 1.1 C API (names and enum values from the
Vulkan specification's registry) that CPVulkanBase/{Base,Config,Formats}.h and CPVulkan/ImageSampler.cpp of the
reference need in order to compile IN PLACE under oracle/Makefile. No Vulkan SDK exists in this image.
TEST INFRASTRUCTURE ONLY. Run once; the generated header is committed."""
import os

sevens = ["UNORM", "SNORM", "USCALED", "SSCALED", "UINT", "SINT"]
names = ["UNDEFINED", "R4G4_UNORM_PACK8", "R4G4B4A4_UNORM_PACK16", "B4G4R4A4_UNORM_PACK16", "R5G6B5_UNORM_PACK16",
         "B5G6R5_UNORM_PACK16", "R5G5B5A1_UNORM_PACK16", "B5G5R5A1_UNORM_PACK16", "A1R5G5B5_UNORM_PACK16"]
for fam in ["R8", "R8G8", "R8G8B8", "B8G8R8", "R8G8B8A8", "B8G8R8A8"]:
    names += ["%s_%s" % (fam, k) for k in sevens + ["SRGB"]]
names += ["A8B8G8R8_%s_PACK32" % k for k in sevens + ["SRGB"]]
for fam in ["A2R10G10B10", "A2B10G10R10"]:
    names += ["%s_%s_PACK32" % (fam, k) for k in sevens]
for fam in ["R16", "R16G16", "R16G16B16", "R16G16B16A16"]:
    names += ["%s_%s" % (fam, k) for k in sevens + ["SFLOAT"]]
for fam in ["R32", "R32G32", "R32G32B32", "R32G32B32A32", "R64", "R64G64", "R64G64B64", "R64G64B64A64"]:
    names += ["%s_%s" % (fam, k) for k in ["UINT", "SINT", "SFLOAT"]]
names += ["B10G11R11_UFLOAT_PACK32", "E5B9G9R9_UFLOAT_PACK32", "D16_UNORM", "X8_D24_UNORM_PACK32", "D32_SFLOAT", "S8_UINT",
          "D16_UNORM_S8_UINT", "D24_UNORM_S8_UINT", "D32_SFLOAT_S8_UINT"]
names += ["BC1_RGB_UNORM_BLOCK", "BC1_RGB_SRGB_BLOCK", "BC1_RGBA_UNORM_BLOCK", "BC1_RGBA_SRGB_BLOCK", "BC2_UNORM_BLOCK",
          "BC2_SRGB_BLOCK", "BC3_UNORM_BLOCK", "BC3_SRGB_BLOCK", "BC4_UNORM_BLOCK", "BC4_SNORM_BLOCK", "BC5_UNORM_BLOCK",
          "BC5_SNORM_BLOCK", "BC6H_UFLOAT_BLOCK", "BC6H_SFLOAT_BLOCK", "BC7_UNORM_BLOCK", "BC7_SRGB_BLOCK",
          "ETC2_R8G8B8_UNORM_BLOCK", "ETC2_R8G8B8_SRGB_BLOCK", "ETC2_R8G8B8A1_UNORM_BLOCK", "ETC2_R8G8B8A1_SRGB_BLOCK",
          "ETC2_R8G8B8A8_UNORM_BLOCK", "ETC2_R8G8B8A8_SRGB_BLOCK", "EAC_R11_UNORM_BLOCK", "EAC_R11_SNORM_BLOCK",
          "EAC_R11G11_UNORM_BLOCK", "EAC_R11G11_SNORM_BLOCK"]
for s in ["4x4", "5x4", "5x5", "6x5", "6x6", "8x5", "8x6", "8x8", "10x5", "10x6", "10x8", "10x10", "12x10", "12x12"]:
    names += ["ASTC_%s_UNORM_BLOCK" % s, "ASTC_%s_SRGB_BLOCK" % s]
assert len(names) == 185, len(names)
ycbcr = ["G8B8G8R8_422_UNORM", "B8G8R8G8_422_UNORM", "G8_B8_R8_3PLANE_420_UNORM", "G8_B8R8_2PLANE_420_UNORM",
         "G8_B8_R8_3PLANE_422_UNORM", "G8_B8R8_2PLANE_422_UNORM", "G8_B8_R8_3PLANE_444_UNORM"]
for b, x in (("10", "6"), ("12", "4")):
    r = "R%sX%s" % (b, x); g = "G%sX%s" % (b, x); bl = "B%sX%s" % (b, x); a = "A%sX%s" % (b, x)
    ycbcr += [r + "_UNORM_PACK16", r + g + "_UNORM_2PACK16", r + g + bl + a + "_UNORM_4PACK16",
              g + bl + g + r + "_422_UNORM_4PACK16", bl + g + r + g + "_422_UNORM_4PACK16"]
    for planes, sub in (("3", "420"), ("2", "420"), ("3", "422"), ("2", "422"), ("3", "444")):
        ycbcr.append(("%s_%s_%s_3PLANE_%s_UNORM_3PACK16" % (g, bl, r, sub)) if planes == "3" else ("%s_%s%s_2PLANE_%s_UNORM_3PACK16" % (g, bl, r, sub)))
ycbcr += ["G16B16G16R16_422_UNORM", "B16G16R16G16_422_UNORM", "G16_B16_R16_3PLANE_420_UNORM", "G16_B16R16_2PLANE_420_UNORM",
          "G16_B16_R16_3PLANE_422_UNORM", "G16_B16R16_2PLANE_422_UNORM", "G16_B16_R16_3PLANE_444_UNORM"]
assert len(ycbcr) == 34, len(ycbcr)
pvrtc = ["PVRTC1_2BPP_UNORM_BLOCK_IMG", "PVRTC1_4BPP_UNORM_BLOCK_IMG", "PVRTC2_2BPP_UNORM_BLOCK_IMG", "PVRTC2_4BPP_UNORM_BLOCK_IMG",
         "PVRTC1_2BPP_SRGB_BLOCK_IMG", "PVRTC1_4BPP_SRGB_BLOCK_IMG", "PVRTC2_2BPP_SRGB_BLOCK_IMG", "PVRTC2_4BPP_SRGB_BLOCK_IMG"]

handles_d = ["Instance", "PhysicalDevice", "Device", "Queue", "CommandBuffer"]
handles_n = ["Semaphore", "Fence", "DeviceMemory", "Buffer", "Image", "Event", "QueryPool", "BufferView", "ImageView", "ShaderModule",
             "PipelineCache", "PipelineLayout", "RenderPass", "Pipeline", "DescriptorSetLayout", "Sampler", "DescriptorPool",
             "DescriptorSet", "Framebuffer", "CommandPool", "SamplerYcbcrConversion", "DescriptorUpdateTemplate"]
feature_bits = [("SAMPLED_IMAGE_BIT", 0x1), ("STORAGE_IMAGE_BIT", 0x2), ("STORAGE_IMAGE_ATOMIC_BIT", 0x4), ("UNIFORM_TEXEL_BUFFER_BIT", 0x8),
                ("STORAGE_TEXEL_BUFFER_BIT", 0x10), ("STORAGE_TEXEL_BUFFER_ATOMIC_BIT", 0x20), ("VERTEX_BUFFER_BIT", 0x40),
                ("COLOR_ATTACHMENT_BIT", 0x80), ("COLOR_ATTACHMENT_BLEND_BIT", 0x100), ("DEPTH_STENCIL_ATTACHMENT_BIT", 0x200),
                ("BLIT_SRC_BIT", 0x400), ("BLIT_DST_BIT", 0x800), ("SAMPLED_IMAGE_FILTER_LINEAR_BIT", 0x1000),
                ("SAMPLED_IMAGE_FILTER_CUBIC_BIT_IMG", 0x2000), ("TRANSFER_SRC_BIT", 0x4000), ("TRANSFER_DST_BIT", 0x8000),
                ("SAMPLED_IMAGE_FILTER_MINMAX_BIT_EXT", 0x10000), ("MIDPOINT_CHROMA_SAMPLES_BIT", 0x20000),
                ("SAMPLED_IMAGE_YCBCR_CONVERSION_LINEAR_FILTER_BIT", 0x40000),
                ("SAMPLED_IMAGE_YCBCR_CONVERSION_SEPARATE_RECONSTRUCTION_FILTER_BIT", 0x80000),
                ("SAMPLED_IMAGE_YCBCR_CONVERSION_CHROMA_RECONSTRUCTION_EXPLICIT_BIT", 0x100000),
                ("SAMPLED_IMAGE_YCBCR_CONVERSION_CHROMA_RECONSTRUCTION_EXPLICIT_FORCEABLE_BIT", 0x200000), ("DISJOINT_BIT", 0x400000),
                ("COSITED_CHROMA_SAMPLES_BIT", 0x800000), ("FRAGMENT_DENSITY_MAP_BIT_EXT", 0x1000000)]

o = []
o.append("// GENERATED by oracle/shim/gen_vulkan_shim.py — a slice of the public Vulkan 1.1 C API (specification names and values)\n"
         "// standing in for the absent Vulkan SDK so that reference sources compile in place for oracle/_ref. TEST INFRASTRUCTURE ONLY.\n"
         "#pragma once\n#include <stdint.h>\n#include <stddef.h>\n#define VK_VERSION_1_0 1\n#define VK_VERSION_1_1 1\n"
         "#define VK_MAKE_VERSION(major, minor, patch) (((major) << 22) | ((minor) << 12) | (patch))\n"
         "#define VK_API_VERSION_1_0 VK_MAKE_VERSION(1, 0, 0)\n#define VK_API_VERSION_1_1 VK_MAKE_VERSION(1, 1, 0)\n"
         "#define VK_UUID_SIZE 16\n#define VKAPI_ATTR\n#define VKAPI_CALL\n#define VKAPI_PTR\n"
         "typedef uint32_t VkFlags;\ntypedef uint32_t VkBool32;\ntypedef uint64_t VkDeviceSize;\ntypedef uint32_t VkSampleMask;\n")
for h in handles_d + handles_n:
    o.append("typedef struct Vk%s_T* Vk%s;\n" % (h, h))
o.append("typedef enum VkFormat {\n")
for i, n in enumerate(names):
    o.append("    VK_FORMAT_%s = %d,\n" % (n, i))
for i, n in enumerate(ycbcr):
    o.append("    VK_FORMAT_%s = %d,\n" % (n, 1000156000 + i))
for i, n in enumerate(pvrtc):
    o.append("    VK_FORMAT_%s = %d,\n" % (n, 1000054000 + i))
o.append("    VK_FORMAT_BEGIN_RANGE = VK_FORMAT_UNDEFINED,\n    VK_FORMAT_END_RANGE = VK_FORMAT_ASTC_12x12_SRGB_BLOCK,\n"
         "    VK_FORMAT_RANGE_SIZE = (VK_FORMAT_ASTC_12x12_SRGB_BLOCK - VK_FORMAT_UNDEFINED + 1),\n    VK_FORMAT_MAX_ENUM = 0x7FFFFFFF\n} VkFormat;\n")
o.append("typedef enum VkFormatFeatureFlagBits {\n")
for n, v in feature_bits:
    o.append("    VK_FORMAT_FEATURE_%s = 0x%08X,\n" % (n, v))
o.append("    VK_FORMAT_FEATURE_FLAG_BITS_MAX_ENUM = 0x7FFFFFFF\n} VkFormatFeatureFlagBits;\ntypedef VkFlags VkFormatFeatureFlags;\n")
o.append("""typedef enum VkColorComponentFlagBits { VK_COLOR_COMPONENT_R_BIT = 1, VK_COLOR_COMPONENT_G_BIT = 2, VK_COLOR_COMPONENT_B_BIT = 4, VK_COLOR_COMPONENT_A_BIT = 8, VK_COLOR_COMPONENT_FLAG_BITS_MAX_ENUM = 0x7FFFFFFF } VkColorComponentFlagBits;
typedef VkFlags VkColorComponentFlags;
typedef enum VkSampleCountFlagBits { VK_SAMPLE_COUNT_1_BIT = 1, VK_SAMPLE_COUNT_2_BIT = 2, VK_SAMPLE_COUNT_4_BIT = 4, VK_SAMPLE_COUNT_8_BIT = 8, VK_SAMPLE_COUNT_16_BIT = 16, VK_SAMPLE_COUNT_32_BIT = 32, VK_SAMPLE_COUNT_64_BIT = 64, VK_SAMPLE_COUNT_FLAG_BITS_MAX_ENUM = 0x7FFFFFFF } VkSampleCountFlagBits;
typedef VkFlags VkSampleCountFlags;
typedef enum VkPhysicalDeviceType { VK_PHYSICAL_DEVICE_TYPE_OTHER = 0, VK_PHYSICAL_DEVICE_TYPE_INTEGRATED_GPU = 1, VK_PHYSICAL_DEVICE_TYPE_DISCRETE_GPU = 2, VK_PHYSICAL_DEVICE_TYPE_VIRTUAL_GPU = 3, VK_PHYSICAL_DEVICE_TYPE_CPU = 4 } VkPhysicalDeviceType;
typedef enum VkSystemAllocationScope { VK_SYSTEM_ALLOCATION_SCOPE_COMMAND = 0, VK_SYSTEM_ALLOCATION_SCOPE_OBJECT = 1, VK_SYSTEM_ALLOCATION_SCOPE_CACHE = 2, VK_SYSTEM_ALLOCATION_SCOPE_DEVICE = 3, VK_SYSTEM_ALLOCATION_SCOPE_INSTANCE = 4 } VkSystemAllocationScope;
typedef enum VkInternalAllocationType { VK_INTERNAL_ALLOCATION_TYPE_EXECUTABLE = 0 } VkInternalAllocationType;
typedef void* (*PFN_vkAllocationFunction)(void*, size_t, size_t, VkSystemAllocationScope);
typedef void* (*PFN_vkReallocationFunction)(void*, void*, size_t, size_t, VkSystemAllocationScope);
typedef void (*PFN_vkFreeFunction)(void*, void*);
typedef void (*PFN_vkInternalAllocationNotification)(void*, size_t, VkInternalAllocationType, VkSystemAllocationScope);
typedef void (*PFN_vkInternalFreeNotification)(void*, size_t, VkInternalAllocationType, VkSystemAllocationScope);
typedef struct VkAllocationCallbacks { void* pUserData; PFN_vkAllocationFunction pfnAllocation; PFN_vkReallocationFunction pfnReallocation; PFN_vkFreeFunction pfnFree; PFN_vkInternalAllocationNotification pfnInternalAllocation; PFN_vkInternalFreeNotification pfnInternalFree; } VkAllocationCallbacks;
typedef enum VkResult { VK_SUCCESS = 0, VK_NOT_READY = 1, VK_TIMEOUT = 2, VK_INCOMPLETE = 5, VK_ERROR_OUT_OF_HOST_MEMORY = -1, VK_ERROR_FEATURE_NOT_PRESENT = -8 } VkResult;
typedef enum VkFilter { VK_FILTER_NEAREST = 0, VK_FILTER_LINEAR = 1, VK_FILTER_CUBIC_IMG = 1000015000 } VkFilter;
typedef enum VkSamplerMipmapMode { VK_SAMPLER_MIPMAP_MODE_NEAREST = 0, VK_SAMPLER_MIPMAP_MODE_LINEAR = 1 } VkSamplerMipmapMode;
typedef enum VkSamplerAddressMode { VK_SAMPLER_ADDRESS_MODE_REPEAT = 0, VK_SAMPLER_ADDRESS_MODE_MIRRORED_REPEAT = 1, VK_SAMPLER_ADDRESS_MODE_CLAMP_TO_EDGE = 2, VK_SAMPLER_ADDRESS_MODE_CLAMP_TO_BORDER = 3, VK_SAMPLER_ADDRESS_MODE_MIRROR_CLAMP_TO_EDGE = 4 } VkSamplerAddressMode;
typedef enum VkBorderColor { VK_BORDER_COLOR_FLOAT_TRANSPARENT_BLACK = 0, VK_BORDER_COLOR_INT_TRANSPARENT_BLACK = 1, VK_BORDER_COLOR_FLOAT_OPAQUE_BLACK = 2, VK_BORDER_COLOR_INT_OPAQUE_BLACK = 3, VK_BORDER_COLOR_FLOAT_OPAQUE_WHITE = 4, VK_BORDER_COLOR_INT_OPAQUE_WHITE = 5 } VkBorderColor;
typedef enum VkCompareOp { VK_COMPARE_OP_NEVER = 0, VK_COMPARE_OP_LESS = 1, VK_COMPARE_OP_EQUAL = 2, VK_COMPARE_OP_LESS_OR_EQUAL = 3, VK_COMPARE_OP_GREATER = 4, VK_COMPARE_OP_NOT_EQUAL = 5, VK_COMPARE_OP_GREATER_OR_EQUAL = 6, VK_COMPARE_OP_ALWAYS = 7 } VkCompareOp;
typedef enum VkSamplerReductionModeEXT { VK_SAMPLER_REDUCTION_MODE_WEIGHTED_AVERAGE_EXT = 0, VK_SAMPLER_REDUCTION_MODE_MIN_EXT = 1, VK_SAMPLER_REDUCTION_MODE_MAX_EXT = 2 } VkSamplerReductionModeEXT;
typedef VkFlags VkSamplerCreateFlags;
typedef enum VkSamplerCreateFlagBits { VK_SAMPLER_CREATE_SUBSAMPLED_BIT_EXT = 1, VK_SAMPLER_CREATE_SUBSAMPLED_COARSE_RECONSTRUCTION_BIT_EXT = 2 } VkSamplerCreateFlagBits;
typedef struct VkSamplerCreateInfo VkSamplerCreateInfo;
typedef enum VkImageType { VK_IMAGE_TYPE_1D = 0, VK_IMAGE_TYPE_2D = 1, VK_IMAGE_TYPE_3D = 2 } VkImageType;
typedef enum VkImageTiling { VK_IMAGE_TILING_OPTIMAL = 0, VK_IMAGE_TILING_LINEAR = 1 } VkImageTiling;
typedef enum VkImageLayout { VK_IMAGE_LAYOUT_UNDEFINED = 0, VK_IMAGE_LAYOUT_GENERAL = 1 } VkImageLayout;
typedef VkFlags VkImageCreateFlags;
typedef VkFlags VkImageUsageFlags;
typedef VkFlags VkImageAspectFlags;
typedef struct VkExtent2D { uint32_t width, height; } VkExtent2D;
typedef struct VkExtent3D { uint32_t width, height, depth; } VkExtent3D;
typedef struct VkOffset2D { int32_t x, y; } VkOffset2D;
typedef struct VkOffset3D { int32_t x, y, z; } VkOffset3D;
typedef struct VkRect2D { VkOffset2D offset; VkExtent2D extent; } VkRect2D;
typedef struct VkViewport { float x, y, width, height, minDepth, maxDepth; } VkViewport;
#define VK_KHR_swapchain 1
typedef struct VkSwapchainKHR_T* VkSwapchainKHR;
typedef struct VkMemoryRequirements { VkDeviceSize size, alignment; uint32_t memoryTypeBits; } VkMemoryRequirements;
typedef struct VkImageSubresource { VkImageAspectFlags aspectMask; uint32_t mipLevel, arrayLayer; } VkImageSubresource;
typedef struct VkSubresourceLayout { VkDeviceSize offset, size, rowPitch, arrayPitch, depthPitch; } VkSubresourceLayout;
typedef struct VkImageCreateInfo VkImageCreateInfo;
typedef union VkClearColorValue { float float32[4]; int32_t int32[4]; uint32_t uint32[4]; } VkClearColorValue;
typedef struct VkClearDepthStencilValue { float depth; uint32_t stencil; } VkClearDepthStencilValue;
typedef union VkClearValue { VkClearColorValue color; VkClearDepthStencilValue depthStencil; } VkClearValue;
""")

# ---- what CPVulkanBase/PipelineState.h and the sliced parts of CPVulkan/CommandBuffer.Draw.cpp need (oracle/_ref/draw_check) ----
advanced_blend = ['ZERO', 'SRC', 'DST', 'SRC_OVER', 'DST_OVER', 'SRC_IN', 'DST_IN', 'SRC_OUT', 'DST_OUT', 'SRC_ATOP', 'DST_ATOP', 'XOR', 'MULTIPLY', 'SCREEN', 'OVERLAY', 'DARKEN', 'LIGHTEN', 'COLORDODGE', 'COLORBURN', 'HARDLIGHT', 'SOFTLIGHT', 'DIFFERENCE', 'EXCLUSION', 'INVERT', 'INVERT_RGB', 'LINEARDODGE', 'LINEARBURN', 'VIVIDLIGHT', 'LINEARLIGHT', 'PINLIGHT', 'HARDMIX', 'HSL_HUE', 'HSL_SATURATION', 'HSL_COLOR', 'HSL_LUMINOSITY', 'PLUS', 'PLUS_CLAMPED', 'PLUS_CLAMPED_ALPHA', 'PLUS_DARKER', 'MINUS', 'MINUS_CLAMPED', 'CONTRAST', 'INVERT_OVG', 'RED', 'GREEN', 'BLUE']
factors = ["ZERO", "ONE", "SRC_COLOR", "ONE_MINUS_SRC_COLOR", "DST_COLOR", "ONE_MINUS_DST_COLOR", "SRC_ALPHA", "ONE_MINUS_SRC_ALPHA",
           "DST_ALPHA", "ONE_MINUS_DST_ALPHA", "CONSTANT_COLOR", "ONE_MINUS_CONSTANT_COLOR", "CONSTANT_ALPHA", "ONE_MINUS_CONSTANT_ALPHA",
           "SRC_ALPHA_SATURATE", "SRC1_COLOR", "ONE_MINUS_SRC1_COLOR", "SRC1_ALPHA", "ONE_MINUS_SRC1_ALPHA"]
topologies = ["POINT_LIST", "LINE_LIST", "LINE_STRIP", "TRIANGLE_LIST", "TRIANGLE_STRIP", "TRIANGLE_FAN", "LINE_LIST_WITH_ADJACENCY",
              "LINE_STRIP_WITH_ADJACENCY", "TRIANGLE_LIST_WITH_ADJACENCY", "TRIANGLE_STRIP_WITH_ADJACENCY", "PATCH_LIST"]
o.append("typedef enum VkBlendFactor {\n" + "".join("    VK_BLEND_FACTOR_%s = %d,\n" % (n, i) for i, n in enumerate(factors)) + "} VkBlendFactor;\n")
o.append("typedef enum VkBlendOp {\n    VK_BLEND_OP_ADD = 0, VK_BLEND_OP_SUBTRACT = 1, VK_BLEND_OP_REVERSE_SUBTRACT = 2, VK_BLEND_OP_MIN = 3, VK_BLEND_OP_MAX = 4,\n"
         + "".join("    VK_BLEND_OP_%s_EXT = %d,\n" % (n, 1000148000 + i) for i, n in enumerate(advanced_blend)) + "} VkBlendOp;\n")
o.append("typedef enum VkPrimitiveTopology {\n" + "".join("    VK_PRIMITIVE_TOPOLOGY_%s = %d,\n" % (n, i) for i, n in enumerate(topologies)) + "} VkPrimitiveTopology;\n")
o.append("""typedef enum VkPolygonMode { VK_POLYGON_MODE_FILL = 0, VK_POLYGON_MODE_LINE = 1, VK_POLYGON_MODE_POINT = 2 } VkPolygonMode;
typedef enum VkFrontFace { VK_FRONT_FACE_COUNTER_CLOCKWISE = 0, VK_FRONT_FACE_CLOCKWISE = 1 } VkFrontFace;
typedef enum VkCullModeFlagBits { VK_CULL_MODE_NONE = 0, VK_CULL_MODE_FRONT_BIT = 1, VK_CULL_MODE_BACK_BIT = 2, VK_CULL_MODE_FRONT_AND_BACK = 3 } VkCullModeFlagBits;
typedef VkFlags VkCullModeFlags;
#define VK_EXT_line_rasterization 1
typedef enum VkLineRasterizationModeEXT { VK_LINE_RASTERIZATION_MODE_DEFAULT_EXT = 0, VK_LINE_RASTERIZATION_MODE_RECTANGULAR_EXT = 1, VK_LINE_RASTERIZATION_MODE_BRESENHAM_EXT = 2, VK_LINE_RASTERIZATION_MODE_RECTANGULAR_SMOOTH_EXT = 3 } VkLineRasterizationModeEXT;
typedef enum VkVertexInputRate { VK_VERTEX_INPUT_RATE_VERTEX = 0, VK_VERTEX_INPUT_RATE_INSTANCE = 1 } VkVertexInputRate;
typedef struct VkVertexInputBindingDescription { uint32_t binding, stride; VkVertexInputRate inputRate; } VkVertexInputBindingDescription;
typedef struct VkVertexInputAttributeDescription { uint32_t location, binding; VkFormat format; uint32_t offset; } VkVertexInputAttributeDescription;
typedef enum VkStencilOp { VK_STENCIL_OP_KEEP = 0, VK_STENCIL_OP_ZERO = 1, VK_STENCIL_OP_REPLACE = 2, VK_STENCIL_OP_INCREMENT_AND_CLAMP = 3, VK_STENCIL_OP_DECREMENT_AND_CLAMP = 4, VK_STENCIL_OP_INVERT = 5, VK_STENCIL_OP_INCREMENT_AND_WRAP = 6, VK_STENCIL_OP_DECREMENT_AND_WRAP = 7 } VkStencilOp;
typedef struct VkStencilOpState { VkStencilOp failOp, passOp, depthFailOp; VkCompareOp compareOp; uint32_t compareMask, writeMask, reference; } VkStencilOpState;
typedef enum VkLogicOp { VK_LOGIC_OP_CLEAR = 0, VK_LOGIC_OP_AND = 1, VK_LOGIC_OP_AND_REVERSE = 2, VK_LOGIC_OP_COPY = 3, VK_LOGIC_OP_AND_INVERTED = 4, VK_LOGIC_OP_NO_OP = 5, VK_LOGIC_OP_XOR = 6, VK_LOGIC_OP_OR = 7, VK_LOGIC_OP_NOR = 8, VK_LOGIC_OP_EQUIVALENT = 9, VK_LOGIC_OP_INVERT = 10, VK_LOGIC_OP_OR_REVERSE = 11, VK_LOGIC_OP_COPY_INVERTED = 12, VK_LOGIC_OP_OR_INVERTED = 13, VK_LOGIC_OP_NAND = 14, VK_LOGIC_OP_SET = 15 } VkLogicOp;
typedef struct VkPipelineColorBlendAttachmentState { VkBool32 blendEnable; VkBlendFactor srcColorBlendFactor, dstColorBlendFactor; VkBlendOp colorBlendOp; VkBlendFactor srcAlphaBlendFactor, dstAlphaBlendFactor; VkBlendOp alphaBlendOp; VkColorComponentFlags colorWriteMask; } VkPipelineColorBlendAttachmentState;
typedef VkFlags VkAttachmentDescriptionFlags;
typedef enum VkAttachmentLoadOp { VK_ATTACHMENT_LOAD_OP_LOAD = 0, VK_ATTACHMENT_LOAD_OP_CLEAR = 1, VK_ATTACHMENT_LOAD_OP_DONT_CARE = 2 } VkAttachmentLoadOp;
typedef enum VkAttachmentStoreOp { VK_ATTACHMENT_STORE_OP_STORE = 0, VK_ATTACHMENT_STORE_OP_DONT_CARE = 1 } VkAttachmentStoreOp;
typedef VkFlags VkSubpassDescriptionFlags;
typedef enum VkPipelineBindPoint { VK_PIPELINE_BIND_POINT_GRAPHICS = 0, VK_PIPELINE_BIND_POINT_COMPUTE = 1 } VkPipelineBindPoint;
typedef VkFlags VkPipelineStageFlags;
typedef VkFlags VkAccessFlags;
typedef VkFlags VkDependencyFlags;
""")

# ---- what the sliced BlitImageCommand::Process of CPVulkan/CommandBuffer.cpp needs (oracle/_ref/blit_check) ----
o.append("""typedef enum VkImageAspectFlagBits { VK_IMAGE_ASPECT_COLOR_BIT = 1, VK_IMAGE_ASPECT_DEPTH_BIT = 2, VK_IMAGE_ASPECT_STENCIL_BIT = 4, VK_IMAGE_ASPECT_METADATA_BIT = 8 } VkImageAspectFlagBits;
typedef struct VkImageSubresourceLayers { VkImageAspectFlags aspectMask; uint32_t mipLevel, baseArrayLayer, layerCount; } VkImageSubresourceLayers;
typedef struct VkImageBlit { VkImageSubresourceLayers srcSubresource; VkOffset3D srcOffsets[2]; VkImageSubresourceLayers dstSubresource; VkOffset3D dstOffsets[2]; } VkImageBlit;
""")
# ---- CPVulkan/Buffer.h (the index buffer bound in draw_check's "ia" mode): names its declarations mention ----
o.append("""typedef VkFlags VkBufferCreateFlags;
typedef VkFlags VkBufferUsageFlags;
typedef struct VkMemoryRequirements2 VkMemoryRequirements2;
typedef struct VkBufferCreateInfo VkBufferCreateInfo;
""")
# ---- CPVulkan/ImageView.h, BufferView.h, DescriptorSet.h and the sliced image functions of GlslFunctions.cpp (oracle/_ref/image_check) ----
o.append("""typedef enum VkComponentSwizzle { VK_COMPONENT_SWIZZLE_IDENTITY = 0, VK_COMPONENT_SWIZZLE_ZERO = 1, VK_COMPONENT_SWIZZLE_ONE = 2, VK_COMPONENT_SWIZZLE_R = 3, VK_COMPONENT_SWIZZLE_G = 4, VK_COMPONENT_SWIZZLE_B = 5, VK_COMPONENT_SWIZZLE_A = 6 } VkComponentSwizzle;
typedef struct VkComponentMapping { VkComponentSwizzle r, g, b, a; } VkComponentMapping;
typedef struct VkImageSubresourceRange { VkImageAspectFlags aspectMask; uint32_t baseMipLevel, levelCount, baseArrayLayer, layerCount; } VkImageSubresourceRange;
typedef enum VkImageViewType { VK_IMAGE_VIEW_TYPE_1D = 0, VK_IMAGE_VIEW_TYPE_2D = 1, VK_IMAGE_VIEW_TYPE_3D = 2, VK_IMAGE_VIEW_TYPE_CUBE = 3, VK_IMAGE_VIEW_TYPE_1D_ARRAY = 4, VK_IMAGE_VIEW_TYPE_2D_ARRAY = 5, VK_IMAGE_VIEW_TYPE_CUBE_ARRAY = 6 } VkImageViewType;
#define VK_REMAINING_MIP_LEVELS (~0U)
#define VK_REMAINING_ARRAY_LAYERS (~0U)
typedef struct VkImageViewCreateInfo VkImageViewCreateInfo;
typedef struct VkBufferViewCreateInfo VkBufferViewCreateInfo;
typedef enum VkDescriptorType { VK_DESCRIPTOR_TYPE_SAMPLER = 0, VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER = 1, VK_DESCRIPTOR_TYPE_SAMPLED_IMAGE = 2, VK_DESCRIPTOR_TYPE_STORAGE_IMAGE = 3, VK_DESCRIPTOR_TYPE_UNIFORM_TEXEL_BUFFER = 4, VK_DESCRIPTOR_TYPE_STORAGE_TEXEL_BUFFER = 5, VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER = 6, VK_DESCRIPTOR_TYPE_STORAGE_BUFFER = 7, VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER_DYNAMIC = 8, VK_DESCRIPTOR_TYPE_STORAGE_BUFFER_DYNAMIC = 9, VK_DESCRIPTOR_TYPE_INPUT_ATTACHMENT = 10 } VkDescriptorType;
typedef struct VkDescriptorBufferInfo { VkBuffer buffer; VkDeviceSize offset, range; } VkDescriptorBufferInfo;
typedef struct VkWriteDescriptorSet VkWriteDescriptorSet;
typedef struct VkCopyDescriptorSet VkCopyDescriptorSet;
""")
here = os.path.dirname(os.path.abspath(__file__))
open(os.path.join(here, "vulkan", "vulkan_core.h"), "w").write("".join(o))
open(os.path.join(here, "vulkan", "vulkan.h"), "w").write("#pragma once\n#include \"vulkan_core.h\"\n")
open(os.path.join(here, "vulkan", "vk_platform.h"), "w").write("#pragma once\n#include <stdint.h>\n#include <stddef.h>\n")
open(os.path.join(here, "vulkan", "vk_icd.h"), "w").write("#pragma once\n#include \"vulkan_core.h\"\n// loader/ICD interface constant (public LunarG loader header value)\n#define ICD_LOADER_MAGIC 0x01CDC0DE\n")
print("formats:", len(names) + len(ycbcr) + len(pvrtc))
