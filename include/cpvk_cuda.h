/*
 * cpvk_cuda.h — the thin C ABI between the Vulkan ICD host side and the sm_100a draw path.
 *
 * This is the INNER drop-in boundary of SURVEY.md §8(b): it replaces the two reference call sites
 *   DrawCommand::Process        (CPVulkan/CommandBuffer.Draw.cpp:1777-1804)
 *   DrawIndexedCommand::Process (CPVulkan/CommandBuffer.Draw.cpp:1840-1864)
 * plus the pipeline compile they depend on
 *   CompileVertexPipeline / CompileFragmentPipeline (LLVMRuntime/PipelineCompiler.cpp:1801-1832),
 * the render-pass clear (CPVulkan/CommandBuffer.cpp:591-640 -> Draw.cpp:117-149) and the transfer commands
 * (CPVulkan/CommandBuffer.Copy.cpp, CommandBuffer.cpp:57-232).
 *
 * Plain C: pointers, sizes and POD structs only. Every address field (uint64_t) is a *device* address
 * when handed to libcpvk_cuda.so; the CPU oracle (oracle/) re-uses these PODs with *host* addresses.
 * All Vulkan enums (VkFormat, VkCompareOp, ...) are passed as their numeric Vulkan values.
 *
 * Return convention: 0 = success, >0 = cudaError_t / CUresult of the failing call (the ICD maps it to
 * VK_ERROR_DEVICE_LOST), <0 = CPVK_E_* below. Unsupported pipeline state is reported at pipeline creation
 * (CPVK_E_UNSUPPORTED); the ICD turns that into abort(), which is the reference convention
 * (CPVulkanBase/Base.h:73-74 TODO_ERROR).
 */
#ifndef CPVK_CUDA_H
#define CPVK_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPVK_ABI_VERSION 1

#define CPVK_E_UNSUPPORTED (-1) /* state the reference aborts on, or outside the built subset          */
#define CPVK_E_SPIRV (-2)       /* malformed / untranslatable SPIR-V                                   */
#define CPVK_E_COMPILE (-3)     /* NVRTC or nvJitLink failed (see cpvk_cuda_last_error)                */
#define CPVK_E_ARGUMENT (-4)
#define CPVK_E_NO_DEVICE (-5)   /* CUDA driver / device missing: the product path has no CPU fallback */

#define CPVK_MAX_VERTEX_BINDINGS 16   /* MAX_VERTEX_INPUT_BINDINGS, CPVulkanBase/Config.h            */
#define CPVK_MAX_VERTEX_ATTRIBUTES 16
#define CPVK_MAX_COLOR_ATTACHMENTS 8  /* MAX_FRAGMENT_OUTPUT_ATTACHMENTS, CPVulkanBase/Config.h:144  */
#define CPVK_MAX_DESCRIPTORS 16
#define CPVK_MAX_MIRRORS 7 /* the other GPUs of one 8-GPU box */
#define CPVK_MAX_MIP_LEVELS 13        /* MAX_MIP_LEVELS = clog2(4096), CPVulkanBase/Formats.h:105-115 */
#define CPVK_MAX_PUSH_CONSTANT_BYTES 128
#define CPVK_MAX_SPEC_ENTRIES 16

typedef struct CpvkDevice CpvkDevice;     /* one per GPU: stream, scratch arenas, module cache */
typedef struct CpvkPipeline CpvkPipeline; /* one per VkPipeline: linked cubin + layout tables  */

/* ---- pipeline description (== the state GraphicsPipeline::Create parses, CPVulkan/Pipeline.cpp:599-714) ---- */

typedef struct CpvkVertexBinding {
    uint32_t binding;
    uint32_t stride;
    uint32_t inputRate; /* VkVertexInputRate: 0 vertex, 1 instance */
} CpvkVertexBinding;

typedef struct CpvkVertexAttribute {
    uint32_t location;
    uint32_t binding;
    uint32_t format; /* VkFormat */
    uint32_t offset;
} CpvkVertexAttribute;

typedef struct CpvkStencilOpState { /* VkStencilOpState */
    uint32_t failOp, passOp, depthFailOp, compareOp;
    uint32_t compareMask, writeMask, reference;
} CpvkStencilOpState;

typedef struct CpvkBlendAttachment { /* VkPipelineColorBlendAttachmentState */
    uint32_t blendEnable;
    uint32_t srcColorBlendFactor, dstColorBlendFactor, colorBlendOp;
    uint32_t srcAlphaBlendFactor, dstAlphaBlendFactor, alphaBlendOp;
    uint32_t colorWriteMask;
} CpvkBlendAttachment;

typedef struct CpvkSpecEntry { /* VkSpecializationMapEntry resolved to a 32-bit value */
    uint32_t constantId;
    uint32_t value;
} CpvkSpecEntry;

typedef struct CpvkShaderStage {
    const uint32_t* spirv; /* host pointer, SPIR-V words as given to vkCreateShaderModule */
    size_t wordCount;
    const char* entryPoint;
    uint32_t specCount;
    CpvkSpecEntry spec[CPVK_MAX_SPEC_ENTRIES];
} CpvkShaderStage;

typedef struct CpvkPipelineDesc {
    CpvkShaderStage vertex;
    CpvkShaderStage fragment; /* spirv == NULL: no fragment stage (vkCmdDraw then skips raster, Draw.cpp:1799-1802) */

    uint32_t bindingCount;
    CpvkVertexBinding bindings[CPVK_MAX_VERTEX_BINDINGS];
    uint32_t attributeCount;
    CpvkVertexAttribute attributes[CPVK_MAX_VERTEX_ATTRIBUTES];

    uint32_t topology; /* VkPrimitiveTopology */
    uint32_t primitiveRestartEnable;

    uint32_t depthClampEnable, rasterizerDiscardEnable;
    uint32_t polygonMode, cullMode, frontFace;
    uint32_t depthBiasEnable;
    float lineWidth;

    uint32_t rasterizationSamples; /* must be 1 */

    uint32_t depthTestEnable, depthWriteEnable, depthCompareOp;
    uint32_t depthBoundsTestEnable, stencilTestEnable;
    CpvkStencilOpState front, back;
    float minDepthBounds, maxDepthBounds;

    uint32_t logicOpEnable;
    uint32_t colorAttachmentCount;                               /* subpass colour attachment count */
    uint32_t colorFormats[CPVK_MAX_COLOR_ATTACHMENTS];           /* VkFormat, 0 = VK_ATTACHMENT_UNUSED */
    CpvkBlendAttachment blend[CPVK_MAX_COLOR_ATTACHMENTS];
    float blendConstants[4];
    uint32_t depthStencilFormat;                                 /* VkFormat, 0 = none */

    uint32_t dynamicViewport; /* informational: the viewport always arrives through CpvkDrawState */
} CpvkPipelineDesc;

/* ---- per-draw state (== what DeviceState.graphicsPipelineState holds at Process() time, DeviceState.h:49-121) ---- */

typedef struct CpvkViewport { /* VkViewport */
    float x, y, width, height, minDepth, maxDepth;
} CpvkViewport;

/* One 2-D subresource of a linear image: Stride = texel * width (CPVulkanBase/Formats.cpp:455-483). */
typedef struct CpvkAttachment {
    uint64_t address; /* address of texel (0,0) of the view's baseMipLevel / baseArrayLayer */
    uint32_t width, height;
    uint32_t rowPitch; /* bytes */
    uint32_t format;   /* VkFormat; must equal the format baked into the pipeline */
} CpvkAttachment;

typedef struct CpvkMipLevel {
    uint64_t address; /* address of texel (0,0,0) of this level for the view's base layer */
    uint32_t width, height, depth;
    uint32_t pad;
} CpvkMipLevel;

typedef struct CpvkSampler { /* the VkSamplerCreateInfo fields Sampler.h keeps */
    uint32_t magFilter, minFilter, mipmapMode;
    uint32_t addressModeU, addressModeV, addressModeW;
    float mipLodBias;
    uint32_t anisotropyEnable;
    uint32_t compareEnable, compareOp;
    float minLod, maxLod;
    uint32_t borderColor;
    uint32_t unnormalizedCoordinates;
    uint32_t flags;
    uint32_t reductionMode;
} CpvkSampler;

enum {
    CPVK_DESC_NONE = 0,
    CPVK_DESC_BUFFER = 1,        /* uniform / storage buffer (LoadUniforms, Draw.cpp:379-396) */
    CPVK_DESC_IMAGE = 2,         /* ImageDescriptorType::Image (+ sampler when combined)      */
    CPVK_DESC_TEXEL_BUFFER = 3,  /* ImageDescriptorType::Buffer                                */
    CPVK_DESC_SAMPLER = 4        /* a sampler object on its own (VK_DESCRIPTOR_TYPE_SAMPLER): only `sampler` is read; OpSampledImage
                                    combines it with a CPVK_DESC_IMAGE (ImageCombine, GlslFunctions.cpp:812-820) */
};

typedef struct CpvkDescriptor {
    uint32_t set, binding, arrayElement;
    uint32_t type; /* CPVK_DESC_* */
    /* CPVK_DESC_BUFFER: address = buffer + offset (+ dynamic offset), range in bytes.
       CPVK_DESC_TEXEL_BUFFER: address = buffer + view offset, range = view range, format = view format. */
    uint64_t address;
    uint64_t range;
    /* CPVK_DESC_IMAGE */
    uint32_t format;       /* view format */
    uint32_t dimensions;   /* 1, 2 or 3 */
    uint32_t levelCount;   /* levels visible through the view (>= 1) */
    uint32_t swizzle[4];   /* VkComponentSwizzle r,g,b,a */
    CpvkMipLevel levels[CPVK_MAX_MIP_LEVELS];
    CpvkSampler sampler;
} CpvkDescriptor;

typedef struct CpvkDrawState {
    const CpvkPipeline* pipeline;
    CpvkViewport viewport;

    uint64_t vertexBuffers[CPVK_MAX_VERTEX_BINDINGS]; /* buffer address + bind offset (Binding.cpp:180-199) */

    uint64_t indexBuffer;  /* buffer address + bind offset; ignored when indexStride == 0 */
    uint32_t indexStride;  /* 0 = vkCmdDraw, 1/2/4 = vkCmdDrawIndexed with u8/u16/u32 (Binding.cpp:113-134) */

    uint32_t count;         /* vertexCount or indexCount */
    uint32_t instanceCount;
    uint32_t first;         /* firstVertex or firstIndex */
    int32_t vertexOffset;
    uint32_t firstInstance;

    uint32_t descriptorCount;
    CpvkDescriptor descriptors[CPVK_MAX_DESCRIPTORS];

    uint32_t pushConstantSize;
    uint8_t pushConstants[CPVK_MAX_PUSH_CONSTANT_BYTES];

    CpvkAttachment color[CPVK_MAX_COLOR_ATTACHMENTS]; /* address == 0: attachment unused / null view */
    CpvkAttachment depthStencil;                      /* address == 0: none                          */

    /* Sort-first band owned by this GPU: rows [bandY0, bandY1). 0,0 = whole target. */
    uint32_t bandY0, bandY1;
    /* Gather fused into rasterisation: device addresses of the other GPUs' copies of color[0] (peer-mapped over
       NVLink, same layout as color[0]). Every finished tile of this GPU's band is stored into each of them as well,
       so the exchange of §8(e) overlaps the raster kernel tile by tile and no collective follows the draw; the caller
       only orders the GPUs with a barrier before anyone reads a frame. 0 = off (gather the bands afterwards). */
    uint32_t mirrorCount;
    uint32_t mirrorPad;
    uint64_t mirrorColor0[CPVK_MAX_MIRRORS];
} CpvkDrawState;

typedef struct CpvkDrawStats {
    uint64_t primitives;        /* assembled primitives (all instances)                          */
    uint64_t fragmentsCovered;  /* N_cov: GetFragmentInput() returned true (Draw.cpp:874-954)    */
    uint64_t fragmentsWritten;  /* N_pass: not discarded and passed depth/stencil                */
    uint64_t binEntries;        /* (primitive, tile) pairs rasterised; counted on the device, 0 with statistics off */
    float msVertex, msSetup, msBin, msRaster, msTotal; /* CUDA-event times of the last draw when timing is on */
} CpvkDrawStats;

typedef union CpvkClearValue { /* VkClearValue */
    float float32[4];
    int32_t int32[4];
    uint32_t uint32[4];
    struct {
        float depth;
        uint32_t stencil;
    } depthStencil;
} CpvkClearValue;

/* ---- entry points ---- */

int cpvk_cuda_abi_version(void);
/* sizeof() of an ABI struct by name ("CpvkDrawState", ...), 0 if unknown: lets a binding in another language check its mirror
   of the PODs above against what this library was compiled with. */
size_t cpvk_cuda_abi_sizeof(const char* name);
const char* cpvk_cuda_last_error(void); /* thread-local, human readable */

/* Device bring-up == `new CPJit()` + AddGlslFunctions in Device::Device (CPVulkan/Device.cpp:17-27). */
int cpvk_cuda_device_create(int cudaOrdinal, CpvkDevice** outDevice);
void cpvk_cuda_device_destroy(CpvkDevice* device);
/* Use an externally owned stream (e.g. torch's current stream) for every launch; 0 = the device's own. */
int cpvk_cuda_device_set_stream(CpvkDevice* device, void* cudaStream);
int cpvk_cuda_device_set_timing(CpvkDevice* device, int enabled);
/* Count fragments (N_cov / N_pass) inside the raster kernel; off by default only costs nothing. */
int cpvk_cuda_device_set_stats(CpvkDevice* device, int enabled);
int cpvk_cuda_sync(CpvkDevice* device);

/* Memory == DeviceMemory (CPVulkan/Util.h:8-37, Device.cpp:181-201): one host-visible, host-coherent type.
   The allocation has an HBM-resident body and a pinned host shadow the application maps; the ICD mirrors
   shadow -> HBM before a submit's commands and HBM -> shadow before the fence signals. */
int cpvk_cuda_mem_alloc(CpvkDevice* device, size_t size, uint64_t* outDeviceAddress, void** outHostShadow);
int cpvk_cuda_mem_free(CpvkDevice* device, uint64_t deviceAddress);
int cpvk_cuda_mem_upload(CpvkDevice* device, uint64_t deviceAddress, const void* host, size_t size);
int cpvk_cuda_mem_download(CpvkDevice* device, void* host, uint64_t deviceAddress, size_t size);
/* The same without waiting: the bytes are in `host` (pinned memory) once a later cpvk_cuda_sync returns. Lets a frame
   loop overlap the read-back of frame k with the upload and rendering of frame k+1 on a second device object. */
int cpvk_cuda_mem_download_async(CpvkDevice* device, void* host, uint64_t deviceAddress, size_t size);

/* vkCreateGraphicsPipelines: SPIR-V -> CUDA C++ device functions -> NVRTC (LTO-IR) -> nvJitLink with the
   prebuilt stage kernels -> cubin. Replaces CompileVertexPipeline/CompileFragmentPipeline + the x86 ORC JIT. */
int cpvk_cuda_pipeline_create(CpvkDevice* device, const CpvkPipelineDesc* desc, CpvkPipeline** outPipeline);
void cpvk_cuda_pipeline_destroy(CpvkDevice* device, CpvkPipeline* pipeline);
/* Generated CUDA C++ for inspection/tests (NUL terminated, owned by the pipeline). */
const char* cpvk_cuda_pipeline_source(const CpvkPipeline* pipeline);
/* Linked cubin image (for cuobjdump -sass / ncu source correlation). */
const void* cpvk_cuda_pipeline_cubin(const CpvkPipeline* pipeline, size_t* outSize);
/* Compile-only variant used on machines without a GPU (build check, CPU tests): no module load. */
int cpvk_cuda_pipeline_compile_only(const CpvkPipelineDesc* desc, CpvkPipeline** outPipeline);

/* vkCmdDraw / vkCmdDrawIndexed execution. Asynchronous on the device stream. */
int cpvk_cuda_draw(CpvkDevice* device, const CpvkDrawState* state);
int cpvk_cuda_last_draw_stats(CpvkDevice* device, CpvkDrawStats* outStats); /* synchronises */
/* Number of kernels this library launched on the device since creation (bench.py "gpu_launches"). */
uint64_t cpvk_cuda_launch_count(const CpvkDevice* device);

/* Render-pass loadOp CLEAR / vkCmdClear*Image: whole subresource, per-texel SetPixel semantics
   (CommandBuffer.cpp:591-640, Draw.cpp:117-149). isDepthStencil selects VkClearDepthStencilValue. */
int cpvk_cuda_clear(CpvkDevice* device, const CpvkAttachment* image, const CpvkClearValue* value, int isDepthStencil);
/* Clears are deferred by default: a clear is recorded and folded into the next cpvk_cuda_draw that renders to the same
   attachment (its tiles start from the clear value, so the clear costs no pass over HBM); every other entry point of
   this library that could observe the memory materialises pending clears first. Code that reads the memory behind the
   library's back in stream order (e.g. a collective enqueued on the same stream right after a clear) calls
   cpvk_cuda_flush first; cpvk_cuda_device_set_lazy_clear(device, 0) turns the deferral off.
   The library also remembers the index range of the last indexed draw and forgets it when one of its own commands writes
   the index bytes; code that rewrites an index buffer behind the library's back (its own kernels, a collective) calls
   cpvk_cuda_flush afterwards as well. */
int cpvk_cuda_flush(CpvkDevice* device);
/* Draws are enqueued to the end without waiting for the binning counts: list capacity, sort mode and the large-
   primitive passes are guessed from the previous draw, checked on the device, and the tail of the draw is replayed
   with exact sizes when the guess was wrong (the check happens before this library enqueues or reads anything else,
   so results never depend on the guess). On a stream supplied through cpvk_cuda_device_set_stream the check (and the
   replay, if any) happens before cpvk_cuda_draw returns, so that work the caller enqueues on that stream right after the
   call — a collective, an event, a copy — is ordered after the complete draw; on the device's own stream it happens at the
   next entry point. 0 turns speculation off: every draw then waits for its counts (one host round trip). */
int cpvk_cuda_device_set_speculation(CpvkDevice* device, int enable);
/* Front-end overlap: the front end of a draw (descriptor copy, vertex stage, primitive setup, binning) reads only the draw's
   inputs and writes scratch memory of its own, so it runs on an internal second stream while the PREVIOUS draw's raster
   kernel is still busy — whenever the library can tell that nothing in flight writes those inputs: back-to-back draws with
   no other command of this header in between (cpvk_cuda_peer_barrier excepted), no shader stores, the remembered index
   range, and the previous draw's attachments outside every allocation (cpvk_cuda_mem_alloc) this draw's vertex buffers,
   index buffer and descriptors point into. Everything else is ordered exactly as before; results never differ.
   On the device's own stream this is on by default. On a stream supplied through cpvk_cuda_device_set_stream foreign work
   may sit between two draws where the library cannot see it, so it is off unless enabled here. Enabling it makes the
   library treat that stream like its own, and the caller promises two things: (1) cpvk_cuda_flush after foreign work on
   the stream that writes anything a later draw reads (the promise the remembered index range asks for as well), and (2)
   cpvk_cuda_flush BEFORE enqueuing foreign work that must come behind a draw (an event, a collective, a copy): the
   validation of a speculative draw — and its replay — then waits for the next entry point of this header instead of
   happening before cpvk_cuda_draw returns, and cpvk_cuda_flush is such an entry point.
   CPVK_OVERLAP=0 / 1 in the environment overrides the default for every device. */
int cpvk_cuda_device_set_overlap(CpvkDevice* device, int enable);
int cpvk_cuda_device_set_lazy_clear(CpvkDevice* device, int enable);

/* ---- multi-GPU (the reference's hook is the device-group plumbing it accepts and ignores, Queue.cpp:27-29, :52-74) ----
   One process, one handle, N GPUs of one box. cpvk_cuda_device_create_group returns the LEADER device; every entry point of
   this header accepts it and fans out: allocations exist on every member (the leader's address is the handle), uploads,
   clears, copies and blits run on every member (replicated resources, one PCIe link each), cpvk_cuda_draw renders one
   sort-first band of tile rows per member (vertex work replicated), and cpvk_cuda_gather copies every member's band of the
   named images into every other member's replica over NVLink (peer copies ordered by events), after which all replicas hold
   the whole images again. Between a draw and its gather only the member's own band of the attachments is defined.
   Downloads read the leader's replica. CPVK_GROUP_SERIAL=1 enqueues the members one after the other from the calling thread
   instead of from one worker thread per member. A group runs on its members' own streams (set_stream is refused). An ordinal
   may be listed more than once: each occurrence is a member of its own (replicas, stream, band) on that GPU. */
int cpvk_cuda_device_create_group(const int* ordinals, uint32_t count, CpvkDevice** out);
uint32_t cpvk_cuda_group_size(const CpvkDevice* device); /* 1 for a plain device */
int cpvk_cuda_gather(CpvkDevice* device, const CpvkAttachment* images, uint32_t count); /* no-op on a plain device */
/* One process per GPU (torchrun): whole allocations made by cpvk_cuda_mem_alloc can be mapped into another process's device
   (cudaIpc): export fills a 64-byte handle, import returns the address in the importing process, to be used as a
   CpvkDrawState.mirrorColor0[] target. */
int cpvk_cuda_mem_export(CpvkDevice* device, uint64_t dev, void* handle64);
int cpvk_cuda_mem_import(CpvkDevice* device, const void* handle64, uint64_t* outDev);
int cpvk_cuda_mem_unimport(CpvkDevice* device, uint64_t dev);
/* The ordering step of such a run, on the device instead of through a collective library: one small kernel on the device's
   stream that (1) stores `sequence` into word `self` of every participant's flag array — a system-scope release, so
   everything this stream did before, k_raster's stores into peers' frames included, is visible to whoever sees the flag —
   and (2) waits until every word of its OWN flag array has reached `sequence`. Work enqueued behind it therefore runs after
   every participant's stream has reached its own call with this sequence number. flagArrays[i] = participant i's array as
   addressable from this device (own allocation for i == self, cpvk_cuda_mem_import for the others), `count` <= 16 words
   each, zero-filled (cpvk_cuda_mem_alloc does that); sequence starts at 1 and grows by one per call on every participant.
   Called right behind a draw whose validation is still pending it costs no host round trip: the kernel reads the draw's
   verdict on the device, does nothing if the draw has to be replayed, and is issued again behind the replay.
   A participant that never arrives makes the kernel trap after 10 s (every later call on the device then fails) instead
   of hanging the GPU. */
int cpvk_cuda_peer_barrier(CpvkDevice* device, const uint64_t* flagArrays, uint32_t count, uint32_t self, uint32_t sequence);

/* Diagnostics. The kernels divide several numerators by one denominator through one shared reciprocal (edge weights / area,
   interpolated components / denominator, position / w), claiming the bits of the IEEE `/` operator: this runs both on device
   arrays a[3 n], b[n] -> outShared[3 n], outPlain[3 n] so that a test can hold them to a host-side IEEE division. */
int cpvk_cuda_selftest_div(CpvkDevice* device, uint64_t a, uint64_t b, uint32_t n, uint64_t outShared, uint64_t outPlain);

/* vkCmdCopyImage / CopyBufferToImage / CopyImageToBuffer: raw row memcpy (CommandBuffer.Copy.cpp:77-200). */
int cpvk_cuda_copy_rows(CpvkDevice* device, uint64_t dst, uint32_t dstPitch, uint64_t src, uint32_t srcPitch,
                        uint32_t rowBytes, uint32_t rows);

/* vkCmdBlitImage for one 2-D region (CommandBuffer.cpp:57-232): SampleImage(filter) + SetPixel per dst texel. */
typedef struct CpvkBlit {
    CpvkAttachment src, dst;
    int32_t srcX0, srcY0, srcX1, srcY1;
    int32_t dstX0, dstY0, dstX1, dstY1;
    uint32_t filter; /* VkFilter */
} CpvkBlit;
int cpvk_cuda_blit(CpvkDevice* device, const CpvkBlit* blit);

#ifdef __cplusplus
}
#endif
#endif /* CPVK_CUDA_H */
