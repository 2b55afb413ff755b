// cpvk_icd.cpp — the host side of the B200 Vulkan ICD: the reference's object model and command recording,
// with every Command::Process on the draw path forwarded to the CUDA C ABI (include/cpvk_cuda.h).
//
// Mirrors, for the draw path only (SURVEY §2.2 "★" rows, App. C entry points):
//   CPVulkan/CPVulkan.cpp:14-104            vk_icd* exports, name -> entry table (Extensions.cpp / VulkanFunctions.h)
//   CPVulkanBase/Base.h:256-293,346-375     handle layout: dispatchable = [16-byte header (ICD_LOADER_MAGIC) | object]
//   CPVulkan/Device.cpp:181-201, Util.h:8-37  one DEVICE_LOCAL|HOST_VISIBLE|HOST_COHERENT memory type, raw mapped pointer
//   CPVulkan/Image.cpp, Formats.cpp:455-483  linear images, Stride = texel * width, mips then layers
//   CPVulkan/CommandBuffer*.cpp              vkCmd* record a command; vkQueueSubmit runs them in order, synchronously
//   CPVulkan/Queue.cpp:11-77                 submit = execute inline, then signal the fence
// Memory semantics: VkDeviceMemory = HBM allocation + pinned host shadow. The application maps the
// shadow. At submit every allocation the host may have written (mapped now, or mapped since the last upload) is
// copied to HBM; after the commands every allocation the GPU wrote is copied back if it is mapped (or lazily at
// the next vkMapMemory) — so a host-coherent mapping observes results after the fence exactly as with the
// reference's malloc'ed memory.
// Unsupported state follows the reference convention (Base.h:73-74): abort().
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/cpvk_cuda.h"
#include "../../include/cpvk_vulkan.h"

namespace {

[[noreturn]] void Fatal(const char* what) {
    fprintf(stderr, "CPVulkan_b200: %s (%s)\n", what, cpvk_cuda_last_error());
    abort(); // TODO_ERROR / FATAL_ERROR convention of the reference
}
#define CU_CHECK(expr) do { int rc_ = (expr); if (rc_ != 0) Fatal(#expr); } while (0)

// ---- handles ----
struct DispatchHeader { uintptr_t loaderMagic; uintptr_t pad; };
static_assert(sizeof(DispatchHeader) == 16, "dispatchable header is 16 bytes (Base.h:256-276)");
template <class T> struct Dispatchable { DispatchHeader hdr{ICD_LOADER_MAGIC, 0}; T obj; };
template <class T, class H> T* Unwrap(H handle) { return handle ? reinterpret_cast<T*>(reinterpret_cast<char*>(handle) + sizeof(DispatchHeader)) : nullptr; }
template <class H, class T> H Wrap(Dispatchable<T>* d) { return reinterpret_cast<H>(d); }
template <class T> Dispatchable<T>* Outer(T* obj) { return reinterpret_cast<Dispatchable<T>*>(reinterpret_cast<char*>(obj) - sizeof(DispatchHeader)); }

uint32_t TexelSize(uint32_t f) { // Formats.cpp:219-341 TotalSize
    if (f >= 9 && f <= 50) { const uint32_t fam = (f - 9) / 7; return fam == 0 ? 1 : fam == 1 ? 2 : (fam == 2 || fam == 3) ? 3 : 4; }
    if (f >= 51 && f <= 69) return 4;
    if (f >= 70 && f <= 97) return 2 * ((f - 70) / 7 + 1);
    if (f >= 98 && f <= 109) return 4 * ((f - 98) / 3 + 1);
    switch (f) { case 124: return 2; case 125: case 126: case 129: return 4; case 127: return 1; case 128: return 3; case 130: return 8; default: return 0; }
}
bool IsDepthStencil(uint32_t f) { return f >= 124 && f <= 130; }

struct Device;
struct DeviceMemory {
    Device* device = nullptr;
    VkDeviceSize size = 0;
    uint64_t devAddr = 0;
    void* host = nullptr;
    bool mapped = false;
    bool hostDirty = true;    // host may hold bytes HBM lacks
    bool deviceDirty = false; // HBM holds bytes the host shadow lacks
    bool gpuReadable = false; // some bound resource can be read by the GPU; pure TRANSFER_DST (readback) memory never needs uploading
};
struct Buffer { VkDeviceSize size = 0; VkBufferUsageFlags usage = 0; DeviceMemory* mem = nullptr; VkDeviceSize memOffset = 0;
    uint64_t address(VkDeviceSize off = 0) const { return mem ? mem->devAddr + memOffset + off : 0; } };
struct MipInfo { uint64_t offset, levelSize, planeSize, stride; uint32_t width, height, depth; };
struct Image {
    VkFormat format = VK_FORMAT_UNDEFINED; VkExtent3D extent{}; uint32_t mipLevels = 1, arrayLayers = 1; VkImageUsageFlags usage = 0;
    std::vector<MipInfo> levels; uint64_t layerSize = 0, totalSize = 0;
    DeviceMemory* mem = nullptr; VkDeviceSize memOffset = 0;
    uint64_t address(uint32_t level, uint32_t layer) const { return mem->devAddr + memOffset + layerSize * layer + levels[level].offset; }
};
struct ImageView { Image* image = nullptr; VkFormat format = VK_FORMAT_UNDEFINED; VkImageViewType viewType = VK_IMAGE_VIEW_TYPE_2D; VkComponentMapping components{}; VkImageSubresourceRange range{}; };
struct BufferView { Buffer* buffer = nullptr; VkFormat format = VK_FORMAT_UNDEFINED; VkDeviceSize offset = 0, range = 0; };
struct Sampler { CpvkSampler s{}; };
struct ShaderModule { std::vector<uint32_t> code; };
// pImmutableSamplers are copied at creation (DescriptorSetLayout.cpp:44-58): the application's array need not outlive the call
struct DescriptorSetLayout { std::vector<VkDescriptorSetLayoutBinding> bindings; std::map<uint32_t, std::vector<Sampler*>> immutableSamplers; };
struct PipelineLayout { int unused = 0; };
struct DescriptorPool { int unused = 0; };
struct DescriptorValue { VkDescriptorType type = VK_DESCRIPTOR_TYPE_MAX_ENUM; Buffer* buffer = nullptr; VkDeviceSize offset = 0, range = 0; ImageView* view = nullptr; Sampler* sampler = nullptr; BufferView* bufferView = nullptr; };
struct DescriptorSet { DescriptorSetLayout* layout = nullptr; std::map<uint32_t, std::vector<DescriptorValue>> bindings; std::map<uint32_t, bool> immutable; };
struct Subpass { std::vector<VkAttachmentReference> color; VkAttachmentReference depthStencil{VK_ATTACHMENT_UNUSED, VK_IMAGE_LAYOUT_UNDEFINED}; };
struct RenderPass { std::vector<VkAttachmentDescription> attachments; std::vector<Subpass> subpasses; };
struct Framebuffer { std::vector<ImageView*> views; uint32_t width = 0, height = 0; };
struct Pipeline { CpvkPipeline* cuda = nullptr; bool dynamicViewport = false; VkViewport staticViewport{}; };
struct PipelineCache { int unused = 0; };
struct CommandPool { Device* device = nullptr; };
struct Fence { std::atomic<bool> signaled{false}; };
struct Semaphore { int unused = 0; };

// The mutable execution context commands read and write (CPVulkan/DeviceState.h:92-121).
struct DeviceState {
    Pipeline* pipeline = nullptr;
    struct { Buffer* buffer = nullptr; VkDeviceSize offset = 0; } vertex[CPVK_MAX_VERTEX_BINDINGS];
    Buffer* indexBuffer = nullptr; VkDeviceSize indexOffset = 0; uint32_t indexStride = 0;
    DescriptorSet* sets[8] = {}; std::vector<uint32_t> dynamicOffsets[8];
    VkViewport viewport{};
    uint8_t pushConstants[CPVK_MAX_PUSH_CONSTANT_BYTES] = {};
    RenderPass* renderPass = nullptr; Framebuffer* framebuffer = nullptr; uint32_t subpass = 0;
};

struct Queue;
struct PhysicalDevice;
struct Instance { Dispatchable<PhysicalDevice>* physical = nullptr; };
struct PhysicalDevice { Instance* instance = nullptr; };
struct Device {
    CpvkDevice* cuda = nullptr;
    Dispatchable<Queue>* queue = nullptr;
    std::vector<DeviceMemory*> memories;
    DeviceState state;
    uint32_t bandY0 = 0, bandY1 = 0; // sort-first band of this process (CPVK_BAND="rank/world"), SURVEY §8(e)
};
struct Queue { Device* device = nullptr; };
using Command = std::function<void(Device&)>;
struct CommandBuffer { Device* device = nullptr; std::vector<Command> commands; };

void TouchedByGpu(DeviceMemory* m) { if (m) m->deviceDirty = true; }

void FillImageLayout(Image& img) { // GetNormalImageSize (Formats.cpp:455-483)
    const uint32_t texel = TexelSize(img.format);
    if (!texel) Fatal("unsupported image format");
    uint32_t w = img.extent.width, h = img.extent.height, d = img.extent.depth;
    img.levels.clear(); img.layerSize = 0;
    for (uint32_t i = 0; i < img.mipLevels; i++) {
        MipInfo m{};
        m.offset = img.layerSize; m.width = w; m.height = h; m.depth = d;
        m.stride = (uint64_t)texel * w; m.planeSize = m.stride * h; m.levelSize = m.planeSize * d;
        img.layerSize += m.levelSize; img.levels.push_back(m);
        w = w / 2 ? w / 2 : 1; h = h / 2 ? h / 2 : 1; d = d / 2 ? d / 2 : 1;
    }
    img.totalSize = img.layerSize * img.arrayLayers;
}

CpvkAttachment AttachmentOf(const ImageView* v) {
    CpvkAttachment a{};
    const Image* img = v->image;
    if (!img->mem) Fatal("attachment image has no memory bound");
    const MipInfo& m = img->levels[v->range.baseMipLevel];
    a.address = img->address(v->range.baseMipLevel, v->range.baseArrayLayer);
    a.width = m.width; a.height = m.height; a.rowPitch = (uint32_t)m.stride; a.format = (uint32_t)v->format;
    return a;
}

// ---- command execution: the Process() bodies ----
void ExecBeginRenderPass(Device& d, RenderPass* rp, Framebuffer* fb, const std::vector<VkClearValue>& clears) { // CommandBuffer.cpp:591-640
    DeviceState& s = d.state;
    s.renderPass = rp; s.framebuffer = fb; s.subpass = 0;
    const Subpass& sp = rp->subpasses[0];
    for (const VkAttachmentReference& ref : sp.color) {
        if (ref.attachment == VK_ATTACHMENT_UNUSED) continue;
        const VkAttachmentDescription& ad = rp->attachments[ref.attachment];
        ImageView* v = fb->views[ref.attachment];
        if (ad.loadOp == VK_ATTACHMENT_LOAD_OP_CLEAR) { // whole subresource, renderArea ignored, like ClearImage
            CpvkClearValue cv; memcpy(&cv, &clears[ref.attachment], sizeof cv);
            for (uint32_t layer = 0; layer < (v->range.layerCount == VK_REMAINING_ARRAY_LAYERS ? v->image->arrayLayers - v->range.baseArrayLayer : v->range.layerCount); layer++)
                for (uint32_t level = 0; level < (v->range.levelCount == VK_REMAINING_MIP_LEVELS ? v->image->mipLevels - v->range.baseMipLevel : v->range.levelCount); level++) {
                    ImageView sub = *v; sub.range.baseMipLevel += level; sub.range.baseArrayLayer += layer;
                    CpvkAttachment a = AttachmentOf(&sub); a.format = (uint32_t)ad.format;
                    CU_CHECK(cpvk_cuda_clear(d.cuda, &a, &cv, 0));
                }
        }
        TouchedByGpu(v->image->mem);
    }
    if (sp.depthStencil.attachment != VK_ATTACHMENT_UNUSED) {
        const VkAttachmentDescription& ad = rp->attachments[sp.depthStencil.attachment];
        ImageView* v = fb->views[sp.depthStencil.attachment];
        if (ad.loadOp == VK_ATTACHMENT_LOAD_OP_CLEAR) {
            CpvkClearValue cv; memcpy(&cv, &clears[sp.depthStencil.attachment], sizeof cv);
            CpvkAttachment a = AttachmentOf(v); a.format = (uint32_t)ad.format;
            CU_CHECK(cpvk_cuda_clear(d.cuda, &a, &cv, 1));
        }
        TouchedByGpu(v->image->mem);
    }
}

// Several GPUs behind one VkDevice (CPVK_CUDA_DEVICES): each renders its sort-first band of the attachments; when a subpass
// ends, the bands are exchanged over NVLink so that every GPU's replica of every attachment is whole again for whatever
// reads it next (the next subpass's input attachments, a copy, a blit, a sampled read, the host). No-op on one GPU.
void GatherSubpassAttachments(Device& d) {
    DeviceState& s = d.state;
    if (!s.renderPass || !s.framebuffer || cpvk_cuda_group_size(d.cuda) < 2) return;
    const Subpass& sp = s.renderPass->subpasses[s.subpass];
    std::vector<CpvkAttachment> atts;
    for (const VkAttachmentReference& ref : sp.color) if (ref.attachment != VK_ATTACHMENT_UNUSED) atts.push_back(AttachmentOf(s.framebuffer->views[ref.attachment]));
    if (sp.depthStencil.attachment != VK_ATTACHMENT_UNUSED) atts.push_back(AttachmentOf(s.framebuffer->views[sp.depthStencil.attachment]));
    if (!atts.empty()) CU_CHECK(cpvk_cuda_gather(d.cuda, atts.data(), (uint32_t)atts.size()));
}

void FillDescriptor(CpvkDescriptor& out, uint32_t set, uint32_t binding, uint32_t element, const DescriptorValue& v, uint32_t dynamicOffset) { // LoadUniforms, Draw.cpp:356-408
    memset(&out, 0, sizeof out);
    out.set = set; out.binding = binding; out.arrayElement = element;
    switch (v.type) {
    case VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER: case VK_DESCRIPTOR_TYPE_STORAGE_BUFFER: case VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER_DYNAMIC: case VK_DESCRIPTOR_TYPE_STORAGE_BUFFER_DYNAMIC:
        out.type = CPVK_DESC_BUFFER; out.address = v.buffer->address(v.offset + dynamicOffset);
        out.range = v.range == VK_WHOLE_SIZE ? v.buffer->size - v.offset : v.range;
        if (v.type == VK_DESCRIPTOR_TYPE_STORAGE_BUFFER || v.type == VK_DESCRIPTOR_TYPE_STORAGE_BUFFER_DYNAMIC) TouchedByGpu(v.buffer->mem);
        break;
    case VK_DESCRIPTOR_TYPE_UNIFORM_TEXEL_BUFFER: case VK_DESCRIPTOR_TYPE_STORAGE_TEXEL_BUFFER:
        out.type = CPVK_DESC_TEXEL_BUFFER; out.address = v.bufferView->buffer->address(v.bufferView->offset);
        out.range = v.bufferView->range == VK_WHOLE_SIZE ? v.bufferView->buffer->size - v.bufferView->offset : v.bufferView->range;
        out.format = (uint32_t)v.bufferView->format; out.dimensions = 1; out.levelCount = 1;
        break;
    case VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER: case VK_DESCRIPTOR_TYPE_SAMPLED_IMAGE: case VK_DESCRIPTOR_TYPE_INPUT_ATTACHMENT: {
        out.type = CPVK_DESC_IMAGE;
        const ImageView* iv = v.view; const Image* img = iv->image;
        out.format = (uint32_t)iv->format;
        out.dimensions = iv->viewType == VK_IMAGE_VIEW_TYPE_1D ? 1 : iv->viewType == VK_IMAGE_VIEW_TYPE_3D ? 3 : 2;
        uint32_t levels = iv->range.levelCount == VK_REMAINING_MIP_LEVELS ? img->mipLevels - iv->range.baseMipLevel : iv->range.levelCount;
        if (levels > CPVK_MAX_MIP_LEVELS) levels = CPVK_MAX_MIP_LEVELS;
        out.levelCount = levels;
        for (uint32_t l = 0; l < levels; l++) { // GetImageData (GlslFunctions.cpp:378-421): levels relative to the view's base
            const MipInfo& m = img->levels[iv->range.baseMipLevel + l];
            out.levels[l].address = img->address(iv->range.baseMipLevel + l, iv->range.baseArrayLayer);
            out.levels[l].width = m.width; out.levels[l].height = m.height; out.levels[l].depth = m.depth;
        }
        out.swizzle[0] = iv->components.r; out.swizzle[1] = iv->components.g; out.swizzle[2] = iv->components.b; out.swizzle[3] = iv->components.a;
        if (v.sampler) out.sampler = v.sampler->s;
        break; }
    case VK_DESCRIPTOR_TYPE_SAMPLER: // combined with a SAMPLED_IMAGE by OpSampledImage in the shader (Samples/separate_image_sampler)
        out.type = CPVK_DESC_SAMPLER;
        if (v.sampler) out.sampler = v.sampler->s;
        break;
    default: Fatal("descriptor type not built");
    }
}

void ExecDraw(Device& d, uint32_t count, uint32_t instanceCount, uint32_t first, int32_t vertexOffset, uint32_t firstInstance, bool indexed) { // Draw.cpp:1777-1864
    DeviceState& s = d.state;
    if (!s.pipeline || !s.renderPass || !s.framebuffer) Fatal("draw outside a render pass or without a pipeline");
    auto st = std::make_unique<CpvkDrawState>();
    memset(st.get(), 0, sizeof(CpvkDrawState));
    st->pipeline = s.pipeline->cuda;
    const VkViewport& vp = s.pipeline->dynamicViewport ? s.viewport : s.pipeline->staticViewport;
    st->viewport = CpvkViewport{vp.x, vp.y, vp.width, vp.height, vp.minDepth, vp.maxDepth};
    for (uint32_t b = 0; b < CPVK_MAX_VERTEX_BINDINGS; b++) if (s.vertex[b].buffer) st->vertexBuffers[b] = s.vertex[b].buffer->address(s.vertex[b].offset);
    if (indexed) {
        if (!s.indexBuffer) Fatal("vkCmdDrawIndexed without an index buffer");
        st->indexBuffer = s.indexBuffer->address(s.indexOffset); st->indexStride = s.indexStride;
    }
    st->count = count; st->instanceCount = instanceCount; st->first = first; st->vertexOffset = vertexOffset; st->firstInstance = firstInstance;
    uint32_t nd = 0;
    for (uint32_t set = 0; set < 8; set++) {
        if (!s.sets[set]) continue;
        uint32_t dyn = 0;
        for (auto& kv : s.sets[set]->bindings)
            for (uint32_t e = 0; e < kv.second.size(); e++) {
                const DescriptorValue& v = kv.second[e];
                if (v.type == VK_DESCRIPTOR_TYPE_MAX_ENUM) continue;
                uint32_t dynOff = 0;
                if (v.type == VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER_DYNAMIC || v.type == VK_DESCRIPTOR_TYPE_STORAGE_BUFFER_DYNAMIC) dynOff = dyn < s.dynamicOffsets[set].size() ? s.dynamicOffsets[set][dyn++] : 0;
                if (nd >= CPVK_MAX_DESCRIPTORS) Fatal("too many descriptors bound");
                FillDescriptor(st->descriptors[nd++], set, kv.first, e, v, dynOff);
            }
    }
    st->descriptorCount = nd;
    st->pushConstantSize = CPVK_MAX_PUSH_CONSTANT_BYTES;
    memcpy(st->pushConstants, s.pushConstants, CPVK_MAX_PUSH_CONSTANT_BYTES);
    const Subpass& sp = s.renderPass->subpasses[s.subpass];
    for (uint32_t i = 0; i < sp.color.size() && i < CPVK_MAX_COLOR_ATTACHMENTS; i++) { // ProcessFragmentShader, Draw.cpp:1627-1640
        if (sp.color[i].attachment == VK_ATTACHMENT_UNUSED) continue;
        ImageView* v = s.framebuffer->views[sp.color[i].attachment];
        st->color[i] = AttachmentOf(v);
        TouchedByGpu(v->image->mem);
    }
    if (sp.depthStencil.attachment != VK_ATTACHMENT_UNUSED) {
        ImageView* v = s.framebuffer->views[sp.depthStencil.attachment];
        st->depthStencil = AttachmentOf(v);
        TouchedByGpu(v->image->mem);
    }
    st->bandY0 = d.bandY0; st->bandY1 = d.bandY1;
    CU_CHECK(cpvk_cuda_draw(d.cuda, st.get()));
}

void UploadHostWrites(Device& d) {
    for (DeviceMemory* m : d.memories)
        if (m->gpuReadable && (m->hostDirty || m->mapped)) {
            CU_CHECK(cpvk_cuda_mem_upload(d.cuda, m->devAddr, m->host, (size_t)m->size));
            m->hostDirty = false;
        }
}
void DownloadDeviceWrites(Device& d) {
    for (DeviceMemory* m : d.memories)
        if (m->deviceDirty && m->mapped) { CU_CHECK(cpvk_cuda_mem_download(d.cuda, m->host, m->devAddr, (size_t)m->size)); m->deviceDirty = false; }
}

// ================================== entry points ==================================
#define VKFN(ret) VKAPI_ATTR ret VKAPI_CALL

VKFN(VkResult) CreateInstance(const VkInstanceCreateInfo*, const VkAllocationCallbacks*, VkInstance* pInstance) {
    auto* inst = new Dispatchable<Instance>();
    inst->obj.physical = new Dispatchable<PhysicalDevice>();
    inst->obj.physical->obj.instance = &inst->obj;
    *pInstance = Wrap<VkInstance>(inst);
    return VK_SUCCESS;
}
VKFN(void) DestroyInstance(VkInstance instance, const VkAllocationCallbacks*) {
    if (!instance) return;
    Instance* i = Unwrap<Instance>(instance);
    delete i->physical; delete Outer(i);
}
VKFN(VkResult) EnumeratePhysicalDevices(VkInstance instance, uint32_t* pCount, VkPhysicalDevice* pDevices) { // Instance.cpp:34-49: exactly one
    if (!pDevices) { *pCount = 1; return VK_SUCCESS; }
    if (*pCount < 1) return VK_INCOMPLETE;
    pDevices[0] = Wrap<VkPhysicalDevice>(Unwrap<Instance>(instance)->physical); *pCount = 1;
    return VK_SUCCESS;
}
VKFN(VkResult) EnumerateInstanceExtensionProperties(const char*, uint32_t* pCount, VkExtensionProperties*) { *pCount = 0; return VK_SUCCESS; }
VKFN(VkResult) EnumerateInstanceLayerProperties(uint32_t* pCount, VkLayerProperties*) { *pCount = 0; return VK_SUCCESS; }
VKFN(VkResult) EnumerateDeviceExtensionProperties(VkPhysicalDevice, const char*, uint32_t* pCount, VkExtensionProperties*) { *pCount = 0; return VK_SUCCESS; }
VKFN(VkResult) EnumerateInstanceVersion(uint32_t* v) { *v = VK_MAKE_VERSION(1, 1, 121); return VK_SUCCESS; }
VKFN(void) GetPhysicalDeviceProperties(VkPhysicalDevice, VkPhysicalDeviceProperties* p) { // PhysicalDevice.cpp + Config.h limits
    memset(p, 0, sizeof *p);
    p->apiVersion = VK_MAKE_VERSION(1, 1, 121); p->driverVersion = 1; p->vendorID = 0x10DE; p->deviceID = 0xB200; p->deviceType = VK_PHYSICAL_DEVICE_TYPE_DISCRETE_GPU;
    snprintf(p->deviceName, sizeof p->deviceName, "CPVulkan B200 draw path");
    VkPhysicalDeviceLimits& l = p->limits;
    l.maxImageDimension1D = l.maxImageDimension2D = l.maxImageDimensionCube = 16384; l.maxImageDimension3D = 256; l.maxImageArrayLayers = 256;
    l.maxTexelBufferElements = 1u << 27; l.maxUniformBufferRange = 1u << 16; l.maxStorageBufferRange = 1u << 27; l.maxPushConstantsSize = CPVK_MAX_PUSH_CONSTANT_BYTES;
    l.maxMemoryAllocationCount = 4096; l.maxSamplerAllocationCount = 4000; l.bufferImageGranularity = 1; l.maxBoundDescriptorSets = 8;
    l.maxPerStageDescriptorSamplers = l.maxPerStageDescriptorUniformBuffers = l.maxPerStageDescriptorStorageBuffers = l.maxPerStageDescriptorSampledImages = CPVK_MAX_DESCRIPTORS;
    l.maxPerStageResources = CPVK_MAX_DESCRIPTORS; l.maxDescriptorSetSamplers = l.maxDescriptorSetUniformBuffers = l.maxDescriptorSetSampledImages = CPVK_MAX_DESCRIPTORS;
    l.maxVertexInputAttributes = CPVK_MAX_VERTEX_ATTRIBUTES; l.maxVertexInputBindings = CPVK_MAX_VERTEX_BINDINGS; l.maxVertexInputAttributeOffset = 2047; l.maxVertexInputBindingStride = 2048;
    l.maxVertexOutputComponents = 64; l.maxFragmentInputComponents = 64; l.maxFragmentOutputAttachments = CPVK_MAX_COLOR_ATTACHMENTS; l.maxColorAttachments = CPVK_MAX_COLOR_ATTACHMENTS;
    l.subPixelPrecisionBits = 4; l.subTexelPrecisionBits = 4; l.mipmapPrecisionBits = 4; l.maxDrawIndexedIndexValue = 0xFFFFFFFFu; l.maxDrawIndirectCount = 1;
    l.maxSamplerLodBias = 32.0f; l.maxSamplerAnisotropy = 1.0f; l.maxViewports = 1; l.maxViewportDimensions[0] = l.maxViewportDimensions[1] = 16384;
    l.viewportBoundsRange[0] = -32768.0f; l.viewportBoundsRange[1] = 32767.0f; l.minMemoryMapAlignment = 64;
    l.minTexelBufferOffsetAlignment = 16; l.minUniformBufferOffsetAlignment = 16; l.minStorageBufferOffsetAlignment = 16;
    l.maxFramebufferWidth = l.maxFramebufferHeight = 16384; l.maxFramebufferLayers = 256;
    l.framebufferColorSampleCounts = l.framebufferDepthSampleCounts = l.framebufferStencilSampleCounts = l.framebufferNoAttachmentsSampleCounts = 1;
    l.sampledImageColorSampleCounts = l.sampledImageIntegerSampleCounts = l.sampledImageDepthSampleCounts = l.sampledImageStencilSampleCounts = l.storageImageSampleCounts = 1;
    l.maxSampleMaskWords = 1; l.timestampPeriod = 1.0f; l.maxClipDistances = 1; l.discreteQueuePriorities = 2;
    l.pointSizeRange[0] = l.pointSizeRange[1] = 1.0f; l.lineWidthRange[0] = l.lineWidthRange[1] = 1.0f; l.strictLines = VK_TRUE; l.standardSampleLocations = VK_TRUE;
    l.optimalBufferCopyOffsetAlignment = 16; l.optimalBufferCopyRowPitchAlignment = 16; l.nonCoherentAtomSize = 64;
}
VKFN(void) GetPhysicalDeviceFeatures(VkPhysicalDevice, VkPhysicalDeviceFeatures* f) { memset(f, 0, sizeof *f); f->fullDrawIndexUint32 = VK_TRUE; f->independentBlend = VK_TRUE; }
VKFN(void) GetPhysicalDeviceQueueFamilyProperties(VkPhysicalDevice, uint32_t* pCount, VkQueueFamilyProperties* p) {
    if (!p) { *pCount = 1; return; }
    if (*pCount >= 1) { p[0] = VkQueueFamilyProperties{VK_QUEUE_GRAPHICS_BIT | VK_QUEUE_TRANSFER_BIT, 1, 0, {1, 1, 1}}; *pCount = 1; }
}
VKFN(void) GetPhysicalDeviceMemoryProperties(VkPhysicalDevice, VkPhysicalDeviceMemoryProperties* p) { // PhysicalDevice.cpp:436-440: one type, all three bits
    memset(p, 0, sizeof *p);
    p->memoryTypeCount = 1; p->memoryTypes[0] = VkMemoryType{VK_MEMORY_PROPERTY_DEVICE_LOCAL_BIT | VK_MEMORY_PROPERTY_HOST_VISIBLE_BIT | VK_MEMORY_PROPERTY_HOST_COHERENT_BIT, 0};
    p->memoryHeapCount = 1; p->memoryHeaps[0] = VkMemoryHeap{160ull << 30, VK_MEMORY_HEAP_DEVICE_LOCAL_BIT};
}
VKFN(void) GetPhysicalDeviceFormatProperties(VkPhysicalDevice, VkFormat format, VkFormatProperties* p) {
    memset(p, 0, sizeof *p);
    if (!TexelSize((uint32_t)format)) return;
    VkFormatFeatureFlags f = VK_FORMAT_FEATURE_SAMPLED_IMAGE_BIT | VK_FORMAT_FEATURE_SAMPLED_IMAGE_FILTER_LINEAR_BIT | VK_FORMAT_FEATURE_TRANSFER_SRC_BIT | VK_FORMAT_FEATURE_TRANSFER_DST_BIT |
                             VK_FORMAT_FEATURE_BLIT_SRC_BIT | VK_FORMAT_FEATURE_BLIT_DST_BIT;
    f |= IsDepthStencil((uint32_t)format) ? VK_FORMAT_FEATURE_DEPTH_STENCIL_ATTACHMENT_BIT : (VK_FORMAT_FEATURE_COLOR_ATTACHMENT_BIT | VK_FORMAT_FEATURE_COLOR_ATTACHMENT_BLEND_BIT);
    p->linearTilingFeatures = p->optimalTilingFeatures = f;
    p->bufferFeatures = IsDepthStencil((uint32_t)format) ? 0 : (VK_FORMAT_FEATURE_VERTEX_BUFFER_BIT | VK_FORMAT_FEATURE_UNIFORM_TEXEL_BUFFER_BIT);
}

VKFN(VkResult) CreateDevice(VkPhysicalDevice, const VkDeviceCreateInfo*, const VkAllocationCallbacks*, VkDevice* pDevice) { // Device.cpp:17-27
    auto* dev = new Dispatchable<Device>();
    int ordinal = 0;
    if (const char* e = getenv("CPVK_CUDA_DEVICE")) ordinal = atoi(e);
    // CPVK_CUDA_DEVICES=0,1,...: one VkDevice over several GPUs of the box — the group (peer access between all of them) is set up
    // here, at vkCreateDevice, where the reference accepts and ignores device groups (Queue.cpp:27-29)
    std::vector<int> ordinals;
    if (const char* e = getenv("CPVK_CUDA_DEVICES")) { for (const char* q = e; *q;) { char* end; const long v = strtol(q, &end, 10); if (end == q) break; ordinals.push_back((int)v); q = *end == ',' ? end + 1 : end; } }
    const int rc = ordinals.size() > 1 ? cpvk_cuda_device_create_group(ordinals.data(), (uint32_t)ordinals.size(), &dev->obj.cuda)
                                       : cpvk_cuda_device_create(ordinals.size() == 1 ? ordinals[0] : ordinal, &dev->obj.cuda);
    if (rc != 0) { fprintf(stderr, "CPVulkan_b200: %s\n", cpvk_cuda_last_error()); delete dev; return VK_ERROR_INITIALIZATION_FAILED; }
    if (const char* b = ordinals.size() > 1 ? nullptr : getenv("CPVK_BAND")) { unsigned r = 0, w = 1, h = 0; if (sscanf(b, "%u/%u/%u", &r, &w, &h) == 3 && w > 0 && r < w) { dev->obj.bandY0 = r * (h / w); dev->obj.bandY1 = (r + 1 == w) ? h : (r + 1) * (h / w); } }
    dev->obj.queue = new Dispatchable<Queue>();
    dev->obj.queue->obj.device = &dev->obj;
    *pDevice = Wrap<VkDevice>(dev);
    return VK_SUCCESS;
}
VKFN(void) DestroyDevice(VkDevice device, const VkAllocationCallbacks*) {
    if (!device) return;
    Device* d = Unwrap<Device>(device);
    cpvk_cuda_device_destroy(d->cuda);
    delete d->queue; delete Outer(d);
}
VKFN(void) GetDeviceQueue(VkDevice device, uint32_t, uint32_t, VkQueue* pQueue) { *pQueue = Wrap<VkQueue>(Unwrap<Device>(device)->queue); }
VKFN(VkResult) DeviceWaitIdle(VkDevice device) { return cpvk_cuda_sync(Unwrap<Device>(device)->cuda) == 0 ? VK_SUCCESS : VK_ERROR_DEVICE_LOST; }
VKFN(VkResult) QueueWaitIdle(VkQueue queue) { return cpvk_cuda_sync(Unwrap<Queue>(queue)->device->cuda) == 0 ? VK_SUCCESS : VK_ERROR_DEVICE_LOST; }

// ---- memory ----
VKFN(VkResult) AllocateMemory(VkDevice device, const VkMemoryAllocateInfo* info, const VkAllocationCallbacks*, VkDeviceMemory* pMemory) {
    Device* d = Unwrap<Device>(device);
    auto* m = new DeviceMemory();
    m->device = d; m->size = info->allocationSize;
    if (cpvk_cuda_mem_alloc(d->cuda, (size_t)m->size, &m->devAddr, &m->host) != 0) { delete m; return VK_ERROR_OUT_OF_DEVICE_MEMORY; }
    memset(m->host, 0, (size_t)m->size);
    d->memories.push_back(m);
    *pMemory = reinterpret_cast<VkDeviceMemory>(m);
    return VK_SUCCESS;
}
VKFN(void) FreeMemory(VkDevice device, VkDeviceMemory memory, const VkAllocationCallbacks*) {
    if (!memory) return;
    Device* d = Unwrap<Device>(device); auto* m = reinterpret_cast<DeviceMemory*>(memory);
    for (size_t i = 0; i < d->memories.size(); i++) if (d->memories[i] == m) { d->memories.erase(d->memories.begin() + i); break; }
    cpvk_cuda_mem_free(d->cuda, m->devAddr);
    delete m;
}
VKFN(VkResult) MapMemory(VkDevice device, VkDeviceMemory memory, VkDeviceSize offset, VkDeviceSize, VkMemoryMapFlags, void** ppData) {
    Device* d = Unwrap<Device>(device); auto* m = reinterpret_cast<DeviceMemory*>(memory);
    if (m->deviceDirty) { if (cpvk_cuda_mem_download(d->cuda, m->host, m->devAddr, (size_t)m->size) != 0) return VK_ERROR_MEMORY_MAP_FAILED; m->deviceDirty = false; }
    m->mapped = true; m->hostDirty = true;
    *ppData = static_cast<char*>(m->host) + offset;
    return VK_SUCCESS;
}
VKFN(void) UnmapMemory(VkDevice, VkDeviceMemory memory) { auto* m = reinterpret_cast<DeviceMemory*>(memory); m->mapped = false; m->hostDirty = true; }
VKFN(VkResult) FlushMappedMemoryRanges(VkDevice, uint32_t, const VkMappedMemoryRange*) { return VK_SUCCESS; }      // no-ops in the reference too (F7)
VKFN(VkResult) InvalidateMappedMemoryRanges(VkDevice, uint32_t, const VkMappedMemoryRange*) { return VK_SUCCESS; }

// ---- buffers / images ----
VKFN(VkResult) CreateBuffer(VkDevice, const VkBufferCreateInfo* info, const VkAllocationCallbacks*, VkBuffer* pBuffer) {
    auto* b = new Buffer(); b->size = info->size; b->usage = info->usage; *pBuffer = reinterpret_cast<VkBuffer>(b); return VK_SUCCESS;
}
VKFN(void) DestroyBuffer(VkDevice, VkBuffer b, const VkAllocationCallbacks*) { delete reinterpret_cast<Buffer*>(b); }
VKFN(void) GetBufferMemoryRequirements(VkDevice, VkBuffer buffer, VkMemoryRequirements* r) { // Buffer.cpp: 256 for uniform/storage/texel usages else 16
    auto* b = reinterpret_cast<Buffer*>(buffer);
    const bool big = b->usage & (VK_BUFFER_USAGE_UNIFORM_BUFFER_BIT | VK_BUFFER_USAGE_STORAGE_BUFFER_BIT | VK_BUFFER_USAGE_UNIFORM_TEXEL_BUFFER_BIT | VK_BUFFER_USAGE_STORAGE_TEXEL_BUFFER_BIT);
    r->size = b->size; r->alignment = big ? 256 : 16; r->memoryTypeBits = 1;
}
VKFN(VkResult) BindBufferMemory(VkDevice, VkBuffer buffer, VkDeviceMemory memory, VkDeviceSize offset) {
    auto* b = reinterpret_cast<Buffer*>(buffer); b->mem = reinterpret_cast<DeviceMemory*>(memory); b->memOffset = offset;
    if (b->usage & ~(VkBufferUsageFlags)VK_BUFFER_USAGE_TRANSFER_DST_BIT) b->mem->gpuReadable = true;
    return VK_SUCCESS;
}
VKFN(VkResult) CreateBufferView(VkDevice, const VkBufferViewCreateInfo* info, const VkAllocationCallbacks*, VkBufferView* pView) {
    auto* v = new BufferView(); v->buffer = reinterpret_cast<Buffer*>(info->buffer); v->format = info->format; v->offset = info->offset; v->range = info->range;
    *pView = reinterpret_cast<VkBufferView>(v); return VK_SUCCESS;
}
VKFN(void) DestroyBufferView(VkDevice, VkBufferView v, const VkAllocationCallbacks*) { delete reinterpret_cast<BufferView*>(v); }
VKFN(VkResult) CreateImage(VkDevice, const VkImageCreateInfo* info, const VkAllocationCallbacks*, VkImage* pImage) {
    if (info->samples != VK_SAMPLE_COUNT_1_BIT) Fatal("multisampled images");
    auto* img = new Image();
    img->format = info->format; img->extent = info->extent; img->mipLevels = info->mipLevels; img->arrayLayers = info->arrayLayers; img->usage = info->usage;
    FillImageLayout(*img);
    *pImage = reinterpret_cast<VkImage>(img);
    return VK_SUCCESS;
}
VKFN(void) DestroyImage(VkDevice, VkImage i, const VkAllocationCallbacks*) { delete reinterpret_cast<Image*>(i); }
VKFN(void) GetImageMemoryRequirements(VkDevice, VkImage image, VkMemoryRequirements* r) { // Image.cpp:28-33
    r->size = reinterpret_cast<Image*>(image)->totalSize; r->alignment = 16; r->memoryTypeBits = 1;
}
VKFN(VkResult) BindImageMemory(VkDevice, VkImage image, VkDeviceMemory memory, VkDeviceSize offset) {
    auto* i = reinterpret_cast<Image*>(image); i->mem = reinterpret_cast<DeviceMemory*>(memory); i->memOffset = offset;
    if (i->usage & ~(VkImageUsageFlags)VK_IMAGE_USAGE_TRANSFER_DST_BIT) i->mem->gpuReadable = true;
    return VK_SUCCESS;
}
VKFN(void) GetImageSubresourceLayout(VkDevice, VkImage image, const VkImageSubresource* sub, VkSubresourceLayout* out) { // Image.cpp:45-83
    auto* i = reinterpret_cast<Image*>(image);
    const MipInfo& m = i->levels[sub->mipLevel];
    out->offset = i->layerSize * sub->arrayLayer + m.offset; out->size = m.levelSize; out->rowPitch = m.stride; out->arrayPitch = i->layerSize; out->depthPitch = m.planeSize;
}
VKFN(VkResult) CreateImageView(VkDevice, const VkImageViewCreateInfo* info, const VkAllocationCallbacks*, VkImageView* pView) {
    auto* v = new ImageView(); v->image = reinterpret_cast<Image*>(info->image); v->format = info->format; v->viewType = info->viewType; v->components = info->components; v->range = info->subresourceRange;
    *pView = reinterpret_cast<VkImageView>(v); return VK_SUCCESS;
}
VKFN(void) DestroyImageView(VkDevice, VkImageView v, const VkAllocationCallbacks*) { delete reinterpret_cast<ImageView*>(v); }
VKFN(VkResult) CreateSampler(VkDevice, const VkSamplerCreateInfo* i, const VkAllocationCallbacks*, VkSampler* pSampler) {
    auto* s = new Sampler();
    s->s.magFilter = i->magFilter; s->s.minFilter = i->minFilter; s->s.mipmapMode = i->mipmapMode;
    s->s.addressModeU = i->addressModeU; s->s.addressModeV = i->addressModeV; s->s.addressModeW = i->addressModeW;
    s->s.mipLodBias = i->mipLodBias; s->s.anisotropyEnable = i->anisotropyEnable; s->s.compareEnable = i->compareEnable; s->s.compareOp = i->compareOp;
    s->s.minLod = i->minLod; s->s.maxLod = i->maxLod; s->s.borderColor = i->borderColor; s->s.unnormalizedCoordinates = i->unnormalizedCoordinates; s->s.flags = i->flags;
    *pSampler = reinterpret_cast<VkSampler>(s); return VK_SUCCESS;
}
VKFN(void) DestroySampler(VkDevice, VkSampler s, const VkAllocationCallbacks*) { delete reinterpret_cast<Sampler*>(s); }

// ---- shaders, layouts, descriptors ----
VKFN(VkResult) CreateShaderModule(VkDevice, const VkShaderModuleCreateInfo* info, const VkAllocationCallbacks*, VkShaderModule* pModule) { // ShaderModule.cpp
    auto* m = new ShaderModule(); m->code.assign(info->pCode, info->pCode + info->codeSize / 4); *pModule = reinterpret_cast<VkShaderModule>(m); return VK_SUCCESS;
}
VKFN(void) DestroyShaderModule(VkDevice, VkShaderModule m, const VkAllocationCallbacks*) { delete reinterpret_cast<ShaderModule*>(m); }
VKFN(VkResult) CreateDescriptorSetLayout(VkDevice, const VkDescriptorSetLayoutCreateInfo* info, const VkAllocationCallbacks*, VkDescriptorSetLayout* pLayout) {
    auto* l = new DescriptorSetLayout(); l->bindings.assign(info->pBindings, info->pBindings + info->bindingCount);
    for (auto& b : l->bindings) {
        if ((b.descriptorType == VK_DESCRIPTOR_TYPE_SAMPLER || b.descriptorType == VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER) && b.pImmutableSamplers) {
            auto& v = l->immutableSamplers[b.binding];
            for (uint32_t i = 0; i < b.descriptorCount; i++) v.push_back(reinterpret_cast<Sampler*>(b.pImmutableSamplers[i]));
        }
        b.pImmutableSamplers = nullptr;
    }
    *pLayout = reinterpret_cast<VkDescriptorSetLayout>(l); return VK_SUCCESS;
}
VKFN(void) DestroyDescriptorSetLayout(VkDevice, VkDescriptorSetLayout l, const VkAllocationCallbacks*) { delete reinterpret_cast<DescriptorSetLayout*>(l); }
VKFN(VkResult) CreatePipelineLayout(VkDevice, const VkPipelineLayoutCreateInfo*, const VkAllocationCallbacks*, VkPipelineLayout* p) { *p = reinterpret_cast<VkPipelineLayout>(new PipelineLayout()); return VK_SUCCESS; }
VKFN(void) DestroyPipelineLayout(VkDevice, VkPipelineLayout l, const VkAllocationCallbacks*) { delete reinterpret_cast<PipelineLayout*>(l); }
VKFN(VkResult) CreateDescriptorPool(VkDevice, const VkDescriptorPoolCreateInfo*, const VkAllocationCallbacks*, VkDescriptorPool* p) { *p = reinterpret_cast<VkDescriptorPool>(new DescriptorPool()); return VK_SUCCESS; }
VKFN(void) DestroyDescriptorPool(VkDevice, VkDescriptorPool p, const VkAllocationCallbacks*) { delete reinterpret_cast<DescriptorPool*>(p); }
VKFN(VkResult) AllocateDescriptorSets(VkDevice, const VkDescriptorSetAllocateInfo* info, VkDescriptorSet* pSets) {
    for (uint32_t i = 0; i < info->descriptorSetCount; i++) {
        auto* s = new DescriptorSet(); s->layout = reinterpret_cast<DescriptorSetLayout*>(info->pSetLayouts[i]);
        for (auto& b : s->layout->bindings) {
            auto& vec = s->bindings[b.binding]; vec.resize(b.descriptorCount);
            auto im = s->layout->immutableSamplers.find(b.binding); // DescriptorSet.cpp:38-48: the set starts out holding them
            if (im != s->layout->immutableSamplers.end()) {
                s->immutable[b.binding] = true;
                for (uint32_t e = 0; e < b.descriptorCount && e < im->second.size(); e++) { vec[e].sampler = im->second[e]; if (b.descriptorType == VK_DESCRIPTOR_TYPE_SAMPLER) vec[e].type = VK_DESCRIPTOR_TYPE_SAMPLER; }
            }
        }
        pSets[i] = reinterpret_cast<VkDescriptorSet>(s);
    }
    return VK_SUCCESS;
}
VKFN(VkResult) FreeDescriptorSets(VkDevice, VkDescriptorPool, uint32_t n, const VkDescriptorSet* sets) { for (uint32_t i = 0; i < n; i++) delete reinterpret_cast<DescriptorSet*>(sets[i]); return VK_SUCCESS; }
VKFN(void) UpdateDescriptorSets(VkDevice, uint32_t writeCount, const VkWriteDescriptorSet* writes, uint32_t copyCount, const VkCopyDescriptorSet*) {
    if (copyCount) Fatal("descriptor copies are not built");
    for (uint32_t w = 0; w < writeCount; w++) {
        const VkWriteDescriptorSet& wr = writes[w];
        auto* set = reinterpret_cast<DescriptorSet*>(wr.dstSet);
        auto& vec = set->bindings[wr.dstBinding];
        if (vec.size() < wr.dstArrayElement + wr.descriptorCount) vec.resize(wr.dstArrayElement + wr.descriptorCount);
        for (uint32_t k = 0; k < wr.descriptorCount; k++) {
            DescriptorValue& v = vec[wr.dstArrayElement + k];
            const bool immutable = set->immutable.count(wr.dstBinding) != 0; // the layout's sampler survives every write (DescriptorSet.cpp:79-101)
            Sampler* const keep = immutable ? v.sampler : nullptr;
            v = DescriptorValue(); v.type = wr.descriptorType;
            switch (wr.descriptorType) {
            case VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER: case VK_DESCRIPTOR_TYPE_STORAGE_BUFFER: case VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER_DYNAMIC: case VK_DESCRIPTOR_TYPE_STORAGE_BUFFER_DYNAMIC:
                v.buffer = reinterpret_cast<Buffer*>(wr.pBufferInfo[k].buffer); v.offset = wr.pBufferInfo[k].offset; v.range = wr.pBufferInfo[k].range; break;
            case VK_DESCRIPTOR_TYPE_UNIFORM_TEXEL_BUFFER: case VK_DESCRIPTOR_TYPE_STORAGE_TEXEL_BUFFER:
                v.bufferView = reinterpret_cast<BufferView*>(wr.pTexelBufferView[k]); break;
            default:
                v.view = reinterpret_cast<ImageView*>(wr.pImageInfo[k].imageView);
                v.sampler = immutable ? keep : (wr.descriptorType == VK_DESCRIPTOR_TYPE_SAMPLED_IMAGE || wr.descriptorType == VK_DESCRIPTOR_TYPE_STORAGE_IMAGE) ? nullptr : reinterpret_cast<Sampler*>(wr.pImageInfo[k].sampler);
                break;
            }
        }
    }
}

// ---- render pass, framebuffer, pipeline ----
VKFN(VkResult) CreateRenderPass(VkDevice, const VkRenderPassCreateInfo* info, const VkAllocationCallbacks*, VkRenderPass* pRenderPass) {
    auto* rp = new RenderPass();
    rp->attachments.assign(info->pAttachments, info->pAttachments + info->attachmentCount);
    for (uint32_t s = 0; s < info->subpassCount; s++) {
        Subpass sp; const VkSubpassDescription& sd = info->pSubpasses[s];
        sp.color.assign(sd.pColorAttachments, sd.pColorAttachments + sd.colorAttachmentCount);
        if (sd.pDepthStencilAttachment) sp.depthStencil = *sd.pDepthStencilAttachment;
        rp->subpasses.push_back(sp);
    }
    *pRenderPass = reinterpret_cast<VkRenderPass>(rp);
    return VK_SUCCESS;
}
VKFN(void) DestroyRenderPass(VkDevice, VkRenderPass r, const VkAllocationCallbacks*) { delete reinterpret_cast<RenderPass*>(r); }
VKFN(VkResult) CreateFramebuffer(VkDevice, const VkFramebufferCreateInfo* info, const VkAllocationCallbacks*, VkFramebuffer* pFramebuffer) {
    auto* fb = new Framebuffer(); fb->width = info->width; fb->height = info->height;
    for (uint32_t i = 0; i < info->attachmentCount; i++) fb->views.push_back(reinterpret_cast<ImageView*>(info->pAttachments[i]));
    *pFramebuffer = reinterpret_cast<VkFramebuffer>(fb); return VK_SUCCESS;
}
VKFN(void) DestroyFramebuffer(VkDevice, VkFramebuffer f, const VkAllocationCallbacks*) { delete reinterpret_cast<Framebuffer*>(f); }
VKFN(VkResult) CreatePipelineCache(VkDevice, const VkPipelineCacheCreateInfo*, const VkAllocationCallbacks*, VkPipelineCache* p) { *p = reinterpret_cast<VkPipelineCache>(new PipelineCache()); return VK_SUCCESS; }
VKFN(void) DestroyPipelineCache(VkDevice, VkPipelineCache c, const VkAllocationCallbacks*) { delete reinterpret_cast<PipelineCache*>(c); }

void FillStage(CpvkShaderStage& out, const VkPipelineShaderStageCreateInfo& st) {
    auto* mod = reinterpret_cast<ShaderModule*>(st.module);
    out.spirv = mod->code.data(); out.wordCount = mod->code.size(); out.entryPoint = st.pName;
    if (const VkSpecializationInfo* si = st.pSpecializationInfo) {
        for (uint32_t i = 0; i < si->mapEntryCount && out.specCount < CPVK_MAX_SPEC_ENTRIES; i++) {
            const VkSpecializationMapEntry& e = si->pMapEntries[i];
            uint32_t v = 0; memcpy(&v, static_cast<const char*>(si->pData) + e.offset, e.size < 4 ? e.size : 4);
            out.spec[out.specCount++] = CpvkSpecEntry{e.constantID, v};
        }
    }
}
VKFN(VkResult) CreateGraphicsPipelines(VkDevice device, VkPipelineCache, uint32_t count, const VkGraphicsPipelineCreateInfo* infos, const VkAllocationCallbacks*, VkPipeline* pPipelines) { // Pipeline.cpp:599-714
    Device* d = Unwrap<Device>(device);
    for (uint32_t n = 0; n < count; n++) {
        const VkGraphicsPipelineCreateInfo& ci = infos[n];
        auto desc = std::make_unique<CpvkPipelineDesc>();
        memset(desc.get(), 0, sizeof(CpvkPipelineDesc));
        for (uint32_t s = 0; s < ci.stageCount; s++) {
            if (ci.pStages[s].stage == VK_SHADER_STAGE_VERTEX_BIT) FillStage(desc->vertex, ci.pStages[s]);
            else if (ci.pStages[s].stage == VK_SHADER_STAGE_FRAGMENT_BIT) FillStage(desc->fragment, ci.pStages[s]);
            else Fatal("tessellation/geometry stages (TODO_ERROR, Draw.cpp:1784-1797)");
        }
        const auto* vi = ci.pVertexInputState;
        desc->bindingCount = vi->vertexBindingDescriptionCount; desc->attributeCount = vi->vertexAttributeDescriptionCount;
        if (desc->bindingCount > CPVK_MAX_VERTEX_BINDINGS || desc->attributeCount > CPVK_MAX_VERTEX_ATTRIBUTES) Fatal("too many vertex bindings/attributes");
        for (uint32_t i = 0; i < desc->bindingCount; i++) desc->bindings[i] = CpvkVertexBinding{vi->pVertexBindingDescriptions[i].binding, vi->pVertexBindingDescriptions[i].stride, (uint32_t)vi->pVertexBindingDescriptions[i].inputRate};
        for (uint32_t i = 0; i < desc->attributeCount; i++) { const auto& a = vi->pVertexAttributeDescriptions[i]; desc->attributes[i] = CpvkVertexAttribute{a.location, a.binding, (uint32_t)a.format, a.offset}; }
        desc->topology = ci.pInputAssemblyState->topology; desc->primitiveRestartEnable = ci.pInputAssemblyState->primitiveRestartEnable;
        const auto* rs = ci.pRasterizationState;
        desc->depthClampEnable = rs->depthClampEnable; desc->rasterizerDiscardEnable = rs->rasterizerDiscardEnable; desc->polygonMode = rs->polygonMode; desc->cullMode = rs->cullMode;
        desc->frontFace = rs->frontFace; desc->depthBiasEnable = rs->depthBiasEnable; desc->lineWidth = rs->lineWidth;
        desc->rasterizationSamples = ci.pMultisampleState ? ci.pMultisampleState->rasterizationSamples : 1;
        if (const auto* ds = ci.pDepthStencilState) {
            desc->depthTestEnable = ds->depthTestEnable; desc->depthWriteEnable = ds->depthWriteEnable; desc->depthCompareOp = ds->depthCompareOp;
            desc->depthBoundsTestEnable = ds->depthBoundsTestEnable; desc->stencilTestEnable = ds->stencilTestEnable;
            auto cp = [](const VkStencilOpState& s) { return CpvkStencilOpState{(uint32_t)s.failOp, (uint32_t)s.passOp, (uint32_t)s.depthFailOp, (uint32_t)s.compareOp, s.compareMask, s.writeMask, s.reference}; };
            desc->front = cp(ds->front); desc->back = cp(ds->back); desc->minDepthBounds = ds->minDepthBounds; desc->maxDepthBounds = ds->maxDepthBounds;
        }
        auto* rp = reinterpret_cast<RenderPass*>(ci.renderPass);
        const Subpass& sp = rp->subpasses[ci.subpass];
        desc->colorAttachmentCount = (uint32_t)sp.color.size();
        if (desc->colorAttachmentCount > CPVK_MAX_COLOR_ATTACHMENTS) Fatal("too many colour attachments");
        for (uint32_t i = 0; i < desc->colorAttachmentCount; i++) {
            desc->colorFormats[i] = sp.color[i].attachment == VK_ATTACHMENT_UNUSED ? 0 : (uint32_t)rp->attachments[sp.color[i].attachment].format;
            desc->blend[i].colorWriteMask = 0xF;
        }
        if (const auto* cb = ci.pColorBlendState) {
            desc->logicOpEnable = cb->logicOpEnable; memcpy(desc->blendConstants, cb->blendConstants, 16);
            for (uint32_t i = 0; i < cb->attachmentCount && i < CPVK_MAX_COLOR_ATTACHMENTS; i++) {
                const auto& b = cb->pAttachments[i];
                desc->blend[i] = CpvkBlendAttachment{b.blendEnable, (uint32_t)b.srcColorBlendFactor, (uint32_t)b.dstColorBlendFactor, (uint32_t)b.colorBlendOp, (uint32_t)b.srcAlphaBlendFactor, (uint32_t)b.dstAlphaBlendFactor, (uint32_t)b.alphaBlendOp, b.colorWriteMask};
            }
        }
        if (sp.depthStencil.attachment != VK_ATTACHMENT_UNUSED) desc->depthStencilFormat = (uint32_t)rp->attachments[sp.depthStencil.attachment].format;
        auto* pipe = new Pipeline();
        if (ci.pDynamicState) for (uint32_t i = 0; i < ci.pDynamicState->dynamicStateCount; i++) {
            const VkDynamicState ds = ci.pDynamicState->pDynamicStates[i];
            if (ds == VK_DYNAMIC_STATE_VIEWPORT) pipe->dynamicViewport = true;
            else if (ds == VK_DYNAMIC_STATE_SCISSOR) { /* recorded, never read (F2) */ }
            else if (ds == VK_DYNAMIC_STATE_LINE_WIDTH || ds == VK_DYNAMIC_STATE_DEPTH_BIAS || ds == VK_DYNAMIC_STATE_BLEND_CONSTANTS) { /* unused by the built subset */ }
            else Fatal("dynamic depth-bounds / stencil state (TODO_ERROR, PipelineCompiler.cpp:1141-1149, :1201-1214)");
        }
        desc->dynamicViewport = pipe->dynamicViewport;
        if (ci.pViewportState) {
            if (ci.pViewportState->viewportCount != 1) Fatal("exactly one viewport (TODO_ERROR, Draw.cpp:1515-1518)");
            if (!pipe->dynamicViewport && ci.pViewportState->pViewports) pipe->staticViewport = ci.pViewportState->pViewports[0];
        }
        if (cpvk_cuda_pipeline_create(d->cuda, desc.get(), &pipe->cuda) != 0) Fatal("vkCreateGraphicsPipelines");
        pPipelines[n] = reinterpret_cast<VkPipeline>(pipe);
    }
    return VK_SUCCESS;
}
VKFN(void) DestroyPipeline(VkDevice device, VkPipeline p, const VkAllocationCallbacks*) {
    if (!p) return;
    auto* pipe = reinterpret_cast<Pipeline*>(p);
    cpvk_cuda_pipeline_destroy(Unwrap<Device>(device)->cuda, pipe->cuda);
    delete pipe;
}

// ---- command buffers ----
VKFN(VkResult) CreateCommandPool(VkDevice device, const VkCommandPoolCreateInfo*, const VkAllocationCallbacks*, VkCommandPool* pPool) {
    auto* p = new CommandPool(); p->device = Unwrap<Device>(device); *pPool = reinterpret_cast<VkCommandPool>(p); return VK_SUCCESS;
}
VKFN(void) DestroyCommandPool(VkDevice, VkCommandPool p, const VkAllocationCallbacks*) { delete reinterpret_cast<CommandPool*>(p); }
VKFN(VkResult) AllocateCommandBuffers(VkDevice device, const VkCommandBufferAllocateInfo* info, VkCommandBuffer* pBuffers) {
    for (uint32_t i = 0; i < info->commandBufferCount; i++) { auto* cb = new Dispatchable<CommandBuffer>(); cb->obj.device = Unwrap<Device>(device); pBuffers[i] = Wrap<VkCommandBuffer>(cb); }
    return VK_SUCCESS;
}
VKFN(void) FreeCommandBuffers(VkDevice, VkCommandPool, uint32_t n, const VkCommandBuffer* bufs) { for (uint32_t i = 0; i < n; i++) if (bufs[i]) delete Outer(Unwrap<CommandBuffer>(bufs[i])); }
VKFN(VkResult) BeginCommandBuffer(VkCommandBuffer cb, const VkCommandBufferBeginInfo*) { Unwrap<CommandBuffer>(cb)->commands.clear(); return VK_SUCCESS; }
VKFN(VkResult) EndCommandBuffer(VkCommandBuffer) { return VK_SUCCESS; }
VKFN(VkResult) ResetCommandBuffer(VkCommandBuffer cb, VkCommandBufferResetFlags) { Unwrap<CommandBuffer>(cb)->commands.clear(); return VK_SUCCESS; }
#define RECORD(cb) Unwrap<CommandBuffer>(cb)->commands.push_back

VKFN(void) CmdBindPipeline(VkCommandBuffer cb, VkPipelineBindPoint bp, VkPipeline pipeline) {
    if (bp != VK_PIPELINE_BIND_POINT_GRAPHICS) Fatal("compute pipelines are outside the draw path");
    auto* p = reinterpret_cast<Pipeline*>(pipeline);
    RECORD(cb)([p](Device& d) { d.state.pipeline = p; });
}
VKFN(void) CmdSetViewport(VkCommandBuffer cb, uint32_t first, uint32_t count, const VkViewport* vps) { // Binding.cpp:202-218
    if (first != 0 || count < 1) return;
    const VkViewport vp = vps[0];
    RECORD(cb)([vp](Device& d) { d.state.viewport = vp; });
}
VKFN(void) CmdSetScissor(VkCommandBuffer, uint32_t, uint32_t, const VkRect2D*) {} // recorded but never read by the reference (F2)
VKFN(void) CmdBindDescriptorSets(VkCommandBuffer cb, VkPipelineBindPoint, VkPipelineLayout, uint32_t firstSet, uint32_t count, const VkDescriptorSet* sets, uint32_t dynCount, const uint32_t* dynOffsets) { // Binding.cpp:58-80
    std::vector<DescriptorSet*> v; for (uint32_t i = 0; i < count; i++) v.push_back(reinterpret_cast<DescriptorSet*>(sets[i]));
    std::vector<uint32_t> dyn(dynOffsets, dynOffsets + dynCount);
    RECORD(cb)([firstSet, v, dyn](Device& d) {
        size_t used = 0;
        for (size_t i = 0; i < v.size() && firstSet + i < 8; i++) {
            d.state.sets[firstSet + i] = v[i];
            size_t need = 0;
            for (auto& kv : v[i]->bindings) for (auto& dv : kv.second) if (dv.type == VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER_DYNAMIC || dv.type == VK_DESCRIPTOR_TYPE_STORAGE_BUFFER_DYNAMIC) need++;
            d.state.dynamicOffsets[firstSet + i].assign(dyn.begin() + std::min(used, dyn.size()), dyn.begin() + std::min(used + need, dyn.size()));
            used += need;
        }
    });
}
VKFN(void) CmdBindVertexBuffers(VkCommandBuffer cb, uint32_t first, uint32_t count, const VkBuffer* bufs, const VkDeviceSize* offsets) { // Binding.cpp:180-199
    std::vector<std::pair<Buffer*, VkDeviceSize>> v; for (uint32_t i = 0; i < count; i++) v.push_back({reinterpret_cast<Buffer*>(bufs[i]), offsets[i]});
    RECORD(cb)([first, v](Device& d) { for (size_t i = 0; i < v.size() && first + i < CPVK_MAX_VERTEX_BINDINGS; i++) { d.state.vertex[first + i].buffer = v[i].first; d.state.vertex[first + i].offset = v[i].second; } });
}
VKFN(void) CmdBindIndexBuffer(VkCommandBuffer cb, VkBuffer buffer, VkDeviceSize offset, VkIndexType type) { // Binding.cpp:113-134
    auto* b = reinterpret_cast<Buffer*>(buffer);
    const uint32_t stride = type == VK_INDEX_TYPE_UINT16 ? 2 : type == VK_INDEX_TYPE_UINT32 ? 4 : type == VK_INDEX_TYPE_UINT8_EXT ? 1 : 0;
    if (!stride) Fatal("index type");
    RECORD(cb)([b, offset, stride](Device& d) { d.state.indexBuffer = b; d.state.indexOffset = offset; d.state.indexStride = stride; });
}
VKFN(void) CmdPushConstants(VkCommandBuffer cb, VkPipelineLayout, VkShaderStageFlags, uint32_t offset, uint32_t size, const void* values) { // CommandBuffer.cpp:552-556
    std::vector<uint8_t> v(static_cast<const uint8_t*>(values), static_cast<const uint8_t*>(values) + size);
    RECORD(cb)([offset, v](Device& d) { if (offset + v.size() <= CPVK_MAX_PUSH_CONSTANT_BYTES) memcpy(d.state.pushConstants + offset, v.data(), v.size()); });
}
VKFN(void) CmdBeginRenderPass(VkCommandBuffer cb, const VkRenderPassBeginInfo* info, VkSubpassContents) {
    auto* rp = reinterpret_cast<RenderPass*>(info->renderPass); auto* fb = reinterpret_cast<Framebuffer*>(info->framebuffer);
    std::vector<VkClearValue> clears(rp->attachments.size());
    for (uint32_t i = 0; i < info->clearValueCount && i < clears.size(); i++) clears[i] = info->pClearValues[i];
    RECORD(cb)([rp, fb, clears](Device& d) { ExecBeginRenderPass(d, rp, fb, clears); });
}
VKFN(void) CmdNextSubpass(VkCommandBuffer cb, VkSubpassContents) { RECORD(cb)([](Device& d) { GatherSubpassAttachments(d); d.state.subpass++; }); }
VKFN(void) CmdEndRenderPass(VkCommandBuffer cb) { RECORD(cb)([](Device& d) { GatherSubpassAttachments(d); d.state.renderPass = nullptr; d.state.framebuffer = nullptr; }); }
VKFN(void) CmdDraw(VkCommandBuffer cb, uint32_t vertexCount, uint32_t instanceCount, uint32_t firstVertex, uint32_t firstInstance) { // Draw.cpp:2350-2354
    RECORD(cb)([=](Device& d) { ExecDraw(d, vertexCount, instanceCount, firstVertex, 0, firstInstance, false); });
}
VKFN(void) CmdDrawIndexed(VkCommandBuffer cb, uint32_t indexCount, uint32_t instanceCount, uint32_t firstIndex, int32_t vertexOffset, uint32_t firstInstance) {
    RECORD(cb)([=](Device& d) { ExecDraw(d, indexCount, instanceCount, firstIndex, vertexOffset, firstInstance, true); });
}
// Indirect draws (Draw.cpp:1874-1996): the parameters are read from the buffer when the command executes.
VKFN(void) CmdDrawIndirect(VkCommandBuffer cb, VkBuffer buffer, VkDeviceSize offset, uint32_t drawCount, uint32_t stride) {
    auto* b = reinterpret_cast<Buffer*>(buffer);
    RECORD(cb)([=](Device& d) {
        for (uint32_t j = 0; j < drawCount; j++) {
            VkDrawIndirectCommand c;
            CU_CHECK(cpvk_cuda_mem_download(d.cuda, &c, b->address(offset + (VkDeviceSize)j * stride), sizeof c));
            ExecDraw(d, c.vertexCount, c.instanceCount, c.firstVertex, 0, c.firstInstance, false);
        }
    });
}
VKFN(void) CmdDrawIndexedIndirect(VkCommandBuffer cb, VkBuffer buffer, VkDeviceSize offset, uint32_t drawCount, uint32_t stride) {
    auto* b = reinterpret_cast<Buffer*>(buffer);
    RECORD(cb)([=](Device& d) {
        for (uint32_t j = 0; j < drawCount; j++) {
            VkDrawIndexedIndirectCommand c;
            CU_CHECK(cpvk_cuda_mem_download(d.cuda, &c, b->address(offset + (VkDeviceSize)j * stride), sizeof c));
            ExecDraw(d, c.indexCount, c.instanceCount, c.firstIndex, c.vertexOffset, c.firstInstance, true);
        }
    });
}
// Indirect-count draws (Draw.cpp:1998-2030): the number of draws comes from a second buffer, capped by maxDrawCount.
VKFN(void) CmdDrawIndirectCount(VkCommandBuffer cb, VkBuffer buffer, VkDeviceSize offset, VkBuffer countBuffer, VkDeviceSize countOffset, uint32_t maxDrawCount, uint32_t stride) {
    auto* b = reinterpret_cast<Buffer*>(buffer); auto* cbuf = reinterpret_cast<Buffer*>(countBuffer);
    RECORD(cb)([=](Device& d) {
        uint32_t count = 0; CU_CHECK(cpvk_cuda_mem_download(d.cuda, &count, cbuf->address(countOffset), 4));
        for (uint32_t j = 0; j < std::min(count, maxDrawCount); j++) {
            VkDrawIndirectCommand c; CU_CHECK(cpvk_cuda_mem_download(d.cuda, &c, b->address(offset + (VkDeviceSize)j * stride), sizeof c));
            ExecDraw(d, c.vertexCount, c.instanceCount, c.firstVertex, 0, c.firstInstance, false);
        }
    });
}
VKFN(void) CmdDrawIndexedIndirectCount(VkCommandBuffer cb, VkBuffer buffer, VkDeviceSize offset, VkBuffer countBuffer, VkDeviceSize countOffset, uint32_t maxDrawCount, uint32_t stride) {
    auto* b = reinterpret_cast<Buffer*>(buffer); auto* cbuf = reinterpret_cast<Buffer*>(countBuffer);
    RECORD(cb)([=](Device& d) {
        uint32_t count = 0; CU_CHECK(cpvk_cuda_mem_download(d.cuda, &count, cbuf->address(countOffset), 4));
        for (uint32_t j = 0; j < std::min(count, maxDrawCount); j++) {
            VkDrawIndexedIndirectCommand c; CU_CHECK(cpvk_cuda_mem_download(d.cuda, &c, b->address(offset + (VkDeviceSize)j * stride), sizeof c));
            ExecDraw(d, c.indexCount, c.instanceCount, c.firstIndex, c.vertexOffset, c.firstInstance, true);
        }
    });
}
// Secondary command buffers (CommandBuffer.cpp:704-731): their commands run in place, on the same device state.
VKFN(void) CmdExecuteCommands(VkCommandBuffer cb, uint32_t n, const VkCommandBuffer* buffers) {
    std::vector<CommandBuffer*> v; for (uint32_t i = 0; i < n; i++) v.push_back(Unwrap<CommandBuffer>(buffers[i]));
    RECORD(cb)([v](Device& d) { for (CommandBuffer* c : v) for (Command& k : c->commands) k(d); });
}
VKFN(void) CmdFillBuffer(VkCommandBuffer cb, VkBuffer dst, VkDeviceSize offset, VkDeviceSize size, uint32_t data) { // CommandBuffer.cpp:278-330: 32-bit words
    auto* t = reinterpret_cast<Buffer*>(dst);
    RECORD(cb)([=](Device& d) {
        const VkDeviceSize bytes = (size == VK_WHOLE_SIZE ? t->size - offset : size) & ~(VkDeviceSize)3;
        CpvkClearValue cv{}; cv.uint32[0] = data;
        // one "image" row of R32_UINT words per piece of at most 1 GiB (the C ABI's sizes are 32-bit; attachments end at 32767 texels per
        // axis only for draws, a clear takes any width)
        for (VkDeviceSize done = 0; done < bytes;) {
            const uint32_t piece = (uint32_t)std::min<VkDeviceSize>(bytes - done, (VkDeviceSize)1 << 30);
            CpvkAttachment a{t->address(offset + done), piece / 4, 1, piece, 98 /* R32_UINT */};
            CU_CHECK(cpvk_cuda_clear(d.cuda, &a, &cv, 0));
            done += piece;
        }
        TouchedByGpu(t->mem);
    });
}
VKFN(void) CmdUpdateBuffer(VkCommandBuffer cb, VkBuffer dst, VkDeviceSize offset, VkDeviceSize size, const void* data) { // CommandBuffer.cpp:241-276: data captured at record time
    auto* t = reinterpret_cast<Buffer*>(dst);
    std::vector<uint8_t> v(static_cast<const uint8_t*>(data), static_cast<const uint8_t*>(data) + size);
    RECORD(cb)([t, offset, v](Device& d) { CU_CHECK(cpvk_cuda_mem_upload(d.cuda, t->address(offset), v.data(), v.size())); CU_CHECK(cpvk_cuda_sync(d.cuda)); TouchedByGpu(t->mem); });
}
VKFN(void) CmdPipelineBarrier(VkCommandBuffer, VkPipelineStageFlags, VkPipelineStageFlags, VkDependencyFlags, uint32_t, const VkMemoryBarrier*, uint32_t, const VkBufferMemoryBarrier*, uint32_t, const VkImageMemoryBarrier*) {} // no-op (F13)

// transfer path: raw row copies and the blit
VKFN(void) CmdCopyBuffer(VkCommandBuffer cb, VkBuffer src, VkBuffer dst, uint32_t n, const VkBufferCopy* regions) {
    auto* s = reinterpret_cast<Buffer*>(src); auto* t = reinterpret_cast<Buffer*>(dst); std::vector<VkBufferCopy> r(regions, regions + n);
    RECORD(cb)([s, t, r](Device& d) {
        // the C ABI's row copies take 32-bit sizes: a region of any VkDeviceSize goes as pieces of at most 1 GiB
        for (auto& c : r)
            for (VkDeviceSize done = 0; done < c.size;) {
                const uint32_t piece = (uint32_t)std::min<VkDeviceSize>(c.size - done, (VkDeviceSize)1 << 30);
                CU_CHECK(cpvk_cuda_copy_rows(d.cuda, t->address(c.dstOffset + done), piece, s->address(c.srcOffset + done), piece, piece, 1));
                done += piece;
            }
        TouchedByGpu(t->mem);
    });
}
VKFN(void) CmdCopyImage(VkCommandBuffer cb, VkImage src, VkImageLayout, VkImage dst, VkImageLayout, uint32_t n, const VkImageCopy* regions) { // CommandBuffer.Copy.cpp:77-200
    auto* s = reinterpret_cast<Image*>(src); auto* t = reinterpret_cast<Image*>(dst); std::vector<VkImageCopy> r(regions, regions + n);
    RECORD(cb)([s, t, r](Device& d) {
        const uint32_t texel = TexelSize(s->format);
        if (texel != TexelSize(t->format)) Fatal("vkCmdCopyImage between different texel sizes");
        for (auto& c : r) for (uint32_t layer = 0; layer < c.srcSubresource.layerCount; layer++) {
            const MipInfo& sm = s->levels[c.srcSubresource.mipLevel]; const MipInfo& tm = t->levels[c.dstSubresource.mipLevel];
            const uint64_t sa = s->address(c.srcSubresource.mipLevel, c.srcSubresource.baseArrayLayer + layer) + c.srcOffset.y * sm.stride + (uint64_t)c.srcOffset.x * texel;
            const uint64_t ta = t->address(c.dstSubresource.mipLevel, c.dstSubresource.baseArrayLayer + layer) + c.dstOffset.y * tm.stride + (uint64_t)c.dstOffset.x * texel;
            CU_CHECK(cpvk_cuda_copy_rows(d.cuda, ta, (uint32_t)tm.stride, sa, (uint32_t)sm.stride, c.extent.width * texel, c.extent.height));
        }
        TouchedByGpu(t->mem);
    });
}
void CopyBufferImage(Device& d, Buffer* b, Image* img, const VkBufferImageCopy& c, bool toImage) { // CommandBuffer.Copy.cpp:461-1083
    uint32_t texel = TexelSize(img->format);
    if (c.imageSubresource.aspectMask == VK_IMAGE_ASPECT_DEPTH_BIT && img->format == VK_FORMAT_D32_SFLOAT_S8_UINT) Fatal("aspect copies of D32_S8");
    const uint32_t rowLen = c.bufferRowLength ? c.bufferRowLength : c.imageExtent.width;
    const uint32_t imgH = c.bufferImageHeight ? c.bufferImageHeight : c.imageExtent.height;
    for (uint32_t layer = 0; layer < c.imageSubresource.layerCount; layer++) {
        const MipInfo& m = img->levels[c.imageSubresource.mipLevel];
        const uint64_t ia = img->address(c.imageSubresource.mipLevel, c.imageSubresource.baseArrayLayer + layer) + c.imageOffset.y * m.stride + (uint64_t)c.imageOffset.x * texel;
        const uint64_t ba = b->address(c.bufferOffset + (uint64_t)layer * rowLen * imgH * texel);
        if (toImage) CU_CHECK(cpvk_cuda_copy_rows(d.cuda, ia, (uint32_t)m.stride, ba, rowLen * texel, c.imageExtent.width * texel, c.imageExtent.height));
        else CU_CHECK(cpvk_cuda_copy_rows(d.cuda, ba, rowLen * texel, ia, (uint32_t)m.stride, c.imageExtent.width * texel, c.imageExtent.height));
    }
    TouchedByGpu(toImage ? img->mem : b->mem);
}
VKFN(void) CmdCopyBufferToImage(VkCommandBuffer cb, VkBuffer src, VkImage dst, VkImageLayout, uint32_t n, const VkBufferImageCopy* regions) {
    auto* b = reinterpret_cast<Buffer*>(src); auto* i = reinterpret_cast<Image*>(dst); std::vector<VkBufferImageCopy> r(regions, regions + n);
    RECORD(cb)([b, i, r](Device& d) { for (auto& c : r) CopyBufferImage(d, b, i, c, true); });
}
VKFN(void) CmdCopyImageToBuffer(VkCommandBuffer cb, VkImage src, VkImageLayout, VkBuffer dst, uint32_t n, const VkBufferImageCopy* regions) {
    auto* b = reinterpret_cast<Buffer*>(dst); auto* i = reinterpret_cast<Image*>(src); std::vector<VkBufferImageCopy> r(regions, regions + n);
    RECORD(cb)([b, i, r](Device& d) { for (auto& c : r) CopyBufferImage(d, b, i, c, false); });
}
VKFN(void) CmdBlitImage(VkCommandBuffer cb, VkImage src, VkImageLayout, VkImage dst, VkImageLayout, uint32_t n, const VkImageBlit* regions, VkFilter filter) { // CommandBuffer.cpp:57-232
    auto* s = reinterpret_cast<Image*>(src); auto* t = reinterpret_cast<Image*>(dst); std::vector<VkImageBlit> r(regions, regions + n);
    RECORD(cb)([s, t, r, filter](Device& d) {
        for (auto& c : r) for (uint32_t layer = 0; layer < c.srcSubresource.layerCount; layer++) {
            if (c.srcOffsets[0].z != 0 || c.srcOffsets[1].z != 1 || c.dstOffsets[0].z != 0 || c.dstOffsets[1].z != 1) Fatal("3-D blits are not built");
            const MipInfo& sm = s->levels[c.srcSubresource.mipLevel]; const MipInfo& tm = t->levels[c.dstSubresource.mipLevel];
            CpvkBlit b{};
            b.src = CpvkAttachment{s->address(c.srcSubresource.mipLevel, c.srcSubresource.baseArrayLayer + layer), sm.width, sm.height, (uint32_t)sm.stride, (uint32_t)s->format};
            b.dst = CpvkAttachment{t->address(c.dstSubresource.mipLevel, c.dstSubresource.baseArrayLayer + layer), tm.width, tm.height, (uint32_t)tm.stride, (uint32_t)t->format};
            b.srcX0 = c.srcOffsets[0].x; b.srcY0 = c.srcOffsets[0].y; b.srcX1 = c.srcOffsets[1].x; b.srcY1 = c.srcOffsets[1].y;
            b.dstX0 = c.dstOffsets[0].x; b.dstY0 = c.dstOffsets[0].y; b.dstX1 = c.dstOffsets[1].x; b.dstY1 = c.dstOffsets[1].y; b.filter = filter;
            CU_CHECK(cpvk_cuda_blit(d.cuda, &b));
        }
        TouchedByGpu(t->mem);
    });
}
VKFN(void) CmdClearColorImage(VkCommandBuffer cb, VkImage image, VkImageLayout, const VkClearColorValue* color, uint32_t n, const VkImageSubresourceRange* ranges) { // Draw.cpp:2032-2100
    auto* img = reinterpret_cast<Image*>(image); CpvkClearValue cv; memcpy(&cv, color, sizeof *color); std::vector<VkImageSubresourceRange> r(ranges, ranges + n);
    RECORD(cb)([img, cv, r](Device& d) {
        for (auto& rg : r) {
            const uint32_t levels = rg.levelCount == VK_REMAINING_MIP_LEVELS ? img->mipLevels - rg.baseMipLevel : rg.levelCount;
            const uint32_t layers = rg.layerCount == VK_REMAINING_ARRAY_LAYERS ? img->arrayLayers - rg.baseArrayLayer : rg.layerCount;
            for (uint32_t la = 0; la < layers; la++) for (uint32_t le = 0; le < levels; le++) {
                const MipInfo& m = img->levels[rg.baseMipLevel + le];
                CpvkAttachment a{img->address(rg.baseMipLevel + le, rg.baseArrayLayer + la), m.width, m.height * m.depth, (uint32_t)m.stride, (uint32_t)img->format};
                CU_CHECK(cpvk_cuda_clear(d.cuda, &a, &cv, 0));
            }
        }
        TouchedByGpu(img->mem);
    });
}

VKFN(void) CmdClearDepthStencilImage(VkCommandBuffer cb, VkImage image, VkImageLayout, const VkClearDepthStencilValue* value, uint32_t n, const VkImageSubresourceRange* ranges) { // Draw.cpp:2102-2224
    auto* img = reinterpret_cast<Image*>(image); CpvkClearValue cv{}; cv.depthStencil.depth = value->depth; cv.depthStencil.stencil = value->stencil;
    std::vector<VkImageSubresourceRange> r(ranges, ranges + n);
    RECORD(cb)([img, cv, r](Device& d) {
        for (auto& rg : r) {
            const uint32_t f = (uint32_t)img->format;
            const VkImageAspectFlags all = (f == 127 ? 0u : (uint32_t)VK_IMAGE_ASPECT_DEPTH_BIT) | (f >= 127 && f <= 130 ? (uint32_t)VK_IMAGE_ASPECT_STENCIL_BIT : 0u);
            if ((rg.aspectMask & all) != all) Fatal("clearing one aspect of a combined depth/stencil image is not built");
            const uint32_t levels = rg.levelCount == VK_REMAINING_MIP_LEVELS ? img->mipLevels - rg.baseMipLevel : rg.levelCount;
            const uint32_t layers = rg.layerCount == VK_REMAINING_ARRAY_LAYERS ? img->arrayLayers - rg.baseArrayLayer : rg.layerCount;
            for (uint32_t la = 0; la < layers; la++) for (uint32_t le = 0; le < levels; le++) {
                const MipInfo& m = img->levels[rg.baseMipLevel + le];
                CpvkAttachment a{img->address(rg.baseMipLevel + le, rg.baseArrayLayer + la), m.width, m.height * m.depth, (uint32_t)m.stride, f};
                CU_CHECK(cpvk_cuda_clear(d.cuda, &a, &cv, 1));
            }
        }
        TouchedByGpu(img->mem);
    });
}
// vkCmdClearAttachments (Draw.cpp:2226-2348): rectangles of the current subpass's attachments, SetPixel per texel.
VKFN(void) CmdClearAttachments(VkCommandBuffer cb, uint32_t na, const VkClearAttachment* atts, uint32_t nr, const VkClearRect* rects) {
    std::vector<VkClearAttachment> a(atts, atts + na); std::vector<VkClearRect> r(rects, rects + nr);
    RECORD(cb)([a, r](Device& d) {
        DeviceState& s = d.state;
        if (!s.renderPass || !s.framebuffer) Fatal("vkCmdClearAttachments outside a render pass");
        const Subpass& sp = s.renderPass->subpasses[s.subpass];
        for (const VkClearAttachment& ca : a) {
            const bool colour = (ca.aspectMask & VK_IMAGE_ASPECT_COLOR_BIT) != 0;
            const uint32_t index = colour ? (ca.colorAttachment < sp.color.size() ? sp.color[ca.colorAttachment].attachment : VK_ATTACHMENT_UNUSED) : sp.depthStencil.attachment;
            if (index == VK_ATTACHMENT_UNUSED) continue;
            ImageView* v = s.framebuffer->views[index];
            const uint32_t texel = TexelSize((uint32_t)s.renderPass->attachments[index].format);
            for (const VkClearRect& cr : r) {
                CpvkAttachment full = AttachmentOf(v); full.format = (uint32_t)s.renderPass->attachments[index].format;
                const uint32_t x0 = (uint32_t)std::max(cr.rect.offset.x, 0), y0 = (uint32_t)std::max(cr.rect.offset.y, 0);
                const uint32_t x1 = std::min(full.width, x0 + cr.rect.extent.width), y1 = std::min(full.height, y0 + cr.rect.extent.height);
                if (x1 <= x0 || y1 <= y0) continue;
                CpvkAttachment sub{full.address + (uint64_t)y0 * full.rowPitch + (uint64_t)x0 * texel, x1 - x0, y1 - y0, full.rowPitch, full.format};
                CpvkClearValue cv; memcpy(&cv, &ca.clearValue, sizeof cv);
                CU_CHECK(cpvk_cuda_clear(d.cuda, &sub, &cv, colour ? 0 : 1));
            }
            TouchedByGpu(v->image->mem);
        }
    });
}

// ---- sync + submit ----
VKFN(VkResult) CreateFence(VkDevice, const VkFenceCreateInfo* info, const VkAllocationCallbacks*, VkFence* pFence) {
    auto* f = new Fence(); f->signaled = (info->flags & VK_FENCE_CREATE_SIGNALED_BIT) != 0; *pFence = reinterpret_cast<VkFence>(f); return VK_SUCCESS;
}
VKFN(void) DestroyFence(VkDevice, VkFence f, const VkAllocationCallbacks*) { delete reinterpret_cast<Fence*>(f); }
VKFN(VkResult) ResetFences(VkDevice, uint32_t n, const VkFence* fences) { for (uint32_t i = 0; i < n; i++) reinterpret_cast<Fence*>(fences[i])->signaled = false; return VK_SUCCESS; }
VKFN(VkResult) GetFenceStatus(VkDevice, VkFence f) { return reinterpret_cast<Fence*>(f)->signaled ? VK_SUCCESS : VK_NOT_READY; }
VKFN(VkResult) WaitForFences(VkDevice, uint32_t n, const VkFence* fences, VkBool32, uint64_t) { // submits are synchronous: a submitted fence is already signalled
    for (uint32_t i = 0; i < n; i++) if (!reinterpret_cast<Fence*>(fences[i])->signaled) return VK_TIMEOUT;
    return VK_SUCCESS;
}
VKFN(VkResult) CreateSemaphore(VkDevice, const VkSemaphoreCreateInfo*, const VkAllocationCallbacks*, VkSemaphore* p) { *p = reinterpret_cast<VkSemaphore>(new Semaphore()); return VK_SUCCESS; }
VKFN(void) DestroySemaphore(VkDevice, VkSemaphore s, const VkAllocationCallbacks*) { delete reinterpret_cast<Semaphore*>(s); }

VKFN(VkResult) QueueSubmit(VkQueue queue, uint32_t submitCount, const VkSubmitInfo* submits, VkFence fence) { // Queue.cpp:11-77: execute inline, then signal
    Device& d = *Unwrap<Queue>(queue)->device;
    UploadHostWrites(d);
    for (uint32_t s = 0; s < submitCount; s++)
        for (uint32_t c = 0; c < submits[s].commandBufferCount; c++)
            for (const Command& cmd : Unwrap<CommandBuffer>(submits[s].pCommandBuffers[c])->commands) cmd(d); // RunCommands (CommandBuffer.cpp:21-30)
    DownloadDeviceWrites(d);
    if (cpvk_cuda_sync(d.cuda) != 0) return VK_ERROR_DEVICE_LOST;
    if (fence) reinterpret_cast<Fence*>(fence)->signaled = true;
    return VK_SUCCESS;
}

// ---- proc-addr table (Extensions.cpp / VulkanFunctions.h) ----
VKFN(PFN_vkVoidFunction) GetDeviceProcAddr(VkDevice, const char* name);
VKFN(PFN_vkVoidFunction) GetInstanceProcAddr(VkInstance, const char* name);
struct Entry { const char* name; PFN_vkVoidFunction fn; };
#define E(n) {"vk" #n, reinterpret_cast<PFN_vkVoidFunction>(n)}
const Entry kEntries[] = {
    E(CreateInstance), E(DestroyInstance), E(EnumeratePhysicalDevices), E(EnumerateInstanceExtensionProperties), E(EnumerateInstanceLayerProperties),
    E(EnumerateDeviceExtensionProperties), E(EnumerateInstanceVersion), E(GetPhysicalDeviceProperties), E(GetPhysicalDeviceFeatures),
    E(GetPhysicalDeviceQueueFamilyProperties), E(GetPhysicalDeviceMemoryProperties), E(GetPhysicalDeviceFormatProperties), E(CreateDevice), E(DestroyDevice),
    E(GetDeviceQueue), E(DeviceWaitIdle), E(QueueWaitIdle), E(AllocateMemory), E(FreeMemory), E(MapMemory), E(UnmapMemory), E(FlushMappedMemoryRanges),
    E(InvalidateMappedMemoryRanges), E(CreateBuffer), E(DestroyBuffer), E(GetBufferMemoryRequirements), E(BindBufferMemory), E(CreateBufferView), E(DestroyBufferView),
    E(CreateImage), E(DestroyImage), E(GetImageMemoryRequirements), E(BindImageMemory), E(GetImageSubresourceLayout), E(CreateImageView), E(DestroyImageView),
    E(CreateSampler), E(DestroySampler), E(CreateShaderModule), E(DestroyShaderModule), E(CreateDescriptorSetLayout), E(DestroyDescriptorSetLayout),
    E(CreatePipelineLayout), E(DestroyPipelineLayout), E(CreateDescriptorPool), E(DestroyDescriptorPool), E(AllocateDescriptorSets), E(FreeDescriptorSets),
    E(UpdateDescriptorSets), E(CreateRenderPass), E(DestroyRenderPass), E(CreateFramebuffer), E(DestroyFramebuffer), E(CreatePipelineCache), E(DestroyPipelineCache),
    E(CreateGraphicsPipelines), E(DestroyPipeline), E(CreateCommandPool), E(DestroyCommandPool), E(AllocateCommandBuffers), E(FreeCommandBuffers),
    E(BeginCommandBuffer), E(EndCommandBuffer), E(ResetCommandBuffer), E(CmdBindPipeline), E(CmdSetViewport), E(CmdSetScissor), E(CmdBindDescriptorSets),
    E(CmdBindVertexBuffers), E(CmdBindIndexBuffer), E(CmdPushConstants), E(CmdBeginRenderPass), E(CmdNextSubpass), E(CmdEndRenderPass), E(CmdDraw), E(CmdDrawIndexed),
    E(CmdDrawIndirect), E(CmdDrawIndexedIndirect), E(CmdDrawIndirectCount), E(CmdDrawIndexedIndirectCount),
    {"vkCmdDrawIndirectCountKHR", reinterpret_cast<PFN_vkVoidFunction>(CmdDrawIndirectCount)}, {"vkCmdDrawIndexedIndirectCountKHR", reinterpret_cast<PFN_vkVoidFunction>(CmdDrawIndexedIndirectCount)}, E(CmdExecuteCommands), E(CmdFillBuffer), E(CmdUpdateBuffer), E(CmdClearDepthStencilImage), E(CmdClearAttachments),
    E(CmdPipelineBarrier), E(CmdCopyBuffer), E(CmdCopyImage), E(CmdCopyBufferToImage), E(CmdCopyImageToBuffer), E(CmdBlitImage), E(CmdClearColorImage),
    E(CreateFence), E(DestroyFence), E(ResetFences), E(GetFenceStatus), E(WaitForFences), E(CreateSemaphore), E(DestroySemaphore), E(QueueSubmit),
    E(GetDeviceProcAddr), E(GetInstanceProcAddr),
};
PFN_vkVoidFunction Lookup(const char* name) {
    for (const Entry& e : kEntries) if (!strcmp(e.name, name)) return e.fn;
    return nullptr;
}
VKFN(PFN_vkVoidFunction) GetDeviceProcAddr(VkDevice, const char* name) { return Lookup(name); }
VKFN(PFN_vkVoidFunction) GetInstanceProcAddr(VkInstance, const char* name) { return Lookup(name); }

} // namespace

extern "C" {
__attribute__((visibility("default"))) VkResult vk_icdNegotiateLoaderICDInterfaceVersion(uint32_t* pSupportedVersion) { // CPVulkan.cpp:95-104
    if (*pSupportedVersion > 5) *pSupportedVersion = 5;
    return VK_SUCCESS;
}
__attribute__((visibility("default"))) PFN_vkVoidFunction vk_icdGetInstanceProcAddr(VkInstance, const char* pName) { return Lookup(pName); } // CPVulkan.cpp:14-88
__attribute__((visibility("default"))) PFN_vkVoidFunction vk_icdGetPhysicalDeviceProcAddr(VkInstance, const char*) { return nullptr; }     // CPVulkan.cpp:90-93
}
