// cpvk_harness.cpp — stands in for the Vulkan loader AND for the LunarG sample applications (Samples/15-draw_cube,
// Samples/draw_textured_cube, Samples/utils/util_init.cpp), neither of which can be built here (no Vulkan SDK,
// no glslang, no window system: SURVEY App. B, H5).
//
// Loader part: reads the ICD manifest named by VK_ICD_FILENAMES (CPVulkan/CPVulkan.json:1-6), dlopen()s its
// library_path, negotiates the interface version and resolves every entry point through vk_icdGetInstanceProcAddr —
// the outer drop-in boundary of SURVEY §8(b).
// Application part: the sample call sequence (init_instance .. init_pipeline, record, submit, wait, read back),
// driven by a scene description exported by cpvulkan_b200/scenes.py so tests can render the very same inputs with
// the CPU oracle and byte-compare. Off-screen: the "swapchain image" is an ordinary colour image (Image.cpp:17-21).
//
//   cpvk_harness <scene dir> <out dir> [--frames K] [--blit W H FORMAT FILTER]
//   [--indirect] draw through vkCmdDraw[Indexed]Indirect (parameters written through a mapping, slot 1 of a 3-slot buffer)
//   [--secondary] record the bind + draw commands into a secondary command buffer and vkCmdExecuteCommands it
//   [--update-buffers] fill the uniform buffers with vkCmdFillBuffer(0) + vkCmdUpdateBuffer instead of a mapping
//   [--clear-rect X Y W H] vkCmdClearAttachments of that rectangle (colour 0.5,0.25,0.75,1 / depth 0.5) after the draw
//   --blit: after the render pass, vkCmdBlitImage the colour image to a W x H image of FORMAT (copy_blit_image.cpp:146-190),
//   vkCmdCopyImage that to a second image (:192-222) and read the copy back as blit.bin
#include <dlfcn.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <algorithm>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/cpvk_vulkan.h"

#define DIE(...) do { fprintf(stderr, "cpvk_harness: " __VA_ARGS__); fprintf(stderr, "\n"); exit(2); } while (0)
#define VK(expr) do { VkResult r_ = (expr); if (r_ != VK_SUCCESS) DIE("%s = %d", #expr, (int)r_); } while (0)

static PFN_vkVoidFunction (*gipa)(VkInstance, const char*);
#define DECL(name, ret, ...) typedef ret (*PFN_##name)(__VA_ARGS__); static PFN_##name name;
DECL(vkCreateInstance, VkResult, const VkInstanceCreateInfo*, const VkAllocationCallbacks*, VkInstance*)
DECL(vkDestroyInstance, void, VkInstance, const VkAllocationCallbacks*)
DECL(vkEnumeratePhysicalDevices, VkResult, VkInstance, uint32_t*, VkPhysicalDevice*)
DECL(vkGetPhysicalDeviceProperties, void, VkPhysicalDevice, VkPhysicalDeviceProperties*)
DECL(vkGetPhysicalDeviceMemoryProperties, void, VkPhysicalDevice, VkPhysicalDeviceMemoryProperties*)
DECL(vkGetPhysicalDeviceQueueFamilyProperties, void, VkPhysicalDevice, uint32_t*, VkQueueFamilyProperties*)
DECL(vkCreateDevice, VkResult, VkPhysicalDevice, const VkDeviceCreateInfo*, const VkAllocationCallbacks*, VkDevice*)
DECL(vkDestroyDevice, void, VkDevice, const VkAllocationCallbacks*)
DECL(vkGetDeviceQueue, void, VkDevice, uint32_t, uint32_t, VkQueue*)
DECL(vkAllocateMemory, VkResult, VkDevice, const VkMemoryAllocateInfo*, const VkAllocationCallbacks*, VkDeviceMemory*)
DECL(vkMapMemory, VkResult, VkDevice, VkDeviceMemory, VkDeviceSize, VkDeviceSize, VkMemoryMapFlags, void**)
DECL(vkUnmapMemory, void, VkDevice, VkDeviceMemory)
DECL(vkCreateBuffer, VkResult, VkDevice, const VkBufferCreateInfo*, const VkAllocationCallbacks*, VkBuffer*)
DECL(vkGetBufferMemoryRequirements, void, VkDevice, VkBuffer, VkMemoryRequirements*)
DECL(vkBindBufferMemory, VkResult, VkDevice, VkBuffer, VkDeviceMemory, VkDeviceSize)
DECL(vkCreateImage, VkResult, VkDevice, const VkImageCreateInfo*, const VkAllocationCallbacks*, VkImage*)
DECL(vkGetImageMemoryRequirements, void, VkDevice, VkImage, VkMemoryRequirements*)
DECL(vkBindImageMemory, VkResult, VkDevice, VkImage, VkDeviceMemory, VkDeviceSize)
DECL(vkGetImageSubresourceLayout, void, VkDevice, VkImage, const VkImageSubresource*, VkSubresourceLayout*)
DECL(vkCreateImageView, VkResult, VkDevice, const VkImageViewCreateInfo*, const VkAllocationCallbacks*, VkImageView*)
DECL(vkCreateSampler, VkResult, VkDevice, const VkSamplerCreateInfo*, const VkAllocationCallbacks*, VkSampler*)
DECL(vkCreateBufferView, VkResult, VkDevice, const VkBufferViewCreateInfo*, const VkAllocationCallbacks*, VkBufferView*)
DECL(vkCreateShaderModule, VkResult, VkDevice, const VkShaderModuleCreateInfo*, const VkAllocationCallbacks*, VkShaderModule*)
DECL(vkCreateDescriptorSetLayout, VkResult, VkDevice, const VkDescriptorSetLayoutCreateInfo*, const VkAllocationCallbacks*, VkDescriptorSetLayout*)
DECL(vkCreatePipelineLayout, VkResult, VkDevice, const VkPipelineLayoutCreateInfo*, const VkAllocationCallbacks*, VkPipelineLayout*)
DECL(vkCreateDescriptorPool, VkResult, VkDevice, const VkDescriptorPoolCreateInfo*, const VkAllocationCallbacks*, VkDescriptorPool*)
DECL(vkAllocateDescriptorSets, VkResult, VkDevice, const VkDescriptorSetAllocateInfo*, VkDescriptorSet*)
DECL(vkUpdateDescriptorSets, void, VkDevice, uint32_t, const VkWriteDescriptorSet*, uint32_t, const VkCopyDescriptorSet*)
DECL(vkCreateRenderPass, VkResult, VkDevice, const VkRenderPassCreateInfo*, const VkAllocationCallbacks*, VkRenderPass*)
DECL(vkCreateFramebuffer, VkResult, VkDevice, const VkFramebufferCreateInfo*, const VkAllocationCallbacks*, VkFramebuffer*)
DECL(vkCreateGraphicsPipelines, VkResult, VkDevice, VkPipelineCache, uint32_t, const VkGraphicsPipelineCreateInfo*, const VkAllocationCallbacks*, VkPipeline*)
DECL(vkCreateCommandPool, VkResult, VkDevice, const VkCommandPoolCreateInfo*, const VkAllocationCallbacks*, VkCommandPool*)
DECL(vkAllocateCommandBuffers, VkResult, VkDevice, const VkCommandBufferAllocateInfo*, VkCommandBuffer*)
DECL(vkBeginCommandBuffer, VkResult, VkCommandBuffer, const VkCommandBufferBeginInfo*)
DECL(vkEndCommandBuffer, VkResult, VkCommandBuffer)
DECL(vkCmdBeginRenderPass, void, VkCommandBuffer, const VkRenderPassBeginInfo*, VkSubpassContents)
DECL(vkCmdEndRenderPass, void, VkCommandBuffer)
DECL(vkCmdBindPipeline, void, VkCommandBuffer, VkPipelineBindPoint, VkPipeline)
DECL(vkCmdBindDescriptorSets, void, VkCommandBuffer, VkPipelineBindPoint, VkPipelineLayout, uint32_t, uint32_t, const VkDescriptorSet*, uint32_t, const uint32_t*)
DECL(vkCmdPushConstants, void, VkCommandBuffer, VkPipelineLayout, VkShaderStageFlags, uint32_t, uint32_t, const void*)
DECL(vkCmdBindVertexBuffers, void, VkCommandBuffer, uint32_t, uint32_t, const VkBuffer*, const VkDeviceSize*)
DECL(vkCmdBindIndexBuffer, void, VkCommandBuffer, VkBuffer, VkDeviceSize, VkIndexType)
DECL(vkCmdSetViewport, void, VkCommandBuffer, uint32_t, uint32_t, const VkViewport*)
DECL(vkCmdSetScissor, void, VkCommandBuffer, uint32_t, uint32_t, const VkRect2D*)
DECL(vkCmdDraw, void, VkCommandBuffer, uint32_t, uint32_t, uint32_t, uint32_t)
DECL(vkCmdDrawIndexed, void, VkCommandBuffer, uint32_t, uint32_t, uint32_t, int32_t, uint32_t)
DECL(vkCmdCopyImageToBuffer, void, VkCommandBuffer, VkImage, VkImageLayout, VkBuffer, uint32_t, const VkBufferImageCopy*)
DECL(vkCmdDrawIndirect, void, VkCommandBuffer, VkBuffer, VkDeviceSize, uint32_t, uint32_t)
DECL(vkCmdDrawIndexedIndirect, void, VkCommandBuffer, VkBuffer, VkDeviceSize, uint32_t, uint32_t)
DECL(vkCmdDrawIndirectCount, void, VkCommandBuffer, VkBuffer, VkDeviceSize, VkBuffer, VkDeviceSize, uint32_t, uint32_t)
DECL(vkCmdDrawIndexedIndirectCount, void, VkCommandBuffer, VkBuffer, VkDeviceSize, VkBuffer, VkDeviceSize, uint32_t, uint32_t)
DECL(vkCmdExecuteCommands, void, VkCommandBuffer, uint32_t, const VkCommandBuffer*)
DECL(vkCmdUpdateBuffer, void, VkCommandBuffer, VkBuffer, VkDeviceSize, VkDeviceSize, const void*)
DECL(vkCmdFillBuffer, void, VkCommandBuffer, VkBuffer, VkDeviceSize, VkDeviceSize, uint32_t)
DECL(vkCmdClearAttachments, void, VkCommandBuffer, uint32_t, const VkClearAttachment*, uint32_t, const VkClearRect*)
DECL(vkCmdBlitImage, void, VkCommandBuffer, VkImage, VkImageLayout, VkImage, VkImageLayout, uint32_t, const VkImageBlit*, VkFilter)
DECL(vkCmdCopyImage, void, VkCommandBuffer, VkImage, VkImageLayout, VkImage, VkImageLayout, uint32_t, const VkImageCopy*)
DECL(vkCreateFence, VkResult, VkDevice, const VkFenceCreateInfo*, const VkAllocationCallbacks*, VkFence*)
DECL(vkResetFences, VkResult, VkDevice, uint32_t, const VkFence*)
DECL(vkWaitForFences, VkResult, VkDevice, uint32_t, const VkFence*, VkBool32, uint64_t)
DECL(vkQueueSubmit, VkResult, VkQueue, uint32_t, const VkSubmitInfo*, VkFence)
DECL(vkDeviceWaitIdle, VkResult, VkDevice)

static void LoadIcd() {
    const char* manifest = getenv("VK_ICD_FILENAMES");
    if (!manifest) DIE("VK_ICD_FILENAMES is not set");
    std::ifstream f(manifest);
    if (!f) DIE("cannot open manifest %s", manifest);
    std::stringstream ss; ss << f.rdbuf();
    const std::string text = ss.str();
    size_t k = text.find("\"library_path\"");
    if (k == std::string::npos) DIE("manifest has no library_path");
    k = text.find(':', k); size_t a = text.find('"', k), b = text.find('"', a + 1);
    std::string lib = text.substr(a + 1, b - a - 1);
    if (lib[0] != '/') { std::string dir(manifest); size_t s = dir.rfind('/'); lib = (s == std::string::npos ? std::string(".") : dir.substr(0, s)) + "/" + lib; }
    void* h = dlopen(lib.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h) DIE("dlopen %s: %s", lib.c_str(), dlerror());
    auto negotiate = (VkResult (*)(uint32_t*))dlsym(h, "vk_icdNegotiateLoaderICDInterfaceVersion");
    gipa = (PFN_vkVoidFunction (*)(VkInstance, const char*))dlsym(h, "vk_icdGetInstanceProcAddr");
    if (!negotiate || !gipa) DIE("ICD lacks the vk_icd* exports");
    uint32_t v = 5; VK(negotiate(&v));
#define GET(name) name = (PFN_##name)gipa(nullptr, #name); if (!name) DIE("ICD does not implement %s", #name);
    GET(vkCreateInstance) GET(vkDestroyInstance) GET(vkEnumeratePhysicalDevices) GET(vkGetPhysicalDeviceProperties) GET(vkGetPhysicalDeviceMemoryProperties)
    GET(vkGetPhysicalDeviceQueueFamilyProperties) GET(vkCreateDevice) GET(vkDestroyDevice) GET(vkGetDeviceQueue) GET(vkAllocateMemory) GET(vkMapMemory) GET(vkUnmapMemory)
    GET(vkCreateBuffer) GET(vkGetBufferMemoryRequirements) GET(vkBindBufferMemory) GET(vkCreateImage) GET(vkGetImageMemoryRequirements) GET(vkBindImageMemory)
    GET(vkGetImageSubresourceLayout) GET(vkCreateImageView) GET(vkCreateSampler) GET(vkCreateBufferView) GET(vkCreateShaderModule) GET(vkCreateDescriptorSetLayout) GET(vkCreatePipelineLayout)
    GET(vkCreateDescriptorPool) GET(vkAllocateDescriptorSets) GET(vkUpdateDescriptorSets) GET(vkCreateRenderPass) GET(vkCreateFramebuffer) GET(vkCreateGraphicsPipelines)
    GET(vkCreateCommandPool) GET(vkAllocateCommandBuffers) GET(vkBeginCommandBuffer) GET(vkEndCommandBuffer) GET(vkCmdBeginRenderPass) GET(vkCmdEndRenderPass)
    GET(vkCmdBindPipeline) GET(vkCmdBindDescriptorSets) GET(vkCmdPushConstants) GET(vkCmdBindVertexBuffers) GET(vkCmdBindIndexBuffer) GET(vkCmdSetViewport) GET(vkCmdSetScissor) GET(vkCmdDraw)
    GET(vkCmdDrawIndexed) GET(vkCmdCopyImageToBuffer) GET(vkCmdBlitImage) GET(vkCmdCopyImage) GET(vkCmdDrawIndirect) GET(vkCmdDrawIndexedIndirect) GET(vkCmdDrawIndirectCount) GET(vkCmdDrawIndexedIndirectCount) GET(vkCmdExecuteCommands) GET(vkCmdUpdateBuffer) GET(vkCmdFillBuffer) GET(vkCmdClearAttachments) GET(vkCreateFence) GET(vkResetFences) GET(vkWaitForFences) GET(vkQueueSubmit) GET(vkDeviceWaitIdle)
}

static std::vector<uint8_t> ReadFile(const std::string& p) {
    std::ifstream f(p, std::ios::binary);
    if (!f) DIE("cannot read %s", p.c_str());
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static uint32_t TexelSize(uint32_t f) {
    if (f >= 9 && f <= 50) { const uint32_t fam = (f - 9) / 7; return fam == 0 ? 1 : fam == 1 ? 2 : (fam == 2 || fam == 3) ? 3 : 4; }
    if (f >= 51 && f <= 69) return 4;
    if (f >= 70 && f <= 97) return 2 * ((f - 70) / 7 + 1);
    if (f >= 98 && f <= 109) return 4 * ((f - 98) / 3 + 1);
    switch (f) { case 124: return 2; case 125: case 126: case 129: return 4; case 127: return 1; case 128: return 3; case 130: return 8; default: return 0; }
}

struct SceneDesc {
    std::string vs, fs;
    std::vector<VkVertexInputBindingDescription> bindings;
    std::vector<VkVertexInputAttributeDescription> attributes;
    uint32_t topology = 3, cull = 0, front = 0, depthTest = 0, depthWrite = 0, depthOp = 3, writeMask = 0xF;
    float lineWidth = 1.0f;
    bool blend = false; uint32_t bl[6] = {};
    struct Buf { std::string file; uint64_t size; }; std::map<std::string, Buf> buffers;
    std::map<uint32_t, std::string> vertexBuffers;
    std::string indexBuffer; uint32_t indexStride = 0;
    // dynRange != 0: bound as UNIFORM_BUFFER_DYNAMIC over [0, dynRange) with dynOffset passed to vkCmdBindDescriptorSets (Samples/dynamic_uniform)
    struct Uni { uint32_t set, binding; std::string name; uint32_t dynOffset = 0, dynRange = 0; }; std::vector<Uni> uniforms;
    std::vector<uint8_t> pushConstants;                                        // vkCmdPushConstants (Samples/push_constants)
    struct Spec { uint32_t stage, id, value; }; std::vector<Spec> specs;       // VkSpecializationInfo (Samples/spirv_specialization)
    // samplerBinding >= 0: texture2D + sampler bound separately (Samples/separate_image_sampler); immutableSampler: the sampler
    // lives in the set layout and the descriptor write carries none (Samples/immutable_sampler)
    struct Tex { uint32_t set, binding, format, w, h, filter, address; std::string file; int samplerBinding = -1; int immutableSampler = 0; int inputAttachment = 0; }; std::vector<Tex> textures;
    struct TexelBuf { uint32_t set, binding, format; std::string name; }; std::vector<TexelBuf> texelBuffers;
    uint32_t colorFormat = 0, width = 0, height = 0; float clearColor[4] = {};
    uint32_t depthFormat = 0; float clearDepth = 1; uint32_t clearStencil = 0;
    float viewport[6] = {};
    uint32_t count = 0, instances = 1, first = 0, firstInstance = 0; int32_t vertexOffset = 0;
};

static SceneDesc ParseScene(const std::string& dir) {
    SceneDesc s;
    std::ifstream f(dir + "/scene.txt");
    if (!f) DIE("cannot open %s/scene.txt", dir.c_str());
    std::string line;
    while (std::getline(f, line)) {
        std::istringstream is(line); std::string key; is >> key;
        if (key == "vs") is >> s.vs; else if (key == "fs") is >> s.fs;
        else if (key == "binding") { VkVertexInputBindingDescription b{}; uint32_t rate; is >> b.binding >> b.stride >> rate; b.inputRate = (VkVertexInputRate)rate; s.bindings.push_back(b); }
        else if (key == "attribute") { VkVertexInputAttributeDescription a{}; uint32_t fmt; is >> a.location >> a.binding >> fmt >> a.offset; a.format = (VkFormat)fmt; s.attributes.push_back(a); }
        else if (key == "line_width") is >> s.lineWidth;
        else if (key == "topology") is >> s.topology; else if (key == "cull") is >> s.cull; else if (key == "front") is >> s.front;
        else if (key == "depth_test") is >> s.depthTest; else if (key == "depth_write") is >> s.depthWrite; else if (key == "depth_op") is >> s.depthOp;
        else if (key == "write_mask") is >> s.writeMask;
        else if (key == "blend") { s.blend = true; for (auto& v : s.bl) is >> v; }
        else if (key == "buffer") { std::string n; SceneDesc::Buf b; is >> n >> b.file >> b.size; s.buffers[n] = b; }
        else if (key == "vertex_buffer") { uint32_t b; std::string n; is >> b >> n; s.vertexBuffers[b] = n; }
        else if (key == "index_buffer") is >> s.indexBuffer >> s.indexStride;
        else if (key == "uniform") { SceneDesc::Uni u; is >> u.set >> u.binding >> u.name; if (!(is >> u.dynOffset >> u.dynRange)) { u.dynOffset = 0; u.dynRange = 0; } s.uniforms.push_back(u); }
        else if (key == "push_constants") { size_t n; std::string hex; is >> n >> hex; for (size_t i = 0; i + 1 < hex.size() && i / 2 < n; i += 2) s.pushConstants.push_back((uint8_t)std::stoul(hex.substr(i, 2), nullptr, 16)); }
        else if (key == "spec") { SceneDesc::Spec sp; is >> sp.stage >> sp.id >> sp.value; s.specs.push_back(sp); }
        else if (key == "texel_buffer") { SceneDesc::TexelBuf t; is >> t.set >> t.binding >> t.name >> t.format; s.texelBuffers.push_back(t); }
        else if (key == "texture") { SceneDesc::Tex t; is >> t.set >> t.binding >> t.format >> t.w >> t.h >> t.filter >> t.address >> t.file; if (!(is >> t.samplerBinding)) t.samplerBinding = -1; if (!(is >> t.immutableSampler)) t.immutableSampler = 0; if (!(is >> t.inputAttachment)) t.inputAttachment = 0; s.textures.push_back(t); }
        else if (key == "color") { is >> s.colorFormat >> s.width >> s.height; for (auto& c : s.clearColor) is >> c; }
        else if (key == "depth") is >> s.depthFormat >> s.clearDepth >> s.clearStencil;
        else if (key == "viewport") for (auto& v : s.viewport) is >> v;
        else if (key == "draw") is >> s.count >> s.instances >> s.first >> s.vertexOffset >> s.firstInstance;
    }
    return s;
}

struct App {
    VkInstance instance; VkPhysicalDevice gpu; VkDevice device; VkQueue queue; VkCommandPool pool; VkCommandBuffer cmd;
    VkDeviceMemory Alloc(VkDeviceSize size) {
        VkMemoryAllocateInfo ai{VK_STRUCTURE_TYPE_MEMORY_ALLOCATE_INFO, nullptr, size, 0};
        VkDeviceMemory m; VK(vkAllocateMemory(device, &ai, nullptr, &m)); return m;
    }
    VkBuffer MakeBuffer(VkDeviceSize size, VkBufferUsageFlags usage, VkDeviceMemory* mem) {
        VkBufferCreateInfo bi{VK_STRUCTURE_TYPE_BUFFER_CREATE_INFO, nullptr, 0, size, usage, VK_SHARING_MODE_EXCLUSIVE, 0, nullptr};
        VkBuffer b; VK(vkCreateBuffer(device, &bi, nullptr, &b));
        VkMemoryRequirements r; vkGetBufferMemoryRequirements(device, b, &r);
        *mem = Alloc(r.size); VK(vkBindBufferMemory(device, b, *mem, 0));
        return b;
    }
    VkImage MakeImage(uint32_t format, uint32_t w, uint32_t h, VkImageUsageFlags usage, VkImageTiling tiling, VkDeviceMemory* mem) {
        VkImageCreateInfo ii{VK_STRUCTURE_TYPE_IMAGE_CREATE_INFO, nullptr, 0, VK_IMAGE_TYPE_2D, (VkFormat)format, {w, h, 1}, 1, 1, VK_SAMPLE_COUNT_1_BIT, tiling, usage,
                             VK_SHARING_MODE_EXCLUSIVE, 0, nullptr, VK_IMAGE_LAYOUT_UNDEFINED};
        VkImage img; VK(vkCreateImage(device, &ii, nullptr, &img));
        VkMemoryRequirements r; vkGetImageMemoryRequirements(device, img, &r);
        *mem = Alloc(r.size); VK(vkBindImageMemory(device, img, *mem, 0));
        return img;
    }
    VkImageView MakeView(VkImage img, uint32_t format, VkImageAspectFlags aspect) {
        VkImageViewCreateInfo vi{VK_STRUCTURE_TYPE_IMAGE_VIEW_CREATE_INFO, nullptr, 0, img, VK_IMAGE_VIEW_TYPE_2D, (VkFormat)format,
                                 {VK_COMPONENT_SWIZZLE_R, VK_COMPONENT_SWIZZLE_G, VK_COMPONENT_SWIZZLE_B, VK_COMPONENT_SWIZZLE_A}, {aspect, 0, 1, 0, 1}};
        VkImageView v; VK(vkCreateImageView(device, &vi, nullptr, &v)); return v;
    }
};

int main(int argc, char** argv) {
    if (argc < 3) DIE("usage: cpvk_harness <scene dir> <out dir> [--frames K]");
    const std::string sceneDir = argv[1], outDir = argv[2];
    int frames = 1;
    for (int i = 3; i + 1 < argc; i++) if (!strcmp(argv[i], "--frames")) frames = atoi(argv[i + 1]);
    bool indirect = false, indirectCount = false, secondary = false, updateBuffers = false; int clearRect[4] = {0, 0, 0, 0};
    for (int i = 3; i < argc; i++) {
        if (!strcmp(argv[i], "--indirect")) indirect = true;
        if (!strcmp(argv[i], "--indirect-count")) indirect = indirectCount = true; // the draw count (1, capped at 3) is word 0 of the same buffer
        if (!strcmp(argv[i], "--secondary")) secondary = true;
        if (!strcmp(argv[i], "--update-buffers")) updateBuffers = true;
        if (!strcmp(argv[i], "--clear-rect") && i + 4 < argc) for (int k = 0; k < 4; k++) clearRect[k] = atoi(argv[i + 1 + k]);
    }
    uint32_t blitW = 0, blitH = 0, blitFormat = 0, blitFilter = 0;
    for (int i = 3; i + 4 < argc; i++) if (!strcmp(argv[i], "--blit")) { blitW = (uint32_t)atoi(argv[i + 1]); blitH = (uint32_t)atoi(argv[i + 2]); blitFormat = (uint32_t)atoi(argv[i + 3]); blitFilter = (uint32_t)atoi(argv[i + 4]); }
    LoadIcd();
    const SceneDesc sc = ParseScene(sceneDir);
    App app{};

    // init_instance / init_enumerate_device / init_device / init_command_pool / init_command_buffer
    VkApplicationInfo ai{VK_STRUCTURE_TYPE_APPLICATION_INFO, nullptr, "cpvk_harness", 1, "cpvk", 1, VK_API_VERSION_1_0};
    VkInstanceCreateInfo ici{VK_STRUCTURE_TYPE_INSTANCE_CREATE_INFO, nullptr, 0, &ai, 0, nullptr, 0, nullptr};
    VK(vkCreateInstance(&ici, nullptr, &app.instance));
    if (*reinterpret_cast<uintptr_t*>(app.instance) != ICD_LOADER_MAGIC) DIE("dispatchable handle lacks ICD_LOADER_MAGIC");
    uint32_t n = 1; VK(vkEnumeratePhysicalDevices(app.instance, &n, &app.gpu));
    VkPhysicalDeviceProperties props; vkGetPhysicalDeviceProperties(app.gpu, &props);
    VkPhysicalDeviceMemoryProperties mp; vkGetPhysicalDeviceMemoryProperties(app.gpu, &mp);
    if (mp.memoryTypeCount != 1 || (mp.memoryTypes[0].propertyFlags & 7) != 7) DIE("expected one DEVICE_LOCAL|HOST_VISIBLE|HOST_COHERENT memory type");
    float prio = 0; VkDeviceQueueCreateInfo qi{VK_STRUCTURE_TYPE_DEVICE_QUEUE_CREATE_INFO, nullptr, 0, 0, 1, &prio};
    VkDeviceCreateInfo di{VK_STRUCTURE_TYPE_DEVICE_CREATE_INFO, nullptr, 0, 1, &qi, 0, nullptr, 0, nullptr, nullptr};
    VK(vkCreateDevice(app.gpu, &di, nullptr, &app.device));
    vkGetDeviceQueue(app.device, 0, 0, &app.queue);
    VkCommandPoolCreateInfo cpi{VK_STRUCTURE_TYPE_COMMAND_POOL_CREATE_INFO, nullptr, 0, 0};
    VK(vkCreateCommandPool(app.device, &cpi, nullptr, &app.pool));
    VkCommandBufferAllocateInfo cbi{VK_STRUCTURE_TYPE_COMMAND_BUFFER_ALLOCATE_INFO, nullptr, app.pool, VK_COMMAND_BUFFER_LEVEL_PRIMARY, 1};
    VK(vkAllocateCommandBuffers(app.device, &cbi, &app.cmd));

    // init_vertex_buffer / init_uniform_buffer: data goes in through a mapping, no flush (host coherent)
    std::map<std::string, VkBuffer> bufs; std::map<std::string, VkDeviceMemory> bufMem; std::map<std::string, std::vector<uint8_t>> bufData;
    for (auto& kv : sc.buffers) {
        std::vector<uint8_t> data = ReadFile(sceneDir + "/" + kv.second.file);
        VkDeviceMemory mem;
        VkBuffer b = app.MakeBuffer(data.size(), VK_BUFFER_USAGE_VERTEX_BUFFER_BIT | VK_BUFFER_USAGE_INDEX_BUFFER_BIT | VK_BUFFER_USAGE_UNIFORM_BUFFER_BIT | VK_BUFFER_USAGE_UNIFORM_TEXEL_BUFFER_BIT | VK_BUFFER_USAGE_TRANSFER_DST_BIT, &mem);
        bool isUniform = false; for (auto& u : sc.uniforms) if (u.name == kv.first) isUniform = true;
        if (!(updateBuffers && isUniform)) { void* p; VK(vkMapMemory(app.device, mem, 0, data.size(), 0, &p)); memcpy(p, data.data(), data.size()); vkUnmapMemory(app.device, mem); }
        bufs[kv.first] = b; bufMem[kv.first] = mem; bufData[kv.first] = std::move(data);
    }
    // init_texture: linear image written through vkGetImageSubresourceLayout + map (util_init.cpp:1783-1799)
    struct TexObj { VkImage img; VkImageView view; VkSampler sampler; };
    std::vector<TexObj> texObjs;
    for (auto& t : sc.textures) {
        VkDeviceMemory mem; TexObj o{};
        o.img = app.MakeImage(t.format, t.w, t.h, VK_IMAGE_USAGE_SAMPLED_BIT, VK_IMAGE_TILING_LINEAR, &mem);
        VkImageSubresource sub{VK_IMAGE_ASPECT_COLOR_BIT, 0, 0}; VkSubresourceLayout lay; vkGetImageSubresourceLayout(app.device, o.img, &sub, &lay);
        std::vector<uint8_t> data = ReadFile(sceneDir + "/" + t.file);
        uint8_t* p; VK(vkMapMemory(app.device, mem, 0, lay.size, 0, (void**)&p));
        const uint32_t rowBytes = TexelSize(t.format) * t.w;
        for (uint32_t y = 0; y < t.h; y++) memcpy(p + lay.offset + y * lay.rowPitch, data.data() + (size_t)y * rowBytes, rowBytes);
        vkUnmapMemory(app.device, mem);
        o.view = app.MakeView(o.img, t.format, VK_IMAGE_ASPECT_COLOR_BIT);
        VkSamplerCreateInfo si{VK_STRUCTURE_TYPE_SAMPLER_CREATE_INFO, nullptr, 0, (VkFilter)t.filter, (VkFilter)t.filter, VK_SAMPLER_MIPMAP_MODE_NEAREST,
                               (VkSamplerAddressMode)t.address, (VkSamplerAddressMode)t.address, (VkSamplerAddressMode)t.address, 0.0f, VK_FALSE, 1.0f, VK_FALSE, VK_COMPARE_OP_NEVER,
                               0.0f, 0.0f, VK_BORDER_COLOR_FLOAT_OPAQUE_WHITE, VK_FALSE};
        VK(vkCreateSampler(app.device, &si, nullptr, &o.sampler));
        texObjs.push_back(o);
    }
    // colour target (the "swapchain image") + init_depth_buffer
    VkDeviceMemory colorMem, depthMem = nullptr;
    VkImage colorImg = app.MakeImage(sc.colorFormat, sc.width, sc.height, VK_IMAGE_USAGE_COLOR_ATTACHMENT_BIT | VK_IMAGE_USAGE_TRANSFER_SRC_BIT, VK_IMAGE_TILING_OPTIMAL, &colorMem);
    VkImageView colorView = app.MakeView(colorImg, sc.colorFormat, VK_IMAGE_ASPECT_COLOR_BIT);
    VkImage depthImg = nullptr; VkImageView depthView = nullptr;
    if (sc.depthFormat) {
        depthImg = app.MakeImage(sc.depthFormat, sc.width, sc.height, VK_IMAGE_USAGE_DEPTH_STENCIL_ATTACHMENT_BIT | VK_IMAGE_USAGE_TRANSFER_SRC_BIT, VK_IMAGE_TILING_OPTIMAL, &depthMem);
        depthView = app.MakeView(depthImg, sc.depthFormat, VK_IMAGE_ASPECT_DEPTH_BIT);
    }
    // init_descriptor_and_pipeline_layouts
    // one set layout per descriptor set number the scene uses (Samples/multiple_sets: the uniform buffer in set 0, the texture in set 1)
    uint32_t nSets = 1;
    for (auto& u : sc.uniforms) nSets = std::max(nSets, u.set + 1);
    for (auto& t : sc.textures) nSets = std::max(nSets, t.set + 1);
    for (auto& t : sc.texelBuffers) nSets = std::max(nSets, t.set + 1);
    std::vector<std::vector<VkDescriptorSetLayoutBinding>> lbs(nSets);
    for (auto& u : sc.uniforms) lbs[u.set].push_back({u.binding, u.dynRange ? VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER_DYNAMIC : VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER, 1, VK_SHADER_STAGE_VERTEX_BIT | VK_SHADER_STAGE_FRAGMENT_BIT, nullptr});
    for (size_t i = 0; i < sc.textures.size(); i++) {
        auto& t = sc.textures[i];
        auto& lb = lbs[t.set];
        const VkSampler* immutable = t.immutableSampler ? &texObjs[i].sampler : nullptr;
        if (t.inputAttachment) lb.push_back({t.binding, VK_DESCRIPTOR_TYPE_INPUT_ATTACHMENT, 1, VK_SHADER_STAGE_FRAGMENT_BIT, nullptr}); // input_attachment.cpp:194-201
        else if (t.samplerBinding < 0) lb.push_back({t.binding, VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER, 1, VK_SHADER_STAGE_FRAGMENT_BIT, immutable});
        else {
            lb.push_back({t.binding, VK_DESCRIPTOR_TYPE_SAMPLED_IMAGE, 1, VK_SHADER_STAGE_FRAGMENT_BIT, nullptr});
            lb.push_back({(uint32_t)t.samplerBinding, VK_DESCRIPTOR_TYPE_SAMPLER, 1, VK_SHADER_STAGE_FRAGMENT_BIT, immutable});
        }
    }
    for (auto& t : sc.texelBuffers) lbs[t.set].push_back({t.binding, VK_DESCRIPTOR_TYPE_UNIFORM_TEXEL_BUFFER, 1, VK_SHADER_STAGE_VERTEX_BIT | VK_SHADER_STAGE_FRAGMENT_BIT, nullptr});
    std::vector<VkDescriptorSetLayout> setLayouts(nSets);
    for (uint32_t k = 0; k < nSets; k++) {
        VkDescriptorSetLayoutCreateInfo li{VK_STRUCTURE_TYPE_DESCRIPTOR_SET_LAYOUT_CREATE_INFO, nullptr, 0, (uint32_t)lbs[k].size(), lbs[k].data()};
        VK(vkCreateDescriptorSetLayout(app.device, &li, nullptr, &setLayouts[k]));
    }
    VkPushConstantRange pcr{VK_SHADER_STAGE_VERTEX_BIT | VK_SHADER_STAGE_FRAGMENT_BIT, 0, (uint32_t)sc.pushConstants.size()};
    VkPipelineLayoutCreateInfo pli{VK_STRUCTURE_TYPE_PIPELINE_LAYOUT_CREATE_INFO, nullptr, 0, nSets, setLayouts.data(), sc.pushConstants.empty() ? 0u : 1u, sc.pushConstants.empty() ? nullptr : &pcr};
    VkPipelineLayout pipeLayout; VK(vkCreatePipelineLayout(app.device, &pli, nullptr, &pipeLayout));
    // init_renderpass (loadOp CLEAR / storeOp STORE) + init_framebuffers
    std::vector<VkAttachmentDescription> atts;
    atts.push_back({0, (VkFormat)sc.colorFormat, VK_SAMPLE_COUNT_1_BIT, VK_ATTACHMENT_LOAD_OP_CLEAR, VK_ATTACHMENT_STORE_OP_STORE, VK_ATTACHMENT_LOAD_OP_DONT_CARE, VK_ATTACHMENT_STORE_OP_DONT_CARE,
                    VK_IMAGE_LAYOUT_UNDEFINED, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL});
    if (sc.depthFormat) atts.push_back({0, (VkFormat)sc.depthFormat, VK_SAMPLE_COUNT_1_BIT, VK_ATTACHMENT_LOAD_OP_CLEAR, VK_ATTACHMENT_STORE_OP_STORE, VK_ATTACHMENT_LOAD_OP_CLEAR, VK_ATTACHMENT_STORE_OP_STORE,
                                        VK_IMAGE_LAYOUT_UNDEFINED, VK_IMAGE_LAYOUT_DEPTH_STENCIL_ATTACHMENT_OPTIMAL});
    VkAttachmentReference cref{0, VK_IMAGE_LAYOUT_COLOR_ATTACHMENT_OPTIMAL}, dref{1, VK_IMAGE_LAYOUT_DEPTH_STENCIL_ATTACHMENT_OPTIMAL};
    VkSubpassDescription sp{0, VK_PIPELINE_BIND_POINT_GRAPHICS, 0, nullptr, 1, &cref, nullptr, sc.depthFormat ? &dref : nullptr, 0, nullptr};
    VkRenderPassCreateInfo rpi{VK_STRUCTURE_TYPE_RENDER_PASS_CREATE_INFO, nullptr, 0, (uint32_t)atts.size(), atts.data(), 1, &sp, 0, nullptr};
    VkRenderPass renderPass; VK(vkCreateRenderPass(app.device, &rpi, nullptr, &renderPass));
    VkImageView fbViews[2] = {colorView, depthView};
    VkFramebufferCreateInfo fbi{VK_STRUCTURE_TYPE_FRAMEBUFFER_CREATE_INFO, nullptr, 0, renderPass, (uint32_t)atts.size(), fbViews, sc.width, sc.height, 1};
    VkFramebuffer framebuffer; VK(vkCreateFramebuffer(app.device, &fbi, nullptr, &framebuffer));
    // init_descriptor_pool / init_descriptor_set
    VkDescriptorPoolSize ps[6] = {{VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER, 8}, {VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER, 8}, {VK_DESCRIPTOR_TYPE_UNIFORM_TEXEL_BUFFER, 8},
                                  {VK_DESCRIPTOR_TYPE_SAMPLED_IMAGE, 8}, {VK_DESCRIPTOR_TYPE_SAMPLER, 8}, {VK_DESCRIPTOR_TYPE_INPUT_ATTACHMENT, 8}};
    VkDescriptorPoolCreateInfo dpi{VK_STRUCTURE_TYPE_DESCRIPTOR_POOL_CREATE_INFO, nullptr, 0, nSets, 6, ps};
    VkDescriptorPool dpool; VK(vkCreateDescriptorPool(app.device, &dpi, nullptr, &dpool));
    VkDescriptorSetAllocateInfo dsa{VK_STRUCTURE_TYPE_DESCRIPTOR_SET_ALLOCATE_INFO, nullptr, dpool, nSets, setLayouts.data()};
    std::vector<VkDescriptorSet> dsets(nSets); VK(vkAllocateDescriptorSets(app.device, &dsa, dsets.data()));
    std::vector<VkDescriptorBufferInfo> binfos(sc.uniforms.size()); std::vector<VkDescriptorImageInfo> iinfos(sc.textures.size() * 2); std::vector<VkWriteDescriptorSet> writes;
    for (size_t i = 0; i < sc.uniforms.size(); i++) {
        const bool dyn = sc.uniforms[i].dynRange != 0; // dynamic_uniform.cpp:196-215: the descriptor covers one element, the offset picks it at bind time
        binfos[i] = {bufs.at(sc.uniforms[i].name), 0, dyn ? (VkDeviceSize)sc.uniforms[i].dynRange : sc.buffers.at(sc.uniforms[i].name).size};
        writes.push_back({VK_STRUCTURE_TYPE_WRITE_DESCRIPTOR_SET, nullptr, dsets[sc.uniforms[i].set], sc.uniforms[i].binding, 0, 1, dyn ? VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER_DYNAMIC : VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER, nullptr, &binfos[i], nullptr});
    }
    for (size_t i = 0; i < sc.textures.size(); i++) {
        auto& t = sc.textures[i];
        const VkSampler written = t.immutableSampler ? VkSampler(VK_NULL_HANDLE) : texObjs[i].sampler; // immutable: image_info.sampler = 0 (immutable_sampler.cpp:106)
        if (t.inputAttachment) {
            iinfos[2 * i] = {VK_NULL_HANDLE, texObjs[i].view, VK_IMAGE_LAYOUT_SHADER_READ_ONLY_OPTIMAL};
            writes.push_back({VK_STRUCTURE_TYPE_WRITE_DESCRIPTOR_SET, nullptr, dsets[t.set], t.binding, 0, 1, VK_DESCRIPTOR_TYPE_INPUT_ATTACHMENT, &iinfos[2 * i], nullptr, nullptr});
        } else if (t.samplerBinding < 0) {
            iinfos[2 * i] = {written, texObjs[i].view, VK_IMAGE_LAYOUT_SHADER_READ_ONLY_OPTIMAL};
            writes.push_back({VK_STRUCTURE_TYPE_WRITE_DESCRIPTOR_SET, nullptr, dsets[t.set], t.binding, 0, 1, VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER, &iinfos[2 * i], nullptr, nullptr});
        } else { // separate_image_sampler.cpp:114-125: image_info.sampler = 0 for the texture, a second info for the sampler
            iinfos[2 * i] = {VK_NULL_HANDLE, texObjs[i].view, VK_IMAGE_LAYOUT_SHADER_READ_ONLY_OPTIMAL};
            writes.push_back({VK_STRUCTURE_TYPE_WRITE_DESCRIPTOR_SET, nullptr, dsets[t.set], t.binding, 0, 1, VK_DESCRIPTOR_TYPE_SAMPLED_IMAGE, &iinfos[2 * i], nullptr, nullptr});
            if (!t.immutableSampler) {
                iinfos[2 * i + 1] = {texObjs[i].sampler, VK_NULL_HANDLE, VK_IMAGE_LAYOUT_UNDEFINED};
                writes.push_back({VK_STRUCTURE_TYPE_WRITE_DESCRIPTOR_SET, nullptr, dsets[t.set], (uint32_t)t.samplerBinding, 0, 1, VK_DESCRIPTOR_TYPE_SAMPLER, &iinfos[2 * i + 1], nullptr, nullptr});
            }
        }
    }
    // texel_buffer.cpp:146-158, :226-236: a buffer view over the whole buffer bound as UNIFORM_TEXEL_BUFFER
    std::vector<VkBufferView> bufferViews(sc.texelBuffers.size());
    for (size_t i = 0; i < sc.texelBuffers.size(); i++) {
        VkBufferViewCreateInfo bvi{VK_STRUCTURE_TYPE_BUFFER_VIEW_CREATE_INFO, nullptr, 0, bufs.at(sc.texelBuffers[i].name), (VkFormat)sc.texelBuffers[i].format, 0, sc.buffers.at(sc.texelBuffers[i].name).size};
        VK(vkCreateBufferView(app.device, &bvi, nullptr, &bufferViews[i]));
        writes.push_back({VK_STRUCTURE_TYPE_WRITE_DESCRIPTOR_SET, nullptr, dsets[sc.texelBuffers[i].set], sc.texelBuffers[i].binding, 0, 1, VK_DESCRIPTOR_TYPE_UNIFORM_TEXEL_BUFFER, nullptr, nullptr, &bufferViews[i]});
    }
    vkUpdateDescriptorSets(app.device, (uint32_t)writes.size(), writes.data(), 0, nullptr);
    // init_shaders (SPIR-V words exported by Python; the samples run glslang here) + init_pipeline
    std::vector<uint8_t> vsb = ReadFile(sceneDir + "/" + sc.vs), fsb = ReadFile(sceneDir + "/" + sc.fs);
    VkShaderModuleCreateInfo smi{VK_STRUCTURE_TYPE_SHADER_MODULE_CREATE_INFO, nullptr, 0, vsb.size(), (const uint32_t*)vsb.data()};
    VkShaderModule vsm, fsm; VK(vkCreateShaderModule(app.device, &smi, nullptr, &vsm));
    smi.codeSize = fsb.size(); smi.pCode = (const uint32_t*)fsb.data(); VK(vkCreateShaderModule(app.device, &smi, nullptr, &fsm));
    // spirv_specialization.cpp: one 32-bit value per constant id, packed back to back
    VkSpecializationInfo specInfo[2] = {}; std::vector<VkSpecializationMapEntry> specEntries[2]; std::vector<uint32_t> specData[2];
    for (auto& sp : sc.specs) { const uint32_t st = sp.stage ? 1u : 0u; specEntries[st].push_back({sp.id, (uint32_t)(specData[st].size() * 4), 4}); specData[st].push_back(sp.value); }
    for (int st = 0; st < 2; st++) specInfo[st] = {(uint32_t)specEntries[st].size(), specEntries[st].data(), specData[st].size() * 4, specData[st].data()};
    VkPipelineShaderStageCreateInfo stages[2] = {{VK_STRUCTURE_TYPE_PIPELINE_SHADER_STAGE_CREATE_INFO, nullptr, 0, VK_SHADER_STAGE_VERTEX_BIT, vsm, "main", specEntries[0].empty() ? nullptr : &specInfo[0]},
                                                 {VK_STRUCTURE_TYPE_PIPELINE_SHADER_STAGE_CREATE_INFO, nullptr, 0, VK_SHADER_STAGE_FRAGMENT_BIT, fsm, "main", specEntries[1].empty() ? nullptr : &specInfo[1]}};
    VkPipelineVertexInputStateCreateInfo vis{VK_STRUCTURE_TYPE_PIPELINE_VERTEX_INPUT_STATE_CREATE_INFO, nullptr, 0, (uint32_t)sc.bindings.size(), sc.bindings.data(), (uint32_t)sc.attributes.size(), sc.attributes.data()};
    VkPipelineInputAssemblyStateCreateInfo ias{VK_STRUCTURE_TYPE_PIPELINE_INPUT_ASSEMBLY_STATE_CREATE_INFO, nullptr, 0, (VkPrimitiveTopology)sc.topology, VK_FALSE};
    VkPipelineViewportStateCreateInfo vps{VK_STRUCTURE_TYPE_PIPELINE_VIEWPORT_STATE_CREATE_INFO, nullptr, 0, 1, nullptr, 1, nullptr};
    VkPipelineRasterizationStateCreateInfo rss{VK_STRUCTURE_TYPE_PIPELINE_RASTERIZATION_STATE_CREATE_INFO, nullptr, 0, VK_FALSE, VK_FALSE, VK_POLYGON_MODE_FILL, sc.cull, (VkFrontFace)sc.front, VK_FALSE, 0, 0, 0, sc.lineWidth};
    VkPipelineMultisampleStateCreateInfo mss{VK_STRUCTURE_TYPE_PIPELINE_MULTISAMPLE_STATE_CREATE_INFO, nullptr, 0, VK_SAMPLE_COUNT_1_BIT, VK_FALSE, 0.0f, nullptr, VK_FALSE, VK_FALSE};
    VkStencilOpState sop{VK_STENCIL_OP_KEEP, VK_STENCIL_OP_KEEP, VK_STENCIL_OP_KEEP, VK_COMPARE_OP_ALWAYS, 0, 0, 0};
    VkPipelineDepthStencilStateCreateInfo dss{VK_STRUCTURE_TYPE_PIPELINE_DEPTH_STENCIL_STATE_CREATE_INFO, nullptr, 0, sc.depthTest, sc.depthWrite, (VkCompareOp)sc.depthOp, VK_FALSE, VK_FALSE, sop, sop, 0.0f, 1.0f};
    VkPipelineColorBlendAttachmentState cba{sc.blend ? VK_TRUE : VK_FALSE, (VkBlendFactor)sc.bl[0], (VkBlendFactor)sc.bl[1], (VkBlendOp)sc.bl[2], (VkBlendFactor)sc.bl[3], (VkBlendFactor)sc.bl[4], (VkBlendOp)sc.bl[5], sc.writeMask};
    VkPipelineColorBlendStateCreateInfo cbs{VK_STRUCTURE_TYPE_PIPELINE_COLOR_BLEND_STATE_CREATE_INFO, nullptr, 0, VK_FALSE, VK_LOGIC_OP_COPY, 1, &cba, {1, 1, 1, 1}};
    VkDynamicState dyn[2] = {VK_DYNAMIC_STATE_VIEWPORT, VK_DYNAMIC_STATE_SCISSOR};
    VkPipelineDynamicStateCreateInfo dys{VK_STRUCTURE_TYPE_PIPELINE_DYNAMIC_STATE_CREATE_INFO, nullptr, 0, 2, dyn};
    VkGraphicsPipelineCreateInfo gpi{VK_STRUCTURE_TYPE_GRAPHICS_PIPELINE_CREATE_INFO, nullptr, 0, 2, stages, &vis, &ias, nullptr, &vps, &rss, &mss, &dss, &cbs, &dys, pipeLayout, renderPass, 0, nullptr, 0};
    VkPipeline pipeline;
    const auto tp0 = std::chrono::steady_clock::now();
    VK(vkCreateGraphicsPipelines(app.device, nullptr, 1, &gpi, nullptr, &pipeline));
    const double pipelineMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tp0).count();

    // readback buffers (the samples' write_ppm copies the image out and maps it, util.cpp:595-622)
    const uint64_t colorBytes = (uint64_t)TexelSize(sc.colorFormat) * sc.width * sc.height;
    const uint64_t depthBytes = sc.depthFormat ? (uint64_t)TexelSize(sc.depthFormat) * sc.width * sc.height : 0;
    VkDeviceMemory rbMem, rbDepthMem = nullptr;
    VkBuffer rb = app.MakeBuffer(colorBytes, VK_BUFFER_USAGE_TRANSFER_DST_BIT, &rbMem), rbDepth = nullptr;
    if (depthBytes) rbDepth = app.MakeBuffer(depthBytes, VK_BUFFER_USAGE_TRANSFER_DST_BIT, &rbDepthMem);
    uint8_t* rbPtr; VK(vkMapMemory(app.device, rbMem, 0, colorBytes, 0, (void**)&rbPtr)); // persistently mapped: coherent reads after the fence

    // record: the 15-draw_cube command sequence (15-draw_cube.cpp:100-160)
    VkCommandBufferBeginInfo bi{VK_STRUCTURE_TYPE_COMMAND_BUFFER_BEGIN_INFO, nullptr, 0, nullptr};
    VK(vkBeginCommandBuffer(app.cmd, &bi));
    if (updateBuffers) // vkCmdFillBuffer + vkCmdUpdateBuffer outside the render pass (CommandBuffer.cpp:241-330)
        for (auto& u : sc.uniforms) {
            const auto& data = bufData.at(u.name);
            vkCmdFillBuffer(app.cmd, bufs.at(u.name), 0, VK_WHOLE_SIZE, 0xDEADBEEFu);
            vkCmdUpdateBuffer(app.cmd, bufs.at(u.name), 0, data.size(), data.data());
        }
    // indirect parameters: slot 1 of three (offset = stride = 32), the other slots hold garbage that must not be read
    VkBuffer indirectBuf = VK_NULL_HANDLE; VkDeviceMemory indirectMem = VK_NULL_HANDLE;
    if (indirect) {
        indirectBuf = app.MakeBuffer(96, VK_BUFFER_USAGE_INDIRECT_BUFFER_BIT, &indirectMem);
        uint32_t* w; VK(vkMapMemory(app.device, indirectMem, 0, 96, 0, (void**)&w));
        for (int i = 0; i < 24; i++) w[i] = 0x7FFFFFFFu;
        w[0] = 1;
        if (sc.indexStride) { w[8] = sc.count; w[9] = sc.instances; w[10] = sc.first; w[11] = (uint32_t)sc.vertexOffset; w[12] = sc.firstInstance; }
        else { w[8] = sc.count; w[9] = sc.instances; w[10] = sc.first; w[11] = sc.firstInstance; }
        vkUnmapMemory(app.device, indirectMem);
    }
    VkClearValue clears[2]; memcpy(clears[0].color.float32, sc.clearColor, 16); clears[1].depthStencil = {sc.clearDepth, sc.clearStencil};
    VkRenderPassBeginInfo rbi{VK_STRUCTURE_TYPE_RENDER_PASS_BEGIN_INFO, nullptr, renderPass, framebuffer, {{0, 0}, {sc.width, sc.height}}, (uint32_t)atts.size(), clears};
    vkCmdBeginRenderPass(app.cmd, &rbi, secondary ? VK_SUBPASS_CONTENTS_SECONDARY_COMMAND_BUFFERS : VK_SUBPASS_CONTENTS_INLINE);
    VkCommandBuffer rec = app.cmd;
    if (secondary) { // vkCmdExecuteCommands (CommandBuffer.cpp:704-731): the draw lives in a secondary command buffer
        VkCommandBufferAllocateInfo sbi{VK_STRUCTURE_TYPE_COMMAND_BUFFER_ALLOCATE_INFO, nullptr, app.pool, VK_COMMAND_BUFFER_LEVEL_SECONDARY, 1};
        VK(vkAllocateCommandBuffers(app.device, &sbi, &rec));
        VkCommandBufferInheritanceInfo inh{VK_STRUCTURE_TYPE_COMMAND_BUFFER_INHERITANCE_INFO, nullptr, renderPass, 0, framebuffer, VK_FALSE, 0, 0};
        VkCommandBufferBeginInfo sbeg{VK_STRUCTURE_TYPE_COMMAND_BUFFER_BEGIN_INFO, nullptr, 0x2 /* VK_COMMAND_BUFFER_USAGE_RENDER_PASS_CONTINUE_BIT */, &inh};
        VK(vkBeginCommandBuffer(rec, &sbeg));
    }
    vkCmdBindPipeline(rec, VK_PIPELINE_BIND_POINT_GRAPHICS, pipeline);
    // one vkCmdBindDescriptorSets per set (firstSet = the set number, Binding.cpp:58-80), each with its own dynamic offsets in
    // binding order, as the API defines for the dynamic descriptors of a set
    for (uint32_t k = 0; k < nSets; k++) {
        std::vector<std::pair<uint32_t, uint32_t>> byBinding;
        for (auto& u : sc.uniforms) if (u.dynRange && u.set == k) byBinding.push_back({u.binding, u.dynOffset});
        std::sort(byBinding.begin(), byBinding.end());
        std::vector<uint32_t> dynOffsets; for (auto& kv : byBinding) dynOffsets.push_back(kv.second);
        vkCmdBindDescriptorSets(rec, VK_PIPELINE_BIND_POINT_GRAPHICS, pipeLayout, k, 1, &dsets[k], (uint32_t)dynOffsets.size(), dynOffsets.empty() ? nullptr : dynOffsets.data());
    }
    if (!sc.pushConstants.empty()) vkCmdPushConstants(rec, pipeLayout, VK_SHADER_STAGE_VERTEX_BIT | VK_SHADER_STAGE_FRAGMENT_BIT, 0, (uint32_t)sc.pushConstants.size(), sc.pushConstants.data());
    for (auto& kv : sc.vertexBuffers) { VkDeviceSize off = 0; VkBuffer b = bufs.at(kv.second); vkCmdBindVertexBuffers(rec, kv.first, 1, &b, &off); }
    VkViewport vp{sc.viewport[0], sc.viewport[1], sc.viewport[2], sc.viewport[3], sc.viewport[4], sc.viewport[5]};
    vkCmdSetViewport(rec, 0, 1, &vp);
    VkRect2D scissor{{0, 0}, {sc.width, sc.height}}; vkCmdSetScissor(rec, 0, 1, &scissor);
    if (sc.indexStride) {
        vkCmdBindIndexBuffer(rec, bufs.at(sc.indexBuffer), 0, sc.indexStride == 2 ? VK_INDEX_TYPE_UINT16 : sc.indexStride == 4 ? VK_INDEX_TYPE_UINT32 : VK_INDEX_TYPE_UINT8_EXT);
        if (indirectCount) vkCmdDrawIndexedIndirectCount(rec, indirectBuf, 32, indirectBuf, 0, 3, 32);
        else if (indirect) vkCmdDrawIndexedIndirect(rec, indirectBuf, 32, 1, 32);
        else vkCmdDrawIndexed(rec, sc.count, sc.instances, sc.first, sc.vertexOffset, sc.firstInstance);
    } else if (indirectCount) vkCmdDrawIndirectCount(rec, indirectBuf, 32, indirectBuf, 0, 3, 32);
    else if (indirect) vkCmdDrawIndirect(rec, indirectBuf, 32, 1, 32);
    else vkCmdDraw(rec, sc.count, sc.instances, sc.first, sc.firstInstance);
    if (clearRect[2] > 0) { // vkCmdClearAttachments (Draw.cpp:2226-2348) after the draw, inside the pass
        VkClearAttachment ca[2]; uint32_t nca = 1;
        ca[0].aspectMask = VK_IMAGE_ASPECT_COLOR_BIT; ca[0].colorAttachment = 0;
        ca[0].clearValue.color.float32[0] = 0.5f; ca[0].clearValue.color.float32[1] = 0.25f; ca[0].clearValue.color.float32[2] = 0.75f; ca[0].clearValue.color.float32[3] = 1.0f;
        if (sc.depthFormat) { ca[1].aspectMask = VK_IMAGE_ASPECT_DEPTH_BIT | VK_IMAGE_ASPECT_STENCIL_BIT; ca[1].colorAttachment = 0; ca[1].clearValue.depthStencil = {0.5f, 0}; nca = 2; }
        VkClearRect cr{{{clearRect[0], clearRect[1]}, {(uint32_t)clearRect[2], (uint32_t)clearRect[3]}}, 0, 1};
        vkCmdClearAttachments(rec, nca, ca, 1, &cr);
    }
    if (secondary) { VK(vkEndCommandBuffer(rec)); vkCmdExecuteCommands(app.cmd, 1, &rec); }
    vkCmdEndRenderPass(app.cmd);
    VkBufferImageCopy cp{0, 0, 0, {VK_IMAGE_ASPECT_COLOR_BIT, 0, 0, 1}, {0, 0, 0}, {sc.width, sc.height, 1}};
    vkCmdCopyImageToBuffer(app.cmd, colorImg, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL, rb, 1, &cp);
    if (depthBytes) { cp.imageSubresource.aspectMask = VK_IMAGE_ASPECT_DEPTH_BIT; vkCmdCopyImageToBuffer(app.cmd, depthImg, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL, rbDepth, 1, &cp); }
    VkBuffer rbBlit = VK_NULL_HANDLE; VkDeviceMemory rbBlitMem = VK_NULL_HANDLE; uint64_t blitBytes = 0;
    if (blitW) {
        VkDeviceMemory m1, m2;
        VkImage blitImg = app.MakeImage(blitFormat, blitW, blitH, VK_IMAGE_USAGE_TRANSFER_DST_BIT | VK_IMAGE_USAGE_TRANSFER_SRC_BIT, VK_IMAGE_TILING_OPTIMAL, &m1);
        VkImage copyImg = app.MakeImage(blitFormat, blitW, blitH, VK_IMAGE_USAGE_TRANSFER_DST_BIT | VK_IMAGE_USAGE_TRANSFER_SRC_BIT, VK_IMAGE_TILING_LINEAR, &m2);
        VkImageBlit region{{VK_IMAGE_ASPECT_COLOR_BIT, 0, 0, 1}, {{0, 0, 0}, {(int32_t)sc.width, (int32_t)sc.height, 1}}, {VK_IMAGE_ASPECT_COLOR_BIT, 0, 0, 1}, {{0, 0, 0}, {(int32_t)blitW, (int32_t)blitH, 1}}};
        vkCmdBlitImage(app.cmd, colorImg, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL, blitImg, VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL, 1, &region, (VkFilter)blitFilter);
        VkImageCopy copy{{VK_IMAGE_ASPECT_COLOR_BIT, 0, 0, 1}, {0, 0, 0}, {VK_IMAGE_ASPECT_COLOR_BIT, 0, 0, 1}, {0, 0, 0}, {blitW, blitH, 1}};
        vkCmdCopyImage(app.cmd, blitImg, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL, copyImg, VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL, 1, &copy);
        blitBytes = (uint64_t)TexelSize(blitFormat) * blitW * blitH;
        rbBlit = app.MakeBuffer(blitBytes, VK_BUFFER_USAGE_TRANSFER_DST_BIT, &rbBlitMem);
        VkBufferImageCopy bc{0, 0, 0, {VK_IMAGE_ASPECT_COLOR_BIT, 0, 0, 1}, {0, 0, 0}, {blitW, blitH, 1}};
        vkCmdCopyImageToBuffer(app.cmd, copyImg, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL, rbBlit, 1, &bc);
    }
    VK(vkEndCommandBuffer(app.cmd));

    VkFenceCreateInfo fi{VK_STRUCTURE_TYPE_FENCE_CREATE_INFO, nullptr, 0};
    VkFence fence; VK(vkCreateFence(app.device, &fi, nullptr, &fence));
    VkSubmitInfo si{VK_STRUCTURE_TYPE_SUBMIT_INFO, nullptr, 0, nullptr, nullptr, 1, &app.cmd, 0, nullptr};
    std::vector<double> frameMs, submitMs; // whole frame (incl. the application's vertex rewrite) / vkQueueSubmit .. fence only
    uint64_t checksum = 0;
    for (int f = 0; f < frames; f++) {
        const auto t0 = std::chrono::steady_clock::now();
        if (frames > 1) { // a real frame loop: the application rewrites its vertex data through the mapping every frame
            for (auto& kv : sc.vertexBuffers) {
                void* p; const auto& data = bufData.at(kv.second);
                VK(vkMapMemory(app.device, bufMem.at(kv.second), 0, data.size(), 0, &p)); memcpy(p, data.data(), data.size()); vkUnmapMemory(app.device, bufMem.at(kv.second));
            }
        }
        VK(vkResetFences(app.device, 1, &fence));
        const auto ts = std::chrono::steady_clock::now();
        VK(vkQueueSubmit(app.queue, 1, &si, fence));
        VK(vkWaitForFences(app.device, 1, &fence, VK_TRUE, ~0ull));
        submitMs.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ts).count());
        checksum += rbPtr[0] + rbPtr[colorBytes - 1]; // the result is host-visible right after the fence
        frameMs.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    { std::ofstream o(outDir + "/color.bin", std::ios::binary); o.write((const char*)rbPtr, (std::streamsize)colorBytes); }
    if (depthBytes) {
        uint8_t* p; VK(vkMapMemory(app.device, rbDepthMem, 0, depthBytes, 0, (void**)&p));
        std::ofstream o(outDir + "/depth.bin", std::ios::binary); o.write((const char*)p, (std::streamsize)depthBytes);
    }
    if (blitBytes) {
        uint8_t* p; VK(vkMapMemory(app.device, rbBlitMem, 0, blitBytes, 0, (void**)&p));
        std::ofstream o(outDir + "/blit.bin", std::ios::binary); o.write((const char*)p, (std::streamsize)blitBytes);
    }
    double sum = 0, best = 1e30; const size_t skip = frameMs.size() > 3 ? 3 : 0;
    for (size_t i = skip; i < frameMs.size(); i++) { sum += frameMs[i]; if (frameMs[i] < best) best = frameMs[i]; }
    std::vector<double> sorted(submitMs.begin() + (std::ptrdiff_t)skip, submitMs.end());
    std::sort(sorted.begin(), sorted.end());
    const double submitMedian = sorted.empty() ? 0.0 : sorted[sorted.size() / 2]; // wall clock vkQueueSubmit -> fence signalled
    printf("{\"device\": \"%s\", \"frames\": %d, \"ms_per_frame\": %.6f, \"ms_best\": %.6f, \"ms_submit_to_fence\": %.6f, \"pipeline_create_ms\": %.3f, \"checksum\": %llu}\n", props.deviceName, frames,
           sum / (double)(frameMs.size() - skip), best, submitMedian, pipelineMs, (unsigned long long)checksum);
    VK(vkDeviceWaitIdle(app.device));
    vkDestroyDevice(app.device, nullptr);
    vkDestroyInstance(app.instance, nullptr);
    return 0;
}
