// ref_sampler_check.cpp — runs the REFERENCE's own texture sampler, CPVulkan/ImageSampler.cpp (wrap, NEAREST / LINEAR
// coordinate arithmetic, the double-precision lerp chain, mip selection, border colours, missing-channel defaults:
// SURVEY §8(a) a13), compiled IN PLACE from /root/reference by oracle/Makefile into oracle/_ref/sampler_check.
// What is absent from this image is replaced by oracle/shim/: Vulkan and GSL headers (public API names), glm >= 0.9.9
// (per-component operators), and LLVMRuntime/Compilers.h — the JIT that would emit the per-format texel load. The texel
// function supplied here is a raw 16-byte copy, which is what ImageCompiler.cpp emits for R32G32B32A32_SFLOAT, so the
// sampler logic runs on exact texel values with no codec in between.
// TEST INFRASTRUCTURE ONLY: tests/golden/make_ref_golden.py stores its output, tests/test_reference_sampler.py compares
// the oracle's sampler with it.
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <unordered_map>
#include <vector>

#include <ImageSampler.cpp> // /root/reference/CPVulkan — included as a translation unit so its file-local templates are callable

ImageFunctions::ImageFunctions(CPJit* j) : jit(j) {}
ImageFunctions::~ImageFunctions() = default;

static void GetRGBA32F(const void* ptr, void* values) { std::memcpy(values, ptr, 16); }
static void SetRGBA32F(void* ptr, const float* values) { std::memcpy(ptr, values, 16); }
static FunctionPointer Unsupported() { std::fprintf(stderr, "sampler_check: only R32G32B32A32_SFLOAT texel functions exist\n"); std::abort(); }
FunctionPointer CompileGetPixelDepth(CPJit*, const FormatInformation*) { return Unsupported(); }
FunctionPointer CompileGetPixelStencil(CPJit*, const FormatInformation*) { return Unsupported(); }
FunctionPointer CompileGetPixelF32(CPJit*, const FormatInformation* f) { return f->Format == VK_FORMAT_R32G32B32A32_SFLOAT ? reinterpret_cast<FunctionPointer>(GetRGBA32F) : Unsupported(); }
FunctionPointer CompileGetPixelI32(CPJit*, const FormatInformation*) { return Unsupported(); }
FunctionPointer CompileGetPixelU32(CPJit*, const FormatInformation*) { return Unsupported(); }
FunctionPointer CompileSetPixelDepthStencil(CPJit*, const FormatInformation*) { return Unsupported(); }
FunctionPointer CompileSetPixelF32(CPJit*, const FormatInformation* f) { return f->Format == VK_FORMAT_R32G32B32A32_SFLOAT ? reinterpret_cast<FunctionPointer>(SetRGBA32F) : Unsupported(); }
FunctionPointer CompileSetPixelI32(CPJit*, const FormatInformation*) { return Unsupported(); }
FunctionPointer CompileSetPixelU32(CPJit*, const FormatInformation*) { return Unsupported(); }

struct Config { uint32_t mag, min, mipmap, addressU, addressV, border; float lod; };

// "3d" mode — the overload vkCmdBlitImage uses (CommandBuffer.cpp:75-226): SampleImage(state, format, data, uvec3 range, fvec3
// coordinates, filter) = lod 1, CLAMP_TO_EDGE, so the MINIFICATION filter on the single level, eight taps when LINEAR.
// input: u32 w, h, d, nCoords; w*h*d*4 floats; coords (u, v, w). Output: for filter NEAREST then LINEAR, nCoords lines.
static int Main3D(const char* path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return 2;
    uint32_t hdr[4];
    in.read(reinterpret_cast<char*>(hdr), 16);
    std::vector<float> texels((size_t)hdr[0] * hdr[1] * hdr[2] * 4);
    in.read(reinterpret_cast<char*>(texels.data()), texels.size() * 4);
    std::vector<float> coords((size_t)hdr[3] * 3);
    in.read(reinterpret_cast<char*>(coords.data()), coords.size() * 4);
    if (!in) return 2;
    auto state = std::make_unique<DeviceState>();
    state->jit = nullptr;
    gsl::span<uint8_t> data(reinterpret_cast<uint8_t*>(texels.data()), (std::ptrdiff_t)(texels.size() * 4));
    for (VkFilter filter : {VK_FILTER_NEAREST, VK_FILTER_LINEAR})
        for (uint32_t i = 0; i < hdr[3]; i++) {
            const glm::fvec4 r = SampleImage<glm::fvec4>(state.get(), VK_FORMAT_R32G32B32A32_SFLOAT, data, glm::uvec3(hdr[0], hdr[1], hdr[2]),
                                                         glm::fvec3(coords[3 * i], coords[3 * i + 1], coords[3 * i + 2]), filter);
            uint32_t b[4];
            std::memcpy(b, &r.x, 16);
            std::printf("%u %u %u %u\n", b[0], b[1], b[2], b[3]);
        }
    return 0;
}

// input file (little endian): u32 levels, nConfigs, nCoords; per level u32 w, h then w*h*4 floats; configs; coords (u, v)
// output on stdout: nConfigs * nCoords lines of four float bit patterns
int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: sampler_check input.bin | sampler_check 3d input.bin\n"); return 2; }
    if (argc > 2 && !std::strcmp(argv[1], "3d")) return Main3D(argv[2]);
    std::ifstream in(argv[1], std::ios::binary);
    if (!in) return 2;
    uint32_t hdr[3];
    in.read(reinterpret_cast<char*>(hdr), 12);
    const uint32_t levels = hdr[0], nConfigs = hdr[1], nCoords = hdr[2];
    if (levels == 0 || levels > MAX_MIP_LEVELS) return 2;
    std::vector<std::vector<float>> texels(levels);
    gsl::span<uint8_t> data[MAX_MIP_LEVELS];
    glm::uvec2 range[MAX_MIP_LEVELS];
    for (uint32_t l = 0; l < levels; l++) {
        uint32_t wh[2];
        in.read(reinterpret_cast<char*>(wh), 8);
        texels[l].resize((size_t)wh[0] * wh[1] * 4);
        in.read(reinterpret_cast<char*>(texels[l].data()), texels[l].size() * 4);
        data[l] = gsl::span<uint8_t>(reinterpret_cast<uint8_t*>(texels[l].data()), (std::ptrdiff_t)(texels[l].size() * 4));
        range[l] = glm::uvec2(wh[0], wh[1]);
    }
    std::vector<Config> configs(nConfigs);
    in.read(reinterpret_cast<char*>(configs.data()), nConfigs * sizeof(Config));
    std::vector<float> coords((size_t)nCoords * 2);
    in.read(reinterpret_cast<char*>(coords.data()), coords.size() * 4);
    if (!in) return 2;
    auto state = std::make_unique<DeviceState>();
    state->jit = nullptr;
    for (const Config& c : configs)
        for (uint32_t i = 0; i < nCoords; i++) {
            const glm::fvec4 r = SampleImage<glm::fvec4>(state.get(), VK_FORMAT_R32G32B32A32_SFLOAT, data, range, 0u, levels, glm::fvec2(coords[2 * i], coords[2 * i + 1]), c.lod,
                                                         static_cast<VkFilter>(c.mag), static_cast<VkFilter>(c.min), static_cast<VkSamplerMipmapMode>(c.mipmap),
                                                         static_cast<VkSamplerAddressMode>(c.addressU), static_cast<VkSamplerAddressMode>(c.addressV), VK_SAMPLER_ADDRESS_MODE_REPEAT,
                                                         false, false, VK_COMPARE_OP_NEVER, static_cast<VkBorderColor>(c.border), false, VK_SAMPLER_REDUCTION_MODE_WEIGHTED_AVERAGE_EXT);
            uint32_t b[4];
            std::memcpy(b, &r.x, 16);
            std::printf("%u %u %u %u\n", b[0], b[1], b[2], b[3]);
        }
    return 0;
}
