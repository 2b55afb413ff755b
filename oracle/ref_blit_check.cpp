// ref_blit_check.cpp — runs the REFERENCE's own vkCmdBlitImage: BlitImageCommand::Process, CPVulkan/CommandBuffer.cpp:57-232 (the
// region loops, flipped extents, the u / v / w arithmetic, the per-texel SampleImage + SetPixel: SURVEY §8(f) f2, config C5),
// compiled IN PLACE from /root/reference by oracle/Makefile into oracle/_ref/blit_check. The method body is lifted out of the
// file where it lies by ref_slice.py (the translation unit as a whole needs the entire ICD) and becomes the body of the same
// method of a stand-in class with the same four members. It calls the reference's real ImageSampler.cpp (included below as a
// translation unit, as in ref_sampler_check.cpp) on real `Image` objects (CPVulkan/Image.h) whose fields are filled in here
// with the reference's own GetImageSize — Image::Create needs the ICD's allocator. Both images are R32G32B32A32_SFLOAT, whose
// JIT-compiled texel functions are raw 16-byte copies (ImageCompiler.cpp), so the arithmetic runs on exact texel values.
// TEST INFRASTRUCTURE ONLY: tests/golden/make_ref_golden.py stores its output, tests/test_reference_blit.py compares the
// oracle's cpvk_oracle_blit with it.
#include <array>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <memory>
#include <unordered_map>
#include <vector>

#include <Base.h>
#include <Formats.h>
#define private public // Image keeps its fields private and offers no constructor outside the ICD
#include <Image.h>
#undef private

#include <ImageSampler.cpp> // /root/reference/CPVulkan

ImageFunctions::ImageFunctions(CPJit* j) : jit(j) {}
ImageFunctions::~ImageFunctions() = default;

static void GetRGBA32F(const void* ptr, void* values) { std::memcpy(values, ptr, 16); }
static void SetRGBA32F(void* ptr, const float* values) { std::memcpy(ptr, values, 16); }
static FunctionPointer Unsupported() { std::fprintf(stderr, "blit_check: only R32G32B32A32_SFLOAT texel functions exist\n"); std::abort(); }
FunctionPointer CompileGetPixelDepth(CPJit*, const FormatInformation*) { return Unsupported(); }
FunctionPointer CompileGetPixelStencil(CPJit*, const FormatInformation*) { return Unsupported(); }
FunctionPointer CompileGetPixelF32(CPJit*, const FormatInformation* f) { return f->Format == VK_FORMAT_R32G32B32A32_SFLOAT ? reinterpret_cast<FunctionPointer>(GetRGBA32F) : Unsupported(); }
FunctionPointer CompileGetPixelI32(CPJit*, const FormatInformation*) { return Unsupported(); }
FunctionPointer CompileGetPixelU32(CPJit*, const FormatInformation*) { return Unsupported(); }
FunctionPointer CompileSetPixelDepthStencil(CPJit*, const FormatInformation*) { return Unsupported(); }
FunctionPointer CompileSetPixelF32(CPJit*, const FormatInformation* f) { return f->Format == VK_FORMAT_R32G32B32A32_SFLOAT ? reinterpret_cast<FunctionPointer>(SetRGBA32F) : Unsupported(); }
FunctionPointer CompileSetPixelI32(CPJit*, const FormatInformation*) { return Unsupported(); }
FunctionPointer CompileSetPixelU32(CPJit*, const FormatInformation*) { return Unsupported(); }

struct Command {
    virtual ~Command() = default;
    virtual void Process(DeviceState* deviceState) = 0;
};
// CommandBuffer.cpp:31-43, :234-238: the command's four members, by the reference's names
class BlitImageCommand final : public Command {
public:
    BlitImageCommand(Image* srcImage, Image* dstImage, std::vector<VkImageBlit> regions, VkFilter filter) : srcImage{srcImage}, dstImage{dstImage}, regions{std::move(regions)}, filter{filter} {}
#include "blit_slices.inc" // void Process(DeviceState* deviceState) override { ... }
private:
    Image* srcImage;
    Image* dstImage;
    std::vector<VkImageBlit> regions;
    VkFilter filter;
};

static void Fill(Image& image, uint32_t width, uint32_t height, std::vector<float>& texels) {
    image.imageType = VK_IMAGE_TYPE_2D;
    image.format = VK_FORMAT_R32G32B32A32_SFLOAT;
    image.extent = VkExtent3D{width, height, 1};
    image.mipLevels = 1;
    image.arrayLayers = 1;
    image.samples = VK_SAMPLE_COUNT_1_BIT;
    image.imageSize = GetImageSize(GetFormatInformation(image.format), width, height, 1, 1, 1); // Image::Create, Image.cpp:150
    image.data = gsl::span<uint8_t>(reinterpret_cast<uint8_t*>(texels.data()), (std::ptrdiff_t)(texels.size() * 4));
}

// input (little endian): u32 nCases; per case 13 x i32 {srcW, srcH, dstW, dstH, filter, srcX0, srcY0, srcX1, srcY1, dstX0, dstY0, dstX1, dstY1},
// srcW*srcH*4 floats, dstW*dstH*4 floats (what the destination holds before). output file: per case the destination's floats after.
int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: blit_check input.bin output.bin\n"); return 2; }
    std::ifstream in(argv[1], std::ios::binary);
    std::ofstream out(argv[2], std::ios::binary);
    if (!in || !out) return 2;
    uint32_t nCases = 0;
    in.read(reinterpret_cast<char*>(&nCases), 4);
    auto state = std::make_unique<DeviceState>();
    state->jit = nullptr;
    for (uint32_t c = 0; c < nCases; c++) {
        int32_t h[13];
        in.read(reinterpret_cast<char*>(h), sizeof h);
        std::vector<float> src((size_t)h[0] * h[1] * 4), dst((size_t)h[2] * h[3] * 4);
        in.read(reinterpret_cast<char*>(src.data()), src.size() * 4);
        in.read(reinterpret_cast<char*>(dst.data()), dst.size() * 4);
        if (!in) return 2;
        Image srcImage, dstImage;
        Fill(srcImage, (uint32_t)h[0], (uint32_t)h[1], src);
        Fill(dstImage, (uint32_t)h[2], (uint32_t)h[3], dst);
        VkImageBlit region{};
        region.srcSubresource = VkImageSubresourceLayers{VK_IMAGE_ASPECT_COLOR_BIT, 0, 0, 1};
        region.dstSubresource = region.srcSubresource;
        region.srcOffsets[0] = VkOffset3D{h[5], h[6], 0}; region.srcOffsets[1] = VkOffset3D{h[7], h[8], 1};
        region.dstOffsets[0] = VkOffset3D{h[9], h[10], 0}; region.dstOffsets[1] = VkOffset3D{h[11], h[12], 1};
        BlitImageCommand command(&srcImage, &dstImage, std::vector<VkImageBlit>{region}, static_cast<VkFilter>(h[4]));
        command.Process(state.get());
        out.write(reinterpret_cast<const char*>(dst.data()), dst.size() * 4);
    }
    return out ? 0 : 2;
}
