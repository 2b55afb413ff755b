// ref_image_check.cpp — runs the REFERENCE's own shader-side image functions, CPVulkan/GlslFunctions.cpp:324-737: GetFormatOffset,
// GetImageRange, GetImageData (which levels and bytes of an image a view selects; texel-buffer views), Swizzle,
// ImageSampleExplicitLod / ImageSampleImplicitLod (LOD bias and clamps in front of SampleImage, the component swizzle behind it) and
// ImageFetch (texelFetch, also from a uniform texel buffer: BASELINE config C5's third item) — SURVEY §8(a) a13 — compiled IN PLACE
// from /root/reference by oracle/Makefile into oracle/_ref/image_check. The functions are lifted out of the file where it lies by
// ref_slice.py (the rest of that file needs the LLVM JIT's symbol table) and run against the reference's real ImageSampler.cpp
// (included as a translation unit), Image.h, ImageView.h, Buffer.h, BufferView.h, Sampler.h and DescriptorSet.h; the objects'
// private fields are filled in here, their Create functions need the ICD. Texels are R32G32B32A32_SFLOAT (raw 16-byte texel functions,
// as in ref_sampler_check.cpp). TEST INFRASTRUCTURE ONLY: tests/golden/make_ref_golden.py stores its output,
// tests/test_reference_image.py compares the oracle's ImageSampleExplicitLod / ImageFetch (oracle_sampler.h) with it.
#include <algorithm>
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
#define private public // these classes keep their fields private and are constructed only by the ICD's Create functions
#include <Buffer.h>
#include <BufferView.h>
#include <Image.h>
#include <ImageView.h>
#include <Sampler.h>
#undef private
#include <DescriptorSet.h>

#include <ImageSampler.cpp> // /root/reference/CPVulkan

ImageFunctions::ImageFunctions(CPJit* j) : jit(j) {}
ImageFunctions::~ImageFunctions() = default;

static void GetRGBA32F(const void* ptr, void* values) { std::memcpy(values, ptr, 16); }
static void SetRGBA32F(void* ptr, const float* values) { std::memcpy(ptr, values, 16); }
static FunctionPointer Unsupported() { std::fprintf(stderr, "image_check: only R32G32B32A32_SFLOAT texel functions exist\n"); std::abort(); }
FunctionPointer CompileGetPixelDepth(CPJit*, const FormatInformation*) { return Unsupported(); }
FunctionPointer CompileGetPixelStencil(CPJit*, const FormatInformation*) { return Unsupported(); }
FunctionPointer CompileGetPixelF32(CPJit*, const FormatInformation* f) { return f->Format == VK_FORMAT_R32G32B32A32_SFLOAT ? reinterpret_cast<FunctionPointer>(GetRGBA32F) : Unsupported(); }
FunctionPointer CompileGetPixelI32(CPJit*, const FormatInformation*) { return Unsupported(); }
FunctionPointer CompileGetPixelU32(CPJit*, const FormatInformation*) { return Unsupported(); }
FunctionPointer CompileSetPixelDepthStencil(CPJit*, const FormatInformation*) { return Unsupported(); }
FunctionPointer CompileSetPixelF32(CPJit*, const FormatInformation* f) { return f->Format == VK_FORMAT_R32G32B32A32_SFLOAT ? reinterpret_cast<FunctionPointer>(SetRGBA32F) : Unsupported(); }
FunctionPointer CompileSetPixelI32(CPJit*, const FormatInformation*) { return Unsupported(); }
FunctionPointer CompileSetPixelU32(CPJit*, const FormatInformation*) { return Unsupported(); }

namespace glm {
template <int L, typename T> vec<L, T> abs(const vec<L, T>& a) { vec<L, T> r; for (int i = 0; i < L; i++) r[i] = a[i] < 0 ? -a[i] : a[i]; return r; } // GetImageDataCube (not exercised)
}

#include "image_slices.inc" // written by oracle/ref_slice.py into the scratch build directory (-I)

template <typename T> static bool Read(std::ifstream& in, T* v, size_t n = 1) { in.read(reinterpret_cast<char*>(v), (std::streamsize)(sizeof(T) * n)); return (bool)in; }

// input (little endian): u32 nCases; per case
//   u32 kind (0 = sample an image, 1 = fetch from an image, 2 = fetch from a texel buffer), width, height, mipLevels, baseMipLevel,
//       levelCount (0xFFFFFFFF = VK_REMAINING_MIP_LEVELS), swizzle r, g, b, a, magFilter, minFilter, mipmapMode, addressU, addressV,
//       borderColor, nCoords, bufferViewOffsetBytes, bufferViewRangeBytes; f32 mipLodBias, minLod, maxLod;
//   the texel bytes (the whole mip chain as the reference lays it out, or the buffer), u32 byte count first;
//   nCoords x {f32 u, v, lod} (kind 0) or {i32 x, y, unused} (kinds 1, 2)
// output file: per coordinate four result words
int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: image_check input.bin output.bin\n"); return 2; }
    std::ifstream in(argv[1], std::ios::binary);
    std::ofstream out(argv[2], std::ios::binary);
    if (!in || !out) return 2;
    uint32_t nCases = 0;
    if (!Read(in, &nCases)) return 2;
    auto state = std::make_unique<DeviceState>();
    state->jit = nullptr;
    for (uint32_t c = 0; c < nCases; c++) {
        uint32_t u[19]; float f[3]; uint32_t nBytes;
        if (!Read(in, u, 19) || !Read(in, f, 3) || !Read(in, &nBytes)) return 2;
        std::vector<uint8_t> bytes(nBytes + 16);
        if (nBytes && !Read(in, bytes.data(), nBytes)) return 2;
        std::vector<uint32_t> coords((size_t)u[16] * 3);
        if (u[16] && !Read(in, coords.data(), coords.size())) return 2;

        Image image;
        image.imageType = VK_IMAGE_TYPE_2D;
        image.format = VK_FORMAT_R32G32B32A32_SFLOAT;
        image.extent = VkExtent3D{u[1], u[2], 1};
        image.mipLevels = u[3];
        image.arrayLayers = 1;
        image.samples = VK_SAMPLE_COUNT_1_BIT;
        image.imageSize = GetImageSize(GetFormatInformation(image.format), u[1], u[2], 1, 1, u[3]); // Image::Create, Image.cpp:150
        image.data = gsl::span<uint8_t>(bytes.data(), (std::ptrdiff_t)bytes.size());
        ImageView view;
        view.image = &image;
        view.viewType = VK_IMAGE_VIEW_TYPE_2D;
        view.format = image.format;
        view.components = VkComponentMapping{static_cast<VkComponentSwizzle>(u[6]), static_cast<VkComponentSwizzle>(u[7]), static_cast<VkComponentSwizzle>(u[8]), static_cast<VkComponentSwizzle>(u[9])};
        view.subresourceRange = VkImageSubresourceRange{VK_IMAGE_ASPECT_COLOR_BIT, u[4], u[5], 0, 1};
        Buffer buffer;
        buffer.data = gsl::span<uint8_t>(bytes.data(), (std::ptrdiff_t)bytes.size());
        buffer.size = bytes.size();
        BufferView bufferView;
        bufferView.buffer = &buffer;
        bufferView.format = VK_FORMAT_R32G32B32A32_SFLOAT;
        bufferView.offset = u[17];
        bufferView.range = u[18];
        Sampler sampler;
        sampler.magFilter = static_cast<VkFilter>(u[10]);
        sampler.minFilter = static_cast<VkFilter>(u[11]);
        sampler.mipmapMode = static_cast<VkSamplerMipmapMode>(u[12]);
        sampler.addressModeU = static_cast<VkSamplerAddressMode>(u[13]);
        sampler.addressModeV = static_cast<VkSamplerAddressMode>(u[14]);
        sampler.addressModeW = VK_SAMPLER_ADDRESS_MODE_REPEAT;
        sampler.borderColour = static_cast<VkBorderColor>(u[15]);
        sampler.mipLodBias = f[0];
        sampler.minLod = f[1];
        sampler.maxLod = f[2];
        ImageDescriptor descriptor{};
        descriptor.ImageSampler = &sampler;
        if (u[0] == 2) { descriptor.Type = ImageDescriptorType::Buffer; descriptor.Data.Buffer = &bufferView; }
        else { descriptor.Type = ImageDescriptorType::Image; descriptor.Data.Image = &view; }

        for (uint32_t i = 0; i < u[16]; i++) {
            glm::fvec4 r(0.0f);
            if (u[0] == 0) {
                float uvl[3];
                std::memcpy(uvl, &coords[3 * i], 12);
                glm::fvec2 uv(uvl[0], uvl[1]);
                ImageSampleExplicitLod<glm::fvec4, glm::fvec2>(state.get(), &r, &descriptor, &uv, uvl[2]);
            } else if (u[0] == 1) {
                glm::ivec2 xy((int32_t)coords[3 * i], (int32_t)coords[3 * i + 1]);
                ImageFetch<glm::fvec4, glm::ivec2>(state.get(), &r, &descriptor, &xy);
            } else {
                ImageFetch<glm::fvec4, int32_t>(state.get(), &r, &descriptor, (int32_t)coords[3 * i]);
            }
            out.write(reinterpret_cast<const char*>(&r), 16);
        }
    }
    return out ? 0 : 2;
}
