// ref_formats_check.cpp — prints what the REFERENCE's own CPVulkanBase/Formats.cpp (format table, image layout) and
// CPVulkanBase/FloatFormat.h (half <-> float) compute. Both are compiled IN PLACE from /root/reference by
// oracle/Makefile into oracle/_ref/formats_check; the Vulkan SDK and MS-GSL they include are absent from this image, so
// oracle/shim/ supplies just the public API names they use (spec enum values, a minimal gsl::span).
// TEST INFRASTRUCTURE ONLY: tests/test_reference_formats.py compares these lines with the oracle's format table,
// GetImagePixelOffset restatement and half codec — the only arithmetic of the path, besides
// the SPIR-V reader, that the reference can execute here.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <type_traits>
#include <algorithm>
#include <cassert>

#include <Formats.h>      // /root/reference/CPVulkanBase
#include <FloatFormat.h>  // /root/reference/CPVulkanBase

static uint32_t FloatBits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static float BitsFloat(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

int main(int argc, char** argv) {
    const char* what = argc > 1 ? argv[1] : "formats";
    if (!std::strcmp(what, "formats")) {
        // one line per uncompressed core format: F id type total element base v0 v1 v2 v3 b0 b1 b2 b3
        for (int f = 1; f <= 130; f++) {
            const FormatInformation& fi = GetFormatInformation(static_cast<VkFormat>(f));
            std::printf("F %d %d %u %u %d", f, (int)fi.Type, fi.TotalSize, fi.ElementSize, (int)fi.Base);
            if (fi.Type == FormatType::Normal) std::printf(" %u %u %u %u 0 0 0 0", fi.Normal.RedOffset, fi.Normal.GreenOffset, fi.Normal.BlueOffset, fi.Normal.AlphaOffset);
            else if (fi.Type == FormatType::Packed) std::printf(" %u %u %u %u %u %u %u %u", fi.Packed.RedOffset, fi.Packed.GreenOffset, fi.Packed.BlueOffset, fi.Packed.AlphaOffset,
                                                                fi.Packed.RedBits, fi.Packed.GreenBits, fi.Packed.BlueBits, fi.Packed.AlphaBits);
            else if (fi.Type == FormatType::DepthStencil) std::printf(" %u %u 0 0 0 0 0 0", fi.DepthStencil.DepthOffset, fi.DepthStencil.StencilOffset);
            else std::printf(" 0 0 0 0 0 0 0 0");
            std::printf("\n");
        }
    } else if (!std::strcmp(what, "layout")) {
        // L format w h d layers mips | total layerSize pixelSize | per level: offset stride planeSize w h d | probes
        const int fmts[] = {9, 37, 44, 97, 109, 124, 126, 129, 130, 64};
        const uint32_t dims[][5] = {{500, 500, 1, 1, 1}, {256, 256, 1, 1, 9}, {3840, 2160, 1, 1, 1}, {7680, 4320, 1, 2, 3}, {17, 5, 3, 4, 3}, {1, 1, 1, 1, 1}, {33, 1, 1, 6, 2}};
        for (int f : fmts)
            for (const auto& d : dims) {
                const FormatInformation& fi = GetFormatInformation(static_cast<VkFormat>(f));
                const ImageSize s = GetImageSize(fi, d[0], d[1], d[2], d[3], d[4]);
                std::printf("L %d %u %u %u %u %u | %llu %llu %llu |", f, d[0], d[1], d[2], d[3], d[4], (unsigned long long)s.TotalSize, (unsigned long long)s.LayerSize, (unsigned long long)s.PixelSize);
                for (uint32_t l = 0; l < d[4]; l++)
                    std::printf(" %llu %llu %llu %u %u %u", (unsigned long long)s.Level[l].Offset, (unsigned long long)s.Level[l].Stride, (unsigned long long)s.Level[l].PlaneSize,
                                s.Level[l].Width, s.Level[l].Height, s.Level[l].Depth);
                std::printf(" |");
                for (uint32_t l = 0; l < d[4]; l++) {
                    const int32_t i = (int32_t)(s.Level[l].Width - 1), j = (int32_t)(s.Level[l].Height / 2), k = (int32_t)(s.Level[l].Depth - 1);
                    std::printf(" %llu", (unsigned long long)GetImagePixelOffset(s, i, j, k, l, d[3] - 1));
                }
                std::printf("\n");
            }
    } else if (!std::strcmp(what, "half")) {
        // every half code -> float bits
        for (uint32_t h = 0; h < 65536; h++) std::printf("H %u %u\n", h, FloatBits(ConvertBits<half, float>((uint16_t)h)));
    } else if (!std::strcmp(what, "tohalf")) {
        // float bit patterns from stdin (one decimal u32 per line) -> half code
        unsigned long long u;
        while (std::scanf("%llu", &u) == 1) std::printf("T %llu %u\n", u, (unsigned)ConvertBits<float, half>(BitsFloat((uint32_t)u)));
    } else {
        return 2;
    }
    return 0;
}
