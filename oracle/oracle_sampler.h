// oracle_sampler.h — CPU restatement of the reference's texel fetch and filtering.
// TEST INFRASTRUCTURE ONLY (see oracle_formats.h).
// PARITY: SampleImage / SampleImageOfLevel / wrap / lerp are PINNED bit for bit against the reference's own
// CPVulkan/ImageSampler.cpp compiled in place (oracle/_ref/sampler_check; tests/test_reference_sampler.py,
// tests/golden/ref_sampler.npz: 90 sampler states x 128 coordinates). The GlslFunctions.cpp wrapper (LOD bias / clamp,
// swizzle, texel fetch) is PARITY UNPINNED (needs the whole ICD to compile).
//
// Follows:
//   CPVulkan/ImageSampler.cpp:12-38 (wrap), :40-55 (frac, lerp in double), :83-145 (GetPixel + border),
//     :363-410 (GetPixelLinear), :461-579 (SampleImageOfLevel), :581-673 (SampleImage: LOD / mip select,
//     missing-channel fix-up), :695-708 (sampler unpack)
//   CPVulkan/GlslFunctions.cpp:378-421 (GetImageData), :539-555 (Swizzle), :598-654 (ImageSampleExplicitLod),
//     :674-737 (ImageFetch)
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>

#include "../include/cpvk_cuda.h"
#include "oracle_formats.h"

namespace oracle {

constexpr float MAX_SAMPLER_LOD_BIAS = 32.0f; // CPVulkanBase/Config.h:156

struct Vec4f { float v[4]; };

inline int32_t Wrap(int32_t v, int32_t size, uint32_t mode) {
    switch (mode) {
    case 0: return (v % size + size) % size;                                    // REPEAT
    case 1: { const int32_t two = 2 * size; const int32_t n = (v % two + two) % two - size;
              return size - 1 - (n >= 0 ? n : -(1 + n)); }                     // MIRRORED_REPEAT
    case 2: return std::clamp(v, 0, size - 1);                                  // CLAMP_TO_EDGE
    case 3: return std::clamp(v, -1, size);                                     // CLAMP_TO_BORDER
    case 4: return std::clamp(v >= 0 ? v : -(1 + v), 0, size - 1);              // MIRROR_CLAMP_TO_EDGE
    default: return 0;
    }
}

// lerp(): float subtraction, product and sum in double, one rounding back to float (ImageSampler.cpp:51-55).
inline Vec4f Lerp(const Vec4f& mn, const Vec4f& mx, double delta) {
    Vec4f r;
    for (int i = 0; i < 4; i++) {
        const float d = mx.v[i] - mn.v[i];
        r.v[i] = (float)((double)mn.v[i] + (double)d * delta);
    }
    return r;
}

inline Vec4f BorderColour(uint32_t border) {
    static const float t[6][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 1}, {0, 0, 0, 1}, {1, 1, 1, 1}, {1, 1, 1, 1}};
    Vec4f r; for (int i = 0; i < 4; i++) r.v[i] = t[border < 6 ? border : 0][i]; return r;
}

// GetPixel<fvec4>(…, uvec2 range, ivec2 coordinates, border) ImageSampler.cpp:105-145 (1-D/3-D analogous).
inline Vec4f GetTexelF32(uint32_t format, const CpvkMipLevel& lvl, int dims, const int32_t c[3], const Vec4f& border) {
    const uint32_t range[3] = {lvl.width, lvl.height, lvl.depth};
    for (int i = 0; i < dims; i++)
        if (c[i] < 0 || (uint32_t)c[i] >= range[i]) return border;
    const FormatInfo fi = GetFormatInformation(format);
    const uint64_t stride = (uint64_t)fi.totalSize * lvl.width;
    const uint64_t plane = stride * lvl.height;
    const uint64_t off = (dims > 2 ? (uint64_t)c[2] * plane : 0) + (dims > 1 ? (uint64_t)c[1] * stride : 0) + (uint64_t)c[0] * fi.totalSize;
    Vec4f r;
    GetPixelF32(fi, format, (const uint8_t*)(uintptr_t)lvl.address + off, r.v);
    return r;
}

// SampleImageOfLevel (ImageSampler.cpp:461-579), float result, non-compressed formats.
inline Vec4f SampleLevel(uint32_t format, const CpvkMipLevel& lvl, int dims, const float coord[3], uint32_t filter,
                         const uint32_t addressMode[3], uint32_t borderColour) {
    const uint32_t range[3] = {lvl.width, lvl.height, lvl.depth};
    const Vec4f border = BorderColour(borderColour);
    if (filter == 0) { // NEAREST
        int32_t nc[3] = {0, 0, 0};
        for (int i = 0; i < dims; i++) {
            nc[i] = (int32_t)std::floor(coord[i] * (float)range[i] + 0.0f);
            nc[i] = Wrap(nc[i], (int32_t)range[i], addressMode[i]);
        }
        return GetTexelF32(format, lvl, dims, nc, border);
    }
    // LINEAR, weighted average
    int32_t c0[3] = {0, 0, 0}, c1[3] = {0, 0, 0};
    float interp[3] = {0, 0, 0};
    for (int i = 0; i < dims; i++) {
        c0[i] = (int32_t)std::floor(coord[i] * (float)range[i] - 0.5f);
        c1[i] = Wrap(c0[i] + 1, (int32_t)range[i], addressMode[i]);
        c0[i] = Wrap(c0[i], (int32_t)range[i], addressMode[i]);
        const float t = coord[i] * (float)range[i] - 0.5f;
        interp[i] = t - std::floor(t);
    }
    if (dims == 1) {
        const Vec4f i0 = GetTexelF32(format, lvl, 1, c0, border);
        const Vec4f i1 = GetTexelF32(format, lvl, 1, c1, border);
        return Lerp(i0, i1, interp[0]);
    }
    if (dims == 2) {
        const int32_t p00[3] = {c0[0], c0[1], 0}, p01[3] = {c0[0], c1[1], 0}, p10[3] = {c1[0], c0[1], 0}, p11[3] = {c1[0], c1[1], 0};
        const Vec4f i0j0 = GetTexelF32(format, lvl, 2, p00, border);
        const Vec4f i0j1 = GetTexelF32(format, lvl, 2, p01, border);
        const Vec4f i1j0 = GetTexelF32(format, lvl, 2, p10, border);
        const Vec4f i1j1 = GetTexelF32(format, lvl, 2, p11, border);
        const Vec4f ij0 = Lerp(i0j0, i1j0, interp[0]);
        const Vec4f ij1 = Lerp(i0j1, i1j1, interp[0]);
        return Lerp(ij0, ij1, interp[1]);
    }
    Vec4f t[2][2][2];
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
        const int32_t p[3] = {i ? c1[0] : c0[0], j ? c1[1] : c0[1], k ? c1[2] : c0[2]};
        t[i][j][k] = GetTexelF32(format, lvl, 3, p, border);
    }
    const Vec4f ij0k0 = Lerp(t[0][0][0], t[1][0][0], interp[0]);
    const Vec4f ij0k1 = Lerp(t[0][0][1], t[1][0][1], interp[0]);
    const Vec4f ij1k0 = Lerp(t[0][1][0], t[1][1][0], interp[0]);
    const Vec4f ij1k1 = Lerp(t[0][1][1], t[1][1][1], interp[0]);
    const Vec4f ijk0 = Lerp(ij0k0, ij1k0, interp[1]);
    const Vec4f ijk1 = Lerp(ij0k1, ij1k1, interp[1]);
    return Lerp(ijk0, ijk1, interp[2]);
}

// SampleImage (ImageSampler.cpp:581-673) with the sampler unpacked as in :695-708.
inline Vec4f SampleImage(const CpvkDescriptor& d, int dims, const float coord[3], float lod, const CpvkSampler& s,
                         uint32_t magFilter, uint32_t minFilter) {
    const uint32_t addressMode[3] = {s.addressModeU, s.addressModeV, s.addressModeW};
    Vec4f result;
    if (lod <= 0) {
        result = SampleLevel(d.format, d.levels[0], dims, coord, magFilter, addressMode, s.borderColor);
    } else {
        const float maxLevel = (float)(d.levelCount - 1);
        const float mipLevel = std::clamp(lod, 0.0f, maxLevel);
        if (s.mipmapMode == 0) {
            const uint32_t real = (uint32_t)std::ceil(mipLevel + 0.5f) - 1;
            result = SampleLevel(d.format, d.levels[real], dims, coord, minFilter, addressMode, s.borderColor);
        } else {
            const uint32_t l1 = (uint32_t)std::floor(mipLevel);
            const float delta = mipLevel - (float)l1;
            const Vec4f p1 = SampleLevel(d.format, d.levels[l1], dims, coord, minFilter, addressMode, s.borderColor);
            if (delta == 0) result = p1;
            else {
                const Vec4f p2 = SampleLevel(d.format, d.levels[l1 + 1], dims, coord, minFilter, addressMode, s.borderColor);
                result = Lerp(p1, p2, delta);
            }
        }
    }
    const FormatInfo fi = GetFormatInformation(d.format);
    if (!(fi.channels & 1)) result.v[0] = 0;
    if (!(fi.channels & 2)) result.v[1] = 0;
    if (!(fi.channels & 4)) result.v[2] = 0;
    if (!(fi.channels & 8)) result.v[3] = 1;
    return result;
}

inline float SwizzleOne(const Vec4f& v, uint32_t swz, int index) {
    switch (swz) {
    case 0: return v.v[index]; // IDENTITY
    case 1: return 0;          // ZERO
    case 2: return 1;          // ONE
    case 3: return v.v[0]; case 4: return v.v[1]; case 5: return v.v[2]; case 6: return v.v[3];
    default: return 0;
    }
}

inline void ApplySwizzle(const CpvkDescriptor& d, Vec4f& r) {
    const uint32_t* s = d.swizzle;
    if ((s[0] != 0 && s[0] != 3) || (s[1] != 0 && s[1] != 4) || (s[2] != 0 && s[2] != 5) || (s[3] != 0 && s[3] != 6)) {
        const Vec4f old = r;
        for (int i = 0; i < 4; i++) r.v[i] = SwizzleOne(old, s[i], i);
    }
}

// ImageSampleExplicitLod<fvec4, fvecN> (GlslFunctions.cpp:598-654); ImplicitLod = lod 0 (:656-672).
inline Vec4f ImageSampleExplicitLod(const CpvkDescriptor& d, const float coord[3], float lod) {
    const CpvkSampler& s = d.sampler;
    const float lambdaBase = lod;
    const float lambdaPrime = lambdaBase + std::clamp(s.mipLodBias + 0.0f, -MAX_SAMPLER_LOD_BIAS, MAX_SAMPLER_LOD_BIAS);
    const float lambda = std::clamp(lambdaPrime, s.minLod, s.maxLod);
    Vec4f r = SampleImage(d, (int)d.dimensions, coord, lambda, s, s.magFilter, s.minFilter);
    if (d.type == CPVK_DESC_IMAGE) ApplySwizzle(d, r);
    return r;
}

// ImageFetch<fvec4, ivecN> (GlslFunctions.cpp:674-737): level 0 of the view, transparent-black border,
// no missing-channel fix-up beyond what the codec does; texel buffers are 1-D with range = bytes / texel size.
inline Vec4f ImageFetch(const CpvkDescriptor& d, const int32_t coord[3]) {
    Vec4f border; border.v[0] = border.v[1] = border.v[2] = border.v[3] = 0;
    Vec4f r;
    if (d.type == CPVK_DESC_TEXEL_BUFFER) {
        const FormatInfo fi = GetFormatInformation(d.format);
        CpvkMipLevel lvl{};
        lvl.address = d.address; lvl.width = (uint32_t)d.range / fi.totalSize; lvl.height = 1; lvl.depth = 1;
        r = GetTexelF32(d.format, lvl, 1, coord, border);
    } else {
        r = GetTexelF32(d.format, d.levels[0], (int)d.dimensions, coord, border);
        ApplySwizzle(d, r);
    }
    return r;
}

} // namespace oracle
