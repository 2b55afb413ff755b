// oracle_formats.h — CPU restatement of the reference's format table and per-format pixel codec.
//
// TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product path; it may be used only by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
// PARITY: GetFormatInformation, the image layout and FloatToHalf / HalfToFloat are PINNED against the reference's own
// CPVulkanBase/Formats.cpp and FloatFormat.h compiled in place (oracle/_ref/formats_check; tests/test_reference_formats.py,
// tests/golden/ref_formats.txt, ref_layout.txt, ref_half.npz). The pack / unpack arithmetic (UNORM / SNORM / sRGB / depth)
// is PARITY UNPINNED: the reference emits it as LLVM IR (ImageCompiler.cpp), which cannot run here; it follows the
// source line by line and is cross-checked by the numpy KATs in tests/test_oracle_kats.py.
//
// Follows:
//   CPVulkanBase/Formats.cpp:210-443 (table), :455-483 (GetNormalImageSize), :583-587 (GetImagePixelOffset)
//   LLVMRuntime/ImageCompiler.cpp:15-53 (conversions), :103-158 (sRGB->linear), :160-301 (unpack),
//     :452-493 / :541-600 (depth get), :608-640 (stencil get), :869-947 (depth/stencil set),
//     :1010-1348 (pack), :1350-1383 (linear->sRGB)
//   CPVulkanBase/FloatFormat.h:138-335 (half <-> float)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace oracle {

enum class FmtType { Invalid, Normal, Packed, DepthStencil };
enum class Base { Unknown, UNorm, SNorm, UScaled, SScaled, UInt, SInt, UFloat, SFloat, SRGB };
constexpr uint32_t INVALID_OFFSET = 0xFFFFFFFFu;

struct FormatInfo {
    FmtType type = FmtType::Invalid;
    Base base = Base::Unknown;
    uint32_t totalSize = 0;
    uint32_t elementSize = 0;
    uint32_t channels = 0;      // VK_COLOR_COMPONENT bits
    uint32_t offset[4] = {INVALID_OFFSET, INVALID_OFFSET, INVALID_OFFSET, INVALID_OFFSET}; // Normal: bytes; Packed: bit offset
    uint32_t bits[4] = {0, 0, 0, 0};                                                      // Packed only
    uint32_t depthOffset = INVALID_OFFSET, stencilOffset = INVALID_OFFSET;
};

// VkFormat numeric values (the reference table is indexed by the enum, Formats.cpp:444-453).
enum : uint32_t {
    F_UNDEFINED = 0,
    F_R8_UNORM = 9, F_R8G8_UNORM = 16, F_R8G8B8_UNORM = 23, F_B8G8R8_UNORM = 30, F_R8G8B8A8_UNORM = 37,
    F_B8G8R8A8_UNORM = 44, F_A8B8G8R8_UNORM_PACK32 = 51, F_A2R10G10B10_UNORM_PACK32 = 58,
    F_A2B10G10R10_UNORM_PACK32 = 64, F_R16_UNORM = 70, F_R16G16_UNORM = 77, F_R16G16B16_UNORM = 84,
    F_R16G16B16A16_UNORM = 91, F_R16G16B16A16_SFLOAT = 97, F_R32_UINT = 98, F_R32G32_UINT = 101,
    F_R32G32B32_UINT = 104, F_R32G32B32A32_UINT = 107, F_R32G32B32A32_SFLOAT = 109,
    F_D16_UNORM = 124, F_X8_D24_UNORM_PACK32 = 125, F_D32_SFLOAT = 126, F_S8_UINT = 127,
    F_D16_UNORM_S8_UINT = 128, F_D24_UNORM_S8_UINT = 129, F_D32_SFLOAT_S8_UINT = 130,
};

// Formats.cpp:219-341 restated by rule instead of by row: every family below is a run of consecutive enum
// values whose rows differ only in BaseType.
inline FormatInfo GetFormatInformation(uint32_t f) {
    FormatInfo r;
    static const Base seven8[7] = {Base::UNorm, Base::SNorm, Base::UScaled, Base::SScaled, Base::UInt, Base::SInt, Base::SRGB};
    static const Base seven16[7] = {Base::UNorm, Base::SNorm, Base::UScaled, Base::SScaled, Base::UInt, Base::SInt, Base::SFloat};
    static const Base six[6] = {Base::UNorm, Base::SNorm, Base::UScaled, Base::SScaled, Base::UInt, Base::SInt};
    static const Base three[3] = {Base::UInt, Base::SInt, Base::SFloat};
    auto normal = [&](Base b, uint32_t elem, int comps, bool bgr) {
        r.type = FmtType::Normal; r.base = b; r.elementSize = elem; r.totalSize = elem * comps;
        r.channels = (1u << comps) - 1;
        for (int c = 0; c < comps; c++) r.offset[c] = elem * c;
        if (bgr) { r.offset[0] = elem * 2; r.offset[2] = 0; }
    };
    auto packed = [&](Base b, uint32_t size, uint32_t ro, uint32_t go, uint32_t bo, uint32_t ao,
                      uint32_t rb, uint32_t gb, uint32_t bb, uint32_t ab) {
        r.type = FmtType::Packed; r.base = b; r.totalSize = size; r.elementSize = 0; r.channels = 0xF;
        r.offset[0] = ro; r.offset[1] = go; r.offset[2] = bo; r.offset[3] = ao;
        r.bits[0] = rb; r.bits[1] = gb; r.bits[2] = bb; r.bits[3] = ab;
    };
    auto depth = [&](uint32_t size, uint32_t elem, Base b, uint32_t d, uint32_t s) {
        r.type = FmtType::DepthStencil; r.base = b; r.totalSize = size; r.elementSize = elem; r.channels = 1;
        r.depthOffset = d; r.stencilOffset = s;
    };
    if (f >= 9 && f <= 15) normal(seven8[f - 9], 1, 1, false);
    else if (f >= 16 && f <= 22) normal(seven8[f - 16], 1, 2, false);
    else if (f >= 23 && f <= 29) normal(seven8[f - 23], 1, 3, false);
    else if (f >= 30 && f <= 36) normal(seven8[f - 30], 1, 3, true);
    else if (f >= 37 && f <= 43) normal(seven8[f - 37], 1, 4, false);
    else if (f >= 44 && f <= 50) normal(seven8[f - 44], 1, 4, true);
    else if (f >= 51 && f <= 57) packed(seven8[f - 51], 4, 0, 8, 16, 24, 8, 8, 8, 8);
    else if (f >= 58 && f <= 63) packed(six[f - 58], 4, 20, 10, 0, 30, 10, 10, 10, 2);
    else if (f >= 64 && f <= 69) packed(six[f - 64], 4, 0, 10, 20, 30, 10, 10, 10, 2);
    else if (f >= 70 && f <= 76) normal(seven16[f - 70], 2, 1, false);
    else if (f >= 77 && f <= 83) normal(seven16[f - 77], 2, 2, false);
    else if (f >= 84 && f <= 90) normal(seven16[f - 84], 2, 3, false);
    else if (f >= 91 && f <= 97) normal(seven16[f - 91], 2, 4, false);
    else if (f >= 98 && f <= 100) normal(three[f - 98], 4, 1, false);
    else if (f >= 101 && f <= 103) normal(three[f - 101], 4, 2, false);
    else if (f >= 104 && f <= 106) normal(three[f - 104], 4, 3, false);
    else if (f >= 107 && f <= 109) normal(three[f - 107], 4, 4, false);
    else if (f == 124) depth(2, 2, Base::UNorm, 0, INVALID_OFFSET);
    else if (f == 125) depth(4, 4, Base::UNorm, 0, INVALID_OFFSET);
    else if (f == 126) depth(4, 4, Base::SFloat, 0, INVALID_OFFSET);
    else if (f == 127) depth(1, 1, Base::UInt, INVALID_OFFSET, 0);
    else if (f == 128) depth(3, 2, Base::UNorm, 0, 2);
    else if (f == 129) depth(4, 3, Base::UNorm, 0, 3);
    else if (f == 130) depth(8, 4, Base::SFloat, 0, 4);
    return r;
}

// ---- half <-> float: FloatFormat.h:138-255 (Truncate), :257-335 (Extend) ----
inline uint16_t FloatToHalf(float value) {
    uint32_t aRep; std::memcpy(&aRep, &value, 4);
    const uint32_t aAbs = aRep & 0x7FFFFFFFu;
    const uint32_t sign = aRep & 0x80000000u;
    const uint32_t underflow = (127u + 1 - 15) << 23, overflow = (127u + 31 - 15) << 23;
    uint16_t absResult;
    if (aAbs - underflow < aAbs - overflow) {
        absResult = (uint16_t)(aAbs >> 13);
        absResult -= (uint16_t)((uint16_t)(127 - 15) << 10);
        const uint32_t roundBits = aAbs & 0x1FFFu;
        if (roundBits > 0x1000u) ++absResult;
        else if (roundBits == 0x1000u) absResult += absResult & 1;
    } else if (aAbs > 0x7F800000u) {
        absResult = (uint16_t)(31u << 10);
        absResult |= 0x200;
        absResult |= (uint16_t)(((aAbs & 0x3FFFFFu) >> 13) & 0x1FFu);
    } else if (aAbs >= overflow) {
        absResult = (uint16_t)(31u << 10);
    } else {
        const int aExp = (int)(aAbs >> 23);
        const int shift = 127 - 15 - aExp + 1;
        const uint32_t significand = (aRep & 0x7FFFFFu) | 0x800000u;
        if (shift > 23) {
            absResult = 0;
        } else {
            const uint32_t sticky = (significand << (32 - shift)) ? 1 : 0;
            const uint32_t den = (significand >> shift) | sticky;
            absResult = (uint16_t)(den >> 13);
            const uint32_t roundBits = den & 0x1FFFu;
            if (roundBits > 0x1000u) ++absResult;
            else if (roundBits == 0x1000u) absResult += absResult & 1;
        }
    }
    return (uint16_t)(absResult | (sign >> 16));
}

inline float HalfToFloat(uint16_t value) {
    const uint32_t aAbs = value & 0x7FFFu;
    const uint32_t sign = value & 0x8000u;
    uint32_t absResult;
    if ((uint16_t)(aAbs - 0x400u) < (uint16_t)(0x7C00u - 0x400u)) {
        absResult = aAbs << 13;
        absResult += (uint32_t)(127 - 15) << 23;
    } else if (aAbs >= 0x7C00u) {
        absResult = 0xFFu << 23;
        absResult |= (aAbs & 0x200u) << 13;
        absResult |= (aAbs & 0x1FFu) << 13;
    } else if (aAbs) {
        const int scale = __builtin_clz(aAbs) - __builtin_clz(0x400u);
        absResult = aAbs << (13 + scale);
        absResult ^= 0x800000u;
        const uint32_t resultExponent = (uint32_t)(127 - 15 - scale + 1);
        absResult |= resultExponent << 23;
    } else {
        absResult = 0;
    }
    const uint32_t result = absResult | (sign << 16);
    float f; std::memcpy(&f, &result, 4);
    return f;
}

// llvm.maxnum / llvm.minnum: if one operand is NaN return the other (ImageCompiler.cpp:524-538).
inline float MaxNum(float a, float b) { if (std::isnan(a)) return b; if (std::isnan(b)) return a; return a > b ? a : b; }
inline float MinNum(float a, float b) { if (std::isnan(a)) return b; if (std::isnan(b)) return a; return a < b ? a : b; }

// ImageCompiler.cpp:103-158
inline float SRGBToLinear(float v) {
    if (!(v <= 0.04045f)) { // FCmpUGT: true when unordered
        float t = v + 0.055f;
        t = t / 1.055f;
        return powf(t, 2.4f);
    }
    return v / 12.92f;
}
// ImageCompiler.cpp:1350-1383
inline float LinearToSRGB(float v) {
    if (!(v <= 0.0031308f)) {
        float t = powf(v, 1.0f / 2.4f);
        t = t * 1.055f;
        return t + -0.055f;
    }
    return v * 12.92f;
}

// EmitConvertFloatUInt / EmitConvertFloatInt (ImageCompiler.cpp:20-32): fmul, llvm.round, fptoui/fptosi.
inline uint32_t FloatToUNormBits(float v, uint32_t maxValue) {
    float t = v * (float)(double)maxValue;
    t = roundf(t);
    return (uint32_t)(int64_t)t;
}
inline int32_t FloatToSNormBits(float v, int32_t maxValue) {
    float t = v * (float)(double)maxValue;
    t = roundf(t);
    return (int32_t)t;
}

inline uint32_t LoadBits(const uint8_t* p, uint32_t size) {
    uint32_t v = 0; std::memcpy(&v, p, size); return v;
}
inline void StoreBits(uint8_t* p, uint32_t v, uint32_t size) { std::memcpy(p, &v, size); }

// GetPixelF32 (ImageCompiler.cpp:648-849 -> EmitGetPixel :495-511). Missing channels read 0,0,0,1.
inline void GetPixelF32(const FormatInfo& fi, uint32_t format, const uint8_t* src, float out[4]) {
    out[0] = 0; out[1] = 0; out[2] = 0; out[3] = 1;
    if (fi.type == FmtType::Normal) {
        for (int c = 0; c < 4; c++) {
            if (fi.offset[c] == INVALID_OFFSET) continue;
            const uint8_t* p = src + (fi.offset[c] / fi.elementSize) * fi.elementSize;
            const uint32_t raw = LoadBits(p, fi.elementSize);
            const uint32_t umax = fi.elementSize == 1 ? 0xFFu : fi.elementSize == 2 ? 0xFFFFu : 0xFFFFFFFFu;
            float v = 0;
            switch (fi.base) {
            case Base::UNorm: case Base::UScaled: case Base::UInt: // EmitConvertUIntFloat: uitofp / max
                v = (float)raw / (float)(double)umax; break;
            case Base::SNorm: case Base::SScaled: case Base::SInt: { // EmitConvertIntFloat: sitofp / max
                int32_t s = fi.elementSize == 1 ? (int32_t)(int8_t)raw : fi.elementSize == 2 ? (int32_t)(int16_t)raw : (int32_t)raw;
                v = (float)s / (float)(double)(umax >> 1); break; }
            case Base::SFloat:
                if (fi.elementSize == 2) v = HalfToFloat((uint16_t)raw); else std::memcpy(&v, &raw, 4);
                break;
            case Base::SRGB:
                v = (float)raw / (float)(double)umax;
                if (c != 3) v = SRGBToLinear(v);
                break;
            default: break;
            }
            out[c] = v;
        }
    } else if (fi.type == FmtType::Packed) {
        const uint32_t source = LoadBits(src, fi.totalSize);
        for (int c = 0; c < 4; c++) {
            const uint32_t bits = fi.bits[c];
            const uint32_t mask = (uint32_t)((1ull << bits) - 1);
            uint32_t value = (source >> fi.offset[c]) & mask;
            float v = 0;
            switch (fi.base) {
            case Base::UNorm: v = (float)value / (float)mask; break;
            case Base::SNorm: {
                int32_t s;
                if (bits == 8) s = (int8_t)value; else if (bits == 16) s = (int16_t)value;
                else s = ((int32_t)(value << (32 - bits))) >> (32 - bits);
                v = (float)s / (float)(mask >> 1); break; }
            case Base::SRGB: v = (float)value / (float)mask; if (c != 3) v = SRGBToLinear(v); break;
            default: v = 0; break;
            }
            out[c] = v;
        }
    } else if (fi.type == FmtType::DepthStencil) { // EmitGetDepthStencilPixel :452-493
        float v = 0;
        if (format == F_D16_UNORM || format == F_D16_UNORM_S8_UINT) v = (float)LoadBits(src, 2) / 65535.0f;
        else if (format == F_D24_UNORM_S8_UINT || format == F_X8_D24_UNORM_PACK32) v = (float)(LoadBits(src, 4) & 0xFFFFFFu) / 16777215.0f;
        else if (format == F_D32_SFLOAT || format == F_D32_SFLOAT_S8_UINT) std::memcpy(&v, src, 4);
        out[0] = v;
    }
}

// GetPixelI32 / GetPixelU32: integer formats only; sign by format base (EmitGetNormalPixel :358-381).
inline void GetPixelInt(const FormatInfo& fi, const uint8_t* src, uint32_t out[4]) {
    out[0] = 0; out[1] = 0; out[2] = 0; out[3] = 1;
    const bool isSigned = fi.base == Base::SInt;
    if (fi.type == FmtType::Normal) {
        for (int c = 0; c < 4; c++) {
            if (fi.offset[c] == INVALID_OFFSET) continue;
            const uint32_t raw = LoadBits(src + fi.offset[c], fi.elementSize);
            if (isSigned) out[c] = fi.elementSize == 1 ? (uint32_t)(int32_t)(int8_t)raw : fi.elementSize == 2 ? (uint32_t)(int32_t)(int16_t)raw : raw;
            else out[c] = raw;
        }
    } else if (fi.type == FmtType::Packed) {
        const uint32_t source = LoadBits(src, fi.totalSize);
        for (int c = 0; c < 4; c++) {
            const uint32_t bits = fi.bits[c];
            const uint32_t mask = (uint32_t)((1ull << bits) - 1);
            uint32_t value = (source >> fi.offset[c]) & mask;
            if (isSigned && bits != 32) value = (uint32_t)(((int32_t)(value << (32 - bits))) >> (32 - bits));
            out[c] = value;
        }
    }
}

// SetPixelF32 (ImageCompiler.cpp:956-1348).
inline void SetPixelF32(const FormatInfo& fi, uint8_t* dst, const float in[4]) {
    if (fi.type == FmtType::Normal) {
        for (int c = 0; c < 4; c++) {
            if (fi.offset[c] == INVALID_OFFSET) continue;
            uint8_t* p = dst + (fi.offset[c] / fi.elementSize) * fi.elementSize;
            const uint32_t umax = fi.elementSize == 1 ? 0xFFu : fi.elementSize == 2 ? 0xFFFFu : 0xFFFFFFFFu;
            switch (fi.base) {
            case Base::UNorm: {
                float v = MinNum(MaxNum(in[c], 0.0f), 1.0f);
                StoreBits(p, FloatToUNormBits(v, umax), fi.elementSize); break; }
            case Base::SNorm: {
                float v = MinNum(MaxNum(in[c], -1.0f), 1.0f);
                StoreBits(p, (uint32_t)FloatToSNormBits(v, (int32_t)(umax >> 1)), fi.elementSize); break; }
            case Base::SFloat:
                if (fi.elementSize == 2) StoreBits(p, FloatToHalf(in[c]), 2); else std::memcpy(p, &in[c], 4);
                break;
            case Base::SRGB: {
                float v = MinNum(MaxNum(in[c], 0.0f), 1.0f);
                if (c != 3) v = LinearToSRGB(v);
                StoreBits(p, FloatToUNormBits(v, umax), fi.elementSize); break; }
            default: break; // UScaled/SScaled/UFloat: TODO_ERROR in the reference
            }
        }
    } else if (fi.type == FmtType::Packed) {
        uint32_t value = 0;
        for (int c = 0; c < 4; c++) {
            const uint32_t bits = fi.bits[c];
            if (!bits) continue;
            const uint32_t mask = (uint32_t)((1ull << bits) - 1);
            uint32_t ch = 0;
            switch (fi.base) {
            case Base::UNorm: { float v = MinNum(MaxNum(in[c], 0.0f), 1.0f); ch = (uint32_t)(int64_t)roundf(v * (float)mask); break; }
            case Base::SNorm: { float v = MinNum(MaxNum(in[c], -1.0f), 1.0f); ch = (uint32_t)(int64_t)roundf(v * (float)(mask >> 1)); break; }
            case Base::SRGB: { float v = MinNum(MaxNum(in[c], 0.0f), 1.0f); if (c != 3) v = LinearToSRGB(v); ch = (uint32_t)(int64_t)roundf(v * (float)mask); break; }
            default: break;
            }
            value |= ch << fi.offset[c];
        }
        StoreBits(dst, value, fi.totalSize);
    }
}

// SetPixelU32 / SetPixelI32 (ImageCompiler.cpp:1121-1177, :1296-1317): clamp to range, truncate.
inline void SetPixelInt(const FormatInfo& fi, uint8_t* dst, const uint32_t in[4]) {
    const bool isSigned = fi.base == Base::SInt;
    if (fi.type == FmtType::Normal) {
        for (int c = 0; c < 4; c++) {
            if (fi.offset[c] == INVALID_OFFSET) continue;
            uint32_t v = in[c];
            if (isSigned) {
                int32_t s = (int32_t)v;
                if (fi.elementSize == 1) { s = s > -128 ? s : -128; s = s < 127 ? s : 127; }
                else if (fi.elementSize == 2) { s = s > -32768 ? s : -32768; s = s < 32767 ? s : 32767; }
                v = (uint32_t)s;
            } else {
                if (fi.elementSize == 1) v = v < 255u ? v : 255u;
                else if (fi.elementSize == 2) v = v < 65535u ? v : 65535u;
            }
            StoreBits(dst + fi.offset[c], v, fi.elementSize);
        }
    } else if (fi.type == FmtType::Packed) {
        uint32_t value = 0;
        for (int c = 0; c < 4; c++) {
            const uint32_t bits = fi.bits[c];
            if (!bits) continue;
            const uint32_t mask = (uint32_t)((1ull << bits) - 1);
            uint32_t v = in[c];
            if (isSigned) {
                const int32_t mn = -(int32_t)(1u << (bits - 1)), mx = (int32_t)((1u << (bits - 1)) - 1);
                int32_t s = (int32_t)v; s = s > mn ? s : mn; s = s < mx ? s : mx; v = (uint32_t)s;
            } else {
                v = v < mask ? v : mask;
            }
            value |= (v & mask) << fi.offset[c];
        }
        StoreBits(dst, value, fi.totalSize);
    }
}

// GetPixelDepth (ImageCompiler.cpp:541-600)
inline float GetDepth(uint32_t format, const uint8_t* src) {
    switch (format) {
    case F_D16_UNORM: case F_D16_UNORM_S8_UINT: return (float)LoadBits(src, 2) / 65535.0f;
    case F_D24_UNORM_S8_UINT: case F_X8_D24_UNORM_PACK32: return (float)(LoadBits(src, 4) & 0xFFFFFFu) / 16777215.0f;
    case F_D32_SFLOAT: case F_D32_SFLOAT_S8_UINT: { float v; std::memcpy(&v, src, 4); return v; }
    default: return 0.0f;
    }
}
// GetPixelStencil (ImageCompiler.cpp:608-640): byte at StencilOffset.
inline uint8_t GetStencil(const FormatInfo& fi, const uint8_t* src) { return src[fi.stencilOffset]; }

// SetPixelDepthStencil (ImageCompiler.cpp:869-947)
inline void SetDepthStencil(const FormatInfo& fi, uint32_t format, uint8_t* dst, float depth, uint8_t stencil) {
    if (fi.depthOffset != INVALID_OFFSET) {
        switch (format) {
        case F_D16_UNORM: case F_D16_UNORM_S8_UINT: {
            float v = MinNum(MaxNum(depth, 0.0f), 1.0f);
            StoreBits(dst, FloatToUNormBits(v, 0xFFFFu), 2); break; }
        case F_D24_UNORM_S8_UINT: case F_X8_D24_UNORM_PACK32: {
            float v = MinNum(MaxNum(depth, 0.0f), 1.0f);
            v = v * 16777215.0f; v = roundf(v);
            uint32_t u = (uint32_t)(int64_t)v;
            if (format == F_D24_UNORM_S8_UINT) u |= (uint32_t)stencil << 24;
            StoreBits(dst, u, 4); break; }
        case F_D32_SFLOAT: case F_D32_SFLOAT_S8_UINT: std::memcpy(dst, &depth, 4); break;
        default: break;
        }
    }
    if (fi.stencilOffset != INVALID_OFFSET && format != F_D24_UNORM_S8_UINT) dst[fi.stencilOffset] = stencil;
}

} // namespace oracle
