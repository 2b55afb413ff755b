// oracle_draw.cpp — CPU restatement of CPVulkan's draw hot path (vkCmdDraw / vkCmdDrawIndexed execution).
//
// TEST INFRASTRUCTURE ONLY. May be used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs — as the checker, never as the thing shipped. The product path (libcpvk_cuda.so)
// never links or calls this.
// PARITY: the reference has no tests, golden vectors or published outputs for this path and the ICD cannot
// be built in this image (LLVM-8, Vulkan SDK, GSL, glm >= 0.9.9: SURVEY F10). What of the reference CAN execute here is compiled
// in place into oracle/_ref/ and pins this oracle bit for bit (tests/test_reference_*.py, fixtures under tests/golden/):
//   draw_check    CommandBuffer.Draw.cpp's own EdgeFunction, CalculatePrimitives, SetDatum, GetFragmentInput, DrawPixel, ProcessPoints,
//                 ProcessLines, ProcessTriangles (fragment streams of 33 draws, 137 612 fragments) and ApplyBlendFactor / ApplyBlend
//                 (5 050 states)  -> ProcessTriangles / ProcessLines / ProcessPoints / Fragment()'s interpolation / ApplyBlend below
//   math_check    SpirvFunctions.cpp + the GLSL.std.450 templates of GlslFunctions.cpp -> oracle_spirv.h's dot / matrix / GLSL code
//   image_check   GlslFunctions.cpp's own GetImageData / Swizzle / ImageSampleExplicitLod / ImageFetch (:324-737, 80 bindings x 24 coordinates)
//                 -> oracle_sampler.h's ImageSampleExplicitLod / ImageFetch
//   sampler_check ImageSampler.cpp -> oracle_sampler.h        formats_check  Formats.cpp + FloatFormat.h -> oracle_formats.h
//                 draw_check ia: ProcessInputAssembler / ProcessInputAssemblerIndexed (109 draws) -> AssembledVertexId below
//   blit_check    CommandBuffer.cpp's own BlitImageCommand::Process (:57-232) on real Image objects -> cpvk_oracle_blit below (36 blits:
//                 scaled, flipped, offset, one-texel, both filters)
//   interface_check  Draw.cpp's own GetVariableFormat / GetVariableSize / GetVariablePointers on modules loaded by SPIRVParser/ -> Reflect's
//                 fragment inputs (Location, format, interpolation, size, offset) for every fragment shader in the tree
//   spirv_check   SPIRVParser/ -> the hand-assembled shaders
// NOT pinned by execution (the reference emits them as LLVM IR, which needs LLVM to run): the late depth / stencil epilogue and
// attachment write of the fragment wrapper (PipelineCompiler.cpp) and the per-format UNORM / SNORM / sRGB pack / unpack arithmetic
// (ImageCompiler.cpp) — held by the independent numpy KATs in tests/test_oracle_kats.py. glm: the checkers link the 0.9.5.3 copy
// vendored with the reference; min / max with NaN or +-0 and normalize(vec4) follow glm >= 0.9.9 here.
//
// Follows, in execution order:
//   CPVulkan/CommandBuffer.Draw.cpp:675-760  ProcessInputAssembler[Indexed]          (IA)
//   CPVulkan/CommandBuffer.Draw.cpp:567-673  CalculatePrimitives                      (topology)
//   CPVulkan/CommandBuffer.Draw.cpp:776-814 + LLVMRuntime/PipelineCompiler.cpp:821-981 (vertex fetch, VS, record store)
//   CPVulkan/CommandBuffer.Draw.cpp:1510-1594 ProcessTriangles                        (setup, bbox, pixel loop)
//   CPVulkan/CommandBuffer.Draw.cpp:410-418, 874-954 EdgeFunction / GetFragmentInput  (coverage, interpolation)
//   CPVulkan/CommandBuffer.Draw.cpp:1300-1313 DrawPixel                               (viewport depth transform)
//   LLVMRuntime/PipelineCompiler.cpp:1020-1527 fragment wrapper                       (late depth/stencil)
//   LLVMRuntime/PipelineCompiler.cpp:1528-1727 + CPVulkan/GlslFunctions.cpp:842-928   (attachment write)
//   CPVulkan/CommandBuffer.Draw.cpp:956-1262  ApplyBlendFactor / ApplyBlend           (blend: dead code in the
//       reference, PipelineCompiler.cpp:1674-1683 aborts on blendEnable — intended semantics only)
// Arithmetic: IEEE binary32, one rounding per operator, no FMA (built with -ffp-contract=off, no -march).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../include/cpvk_cuda.h"
#include "oracle_formats.h"
#include "oracle_sampler.h"
#include "oracle_spirv.h"

using namespace oracle;
using namespace oracle::spv;

namespace {

thread_local std::string g_error;

// ---- reflection of the shader interface, as GetVariablePointers does it (Draw.cpp:420-565) ----

// GetVariableSize (Draw.cpp:297-354): the FRAGMENT side's idea of a varying's size.
uint32_t VariableSize(const Module& m, uint32_t ty) {
    const Type& t = m.types[ty];
    switch (t.kind) {
    case Type::Array: return VariableSize(m, t.elem) * t.count;
    case Type::Struct: { uint32_t s = 0; for (uint32_t k = 0; k < t.members.size(); k++) { uint32_t off;
        if (m.memberDecoVal(ty, k, DecoOffset, off) && off > s) s = off; s += VariableSize(m, t.members[k]); } return s; }
    case Type::Matrix: return 4 * t.count * m.types[t.elem].count;
    case Type::Vector: return 4 * t.count;
    case Type::Int: case Type::Float: return t.width / 8;
    default: Fail("GetVariableSize: unsupported type");
    }
}

// Size of a member of the VERTEX wrapper's packed LLVM `_Output` struct (PipelineCompiler.cpp:532-547,
// :585-601): LLVM alloc size, i.e. <3 x float> occupies 16 bytes, matrices are structs of column vectors.
uint32_t AllocSize(const Module& m, uint32_t ty) {
    const Type& t = m.types[ty];
    switch (t.kind) {
    case Type::Array: return AllocSize(m, t.elem) * t.count;
    case Type::Matrix: return AllocSize(m, t.elem) * t.count;
    case Type::Vector: return t.count == 3 ? 16 : 4 * t.count;
    case Type::Int: case Type::Float: case Type::Bool: return 4;
    default: Fail("vertex output: unsupported type");
    }
}

// Copies logical words of a value into its LLVM memory image (vec3 padded to 16 bytes).
void StoreAlloc(const Module& m, uint32_t ty, const uint32_t* src, uint8_t* dst) {
    const Type& t = m.types[ty];
    switch (t.kind) {
    case Type::Array: case Type::Matrix: {
        const uint32_t ew = m.types[t.elem].words, es = AllocSize(m, t.elem);
        for (uint32_t k = 0; k < t.count; k++) StoreAlloc(m, t.elem, src + k * ew, dst + k * es);
        break; }
    default: std::memcpy(dst, src, t.words * 4); break;
    }
}

struct InOut {
    uint32_t var;        // variable id
    uint32_t type;       // pointee type
    uint32_t location;
    uint32_t interpolation; // 0 perspective, 1 linear, 2 flat
    uint32_t size;       // bytes (side-specific)
    uint32_t offset;     // bytes within the vertex record (24 + ...)
    uint32_t comps = 0;  // fragment inputs: 32-bit components SetDatum walks (TotalSize / ElementSize of the variable's format)
    bool floats = false; // fragment inputs: 32-bit SFLOAT components (anything else with Perspective / Linear is FATAL_ERROR)
};

struct StageInfo {
    Module mod;
    std::vector<InOut> inputs, outputs;
    uint32_t outputStride = 24;     // VS: sizeof packed _Output
    // builtin plumbing
    uint32_t perVertexVar = 0; int positionMember = -1, pointSizeMember = -1, clipMember = -1; // VS output block
    uint32_t positionVar = 0, pointSizeVar = 0;                                                // loose builtins
    uint32_t vertexIndexVar = 0, instanceIndexVar = 0, fragCoordVar = 0;
};

void Reflect(StageInfo& s, bool vertex) {
    Module& m = s.mod;
    uint32_t inOff = 24, outOff = vertex ? 24 : 0;
    for (const Variable& v : m.variables) {
        const uint32_t pointee = m.types[v.ptrType].elem;
        if (v.storage == ScInput || v.storage == ScOutput) {
            // builtins
            if (m.hasDeco(v.id, DecoBuiltIn)) {
                const uint32_t b = m.decoVal(v.id, DecoBuiltIn);
                if (b == BiVertexIndex || b == BiVertexId) s.vertexIndexVar = v.id;
                else if (b == BiInstanceIndex || b == BiInstanceId) s.instanceIndexVar = v.id;
                else if (b == BiFragCoord) s.fragCoordVar = v.id;
                else if (b == BiPosition) s.positionVar = v.id;
                else if (b == BiPointSize) s.pointSizeVar = v.id;
                continue;
            }
            if (m.types[pointee].kind == Type::Struct) {
                bool isBuiltinBlock = false;
                for (uint32_t k = 0; k < m.types[pointee].members.size(); k++) {
                    uint32_t b;
                    if (m.memberDecoVal(pointee, k, DecoBuiltIn, b)) {
                        isBuiltinBlock = true;
                        if (v.storage == ScOutput) {
                            if (b == BiPosition) s.positionMember = (int)k;
                            else if (b == BiPointSize) s.pointSizeMember = (int)k;
                            else if (b == BiClipDistance) s.clipMember = (int)k;
                        }
                    }
                }
                if (isBuiltinBlock) { if (v.storage == ScOutput) s.perVertexVar = v.id; continue; }
            }
            if (!m.hasDeco(v.id, DecoLocation)) continue;
            InOut io{};
            io.var = v.id; io.type = pointee; io.location = m.decoVal(v.id, DecoLocation);
            if (v.storage == ScInput) {
                io.interpolation = m.hasDeco(v.id, DecoFlat) ? 2 : (m.hasDeco(v.id, DecoNoPerspective) ? 1 : 0);
                io.size = vertex ? 0 : VariableSize(m, pointee);
                io.offset = inOff; inOff += io.size;
                if (!vertex) {
                    const Type& t = m.types[pointee];
                    const Type& et = t.kind == Type::Vector ? m.types[t.elem] : t;
                    io.floats = et.kind == Type::Float && (t.kind == Type::Vector || t.kind == Type::Float);
                    io.comps = t.kind == Type::Vector ? t.count : 1;
                }
                s.inputs.push_back(io);
            } else {
                io.size = vertex ? AllocSize(m, pointee) : 0;
                io.offset = outOff; outOff += io.size;
                s.outputs.push_back(io);
            }
        }
    }
    if (vertex) s.outputStride = outOff;
}

// ---- vertex attribute fetch: EmitCopyInput (PipelineCompiler.cpp:821-896) ----

// GetVariableFormat (Draw.cpp:151-295) for the cases 32-bit shaders can produce.
uint32_t VariableFormat(const Module& m, uint32_t ty) {
    const Type& t = m.types[ty];
    const Type& e = t.kind == Type::Vector ? m.types[t.elem] : t;
    const uint32_t n = t.kind == Type::Vector ? t.count : 1;
    static const uint32_t fl[5] = {0, 100, 103, 106, 109}, si[5] = {0, 99, 102, 105, 108}, ui[5] = {0, 98, 101, 104, 107};
    if (e.kind == Type::Float) return fl[n];
    if (e.kind == Type::Int) return e.isSigned ? si[n] : ui[n];
    return 0;
}

// GetTypeFromFormat (PipelineCompiler.cpp:627-710): formats that are plain int/float vectors.
bool SimpleFormat(uint32_t f, uint32_t& elemBytes, uint32_t& comps, bool& isFloat, bool& isSignedInt) {
    const FormatInfo fi = GetFormatInformation(f);
    if (fi.type != FmtType::Normal) return false;
    if (fi.base != Base::UInt && fi.base != Base::SInt && fi.base != Base::SFloat) return false;
    if (fi.base == Base::SFloat && fi.elementSize == 1) return false;
    if (f >= 30 && f <= 36) return false; // B8G8R8_*: not in the switch
    if (f >= 44 && f <= 50) return false; // B8G8R8A8_*: not in the switch
    elemBytes = fi.elementSize; comps = fi.totalSize / fi.elementSize;
    isFloat = fi.base == Base::SFloat; isSignedInt = fi.base == Base::SInt;
    return true;
}

void FetchAttribute(const Module& m, const CpvkPipelineDesc& desc, const CpvkDrawState& st, uint32_t location,
                    uint32_t vertexId, uint32_t instanceId, uint32_t spirvType, uint32_t* dst) {
    const CpvkVertexAttribute* attr = nullptr;
    for (uint32_t i = 0; i < desc.attributeCount; i++) if (desc.attributes[i].location == location) { attr = &desc.attributes[i]; break; }
    if (!attr) Fail("FindAttribute: no attribute for location " + std::to_string(location));
    const CpvkVertexBinding* bind = nullptr;
    for (uint32_t i = 0; i < desc.bindingCount; i++) if (desc.bindings[i].binding == attr->binding) { bind = &desc.bindings[i]; break; }
    if (!bind) Fail("FindBinding: no binding");
    const uint64_t index = bind->inputRate == 0 ? vertexId : instanceId;
    const uint8_t* src = (const uint8_t*)(uintptr_t)st.vertexBuffers[bind->binding] + (uint64_t)bind->stride * index + attr->offset;
    const Type& t = m.types[spirvType];
    const uint32_t shaderComps = t.kind == Type::Vector ? t.count : 1;
    const Type& et = t.kind == Type::Vector ? m.types[t.elem] : t;
    const uint32_t shaderFormat = VariableFormat(m, spirvType);
    if (shaderFormat == attr->format) { std::memcpy(dst, src, shaderComps * 4); return; }
    uint32_t eb, comps; bool isF, isS;
    if (SimpleFormat(attr->format, eb, comps, isF, isS)) {
        if (comps != shaderComps) Fail("EmitVectorConversion: component count mismatch");
        if (isF != (et.kind == Type::Float)) Fail("EmitConversion: float<->int attribute conversion (TODO_ERROR)");
        for (uint32_t c = 0; c < comps; c++) {
            if (isF) {
                float v;
                if (eb == 2) { uint16_t h; std::memcpy(&h, src + 2 * c, 2); v = HalfToFloat(h); }
                else if (eb == 4) std::memcpy(&v, src + 4 * c, 4);
                else { double d; std::memcpy(&d, src + 8 * c, 8); v = (float)d; }
                std::memcpy(&dst[c], &v, 4);
            } else {
                uint64_t raw = 0; std::memcpy(&raw, src + eb * c, eb);
                if (et.isSigned && eb < 4) { // CreateSExtOrTrunc by *target* signedness
                    const int shift = 64 - 8 * (int)eb; raw = (uint64_t)(((int64_t)(raw << shift)) >> shift);
                }
                dst[c] = (uint32_t)raw;
            }
        }
        return;
    }
    // packed / normalised formats: EmitGetPixel to a 4-vector of the shader's element type, keep the first N
    const FormatInfo fi = GetFormatInformation(attr->format);
    if (fi.type == FmtType::Invalid) Fail("vertex attribute format unsupported");
    if (et.kind == Type::Float) { float px[4]; GetPixelF32(fi, attr->format, src, px); std::memcpy(dst, px, shaderComps * 4); }
    else { uint32_t px[4]; GetPixelInt(fi, src, px); std::memcpy(dst, px, shaderComps * 4); }
}

// ---- fragment epilogue pieces ----

bool FCompare(float reference, float value, uint32_t op) { // CompileFCompareTest: ordered compares
    switch (op) {
    case 0: return false; case 1: return reference < value; case 2: return reference == value; case 3: return reference <= value;
    case 4: return reference > value; case 5: return reference < value || reference > value; case 6: return reference >= value;
    case 7: return true; default: Fail("bad compare op");
    }
}
bool ICompare(uint8_t reference, uint8_t value, uint32_t op) {
    switch (op) {
    case 0: return false; case 1: return reference < value; case 2: return reference == value; case 3: return reference <= value;
    case 4: return reference > value; case 5: return reference != value; case 6: return reference >= value;
    case 7: return true; default: Fail("bad compare op");
    }
}
uint8_t StencilResult(uint32_t op, uint8_t cur, uint8_t ref) { // CompileGetStencilResult :1382-1413
    switch (op) {
    case 0: return cur; case 1: return 0; case 2: return ref;
    case 3: { int v = (int8_t)cur + 1; if (v > 127) v = 127; return (uint8_t)(int8_t)v; }   // sadd_sat on i8 (reference quirk)
    case 4: { int v = (int8_t)cur - 1; if (v < -128) v = -128; return (uint8_t)(int8_t)v; } // ssub_sat on i8
    case 5: return (uint8_t)~cur; case 6: return (uint8_t)(cur + 1); case 7: return (uint8_t)(cur - 1);
    default: Fail("bad stencil op");
    }
}

struct F4 { float v[4]; };

F4 BlendFactor(const F4& s, const F4& d, const F4& c, uint32_t colourFactor, uint32_t alphaFactor) { // ApplyBlendFactor
    F4 v{{0, 0, 0, 0}};
    auto splat = [&](float x) { v = F4{{x, x, x, x}}; };
    switch (colourFactor) {
    case 0: break;
    case 1: splat(1); break;
    case 2: v = s; break;
    case 3: for (int i = 0; i < 4; i++) v.v[i] = 1.0f - s.v[i]; break;
    case 4: v = d; break;
    case 5: for (int i = 0; i < 4; i++) v.v[i] = 1.0f - d.v[i]; break;
    case 6: splat(s.v[3]); break;
    case 7: splat(1 - s.v[3]); break;
    case 8: splat(d.v[3]); break;
    case 9: splat(1 - d.v[3]); break;
    case 10: v = c; break;
    case 11: for (int i = 0; i < 4; i++) v.v[i] = 1.0f - c.v[i]; break;
    case 12: splat(c.v[3]); break;
    case 13: splat(1 - c.v[3]); break;
    case 14: { const float f = std::min(s.v[3], 1 - d.v[3]); v = F4{{f, f, f, 1}}; break; }
    default: Fail("blend factor unsupported (SRC1_*: TODO_ERROR)");
    }
    if (colourFactor != alphaFactor) {
        switch (alphaFactor) {
        case 0: v.v[3] = 0; break;
        case 1: case 14: v.v[3] = 1; break;
        case 2: case 6: v.v[3] = s.v[3]; break;
        case 3: case 7: v.v[3] = 1 - s.v[3]; break;
        case 4: case 8: v.v[3] = d.v[3]; break;
        case 5: case 9: v.v[3] = 1 - d.v[3]; break;
        case 10: case 12: v.v[3] = c.v[3]; break;
        case 11: case 13: v.v[3] = 1 - c.v[3]; break;
        default: Fail("blend factor unsupported");
        }
    }
    return v;
}

F4 ApplyBlend(const F4& s, const F4& d, const F4& c, const CpvkBlendAttachment& b) { // Draw.cpp:1105-1262
    const F4 sf = BlendFactor(s, d, c, b.srcColorBlendFactor, b.srcAlphaBlendFactor);
    const F4 df = BlendFactor(s, d, c, b.dstColorBlendFactor, b.dstAlphaBlendFactor);
    F4 v{};
    for (int i = 0; i < 4; i++) {
        switch (b.colorBlendOp) {
        case 0: v.v[i] = s.v[i] * sf.v[i] + d.v[i] * df.v[i]; break;
        case 1: v.v[i] = s.v[i] * sf.v[i] - d.v[i] * df.v[i]; break;
        case 2: v.v[i] = d.v[i] * df.v[i] - s.v[i] * sf.v[i]; break;
        case 3: v.v[i] = (d.v[i] < s.v[i]) ? d.v[i] : s.v[i]; break; // glm::min(x,y) = (y < x) ? y : x
        case 4: v.v[i] = (s.v[i] < d.v[i]) ? d.v[i] : s.v[i]; break; // glm::max(x,y) = (x < y) ? y : x
        default: Fail("blend op unsupported (TODO_ERROR)");
        }
    }
    if (b.colorBlendOp != b.alphaBlendOp) {
        switch (b.alphaBlendOp) {
        case 0: v.v[3] = s.v[3] * sf.v[3] + d.v[3] * df.v[3]; break;
        case 1: v.v[3] = s.v[3] * sf.v[3] - d.v[3] * df.v[3]; break;
        case 2: v.v[3] = d.v[3] * df.v[3] - s.v[3] * sf.v[3]; break;
        case 3: v.v[3] = std::min(s.v[3], d.v[3]); break;
        case 4: v.v[3] = std::max(s.v[3], d.v[3]); break;
        default: Fail("blend op unsupported (TODO_ERROR)");
        }
    }
    return v;
}

inline float EdgeFunction(const float a[4], const float b[4], const float c[2]) { // Draw.cpp:415-418
    return (c[0] - a[0]) * (b[1] - a[1]) - (c[1] - a[1]) * (b[0] - a[0]);
}

struct Window { int32_t x0, y0, x1, y1; };

struct DrawContext {
    const CpvkPipelineDesc* desc;
    const CpvkDrawState* st;
    StageInfo vs, fs;
    bool hasFs = false;
    Interp vsi, fsi;
    std::vector<uint8_t> vertexStorage;
    Window win;
    CpvkDrawStats stats{};
    // fragment-stream mode (cpvk_oracle_raster_records): no shaders; every fragment the rasteriser emits is written down
    // with the arguments the reference would call the fragment wrapper with and the interpolated inputs
    std::vector<uint32_t>* trace = nullptr;
    bool traceOriginUpper = true;
};

void CheckSupported(const CpvkPipelineDesc& d) {
    // Everything the reference aborts on is rejected here as well.
    if (d.rasterizerDiscardEnable) Fail("RasterizerDiscardEnable (TODO_ERROR, Draw.cpp:1602-1605)");
    if (d.polygonMode != 0) Fail("PolygonMode != FILL (TODO_ERROR, Draw.cpp:1675-1678)");
    if (d.primitiveRestartEnable) Fail("primitive restart (TODO_ERROR, Draw.cpp:697-700)");
    if (d.logicOpEnable) Fail("logic op (TODO_ERROR, PipelineCompiler.cpp:1687-1690)");
    if (d.depthClampEnable && d.depthTestEnable) Fail("depth clamp (TODO_ERROR, PipelineCompiler.cpp:1232-1236)");
    if (d.rasterizationSamples > 1) Fail("multisampling");
}

// IA: ProcessInputAssembler (Draw.cpp:675-688) / ProcessIndexedVertices + ProcessInputAssemblerIndexed (:690-760): the vertex id the
// vertex stage sees for raw vertex i. Pinned against those functions themselves (tests/test_reference_draw.py, draw_check ia).
uint32_t AssembledVertexId(const CpvkDrawState& st, uint32_t i) {
    if (st.indexStride == 0) return st.first + i;
    const uint8_t* ib = (const uint8_t*)(uintptr_t)st.indexBuffer;
    uint32_t index = 0;
    const uint64_t k = (uint64_t)st.first + i;
    if (st.indexStride == 1) index = ib[k];
    else if (st.indexStride == 2) { uint16_t t; std::memcpy(&t, ib + 2 * k, 2); index = t; }
    else { std::memcpy(&index, ib + 4 * k, 4); }
    return (uint32_t)st.vertexOffset + index;
}

void RunVertexStage(DrawContext& c, uint32_t instance, uint32_t n) {
    const CpvkDrawState& st = *c.st;
    Module& m = c.vs.mod;
    c.vertexStorage.assign((size_t)n * c.vs.outputStride, 0);
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t vertexId = AssembledVertexId(st, i);
        c.vsi.BeginInvocation();
        if (c.vs.vertexIndexVar) *c.vsi.VarData(c.vs.vertexIndexVar) = vertexId;
        if (c.vs.instanceIndexVar) *c.vsi.VarData(c.vs.instanceIndexVar) = instance;
        for (const InOut& in : c.vs.inputs) {
            const Type& t = m.types[in.type];
            uint32_t* dst = c.vsi.VarData(in.var);
            if (t.kind == Type::Array || t.kind == Type::Matrix) { // consecutive locations (x2 if element > 16 bytes)
                const uint32_t ew = m.types[t.elem].words;
                const uint32_t mult = VariableSize(m, t.elem) > 16 ? 2 : 1;
                for (uint32_t j = 0; j < t.count; j++) FetchAttribute(m, *c.desc, st, in.location + j * mult, vertexId, instance, t.elem, dst + j * ew);
            } else FetchAttribute(m, *c.desc, st, in.location, vertexId, instance, in.type, dst);
        }
        c.vsi.Call(m.entryPoint, nullptr, nullptr, 0);
        // record store: {vec4 position, float pointSize, float clip[1]} + outputs in declaration order
        uint8_t* rec = c.vertexStorage.data() + (size_t)i * c.vs.outputStride;
        if (c.vs.perVertexVar) {
            const uint32_t* pv = c.vsi.VarData(c.vs.perVertexVar);
            const Type& bt = m.types[m.types[m.variables[m.varIndex.at(c.vs.perVertexVar)].ptrType].elem];
            uint32_t off = 0;
            for (uint32_t k = 0; k < bt.members.size(); k++) {
                if ((int)k == c.vs.positionMember) std::memcpy(rec, pv + off, 16);
                if ((int)k == c.vs.pointSizeMember) std::memcpy(rec + 16, pv + off, 4);
                if ((int)k == c.vs.clipMember && m.types[bt.members[k]].words) std::memcpy(rec + 20, pv + off, 4);
                off += m.types[bt.members[k]].words;
            }
        }
        if (c.vs.positionVar) std::memcpy(rec, c.vsi.VarData(c.vs.positionVar), 16);
        if (c.vs.pointSizeVar) std::memcpy(rec + 16, c.vsi.VarData(c.vs.pointSizeVar), 4);
        for (const InOut& out : c.vs.outputs) StoreAlloc(m, out.type, c.vsi.VarData(out.var), rec + out.offset);
    }
}

// One fragment: FS call + epilogue (PipelineCompiler.cpp:1058-1080). Returns nothing; updates attachments.
// nv = vertices of the primitive: 3 triangle, 2 line (SetDatum<.., 2>, Draw.cpp:1458-1482), 1 point (every input is a
// plain copy of the vertex's value whatever its interpolation qualifier, Draw.cpp:1366-1370).
void Fragment(DrawContext& c, int32_t x, int32_t y, float depth, bool front, const float w[3], const float pw[3],
              const uint32_t idx[3], uint32_t provoking, int nv = 3) {
    const CpvkPipelineDesc& d = *c.desc;
    const CpvkDrawState& st = *c.st;
    Module& m = c.fs.mod;
    c.stats.fragmentsCovered++;
    size_t traceAt = 0;
    if (c.trace) { traceAt = c.trace->size(); uint32_t words = 8; for (const InOut& in : c.fs.inputs) words += in.size / 4; c.trace->resize(traceAt + words, 0); }
    else c.fsi.BeginInvocation();
    uint32_t traceWord = 8;
    // interpolants (Draw.cpp:816-872, 911-951)
    for (const InOut& in : c.fs.inputs) {
        uint32_t* dst = c.trace ? c.trace->data() + traceAt + traceWord : c.fsi.VarData(in.var);
        traceWord += in.size / 4;
        const uint8_t* data[3];
        for (int k = 0; k < nv; k++) data[k] = c.vertexStorage.data() + (size_t)idx[k] * c.vs.outputStride + in.offset;
        if (nv == 1) { std::memcpy(dst, data[0], in.size); continue; }
        if (in.interpolation == 2) { std::memcpy(dst, c.vertexStorage.data() + (size_t)provoking * c.vs.outputStride + in.offset, in.size); continue; }
        if (!in.floats) Fail("SetDatum: only 32-bit float inputs interpolate (FATAL_ERROR, Draw.cpp:863-869)");
        for (uint32_t e = 0; e < in.comps; e++) {
            float values[3]; for (int k = 0; k < nv; k++) std::memcpy(&values[k], data[k] + 4 * e, 4);
            float result;
            if (in.interpolation == 0) {
                float numerator = 0.0f, denominator = 0.0f;
                for (int k = 0; k < nv; k++) { numerator += w[k] * values[k] / pw[k]; denominator += w[k] / pw[k]; }
                result = numerator / denominator;
            } else {
                result = 0.0f; for (int k = 0; k < nv; k++) result += w[k] * values[k];
            }
            std::memcpy(&dst[e], &result, 4);
        }
    }
    if (c.fs.fragCoordVar || c.trace) {
        float fc[4];
        fc[0] = (c.trace ? c.traceOriginUpper : m.originUpperLeft) ? (float)x : st.viewport.width - (float)x - 1; // Draw.cpp:1579
        fc[1] = (float)y; fc[2] = depth; fc[3] = 1.0f;
        std::memcpy(c.trace ? c.trace->data() + traceAt + 4 : c.fsi.VarData(c.fs.fragCoordVar), fc, 16);
    }
    if (c.trace) {
        const float shaderDepth = (st.viewport.maxDepth - st.viewport.minDepth) * depth + st.viewport.minDepth; // Draw.cpp:1310
        uint32_t* t = c.trace->data() + traceAt;
        t[0] = (uint32_t)x; t[1] = (uint32_t)y; t[2] = front ? 1u : 0u; std::memcpy(&t[3], &shaderDepth, 4);
        return;
    }
    depth = (st.viewport.maxDepth - st.viewport.minDepth) * depth + st.viewport.minDepth; // Draw.cpp:1310
    c.fsi.Call(m.entryPoint, nullptr, nullptr, 0);
    if (c.fsi.killed) return;

    const bool hasDS = st.depthStencil.address != 0 && d.depthStencilFormat != 0;
    const FormatInfo dsf = GetFormatInformation(d.depthStencilFormat);
    uint8_t* dsPtr = hasDS ? (uint8_t*)(uintptr_t)st.depthStencil.address + (uint64_t)y * st.depthStencil.rowPitch + (uint64_t)x * dsf.totalSize : nullptr;
    const bool fmtDepth = d.depthStencilFormat != 0 && dsf.depthOffset != INVALID_OFFSET;
    const bool fmtStencil = d.depthStencilFormat != 0 && dsf.stencilOffset != INVALID_OFFSET;
    // CompileGetCurrentData
    float currentDepth = 0; uint8_t currentStencil = 0;
    if ((d.depthBoundsTestEnable || d.depthTestEnable) && fmtDepth && hasDS) currentDepth = GetDepth(d.depthStencilFormat, dsPtr);
    if (d.stencilTestEnable && fmtStencil && hasDS) currentStencil = GetStencil(dsf, dsPtr);
    // CompileDepthBoundsTest: unordered compares
    if (d.depthBoundsTestEnable && d.depthStencilFormat != 0) {
        if (!(currentDepth >= d.minDepthBounds) || !(currentDepth <= d.maxDepthBounds)) return;
    }
    bool stencilResult = true, depthResult = true;
    const bool stencilOn = d.stencilTestEnable && fmtStencil;
    const CpvkStencilOpState& testState = front ? d.front : d.back;
    uint8_t stencilRef = 0;
    if (stencilOn) {
        stencilRef = (uint8_t)testState.reference;
        stencilResult = ICompare(stencilRef & (uint8_t)testState.compareMask, currentStencil & (uint8_t)testState.compareMask, testState.compareOp);
    }
    if (d.depthTestEnable && d.depthStencilFormat != 0 && d.depthStencilFormat != F_S8_UINT) depthResult = FCompare(depth, currentDepth, d.depthCompareOp);
    // CompileDepthStencilWrite: both arms pass front=true, so the WRITE ops always come from `front` state
    // (reference defect); kept literally.
    if (stencilOn) {
        const CpvkStencilOpState& ws = d.front;
        const uint8_t failR = StencilResult(ws.failOp, currentStencil, stencilRef);
        const uint8_t dfailR = StencilResult(ws.depthFailOp, currentStencil, stencilRef);
        const uint8_t passR = StencilResult(ws.passOp, currentStencil, stencilRef);
        uint8_t writeValue = stencilResult ? (depthResult ? passR : dfailR) : failR;
        writeValue = (uint8_t)((writeValue & (uint8_t)ws.writeMask) | (currentStencil & (uint8_t)~ws.writeMask));
        const bool attempt = d.depthTestEnable && d.depthWriteEnable && d.depthStencilFormat != F_S8_UINT;
        if (hasDS) {
            if (depthResult && attempt) SetDepthStencil(dsf, d.depthStencilFormat, dsPtr, depth, writeValue);
            else { // SetStencilPixelXXX: re-reads and re-packs depth (GlslFunctions.cpp:898-914)
                const float dd = fmtDepth ? GetDepth(d.depthStencilFormat, dsPtr) : 0.0f;
                SetDepthStencil(dsf, d.depthStencilFormat, dsPtr, dd, writeValue);
            }
        }
    } else if (d.depthTestEnable && d.depthWriteEnable && d.depthStencilFormat != 0 && d.depthStencilFormat != F_S8_UINT) {
        if (depthResult && hasDS) { // SetDepthPixelXXX preserves the stencil byte (GlslFunctions.cpp:880-896)
            const uint8_t s = fmtStencil ? GetStencil(dsf, dsPtr) : 0;
            SetDepthStencil(dsf, d.depthStencilFormat, dsPtr, depth, s);
        }
    }
    if (!(stencilResult && depthResult)) return;
    c.stats.fragmentsWritten++;
    // CompileWriteFragment
    for (uint32_t a = 0; a < d.colorAttachmentCount; a++) {
        if (d.colorFormats[a] == 0 || st.color[a].address == 0) continue;
        const InOut* out = nullptr;
        for (const InOut& o : c.fs.outputs) {
            const Type& ot = m.types[o.type];
            const uint32_t span = (ot.kind == Type::Array || ot.kind == Type::Matrix) ? ot.count : 1;
            if (a >= o.location && a < o.location + span) { out = &o; break; }
        }
        if (!out) Fail("outputLocations.at(): fragment shader has no output for attachment " + std::to_string(a));
        const Type& ot = m.types[out->type];
        const uint32_t* src = c.fsi.VarData(out->var);
        if (ot.kind == Type::Array || ot.kind == Type::Matrix) src += (a - out->location) * m.types[ot.elem].words;
        uint32_t value[4] = {0, 0, 0, 0};
        const uint32_t avail = (ot.kind == Type::Array || ot.kind == Type::Matrix) ? m.types[ot.elem].words : ot.words;
        std::memcpy(value, src, std::min<uint32_t>(avail, 4) * 4);
        const FormatInfo cf = GetFormatInformation(d.colorFormats[a]);
        uint8_t* px = (uint8_t*)(uintptr_t)st.color[a].address + (uint64_t)y * st.color[a].rowPitch + (uint64_t)x * cf.totalSize;
        const CpvkBlendAttachment& b = d.blend[a];
        if (cf.base == Base::UInt || cf.base == Base::SInt) {
            if (b.colorWriteMask != 0xF) { uint32_t dst[4]; GetPixelInt(cf, px, dst); for (int k = 0; k < 4; k++) if (!(b.colorWriteMask & (1u << k))) value[k] = dst[k]; }
            SetPixelInt(cf, px, value);
        } else {
            F4 colour; std::memcpy(colour.v, value, 16);
            if (b.blendEnable || b.colorWriteMask != 0xF) {
                F4 dst; GetPixelF32(cf, d.colorFormats[a], px, dst.v); // ImageFetch (Draw.cpp:1283-1298)
                if (b.blendEnable) { F4 cst; std::memcpy(cst.v, d.blendConstants, 16); colour = ApplyBlend(colour, dst, cst, b); }
                for (int k = 0; k < 4; k++) if (!(b.colorWriteMask & (1u << k))) colour.v[k] = dst.v[k]; // PipelineCompiler.cpp:1695-1698
            }
            SetPixelF32(cf, px, colour.v);
        }
    }
}

// EdgeFunction on vec2 (Draw.cpp:410-413)
static float EdgeFunction2(const float a[2], const float b[2], const float c[2]) { return (c[0] - a[0]) * (b[1] - a[1]) - (c[1] - a[1]) * (b[0] - a[0]); }

// ProcessPoints (Draw.cpp:1315-1380): a square of ceil(pointSize / 2) pixels around the truncated screen position,
// s/t test against the point size, inputs copied from the vertex, depth = p0.z, always front-facing.
void ProcessPoints(DrawContext& c, uint32_t primCount) {
    const CpvkDrawState& st = *c.st;
    const float W = st.viewport.width, H = st.viewport.height;
    for (uint32_t p = 0; p < primCount; p++) {
        float pos[4], pointSize;
        std::memcpy(pos, c.vertexStorage.data() + (size_t)p * c.vs.outputStride, 16);
        std::memcpy(&pointSize, c.vertexStorage.data() + (size_t)p * c.vs.outputStride + 16, 4);
        float P[4]; for (int q = 0; q < 4; q++) P[q] = pos[q] / pos[3];
        P[3] = pos[3];
        const int32_t px = (int32_t)((P[0] + 1) * 0.5f * (W - 1)), py = (int32_t)((P[1] + 1) * 0.5f * (H - 1));
        const int32_t half = (int32_t)std::ceil(pointSize / 2);
        int32_t startX = std::max(0, px - half), startY = std::max(0, py - half);
        int32_t endX = std::min((int32_t)W, px + half + 1), endY = std::min((int32_t)H, py + half + 1);
        startX = std::max(startX, c.win.x0); startY = std::max(startY, c.win.y0);
        endX = std::min(endX, c.win.x1); endY = std::min(endY, c.win.y1);
        const uint32_t idx[3] = {p, p, p}; const float w[3] = {1.0f, 0.0f, 0.0f}; const float pw[3] = {P[3], 1.0f, 1.0f};
        for (int32_t y = startY; y < endY; y++)
            for (int32_t x = startX; x < endX; x++) {
                const float s = 0.5f + (x - px) / pointSize;
                const float t = 0.5f + (y - py) / pointSize;
                if (s >= 0 && t >= 0 && s <= 1 && t <= 1) Fragment(c, x, y, P[2], true, w, pw, idx, p, 1);
            }
    }
}

// ProcessLines (Draw.cpp:1382-1508), rectangular mode: a quad of +-perpendicular * (lineWidth / viewport) around the
// segment in NDC; the perpendicular comes from glm::normalize of the *4-component* difference (so z and w shorten it);
// every pixel of the viewport is tested against the four quad edges; t by projection on the segment;
// attributes weigh (1 - t, t), depth weighs (t, 1 - t) — as written there.
void ProcessLines(DrawContext& c, uint32_t primCount, uint32_t topology) {
    const CpvkDrawState& st = *c.st;
    const CpvkPipelineDesc& d = *c.desc;
    const float W = st.viewport.width, H = st.viewport.height;
    const float halfPixel[2] = {(1.0f / W) * 0.5f, (1.0f / H) * 0.5f};
    for (uint32_t p = 0; p < primCount; p++) {
        const uint32_t i0 = topology == 1 ? p * 2 : p, i1 = i0 + 1, provoking = i0;
        float P[2][4];
        const uint32_t iv[2] = {i0, i1};
        for (int k = 0; k < 2; k++) {
            float pos[4]; std::memcpy(pos, c.vertexStorage.data() + (size_t)iv[k] * c.vs.outputStride, 16);
            for (int q = 0; q < 4; q++) P[k][q] = pos[q] / pos[3];
            P[k][3] = pos[3];
        }
        const float lineWidth[2] = {d.lineWidth / W, d.lineWidth / H};
        float diff[4]; for (int q = 0; q < 4; q++) diff[q] = P[1][q] - P[0][q];
        const float sqr = diff[0] * diff[0] + diff[1] * diff[1] + diff[2] * diff[2] + diff[3] * diff[3]; // glm normalize(vec4), func_geometric.inl:269-278
        const float inv = 1.0f / std::sqrt(sqr);                                                          // glm inversesqrt(float), func_exponential.inl:226-229
        const float dir[2] = {diff[0] * inv, diff[1] * inv};
        const float perp[2] = {dir[1], -dir[0]};
        const float off[2] = {perp[0] * lineWidth[0], perp[1] * lineWidth[1]};
        const float p00[2] = {P[0][0] + off[0], P[0][1] + off[1]}, p01[2] = {P[0][0] - off[0], P[0][1] - off[1]};
        const float p10[2] = {P[1][0] + off[0], P[1][1] + off[1]}, p11[2] = {P[1][0] - off[0], P[1][1] - off[1]};
        const float seg[2] = {P[1][0] - P[0][0], P[1][1] - P[0][1]};
        const float len = std::sqrt(seg[0] * seg[0] + seg[1] * seg[1]); // glm length(vec2)
        const int32_t y1 = std::min((int32_t)std::ceil(H), c.win.y1), x1 = std::min((int32_t)std::ceil(W), c.win.x1); // `y < viewport.height` with y unsigned
        const uint32_t idx[3] = {i0, i1, i1}; const float pw[3] = {P[0][3], P[1][3], 1.0f};
        for (int32_t y = std::max(0, c.win.y0); y < y1; y++) {
            const float yf = ((float)y / H + halfPixel[1]) * 2 - 1;
            for (int32_t x = std::max(0, c.win.x0); x < x1; x++) {
                const float xf = ((float)x / W + halfPixel[0]) * 2 - 1;
                const float pt[2] = {xf, yf};
                if (!(EdgeFunction2(p00, p01, pt) >= 0 && EdgeFunction2(p11, p10, pt) >= 0 && EdgeFunction2(p10, p00, pt) >= 0 && EdgeFunction2(p01, p11, pt) >= 0)) continue;
                const float rel[2] = {pt[0] - P[0][0], pt[1] - P[0][1]};
                const float t = (rel[0] * seg[0] + rel[1] * seg[1]) / (len * len); // glm dot(vec2) = tmp.x + tmp.y
                const float w[3] = {1 - t, t, 0.0f};
                const float depth = P[0][2] * t + P[1][2] * (1 - t);
                Fragment(c, x, y, depth, true, w, pw, idx, provoking, 2);
            }
        }
    }
}

void ProcessTriangles(DrawContext& c, uint32_t primCount, uint32_t topology) {
    const CpvkDrawState& st = *c.st;
    const CpvkPipelineDesc& d = *c.desc;
    const float W = st.viewport.width, H = st.viewport.height;
    const float halfPixel[2] = {(1.0f / W) * 0.5f, (1.0f / H) * 0.5f};
    for (uint32_t p = 0; p < primCount; p++) {
        uint32_t provoking, v0, v1, v2;
        if (topology == 3) { provoking = p * 3; v0 = p * 3; v1 = p * 3 + 1; v2 = p * 3 + 2; }         // TRIANGLE_LIST
        else if (topology == 4) { provoking = p; v0 = p; v1 = p + 1; v2 = p + 2; }                     // TRIANGLE_STRIP
        else { provoking = p + 1; v0 = 0; v1 = p + 1; v2 = p + 2; }                                    // TRIANGLE_FAN
        uint32_t idx[3] = {v0, v1, v2};
        if (d.frontFace == 1) std::swap(idx[0], idx[2]); // VK_FRONT_FACE_CLOCKWISE
        float P[3][4];
        for (int k = 0; k < 3; k++) {
            float pos[4]; std::memcpy(pos, c.vertexStorage.data() + (size_t)idx[k] * c.vs.outputStride, 16);
            for (int q = 0; q < 4; q++) P[k][q] = pos[q] / pos[3];
            P[k][3] = pos[3];
        }
        int32_t sx[3], sy[3];
        for (int k = 0; k < 3; k++) { sx[k] = (int32_t)((P[k][0] + 1) * 0.5f * W); sy[k] = (int32_t)((P[k][1] + 1) * 0.5f * H); }
        int32_t startX = std::max(0, std::min({sx[0], sx[1], sx[2]}));
        int32_t startY = std::max(0, std::min({sy[0], sy[1], sy[2]}));
        int32_t endX = std::min((int32_t)W, std::max({sx[0], sx[1], sx[2]}) + 1);
        int32_t endY = std::min((int32_t)H, std::max({sy[0], sy[1], sy[2]}) + 1);
        // window restriction (oracle-only; pixels are independent in the reference)
        startX = std::max(startX, c.win.x0); startY = std::max(startY, c.win.y0);
        endX = std::min(endX, c.win.x1); endY = std::min(endY, c.win.y1);
        const float pw[3] = {P[0][3], P[1][3], P[2][3]};
        for (int32_t y = startY; y < endY; y++) {
            const float yf = ((float)y / H + halfPixel[1]) * 2 - 1;
            for (int32_t x = startX; x < endX; x++) {
                const float xf = ((float)x / W + halfPixel[0]) * 2 - 1;
                const float pt[2] = {xf, yf};
                // GetFragmentInput
                float area = EdgeFunction(P[0], P[1], P[2]);
                float w0, w1, w2; bool front;
                if (area < 0) { area = -area; front = false; w0 = EdgeFunction(P[2], P[1], pt); w1 = EdgeFunction(P[0], P[2], pt); w2 = EdgeFunction(P[1], P[0], pt); }
                else { front = true; w0 = EdgeFunction(P[1], P[2], pt); w1 = EdgeFunction(P[2], P[0], pt); w2 = EdgeFunction(P[0], P[1], pt); }
                if (w0 < 0 || w1 < 0 || w2 < 0 || ((d.cullMode & 2) && !front) || ((d.cullMode & 1) && front)) continue;
                w0 /= area; w1 /= area; w2 /= area;
                const float depth = P[0][2] * w0 + P[1][2] * w1 + P[2][2] * w2;
                const float w[3] = {w0, w1, w2};
                Fragment(c, x, y, depth, front, w, pw, idx, provoking);
            }
        }
    }
}

int DrawImpl(const CpvkPipelineDesc* desc, const CpvkDrawState* st, Window win, CpvkDrawStats* stats) {
    try {
        CheckSupported(*desc);
        DrawContext c;
        c.desc = desc; c.st = st; c.win = win;
        Parse(c.vs.mod, desc->vertex.spirv, desc->vertex.wordCount, desc->vertex.entryPoint ? desc->vertex.entryPoint : "main", 0,
              desc->vertex.spec, desc->vertex.specCount);
        Reflect(c.vs, true);
        c.hasFs = desc->fragment.spirv != nullptr;
        if (c.hasFs) {
            Parse(c.fs.mod, desc->fragment.spirv, desc->fragment.wordCount, desc->fragment.entryPoint ? desc->fragment.entryPoint : "main", 4,
                  desc->fragment.spec, desc->fragment.specCount);
            Reflect(c.fs, false);
        }
        Env env; env.descriptors = st->descriptors; env.descriptorCount = st->descriptorCount; env.pushConstants = st->pushConstants;
        c.vsi.Bind(&c.vs.mod, env);
        if (c.hasFs) c.fsi.Bind(&c.fs.mod, env);
        if (st->viewport.width <= 0 || st->viewport.height <= 0) Fail("viewport must be positive");
        const uint32_t n = st->count;
        uint32_t primCount = 0;
        switch (desc->topology) {
        case 3: primCount = n / 3; break;
        case 4: case 5: primCount = n > 2 ? n - 2 : 0; break;
        case 0: primCount = n; break;
        case 1: primCount = n / 2; break;
        case 2: primCount = n > 1 ? n - 1 : 0; break;
        default: Fail("topology unsupported (TODO_ERROR, Draw.cpp:663-668)");
        }
        for (uint32_t i = 0; i < st->instanceCount; i++) {
            RunVertexStage(c, st->firstInstance + i, n);
            // vkCmdDraw skips raster without a fragment stage (Draw.cpp:1799-1802)
            if (c.hasFs) {
                if (desc->topology == 0) ProcessPoints(c, primCount);
                else if (desc->topology <= 2) ProcessLines(c, primCount, desc->topology);
                else ProcessTriangles(c, primCount, desc->topology);
            }
            c.stats.primitives += primCount;
        }
        if (stats) *stats = c.stats;
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return CPVK_E_UNSUPPORTED;
    }
}

} // namespace

extern "C" {

const char* cpvk_oracle_last_error(void) { return g_error.c_str(); }

int cpvk_oracle_draw(const CpvkPipelineDesc* desc, const CpvkDrawState* state, CpvkDrawStats* stats) {
    Window w{0, 0, INT32_MAX, INT32_MAX};
    if (state->bandY1 > state->bandY0) { w.y0 = (int32_t)state->bandY0; w.y1 = (int32_t)state->bandY1; }
    return DrawImpl(desc, state, w, stats);
}

int cpvk_oracle_draw_window(const CpvkPipelineDesc* desc, const CpvkDrawState* state, int32_t x0, int32_t y0, int32_t x1, int32_t y1,
                            CpvkDrawStats* stats) {
    return DrawImpl(desc, state, Window{x0, y0, x1, y1}, stats);
}

// Fragment-stream mode, the counterpart of oracle/ref_draw_check.cpp `raster`: the vertex stage's output records are given
// (no shaders run), CalculatePrimitives + ProcessPoints / ProcessLines / ProcessTriangles + GetFragmentInput / SetDatum /
// DrawPixel of THIS file produce, per emitted fragment and in emission order, the words
//   x, y, front, depth as the fragment wrapper receives it (after the viewport transform), fragCoord[4], interpolated inputs.
// inputs = {offset, VkFormat, interpolation (0 perspective, 1 linear, 2 flat), size} per fragment input, as
// VariableInOutData has them (Draw.cpp:97-106). Returns the number of fragments, or a negative error; `out` receives at
// most `capacityWords` words (the count is still exact, so a caller can size the buffer and call again).
int64_t cpvk_oracle_raster_records(float width, float height, float minDepth, float maxDepth, float lineWidth, uint32_t topology,
                                   uint32_t frontFace, uint32_t cullMode, uint32_t originUpper, uint32_t vertexCount, uint32_t stride,
                                   const uint32_t* inputs, uint32_t inputCount, const uint8_t* records, uint32_t* out, uint64_t capacityWords) {
    try {
        CpvkPipelineDesc desc{};
        desc.topology = topology; desc.frontFace = frontFace; desc.cullMode = cullMode; desc.lineWidth = lineWidth;
        CpvkDrawState st{};
        st.viewport.width = width; st.viewport.height = height; st.viewport.minDepth = minDepth; st.viewport.maxDepth = maxDepth;
        DrawContext c;
        c.desc = &desc; c.st = &st; c.win = Window{0, 0, INT32_MAX, INT32_MAX};
        std::vector<uint32_t> trace;
        c.trace = &trace; c.traceOriginUpper = originUpper != 0;
        c.vs.outputStride = stride;
        c.vertexStorage.assign(records, records + (size_t)vertexCount * stride);
        uint32_t words = 8;
        for (uint32_t i = 0; i < inputCount; i++) {
            InOut io{};
            io.location = i; io.offset = inputs[4 * i]; io.interpolation = inputs[4 * i + 2]; io.size = inputs[4 * i + 3];
            const FormatInfo fi = GetFormatInformation(inputs[4 * i + 1]);
            if (fi.type == FmtType::Invalid) Fail("raster_records: unknown input format");
            io.comps = fi.totalSize / fi.elementSize;                      // SetDatum: TotalSize / ElementSize (Draw.cpp:841-842)
            io.floats = fi.base == Base::SFloat && fi.elementSize == 4;   // :845-869
            c.fs.inputs.push_back(io);
            words += io.size / 4;
        }
        const uint32_t n = vertexCount;
        uint32_t primCount = 0;
        switch (topology) {
        case 3: primCount = n / 3; break;
        case 4: case 5: primCount = n > 2 ? n - 2 : 0; break;
        case 0: primCount = n; break;
        case 1: primCount = n / 2; break;
        case 2: primCount = n > 1 ? n - 1 : 0; break;
        default: Fail("topology unsupported (TODO_ERROR, Draw.cpp:663-668)");
        }
        if (topology == 0) ProcessPoints(c, primCount);
        else if (topology <= 2) ProcessLines(c, primCount, topology);
        else ProcessTriangles(c, primCount, topology);
        const uint64_t total = trace.size();
        std::memcpy(out, trace.data(), (size_t)std::min<uint64_t>(total, capacityWords) * 4);
        return (int64_t)(total / words);
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
}

// ApplyBlend on its own (Draw.cpp:1105-1262), the counterpart of oracle/ref_draw_check.cpp `blend`: state = the eight
// VkPipelineColorBlendAttachmentState members in order; source / destination / constant / out are float[4].
// Test hook: the fragment stage's view of its inputs — per input Location, the format SetDatum interpolates it as, interpolation
// kind, size and byte offset inside the vertex stage's record (Reflect above = GetVariablePointers, Draw.cpp:420-565, called
// with inputSize = sizeof(VertexBuiltinOutput) as ProcessFragmentShader does, :1613-1619). out = {count, then 5 words per input,
// then the final inputSize}; returns the number of words written or a negative value.
int cpvk_oracle_fragment_inputs(const uint32_t* spirv, uint32_t wordCount, uint32_t* out, uint32_t capacityWords) {
    try {
        StageInfo s;
        Parse(s.mod, spirv, wordCount, "main", 4, nullptr, 0);
        Reflect(s, false);
        const uint32_t need = 2 + 5 * (uint32_t)s.inputs.size();
        if (need > capacityWords) return -2;
        uint32_t k = 0, end = 24;
        out[k++] = (uint32_t)s.inputs.size();
        for (const InOut& io : s.inputs) {
            out[k++] = io.location; out[k++] = VariableFormat(s.mod, io.type); out[k++] = io.interpolation; out[k++] = io.size; out[k++] = io.offset;
            end = io.offset + io.size;
        }
        out[k++] = end;
        return (int)k;
    } catch (const std::exception& e) { g_error = e.what(); return -1; }
}

// Test hook: the vertex ids input assembly produces for a draw state (only count, first, vertexOffset and the index binding are read).
int cpvk_oracle_input_assembly(const CpvkDrawState* st, uint32_t* outVertexIds) {
    if (!st || !outVertexIds) return 1;
    for (uint32_t i = 0; i < st->count; i++) outVertexIds[i] = AssembledVertexId(*st, i);
    return 0;
}

int cpvk_oracle_apply_blend(const uint32_t state[8], const float* source, const float* destination, const float* constant, float* out) {
    try {
        CpvkBlendAttachment b{};
        b.blendEnable = state[0]; b.srcColorBlendFactor = state[1]; b.dstColorBlendFactor = state[2]; b.colorBlendOp = state[3];
        b.srcAlphaBlendFactor = state[4]; b.dstAlphaBlendFactor = state[5]; b.alphaBlendOp = state[6]; b.colorWriteMask = state[7];
        F4 s, d, k; std::memcpy(s.v, source, 16); std::memcpy(d.v, destination, 16); std::memcpy(k.v, constant, 16);
        const F4 r = ApplyBlend(s, d, k, b);
        std::memcpy(out, r.v, 16);
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return CPVK_E_UNSUPPORTED;
    }
}

// Shader runtime math one operation at a time (a14), the counterpart of oracle/ref_math_check.cpp: the interpreter's own
// GLSL.std.450 / OpDot / OpMatrixTimes* code (oracle_spirv.h) on raw 32-bit lanes. header = {kind, op, type, n, 0} as documented
// there; out receives 16 lanes.
int cpvk_oracle_math(const uint32_t header[5], const uint32_t* a, const uint32_t* b, const uint32_t* c, uint32_t* out) {
    try {
        const uint32_t kind = header[0], op = header[1], n = header[3];
        for (int i = 0; i < 16; i++) out[i] = 0;
        switch (kind) {
        case 0: Interp::GlslOp(op, a, b, c, (op == 66 || op == 67) ? 1 : n, n, out); break;
        case 1: { const float r = Interp::Dot(a, b, n); std::memcpy(out, &r, 4); break; }
        case 2: { float s; std::memcpy(&s, b, 4); for (uint32_t i = 0; i < n * n; i++) { float v; std::memcpy(&v, &a[i], 4); v = v * s; std::memcpy(&out[i], &v, 4); } break; } // OpMatrixTimesScalar
        case 3: Interp::VectorTimesMatrix(a, b, 4, 4, out); break;
        case 4: Interp::MatrixTimesVector(a, b, n, n, out); break;
        case 5: Interp::MatrixTimesMatrix(a, b, 4, 4, 4, out); break;
        default: Fail("cpvk_oracle_math: unknown kind");
        }
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return CPVK_E_UNSUPPORTED;
    }
}

// Render-pass clear == ClearImage: SetPixel on every texel of the subresource (Draw.cpp:117-149, ImageSampler.cpp:723-761).
int cpvk_oracle_clear(const CpvkAttachment* image, const CpvkClearValue* value, int isDepthStencil) {
    const FormatInfo fi = GetFormatInformation(image->format);
    if (fi.type == FmtType::Invalid) { g_error = "clear: unsupported format"; return CPVK_E_UNSUPPORTED; }
    for (uint32_t y = 0; y < image->height; y++)
        for (uint32_t x = 0; x < image->width; x++) {
            uint8_t* px = (uint8_t*)(uintptr_t)image->address + (uint64_t)y * image->rowPitch + (uint64_t)x * fi.totalSize;
            if (isDepthStencil) SetDepthStencil(fi, image->format, px, value->depthStencil.depth, (uint8_t)value->depthStencil.stencil);
            else if (fi.base == Base::UInt || fi.base == Base::SInt) SetPixelInt(fi, px, value->uint32);
            else SetPixelF32(fi, px, value->float32);
        }
    return 0;
}

int cpvk_oracle_copy_rows(uint64_t dst, uint32_t dstPitch, uint64_t src, uint32_t srcPitch, uint32_t rowBytes, uint32_t rows) {
    for (uint32_t r = 0; r < rows; r++) std::memcpy((uint8_t*)(uintptr_t)dst + (uint64_t)r * dstPitch, (const uint8_t*)(uintptr_t)src + (uint64_t)r * srcPitch, rowBytes);
    return 0;
}

// vkCmdBlitImage, one 2-D colour region (CommandBuffer.cpp:57-232): 3-D SampleImage (lod 1 on a 1-level chain ->
// level 0, clamp-to-edge, z taps weight 0) + SetPixel on the destination texel.
static int BlitImpl(const CpvkBlit* b, int32_t wx0, int32_t wy0, int32_t wx1, int32_t wy1) {
    const FormatInfo df = GetFormatInformation(b->dst.format), sf = GetFormatInformation(b->src.format);
    if (df.type == FmtType::Invalid || sf.type == FmtType::Invalid) { g_error = "blit: unsupported format"; return CPVK_E_UNSUPPORTED; }
    if (df.base == Base::UInt || df.base == Base::SInt) { g_error = "blit: integer formats not built yet"; return CPVK_E_UNSUPPORTED; }
    int32_t dstW = b->dstX1 - b->dstX0, dstH = b->dstY1 - b->dstY0;
    const bool negW = dstW < 0, negH = dstH < 0;
    if (negW) dstW = -dstW;
    if (negH) dstH = -dstH;
    CpvkDescriptor d{};
    d.type = CPVK_DESC_IMAGE; d.format = b->src.format; d.dimensions = 3; d.levelCount = 1;
    d.levels[0].address = b->src.address; d.levels[0].width = b->src.width; d.levels[0].height = b->src.height; d.levels[0].depth = 1;
    CpvkSampler s{}; s.magFilter = s.minFilter = b->filter; s.mipmapMode = 0; s.addressModeU = s.addressModeV = s.addressModeW = 2; s.borderColor = 0;
    for (int32_t y = 0; y < dstH; y++)
        for (int32_t x = 0; x < dstW; x++) {
            const int32_t dstX = negW ? x + b->dstX1 : x + b->dstX0;
            const int32_t dstY = negH ? y + b->dstY1 : y + b->dstY0;
            if (dstX < wx0 || dstX >= wx1 || dstY < wy0 || dstY >= wy1) continue; // window restriction (oracle-only: destination texels are independent)
            const float u = (dstX + 0.5f - b->dstX0) * ((float)(b->srcX1 - b->srcX0) / (b->dstX1 - b->dstX0)) + b->srcX0;
            const float v = (dstY + 0.5f - b->dstY0) * ((float)(b->srcY1 - b->srcY0) / (b->dstY1 - b->dstY0)) + b->srcY0;
            const float w = (0 + 0.5f - 0) * ((float)(1 - 0) / (1 - 0)) + 0;
            const float coord[3] = {u / b->src.width, v / b->src.height, w / 1};
            const Vec4f value = SampleImage(d, 3, coord, 1.0f, s, b->filter, b->filter);
            if (dstX < 0 || dstY < 0 || (uint32_t)dstX >= b->dst.width || (uint32_t)dstY >= b->dst.height) continue;
            uint8_t* px = (uint8_t*)(uintptr_t)b->dst.address + (uint64_t)dstY * b->dst.rowPitch + (uint64_t)dstX * df.totalSize;
            SetPixelF32(df, px, value.v);
        }
    return 0;
}
int cpvk_oracle_blit(const CpvkBlit* b) { return BlitImpl(b, INT32_MIN, INT32_MIN, INT32_MAX, INT32_MAX); }
// The same blit restricted to destination texels inside [x0, x1) x [y0, y1): how the 8K configurations are byte-compared
// in seconds (tests/test_fullsize_gpu.py).
int cpvk_oracle_blit_window(const CpvkBlit* b, int32_t x0, int32_t y0, int32_t x1, int32_t y1) { return BlitImpl(b, x0, y0, x1, y1); }

// ---- KAT helpers: expose the codec and the sampler one call at a time ----
int cpvk_oracle_format_info(uint32_t format, uint32_t out[4]) {
    const FormatInfo fi = GetFormatInformation(format);
    out[0] = (uint32_t)fi.type; out[1] = (uint32_t)fi.base; out[2] = fi.totalSize; out[3] = fi.elementSize;
    return fi.type == FmtType::Invalid ? CPVK_E_UNSUPPORTED : 0;
}
void cpvk_oracle_pack_f32(uint32_t format, const float* in, uint32_t count, uint8_t* out) {
    const FormatInfo fi = GetFormatInformation(format);
    for (uint32_t i = 0; i < count; i++) SetPixelF32(fi, out + (size_t)i * fi.totalSize, in + 4 * (size_t)i);
}
void cpvk_oracle_unpack_f32(uint32_t format, const uint8_t* in, uint32_t count, float* out) {
    const FormatInfo fi = GetFormatInformation(format);
    for (uint32_t i = 0; i < count; i++) GetPixelF32(fi, format, in + (size_t)i * fi.totalSize, out + 4 * (size_t)i);
}
void cpvk_oracle_pack_depth(uint32_t format, const float* depth, const uint8_t* stencil, uint32_t count, uint8_t* out) {
    const FormatInfo fi = GetFormatInformation(format);
    for (uint32_t i = 0; i < count; i++) SetDepthStencil(fi, format, out + (size_t)i * fi.totalSize, depth[i], stencil ? stencil[i] : 0);
}
void cpvk_oracle_unpack_depth(uint32_t format, const uint8_t* in, uint32_t count, float* out) {
    const FormatInfo fi = GetFormatInformation(format);
    for (uint32_t i = 0; i < count; i++) out[i] = GetDepth(format, in + (size_t)i * fi.totalSize);
}
void cpvk_oracle_sample(const CpvkDescriptor* d, const float* coords, uint32_t count, float lod, float* out) {
    for (uint32_t i = 0; i < count; i++) {
        const Vec4f r = ImageSampleExplicitLod(*d, coords + 3 * (size_t)i, lod);
        std::memcpy(out + 4 * (size_t)i, r.v, 16);
    }
}
// Test hook: ImageFetch (oracle_sampler.h) on caller-supplied integer coordinates (three per fetch), image or texel-buffer descriptor.
void cpvk_oracle_fetch(const CpvkDescriptor* d, const int32_t* coords, uint32_t count, float* out) {
    for (uint32_t i = 0; i < count; i++) {
        const Vec4f r = ImageFetch(*d, coords + 3 * (size_t)i);
        std::memcpy(out + 4 * (size_t)i, r.v, 16);
    }
}
// Full format-table row in the column order oracle/ref_formats_check.cpp prints the reference's FormatInformation:
// type total element base v0 v1 v2 v3 b0 b1 b2 b3 (Normal: byte offsets; Packed: bit offsets + widths; DepthStencil:
// depth/stencil offsets). FmtType here is {Invalid, Normal, Packed, DepthStencil}; the reference's FormatType starts
// at Normal = 0 (Formats.h:4-12), hence the -1.
int cpvk_oracle_format_row(uint32_t format, uint32_t out[12]) {
    const FormatInfo fi = GetFormatInformation(format);
    for (int i = 0; i < 12; i++) out[i] = 0;
    if (fi.type == FmtType::Invalid) return CPVK_E_UNSUPPORTED;
    out[0] = (uint32_t)fi.type - 1u; out[1] = fi.totalSize; out[2] = fi.elementSize; out[3] = (uint32_t)fi.base;
    if (fi.type == FmtType::Normal) { for (int c = 0; c < 4; c++) out[4 + c] = fi.offset[c]; }
    else if (fi.type == FmtType::Packed) { for (int c = 0; c < 4; c++) { out[4 + c] = fi.offset[c]; out[8 + c] = fi.bits[c]; } }
    else { out[4] = fi.depthOffset; out[5] = fi.stencilOffset; }
    return 0;
}
// GetNormalImageSize (Formats.cpp:455-483) + GetImagePixelOffset (:583-587): linear images, Stride = texel * width, mip
// levels back to back inside a layer, layers back to back. out = {total, layerSize, pixelSize, then per level
// offset, stride, planeSize, width, height, depth}.
int cpvk_oracle_image_layout(uint32_t format, uint32_t width, uint32_t height, uint32_t depth, uint32_t layers, uint32_t mips, uint64_t* out) {
    const FormatInfo fi = GetFormatInformation(format);
    if (fi.type == FmtType::Invalid) return CPVK_E_UNSUPPORTED;
    uint64_t layerSize = 0;
    for (uint32_t i = 0; i < mips; i++) {
        uint64_t* l = out + 3 + 6 * (size_t)i;
        l[0] = layerSize; l[1] = (uint64_t)fi.totalSize * width; l[2] = l[1] * height; l[3] = width; l[4] = height; l[5] = depth;
        layerSize += l[2] * depth;
        width = width / 2 > 1u ? width / 2 : 1u; height = height / 2 > 1u ? height / 2 : 1u; depth = depth / 2 > 1u ? depth / 2 : 1u;
    }
    out[0] = layerSize * layers; out[1] = layerSize; out[2] = fi.totalSize;
    return 0;
}
uint64_t cpvk_oracle_pixel_offset(const uint64_t* layout, int32_t i, int32_t j, int32_t k, uint32_t level, uint32_t layer) {
    const uint64_t* l = layout + 3 + 6 * (size_t)level;
    return layout[1] * layer + l[0] + (uint64_t)(int64_t)k * l[2] + (uint64_t)(int64_t)j * l[1] + (uint64_t)(int64_t)i * layout[2];
}
uint16_t cpvk_oracle_float_to_half(float v) { return FloatToHalf(v); }
float cpvk_oracle_half_to_float(uint16_t v) { return HalfToFloat(v); }

} // extern "C"
