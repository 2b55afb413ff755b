// spirv_to_cuda.cpp — lowers one SPIR-V shader stage to a CUDA C++ device function that nvJitLink later
// inlines into the prebuilt stage kernels (stage_kernels.cu). GPU-side replacement for
//   LLVMRuntime/SPIRVCompiler.cpp            (SPIR-V -> LLVM IR, module-global shader state, F6)
//   LLVMRuntime/PipelineCompiler.cpp:821-981 (vertex attribute fetch + output record store)
//   CPVulkan/CommandBuffer.Draw.cpp:420-565  (GetVariablePointers: interface reflection, record offsets)
// Design: every SSA value is scalarised into 32-bit C scalars (float / unsigned / bool), every variable is a
// per-thread word array, pointers are resolved symbolically at translation time, and function calls are
// inlined (SPIR-V forbids recursion), so the result is one flat function whose control flow is the SPIR-V CFG
// expressed with labels and gotos. Each SPIR-V arithmetic instruction becomes its own C statement and the whole
// pipeline is compiled with -fmad=false, so results carry exactly one IEEE rounding per SPIR-V operation, in the
// operand order the reference's runtime (glm / GlslFunctions.cpp) uses.
#include "spirv_to_cuda.h"

#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace cpvk {
namespace {

struct Unsupported : std::runtime_error { using std::runtime_error::runtime_error; };
struct Malformed : std::runtime_error { using std::runtime_error::runtime_error; };

enum : uint16_t {
    OpUndef = 1, OpExtInstImport = 11, OpExtInst = 12, OpEntryPoint = 15, OpExecutionMode = 16,
    OpTypeVoid = 19, OpTypeBool, OpTypeInt, OpTypeFloat, OpTypeVector, OpTypeMatrix, OpTypeImage, OpTypeSampler,
    OpTypeSampledImage, OpTypeArray, OpTypeRuntimeArray, OpTypeStruct, OpTypePointer = 32, OpTypeFunction = 33,
    OpConstantTrue = 41, OpConstantFalse, OpConstant, OpConstantComposite, OpConstantNull = 46,
    OpSpecConstantTrue = 48, OpSpecConstantFalse, OpSpecConstant, OpSpecConstantComposite,
    OpFunction = 54, OpFunctionParameter, OpFunctionEnd, OpFunctionCall, OpVariable = 59, OpLoad = 61, OpStore,
    OpAccessChain = 65, OpInBoundsAccessChain, OpDecorate = 71, OpMemberDecorate, OpVectorExtractDynamic = 77,
    OpVectorInsertDynamic, OpVectorShuffle, OpCompositeConstruct, OpCompositeExtract, OpCompositeInsert, OpCopyObject,
    OpTranspose, OpSampledImage = 86, OpImageSampleImplicitLod, OpImageSampleExplicitLod, OpImageFetch = 95, OpImageRead = 98,
    OpImage = 100, OpConvertFToU = 109, OpConvertFToS, OpConvertSToF, OpConvertUToF, OpBitcast = 124,
    OpSNegate = 126, OpFNegate, OpIAdd, OpFAdd, OpISub, OpFSub, OpIMul, OpFMul, OpUDiv, OpSDiv, OpFDiv, OpUMod, OpSRem,
    OpSMod, OpFRem, OpFMod, OpVectorTimesScalar, OpMatrixTimesScalar, OpVectorTimesMatrix, OpMatrixTimesVector,
    OpMatrixTimesMatrix, OpDot = 148, OpAny = 154, OpAll, OpIsNan, OpIsInf, OpLogicalEqual = 164, OpLogicalNotEqual,
    OpLogicalOr, OpLogicalAnd, OpLogicalNot, OpSelect, OpIEqual, OpINotEqual, OpUGreaterThan, OpSGreaterThan,
    OpUGreaterThanEqual, OpSGreaterThanEqual, OpULessThan, OpSLessThan, OpULessThanEqual, OpSLessThanEqual,
    OpFOrdEqual, OpFUnordEqual, OpFOrdNotEqual, OpFUnordNotEqual, OpFOrdLessThan, OpFUnordLessThan,
    OpFOrdGreaterThan, OpFUnordGreaterThan, OpFOrdLessThanEqual, OpFUnordLessThanEqual, OpFOrdGreaterThanEqual,
    OpFUnordGreaterThanEqual, OpShiftRightLogical = 194, OpShiftRightArithmetic, OpShiftLeftLogical, OpBitwiseOr,
    OpBitwiseXor, OpBitwiseAnd, OpNot, OpPhi = 245, OpLoopMerge, OpSelectionMerge, OpLabel, OpBranch,
    OpBranchConditional, OpSwitch, OpKill, OpReturn, OpReturnValue, OpUnreachable,
};
enum { DecoSpecId = 1, DecoArrayStride = 6, DecoMatrixStride = 7, DecoBuiltIn = 11, DecoNoPerspective = 13, DecoFlat = 14,
       DecoLocation = 30, DecoBinding = 33, DecoDescriptorSet = 34, DecoOffset = 35 };
enum { ScUniformConstant = 0, ScInput = 1, ScUniform = 2, ScOutput = 3, ScPrivate = 6, ScFunction = 7, ScPushConstant = 9, ScStorageBuffer = 12 };
enum { BiPosition = 0, BiPointSize = 1, BiClipDistance = 3, BiVertexId = 5, BiInstanceId = 6, BiFragCoord = 15, BiVertexIndex = 42, BiInstanceIndex = 43 };

struct Type {
    enum Kind { None, Void, Bool, Int, Float, Vector, Matrix, Array, RuntimeArray, Struct, Pointer, Function, Image, Sampler, SampledImage } kind = None;
    uint32_t width = 0; bool isSigned = false;
    uint32_t elem = 0, count = 0, storage = 0, lengthId = 0;
    std::vector<uint32_t> members;
    uint32_t words = 0;
    std::string leaves; // one char per logical word: f / u / b
};
struct Inst { uint16_t op; uint32_t type, result; const uint32_t* ops; uint32_t nops; };
struct Block { uint32_t label; std::vector<Inst> insts; };
struct Func { uint32_t id = 0, retType = 0; std::vector<uint32_t> params; std::vector<Block> blocks; std::map<uint32_t, size_t> blockIndex; };
struct Var { uint32_t id, ptrType, storage, initializer; };

struct Module {
    std::vector<uint32_t> words;
    uint32_t bound = 0, entry = 0, glsl = 0;
    bool originUpperLeft = false;
    std::vector<Type> types;
    std::vector<std::map<uint32_t, std::vector<uint32_t>>> deco;
    std::map<std::pair<uint32_t, uint32_t>, std::map<uint32_t, std::vector<uint32_t>>> memberDeco;
    std::vector<Var> vars;
    std::map<uint32_t, size_t> varIndex;
    std::map<uint32_t, Func> funcs;
    std::map<uint32_t, std::vector<uint32_t>> constants; // id -> logical words
    std::map<uint32_t, uint32_t> idType;

    bool hasDeco(uint32_t id, uint32_t d) const { return deco[id].count(d) != 0; }
    uint32_t decoVal(uint32_t id, uint32_t d, uint32_t def = 0) const { auto it = deco[id].find(d); return it == deco[id].end() || it->second.empty() ? def : it->second[0]; }
    bool memberDecoVal(uint32_t id, uint32_t mem, uint32_t d, uint32_t& out) const {
        auto it = memberDeco.find({id, mem}); if (it == memberDeco.end()) return false;
        auto jt = it->second.find(d); if (jt == it->second.end()) return false;
        out = jt->second.empty() ? 0 : jt->second[0]; return true;
    }
};

void FinishType(Module& m, uint32_t id) {
    Type& t = m.types[id];
    switch (t.kind) {
    case Type::Bool: t.words = 1; t.leaves = "b"; break;
    case Type::Int: t.words = 1; t.leaves = "u"; break;
    case Type::Float: t.words = 1; t.leaves = "f"; break;
    case Type::Vector: case Type::Matrix: case Type::Array: {
        const Type& e = m.types[t.elem]; t.words = e.words * t.count; t.leaves.clear();
        for (uint32_t k = 0; k < t.count; k++) t.leaves += e.leaves; break; }
    case Type::Struct: t.words = 0; t.leaves.clear(); for (uint32_t mm : t.members) { t.words += m.types[mm].words; t.leaves += m.types[mm].leaves; } break;
    default: break;
    }
}

void Parse(Module& m, const CpvkShaderStage& stage, uint32_t model) {
    const uint32_t* code = stage.spirv; const size_t n = stage.wordCount;
    if (!code || n < 5 || code[0] != 0x07230203u) throw Malformed("not a SPIR-V module");
    m.words.assign(code, code + n);
    const uint32_t* w = m.words.data();
    m.bound = w[3];
    m.types.assign(m.bound, Type()); m.deco.assign(m.bound, {});
    const std::string entryName = stage.entryPoint ? stage.entryPoint : "main";
    Func* fn = nullptr; Block* blk = nullptr;
    struct Pending { uint32_t id; const uint32_t* ins; };
    std::vector<Pending> consts;
    auto checkId = [&](uint32_t id) { if (id >= m.bound) throw Malformed("id out of bounds"); return id; };
    for (size_t i = 5; i < n;) {
        const uint32_t wc = w[i] >> 16, op = w[i] & 0xFFFF;
        if (wc == 0 || i + wc > n) throw Malformed("truncated instruction");
        const uint32_t* o = w + i + 1;
        // operands are only read after their presence was checked (a truncated or malformed module from the application is an
        // error, CPVK_E_SPIRV, not an out-of-bounds read), and every id a type refers to must lie below the module's bound
        auto need = [&](uint32_t words) { if (wc < words) throw Malformed("instruction too short"); };
        switch (op) {
        case OpExtInstImport: case OpDecorate: need(3); break;
        case OpTypePointer: case OpTypeArray: case OpEntryPoint: case OpMemberDecorate: need(4); break;
        case OpTypeInt: case OpTypeVector: case OpTypeMatrix: need(4); break;
        case OpTypeFloat: case OpTypeSampledImage: case OpTypeRuntimeArray: case OpTypeFunction: case OpUndef: case OpFunctionParameter: need(3); break;
        case OpTypeImage: need(9); break;
        case OpTypeVoid: case OpTypeBool: case OpTypeSampler: case OpTypeStruct: case OpLabel: need(2); break;
        case OpConstantTrue: case OpConstantFalse: case OpConstantNull: case OpSpecConstantTrue: case OpSpecConstantFalse: case OpConstantComposite: case OpSpecConstantComposite: need(3); break;
        case OpConstant: case OpSpecConstant: case OpVariable: need(4); break;
        case OpFunction: need(5); break;
        default: break;
        }
        switch (op) {
        case OpTypeVector: case OpTypeMatrix: case OpTypeImage: case OpTypeSampledImage: case OpTypeRuntimeArray: case OpTypeFunction: checkId(o[1]); break;
        case OpTypeArray: checkId(o[1]); checkId(o[2]); break;
        case OpTypePointer: checkId(o[2]); break;
        case OpTypeStruct: for (uint32_t k = 1; k + 1 < wc; k++) checkId(o[k]); break;
        case OpConstant: case OpSpecConstant: case OpConstantTrue: case OpConstantFalse: case OpConstantNull: case OpSpecConstantTrue: case OpSpecConstantFalse:
        case OpConstantComposite: case OpSpecConstantComposite: case OpUndef: case OpVariable: case OpFunction: case OpFunctionParameter: checkId(o[0]); break;
        case OpMemberDecorate: case OpLabel: checkId(o[0]); break;
        default: break;
        }
        if (op == OpTypeFunction) for (uint32_t k = 2; k + 1 < wc; k++) checkId(o[k]);
        if (op == OpConstantComposite || op == OpSpecConstantComposite) for (uint32_t k = 2; k + 1 < wc; k++) checkId(o[k]);
        switch (op) {
        case OpExtInstImport: if (std::string(reinterpret_cast<const char*>(o + 1)) == "GLSL.std.450") m.glsl = o[0]; break;
        case OpEntryPoint: if (o[0] == model && entryName == reinterpret_cast<const char*>(o + 2)) m.entry = o[1]; break;
        case OpDecorate: m.deco[checkId(o[0])][o[1]] = std::vector<uint32_t>(o + 2, o + wc - 1); break;
        case OpMemberDecorate: m.memberDeco[{o[0], o[1]}][o[2]] = std::vector<uint32_t>(o + 3, o + wc - 1); break;
        case OpTypeVoid: m.types[checkId(o[0])].kind = Type::Void; break;
        case OpTypeBool: m.types[checkId(o[0])].kind = Type::Bool; FinishType(m, o[0]); break;
        case OpTypeInt: { Type& t = m.types[checkId(o[0])]; t.kind = Type::Int; t.width = o[1]; t.isSigned = o[2] != 0;
            if (t.width != 32) throw Unsupported("only 32-bit integer types are built"); FinishType(m, o[0]); break; }
        case OpTypeFloat: { Type& t = m.types[checkId(o[0])]; t.kind = Type::Float; t.width = o[1];
            if (t.width != 32) throw Unsupported("only 32-bit float types are built"); FinishType(m, o[0]); break; }
        case OpTypeVector: { Type& t = m.types[checkId(o[0])]; t.kind = Type::Vector; t.elem = o[1]; t.count = o[2]; FinishType(m, o[0]); break; }
        case OpTypeMatrix: { Type& t = m.types[checkId(o[0])]; t.kind = Type::Matrix; t.elem = o[1]; t.count = o[2]; FinishType(m, o[0]); break; }
        case OpTypeImage: { Type& t = m.types[checkId(o[0])]; t.kind = Type::Image; t.elem = o[1]; t.count = o[2]; break; }
        case OpTypeSampler: m.types[checkId(o[0])].kind = Type::Sampler; break;
        case OpTypeSampledImage: { Type& t = m.types[checkId(o[0])]; t.kind = Type::SampledImage; t.elem = o[1]; break; }
        case OpTypeArray: { Type& t = m.types[checkId(o[0])]; t.kind = Type::Array; t.elem = o[1]; t.lengthId = o[2]; break; }
        case OpTypeRuntimeArray: { Type& t = m.types[checkId(o[0])]; t.kind = Type::RuntimeArray; t.elem = o[1]; break; }
        case OpTypeStruct: { Type& t = m.types[checkId(o[0])]; t.kind = Type::Struct; t.members.assign(o + 1, o + wc - 1); break; }
        case OpTypePointer: { Type& t = m.types[checkId(o[0])]; t.kind = Type::Pointer; t.storage = o[1]; t.elem = o[2]; break; }
        case OpTypeFunction: { Type& t = m.types[checkId(o[0])]; t.kind = Type::Function; t.elem = o[1]; t.members.assign(o + 2, o + wc - 1); break; }
        case OpConstantTrue: case OpConstantFalse: case OpConstant: case OpConstantComposite: case OpConstantNull:
        case OpSpecConstantTrue: case OpSpecConstantFalse: case OpSpecConstant: case OpSpecConstantComposite:
            consts.push_back({checkId(o[1]), w + i}); m.idType[o[1]] = o[0]; break;
        case OpUndef:
            m.idType[checkId(o[1])] = o[0];
            if (!fn) consts.push_back({o[1], w + i}); else if (blk) blk->insts.push_back(Inst{(uint16_t)op, o[0], o[1], o + 2, 0});
            break;
        case OpVariable:
            m.idType[checkId(o[1])] = o[0];
            if (o[2] != ScFunction) { m.varIndex[o[1]] = m.vars.size(); m.vars.push_back(Var{o[1], o[0], o[2], wc > 4 ? o[3] : 0}); }
            else if (blk) blk->insts.push_back(Inst{(uint16_t)op, o[0], o[1], o + 2, wc - 3});
            break;
        case OpFunction: { Func f; f.id = checkId(o[1]); f.retType = o[0]; m.funcs[f.id] = f; fn = &m.funcs[f.id]; break; }
        case OpFunctionParameter: if (fn) { fn->params.push_back(checkId(o[1])); m.idType[o[1]] = o[0]; } break;
        case OpFunctionEnd: fn = nullptr; blk = nullptr; break;
        case OpLabel: if (fn) { fn->blockIndex[o[0]] = fn->blocks.size(); fn->blocks.push_back(Block{o[0], {}}); blk = &fn->blocks.back(); } break;
        default:
            if (fn && blk) {
                bool noResult = false;
                switch (op) { case OpStore: case OpLoopMerge: case OpSelectionMerge: case OpBranch: case OpBranchConditional: case OpSwitch:
                              case OpKill: case OpReturn: case OpReturnValue: case OpUnreachable: case 0: case 8 /*OpLine*/: case 317 /*OpNoLine*/: noResult = true; break; default: break; }
                Inst in{};
                in.op = (uint16_t)op;
                if (noResult) { in.ops = o; in.nops = wc - 1; }
                else { if (wc < 3) throw Malformed("instruction too short"); in.type = checkId(o[0]); in.result = checkId(o[1]); in.ops = o + 2; in.nops = wc - 3; m.idType[in.result] = in.type; }
                blk->insts.push_back(in);
            }
            break;
        }
        i += wc;
    }
    if (!m.entry) throw Malformed("entry point '" + entryName + "' not found");
    for (size_t i = 5; i < n; i += w[i] >> 16)
        if ((w[i] & 0xFFFF) == OpExecutionMode && w[i + 1] == m.entry && w[i + 2] == 7) m.originUpperLeft = true;
    // scalar constants first (array lengths), honouring specialisation
    auto specOverride = [&](uint32_t id, uint32_t& v) {
        if (!m.hasDeco(id, DecoSpecId)) return;
        const uint32_t sid = m.decoVal(id, DecoSpecId);
        for (uint32_t k = 0; k < stage.specCount; k++) if (stage.spec[k].constantId == sid) v = stage.spec[k].value;
    };
    for (auto& c : consts) {
        const uint32_t op = c.ins[0] & 0xFFFF;
        uint32_t v;
        switch (op) {
        case OpConstant: m.constants[c.id] = {c.ins[3]}; break;
        case OpSpecConstant: v = c.ins[3]; specOverride(c.id, v); m.constants[c.id] = {v}; break;
        case OpConstantTrue: m.constants[c.id] = {1}; break;
        case OpConstantFalse: m.constants[c.id] = {0}; break;
        case OpSpecConstantTrue: v = 1; specOverride(c.id, v); m.constants[c.id] = {v ? 1u : 0u}; break;
        case OpSpecConstantFalse: v = 0; specOverride(c.id, v); m.constants[c.id] = {v ? 1u : 0u}; break;
        default: break;
        }
    }
    for (uint32_t id = 0; id < m.bound; id++) {
        Type& t = m.types[id];
        if (t.kind == Type::Array) { auto it = m.constants.find(t.lengthId); if (it == m.constants.end()) throw Malformed("array length is not a constant"); t.count = it->second[0]; }
        if (t.kind == Type::Array || t.kind == Type::Struct || t.kind == Type::Matrix || t.kind == Type::Vector) FinishType(m, id);
    }
    for (auto& c : consts) {
        const uint32_t op = c.ins[0] & 0xFFFF, wc = c.ins[0] >> 16, ty = c.ins[1];
        if (op == OpConstantNull || op == OpUndef) m.constants[c.id] = std::vector<uint32_t>(m.types[ty].words ? m.types[ty].words : 1, 0u);
        else if (op == OpConstantComposite || op == OpSpecConstantComposite) {
            std::vector<uint32_t> v;
            for (uint32_t k = 3; k < wc; k++) { auto it = m.constants.find(c.ins[k]); if (it == m.constants.end()) throw Malformed("composite of non-constant"); v.insert(v.end(), it->second.begin(), it->second.end()); }
            m.constants[c.id] = v;
        }
    }
}

// ---- host copy of the few format facts vertex fetch needs (Formats.cpp:219-341, PipelineCompiler.cpp:627-710) ----
struct HostFormat { bool valid = false, simple = false, isFloat = false, isSignedInt = false; uint32_t elemBytes = 0, comps = 0; };
HostFormat ClassifyFormat(uint32_t f) {
    HostFormat r; uint32_t k = 0;
    if (f >= 9 && f <= 50) { const uint32_t fam = (f - 9) / 7; k = (f - 9) % 7; r.valid = true; r.elemBytes = 1; r.comps = fam == 0 ? 1 : fam == 1 ? 2 : (fam == 2 || fam == 3) ? 3 : 4;
        const bool bgr = fam == 3 || fam == 5; r.simple = !bgr && (k == 4 || k == 5); r.isSignedInt = k == 5; }
    else if (f >= 51 && f <= 69) { r.valid = true; r.comps = 4; r.elemBytes = 0; }
    else if (f >= 70 && f <= 97) { k = (f - 70) % 7; r.valid = true; r.elemBytes = 2; r.comps = (f - 70) / 7 + 1; r.simple = k >= 4; r.isSignedInt = k == 5; r.isFloat = k == 6; }
    else if (f >= 98 && f <= 109) { k = (f - 98) % 3; r.valid = true; r.elemBytes = 4; r.comps = (f - 98) / 3 + 1; r.simple = true; r.isSignedInt = k == 1; r.isFloat = k == 2; }
    return r;
}

struct Ptr {
    enum Kind { Local, Buffer, Handle } kind = Local;
    std::string base;       // Local: word array; Buffer: byte pointer expression; Handle: descriptor pointer expression
    uint32_t type = 0;      // pointee
    uint32_t off = 0;       // static offset: words (Local) or bytes (Buffer)
    std::string dyn;        // dynamic offset expression (same unit), may be empty
    uint32_t matStride = 0;
    bool writable = true;
};

class Translator {
public:
    Translator(const CpvkShaderStage& st, uint32_t model, const CpvkPipelineDesc& d, PipelineLayoutInfo& l) : stage(st), model(model), desc(d), layout(l) {}

    std::string Run() {
        Parse(m, stage, model);
        std::ostringstream pre;
        const bool vs = model == 0;
        Prologue(pre, vs);
        EmitFunction(m.entry, 0, {}, 0);
        body << "L_end: ;\n";
        Epilogue(vs);
        std::ostringstream out;
        if (vs) out << "extern \"C\" __device__ void cpvk_vs_main(cpvk_u32 vertexId, cpvk_u32 instanceId, cpvk_u32 rawId, const CpvkDrawParams* dp) {\n";
        else out << "extern \"C\" __device__ bool cpvk_fs_main(const CpvkFragCtx* ctx, CpvkFragOut* out) {\n  const CpvkDrawParams* dp = ctx->dp; (void)dp;\n";
        out << "  bool discard_ = false; (void)discard_;\n";
        out << (vs ? "  const float* lut_ = nullptr; (void)lut_;\n" : "  const float* lut_ = ctx->unorm8; __builtin_assume(lut_ != nullptr);\n");
        for (auto& a : arrays) out << "  " << a << "\n";
        EmitDecls(out, "float", declF); EmitDecls(out, "unsigned", declU); EmitDecls(out, "bool", declB); EmitDecls(out, "CpvkVec4", declV);
        out << pre.str() << body.str();
        out << (vs ? "}\n" : "  return discard_;\n}\n");
        return out.str();
    }

private:
    const CpvkShaderStage& stage; uint32_t model; const CpvkPipelineDesc& desc; PipelineLayoutInfo& layout;
    bool storesToMemory = false;
    Module m;
    std::ostringstream body;
    std::set<std::string> declF, declU, declB, declV;
    std::vector<std::string> arrays;
    std::map<std::pair<int, uint32_t>, Ptr> ptrs;
    std::map<std::pair<int, uint32_t>, std::string> handles;
    std::map<std::pair<int, uint32_t>, std::string> samplerOf; // OpSampledImage results: the descriptor that supplies the sampler state
    void CopyHandle(std::pair<int, uint32_t> dst, std::pair<int, uint32_t> src) {
        handles[dst] = handles.at(src);
        auto s = samplerOf.find(src);
        if (s != samplerOf.end()) samplerOf[dst] = s->second; else samplerOf.erase(dst);
    }
    int ctxCounter = 0, tmpCounter = 0;
    uint32_t perVertexVar = 0, positionVar = 0, pointSizeVar = 0;

    static void EmitDecls(std::ostringstream& o, const char* ty, const std::set<std::string>& names) {
        if (names.empty()) return;
        o << "  " << ty << " "; bool first = true;
        for (auto& n : names) { o << (first ? "" : ", ") << n; first = false; }
        o << ";\n";
    }
    static std::string Hex(uint32_t v) { char b[32]; snprintf(b, sizeof b, "0x%08xu", v); return b; }
    static std::string Lit(char kind, uint32_t bits) {
        if (kind == 'f') return "__uint_as_float(" + Hex(bits) + ")";
        if (kind == 'b') return bits ? "true" : "false";
        return Hex(bits);
    }
    const Type& T(uint32_t id) const { if (id >= m.bound || m.types[id].kind == Type::None) throw Malformed("bad type id"); return m.types[id]; }
    uint32_t TypeOf(uint32_t id) const { auto it = m.idType.find(id); if (it == m.idType.end()) throw Malformed("untyped id " + std::to_string(id)); return it->second; }

    std::string Name(int ctx, uint32_t id, uint32_t k, char kind) {
        std::string n = "v" + std::to_string(ctx) + "_" + std::to_string(id) + "_" + std::to_string(k);
        (kind == 'f' ? declF : kind == 'b' ? declB : declU).insert(n);
        return n;
    }
    // r-value expression of logical word k of value `id`
    std::string W(int ctx, uint32_t id, uint32_t k) {
        auto c = m.constants.find(id);
        const std::string& leaves = T(TypeOf(id)).leaves;
        if (k >= leaves.size()) throw Malformed("word index out of range");
        if (c != m.constants.end()) return Lit(leaves[k], c->second[k]);
        return Name(ctx, id, k, leaves[k]);
    }
    std::string Dst(int ctx, const Inst& in, uint32_t k) { return Name(ctx, in.result, k, T(in.type).leaves[k]); }
    static std::string ToWord(char kind, const std::string& e) { return kind == 'f' ? "__float_as_uint(" + e + ")" : kind == 'b' ? "((" + e + ") ? 1u : 0u)" : e; }
    static std::string FromWord(char kind, const std::string& e) { return kind == 'f' ? "__uint_as_float(" + e + ")" : kind == 'b' ? "((" + e + ") != 0u)" : e; }

    // ---- resources ----
    uint32_t SlotFor(uint32_t set, uint32_t binding, uint32_t count) {
        for (auto& s : layout.slots) if (s.set == set && s.binding == binding) { if (s.count < count) throw Unsupported("descriptor array size mismatch between stages"); return s.slotBase; }
        if (layout.slotCount + count > CPVK_MAX_DESCRIPTORS) throw Unsupported("too many descriptors");
        ResourceSlot s{set, binding, count, layout.slotCount};
        layout.slots.push_back(s); layout.slotCount += count;
        return s.slotBase;
    }

    Ptr VarPtr(int ctx, uint32_t id) {
        auto it = ptrs.find({ctx, id});
        if (it != ptrs.end()) return it->second;
        auto vi = m.varIndex.find(id);
        if (vi == m.varIndex.end()) throw Malformed("pointer id " + std::to_string(id) + " has no definition in scope");
        const Var& v = m.vars[vi->second];
        Ptr p; p.type = T(v.ptrType).elem;
        const Type& pt = T(p.type);
        switch (v.storage) {
        case ScInput: case ScOutput: case ScPrivate: p.kind = Ptr::Local; p.base = "g" + std::to_string(id); break;
        case ScPushConstant: p.kind = Ptr::Buffer; p.base = "((const cpvk_u8*)dp->push)"; p.writable = false; break;
        case ScUniform: case ScStorageBuffer: case ScUniformConstant: {
            const uint32_t set = m.decoVal(id, DecoDescriptorSet), binding = m.decoVal(id, DecoBinding);
            const bool isArray = pt.kind == Type::Array;
            const Type& et = isArray ? T(pt.elem) : pt;
            const uint32_t slot = SlotFor(set, binding, isArray ? pt.count : 1);
            if (et.kind == Type::Image || et.kind == Type::SampledImage || et.kind == Type::Sampler) {
                p.kind = Ptr::Handle; p.base = "(dp->desc + " + std::to_string(slot) + ")";
            } else {
                if (isArray) throw Unsupported("arrays of buffer descriptors");
                p.kind = Ptr::Buffer; p.base = "((const cpvk_u8*)dp->desc[" + std::to_string(slot) + "].address)"; p.writable = v.storage == ScStorageBuffer;
            }
            break; }
        default: throw Unsupported("storage class " + std::to_string(v.storage));
        }
        return p;
    }

    uint32_t ArrayStride(uint32_t typeId) const { return m.decoVal(typeId, DecoArrayStride, T(T(typeId).elem).words * 4); }

    // byte offsets of every logical word of a buffer-resident object
    void BufferLeaves(uint32_t typeId, uint32_t byteOff, uint32_t matStride, std::vector<uint32_t>& out) const {
        const Type& t = T(typeId);
        switch (t.kind) {
        case Type::Bool: case Type::Int: case Type::Float: out.push_back(byteOff); break;
        case Type::Vector: for (uint32_t k = 0; k < t.count; k++) out.push_back(byteOff + 4 * k); break;
        case Type::Matrix: { const uint32_t st = matStride ? matStride : T(t.elem).words * 4; for (uint32_t c = 0; c < t.count; c++) BufferLeaves(t.elem, byteOff + c * st, 0, out); break; }
        case Type::Array: { const uint32_t st = ArrayStride(typeId); for (uint32_t k = 0; k < t.count; k++) BufferLeaves(t.elem, byteOff + k * st, matStride, out); break; }
        case Type::Struct: for (uint32_t k = 0; k < t.members.size(); k++) { uint32_t bo = 0, ms = 0; m.memberDecoVal(typeId, k, DecoOffset, bo); m.memberDecoVal(typeId, k, DecoMatrixStride, ms); BufferLeaves(t.members[k], byteOff + bo, ms, out); } break;
        default: throw Unsupported("buffer access to this type");
        }
    }

    std::string LocalIndex(const Ptr& p, uint32_t k) const { std::string s = std::to_string(p.off + k); if (!p.dyn.empty()) s += " + " + p.dyn; return p.base + "[" + s + "]"; }

    void EmitLoad(int ctx, const Inst& in, const Ptr& p) {
        if (p.kind == Ptr::Handle) { handles[{ctx, in.result}] = p.dyn.empty() ? p.base : "(" + p.base + " + " + p.dyn + ")"; return; }
        const Type& t = T(in.type);
        if (p.kind == Ptr::Local) {
            for (uint32_t k = 0; k < t.words; k++) body << "  " << Dst(ctx, in, k) << " = " << FromWord(t.leaves[k], LocalIndex(p, k)) << ";\n";
        } else {
            std::vector<uint32_t> offs; BufferLeaves(in.type, p.off, p.matStride, offs);
            for (uint32_t k = 0; k < t.words; k++) {
                std::string o = std::to_string(offs[k]) + "ull"; if (!p.dyn.empty()) o += " + (cpvk_u64)(" + p.dyn + ")";
                body << "  " << Dst(ctx, in, k) << " = " << FromWord(t.leaves[k], "cpvk_buf_ld(" + p.base + ", " + o + ")") << ";\n";
            }
        }
    }
    void EmitStore(int ctx, const Ptr& p, uint32_t valueId) {
        const Type& t = T(p.type);
        if (p.kind == Ptr::Handle) throw Unsupported("store to an opaque handle");
        if (p.kind == Ptr::Local) {
            for (uint32_t k = 0; k < t.words; k++) body << "  " << LocalIndex(p, k) << " = " << ToWord(t.leaves[k], W(ctx, valueId, k)) << ";\n";
        } else {
            if (!p.writable) throw Unsupported("store to a read-only buffer");
            storesToMemory = true;
            std::vector<uint32_t> offs; BufferLeaves(p.type, p.off, p.matStride, offs);
            for (uint32_t k = 0; k < t.words; k++) {
                std::string o = std::to_string(offs[k]) + "ull"; if (!p.dyn.empty()) o += " + (cpvk_u64)(" + p.dyn + ")";
                body << "  cpvk_buf_st((cpvk_u8*)" << p.base << ", " << o << ", " << ToWord(t.leaves[k], W(ctx, valueId, k)) << ");\n";
            }
        }
    }

    Ptr AccessChain(int ctx, const Inst& in) {
        Ptr p = VarPtr(ctx, in.ops[0]);
        auto addDyn = [&](const std::string& idx, uint32_t scale) {
            const std::string term = "(" + idx + ") * " + std::to_string(scale) + "u";
            p.dyn = p.dyn.empty() ? term : p.dyn + " + " + term;
        };
        for (uint32_t k = 1; k < in.nops; k++) {
            const uint32_t idxId = in.ops[k];
            auto c = m.constants.find(idxId);
            const bool isConst = c != m.constants.end();
            const uint32_t ci = isConst ? c->second[0] : 0;
            const Type& t = T(p.type);
            if (p.kind == Ptr::Handle) {
                if (t.kind != Type::Array) throw Malformed("access chain into a handle");
                if (isConst) p.base = "(" + p.base + " + " + std::to_string(ci) + ")"; else p.dyn = W(ctx, idxId, 0);
                p.type = t.elem; continue;
            }
            switch (t.kind) {
            case Type::Struct: {
                if (!isConst || ci >= t.members.size()) throw Malformed("struct index must be a constant in range");
                if (p.kind == Ptr::Buffer) { uint32_t bo = 0, ms = 0; m.memberDecoVal(p.type, ci, DecoOffset, bo); m.memberDecoVal(p.type, ci, DecoMatrixStride, ms); p.off += bo; p.matStride = ms; }
                else { for (uint32_t q = 0; q < ci; q++) p.off += T(t.members[q]).words; }
                p.type = t.members[ci]; break; }
            case Type::Array: case Type::RuntimeArray: case Type::Matrix: case Type::Vector: {
                uint32_t scale;
                if (p.kind == Ptr::Buffer) scale = t.kind == Type::Vector ? 4 : t.kind == Type::Matrix ? (p.matStride ? p.matStride : T(t.elem).words * 4) : ArrayStride(p.type);
                else scale = T(t.elem).words;
                if (isConst) p.off += ci * scale; else addDyn(W(ctx, idxId, 0), scale);
                p.type = t.elem; break; }
            default: throw Malformed("access chain into a scalar");
            }
        }
        return p;
    }

    // ---- prologue: shader inputs ----
    uint32_t VariableSize(uint32_t ty) const { // GetVariableSize (Draw.cpp:297-354), fragment side
        const Type& t = T(ty);
        switch (t.kind) {
        case Type::Array: return VariableSize(t.elem) * t.count;
        case Type::Matrix: return 4 * t.count * T(t.elem).count;
        case Type::Vector: return 4 * t.count;
        case Type::Int: case Type::Float: return 4;
        case Type::Struct: { uint32_t s = 0; for (uint32_t k = 0; k < t.members.size(); k++) { uint32_t o; if (m.memberDecoVal(ty, k, DecoOffset, o) && o > s) s = o; s += VariableSize(t.members[k]); } return s; }
        default: throw Unsupported("interface variable type");
        }
    }
    uint32_t AllocSize(uint32_t ty) const { // LLVM alloc size inside the packed _Output struct (PipelineCompiler.cpp:585-601)
        const Type& t = T(ty);
        switch (t.kind) {
        case Type::Array: case Type::Matrix: return AllocSize(t.elem) * t.count;
        case Type::Vector: return t.count == 3 ? 16 : 4 * t.count;
        case Type::Int: case Type::Float: case Type::Bool: return 4;
        default: throw Unsupported("vertex output type");
        }
    }
    uint32_t VariableFormat(uint32_t ty) const { // GetVariableFormat (Draw.cpp:151-295)
        const Type& t = T(ty); const Type& e = t.kind == Type::Vector ? T(t.elem) : t; const uint32_t n = t.kind == Type::Vector ? t.count : 1;
        static const uint32_t fl[5] = {0, 100, 103, 106, 109}, si[5] = {0, 99, 102, 105, 108}, ui[5] = {0, 98, 101, 104, 107};
        if (n > 4) return 0;
        if (e.kind == Type::Float) return fl[n];
        if (e.kind == Type::Int) return e.isSigned ? si[n] : ui[n];
        return 0;
    }

    void EmitFetch(std::ostringstream& o, uint32_t location, uint32_t ty, const std::string& dst) { // EmitCopyInput
        const CpvkVertexAttribute* attr = nullptr; const CpvkVertexBinding* bind = nullptr;
        for (uint32_t i = 0; i < desc.attributeCount; i++) if (desc.attributes[i].location == location) { attr = &desc.attributes[i]; break; }
        if (!attr) throw Unsupported("no vertex attribute for location " + std::to_string(location) + " (FATAL_ERROR, PipelineCompiler.cpp:603-613)");
        for (uint32_t i = 0; i < desc.bindingCount; i++) if (desc.bindings[i].binding == attr->binding) { bind = &desc.bindings[i]; break; }
        if (!bind || bind->binding >= CPVK_MAX_VERTEX_BINDINGS) throw Unsupported("no vertex binding for attribute");
        const Type& t = T(ty); const Type& et = t.kind == Type::Vector ? T(t.elem) : t; const uint32_t comps = t.kind == Type::Vector ? t.count : 1;
        const std::string a = "a" + std::to_string(tmpCounter++);
        o << "  const cpvk_u8* " << a << " = cpvk_attr_ptr(dp, " << bind->binding << "u, " << bind->stride << "u, " << (bind->inputRate == 0 ? "vertexId" : "instanceId")
          << ", " << attr->offset << "u);\n";
        const HostFormat hf = ClassifyFormat(attr->format);
        if (!hf.valid) throw Unsupported("vertex attribute format " + std::to_string(attr->format));
        if (VariableFormat(ty) == attr->format) { o << "  cpvk_fetch_raw(" << a << ", " << comps << ", " << dst << ");\n"; return; }
        if (hf.simple) {
            if (hf.comps != comps) throw Unsupported("attribute/shader component count mismatch (FATAL_ERROR, PipelineCompiler.cpp:719-722)");
            if (hf.isFloat != (et.kind == Type::Float)) throw Unsupported("float<->int attribute conversion (TODO_ERROR, PipelineCompiler.cpp:736-749)");
            if (hf.isFloat) o << "  cpvk_fetch_half(" << a << ", " << comps << ", " << dst << ");\n";
            else o << "  cpvk_fetch_int(" << a << ", " << comps << ", " << hf.elemBytes << ", " << (et.isSigned ? "true" : "false") << ", " << dst << ");\n";
            return;
        }
        if (et.kind == Type::Float) o << "  cpvk_fetch_format_f32(" << attr->format << "u, " << a << ", " << comps << ", " << dst << ");\n";
        else o << "  cpvk_fetch_format_int(" << attr->format << "u, " << a << ", " << comps << ", " << dst << ");\n";
    }

    void Prologue(std::ostringstream& o, bool vs) {
        uint32_t inOff = 24; // FS input record offsets start after {vec4, float, float[1]} (Draw.cpp:1613)
        for (const Var& v : m.vars) {
            const uint32_t pointee = T(v.ptrType).elem;
            const Type& pt = T(pointee);
            if (v.storage == ScInput || v.storage == ScOutput || v.storage == ScPrivate) {
                const std::string g = "g" + std::to_string(v.id);
                arrays.push_back("unsigned " + g + "[" + std::to_string(pt.words ? pt.words : 1) + "] = {0};");
                if (v.initializer) { auto c = m.constants.find(v.initializer); if (c != m.constants.end()) for (uint32_t k = 0; k < pt.words; k++) o << "  " << g << "[" << k << "] = " << Hex(c->second[k]) << ";\n"; }
            }
            if (v.storage == ScOutput && pt.kind == Type::Struct) {
                uint32_t b; for (uint32_t k = 0; k < pt.members.size(); k++) if (m.memberDecoVal(pointee, k, DecoBuiltIn, b)) perVertexVar = v.id;
            }
            if (v.storage == ScOutput && m.hasDeco(v.id, DecoBuiltIn)) {
                const uint32_t b = m.decoVal(v.id, DecoBuiltIn);
                if (b == BiPosition) positionVar = v.id; else if (b == BiPointSize) pointSizeVar = v.id;
            }
            if (v.storage != ScInput) continue;
            const std::string g = "g" + std::to_string(v.id);
            if (m.hasDeco(v.id, DecoBuiltIn)) {
                const uint32_t b = m.decoVal(v.id, DecoBuiltIn);
                if (vs && (b == BiVertexIndex || b == BiVertexId)) o << "  " << g << "[0] = vertexId;\n";
                else if (vs && (b == BiInstanceIndex || b == BiInstanceId)) o << "  " << g << "[0] = instanceId;\n";
                else if (!vs && b == BiFragCoord) for (int k = 0; k < 4; k++) o << "  " << g << "[" << k << "] = __float_as_uint(ctx->fragCoord[" << k << "]);\n";
                continue; // other builtins are not mapped by the reference either (SPIRVCompiler.cpp:3659-3720)
            }
            if (!m.hasDeco(v.id, DecoLocation)) continue;
            const uint32_t location = m.decoVal(v.id, DecoLocation);
            if (vs) {
                if (pt.kind == Type::Array || pt.kind == Type::Matrix) { // consecutive locations, x2 for elements > 16 bytes (PipelineCompiler.cpp:924-944)
                    const uint32_t ew = T(pt.elem).words, mult = VariableSize(pt.elem) > 16 ? 2 : 1;
                    for (uint32_t j = 0; j < pt.count; j++) EmitFetch(o, location + j * mult, pt.elem, g + " + " + std::to_string(j * ew));
                } else EmitFetch(o, location, pointee, g);
            } else {
                const uint32_t size = VariableSize(pointee), word = inOff / 4;
                inOff += size;
                const bool flat = m.hasDeco(v.id, DecoFlat), linear = m.hasDeco(v.id, DecoNoPerspective);
                if (flat) { for (uint32_t k = 0; k < size / 4; k++) o << "  " << g << "[" << k << "] = cpvk_interp_flat(ctx, " << word + k << "u);\n"; continue; }
                const Type& et = pt.kind == Type::Vector ? T(pt.elem) : pt;
                if ((pt.kind != Type::Vector && pt.kind != Type::Float) || et.kind != Type::Float)
                    throw Unsupported("only 32-bit float scalars/vectors interpolate (FATAL_ERROR, Draw.cpp:863-869)");
                o << "  " << (linear ? "cpvk_interp_linear_vec" : "cpvk_interp_perspective_vec") << "(ctx, " << word << "u, " << pt.words << ", " << g << ");\n";
            }
        }
    }

    void Epilogue(bool vs) {
        if (vs) {
            body << "  cpvk_u32* rec_ = (cpvk_u32*)__builtin_assume_aligned(dp->vsOut + (cpvk_u64)rawId * dp->vsStride, 16);\n";
            auto outw = [&](uint32_t word, const std::string& val) { if (word >= 6) body << "  rec_[cpvk_vs_slot(" << word << "u)] = " << val << ";\n"; };
            // builtin block {vec4 position, float pointSize, float clip[1]} = words 0..5 (PipelineCompiler.cpp:532-538, :957-960)
            std::string pos[4] = {"0u", "0u", "0u", "0u"}, psz = "0u", clip = "0u";
            if (perVertexVar) {
                const uint32_t st = T(m.vars[m.varIndex.at(perVertexVar)].ptrType).elem; const Type& bt = T(st);
                uint32_t off = 0; const std::string g = "g" + std::to_string(perVertexVar);
                for (uint32_t k = 0; k < bt.members.size(); k++) {
                    uint32_t b = ~0u; m.memberDecoVal(st, k, DecoBuiltIn, b);
                    if (b == BiPosition) for (int q = 0; q < 4; q++) pos[q] = g + "[" + std::to_string(off + q) + "]";
                    else if (b == BiPointSize) psz = g + "[" + std::to_string(off) + "]";
                    else if (b == BiClipDistance && T(bt.members[k]).words) clip = g + "[" + std::to_string(off) + "]";
                    off += T(bt.members[k]).words;
                }
            }
            if (positionVar) for (int q = 0; q < 4; q++) pos[q] = "g" + std::to_string(positionVar) + "[" + std::to_string(q) + "]";
            if (pointSizeVar) psz = "g" + std::to_string(pointSizeVar) + "[0]";
            body << "  cpvk_store_position(dp, rawId, " << pos[0] << ", " << pos[1] << ", " << pos[2] << ", " << pos[3] << ");\n";
            if (desc.topology == 0) body << "  dp->vsPointSize[rawId] = " << psz << ";\n"; // point lists read the size back (Draw.cpp:1345)
            (void)clip; // clip distances have no consumer (no clipping)
            uint32_t byteOff = 24;
            std::function<void(uint32_t, const std::string&, uint32_t&, uint32_t)> store = [&](uint32_t ty, const std::string& g, uint32_t& srcWord, uint32_t dstByte) {
                const Type& t = T(ty);
                if (t.kind == Type::Array || t.kind == Type::Matrix) { const uint32_t es = AllocSize(t.elem); for (uint32_t k = 0; k < t.count; k++) store(t.elem, g, srcWord, dstByte + k * es); return; }
                for (uint32_t k = 0; k < t.words; k++) outw(dstByte / 4 + k, g + "[" + std::to_string(srcWord++) + "]");
                for (uint32_t k = t.words; k < AllocSize(ty) / 4; k++) outw(dstByte / 4 + k, "0u"); // <3 x float> tail padding
            };
            for (const Var& v : m.vars) {
                if (v.storage != ScOutput || !m.hasDeco(v.id, DecoLocation)) continue; // declaration order, not Location (F5)
                const uint32_t pointee = T(v.ptrType).elem; uint32_t srcWord = 0;
                store(pointee, "g" + std::to_string(v.id), srcWord, byteOff);
                byteOff += AllocSize(pointee);
            }
            layout.recordWords = byteOff / 4;
            layout.vsWritesMemory = storesToMemory;
        } else {
            layout.originUpperLeft = m.originUpperLeft;
            layout.fsWritesMemory = storesToMemory;
            // FindShaderLocations (PipelineCompiler.cpp:1729-1798): output at Location L feeds attachment L
            for (const Var& v : m.vars) {
                if (v.storage != ScOutput || !m.hasDeco(v.id, DecoLocation)) continue;
                const uint32_t pointee = T(v.ptrType).elem; const Type& pt = T(pointee);
                const uint32_t location = m.decoVal(v.id, DecoLocation);
                const bool multi = pt.kind == Type::Array || pt.kind == Type::Matrix;
                const uint32_t n = multi ? pt.count : 1, ew = multi ? T(pt.elem).words : pt.words;
                for (uint32_t j = 0; j < n; j++) {
                    if (location + j >= CPVK_MAX_COLOR_ATTACHMENTS) continue;
                    for (uint32_t k = 0; k < 4; k++)
                        body << "  out->color[" << location + j << "][" << k << "] = " << (k < ew ? "g" + std::to_string(v.id) + "[" + std::to_string(j * ew + k) + "]" : std::string("0u")) << ";\n";
                }
            }
        }
    }

    // ---- function bodies ----
    struct CallFrame { uint32_t resultId; int callerCtx; };

    void PhiCopies(int ctx, const Func& f, uint32_t fromLabel, uint32_t toLabel) {
        auto bi = f.blockIndex.find(toLabel);
        if (bi == f.blockIndex.end()) throw Malformed("branch to unknown label");
        const Block& tb = f.blocks[bi->second];
        std::vector<std::pair<std::string, std::string>> finals;
        for (const Inst& in : tb.insts) {
            if (in.op != OpPhi) break;
            const Type& t = T(in.type);
            for (uint32_t k = 0; k + 1 < in.nops; k += 2) {
                if (in.ops[k + 1] != fromLabel) continue;
                for (uint32_t wv = 0; wv < t.words; wv++) {
                    const std::string tmp = "p" + std::to_string(ctx) + "_" + std::to_string(in.result) + "_" + std::to_string(wv);
                    (t.leaves[wv] == 'f' ? declF : t.leaves[wv] == 'b' ? declB : declU).insert(tmp);
                    body << "  " << tmp << " = " << W(ctx, in.ops[k], wv) << ";\n";
                    finals.push_back({Dst(ctx, in, wv), tmp});
                }
                break;
            }
        }
        for (auto& fc : finals) body << "  " << fc.first << " = " << fc.second << ";\n";
    }
    std::string Label(int ctx, uint32_t id) const { return "L" + std::to_string(ctx) + "_" + std::to_string(id); }
    void Goto(int ctx, const Func& f, uint32_t from, uint32_t to) { PhiCopies(ctx, f, from, to); body << "  goto " << Label(ctx, to) << ";\n"; }

    void EmitFunction(uint32_t fnId, int ctx, const std::vector<uint32_t>& /*unused*/, uint32_t callResult, int callerCtx = 0, uint32_t callResultType = 0) {
        auto fi = m.funcs.find(fnId);
        if (fi == m.funcs.end()) throw Malformed("call to unknown function");
        const Func& f = fi->second;
        if (f.blocks.empty()) throw Malformed("function without body");
        for (const Block& b : f.blocks) {
            body << Label(ctx, b.label) << ": ;\n";
            for (const Inst& in : b.insts) EmitInst(ctx, f, b, in, callResult, callerCtx, callResultType);
        }
    }

    void Comp(int ctx, const Inst& in, const std::function<std::string(uint32_t)>& expr) {
        const Type& t = T(in.type);
        for (uint32_t k = 0; k < t.words; k++) body << "  " << Dst(ctx, in, k) << " = " << expr(k) << ";\n";
    }
    void Bin(int ctx, const Inst& in, const char* fmtPre, const char* mid, const char* post) {
        Comp(ctx, in, [&](uint32_t k) { return std::string(fmtPre) + W(ctx, in.ops[0], k) + mid + W(ctx, in.ops[1], k) + post; });
    }
    void SBin(int ctx, const Inst& in, const char* opstr) { // signed integer binary on unsigned storage
        Comp(ctx, in, [&](uint32_t k) { return "(unsigned)((int)" + W(ctx, in.ops[0], k) + " " + opstr + " (int)" + W(ctx, in.ops[1], k) + ")"; });
    }
    void Cmp(int ctx, const Inst& in, const char* opstr, bool isSigned, bool negate) {
        Comp(ctx, in, [&](uint32_t k) {
            std::string a = W(ctx, in.ops[0], k), b = W(ctx, in.ops[1], k);
            if (isSigned) { a = "(int)" + a; b = "(int)" + b; }
            return std::string(negate ? "!(" : "(") + a + " " + opstr + " " + b + ")";
        });
    }
    uint32_t Count(uint32_t valueId) const { const Type& t = T(TypeOf(valueId)); return t.kind == Type::Vector ? t.count : 1; }
    // glm::dot order: vec2 a+b, vec3 a+b+c, vec4 (a+b)+(c+d) (glm detail/func_geometric.inl compute_dot)
    std::string DotExpr(const std::vector<std::string>& a, const std::vector<std::string>& b) {
        std::vector<std::string> t; for (size_t i = 0; i < a.size(); i++) t.push_back("(" + a[i] + " * " + b[i] + ")");
        if (t.size() == 1) return t[0];
        if (t.size() == 2) return "(" + t[0] + " + " + t[1] + ")";
        if (t.size() == 3) return "((" + t[0] + " + " + t[1] + ") + " + t[2] + ")";
        return "((" + t[0] + " + " + t[1] + ") + (" + t[2] + " + " + t[3] + "))";
    }
    std::vector<std::string> Words(int ctx, uint32_t id, uint32_t first, uint32_t n) { std::vector<std::string> r; for (uint32_t k = 0; k < n; k++) r.push_back(W(ctx, id, first + k)); return r; }
    std::string TempF(const std::string& expr) { const std::string n = "t" + std::to_string(tmpCounter++); declF.insert(n); body << "  " << n << " = " << expr << ";\n"; return n; }

    void EmitExt(int ctx, const Inst& in) {
        if (in.ops[0] != m.glsl) throw Unsupported("extended instruction set other than GLSL.std.450");
        const uint32_t e = in.ops[1];
        auto A = [&](uint32_t i, uint32_t k) { return W(ctx, in.ops[2 + i], k); };
        auto call1 = [&](const char* fn) { Comp(ctx, in, [&](uint32_t k) { return std::string(fn) + "(" + A(0, k) + ")"; }); };
        auto call2 = [&](const char* fn) { Comp(ctx, in, [&](uint32_t k) { return std::string(fn) + "(" + A(0, k) + ", " + A(1, k) + ")"; }); };
        auto call3 = [&](const char* fn) { Comp(ctx, in, [&](uint32_t k) { return std::string(fn) + "(" + A(0, k) + ", " + A(1, k) + ", " + A(2, k) + ")"; }); };
        auto scall2 = [&](const char* fn) { Comp(ctx, in, [&](uint32_t k) { return std::string("(unsigned)") + fn + "((int)" + A(0, k) + ", (int)" + A(1, k) + ")"; }); };
        switch (e) {
        case 4: call1("fabsf"); break;
        case 5: Comp(ctx, in, [&](uint32_t k) { return "(unsigned)((int)" + A(0, k) + " < 0 ? -(int)" + A(0, k) + " : (int)" + A(0, k) + ")"; }); break;
        case 6: call1("cpvk_fsign"); break;
        case 7: Comp(ctx, in, [&](uint32_t k) { return "(unsigned)cpvk_ssign((int)" + A(0, k) + ")"; }); break;
        case 8: call1("floorf"); break;
        case 9: call1("ceilf"); break;
        case 10: Comp(ctx, in, [&](uint32_t k) { return "(" + A(0, k) + " - floorf(" + A(0, k) + "))"; }); break;
        case 13: call1("sinf"); break; case 14: call1("cosf"); break;
        case 26: call2("powf"); break; case 27: call1("expf"); break; case 28: call1("logf"); break;
        case 29: call1("exp2f"); break; case 30: call1("log2f"); break; case 31: call1("sqrtf"); break;
        case 32: Comp(ctx, in, [&](uint32_t k) { return "(1.0f / sqrtf(" + A(0, k) + "))"; }); break;
        case 37: call2("cpvk_fmin"); break; case 38: call2("cpvk_umin"); break; case 39: scall2("cpvk_smin"); break;
        case 40: call2("cpvk_fmax"); break; case 41: call2("cpvk_umax"); break; case 42: scall2("cpvk_smax"); break;
        case 43: call3("cpvk_clampf"); break; case 44: call3("cpvk_uclamp"); break;
        case 45: Comp(ctx, in, [&](uint32_t k) { return "(unsigned)cpvk_sclamp((int)" + A(0, k) + ", (int)" + A(1, k) + ", (int)" + A(2, k) + ")"; }); break;
        case 46: call3("cpvk_fmix"); break;
        case 79: call2("cpvk_nmin"); break; case 80: call2("cpvk_nmax"); break;
        case 81: Comp(ctx, in, [&](uint32_t k) { return "cpvk_nmin(cpvk_nmax(" + A(0, k) + ", " + A(1, k) + "), " + A(2, k) + ")"; }); break;
        case 66: { const uint32_t n = Count(in.ops[2]); auto a = Words(ctx, in.ops[2], 0, n); body << "  " << Dst(ctx, in, 0) << " = sqrtf(" << DotExpr(a, a) << ");\n"; break; }
        case 67: { const uint32_t n = Count(in.ops[2]); std::vector<std::string> d; for (uint32_t k = 0; k < n; k++) d.push_back(TempF(A(1, k) + " - " + A(0, k)));
                   body << "  " << Dst(ctx, in, 0) << " = sqrtf(" << DotExpr(d, d) << ");\n"; break; }
        case 69: { const uint32_t n = Count(in.ops[2]); auto a = Words(ctx, in.ops[2], 0, n); const std::string inv = TempF("1.0f / sqrtf(" + DotExpr(a, a) + ")");
                   for (uint32_t k = 0; k < n; k++) body << "  " << Dst(ctx, in, k) << " = " << a[k] << " * " << inv << ";\n"; break; }
        case 71: { const uint32_t n = Count(in.ops[2]); auto i = Words(ctx, in.ops[2], 0, n), nn = Words(ctx, in.ops[3], 0, n); const std::string d = TempF(DotExpr(nn, i));
                   for (uint32_t k = 0; k < n; k++) { const std::string t1 = TempF(nn[k] + " * " + d); const std::string t2 = TempF(t1 + " * 2.0f"); body << "  " << Dst(ctx, in, k) << " = " << i[k] << " - " << t2 << ";\n"; } break; }
        case 68: { auto a = Words(ctx, in.ops[2], 0, 3), b = Words(ctx, in.ops[3], 0, 3);
                   const int ia[3] = {1, 2, 0}, ib[3] = {2, 0, 1};
                   for (int k = 0; k < 3; k++) { const std::string p = TempF(a[ia[k]] + " * " + b[ib[k]]), q = TempF(b[ia[k]] + " * " + a[ib[k]]); body << "  " << Dst(ctx, in, k) << " = " << p << " - " << q << ";\n"; } break; }
        default: throw Unsupported("GLSL.std.450 instruction " + std::to_string(e));
        }
    }

    void EmitInst(int ctx, const Func& f, const Block& b, const Inst& in, uint32_t callResult, int callerCtx, uint32_t callResultType) {
        switch (in.op) {
        case 0: case 8: case 317: case OpLoopMerge: case OpSelectionMerge: case OpPhi: break;
        case OpUndef: Comp(ctx, in, [&](uint32_t k) { return Lit(T(in.type).leaves[k], 0); }); break;
        case OpVariable: {
            const uint32_t pointee = T(in.type).elem; const Type& pt = T(pointee);
            Ptr p; p.kind = Ptr::Local; p.type = pointee; p.base = "l" + std::to_string(ctx) + "_" + std::to_string(in.result);
            arrays.push_back("unsigned " + p.base + "[" + std::to_string(pt.words ? pt.words : 1) + "];");
            ptrs[{ctx, in.result}] = p;
            if (in.nops > 1) EmitStore(ctx, p, in.ops[1]);
            break; }
        case OpLoad: EmitLoad(ctx, in, VarPtr(ctx, in.ops[0])); break;
        case OpStore: EmitStore(ctx, VarPtr(ctx, in.ops[0]), in.ops[1]); break;
        case OpAccessChain: case OpInBoundsAccessChain: ptrs[{ctx, in.result}] = AccessChain(ctx, in); break;
        case OpFunctionCall: {
            const int nctx = ++ctxCounter;
            auto fi = m.funcs.find(in.ops[0]); if (fi == m.funcs.end()) throw Malformed("call to unknown function");
            const Func& callee = fi->second;
            if (callee.params.size() != in.nops - 1) throw Malformed("argument count mismatch");
            for (size_t a = 0; a < callee.params.size(); a++) {
                const uint32_t pid = callee.params[a], aid = in.ops[1 + a];
                const Type& pt = T(TypeOf(pid));
                if (pt.kind == Type::Pointer) ptrs[{nctx, pid}] = VarPtr(ctx, aid);
                else if (pt.kind == Type::Image || pt.kind == Type::SampledImage || pt.kind == Type::Sampler) CopyHandle({nctx, pid}, {ctx, aid});
                else for (uint32_t k = 0; k < pt.words; k++) body << "  " << Name(nctx, pid, k, pt.leaves[k]) << " = " << W(ctx, aid, k) << ";\n";
            }
            EmitFunction(in.ops[0], nctx, {}, in.result, ctx, in.type);
            body << "R" << nctx << ": ;\n";
            break; }
        case OpReturn: body << "  goto " << (ctx == 0 ? std::string("L_end") : "R" + std::to_string(ctx)) << ";\n"; break;
        case OpReturnValue: {
            if (ctx == 0) throw Malformed("entry point returns a value");
            const Type& t = T(callResultType);
            for (uint32_t k = 0; k < t.words; k++) body << "  " << Name(callerCtx, callResult, k, t.leaves[k]) << " = " << W(ctx, in.ops[0], k) << ";\n";
            body << "  goto R" << ctx << ";\n"; break; }
        case OpKill: body << "  discard_ = true; goto L_end;\n"; break; // "return true" from the fragment entry (SPIRVCompiler.cpp:3308-3310)
        case OpUnreachable: body << "  goto L_end;\n"; break;
        case OpBranch: Goto(ctx, f, b.label, in.ops[0]); break;
        case OpBranchConditional:
            body << "  if (" << W(ctx, in.ops[0], 0) << ") {\n"; Goto(ctx, f, b.label, in.ops[1]);
            body << "  } else {\n"; Goto(ctx, f, b.label, in.ops[2]); body << "  }\n"; break;
        case OpSwitch:
            body << "  switch (" << W(ctx, in.ops[0], 0) << ") {\n";
            for (uint32_t k = 2; k + 1 < in.nops; k += 2) { body << "  case " << Hex(in.ops[k]) << ": {\n"; Goto(ctx, f, b.label, in.ops[k + 1]); body << "  }\n"; }
            body << "  default: {\n"; Goto(ctx, f, b.label, in.ops[1]); body << "  }\n  }\n"; break;
        case OpCopyObject:
            if (T(in.type).kind == Type::Pointer) ptrs[{ctx, in.result}] = VarPtr(ctx, in.ops[0]);
            else if (handles.count({ctx, in.ops[0]})) CopyHandle({ctx, in.result}, {ctx, in.ops[0]});
            else Comp(ctx, in, [&](uint32_t k) { return W(ctx, in.ops[0], k); });
            break;
        case OpVectorShuffle: {
            const uint32_t na = Count(in.ops[0]);
            Comp(ctx, in, [&](uint32_t k) { const uint32_t s = in.ops[2 + k]; return s == 0xFFFFFFFFu ? Lit(T(in.type).leaves[k], 0) : (s < na ? W(ctx, in.ops[0], s) : W(ctx, in.ops[1], s - na)); });
            break; }
        case OpCompositeConstruct: {
            uint32_t k = 0;
            for (uint32_t a = 0; a < in.nops; a++) { const uint32_t wn = T(TypeOf(in.ops[a])).words; for (uint32_t q = 0; q < wn; q++, k++) body << "  " << Dst(ctx, in, k) << " = " << W(ctx, in.ops[a], q) << ";\n"; }
            break; }
        case OpCompositeExtract: case OpCompositeInsert: {
            const bool ins = in.op == OpCompositeInsert;
            const uint32_t comp = ins ? in.ops[1] : in.ops[0];
            uint32_t ty = TypeOf(comp), off = 0;
            for (uint32_t k = ins ? 2 : 1; k < in.nops; k++) {
                const uint32_t idx = in.ops[k]; const Type& t = T(ty);
                if (t.kind == Type::Struct) { if (idx >= t.members.size()) throw Malformed("bad composite index"); for (uint32_t q = 0; q < idx; q++) off += T(t.members[q]).words; ty = t.members[idx]; }
                else { off += idx * T(t.elem).words; ty = t.elem; }
            }
            const uint32_t pw = T(ty).words;
            if (ins) Comp(ctx, in, [&](uint32_t k) { return (k >= off && k < off + pw) ? W(ctx, in.ops[0], k - off) : W(ctx, comp, k); });
            else Comp(ctx, in, [&](uint32_t k) { return W(ctx, comp, off + k); });
            break; }
        case OpVectorExtractDynamic: {
            const uint32_t n = Count(in.ops[0]); std::string e = W(ctx, in.ops[0], n - 1);
            for (int k = (int)n - 2; k >= 0; k--) e = "(" + W(ctx, in.ops[1], 0) + " == " + std::to_string(k) + "u ? " + W(ctx, in.ops[0], k) + " : " + e + ")";
            body << "  " << Dst(ctx, in, 0) << " = " << e << ";\n"; break; }
        case OpVectorInsertDynamic:
            Comp(ctx, in, [&](uint32_t k) { return "(" + W(ctx, in.ops[2], 0) + " == " + std::to_string(k) + "u ? " + W(ctx, in.ops[1], 0) + " : " + W(ctx, in.ops[0], k) + ")"; }); break;
        case OpTranspose: {
            const Type& mt = T(TypeOf(in.ops[0])); const uint32_t cols = mt.count, rows = T(mt.elem).count;
            Comp(ctx, in, [&](uint32_t k) { const uint32_t q = k / cols, c = k % cols; return W(ctx, in.ops[0], c * rows + q); }); break; }
        case OpFNegate: Comp(ctx, in, [&](uint32_t k) { return "-" + W(ctx, in.ops[0], k); }); break;
        case OpSNegate: Comp(ctx, in, [&](uint32_t k) { return "(0u - " + W(ctx, in.ops[0], k) + ")"; }); break;
        case OpFAdd: case OpIAdd: Bin(ctx, in, "", " + ", ""); break;
        case OpFSub: case OpISub: Bin(ctx, in, "", " - ", ""); break;
        case OpFMul: case OpIMul: Bin(ctx, in, "", " * ", ""); break;
        case OpFDiv: Bin(ctx, in, "", " / ", ""); break;
        case OpFRem: Bin(ctx, in, "fmodf(", ", ", ")"); break;
        case OpFMod: Bin(ctx, in, "cpvk_fmod_glsl(", ", ", ")"); break;
        case OpUDiv: Bin(ctx, in, "cpvk_udiv(", ", ", ")"); break;
        case OpUMod: Bin(ctx, in, "cpvk_umod(", ", ", ")"); break;
        case OpSDiv: Comp(ctx, in, [&](uint32_t k) { return "(unsigned)cpvk_sdiv((int)" + W(ctx, in.ops[0], k) + ", (int)" + W(ctx, in.ops[1], k) + ")"; }); break;
        case OpSRem: Comp(ctx, in, [&](uint32_t k) { return "(unsigned)cpvk_srem((int)" + W(ctx, in.ops[0], k) + ", (int)" + W(ctx, in.ops[1], k) + ")"; }); break;
        case OpSMod: Comp(ctx, in, [&](uint32_t k) { return "(unsigned)cpvk_smod((int)" + W(ctx, in.ops[0], k) + ", (int)" + W(ctx, in.ops[1], k) + ")"; }); break;
        case OpVectorTimesScalar: case OpMatrixTimesScalar: Comp(ctx, in, [&](uint32_t k) { return W(ctx, in.ops[0], k) + " * " + W(ctx, in.ops[1], 0); }); break;
        case OpMatrixTimesVector: { // glm mat*vec: 4x4 -> (m0*v0 + m1*v1) + (m2*v2 + m3*v3); otherwise left to right (SpirvFunctions.cpp:37-47)
            const Type& mt = T(TypeOf(in.ops[0])); const uint32_t cols = mt.count, rows = T(mt.elem).count;
            for (uint32_t q = 0; q < rows; q++) {
                std::vector<std::string> pr; for (uint32_t c = 0; c < cols; c++) pr.push_back("(" + W(ctx, in.ops[0], c * rows + q) + " * " + W(ctx, in.ops[1], c) + ")");
                std::string e;
                if (cols == 4 && rows == 4) e = "((" + pr[0] + " + " + pr[1] + ") + (" + pr[2] + " + " + pr[3] + "))";
                else { e = pr[0]; for (uint32_t c = 1; c < cols; c++) e = "(" + e + " + " + pr[c] + ")"; }
                body << "  " << Dst(ctx, in, q) << " = " << e << ";\n";
            }
            break; }
        case OpVectorTimesMatrix: { // glm vec*mat: per column, left-to-right sum over rows (SpirvFunctions.cpp:24-35)
            const Type& mt = T(TypeOf(in.ops[1])); const uint32_t cols = mt.count, rows = T(mt.elem).count;
            for (uint32_t c = 0; c < cols; c++) {
                std::string e = "(" + W(ctx, in.ops[1], c * rows) + " * " + W(ctx, in.ops[0], 0) + ")";
                for (uint32_t k = 1; k < rows; k++) e = "(" + e + " + (" + W(ctx, in.ops[1], c * rows + k) + " * " + W(ctx, in.ops[0], k) + "))";
                body << "  " << Dst(ctx, in, c) << " = " << e << ";\n";
            }
            break; }
        case OpMatrixTimesMatrix: { // glm mat*mat: Result[j] = A0*B[j][0] + A1*B[j][1] + ... left to right (SpirvFunctions.cpp:49-60)
            const Type& lt = T(TypeOf(in.ops[0])); const uint32_t lcols = lt.count, lrows = T(lt.elem).count; const uint32_t rcols = T(TypeOf(in.ops[1])).count;
            for (uint32_t j = 0; j < rcols; j++) for (uint32_t q = 0; q < lrows; q++) {
                std::string e = "(" + W(ctx, in.ops[0], q) + " * " + W(ctx, in.ops[1], j * lcols) + ")";
                for (uint32_t k = 1; k < lcols; k++) e = "(" + e + " + (" + W(ctx, in.ops[0], k * lrows + q) + " * " + W(ctx, in.ops[1], j * lcols + k) + "))";
                body << "  " << Dst(ctx, in, j * lrows + q) << " = " << e << ";\n";
            }
            break; }
        case OpDot: { const uint32_t n = Count(in.ops[0]); body << "  " << Dst(ctx, in, 0) << " = " << DotExpr(Words(ctx, in.ops[0], 0, n), Words(ctx, in.ops[1], 0, n)) << ";\n"; break; }
        case OpConvertFToU: Comp(ctx, in, [&](uint32_t k) { return "cpvk_f2u(" + W(ctx, in.ops[0], k) + ")"; }); break;
        case OpConvertFToS: Comp(ctx, in, [&](uint32_t k) { return "(unsigned)cpvk_f2s(" + W(ctx, in.ops[0], k) + ")"; }); break;
        case OpConvertSToF: Comp(ctx, in, [&](uint32_t k) { return "(float)(int)" + W(ctx, in.ops[0], k); }); break;
        case OpConvertUToF: Comp(ctx, in, [&](uint32_t k) { return "(float)" + W(ctx, in.ops[0], k); }); break;
        case OpBitcast: {
            const std::string& sl = T(TypeOf(in.ops[0])).leaves; const std::string& dl = T(in.type).leaves;
            if (sl.size() != dl.size()) throw Unsupported("bitcast changing component count");
            Comp(ctx, in, [&](uint32_t k) { return FromWord(dl[k], ToWord(sl[k], W(ctx, in.ops[0], k))); }); break; }
        case OpAny: case OpAll: { const uint32_t n = Count(in.ops[0]); std::string e = W(ctx, in.ops[0], 0);
            for (uint32_t k = 1; k < n; k++) e += std::string(in.op == OpAny ? " || " : " && ") + W(ctx, in.ops[0], k);
            body << "  " << Dst(ctx, in, 0) << " = (" << e << ");\n"; break; }
        case OpIsNan: Comp(ctx, in, [&](uint32_t k) { return "cpvk_isnan(" + W(ctx, in.ops[0], k) + ")"; }); break;
        case OpIsInf: Comp(ctx, in, [&](uint32_t k) { return "(fabsf(" + W(ctx, in.ops[0], k) + ") == __uint_as_float(0x7f800000u))"; }); break;
        case OpLogicalEqual: Bin(ctx, in, "(", " == ", ")"); break;
        case OpLogicalNotEqual: Bin(ctx, in, "(", " != ", ")"); break;
        case OpLogicalOr: Bin(ctx, in, "(", " || ", ")"); break;
        case OpLogicalAnd: Bin(ctx, in, "(", " && ", ")"); break;
        case OpLogicalNot: Comp(ctx, in, [&](uint32_t k) { return "!" + W(ctx, in.ops[0], k); }); break;
        case OpSelect: { const bool vc = T(TypeOf(in.ops[0])).kind == Type::Vector;
            Comp(ctx, in, [&](uint32_t k) { return "(" + W(ctx, in.ops[0], vc ? k : 0) + " ? " + W(ctx, in.ops[1], k) + " : " + W(ctx, in.ops[2], k) + ")"; }); break; }
        case OpIEqual: Cmp(ctx, in, "==", false, false); break;
        case OpINotEqual: Cmp(ctx, in, "!=", false, false); break;
        case OpUGreaterThan: Cmp(ctx, in, ">", false, false); break;
        case OpSGreaterThan: Cmp(ctx, in, ">", true, false); break;
        case OpUGreaterThanEqual: Cmp(ctx, in, ">=", false, false); break;
        case OpSGreaterThanEqual: Cmp(ctx, in, ">=", true, false); break;
        case OpULessThan: Cmp(ctx, in, "<", false, false); break;
        case OpSLessThan: Cmp(ctx, in, "<", true, false); break;
        case OpULessThanEqual: Cmp(ctx, in, "<=", false, false); break;
        case OpSLessThanEqual: Cmp(ctx, in, "<=", true, false); break;
        case OpFOrdEqual: Cmp(ctx, in, "==", false, false); break;
        case OpFUnordNotEqual: Cmp(ctx, in, "!=", false, false); break;
        case OpFOrdLessThan: Cmp(ctx, in, "<", false, false); break;
        case OpFOrdGreaterThan: Cmp(ctx, in, ">", false, false); break;
        case OpFOrdLessThanEqual: Cmp(ctx, in, "<=", false, false); break;
        case OpFOrdGreaterThanEqual: Cmp(ctx, in, ">=", false, false); break;
        case OpFUnordLessThan: Cmp(ctx, in, ">=", false, true); break;
        case OpFUnordGreaterThan: Cmp(ctx, in, "<=", false, true); break;
        case OpFUnordLessThanEqual: Cmp(ctx, in, ">", false, true); break;
        case OpFUnordGreaterThanEqual: Cmp(ctx, in, "<", false, true); break;
        case OpFOrdNotEqual: Comp(ctx, in, [&](uint32_t k) { const std::string a = W(ctx, in.ops[0], k), c = W(ctx, in.ops[1], k); return "(" + a + " < " + c + " || " + a + " > " + c + ")"; }); break;
        case OpFUnordEqual: Comp(ctx, in, [&](uint32_t k) { const std::string a = W(ctx, in.ops[0], k), c = W(ctx, in.ops[1], k); return "!(" + a + " < " + c + " || " + a + " > " + c + ")"; }); break;
        case OpShiftRightLogical: Comp(ctx, in, [&](uint32_t k) { return "(" + W(ctx, in.ops[0], k) + " >> (" + W(ctx, in.ops[1], k) + " & 31u))"; }); break;
        case OpShiftRightArithmetic: Comp(ctx, in, [&](uint32_t k) { return "(unsigned)((int)" + W(ctx, in.ops[0], k) + " >> (" + W(ctx, in.ops[1], k) + " & 31u))"; }); break;
        case OpShiftLeftLogical: Comp(ctx, in, [&](uint32_t k) { return "(" + W(ctx, in.ops[0], k) + " << (" + W(ctx, in.ops[1], k) + " & 31u))"; }); break;
        case OpBitwiseOr: Bin(ctx, in, "(", " | ", ")"); break;
        case OpBitwiseXor: Bin(ctx, in, "(", " ^ ", ")"); break;
        case OpBitwiseAnd: Bin(ctx, in, "(", " & ", ")"); break;
        case OpNot: Comp(ctx, in, [&](uint32_t k) { return "~" + W(ctx, in.ops[0], k); }); break;
        case OpExtInst: EmitExt(ctx, in); break;
        case OpSampledImage: { // @Image.Combine (GlslFunctions.cpp:812-820): the image's data with the sampler object's state
            auto hi = handles.find({ctx, in.ops[0]}), hs = handles.find({ctx, in.ops[1]});
            if (hi == handles.end() || hs == handles.end()) throw Malformed("OpSampledImage operands are not loaded handles");
            handles[{ctx, in.result}] = hi->second; samplerOf[{ctx, in.result}] = hs->second; break; }
        case OpImage: handles[{ctx, in.result}] = handles.at({ctx, in.ops[0]}); break;
        case OpImageSampleImplicitLod: case OpImageSampleExplicitLod: {
            auto h = handles.find({ctx, in.ops[0]}); if (h == handles.end()) throw Malformed("sampled image operand is not a loaded handle");
            std::string lod = "0.0f";
            if (in.op == OpImageSampleExplicitLod) { if (in.nops < 4 || in.ops[2] != 2) throw Unsupported("image operands other than Lod (SPIRVCompiler.cpp:1607-1615)"); lod = W(ctx, in.ops[3], 0); }
            else if (in.nops > 2) throw Unsupported("image operands on an implicit-LOD sample");
            const uint32_t cn = T(TypeOf(in.ops[1])).words;
            const std::string t = "s" + std::to_string(tmpCounter++);
            declV.insert(t);
            if (model == 4) layout.fsSamplesImages = true;
            int dimsHint = 0; // OpTypeImage Dim: 1D, 2D, 3D -> 1, 2, 3
            { const Type& st = T(TypeOf(in.ops[0])); const Type& it = st.kind == Type::SampledImage ? T(st.elem) : st; if (it.kind == Type::Image && it.count <= 2) dimsHint = (int)it.count + 1; }
            const auto so = samplerOf.find({ctx, in.ops[0]});
            body << "  " << t << " = cpvk_image_sample(" << h->second << ", " << (so != samplerOf.end() ? so->second : h->second) << ", " << W(ctx, in.ops[1], 0) << ", " << (cn > 1 ? W(ctx, in.ops[1], 1) : "0.0f") << ", "
                 << (cn > 2 ? W(ctx, in.ops[1], 2) : "0.0f") << ", " << lod << ", lut_, " << dimsHint << ");\n";
            if (T(T(in.type).kind == Type::Vector ? T(in.type).elem : in.type).kind != Type::Float) throw Unsupported("integer image sampling");
            Comp(ctx, in, [&](uint32_t k) { return t + ".v[" + std::to_string(k) + "]"; });
            break; }
        // OpImageRead is @Image.Read = ImageFetch with the coordinate as written (GlslFunctions.cpp:739-743) — for
        // subpassLoad() glslang writes ivec2(0, 0), so the reference reads texel (0, 0) of the input attachment for every
        // fragment; kept as is for parity.
        case OpImageFetch: case OpImageRead: {
            auto h = handles.find({ctx, in.ops[0]}); if (h == handles.end()) throw Malformed("image operand is not a loaded handle");
            if (in.nops > 2) throw Unsupported("image operands on OpImageFetch / OpImageRead (TODO_ERROR, SPIRVCompiler.cpp:1660-1663)");
            const uint32_t cn = T(TypeOf(in.ops[1])).words;
            const std::string t = "s" + std::to_string(tmpCounter++);
            declV.insert(t);
            if (model == 4) layout.fsSamplesImages = true;
            body << "  " << t << " = cpvk_image_fetch(" << h->second << ", (int)" << W(ctx, in.ops[1], 0) << ", " << (cn > 1 ? "(int)" + W(ctx, in.ops[1], 1) : "0") << ", "
                 << (cn > 2 ? "(int)" + W(ctx, in.ops[1], 2) : "0") << ", lut_);\n";
            if (T(T(in.type).kind == Type::Vector ? T(in.type).elem : in.type).kind != Type::Float) throw Unsupported("integer image fetch");
            Comp(ctx, in, [&](uint32_t k) { return t + ".v[" + std::to_string(k) + "]"; });
            break; }
        default: throw Unsupported("SPIR-V opcode " + std::to_string(in.op));
        }
    }
};

} // namespace

int TranslateStage(const CpvkShaderStage& stage, uint32_t model, const CpvkPipelineDesc& desc, PipelineLayoutInfo& layout, std::string& out, std::string& error) {
    try {
        Translator t(stage, model, desc, layout);
        out += t.Run();
        return 0;
    } catch (const Unsupported& e) { error = std::string("unsupported: ") + e.what(); return CPVK_E_UNSUPPORTED; }
    catch (const Malformed& e) { error = std::string("malformed SPIR-V: ") + e.what(); return CPVK_E_SPIRV; }
    catch (const std::exception& e) { error = e.what(); return CPVK_E_SPIRV; }
}

} // namespace cpvk
