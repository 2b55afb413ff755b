// oracle_spirv.h — a small SPIR-V interpreter: the oracle's stand-in for the reference's SPIR-V -> LLVM-IR
// lowering + x86 JIT, executing one shader invocation at a time exactly like the reference does.
// TEST INFRASTRUCTURE ONLY (see oracle_formats.h). PARITY UNPINNED by reference tests.
//
// Follows the lowering contract of
//   LLVMRuntime/SPIRVCompiler.cpp: instruction semantics :1938-3320 (plain IEEE ops, no fast-math flags
//     :3548-3578), OpKill -> "return true" :3308-3310, OpReturn -> "return false" in a fragment entry point
//     :3642-3653, variables as zero-initialised module globals :1218-1271, builtins :3659-3720
//   LLVMRuntime/SpirvFunctions.cpp:6-82 (glm dot / matrix products; operand order per glm, e.g.
//     Samples/utils/glm/detail/type_mat4x4.inl:676-687 for mat4*vec4)
//   CPVulkan/GlslFunctions.cpp:19-321 (GLSL.std.450 subset), :598-737 (image sample / fetch)
// Deliberate simplification (documented divergence): private/output globals are re-zeroed per invocation,
// whereas the reference's persist between invocations; shaders that read an output before
// writing it are invocation-order dependent in the reference and are not covered.
// Buffer-backed (Uniform / PushConstant / StorageBuffer) data is addressed through the explicit Offset /
// ArrayStride / MatrixStride decorations (std140/std430 as written by the front end).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "oracle_sampler.h"

namespace oracle {
namespace spv {

enum Op : uint16_t {
    OpNop = 0, OpUndef = 1, OpName = 5, OpExtInstImport = 11, OpExtInst = 12, OpMemoryModel = 14, OpEntryPoint = 15,
    OpExecutionMode = 16, OpCapability = 17, OpTypeVoid = 19, OpTypeBool = 20, OpTypeInt = 21, OpTypeFloat = 22,
    OpTypeVector = 23, OpTypeMatrix = 24, OpTypeImage = 25, OpTypeSampler = 26, OpTypeSampledImage = 27,
    OpTypeArray = 28, OpTypeRuntimeArray = 29, OpTypeStruct = 30, OpTypePointer = 32, OpTypeFunction = 33,
    OpConstantTrue = 41, OpConstantFalse = 42, OpConstant = 43, OpConstantComposite = 44, OpConstantNull = 46,
    OpSpecConstantTrue = 48, OpSpecConstantFalse = 49, OpSpecConstant = 50, OpSpecConstantComposite = 51,
    OpFunction = 54, OpFunctionParameter = 55, OpFunctionEnd = 56, OpFunctionCall = 57, OpVariable = 59,
    OpLoad = 61, OpStore = 62, OpAccessChain = 65, OpInBoundsAccessChain = 66, OpDecorate = 71,
    OpMemberDecorate = 72, OpVectorExtractDynamic = 77, OpVectorInsertDynamic = 78, OpVectorShuffle = 79,
    OpCompositeConstruct = 80, OpCompositeExtract = 81, OpCompositeInsert = 82, OpCopyObject = 83,
    OpTranspose = 84, OpSampledImage = 86, OpImageSampleImplicitLod = 87, OpImageSampleExplicitLod = 88,
    OpImageFetch = 95, OpImageRead = 98, OpImage = 100, OpConvertFToU = 109, OpConvertFToS = 110, OpConvertSToF = 111,
    OpConvertUToF = 112, OpBitcast = 124, OpSNegate = 126, OpFNegate = 127, OpIAdd = 128, OpFAdd = 129,
    OpISub = 130, OpFSub = 131, OpIMul = 132, OpFMul = 133, OpUDiv = 134, OpSDiv = 135, OpFDiv = 136,
    OpUMod = 137, OpSRem = 138, OpSMod = 139, OpFRem = 140, OpFMod = 141, OpVectorTimesScalar = 142,
    OpMatrixTimesScalar = 143, OpVectorTimesMatrix = 144, OpMatrixTimesVector = 145, OpMatrixTimesMatrix = 146,
    OpDot = 148, OpAny = 154, OpAll = 155, OpIsNan = 156, OpIsInf = 157, OpLogicalEqual = 164,
    OpLogicalNotEqual = 165, OpLogicalOr = 166, OpLogicalAnd = 167, OpLogicalNot = 168, OpSelect = 169,
    OpIEqual = 170, OpINotEqual = 171, OpUGreaterThan = 172, OpSGreaterThan = 173, OpUGreaterThanEqual = 174,
    OpSGreaterThanEqual = 175, OpULessThan = 176, OpSLessThan = 177, OpULessThanEqual = 178,
    OpSLessThanEqual = 179, OpFOrdEqual = 180, OpFUnordEqual = 181, OpFOrdNotEqual = 182, OpFUnordNotEqual = 183,
    OpFOrdLessThan = 184, OpFUnordLessThan = 185, OpFOrdGreaterThan = 186, OpFUnordGreaterThan = 187,
    OpFOrdLessThanEqual = 188, OpFUnordLessThanEqual = 189, OpFOrdGreaterThanEqual = 190,
    OpFUnordGreaterThanEqual = 191, OpShiftRightLogical = 194, OpShiftRightArithmetic = 195,
    OpShiftLeftLogical = 196, OpBitwiseOr = 197, OpBitwiseXor = 198, OpBitwiseAnd = 199, OpNot = 200,
    OpPhi = 245, OpLoopMerge = 246, OpSelectionMerge = 247, OpLabel = 248, OpBranch = 249,
    OpBranchConditional = 250, OpSwitch = 251, OpKill = 252, OpReturn = 253, OpReturnValue = 254,
    OpUnreachable = 255,
};

enum Deco { DecoSpecId = 1, DecoBlock = 2, DecoRowMajor = 4, DecoColMajor = 5, DecoArrayStride = 6, DecoMatrixStride = 7,
            DecoBuiltIn = 11, DecoNoPerspective = 13, DecoFlat = 14, DecoLocation = 30, DecoBinding = 33,
            DecoDescriptorSet = 34, DecoOffset = 35 };
enum Storage { ScUniformConstant = 0, ScInput = 1, ScUniform = 2, ScOutput = 3, ScPrivate = 6, ScFunction = 7,
               ScPushConstant = 9, ScStorageBuffer = 12 };
enum BuiltIn { BiPosition = 0, BiPointSize = 1, BiClipDistance = 3, BiVertexId = 5, BiInstanceId = 6, BiFragCoord = 15,
               BiVertexIndex = 42, BiInstanceIndex = 43 };

struct Type {
    enum Kind { Void, Bool, Int, Float, Vector, Matrix, Array, RuntimeArray, Struct, Pointer, Function, Image, Sampler, SampledImage } kind = Void;
    uint32_t width = 0;
    bool isSigned = false;
    uint32_t elem = 0;      // element / column / pointee / image type id
    uint32_t count = 0;     // components / columns / array length
    uint32_t storage = 0;   // pointers
    uint32_t dim = 0;       // images
    std::vector<uint32_t> members;
    uint32_t words = 0;     // logical size in 32-bit words
};

// A pointer value as held in a register slot (4 words).
struct Ptr {
    uint8_t* addr;       // host address (register file word, or external buffer byte)
    uint32_t type;       // pointee type id
    uint16_t buffer;     // 1: external memory with explicit layout
    uint16_t matStride;  // inherited MatrixStride for buffer pointers into/at a matrix
};
static_assert(sizeof(Ptr) == 16, "Ptr must occupy 4 words");

struct Inst {
    uint16_t op;
    uint16_t nops;          // operand words after type/result
    uint32_t type, result;
    const uint32_t* ops;
};

struct Block {
    uint32_t label;
    std::vector<Inst> insts;
};

struct Function {
    uint32_t id = 0, retType = 0;
    std::vector<uint32_t> params;
    std::vector<Block> blocks;
    std::unordered_map<uint32_t, uint32_t> blockIndex;
    uint32_t frameWords = 0;
    std::unordered_map<uint32_t, uint32_t> slot;           // result id -> word offset in frame
    std::vector<std::pair<uint32_t, uint32_t>> localVars;  // (variable id, storage slot)
};

struct Variable {
    uint32_t id, ptrType, storage, initializer;
    uint32_t slot;  // word offset in globals of the *pointer* value; storage follows for logical classes
    uint32_t dataSlot;
};

struct Module {
    std::vector<uint32_t> words;
    uint32_t bound = 0;
    std::vector<Type> types;                                  // indexed by id
    std::vector<std::map<uint32_t, std::vector<uint32_t>>> deco;                     // id -> decoration -> operands
    std::map<std::pair<uint32_t, uint32_t>, std::map<uint32_t, std::vector<uint32_t>>> memberDeco; // (id, member)
    std::vector<Variable> variables;                          // module order (== SPIRVModule::getVariable(i))
    std::unordered_map<uint32_t, uint32_t> varIndex;
    std::unordered_map<uint32_t, Function> functions;
    std::unordered_map<uint32_t, uint32_t> constSlot;         // constant id -> word offset in globals
    std::unordered_map<uint32_t, uint32_t> idType;            // any value id -> type id
    std::vector<uint32_t> globals;                            // constants + variable pointers + logical storage
    uint32_t entryPoint = 0;
    uint32_t executionModel = 0;
    bool originUpperLeft = false;
    uint32_t glslExt = 0;

    bool hasDeco(uint32_t id, uint32_t d) const { return deco[id].count(d) != 0; }
    uint32_t decoVal(uint32_t id, uint32_t d, uint32_t def = 0) const {
        auto it = deco[id].find(d); return it == deco[id].end() || it->second.empty() ? def : it->second[0];
    }
    bool memberDecoVal(uint32_t id, uint32_t m, uint32_t d, uint32_t& out) const {
        auto it = memberDeco.find({id, m}); if (it == memberDeco.end()) return false;
        auto jt = it->second.find(d); if (jt == it->second.end()) return false;
        out = jt->second.empty() ? 0 : jt->second[0]; return true;
    }
};

[[noreturn]] inline void Fail(const std::string& s) { throw std::runtime_error("oracle spirv: " + s); }

inline uint32_t LogicalWords(Module& m, uint32_t id) { return m.types[id].words; }

inline void Parse(Module& m, const uint32_t* code, size_t n, const char* entryName, uint32_t model,
                  const CpvkSpecEntry* spec, uint32_t specCount) {
    if (n < 5 || code[0] != 0x07230203u) Fail("bad magic");
    m.words.assign(code, code + n);
    m.bound = m.words[3];
    m.types.assign(m.bound, Type());
    m.deco.assign(m.bound, {});
    m.executionModel = model;
    const uint32_t* w = m.words.data();
    Function* fn = nullptr;
    Block* blk = nullptr;
    std::vector<std::pair<uint32_t, const uint32_t*>> constants; // (id, instruction)
    size_t i = 5;
    auto str = [&](const uint32_t* p) { return std::string(reinterpret_cast<const char*>(p)); };
    while (i < n) {
        const uint32_t wc = w[i] >> 16, op = w[i] & 0xFFFF;
        if (wc == 0 || i + wc > n) Fail("truncated instruction");
        const uint32_t* o = w + i + 1;
        switch (op) {
        case OpExtInstImport: if (str(o + 1) == "GLSL.std.450") m.glslExt = o[0]; break;
        case OpEntryPoint: {
            if (o[0] == model && str(o + 2) == entryName) m.entryPoint = o[1];
            break; }
        case OpExecutionMode: if (o[1] == 7) { /* resolved after entry known */ } break;
        case OpDecorate: m.deco[o[0]][o[1]] = std::vector<uint32_t>(o + 2, o + wc - 1); break;
        case OpMemberDecorate: m.memberDeco[{o[0], o[1]}][o[2]] = std::vector<uint32_t>(o + 3, o + wc - 1); break;
        case OpTypeVoid: m.types[o[0]].kind = Type::Void; break;
        case OpTypeBool: m.types[o[0]].kind = Type::Bool; m.types[o[0]].words = 1; break;
        case OpTypeInt: { Type& t = m.types[o[0]]; t.kind = Type::Int; t.width = o[1]; t.isSigned = o[2] != 0; t.words = 1;
            if (t.width != 32) Fail("only 32-bit integers"); break; }
        case OpTypeFloat: { Type& t = m.types[o[0]]; t.kind = Type::Float; t.width = o[1]; t.words = 1;
            if (t.width != 32) Fail("only 32-bit floats"); break; }
        case OpTypeVector: { Type& t = m.types[o[0]]; t.kind = Type::Vector; t.elem = o[1]; t.count = o[2]; t.words = o[2]; break; }
        case OpTypeMatrix: { Type& t = m.types[o[0]]; t.kind = Type::Matrix; t.elem = o[1]; t.count = o[2]; t.words = o[2] * m.types[o[1]].words; break; }
        case OpTypeImage: { Type& t = m.types[o[0]]; t.kind = Type::Image; t.elem = o[1]; t.dim = o[2]; t.words = 2; break; }
        case OpTypeSampler: { Type& t = m.types[o[0]]; t.kind = Type::Sampler; t.words = 2; break; }
        case OpTypeSampledImage: { Type& t = m.types[o[0]]; t.kind = Type::SampledImage; t.elem = o[1]; t.words = 2; break; }
        case OpTypeArray: { Type& t = m.types[o[0]]; t.kind = Type::Array; t.elem = o[1]; t.count = 0; t.storage = o[2]; break; } // length resolved below
        case OpTypeRuntimeArray: { Type& t = m.types[o[0]]; t.kind = Type::RuntimeArray; t.elem = o[1]; t.words = 0; break; }
        case OpTypeStruct: { Type& t = m.types[o[0]]; t.kind = Type::Struct; t.members.assign(o + 1, o + wc - 1);
            t.words = 0; for (uint32_t mm : t.members) t.words += m.types[mm].words; break; }
        case OpTypePointer: { Type& t = m.types[o[0]]; t.kind = Type::Pointer; t.storage = o[1]; t.elem = o[2]; t.words = 4; break; }
        case OpTypeFunction: { Type& t = m.types[o[0]]; t.kind = Type::Function; t.elem = o[1]; t.members.assign(o + 2, o + wc - 1); break; }
        case OpConstantTrue: case OpConstantFalse: case OpConstant: case OpConstantComposite: case OpConstantNull:
        case OpSpecConstantTrue: case OpSpecConstantFalse: case OpSpecConstant: case OpSpecConstantComposite: case OpUndef:
            if (!fn) { constants.push_back({o[1], w + i}); m.idType[o[1]] = o[0]; }
            else if (op == OpUndef) { blk->insts.push_back(Inst{(uint16_t)op, 0, o[0], o[1], o + 2}); m.idType[o[1]] = o[0]; }
            break;
        case OpVariable:
            m.idType[o[1]] = o[0];
            if (o[2] != ScFunction) {
                Variable v{o[1], o[0], o[2], wc > 4 ? o[3] : 0, 0, 0};
                m.varIndex[v.id] = (uint32_t)m.variables.size();
                m.variables.push_back(v);
            } else {
                if (!blk) Fail("function variable outside block");
                blk->insts.push_back(Inst{(uint16_t)op, (uint16_t)(wc - 3), o[0], o[1], o + 2});
            }
            break;
        case OpFunction: { Function f; f.id = o[1]; f.retType = o[0]; m.functions[f.id] = f; fn = &m.functions[o[1]]; m.idType[o[1]] = o[0]; break; }
        case OpFunctionParameter: fn->params.push_back(o[1]); m.idType[o[1]] = o[0]; break;
        case OpFunctionEnd: fn = nullptr; blk = nullptr; break;
        case OpLabel: fn->blockIndex[o[0]] = (uint32_t)fn->blocks.size(); fn->blocks.push_back(Block{o[0], {}}); blk = &fn->blocks.back(); break;
        default:
            if (fn && blk) {
                // Generic instruction: figure out whether it has type/result from the opcode class.
                bool hasType = false, hasRes = false;
                switch (op) {
                case OpStore: case OpLoopMerge: case OpSelectionMerge: case OpBranch: case OpBranchConditional: case OpSwitch:
                case OpKill: case OpReturn: case OpReturnValue: case OpUnreachable: case OpNop: break;
                default: hasType = hasRes = true; break;
                }
                Inst in; in.op = (uint16_t)op;
                if (hasType) { in.type = o[0]; in.result = o[1]; in.ops = o + 2; in.nops = (uint16_t)(wc - 3); m.idType[in.result] = in.type; }
                else { in.type = 0; in.result = 0; in.ops = o; in.nops = (uint16_t)(wc - 1); }
                (void)hasRes;
                blk->insts.push_back(in);
            }
            break;
        }
        i += wc;
    }
    if (!m.entryPoint) Fail(std::string("entry point not found: ") + entryName);
    // execution modes
    for (i = 5; i < n; i += w[i] >> 16) if ((w[i] & 0xFFFF) == OpExecutionMode && w[i + 1] == m.entryPoint && w[i + 2] == 7) m.originUpperLeft = true;

    // ---- lay out the globals area: constants, then per variable a pointer slot (+ logical storage) ----
    auto alloc = [&](uint32_t words) { uint32_t s = (uint32_t)m.globals.size(); m.globals.resize(s + words, 0); return s; };
    // Array lengths need scalar constants first.
    std::unordered_map<uint32_t, uint32_t> scalarConst;
    for (auto& c : constants) {
        const uint32_t* ins = c.second; const uint32_t op = ins[0] & 0xFFFF;
        if (op == OpConstant || op == OpSpecConstant) {
            uint32_t v = ins[3];
            if (op == OpSpecConstant && m.hasDeco(c.first, DecoSpecId)) {
                const uint32_t sid = m.decoVal(c.first, DecoSpecId);
                for (uint32_t k = 0; k < specCount; k++) if (spec[k].constantId == sid) v = spec[k].value;
            }
            scalarConst[c.first] = v;
        }
    }
    // Resolve array types in id order (element types always precede).
    for (uint32_t id = 0; id < m.bound; id++) {
        Type& t = m.types[id];
        if (t.kind == Type::Array) { t.count = scalarConst.count(t.storage) ? scalarConst[t.storage] : 0; t.storage = 0; t.words = t.count * m.types[t.elem].words; }
        if (t.kind == Type::Struct) { t.words = 0; for (uint32_t mm : t.members) t.words += m.types[mm].words; }
        if (t.kind == Type::Matrix) t.words = t.count * m.types[t.elem].words;
    }
    for (auto& c : constants) {
        const uint32_t* ins = c.second; const uint32_t op = ins[0] & 0xFFFF, wc = ins[0] >> 16;
        const uint32_t ty = ins[1];
        const uint32_t s = alloc(m.types[ty].words ? m.types[ty].words : 1);
        m.constSlot[c.first] = s;
        auto specBool = [&](bool def) {
            if (m.hasDeco(c.first, DecoSpecId)) { const uint32_t sid = m.decoVal(c.first, DecoSpecId);
                for (uint32_t k = 0; k < specCount; k++) if (spec[k].constantId == sid) return spec[k].value != 0; }
            return def; };
        switch (op) {
        case OpConstantTrue: m.globals[s] = 1; break;
        case OpConstantFalse: m.globals[s] = 0; break;
        case OpSpecConstantTrue: m.globals[s] = specBool(true); break;
        case OpSpecConstantFalse: m.globals[s] = specBool(false); break;
        case OpConstant: case OpSpecConstant: m.globals[s] = scalarConst[c.first]; break;
        case OpConstantNull: case OpUndef: break;
        case OpConstantComposite: case OpSpecConstantComposite: {
            uint32_t off = s;
            for (uint32_t k = 3; k < wc; k++) {
                const uint32_t cid = ins[k]; const uint32_t cw = m.types[m.idType[cid]].words;
                std::memcpy(&m.globals[off], &m.globals[m.constSlot.at(cid)], cw * 4); off += cw;
            }
            break; }
        }
    }
    for (auto& v : m.variables) {
        v.slot = alloc(4);
        const uint32_t pointee = m.types[v.ptrType].elem;
        const bool logical = v.storage == ScInput || v.storage == ScOutput || v.storage == ScPrivate;
        v.dataSlot = logical ? alloc(m.types[pointee].words) : 0;
    }
    // ---- per-function frames ----
    for (auto& kv : m.functions) {
        Function& f = kv.second;
        uint32_t top = 0;
        for (uint32_t p : f.params) { f.slot[p] = top; top += m.types[m.idType[p]].words; }
        for (auto& b : f.blocks) for (auto& in : b.insts) {
            if (!in.result) continue;
            f.slot[in.result] = top; top += m.types[in.type].words ? m.types[in.type].words : 1;
            if (in.op == OpVariable) { const uint32_t pointee = m.types[in.type].elem; f.localVars.push_back({in.result, top}); top += m.types[pointee].words; }
        }
        f.frameWords = top;
    }
}

struct Env {
    const CpvkDescriptor* descriptors = nullptr;
    uint32_t descriptorCount = 0;
    const uint8_t* pushConstants = nullptr;
};

struct Interp {
    Module* m = nullptr;
    std::vector<uint32_t> g;           // working copy of globals for the current invocation
    std::vector<std::vector<uint32_t>> frames;
    size_t depth = 0;
    bool killed = false;
    Env env;

    void Bind(Module* mod, const Env& e) {
        m = mod; env = e;
        frames.resize(64); // fixed: references into `frames` must stay valid across nested calls
    }

    // OpSampledImage = @Image.Combine (GlslFunctions.cpp:812-820): the image descriptor's data with the sampler object's state.
    // One combined record per (image, sampler) pair, kept for the interpreter's lifetime.
    std::deque<CpvkDescriptor> combined;
    std::map<std::pair<const CpvkDescriptor*, const CpvkDescriptor*>, const CpvkDescriptor*> combinedOf;
    const CpvkDescriptor* Combine(const CpvkDescriptor* image, const CpvkDescriptor* sampler) {
        auto it = combinedOf.find({image, sampler});
        if (it != combinedOf.end()) return it->second;
        combined.push_back(*image);
        combined.back().sampler = sampler->sampler;
        return combinedOf[{image, sampler}] = &combined.back();
    }
    const CpvkDescriptor* FindDescriptor(uint32_t set, uint32_t binding, uint32_t element) const {
        for (uint32_t i = 0; i < env.descriptorCount; i++) {
            const CpvkDescriptor& d = env.descriptors[i];
            if (d.set == set && d.binding == binding && d.arrayElement == element) return &d;
        }
        return nullptr;
    }

    // Reset globals for one invocation: zero logical storage, apply initialisers, point buffer variables.
    void BeginInvocation() {
        g = m->globals;
        killed = false;
        for (auto& v : m->variables) {
            Ptr p{}; p.type = m->types[v.ptrType].elem;
            const Type& pt = m->types[p.type];
            if (v.dataSlot || v.storage == ScInput || v.storage == ScOutput || v.storage == ScPrivate) {
                p.addr = reinterpret_cast<uint8_t*>(&g[v.dataSlot]); p.buffer = 0;
                if (v.initializer) std::memcpy(&g[v.dataSlot], &g[m->constSlot.at(v.initializer)], pt.words * 4);
            } else if (v.storage == ScPushConstant) {
                p.addr = const_cast<uint8_t*>(env.pushConstants); p.buffer = 1;
            } else {
                const uint32_t set = m->decoVal(v.id, DecoDescriptorSet), binding = m->decoVal(v.id, DecoBinding);
                const CpvkDescriptor* d = FindDescriptor(set, binding, 0);
                if (pt.kind == Type::Image || pt.kind == Type::SampledImage || pt.kind == Type::Sampler) {
                    // the variable's storage is the descriptor handle itself
                    p.addr = reinterpret_cast<uint8_t*>(const_cast<CpvkDescriptor*>(d)); p.buffer = 2;
                } else {
                    p.addr = d ? reinterpret_cast<uint8_t*>((uintptr_t)d->address) : nullptr; p.buffer = 1;
                }
            }
            std::memcpy(&g[v.slot], &p, 16);
        }
    }

    uint32_t* VarData(uint32_t varId) { return &g[m->variables[m->varIndex.at(varId)].dataSlot]; }

    // ---- value access ----
    uint32_t* Val(Function& f, std::vector<uint32_t>& fr, uint32_t id) {
        auto it = f.slot.find(id);
        if (it != f.slot.end()) return &fr[it->second];
        auto ct = m->constSlot.find(id);
        if (ct != m->constSlot.end()) return &g[ct->second];
        auto vt = m->varIndex.find(id);
        if (vt != m->varIndex.end()) return &g[m->variables[vt->second].slot];
        Fail("unknown id " + std::to_string(id));
    }
    uint32_t TypeOf(uint32_t id) const { auto it = m->idType.find(id); if (it == m->idType.end()) Fail("untyped id"); return it->second; }

    // ---- buffer layout helpers ----
    uint32_t ArrayStride(uint32_t typeId) const { return m->decoVal(typeId, DecoArrayStride, m->types[m->types[typeId].elem].words * 4); }

    void LoadBuffer(const uint8_t* p, uint32_t typeId, uint32_t matStride, uint32_t* out) {
        const Type& t = m->types[typeId];
        switch (t.kind) {
        case Type::Bool: case Type::Int: case Type::Float: std::memcpy(out, p, 4); break;
        case Type::Vector: std::memcpy(out, p, 4 * t.count); break;
        case Type::Matrix: {
            const uint32_t cw = m->types[t.elem].words; const uint32_t stride = matStride ? matStride : cw * 4;
            for (uint32_t c = 0; c < t.count; c++) std::memcpy(out + c * cw, p + c * stride, cw * 4);
            break; }
        case Type::Array: { const uint32_t ew = m->types[t.elem].words, st = ArrayStride(typeId);
            for (uint32_t k = 0; k < t.count; k++) LoadBuffer(p + k * st, t.elem, matStride, out + k * ew); break; }
        case Type::Struct: { uint32_t off = 0;
            for (uint32_t k = 0; k < t.members.size(); k++) {
                uint32_t bo = 0, ms = 0; m->memberDecoVal(typeId, k, DecoOffset, bo); m->memberDecoVal(typeId, k, DecoMatrixStride, ms);
                LoadBuffer(p + bo, t.members[k], ms, out + off); off += m->types[t.members[k]].words; }
            break; }
        default: Fail("unsupported buffer load");
        }
    }
    void StoreBuffer(uint8_t* p, uint32_t typeId, uint32_t matStride, const uint32_t* in) {
        const Type& t = m->types[typeId];
        switch (t.kind) {
        case Type::Bool: case Type::Int: case Type::Float: std::memcpy(p, in, 4); break;
        case Type::Vector: std::memcpy(p, in, 4 * t.count); break;
        case Type::Matrix: { const uint32_t cw = m->types[t.elem].words; const uint32_t stride = matStride ? matStride : cw * 4;
            for (uint32_t c = 0; c < t.count; c++) std::memcpy(p + c * stride, in + c * cw, cw * 4); break; }
        case Type::Array: { const uint32_t ew = m->types[t.elem].words, st = ArrayStride(typeId);
            for (uint32_t k = 0; k < t.count; k++) StoreBuffer(p + k * st, t.elem, matStride, in + k * ew); break; }
        case Type::Struct: { uint32_t off = 0;
            for (uint32_t k = 0; k < t.members.size(); k++) {
                uint32_t bo = 0, ms = 0; m->memberDecoVal(typeId, k, DecoOffset, bo); m->memberDecoVal(typeId, k, DecoMatrixStride, ms);
                StoreBuffer(p + bo, t.members[k], ms, in + off); off += m->types[t.members[k]].words; }
            break; }
        default: Fail("unsupported buffer store");
        }
    }

    static float F(uint32_t v) { float f; std::memcpy(&f, &v, 4); return f; }
    static uint32_t U(float f) { uint32_t v; std::memcpy(&v, &f, 4); return v; }

    // glm::dot operand order (func_geometric.inl compute_dot): vec2 a.x*b.x + a.y*b.y; vec3 left to right;
    // vec4 (tmp.x + tmp.y) + (tmp.z + tmp.w).
    static void GlslOp(uint32_t e, const uint32_t* a, const uint32_t* b, const uint32_t* c, uint32_t n, uint32_t cnt, uint32_t* r);
    static void MatrixTimesVector(const uint32_t* a, const uint32_t* v, uint32_t cols, uint32_t rows, uint32_t* out);
    static void VectorTimesMatrix(const uint32_t* v, const uint32_t* a, uint32_t cols, uint32_t rows, uint32_t* out);
    static void MatrixTimesMatrix(const uint32_t* a, const uint32_t* b, uint32_t lcols, uint32_t lrows, uint32_t rcols, uint32_t* out);
    static float Dot(const uint32_t* a, const uint32_t* b, uint32_t n) {
        float t[4]; for (uint32_t i = 0; i < n; i++) t[i] = F(a[i]) * F(b[i]);
        if (n == 1) return t[0];
        if (n == 2) return t[0] + t[1];
        if (n == 3) return t[0] + t[1] + t[2];
        return (t[0] + t[1]) + (t[2] + t[3]);
    }

    // ---- execution ----
    // Returns pointer to the return value words (valid until next call at same depth), or nullptr.
    uint32_t* Call(uint32_t fnId, const uint32_t* const* args, const uint32_t* argWords, uint32_t nargs) {
        Function& f = m->functions.at(fnId);
        if (depth >= frames.size()) Fail("call depth exceeded");
        std::vector<uint32_t>& fr = frames[depth];
        fr.assign(f.frameWords + 64, 0);
        depth++;
        for (uint32_t a = 0; a < nargs; a++) std::memcpy(&fr[f.slot.at(f.params[a])], args[a], argWords[a] * 4);
        for (auto& lv : f.localVars) {
            Ptr p{}; p.addr = reinterpret_cast<uint8_t*>(&fr[lv.second]); p.type = m->types[TypeOf(lv.first)].elem; p.buffer = 0;
            std::memcpy(&fr[f.slot.at(lv.first)], &p, 16);
        }
        uint32_t cur = 0, prevLabel = 0;
        uint32_t* ret = nullptr;
        for (;;) {
            Block& b = f.blocks[cur];
            uint32_t next = UINT32_MAX;
            // OpPhi nodes read their inputs "simultaneously": evaluate all first, then commit.
            size_t ip = 0;
            {
                std::vector<std::pair<uint32_t*, std::vector<uint32_t>>> pending;
                for (; ip < b.insts.size() && b.insts[ip].op == OpPhi; ip++) {
                    const Inst& in = b.insts[ip];
                    const uint32_t wds = m->types[in.type].words;
                    for (uint32_t k = 0; k + 1 < in.nops; k += 2) if (in.ops[k + 1] == prevLabel) {
                        uint32_t* src = Val(f, fr, in.ops[k]);
                        pending.push_back({Val(f, fr, in.result), std::vector<uint32_t>(src, src + wds)}); break; }
                }
                for (auto& p : pending) std::memcpy(p.first, p.second.data(), p.second.size() * 4);
            }
            for (; ip < b.insts.size(); ip++) {
                const Inst& in = b.insts[ip];
                if (Exec(f, fr, in, next, ret)) { depth--; return ret; }
                if (next != UINT32_MAX) break;
            }
            if (next == UINT32_MAX) Fail("block fell through");
            prevLabel = b.label;
            cur = f.blocks.size() > 0 ? f.blockIndex.at(next) : 0;
        }
    }

    // Executes one instruction. Returns true when the function is finished.
    bool Exec(Function& f, std::vector<uint32_t>& fr, const Inst& in, uint32_t& next, uint32_t*& ret);
    void ExtInst(Function& f, std::vector<uint32_t>& fr, const Inst& in);
};

inline bool Interp::Exec(Function& f, std::vector<uint32_t>& fr, const Inst& in, uint32_t& next, uint32_t*& ret) {
    auto V = [&](uint32_t id) { return Val(f, fr, id); };
    const Type* rt = in.type ? &m->types[in.type] : nullptr;
    uint32_t* r = in.result ? V(in.result) : nullptr;
    const uint32_t n = rt ? (rt->words ? rt->words : 1) : 0;
    #define UN_F(expr) { const uint32_t* a = V(in.ops[0]); for (uint32_t i = 0; i < n; i++) { const float x = F(a[i]); r[i] = U(expr); } return false; }
    #define BIN_F(expr) { const uint32_t* a = V(in.ops[0]); const uint32_t* b = V(in.ops[1]); for (uint32_t i = 0; i < n; i++) { const float x = F(a[i]), y = F(b[i]); r[i] = U(expr); } return false; }
    #define BIN_U(expr) { const uint32_t* a = V(in.ops[0]); const uint32_t* b = V(in.ops[1]); for (uint32_t i = 0; i < n; i++) { const uint32_t x = a[i], y = b[i]; (void)x; (void)y; r[i] = (uint32_t)(expr); } return false; }
    #define BIN_S(expr) { const uint32_t* a = V(in.ops[0]); const uint32_t* b = V(in.ops[1]); for (uint32_t i = 0; i < n; i++) { const int32_t x = (int32_t)a[i], y = (int32_t)b[i]; (void)x; (void)y; r[i] = (uint32_t)(expr); } return false; }
    #define CMP_F(expr) { const uint32_t* a = V(in.ops[0]); const uint32_t* b = V(in.ops[1]); for (uint32_t i = 0; i < n; i++) { const float x = F(a[i]), y = F(b[i]); r[i] = (expr) ? 1u : 0u; } return false; }
    switch (in.op) {
    case OpNop: case OpLoopMerge: case OpSelectionMerge: return false;
    case OpUndef: return false;
    case OpVariable: {
        if (in.nops > 1) { // initializer
            Ptr p; std::memcpy(&p, r, 16);
            std::memcpy(p.addr, V(in.ops[1]), m->types[p.type].words * 4);
        }
        return false; }
    case OpLoad: {
        Ptr p; std::memcpy(&p, V(in.ops[0]), 16);
        if (p.buffer == 2) { uint64_t h = (uint64_t)(uintptr_t)p.addr; std::memcpy(r, &h, 8); }
        else if (p.buffer == 1) LoadBuffer(p.addr, in.type, p.matStride, r);
        else std::memcpy(r, p.addr, n * 4);
        return false; }
    case OpStore: {
        Ptr p; std::memcpy(&p, V(in.ops[0]), 16);
        const uint32_t* src = V(in.ops[1]);
        if (p.buffer == 1) StoreBuffer(p.addr, p.type, p.matStride, src);
        else std::memcpy(p.addr, src, m->types[p.type].words * 4);
        return false; }
    case OpAccessChain: case OpInBoundsAccessChain: {
        Ptr p; std::memcpy(&p, V(in.ops[0]), 16);
        for (uint32_t k = 1; k < in.nops; k++) {
            const uint32_t idx = *V(in.ops[k]);
            const Type& t = m->types[p.type];
            switch (t.kind) {
            case Type::Struct:
                if (p.buffer == 1) { uint32_t bo = 0, ms = 0; m->memberDecoVal(p.type, idx, DecoOffset, bo); m->memberDecoVal(p.type, idx, DecoMatrixStride, ms);
                    p.addr += bo; p.matStride = (uint16_t)ms; }
                else { uint32_t off = 0; for (uint32_t q = 0; q < idx; q++) off += m->types[t.members[q]].words; p.addr += off * 4; }
                p.type = t.members[idx]; break;
            case Type::Array: case Type::RuntimeArray:
                p.addr += (size_t)idx * (p.buffer == 1 ? ArrayStride(p.type) : m->types[t.elem].words * 4);
                p.type = t.elem; break;
            case Type::Matrix:
                p.addr += (size_t)idx * (p.buffer == 1 && p.matStride ? p.matStride : m->types[t.elem].words * 4);
                p.type = t.elem; break;
            case Type::Vector: p.addr += (size_t)idx * 4; p.type = t.elem; break;
            default: Fail("bad access chain");
            }
        }
        std::memcpy(r, &p, 16);
        return false; }
    case OpFunctionCall: {
        const uint32_t nargs = in.nops - 1;
        const uint32_t* args[16]; uint32_t aw[16];
        for (uint32_t a = 0; a < nargs; a++) { args[a] = V(in.ops[1 + a]); aw[a] = m->types[TypeOf(in.ops[1 + a])].words; }
        uint32_t* rv = Call(in.ops[0], args, aw, nargs);
        // frames may have been resized: recompute r
        std::vector<uint32_t>& fr2 = frames[depth - 1];
        if (rv && n) std::memcpy(&fr2[f.slot.at(in.result)], rv, n * 4);
        return false; }
    case OpCopyObject: std::memcpy(r, V(in.ops[0]), n * 4); return false;
    case OpVectorShuffle: {
        const uint32_t* a = V(in.ops[0]); const uint32_t* b = V(in.ops[1]);
        const uint32_t na = m->types[TypeOf(in.ops[0])].count;
        for (uint32_t i = 0; i < n; i++) { const uint32_t s = in.ops[2 + i]; r[i] = s == 0xFFFFFFFFu ? 0 : (s < na ? a[s] : b[s - na]); }
        return false; }
    case OpCompositeConstruct: {
        uint32_t off = 0;
        for (uint32_t k = 0; k < in.nops; k++) { const uint32_t w = m->types[TypeOf(in.ops[k])].words; std::memcpy(r + off, V(in.ops[k]), w * 4); off += w; }
        return false; }
    case OpCompositeExtract: case OpCompositeInsert: {
        const bool ins = in.op == OpCompositeInsert;
        const uint32_t compId = ins ? in.ops[1] : in.ops[0];
        uint32_t ty = TypeOf(compId), off = 0;
        for (uint32_t k = ins ? 2 : 1; k < in.nops; k++) {
            const uint32_t idx = in.ops[k]; const Type& t = m->types[ty];
            if (t.kind == Type::Struct) { for (uint32_t q = 0; q < idx; q++) off += m->types[t.members[q]].words; ty = t.members[idx]; }
            else { off += idx * m->types[t.elem].words; ty = t.elem; }
        }
        if (ins) { std::memcpy(r, V(compId), n * 4); std::memcpy(r + off, V(in.ops[0]), m->types[ty].words * 4); }
        else std::memcpy(r, V(compId) + off, n * 4);
        return false; }
    case OpVectorExtractDynamic: r[0] = V(in.ops[0])[*V(in.ops[1])]; return false;
    case OpVectorInsertDynamic: std::memcpy(r, V(in.ops[0]), n * 4); r[*V(in.ops[2])] = *V(in.ops[1]); return false;
    case OpTranspose: {
        const Type& mt = m->types[TypeOf(in.ops[0])]; const uint32_t cols = mt.count, rows = m->types[mt.elem].count;
        const uint32_t* a = V(in.ops[0]);
        for (uint32_t c = 0; c < cols; c++) for (uint32_t q = 0; q < rows; q++) r[q * cols + c] = a[c * rows + q];
        return false; }
    case OpFNegate: UN_F(-x)
    case OpSNegate: { const uint32_t* a = V(in.ops[0]); for (uint32_t i = 0; i < n; i++) r[i] = 0u - a[i]; return false; }
    case OpFAdd: BIN_F(x + y)
    case OpFSub: BIN_F(x - y)
    case OpFMul: BIN_F(x * y)
    case OpFDiv: BIN_F(x / y)
    case OpFRem: BIN_F(fmodf(x, y))
    case OpFMod: BIN_F(([&] { float q = fmodf(x, y); if (q != 0 && ((q < 0) != (y < 0))) q += y; return q; })())
    case OpIAdd: BIN_U(x + y)
    case OpISub: BIN_U(x - y)
    case OpIMul: BIN_U(x * y)
    case OpUDiv: BIN_U(y ? x / y : 0)
    case OpSDiv: BIN_S((y == 0 || (x == INT32_MIN && y == -1)) ? 0 : x / y)
    case OpUMod: BIN_U(y ? x % y : 0)
    case OpSRem: BIN_S((y == 0 || (x == INT32_MIN && y == -1)) ? 0 : x % y)
    case OpSMod: BIN_S(([&] { if (y == 0 || (x == INT32_MIN && y == -1)) return 0; int32_t q = x % y; if (q != 0 && ((q < 0) != (y < 0))) q += y; return q; })())
    case OpVectorTimesScalar: { const uint32_t* a = V(in.ops[0]); const float s = F(*V(in.ops[1])); for (uint32_t i = 0; i < n; i++) r[i] = U(F(a[i]) * s); return false; }
    case OpMatrixTimesScalar: { const uint32_t* a = V(in.ops[0]); const float s = F(*V(in.ops[1])); for (uint32_t i = 0; i < n; i++) r[i] = U(F(a[i]) * s); return false; }
    case OpMatrixTimesVector: {
        const Type& mt = m->types[TypeOf(in.ops[0])]; const uint32_t cols = mt.count, rows = m->types[mt.elem].count;
        uint32_t out[4];
        MatrixTimesVector(V(in.ops[0]), V(in.ops[1]), cols, rows, out);
        std::memcpy(r, out, rows * 4);
        return false; }
    case OpVectorTimesMatrix: {
        const Type& mt = m->types[TypeOf(in.ops[1])]; const uint32_t cols = mt.count, rows = m->types[mt.elem].count;
        uint32_t out[4];
        VectorTimesMatrix(V(in.ops[0]), V(in.ops[1]), cols, rows, out);
        std::memcpy(r, out, cols * 4);
        return false; }
    case OpMatrixTimesMatrix: {
        const Type& lt = m->types[TypeOf(in.ops[0])]; const uint32_t lcols = lt.count, lrows = m->types[lt.elem].count;
        const Type& rtm = m->types[TypeOf(in.ops[1])]; const uint32_t rcols = rtm.count;
        uint32_t out[16];
        MatrixTimesMatrix(V(in.ops[0]), V(in.ops[1]), lcols, lrows, rcols, out);
        std::memcpy(r, out, rcols * lrows * 4);
        return false; }
    case OpDot: { const uint32_t cnt = m->types[TypeOf(in.ops[0])].count; r[0] = U(Dot(V(in.ops[0]), V(in.ops[1]), cnt)); return false; }
    case OpConvertFToU: { const uint32_t* a = V(in.ops[0]); for (uint32_t i = 0; i < n; i++) { float x = F(a[i]); r[i] = std::isnan(x) || x <= -1.0f ? 0u : (x >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)x); } return false; }
    case OpConvertFToS: { const uint32_t* a = V(in.ops[0]); for (uint32_t i = 0; i < n; i++) { float x = F(a[i]); r[i] = std::isnan(x) ? 0u : (x >= 2147483648.0f ? 0x7FFFFFFFu : (x < -2147483648.0f ? 0x80000000u : (uint32_t)(int32_t)x)); } return false; }
    case OpConvertSToF: { const uint32_t* a = V(in.ops[0]); for (uint32_t i = 0; i < n; i++) r[i] = U((float)(int32_t)a[i]); return false; }
    case OpConvertUToF: { const uint32_t* a = V(in.ops[0]); for (uint32_t i = 0; i < n; i++) r[i] = U((float)a[i]); return false; }
    case OpBitcast: std::memcpy(r, V(in.ops[0]), n * 4); return false;
    case OpAny: { const uint32_t cnt = m->types[TypeOf(in.ops[0])].count; const uint32_t* a = V(in.ops[0]); uint32_t x = 0; for (uint32_t i = 0; i < cnt; i++) x |= a[i]; r[0] = x ? 1 : 0; return false; }
    case OpAll: { const uint32_t cnt = m->types[TypeOf(in.ops[0])].count; const uint32_t* a = V(in.ops[0]); uint32_t x = 1; for (uint32_t i = 0; i < cnt; i++) x &= a[i] ? 1 : 0; r[0] = x; return false; }
    case OpIsNan: { const uint32_t* a = V(in.ops[0]); for (uint32_t i = 0; i < n; i++) r[i] = std::isnan(F(a[i])); return false; }
    case OpIsInf: { const uint32_t* a = V(in.ops[0]); for (uint32_t i = 0; i < n; i++) r[i] = std::isinf(F(a[i])); return false; }
    case OpLogicalEqual: BIN_U((x != 0) == (y != 0))
    case OpLogicalNotEqual: BIN_U((x != 0) != (y != 0))
    case OpLogicalOr: BIN_U((x | y) != 0)
    case OpLogicalAnd: BIN_U((x != 0) && (y != 0))
    case OpLogicalNot: { const uint32_t* a = V(in.ops[0]); for (uint32_t i = 0; i < n; i++) r[i] = a[i] ? 0 : 1; return false; }
    case OpSelect: {
        const uint32_t* c = V(in.ops[0]); const uint32_t* a = V(in.ops[1]); const uint32_t* b = V(in.ops[2]);
        const bool vecCond = m->types[TypeOf(in.ops[0])].kind == Type::Vector;
        for (uint32_t i = 0; i < n; i++) r[i] = (vecCond ? c[i] : c[0]) ? a[i] : b[i];
        return false; }
    case OpIEqual: BIN_U(x == y)
    case OpINotEqual: BIN_U(x != y)
    case OpUGreaterThan: BIN_U(x > y)
    case OpSGreaterThan: BIN_S(x > y)
    case OpUGreaterThanEqual: BIN_U(x >= y)
    case OpSGreaterThanEqual: BIN_S(x >= y)
    case OpULessThan: BIN_U(x < y)
    case OpSLessThan: BIN_S(x < y)
    case OpULessThanEqual: BIN_U(x <= y)
    case OpSLessThanEqual: BIN_S(x <= y)
    case OpFOrdEqual: CMP_F(x == y)
    case OpFUnordEqual: CMP_F(!(x < y || x > y))
    case OpFOrdNotEqual: CMP_F(x < y || x > y)
    case OpFUnordNotEqual: CMP_F(x != y)
    case OpFOrdLessThan: CMP_F(x < y)
    case OpFUnordLessThan: CMP_F(!(x >= y))
    case OpFOrdGreaterThan: CMP_F(x > y)
    case OpFUnordGreaterThan: CMP_F(!(x <= y))
    case OpFOrdLessThanEqual: CMP_F(x <= y)
    case OpFUnordLessThanEqual: CMP_F(!(x > y))
    case OpFOrdGreaterThanEqual: CMP_F(x >= y)
    case OpFUnordGreaterThanEqual: CMP_F(!(x < y))
    case OpShiftRightLogical: BIN_U(x >> (y & 31))
    case OpShiftRightArithmetic: BIN_S(x >> (y & 31))
    case OpShiftLeftLogical: BIN_U(x << (y & 31))
    case OpBitwiseOr: BIN_U(x | y)
    case OpBitwiseXor: BIN_U(x ^ y)
    case OpBitwiseAnd: BIN_U(x & y)
    case OpNot: { const uint32_t* a = V(in.ops[0]); for (uint32_t i = 0; i < n; i++) r[i] = ~a[i]; return false; }
    case OpExtInst: ExtInst(f, fr, in); return false;
    case OpSampledImage: {
        uint64_t hi, hs; std::memcpy(&hi, V(in.ops[0]), 8); std::memcpy(&hs, V(in.ops[1]), 8);
        if (!hi || !hs) Fail("OpSampledImage on an unbound image or sampler");
        const uint64_t h = (uint64_t)(uintptr_t)Combine(reinterpret_cast<const CpvkDescriptor*>((uintptr_t)hi), reinterpret_cast<const CpvkDescriptor*>((uintptr_t)hs));
        std::memcpy(r, &h, 8); return false; }
    case OpImage: std::memcpy(r, V(in.ops[0]), 8); return false;
    case OpImageSampleImplicitLod: case OpImageSampleExplicitLod: {
        uint64_t h; std::memcpy(&h, V(in.ops[0]), 8);
        const CpvkDescriptor* d = reinterpret_cast<const CpvkDescriptor*>((uintptr_t)h);
        if (!d) Fail("unbound image descriptor");
        const uint32_t cw = m->types[TypeOf(in.ops[1])].words;
        const uint32_t* c = V(in.ops[1]);
        float coord[3] = {0, 0, 0}; for (uint32_t i = 0; i < cw && i < 3; i++) coord[i] = F(c[i]);
        float lod = 0;
        if (in.op == OpImageSampleExplicitLod) { if (in.nops < 4 || in.ops[2] != 2) Fail("only Lod image operand"); lod = F(*V(in.ops[3])); }
        else if (in.nops > 2) Fail("image operands on implicit-lod sample");
        const Vec4f s = ImageSampleExplicitLod(*d, coord, lod);
        for (uint32_t i = 0; i < n; i++) r[i] = U(s.v[i]);
        return false; }
    case OpImageFetch: case OpImageRead: { // @Image.Read = ImageFetch, coordinate as written (GlslFunctions.cpp:739-743)
        uint64_t h; std::memcpy(&h, V(in.ops[0]), 8);
        const CpvkDescriptor* d = reinterpret_cast<const CpvkDescriptor*>((uintptr_t)h);
        if (!d) Fail("unbound image descriptor");
        const uint32_t cw = m->types[TypeOf(in.ops[1])].words; const uint32_t* c = V(in.ops[1]);
        int32_t coord[3] = {0, 0, 0}; for (uint32_t i = 0; i < cw && i < 3; i++) coord[i] = (int32_t)c[i];
        const Vec4f s = ImageFetch(*d, coord);
        for (uint32_t i = 0; i < n; i++) r[i] = U(s.v[i]);
        return false; }
    case OpBranch: next = in.ops[0]; return false;
    case OpBranchConditional: next = *V(in.ops[0]) ? in.ops[1] : in.ops[2]; return false;
    case OpSwitch: {
        const uint32_t sel = *V(in.ops[0]); next = in.ops[1];
        for (uint32_t k = 2; k + 1 < in.nops; k += 2) if (in.ops[k] == sel) { next = in.ops[k + 1]; break; }
        return false; }
    case OpKill: killed = true; ret = nullptr; return true;
    case OpReturn: ret = nullptr; return true;
    case OpReturnValue: ret = V(in.ops[0]); return true;
    case OpUnreachable: Fail("reached OpUnreachable");
    default: Fail("unsupported opcode " + std::to_string(in.op));
    }
    #undef UN_F
    #undef BIN_F
    #undef BIN_U
    #undef BIN_S
    #undef CMP_F
}

// @Matrix.Mult.* (SpirvFunctions.cpp:6-60 -> glm operators), column-major, on raw 32-bit lanes.
// glm mat*vec: 4x4 -> (m0*v0 + m1*v1) + (m2*v2 + m3*v3); 2 / 3 columns -> left to right.
inline void Interp::MatrixTimesVector(const uint32_t* a, const uint32_t* v, uint32_t cols, uint32_t rows, uint32_t* out) {
    for (uint32_t q = 0; q < rows; q++) {
        float p[4]; for (uint32_t c = 0; c < cols; c++) p[c] = F(a[c * rows + q]) * F(v[c]);
        float s;
        if (cols == 4 && rows == 4) s = (p[0] + p[1]) + (p[2] + p[3]);
        else { s = p[0]; for (uint32_t c = 1; c < cols; c++) s = s + p[c]; }
        out[q] = U(s);
    }
}
// glm vec*mat: result[c] = left-to-right sum over the rows of m[c][k] * v[k].
inline void Interp::VectorTimesMatrix(const uint32_t* v, const uint32_t* a, uint32_t cols, uint32_t rows, uint32_t* out) {
    for (uint32_t c = 0; c < cols; c++) { float s = F(a[c * rows]) * F(v[0]); for (uint32_t k = 1; k < rows; k++) s = s + F(a[c * rows + k]) * F(v[k]); out[c] = U(s); }
}
// glm mat*mat: Result[j] = A0*B[j][0] + A1*B[j][1] + ... left to right.
inline void Interp::MatrixTimesMatrix(const uint32_t* a, const uint32_t* b, uint32_t lcols, uint32_t lrows, uint32_t rcols, uint32_t* out) {
    for (uint32_t j = 0; j < rcols; j++) for (uint32_t q = 0; q < lrows; q++) {
        float s = F(a[q]) * F(b[j * lcols]);
        for (uint32_t k = 1; k < lcols; k++) s = s + F(a[k * lrows + q]) * F(b[j * lcols + k]);
        out[j * lrows + q] = U(s);
    }
}

// GLSL.std.450 as the reference implements it (GlslFunctions.cpp:19-321): std::min/max/clamp comparison
// forms, Mix = x*(1-a)+y*a, glm::normalize = v * (1/sqrt(dot(v,v))), glm::reflect = I - N*dot(N,I)*2.
inline void Interp::ExtInst(Function& f, std::vector<uint32_t>& fr, const Inst& in) {
    auto V = [&](uint32_t id) { return Val(f, fr, id); };
    if (in.ops[0] != m->glslExt) Fail("unknown extended instruction set");
    const uint32_t* a = in.nops > 2 ? V(in.ops[2]) : nullptr;
    const uint32_t* b = in.nops > 3 ? V(in.ops[3]) : nullptr;
    const uint32_t* c = in.nops > 4 ? V(in.ops[4]) : nullptr;
    GlslOp(in.ops[1], a, b, c, m->types[in.type].words, in.nops > 2 ? m->types[TypeOf(in.ops[2])].words : 0, V(in.result));
}
// One GLSL.std.450 instruction on raw 32-bit lanes: n = words of the result, cnt = words of the first operand (the vector width
// of Length / Distance / Normalize / Reflect / Cross). Also what cpvk_oracle_math runs for tests/test_reference_math.py.
inline void Interp::GlslOp(uint32_t e, const uint32_t* a, const uint32_t* b, const uint32_t* c, uint32_t n, uint32_t cnt, uint32_t* r) {
    auto mn = [](auto x, auto y) { return y < x ? y : x; };
    auto mx = [](auto x, auto y) { return x < y ? y : x; };
    auto cl = [](auto v, auto lo, auto hi) { return v < lo ? lo : (hi < v ? hi : v); };
    for (uint32_t i = 0; i < n; i++) {
        const float x = a ? F(a[i]) : 0, y = b ? F(b[i]) : 0, z = c ? F(c[i]) : 0;
        const int32_t sx = a ? (int32_t)a[i] : 0, sy = b ? (int32_t)b[i] : 0, sz = c ? (int32_t)c[i] : 0;
        const uint32_t ux = a ? a[i] : 0, uy = b ? b[i] : 0, uz = c ? c[i] : 0;
        switch (e) {
        case 4: r[i] = U(std::fabs(x)); break;                                         // FAbs
        case 5: r[i] = (uint32_t)(sx < 0 ? -sx : sx); break;                           // SAbs
        case 6: r[i] = U((float)(0.0f < x) - (float)(x < 0.0f)); break;                // FSign
        case 7: r[i] = (uint32_t)((int32_t)(0 < sx) - (int32_t)(sx < 0)); break;       // SSign
        case 8: r[i] = U(std::floor(x)); break;
        case 9: r[i] = U(std::ceil(x)); break;
        case 10: r[i] = U(x - std::floor(x)); break;                                   // Fract
        case 13: r[i] = U(std::sin(x)); break;
        case 14: r[i] = U(std::cos(x)); break;
        case 26: r[i] = U(std::pow(x, y)); break;
        case 27: r[i] = U(std::exp(x)); break;
        case 28: r[i] = U(std::log(x)); break;
        case 29: r[i] = U(std::exp2(x)); break;
        case 30: r[i] = U(std::log2(x)); break;
        case 31: r[i] = U(std::sqrt(x)); break;
        case 32: r[i] = U(1.0f / std::sqrt(x)); break;                                 // InverseSqrt
        case 37: r[i] = U(mn(x, y)); break; case 38: r[i] = mn(ux, uy); break; case 39: r[i] = (uint32_t)mn(sx, sy); break;
        case 40: r[i] = U(mx(x, y)); break; case 41: r[i] = mx(ux, uy); break; case 42: r[i] = (uint32_t)mx(sx, sy); break;
        case 43: r[i] = U(cl(x, y, z)); break; case 44: r[i] = cl(ux, uy, uz); break; case 45: r[i] = (uint32_t)cl(sx, sy, sz); break;
        case 46: r[i] = U(x * (1 - z) + y * z); break;                                 // FMix(x, y, a)
        case 79: r[i] = U(std::isnan(x) ? y : mn(x, y)); break;                        // NMin
        case 80: r[i] = U(std::isnan(x) ? y : mx(x, y)); break;                        // NMax
        case 81: { float t = std::isnan(x) ? y : mx(x, y); r[i] = U(std::isnan(t) ? z : mn(t, z)); break; } // NClamp
        case 66: case 69: case 71: case 68: case 67: goto vector_ops;
        default: Fail("unsupported GLSL.std.450 instruction " + std::to_string(e));
        }
    }
    return;
vector_ops: {
    switch (e) {
    case 66: r[0] = U(std::sqrt(Dot(a, a, cnt))); break;                              // Length = sqrt(dot(v,v))
    case 67: { uint32_t d[4]; for (uint32_t i = 0; i < cnt; i++) d[i] = U(F(b[i]) - F(a[i])); r[0] = U(std::sqrt(Dot(d, d, cnt))); break; } // glm::distance = length(p1 - p0)
    case 69: { const float inv = 1.0f / std::sqrt(Dot(a, a, cnt)); for (uint32_t i = 0; i < cnt; i++) r[i] = U(F(a[i]) * inv); break; }
    case 71: { const float d = Dot(b, a, cnt); for (uint32_t i = 0; i < cnt; i++) r[i] = U(F(a[i]) - F(b[i]) * d * 2.0f); break; }
    case 68: { const float ax = F(a[0]), ay = F(a[1]), az = F(a[2]), bx = F(b[0]), by = F(b[1]), bz = F(b[2]);
               r[0] = U(ay * bz - by * az); r[1] = U(az * bx - bz * ax); r[2] = U(ax * by - bx * ay); break; }
    }
    }
}

} // namespace spv
} // namespace oracle
