// spirv_to_cuda.h — application SPIR-V -> CUDA C++ device functions (the replacement for the reference's
// SPIR-V -> LLVM-IR -> x86 lowering in LLVMRuntime/SPIRVCompiler.cpp + PipelineCompiler.cpp).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/cpvk_cuda.h"

namespace cpvk {

// One shader resource variable -> a run of consecutive slots in CpvkDrawParams::desc.
struct ResourceSlot {
    uint32_t set, binding, count, slotBase;
};

struct PipelineLayoutInfo {
    std::vector<ResourceSlot> slots; // shared by both stages of a pipeline
    uint32_t slotCount = 0;
    uint32_t recordWords = 6;        // VS output record size in 32-bit words (24-byte builtin block + outputs)
    bool originUpperLeft = false;    // FS OriginUpperLeft execution mode (Pipeline.cpp:976-984)
    bool fsSamplesImages = false;    // the fragment shader samples or fetches images (large software sampler inlined: more registers pay off)
    bool vsWritesMemory = false;     // the vertex shader stores to a buffer: its invocation count is observable
    bool fsWritesMemory = false;     // the fragment shader stores to a buffer
};

// Translates one stage. `model` is the SPIR-V execution model (0 vertex, 4 fragment). Appends the generated
// function (cpvk_vs_main / cpvk_fs_main) to `out`. Returns 0 or a CPVK_E_* code with `error` filled.
int TranslateStage(const CpvkShaderStage& stage, uint32_t model, const CpvkPipelineDesc& desc, PipelineLayoutInfo& layout,
                   std::string& out, std::string& error);

} // namespace cpvk
