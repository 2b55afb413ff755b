// ref_interface_check.cpp — runs the REFERENCE's own shader-interface reflection: GetVariableFormat, GetVariableSize and
// GetVariablePointers (CPVulkan/CommandBuffer.Draw.cpp:151-354, :420-565), lifted out of the file where it lies by ref_slice.py, on
// SPIR-V modules loaded by the reference's own front end (SPIRVParser/, as CPVulkan/ShaderModule.cpp:62-73 does), compiled IN PLACE
// by oracle/Makefile into oracle/_ref/interface_check. This is the code that decides, per fragment-shader input, the Location, the
// format SetDatum interpolates it as, the interpolation kind, its size and its byte offset inside the vertex stage's output
// record (VS -> FS linkage by declaration order; §8(a) a3 / a6). The LLVM-compiled module it asks for variable
// addresses is stood in for by a name -> dummy-address table: only the addresses' identity is used.
// TEST INFRASTRUCTURE ONLY: tests/test_reference_interface.py compares the oracle's Reflect (oracle_draw.cpp) with its output.
#include <SPIRVFunction.h>
#include <SPIRVInstruction.h>
#include <SPIRVModule.h>
#include <SPIRVValue.h>

#include <cassert>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <Base.h>
#include <Formats.h>
#include <PipelineData.h>

class CompiledModule {
public:
    void* getPointer(const std::string& name) const { return &table[name]; }
private:
    mutable std::map<std::string, uint64_t> table;
};
// LLVMRuntime/Compilers.h:18 — any injective name will do here
std::string MangleName(const SPIRV::SPIRVVariable* variable) { return "@" + std::to_string(variable->getId()); }

#include "interface_slices.inc" // written by oracle/ref_slice.py into the scratch build directory (-I)

// interface_check <vertex|fragment> file.spv : prints the reflected interface, one item per line
int main(int argc, char** argv) {
    if (argc < 3) { std::cerr << "usage: interface_check vertex|fragment file.spv\n"; return 2; }
    const bool vertex = std::string(argv[1]) == "vertex";
    std::ifstream f(argv[2], std::ios::binary);
    if (!f) { std::cerr << "cannot open " << argv[2] << "\n"; return 2; }
    SPIRV::TranslatorOptions options{};
    options.EnableAllExtensions();
    SPIRV::SPIRVModule* module = SPIRV::SPIRVModule::createSPIRVModule(options);
    f >> *module;
    if (!module->isModuleValid()) { std::cout << "valid 0\n"; return 1; }
    CompiledModule compiled;
    // ProcessVertexShader (Draw.cpp:785-791) / ProcessFragmentShader (:1613-1619)
    uint32_t inputSize = vertex ? 0u : (uint32_t)sizeof(VertexBuiltinOutput);
    uint32_t outputSize = vertex ? (uint32_t)sizeof(VertexBuiltinOutput) : 0u;
    std::vector<VariableInOutData> inputData{};
    std::vector<VariableUniformData> uniformData{};
    std::vector<VariableInOutData> outputData{};
    std::pair<void*, uint32_t> pushConstant{};
    GetVariablePointers(module, &compiled, inputData, uniformData, outputData, pushConstant, inputSize, outputSize);
    for (const auto& v : inputData) std::cout << "input " << v.location << " " << (int)v.format << " " << (int)v.interpolation << " " << v.size << " " << v.offset << "\n";
    for (const auto& v : outputData) std::cout << "output " << v.location << " " << (int)v.format << " " << v.size << " " << v.offset << "\n";
    for (const auto& v : uniformData) std::cout << "uniform " << v.set << " " << v.binding << "\n";
    std::cout << "push " << (pushConstant.first ? (int)pushConstant.second : -1) << "\n";
    std::cout << "sizes " << inputSize << " " << outputSize << "\n";
    return 0;
}
