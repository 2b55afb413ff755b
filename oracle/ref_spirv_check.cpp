// ref_spirv_check.cpp — loads a SPIR-V binary with the REFERENCE's own front end (the vendored libSPIRV under
// /root/reference/SPIRVParser, compiled in place by oracle/Makefile into oracle/_ref/) exactly the way
// CPVulkan/ShaderModule.cpp:62-73 does, and prints what the reference sees: validity, entry points and the
// module-order variable list (storage class, Location) that VS->FS linkage depends on.
// TEST INFRASTRUCTURE ONLY: used by tests/test_reference_spirv_frontend.py to pin the hand-assembled shaders.
#include <SPIRVModule.h>
#include <SPIRVFunction.h>
#include <SPIRVValue.h>
#include <SPIRVInstruction.h>

#include <fstream>
#include <iostream>
#include <sstream>

int main(int argc, char** argv) {
    if (argc < 2) { std::cerr << "usage: spirv_check file.spv\n"; return 2; }
    std::ifstream f(argv[1], std::ios::binary);
    if (!f) { std::cerr << "cannot open " << argv[1] << "\n"; return 2; }
    SPIRV::TranslatorOptions options{};
    options.EnableAllExtensions();
    SPIRV::SPIRVModule* module = SPIRV::SPIRVModule::createSPIRVModule(options);
    f >> *module;
    if (!module->isModuleValid()) {
        std::string msg; module->getError(msg);
        std::cout << "valid 0 " << msg << "\n";
        return 1;
    }
    std::cout << "valid 1\n";
    std::cout << "memory_model " << (int)module->getMemoryModel() << " addressing " << (int)module->getAddressingModel() << "\n";
    for (int model : {0, 4}) {
        const auto n = module->getNumEntryPoints((spv::ExecutionModel)model);
        for (unsigned i = 0; i < n; i++) {
            auto fn = module->getEntryPoint((spv::ExecutionModel)model, i);
            std::cout << "entry " << model << " " << module->getEntryPointName((spv::ExecutionModel)model, i) << " blocks " << fn->getNumBasicBlock() << "\n";
        }
    }
    std::cout << "functions " << module->getNumFunctions() << "\n";
    for (unsigned i = 0; i < module->getNumVariables(); i++) {
        auto v = module->getVariable(i);
        auto loc = v->getDecorate(spv::DecorationLocation);
        std::cout << "variable " << i << " storage " << (int)v->getStorageClass() << " location " << (loc.empty() ? -1 : (int)*loc.begin()) << " name " << v->getName() << "\n";
    }
    return 0;
}
