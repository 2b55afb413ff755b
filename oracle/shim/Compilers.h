// Stand-in for LLVMRuntime/Compilers.h (needs LLVM 8, absent): only the declarations CPVulkan/ImageSampler.cpp calls.
// The definitions live in oracle/ref_sampler_check.cpp, which returns a raw RGBA32F load/store as the "JIT-compiled"
// texel function — for R32G32B32A32_SFLOAT that is what ImageCompiler.cpp emits (same-type load, no conversion).
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include <Formats.h>
class CPJit;
using FunctionPointer = void (*)();
FunctionPointer CompileGetPixelDepth(CPJit* jit, const FormatInformation* information);
FunctionPointer CompileGetPixelStencil(CPJit* jit, const FormatInformation* information);
FunctionPointer CompileGetPixelF32(CPJit* jit, const FormatInformation* information);
FunctionPointer CompileGetPixelI32(CPJit* jit, const FormatInformation* information);
FunctionPointer CompileGetPixelU32(CPJit* jit, const FormatInformation* information);
FunctionPointer CompileSetPixelDepthStencil(CPJit* jit, const FormatInformation* information);
FunctionPointer CompileSetPixelF32(CPJit* jit, const FormatInformation* information);
FunctionPointer CompileSetPixelI32(CPJit* jit, const FormatInformation* information);
FunctionPointer CompileSetPixelU32(CPJit* jit, const FormatInformation* information);
