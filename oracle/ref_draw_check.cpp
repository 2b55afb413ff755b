// ref_draw_check.cpp — runs the REFERENCE's own rasteriser, interpolator and blend code: the functions of
// CPVulkan/CommandBuffer.Draw.cpp that decide which pixels a primitive covers and what every fragment carries
// (SURVEY §8(a) a2, a5, a6, a7, a11, a17):
//     EdgeFunction (:410-418), CalculatePrimitives (:567-673), SetDatum / GetFragmentInput (:816-954),
//     ApplyBlendFactor / ApplyBlend (:956-1262), DrawPixel / ProcessPoints / ProcessLines / ProcessTriangles (:1300-1594).
// That translation unit as a whole needs the entire ICD (LLVM-8, Vulkan SDK, xcb: SURVEY F10), so oracle/Makefile cuts
// exactly those line ranges out of the file where it lies under /root/reference (oracle/ref_slice.py, anchored on the
// first and last line of each range) into a scratch file outside the repository, and this file #includes it: the
// machine code of oracle/_ref/draw_check for those functions is compiled from the reference's own text.
// Around the slices stand: the reference's real CPVulkanBase headers (Base.h, Formats.h + Formats.cpp, PipelineState.h,
// PipelineData.h) and CPVulkan/DeviceState.h; the glm copy vendored under the reference's Samples/utils (0.9.5.3 — the
// only glm in the tree; glm::xy of gtx/vec_swizzle, absent from that version, is the two-line function below); oracle/shim
// for the Vulkan and GSL headers; and three stand-in classes defined here for what Pipeline.h would need LLVM for:
// GraphicsPipeline (a plain holder of the GraphicsPipelineStateStorage state blocks), FragmentShaderModule (entry point +
// originUpper) and an opaque ImageView. The "JIT-compiled fragment shader" is RecordFragment below: it writes down the
// arguments the reference calls the shader with and the interpolated inputs the reference stored for it.
// TEST INFRASTRUCTURE ONLY: tests/golden/make_ref_golden.py stores its output, tests/test_reference_draw.py compares the
// oracle's fragment stream and blend results with it.
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <unordered_map>
#include <vector>

#include <Base.h>
#define private public // Buffer keeps its fields private and offers no constructor outside the ICD ("ia" mode binds an index buffer)
#include <Buffer.h>
#undef private
#include <DeviceState.h>
#include <Formats.h>
#include <PipelineData.h>
#include <PipelineState.h>

#include <glm/glm.hpp>

namespace glm {
// gtx/vec_swizzle.hpp (glm >= 0.9.9): xy(v) = vec2(v.x, v.y)
template <typename T, precision P> detail::tvec2<T, P> xy(const detail::tvec4<T, P>& v) { return detail::tvec2<T, P>(v.x, v.y); }
}

// `EdgeFunction(p0, p1, p2)` at Draw.cpp:879 passes a vec4 where the function takes a vec2: with glm >= 0.9.9 that is the
// implicit converting constructor vec2(vec4) = (x, y) (explicit in the vendored 0.9.5.3). Spelled out here; the callee is
// the reference's own EdgeFunction(vec4, vec4, vec2) from the slice.
static float EdgeFunction(const glm::vec4& a, const glm::vec4& b, const glm::vec2& c);
static inline float EdgeFunction(const glm::vec4& a, const glm::vec4& b, const glm::vec4& c) { return EdgeFunction(a, b, glm::vec2(c)); }

namespace SPIRV { class SPIRVType; }

ImageFunctions::ImageFunctions(CPJit* j) : jit(j) {}
ImageFunctions::~ImageFunctions() = default;

using EntryPoint = void (*)();

class FragmentShaderModule {
public:
    FragmentShaderModule(EntryPoint e, bool upper) : entryPoint(e), originUpper(upper) {}
    EntryPoint getEntryPoint() const { return entryPoint; }
    bool getOriginUpper() const { return originUpper; }
private:
    EntryPoint entryPoint;
    bool originUpper;
};

class GraphicsPipeline final : public GraphicsPipelineStateStorage {
public:
    const PipelineLayout* getLayout() const override { return nullptr; }
    const VertexInputState& getVertexInputState() const override { return vertexInputState; }
    const InputAssemblyState& getInputAssemblyState() const override { return inputAssemblyState; }
    const TessellationState& getTessellationState() const override { return tessellationState; }
    const ViewportState& getViewportState() const override { return viewportState; }
    const RasterizationState& getRasterizationState() const override { return rasterizationState; }
    const MultisampleState& getMultisampleState() const override { return multisampleState; }
    const DepthStencilState& getDepthStencilState() const override { return depthStencilState; }
    const ColourBlendState& getColourBlendState() const override { return colourBlendState; }
    const DynamicState& getDynamicState() const override { return dynamicState; }
    const std::vector<AttachmentDescription>& getAttachments() const override { return attachments; }
    const SubpassDescription& getSubpass() const override { return subpass; }
    const std::vector<SubpassDependency>& getDependencies() const override { return dependencies; }

    VertexInputState vertexInputState{};
    InputAssemblyState inputAssemblyState{};
    TessellationState tessellationState{};
    ViewportState viewportState{};
    RasterizationState rasterizationState{};
    MultisampleState multisampleState{};
    DepthStencilState depthStencilState{};
    ColourBlendState colourBlendState{};
    DynamicState dynamicState{};
    std::vector<AttachmentDescription> attachments{};
    SubpassDescription subpass{};
    std::vector<SubpassDependency> dependencies{};
};

#include "draw_slices.inc" // written by oracle/ref_slice.py into the scratch build directory (-I)

// ---- the checker proper ----

struct InputSpec { uint32_t offset, format, interpolation, size; };

static std::vector<uint32_t>* g_out;
static FragmentBuiltinInput* g_builtinInput;
static std::vector<VariableInOutData>* g_inputs;
static uint32_t g_fragments;

// what the reference calls per fragment in place of the JIT-compiled wrapper: main(depth, x, y, front) (Draw.cpp:1312)
static void RecordFragment(float depth, uint32_t x, uint32_t y, bool front) {
    uint32_t w[8];
    w[0] = x; w[1] = y; w[2] = front ? 1u : 0u;
    std::memcpy(&w[3], &depth, 4);
    std::memcpy(&w[4], &g_builtinInput->fragCoord, 16);
    g_out->insert(g_out->end(), w, w + 8);
    for (const VariableInOutData& in : *g_inputs) {
        const uint32_t* p = static_cast<const uint32_t*>(in.pointer);
        g_out->insert(g_out->end(), p, p + in.size / 4);
    }
    g_fragments++;
}

template <typename T> static bool Read(std::ifstream& in, T* v, size_t n = 1) { in.read(reinterpret_cast<char*>(v), (std::streamsize)(sizeof(T) * n)); return (bool)in; }

// raster <input> <output>
// input:  u32 nCases, then per case
//           f32 W, H, minDepth, maxDepth, lineWidth; u32 topology, frontFace, cullMode, originUpper, dynamicViewport,
//           vertexCount, stride, inputCount; inputCount x {u32 offset, format, interpolation, size}; vertexCount*stride bytes
//         The bytes are the vertex stage's output records {vec4 position, f32 pointSize, f32 clip[1], outputs...}
//         (PipelineData.h:4-9 + PipelineCompiler.cpp:532-547), indexed by raw vertex position as Draw.cpp does.
// output: per case u32 nFragments, wordsPerFragment; then per fragment, in the order the reference emits them,
//           x, y, front, depth (as passed to the shader = after the viewport depth transform), fragCoord[4], input words
static int MainRaster(const char* inPath, const char* outPath) {
    std::ifstream in(inPath, std::ios::binary);
    if (!in) return 2;
    std::ofstream out(outPath, std::ios::binary);
    uint32_t nCases;
    if (!Read(in, &nCases)) return 2;
    for (uint32_t c = 0; c < nCases; c++) {
        float f[5]; uint32_t u[8];
        if (!Read(in, f, 5) || !Read(in, u, 8)) return 2;
        const uint32_t vertexCount = u[5], stride = u[6], inputCount = u[7];
        std::vector<InputSpec> specs(inputCount);
        if (inputCount && !Read(in, specs.data(), inputCount)) return 2;

        auto state = std::make_unique<DeviceState>();
        state->jit = nullptr;
        GraphicsPipeline pipeline;
        const VkViewport viewport{0, 0, f[0], f[1], f[2], f[3]};
        const VkViewport decoy{0, 0, 1, 1, 0.25f, 0.5f};
        pipeline.inputAssemblyState.Topology = static_cast<VkPrimitiveTopology>(u[0]);
        pipeline.rasterizationState.FrontFace = static_cast<VkFrontFace>(u[1]);
        pipeline.rasterizationState.CullMode = u[2];
        pipeline.rasterizationState.LineWidth = f[4];
        pipeline.rasterizationState.LineRasterizationMode = VK_LINE_RASTERIZATION_MODE_DEFAULT_EXT;
        pipeline.dynamicState.DynamicViewport = u[4] != 0;
        pipeline.viewportState.Viewports.push_back(u[4] ? decoy : viewport);
        state->graphicsPipelineState.dynamicState.viewports[0] = u[4] ? viewport : decoy;
        state->graphicsPipelineState.pipeline = &pipeline;
        state->graphicsPipelineState.vertexOutputStorage.resize((size_t)vertexCount * stride);
        if (vertexCount && !Read(in, state->graphicsPipelineState.vertexOutputStorage.data(), (size_t)vertexCount * stride)) return 2;

        // the storage the reference interpolates INTO (the shader's `_input_*` globals, Draw.cpp:459)
        std::vector<std::vector<uint32_t>> inputStorage(inputCount);
        std::vector<VariableInOutData> inputData(inputCount);
        uint32_t words = 8;
        for (uint32_t i = 0; i < inputCount; i++) {
            inputStorage[i].assign(specs[i].size / 4 + 4, 0);
            inputData[i].pointer = inputStorage[i].data();
            inputData[i].location = i;
            inputData[i].format = static_cast<VkFormat>(specs[i].format);
            inputData[i].type = nullptr;
            inputData[i].interpolation = static_cast<InterpolationType>(specs[i].interpolation);
            inputData[i].size = specs[i].size;
            inputData[i].offset = specs[i].offset;
            words += specs[i].size / 4;
        }

        AssemblerOutput assembler{};
        assembler.vertices.resize(vertexCount);
        for (uint32_t i = 0; i < vertexCount; i++) assembler.vertices[i] = VertexInput{i, i}; // Draw.cpp:675-688
        CalculatePrimitives(state.get(), assembler);

        const VertexOutput vertexOutput{sizeof(VertexBuiltinOutput), stride, vertexCount};
        FragmentBuiltinInput builtinInput{};
        builtinInput.fragCoord = glm::vec4(0, 0, 0, 1); // Draw.cpp:1673
        FragmentBuiltinOutput builtinOutput{};
        const FragmentShaderModule shader(reinterpret_cast<EntryPoint>(RecordFragment), u[3] != 0);
        std::vector<uint32_t> stream;
        g_out = &stream; g_builtinInput = &builtinInput; g_inputs = &inputData; g_fragments = 0;
        std::pair<AttachmentDescription, ImageView*> none{};
        std::vector<std::pair<AttachmentDescription, ImageView*>> images;
        std::vector<VariableInOutData> outputData;
        switch (assembler.primitiveType) { // Draw.cpp:1680-1696
        case PrimitiveType::Point: ProcessPoints(state.get(), assembler, &builtinInput, &builtinOutput, &shader, none, none, images, outputData, vertexOutput, pipeline.rasterizationState, inputData); break;
        case PrimitiveType::Line: ProcessLines(state.get(), assembler, &builtinInput, &builtinOutput, &shader, none, none, images, outputData, vertexOutput, pipeline.rasterizationState, inputData); break;
        case PrimitiveType::Triangle: ProcessTriangles(state.get(), assembler, &builtinInput, &builtinOutput, &shader, none, none, images, outputData, vertexOutput, pipeline.rasterizationState, inputData); break;
        }
        const uint32_t hdr[2] = {g_fragments, words};
        out.write(reinterpret_cast<const char*>(hdr), 8);
        out.write(reinterpret_cast<const char*>(stream.data()), (std::streamsize)(stream.size() * 4));
    }
    return out ? 0 : 2;
}

// ia <input> <output>
// input:  u32 nCases, then per case u32 {indexed, first, count, vertexOffset, indexStride, topology, bindingOffset, bufferBytes} and
//         bufferBytes bytes (the index buffer; none when not indexed).
// output: per case u32 nVertices, then nVertices x {rawId, vertexId}: what ProcessInputAssembler / ProcessInputAssemblerIndexed
//         (Draw.cpp:675-760) hand to the vertex stage.
static int MainIa(const char* inPath, const char* outPath) {
    std::ifstream in(inPath, std::ios::binary);
    if (!in) return 2;
    std::ofstream out(outPath, std::ios::binary);
    uint32_t nCases;
    if (!Read(in, &nCases)) return 2;
    for (uint32_t c = 0; c < nCases; c++) {
        uint32_t u[8];
        if (!Read(in, u, 8)) return 2;
        std::vector<uint8_t> bytes(u[7] + 16);
        if (u[7] && !Read(in, bytes.data(), u[7])) return 2;
        auto state = std::make_unique<DeviceState>();
        state->jit = nullptr;
        GraphicsPipeline pipeline;
        pipeline.inputAssemblyState.Topology = static_cast<VkPrimitiveTopology>(u[5]);
        pipeline.inputAssemblyState.PrimitiveRestartEnable = false; // the reference aborts on it (Draw.cpp:697-700)
        state->graphicsPipelineState.pipeline = &pipeline;
        Buffer buffer;
        buffer.data = gsl::span<uint8_t>(bytes.data(), (std::ptrdiff_t)bytes.size());
        buffer.size = bytes.size();
        state->graphicsPipelineState.indexBinding = &buffer;           // vkCmdBindIndexBuffer, CommandBuffer.cpp
        state->graphicsPipelineState.indexBindingOffset = u[6];
        state->graphicsPipelineState.indexBindingStride = u[4];
        const AssemblerOutput a = u[0] ? ProcessInputAssemblerIndexed(state.get(), u[1], u[2], u[3]) : ProcessInputAssembler(state.get(), u[1], u[2]);
        const uint32_t n = (uint32_t)a.vertices.size();
        out.write(reinterpret_cast<const char*>(&n), 4);
        for (const VertexInput& v : a.vertices) { const uint32_t w[2] = {v.rawId, v.vertexId}; out.write(reinterpret_cast<const char*>(w), 8); }
    }
    return out ? 0 : 2;
}

// blend <input> <output>
// input:  u32 nCases, then per case 8 x u32 (VkPipelineColorBlendAttachmentState in member order) + source[4], destination[4],
//         constant[4] as f32.  output: per case ApplyBlend<glm::vec4>(source, destination, constant, state) as 4 x f32 bits.
static int MainBlend(const char* inPath, const char* outPath) {
    std::ifstream in(inPath, std::ios::binary);
    if (!in) return 2;
    std::ofstream out(outPath, std::ios::binary);
    uint32_t nCases;
    if (!Read(in, &nCases)) return 2;
    for (uint32_t c = 0; c < nCases; c++) {
        uint32_t s[8]; float v[12];
        if (!Read(in, s, 8) || !Read(in, v, 12)) return 2;
        VkPipelineColorBlendAttachmentState b{};
        b.blendEnable = s[0];
        b.srcColorBlendFactor = static_cast<VkBlendFactor>(s[1]); b.dstColorBlendFactor = static_cast<VkBlendFactor>(s[2]);
        b.colorBlendOp = static_cast<VkBlendOp>(s[3]);
        b.srcAlphaBlendFactor = static_cast<VkBlendFactor>(s[4]); b.dstAlphaBlendFactor = static_cast<VkBlendFactor>(s[5]);
        b.alphaBlendOp = static_cast<VkBlendOp>(s[6]);
        b.colorWriteMask = s[7];
        const glm::vec4 r = ApplyBlend<glm::vec4>(glm::vec4(v[0], v[1], v[2], v[3]), glm::vec4(v[4], v[5], v[6], v[7]), glm::vec4(v[8], v[9], v[10], v[11]), b);
        out.write(reinterpret_cast<const char*>(&r), 16);
    }
    return out ? 0 : 2;
}

int main(int argc, char** argv) {
    if (argc == 4 && !std::strcmp(argv[1], "raster")) return MainRaster(argv[2], argv[3]);
    if (argc == 4 && !std::strcmp(argv[1], "blend")) return MainBlend(argv[2], argv[3]);
    if (argc == 4 && !std::strcmp(argv[1], "ia")) return MainIa(argv[2], argv[3]);
    std::fprintf(stderr, "usage: draw_check raster|blend|ia <input> <output>\n");
    return 2;
}
