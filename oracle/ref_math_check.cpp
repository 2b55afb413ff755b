// ref_math_check.cpp — runs the REFERENCE's own shader runtime math: the functions JIT-compiled shaders call by
// name for OpDot / OpMatrixTimes* (LLVMRuntime/SpirvFunctions.cpp, compiled here as a whole translation unit, reached through
// its own name table getSpirvFunctions()) and for GLSL.std.450 (the templates of CPVulkan/GlslFunctions.cpp:19-321, lifted out
// of that file at build time by oracle/ref_slice.py — the rest of it needs the whole ICD), against the glm copy vendored under
// the reference's Samples/utils (0.9.5.3, the only glm in the tree; the reference's build needs glm >= 0.9.9, whose normalize
// sums the squares pairwise for 4 components where 0.9.5.3 sums left to right — only vec4 Normalise can differ).
// The per-component vector forms (VAbs, VMin, ...) loop over T::length(), static in glm >= 0.9.9 only: they are instantiated
// with a plain four-lane container (Lanes<T, N>: storage and operator[] — no arithmetic of its own).
// TEST INFRASTRUCTURE ONLY: tests/golden/make_ref_golden.py stores its output, tests/test_reference_math.py compares the oracle
// (cpvk_oracle_math) with it.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <limits>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include <SpirvFunctions.cpp> // /root/reference/LLVMRuntime, whole translation unit

template <typename T, int N> struct Lanes {
    T v[N];
    Lanes() = default;
    Lanes(T x) { for (int i = 0; i < N; i++) v[i] = x; }
    static constexpr int length() { return N; }
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
};

#include "glsl_slices.inc" // CPVulkan/GlslFunctions.cpp:19-321, written by oracle/ref_slice.py into the scratch build directory

template <typename T> static void Put(uint32_t* out, const T& v, int n) { std::memcpy(out, &v, (size_t)n * 4); }

// One case: kind, op, p0, p1, p2, then a[16], b[16], c[16] (raw 32-bit lanes). Result: 16 lanes.
//   kind 0: GLSL.std.450 instruction `op` on n = p1 lanes of type p0 (0 float, 1 signed, 2 unsigned): the scalar template for
//           n == 1, the V* template over Lanes<T, n> otherwise; Normalise / Reflect on glm vectors of n lanes
//   kind 1: @Vector.Dot.F32.F32[n].F32[n]                         kind 2: @Matrix.Mult mat(n x n) * scalar
//   kind 3: @Matrix.Mult vec4 * mat4                              kind 4: @Matrix.Mult mat(n x n) * vec(n)      kind 5: mat4 * mat4
template <typename T, int N> static bool Glsl(uint32_t op, const uint32_t* a, const uint32_t* b, const uint32_t* c, uint32_t* out) {
    using V = Lanes<T, N>;
    V x, y, z; std::memcpy(&x, a, N * 4); std::memcpy(&y, b, N * 4); std::memcpy(&z, c, N * 4);
    V r{};
    const bool isFloat = std::is_floating_point<T>::value;
    switch (op) {
    case 4: case 5: if constexpr (std::is_signed<T>::value) { if (N == 1) r[0] = Abs(x[0]); else r = VAbs(x); } else return false; break;
    case 7: if constexpr (std::is_signed<T>::value) { if (N == 1) r[0] = SSign(x[0]); else r = VSSign(x); } else return false; break;
    case 37: case 38: case 39: if (N == 1) r[0] = Min(x[0], y[0]); else r = VMin(x, y); break;
    case 40: case 41: case 42: if (N == 1) r[0] = Max(x[0], y[0]); else r = VMax(x, y); break;
    case 43: case 44: case 45: if (N == 1) r[0] = Clamp(x[0], y[0], z[0]); else r = VClamp(x, y, z); break;
    default:
        if (!isFloat) return false;
    }
    if (op == 4 || op == 5 || op == 7 || (op >= 37 && op <= 45)) { Put(out, r, N); return true; }
    return false;
}
template <int N> static bool GlslFloat(uint32_t op, const uint32_t* a, const uint32_t* b, const uint32_t* c, uint32_t* out) {
    if (Glsl<float, N>(op, a, b, c, out)) return true;
    using V = Lanes<float, N>;
    V x, y, z; std::memcpy(&x, a, N * 4); std::memcpy(&y, b, N * 4); std::memcpy(&z, c, N * 4);
    V r{};
    switch (op) {
    case 13: if (N == 1) r[0] = Sin(x[0]); else r = VSin(x); break;
    case 14: if (N == 1) r[0] = Cos(x[0]); else r = VCos(x); break;
    case 26: if (N == 1) r[0] = Pow(x[0], y[0]); else r = VPow(x, y); break;
    case 46: if (N == 1) r[0] = Mix(x[0], y[0], z[0]); else r = VMix(x, y, z); break;
    case 79: if (N == 1) r[0] = NMin(x[0], y[0]); else r = VNMin(x, y); break;
    case 80: if (N == 1) r[0] = NMax(x[0], y[0]); else r = VNMax(x, y); break;
    case 81: if (N == 1) r[0] = NClamp(x[0], y[0], z[0]); else r = VNClamp(x, y, z); break;
    default: return false;
    }
    Put(out, r, N);
    return true;
}
template <typename G, int N> static void GlmOp(uint32_t op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    G x, y; std::memcpy(&x, a, N * 4); std::memcpy(&y, b, N * 4);
    const G r = op == 69 ? VNormalise(x) : Reflect(x, y);
    std::memcpy(out, &r, N * 4);
}

int main(int argc, char** argv) {
    if (argc != 3) { std::fprintf(stderr, "usage: math_check <input> <output>\n"); return 2; }
    std::ifstream in(argv[1], std::ios::binary);
    std::ofstream outf(argv[2], std::ios::binary);
    uint32_t nCases;
    in.read(reinterpret_cast<char*>(&nCases), 4);
    if (!in) return 2;
    auto& table = getSpirvFunctions();
    auto fn = [&](const char* name) -> FunctionPointer { auto it = table.find(name); if (it == table.end()) { std::fprintf(stderr, "math_check: the reference has no %s\n", name); std::abort(); } return it->second; };
    for (uint32_t i = 0; i < nCases; i++) {
        uint32_t h[5], a[16], b[16], c[16], out[16] = {0};
        in.read(reinterpret_cast<char*>(h), 20); in.read(reinterpret_cast<char*>(a), 64); in.read(reinterpret_cast<char*>(b), 64); in.read(reinterpret_cast<char*>(c), 64);
        if (!in) return 2;
        const uint32_t kind = h[0], op = h[1], n = h[3];
        bool ok = true;
        if (kind == 0) {
            if (op == 69 || op == 71) {
                if (n == 2) GlmOp<glm::vec2, 2>(op, a, b, out); else if (n == 3) GlmOp<glm::vec3, 3>(op, a, b, out); else if (n == 4) GlmOp<glm::vec4, 4>(op, a, b, out); else ok = false;
            } else if (h[2] == 0) {
                ok = n == 1 ? GlslFloat<1>(op, a, b, c, out) : n == 2 ? GlslFloat<2>(op, a, b, c, out) : n == 3 ? GlslFloat<3>(op, a, b, c, out) : GlslFloat<4>(op, a, b, c, out);
            } else if (h[2] == 1) {
                ok = n == 1 ? Glsl<int32_t, 1>(op, a, b, c, out) : n == 2 ? Glsl<int32_t, 2>(op, a, b, c, out) : n == 3 ? Glsl<int32_t, 3>(op, a, b, c, out) : Glsl<int32_t, 4>(op, a, b, c, out);
            } else {
                ok = n == 1 ? Glsl<uint32_t, 1>(op, a, b, c, out) : n == 2 ? Glsl<uint32_t, 2>(op, a, b, c, out) : n == 3 ? Glsl<uint32_t, 3>(op, a, b, c, out) : Glsl<uint32_t, 4>(op, a, b, c, out);
            }
        } else if (kind == 1) {
            const char* names[5] = {nullptr, nullptr, "@Vector.Dot.F32.F32[2].F32[2]", "@Vector.Dot.F32.F32[3].F32[3]", "@Vector.Dot.F32.F32[4].F32[4]"};
            if (n < 2 || n > 4) ok = false;
            else { const float r = reinterpret_cast<float (*)(const void*, const void*)>(fn(names[n]))(a, b); std::memcpy(out, &r, 4); }
        } else if (kind == 2) {
            const char* names[5] = {nullptr, nullptr, "@Matrix.Mult.F32[2,2,col].F32[2,2,col].F32", "@Matrix.Mult.F32[3,3,col].F32[3,3,col].F32", "@Matrix.Mult.F32[4,4,col].F32[4,4,col].F32"};
            if (n < 2 || n > 4) ok = false;
            else { float s; std::memcpy(&s, b, 4); reinterpret_cast<void (*)(void*, const void*, float)>(fn(names[n]))(out, a, s); }
        } else if (kind == 3) {
            reinterpret_cast<void (*)(void*, void*, const void*)>(fn("@Matrix.Mult.F32[4].F32[4].F32[4,4,col]"))(out, a, b); // (result, vector, matrix)
        } else if (kind == 4) {
            if (n == 3) reinterpret_cast<void (*)(void*, const void*, void*)>(fn("@Matrix.Mult.F32[3].F32[3,3,col].F32[3]"))(out, a, b);
            else if (n == 4) reinterpret_cast<void (*)(void*, const void*, void*)>(fn("@Matrix.Mult.F32[4].F32[4,4,col].F32[4]"))(out, a, b);
            else ok = false;
        } else if (kind == 5) {
            reinterpret_cast<void (*)(void*, const void*, void*)>(fn("@Matrix.Mult.F32[4,4,col].F32[4,4,col].F32[4,4,col]"))(out, a, b);
        } else ok = false;
        if (!ok) { std::fprintf(stderr, "math_check: case %u (kind %u op %u type %u n %u) has no reference function\n", i, kind, op, h[2], n); return 3; }
        outf.write(reinterpret_cast<const char*>(out), 64);
    }
    return outf ? 0 : 2;
}
