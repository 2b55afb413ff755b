// Minimal stand-in for glm >= 0.9.9 (absent from this image; the copy vendored under the reference's Samples/ is 0.9.5.3,
// which lacks vec<L, T> and static length()). Only what CPVulkan/ImageSampler.cpp uses when it is compiled in place for
// oracle/_ref/sampler_check: vec<1..4, T> with x/r component aliases, component-wise + - * /, converting constructors,
// min/max. Every operator is the plain per-component C++ operator on T — the semantics glm documents; the reference's
// glm version is unpinned. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstdint>
namespace glm {
template <int L, typename T> struct vec;
template <typename T> struct vec<1, T> {
    using value_type = T;
    union { T x, r, s; };
    static constexpr int length() { return 1; }
    vec() = default;
    template <typename A> vec(A a) : x(static_cast<T>(a)) {}
    template <typename U> vec(const vec<1, U>& o) : x(static_cast<T>(o.x)) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};
template <typename T> struct vec<2, T> {
    using value_type = T;
    union { T x, r, s; }; union { T y, g, t; };
    static constexpr int length() { return 2; }
    vec() = default;
    template <typename A> explicit vec(A a) : x(static_cast<T>(a)), y(static_cast<T>(a)) {}
    template <typename A, typename B> vec(A a, B b) : x(static_cast<T>(a)), y(static_cast<T>(b)) {}
    template <typename U> vec(const vec<2, U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};
template <typename T> struct vec<3, T> {
    using value_type = T;
    union { T x, r, s; }; union { T y, g, t; }; union { T z, b, p; };
    static constexpr int length() { return 3; }
    vec() = default;
    template <typename A> explicit vec(A a) : x(static_cast<T>(a)), y(static_cast<T>(a)), z(static_cast<T>(a)) {}
    template <typename A, typename B, typename C> vec(A a, B b_, C c) : x(static_cast<T>(a)), y(static_cast<T>(b_)), z(static_cast<T>(c)) {}
    template <typename U> vec(const vec<3, U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)), z(static_cast<T>(o.z)) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};
template <typename T> struct vec<4, T> {
    using value_type = T;
    union { T x, r, s; }; union { T y, g, t; }; union { T z, b, p; }; union { T w, a, q; };
    static constexpr int length() { return 4; }
    vec() = default;
    template <typename A> explicit vec(A v) : x(static_cast<T>(v)), y(static_cast<T>(v)), z(static_cast<T>(v)), w(static_cast<T>(v)) {}
    template <typename A, typename B, typename C, typename D> vec(A a_, B b_, C c, D d) : x(static_cast<T>(a_)), y(static_cast<T>(b_)), z(static_cast<T>(c)), w(static_cast<T>(d)) {}
    template <typename U> vec(const vec<4, U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)), z(static_cast<T>(o.z)), w(static_cast<T>(o.w)) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};
#define CPVK_GLM_OP(op)                                                                                                        \
    template <int L, typename T> vec<L, T> operator op(const vec<L, T>& a, const vec<L, T>& b) { vec<L, T> r; for (int i = 0; i < L; i++) r[i] = a[i] op b[i]; return r; } \
    template <int L, typename T> vec<L, T> operator op(const vec<L, T>& a, T b) { vec<L, T> r; for (int i = 0; i < L; i++) r[i] = a[i] op b; return r; }                \
    template <int L, typename T> vec<L, T> operator op(T a, const vec<L, T>& b) { vec<L, T> r; for (int i = 0; i < L; i++) r[i] = a op b[i]; return r; }
CPVK_GLM_OP(+) CPVK_GLM_OP(-) CPVK_GLM_OP(*) CPVK_GLM_OP(/)
#undef CPVK_GLM_OP
template <int L, typename T> vec<L, T> min(const vec<L, T>& a, const vec<L, T>& b) { vec<L, T> r; for (int i = 0; i < L; i++) r[i] = b[i] < a[i] ? b[i] : a[i]; return r; }
template <int L, typename T> vec<L, T> max(const vec<L, T>& a, const vec<L, T>& b) { vec<L, T> r; for (int i = 0; i < L; i++) r[i] = a[i] < b[i] ? b[i] : a[i]; return r; }
using fvec1 = vec<1, float>; using fvec2 = vec<2, float>; using fvec3 = vec<3, float>; using fvec4 = vec<4, float>;
using ivec1 = vec<1, int32_t>; using ivec2 = vec<2, int32_t>; using ivec3 = vec<3, int32_t>; using ivec4 = vec<4, int32_t>;
using uvec1 = vec<1, uint32_t>; using uvec2 = vec<2, uint32_t>; using uvec3 = vec<3, uint32_t>; using uvec4 = vec<4, uint32_t>;
using dvec4 = vec<4, double>;
using vec4 = fvec4; using vec3 = fvec3; using vec2 = fvec2;
}
