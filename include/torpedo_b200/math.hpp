// math.hpp — the sliver of torpedo/math the rasterizer's inputs depend on.
//
// The camera block uploaded to the GPU is computed on the host with torpedo's *compensated* fp32
// arithmetic (reference: torpedo/math/include/torpedo/math/common.h:11-40, vec3.h:215-252,
// vec4.h:248-265, mat4.h:257-264). Keys and tile counts are bit-exact functions of those matrices, so
// this header reproduces the same rounding sequence: Kahan's difference of products for cross(),
// the Dot2-style error-free transformations (TwoProduct via fma, TwoSum) for dot(), and
// normalize(v) = v * (1 / sqrt(dot(v, v))). Checked bit-for-bit against the reference's own sources
// by tests/test_host_layer.py (fixtures from oracle/_ref, tests/golden/cameras.json).
#pragma once

#include <array>
#include <cmath>
#include <cstddef>

namespace tpd {

struct vec3 {
    float x{}, y{}, z{};
};
struct vec4 {
    float x{}, y{}, z{}, w{};
};

constexpr vec3 operator-(const vec3& a, const vec3& b) noexcept { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
constexpr vec3 operator+(const vec3& a, const vec3& b) noexcept { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
constexpr vec3 operator*(const vec3& a, float s) noexcept { return { a.x * s, a.y * s, a.z * s }; }

/// Row-major 3x3 / 4x4 matrices (m[row][col]); the GPU consumes the 16 floats as they lie.
struct mat3 {
    std::array<std::array<float, 3>, 3> m{};
};
struct mat4 {
    std::array<std::array<float, 4>, 4> m{};

    constexpr mat4() noexcept = default;
    constexpr explicit mat4(float diagonal) noexcept {
        for (std::size_t i = 0; i < 4; ++i) m[i][i] = diagonal;
    }
    constexpr mat4(std::initializer_list<float> rowMajor) noexcept {
        std::size_t k = 0;
        for (const float v : rowMajor) {
            if (k >= 16) break;
            m[k / 4][k % 4] = v;
            ++k;
        }
    }
    /// [R | t; 0 0 0 1] — Camera(const mat3& R, const vec3& t) in the reference (Camera.h:14)
    constexpr mat4(const mat3& R, const vec3& t) noexcept {
        for (std::size_t i = 0; i < 3; ++i)
            for (std::size_t j = 0; j < 3; ++j) m[i][j] = R.m[i][j];
        m[0][3] = t.x; m[1][3] = t.y; m[2][3] = t.z; m[3][3] = 1.0f;
    }
    [[nodiscard]] const float* data_ptr() const noexcept { return m[0].data(); }
    [[nodiscard]] float* data_ptr() noexcept { return m[0].data(); }
};
static_assert(sizeof(mat4) == 64, "mat4 must be 16 packed floats");

namespace math {

struct ErrorFree {
    float value, error;
};

/// TwoProduct: a*b = value + error exactly (one fma).
inline ErrorFree two_product(float a, float b) noexcept {
    const float p = a * b;
    return { p, std::fma(a, b, -p) };
}
/// TwoSum (Knuth): a+b = value + error exactly.
inline ErrorFree two_sum(float a, float b) noexcept {
    const float s = a + b;
    const float bb = s - a;
    return { s, a - (s - bb) + (b - bb) };
}
/// Kahan: a*b - c*d with one rounding error recovered.
inline float diff_of_products(float a, float b, float c, float d) noexcept {
    const float cd = c * d;
    const float err = std::fma(-c, d, cd);
    const float dop = std::fma(a, b, -cd);
    return dop + err;
}

inline float dot(const vec3& u, const vec3& v) noexcept {
    const auto px = two_product(u.x, v.x);
    const auto py = two_product(u.y, v.y);
    const auto s1 = two_sum(px.value, py.value);
    const float e1 = px.error + (s1.error + py.error);
    const auto pz = two_product(u.z, v.z);
    const auto s2 = two_sum(s1.value, pz.value);
    const float e2 = e1 + (s2.error + pz.error);
    return s2.value + e2;
}

inline float dot(const vec4& u, const vec4& v) noexcept {
    const auto px = two_product(u.x, v.x);
    const auto py = two_product(u.y, v.y);
    const auto s1 = two_sum(px.value, py.value);
    const float e1 = px.error + (s1.error + py.error);
    const auto pz = two_product(u.z, v.z);
    const auto s2 = two_sum(s1.value, pz.value);
    const float e2 = e1 + (s2.error + pz.error);
    const auto pw = two_product(u.w, v.w);
    const auto s3 = two_sum(s2.value, pw.value);
    const float e3 = e2 + (s3.error + pw.error);
    return s3.value + e3;
}

inline vec3 cross(const vec3& u, const vec3& v) noexcept {
    return { diff_of_products(u.y, v.z, u.z, v.y), diff_of_products(u.z, v.x, u.x, v.z), diff_of_products(u.x, v.y, u.y, v.x) };
}

inline vec3 normalize(const vec3& v) noexcept {
    const float inv = 1.0f / std::sqrt(dot(v, v));
    return v * inv;
}

inline mat4 mul(const mat4& a, const mat4& b) noexcept {
    mat4 r;
    for (std::size_t i = 0; i < 4; ++i) {
        const vec4 row{ a.m[i][0], a.m[i][1], a.m[i][2], a.m[i][3] };
        for (std::size_t j = 0; j < 4; ++j) r.m[i][j] = dot(row, vec4{ b.m[0][j], b.m[1][j], b.m[2][j], b.m[3][j] });
    }
    return r;
}

/// Spherical (theta around +z, phi from +z) to cartesian — math/transform.h:10-16.
inline vec3 to_cartesian(float theta, float phi, float radius = 1.0f) noexcept {
    return { radius * std::sin(phi) * std::cos(theta), radius * std::sin(phi) * std::sin(theta), radius * std::cos(phi) };
}

}  // namespace math
}  // namespace tpd
