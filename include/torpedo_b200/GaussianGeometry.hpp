// GaussianGeometry.hpp — tpd::GaussianPoint, the 240-byte input record of the rasterizer
// (torpedo/volumetric/include/torpedo/volumetric/GaussianGeometry.h:10-52, GPU twin splat.slang:24-31).
#pragma once

#include "math.hpp"

#include <array>
#include <cstdint>
#include <filesystem>
#include <random>
#include <vector>

namespace tpd {

struct GaussianPoint {
    static constexpr uint32_t MAX_SH_FLOATS = 48;  // 3 * 16

    vec3 position;
    float opacity;
    vec4 quaternion;  // (x, y, z, w): w LAST, not re-normalised on the GPU (splat/volume.slang:29)
    vec4 scale;       // w = scale modifier
    std::array<float, MAX_SH_FLOATS> sh;  // DC rgb, then 15 R, 15 G, 15 B (splat/common.slang:25-31)

    /// Same distribution as the reference (GaussianGeometry.cpp:8-32): positions U(-1,1)^3 * radius + center,
    /// identity rotation, per-axis scale U(minScale,maxScale), opacity U(minOpacity,maxOpacity), DC colour U(0,1).
    /// The reference seeds from std::random_device; pass `seed` for a reproducible cloud.
    [[nodiscard]] static std::vector<GaussianPoint> random(uint32_t count, float radius = 1.0f, const vec3& center = { 0.f, 0.f, 0.f },
                                                           float minScale = 0.1f, float maxScale = 1.0f, float minOpacity = 0.1f,
                                                           float maxOpacity = 1.0f, uint64_t seed = std::random_device{}());

    /// 3DGS PLY loader (GaussianGeometry.cpp:59-127): see PlyLoader.hpp.
    [[nodiscard]] static std::vector<GaussianPoint> fromModel(const std::filesystem::path& plyFile);
};
static_assert(sizeof(GaussianPoint) == 240 && alignof(GaussianPoint) == 4, "GaussianPoint is the 240-byte wire format");

namespace utils {
[[nodiscard]] constexpr std::array<float, GaussianPoint::MAX_SH_FLOATS> rgb2sh(float r, float g, float b) noexcept {
    constexpr float C0 = 0.28209479177387814f;
    std::array<float, GaussianPoint::MAX_SH_FLOATS> sh{};
    sh[0] = (r - 0.5f) / C0;
    sh[1] = (g - 0.5f) / C0;
    sh[2] = (b - 0.5f) / C0;
    return sh;
}
[[nodiscard]] constexpr vec3 sh2rgb(const std::array<float, GaussianPoint::MAX_SH_FLOATS>& sh) noexcept {
    constexpr float C0 = 0.28209479177387814f;
    return { sh[0] * C0 + 0.5f, sh[1] * C0 + 0.5f, sh[2] * C0 + 0.5f };
}
}  // namespace utils

inline std::vector<GaussianPoint> GaussianPoint::random(uint32_t count, float radius, const vec3& center, float minScale, float maxScale,
                                                        float minOpacity, float maxOpacity, uint64_t seed) {
    // counter-based splitmix64 -> 24-bit uniforms: identical on every platform, unlike std::uniform_real_distribution
    auto u01 = [seed](uint64_t field, uint64_t index) {
        auto mix = [](uint64_t z) {
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            return z ^ (z >> 31);
        };
        const uint64_t key = mix(seed * 0x9E3779B97F4A7C15ull + field * 0xD1B54A32D192ED03ull);
        return static_cast<float>(mix(index * 0x9E3779B97F4A7C15ull + key) >> 40) * (1.0f / 16777216.0f);
    };
    std::vector<GaussianPoint> points(count);
    for (uint32_t i = 0; i < count; ++i) {
        auto& p = points[i];
        p.position = vec3{ u01(1, i) * 2.f - 1.f, u01(2, i) * 2.f - 1.f, u01(3, i) * 2.f - 1.f } * radius + center;
        p.opacity = minOpacity + u01(4, i) * (maxOpacity - minOpacity);
        p.quaternion = { 0.f, 0.f, 0.f, 1.f };
        p.scale = { minScale + u01(5, i) * (maxScale - minScale), minScale + u01(6, i) * (maxScale - minScale),
                    minScale + u01(7, i) * (maxScale - minScale), 1.f };
        p.sh = utils::rgb2sh(u01(8, i), u01(9, i), u01(10, i));
    }
    return points;
}

}  // namespace tpd
