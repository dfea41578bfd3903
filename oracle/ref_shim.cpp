// ref_shim.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin extern "C" shim over the *unmodified* reference sources that compile in this image:
//   /root/reference/torpedo/rendering/src/Camera.cpp                (Camera::lookAt, :3-16)
//   /root/reference/torpedo/extension/src/PerspectiveCamera.cpp     (updateProjectionMatrix, :9-23)
//   /root/reference/torpedo/math/include/torpedo/math/*.h           (compensated dot/cross/mul)
//   /root/reference/torpedo/volumetric/include/.../GaussianGeometry.h (GaussianPoint, rgb2sh)
//   /root/reference/torpedo/volumetric/src/GaussianGeometry.cpp      (GaussianPoint::fromModel, :59-127) + src/miniply.cpp
// The sources are compiled where they lie (see oracle/Makefile, target _ref); nothing is copied.
// The shim packs the camera UBO exactly like GaussianEngine::updateCameraBuffer
// (/root/reference/torpedo/volumetric/src/GaussianEngine.cpp:764-775).
#include <torpedo/extension/PerspectiveCamera.h>
#include <torpedo/math/transform.h>
#include <torpedo/volumetric/GaussianGeometry.h>

#include <cstring>

extern "C" {

// out34 = view(16, row-major) | proj*view (16) | focalNDC (2)
void tpdref_camera_ubo(uint32_t w, uint32_t h, const float eye[3], const float center[3], const float up[3],
                       float fovDegOrZero, float nearOrZero, float farOrZero, float out34[34]) {
    auto camera = tpd::PerspectiveCamera{ w, h };
    if (nearOrZero > 0.f) camera.setNear(nearOrZero);
    if (farOrZero > 0.f) camera.setFar(farOrZero);
    if (fovDegOrZero > 0.f || nearOrZero > 0.f || farOrZero > 0.f)
        camera.setVerticalFov(fovDegOrZero > 0.f ? fovDegOrZero : 60.f);
    camera.lookAt(tpd::vec3{ eye[0], eye[1], eye[2] }, tpd::vec3{ center[0], center[1], center[2] }, tpd::vec3{ up[0], up[1], up[2] });

    // == GaussianEngine::updateCameraBuffer
    auto projection = tpd::mat4{ camera.getProjectionData() };
    const auto fx = projection[0, 0];
    const auto fy = projection[1, 1];
    projection = tpd::math::mul(projection, camera.getViewMatrix());
    std::memcpy(out34, camera.getViewMatrixData(), 64);
    std::memcpy(out34 + 16, projection.data_ptr(), 64);
    out34[32] = fx;
    out34[33] = fy;
}

void tpdref_to_cartesian(float theta, float phi, float radius, float out3[3]) {
    const auto v = tpd::math::to_cartesian(theta, phi, radius);
    out3[0] = v.x; out3[1] = v.y; out3[2] = v.z;
}

void tpdref_rgb2sh(float r, float g, float b, float out48[48]) {
    const auto sh = tpd::utils::rgb2sh(r, g, b);
    std::memcpy(out48, sh.data(), sizeof(float) * 48);
}

// math::normalize(vec4) as used by GaussianPoint::fromModel (volumetric/src/GaussianGeometry.cpp:112-113)
void tpdref_normalize4(const float in4[4], float out4[4]) {
    const auto q = tpd::math::normalize(tpd::vec4{ in4[0], in4[1], in4[2], in4[3] });
    out4[0] = q.x; out4[1] = q.y; out4[2] = q.z; out4[3] = q.w;
}

uint32_t tpdref_sizeof_gaussian_point() { return sizeof(tpd::GaussianPoint); }

// GaussianPoint::fromModel (volumetric/src/GaussianGeometry.cpp:59-127) on a 3DGS .ply file: the reference's own reader
// (miniply) and field transforms. Returns the point count (records beyond `capacity` are not copied), or -1 when the
// reference throws.
int64_t tpdref_from_model(const char* ply_path, float* out_recs60, uint64_t capacity) {
    try {
        const auto points = tpd::GaussianPoint::fromModel(ply_path);
        static_assert(sizeof(tpd::GaussianPoint) == 240);
        const auto n = std::min<uint64_t>(points.size(), capacity);
        if (n && out_recs60) std::memcpy(out_recs60, points.data(), n * sizeof(tpd::GaussianPoint));
        return static_cast<int64_t>(points.size());
    } catch (const std::exception&) {
        return -1;
    }
}

} // extern "C"
