// Camera.hpp — tpd::Camera / tpd::PerspectiveCamera with the reference's semantics
// (torpedo/rendering/include/torpedo/rendering/Camera.h:11-38, rendering/src/Camera.cpp:3-16,
//  extension/include/torpedo/extension/PerspectiveCamera.h:9-44, extension/src/PerspectiveCamera.cpp:3-23).
#pragma once

#include "math.hpp"

#include <cstdint>
#include <numbers>

namespace tpd {

class Camera {
public:
    explicit Camera(const mat4& worldToCamera = mat4{ 1.0f }) noexcept : _view{ worldToCamera } {}
    Camera(const mat3& R, const vec3& t) noexcept : _view{ R, t } {}
    virtual ~Camera() = default;

    /// View basis: z forward, x right, y DOWN; translation -dot(axis, eye) (Camera.cpp:3-16).
    void lookAt(const vec3& eye, const vec3& center, const vec3& up) noexcept {
        const vec3 forward = math::normalize(center - eye);
        const vec3 right = math::normalize(math::cross(forward, up));
        const vec3 down = math::normalize(math::cross(forward, right));
        _view = mat4{
            right.x,   right.y,   right.z,   -math::dot(right, eye),
            down.x,    down.y,    down.z,    -math::dot(down, eye),
            forward.x, forward.y, forward.z, -math::dot(forward, eye),
            0.f,       0.f,       0.f,       1.f,
        };
    }
    void lookAt(const mat3& R, const vec3& t) noexcept { _view = mat4{ R, t }; }

    void setNear(float near) noexcept { _near = near; }
    void setFar(float far) noexcept { _far = far; }

    [[nodiscard]] const mat4& getViewMatrix() const noexcept { return _view; }
    [[nodiscard]] const float* getViewMatrixData() const noexcept { return _view.data_ptr(); }

    [[nodiscard]] virtual const float* getProjectionData() const noexcept = 0;
    [[nodiscard]] virtual uint32_t getProjectionByteSize() const noexcept = 0;
    virtual void onImageSizeChange(uint32_t, uint32_t) noexcept {}

protected:
    float _near{ 0.01f };  // Camera.h:33-34
    float _far{ 100.0f };

private:
    mat4 _view;
};

class PerspectiveCamera final : public Camera {
public:
    PerspectiveCamera(uint32_t imageWidth, uint32_t imageHeight)
        : _aspect{ static_cast<float>(imageWidth) / static_cast<float>(imageHeight) } {
        rebuild(std::numbers::sqrt3_v<float>);  // 60 degree vertical fov (PerspectiveCamera.h:27-32)
    }

    void setVerticalFov(float degrees) noexcept {
        const float fovY = degrees * std::numbers::pi_v<float> / 180.f;
        rebuild(1.f / std::tan(fovY * 0.5f));
    }

    [[nodiscard]] const float* getProjectionData() const noexcept override { return _projection.data_ptr(); }
    [[nodiscard]] uint32_t getProjectionByteSize() const noexcept override { return sizeof(float) * 16; }

    void onImageSizeChange(uint32_t w, uint32_t h) noexcept override {
        _aspect = static_cast<float>(w) / static_cast<float>(h);
        rebuild(_projection.m[1][1]);
    }

private:
    /// Reversed-z perspective: [near, far] -> [1, 0] (PerspectiveCamera.cpp:9-23)
    void rebuild(float fy) noexcept {
        const float fx = fy / _aspect;
        const float za = _near / (_near - _far);
        const float zb = _near * _far / (_far - _near);
        _projection = mat4{
            fx,  0.f, 0.f, 0.f,
            0.f, fy,  0.f, 0.f,
            0.f, 0.f, za,  zb,
            0.f, 0.f, 1.f, 0.f,
        };
    }

    float _aspect;
    mat4 _projection;
};

}  // namespace tpd
