// GaussianEngine.hpp — Vulkan-free drop-in for tpd::GaussianEngine's public surface
// (torpedo/volumetric/include/torpedo/volumetric/GaussianEngine.h:15-29):
//
//     compile(scene, settings) · getTransformHost() · rasterFrame(camera) · draw(...) · waitIdle()
//
// Same names, argument meaning and error behaviour (C++ exceptions; main-thread only); the nine Slang
// dispatches behind rasterFrame are replaced by libtpdcu's CUDA kernels through include/tpdcu.h.
// What the reference obtains from its Renderer — the framebuffer size and the swap image — is passed
// explicitly: the constructor / resize() take the size (GaussianEngine.cpp:99,206-220), draw() copies
// the finished frame to host memory, and bindSwapTarget() attaches a Vulkan allocation exported as an
// opaque fd so that draw()'s target->swap-image copy (GaussianEngine.cpp:865-875) can read CUDA's
// output directly.
#pragma once

#include "../tpdcu.h"
#include "Camera.hpp"
#include "GaussianGeometry.hpp"
#include "PlyLoader.hpp"
#include "Scene.hpp"
#include "TransformHost.hpp"

#include <array>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

namespace tpd {

class GaussianEngine final {
public:
    struct Settings {
        uint32_t sphericalHarmonicsDegree{ 3 };
        [[nodiscard]] static constexpr Settings getDefault() { return {}; }
    };

    GaussianEngine(uint32_t framebufferWidth, uint32_t framebufferHeight, int cudaDevice = 0) {
        check(tpdcu_create(cudaDevice, &_ctx), "tpdcu_create");
        _transformHost = std::make_unique<TransformHost>(_ctx);
        try {
            resize(framebufferWidth, framebufferHeight);
        } catch (...) {
            tpdcu_destroy(_ctx);
            throw;
        }
    }
    GaussianEngine(const GaussianEngine&) = delete;
    GaussianEngine& operator=(const GaussianEngine&) = delete;
    ~GaussianEngine() noexcept { tpdcu_destroy(_ctx); }

    /// GaussianEngine::compile (GaussianEngine.cpp:359-399): snapshot every GaussianPoint of the scene (groups
    /// first, then singles), clamp the SH degree to 3, build the per-Gaussian entity index, identity transforms.
    /// An empty scene is a warning + no-op in the reference; here it is a silent no-op.
    void compile(const Scene& scene, const Settings& settings = Settings::getDefault()) {
        const uint32_t gaussianCount = scene.countAll<GaussianPoint>();
        if (gaussianCount == 0) return;
        _shDegree = settings.sphericalHarmonicsDegree > 3 ? 3 : settings.sphericalHarmonicsDegree;

        auto entityMap = scene.buildEntityMap<GaussianPoint>();
        const auto entityCount = static_cast<uint32_t>(entityMap.size());
        std::vector<uint32_t> indices;
        indices.reserve(gaussianCount);
        uint32_t index = 0;
        for (const uint32_t size : scene.groupSizes<GaussianPoint>()) indices.insert(indices.end(), size, index++);
        for (uint32_t i = 0; i < scene.count<GaussianPoint>(); ++i) indices.push_back(index++);

        const auto bytes = scene.dataAll<GaussianPoint>();
        check(tpdcu_upload_gaussians(_ctx, bytes.data(), gaussianCount, indices.data(), entityCount), "tpdcu_upload_gaussians");
        _transformHost->update(std::move(entityMap));
        _compiled = true;
    }

    /// compile() for a cloud that already sits in THIS GPU's memory as 240-byte records (e.g. after the NCCL broadcast
    /// that replicates a scene across the GPUs of a box): one group entity, slot 0.
    Entity compileDevice(const void* deviceRecords240, uint32_t gaussianCount, const Settings& settings = Settings::getDefault(),
                         void* stream = nullptr) {
        if (gaussianCount == 0) return Entity{ 0 };
        _shDegree = settings.sphericalHarmonicsDegree > 3 ? 3 : settings.sphericalHarmonicsDegree;
        check(tpdcu_upload_gaussians_device(_ctx, deviceRecords240, gaussianCount, nullptr, 1, stream), "tpdcu_upload_gaussians_device");
        const Entity cloud{ 0 };
        _transformHost->update(std::map<Entity, uint32_t>{ { cloud, 0u } });
        _compiled = true;
        return cloud;
    }

    [[nodiscard]] const std::unique_ptr<TransformHost>& getTransformHost() const noexcept { return _transformHost; }

    /// GaussianEngine::updateCameraBuffer (GaussianEngine.cpp:764-775): view | proj*view | (P00, P11).
    [[nodiscard]] static std::array<float, TPDCU_CAMERA_FLOATS> packCameraBuffer(const Camera& camera) noexcept {
        mat4 projection;
        std::memcpy(projection.data_ptr(), camera.getProjectionData(), sizeof(mat4));
        const float fx = projection.m[0][0];
        const float fy = projection.m[1][1];
        projection = math::mul(projection, camera.getViewMatrix());
        std::array<float, TPDCU_CAMERA_FLOATS> ubo{};
        std::memcpy(ubo.data(), camera.getViewMatrixData(), sizeof(mat4));
        std::memcpy(ubo.data() + 16, projection.data_ptr(), sizeof(mat4));
        ubo[32] = fx;
        ubo[33] = fy;
        return ubo;
    }

    /// GaussianEngine::rasterFrame (GaussianEngine.cpp:621-712). Asynchronous on `stream`; no host sync.
    void rasterFrame(const Camera& camera, void* stream = nullptr) {
        if (!_compiled) return;  // the reference records nothing when _pc.count == 0 (:657)
        const auto ubo = packCameraBuffer(camera);
        check(tpdcu_raster(_ctx, ubo.data(), _shDegree, stream), "tpdcu_raster");
    }

    /// GaussianEngine::draw (GaussianEngine.cpp:714-762) for hosts without a swapchain: wait for the frame and
    /// copy the R8G8B8A8_UNORM target to `hostRgba8` (rows `pitchBytes` apart).
    void draw(void* hostRgba8, std::size_t pitchBytes) const {
        check(tpdcu_read_frame(_ctx, hostRgba8, pitchBytes), "tpdcu_read_frame");
    }
    /// draw() without the wait: the copy is enqueued on `stream` (the one passed to rasterFrame) and the caller waits on the
    /// stream later, so the next rasterFrame can be issued while this frame travels to the host — what the swap-chain
    /// fences give the reference's render loop (SurfaceRenderer.cpp:254-319).
    void drawAsync(void* hostRgba8, std::size_t pitchBytes, void* stream) const {
        check(tpdcu_read_frame_async(_ctx, hostRgba8, pitchBytes, stream), "tpdcu_read_frame_async");
    }
    /// With a Vulkan swap target bound (bindSwapTarget) draw() only has to wait for CUDA; the Vulkan side then
    /// records copyBufferToImage into the swap image in place of recordTargetCopy.
    void draw() const { check(tpdcu_finish(_ctx, nullptr), "tpdcu_finish"); }

    void bindSwapTarget(int opaqueFd, std::size_t bytes) { check(tpdcu_bind_output_fd(_ctx, opaqueFd, bytes), "tpdcu_bind_output_fd"); }

    /// onFramebufferResize (GaussianEngine.cpp:206-220)
    void resize(uint32_t width, uint32_t height) {
        check(tpdcu_resize(_ctx, width, height), "tpdcu_resize");
        _width = width;
        _height = height;
    }
    [[nodiscard]] std::pair<uint32_t, uint32_t> getFramebufferSize() const noexcept { return { _width, _height }; }

    /// Engine::waitIdle (rendering/include/torpedo/rendering/Engine.h:19)
    void waitIdle() const noexcept { tpdcu_finish(_ctx, nullptr); }

    /// The C-ABI handle, for introspection (tpdcu_read_keys, tpdcu_stage_times_ms, ...).
    [[nodiscard]] tpdcu_ctx* handle() const noexcept { return _ctx; }

private:
    static void check(int status, const char* what) {
        if (status != TPDCU_OK) throw std::runtime_error(std::string("GaussianEngine - ") + what + ": " + tpdcu_last_error());
    }

    tpdcu_ctx* _ctx{ nullptr };
    std::unique_ptr<TransformHost> _transformHost{};
    uint32_t _shDegree{ 3 };
    uint32_t _width{ 0 }, _height{ 0 };
    bool _compiled{ false };
};

}  // namespace tpd
