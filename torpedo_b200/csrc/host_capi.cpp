// host_capi.cpp — flat C wrapper over the C++ drop-in (include/torpedo_b200/*.hpp) so that the Python
// tests and bench.py drive the SAME host layer a C++ application would (lib/libtpdhost.so, links libtpdcu.so).
#include "../../include/torpedo_b200/GaussianEngine.hpp"

#include <string>
#include <vector>

namespace {
thread_local std::string g_error;
template <typename F>
int guarded(F&& f) noexcept {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    } catch (...) {
        g_error = "unknown C++ exception";
        return -1;
    }
}
}  // namespace

extern "C" {

const char* tpdh_last_error() { return g_error.c_str(); }

// ---- camera ---------------------------------------------------------------------------------------
tpd::PerspectiveCamera* tpdh_camera_create(uint32_t w, uint32_t h) { return new tpd::PerspectiveCamera(w, h); }
void tpdh_camera_destroy(tpd::PerspectiveCamera* c) { delete c; }
void tpdh_camera_look_at(tpd::PerspectiveCamera* c, const float eye[3], const float center[3], const float up[3]) {
    c->lookAt({ eye[0], eye[1], eye[2] }, { center[0], center[1], center[2] }, { up[0], up[1], up[2] });
}
void tpdh_camera_look_at_rt(tpd::PerspectiveCamera* c, const float R[9], const float t[3]) {
    tpd::mat3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[i][j] = R[i * 3 + j];
    c->lookAt(r, { t[0], t[1], t[2] });
}
void tpdh_camera_set_near(tpd::PerspectiveCamera* c, float v) { c->setNear(v); }
void tpdh_camera_set_far(tpd::PerspectiveCamera* c, float v) { c->setFar(v); }
void tpdh_camera_set_vertical_fov(tpd::PerspectiveCamera* c, float degrees) { c->setVerticalFov(degrees); }
void tpdh_camera_on_image_size_change(tpd::PerspectiveCamera* c, uint32_t w, uint32_t h) { c->onImageSizeChange(w, h); }
void tpdh_camera_pack(const tpd::PerspectiveCamera* c, float out34[TPDCU_CAMERA_FLOATS]) {
    const auto ubo = tpd::GaussianEngine::packCameraBuffer(*c);
    std::memcpy(out34, ubo.data(), sizeof(float) * TPDCU_CAMERA_FLOATS);
}
void tpdh_to_cartesian(float theta, float phi, float radius, float out3[3]) {
    const auto v = tpd::math::to_cartesian(theta, phi, radius);
    out3[0] = v.x; out3[1] = v.y; out3[2] = v.z;
}
void tpdh_rgb2sh(float r, float g, float b, float out48[48]) {
    const auto sh = tpd::utils::rgb2sh(r, g, b);
    std::memcpy(out48, sh.data(), sizeof(float) * 48);
}
uint32_t tpdh_sizeof_gaussian_point() { return sizeof(tpd::GaussianPoint); }
int tpdh_random_points(uint32_t count, float radius, float minScale, float maxScale, float minOpacity, float maxOpacity, uint64_t seed,
                       void* out240) {
    return guarded([&] {
        const auto pts = tpd::GaussianPoint::random(count, radius, { 0.f, 0.f, 0.f }, minScale, maxScale, minOpacity, maxOpacity, seed);
        std::memcpy(out240, pts.data(), pts.size() * sizeof(tpd::GaussianPoint));
    });
}

// GaussianPoint::fromModel: returns the point count (or -1) and keeps the cloud until tpdh_model_take copies it out
static thread_local std::vector<tpd::GaussianPoint> g_model;
int64_t tpdh_model_load(const char* path) {
    int64_t n = -1;
    guarded([&] { g_model = tpd::GaussianPoint::fromModel(path); n = static_cast<int64_t>(g_model.size()); });
    return n;
}
void tpdh_model_take(void* out240) {
    std::memcpy(out240, g_model.data(), g_model.size() * sizeof(tpd::GaussianPoint));
    g_model.clear();
    g_model.shrink_to_fit();
}

// ---- scene ----------------------------------------------------------------------------------------
tpd::Scene* tpdh_scene_create() { return new tpd::Scene(); }
void tpdh_scene_destroy(tpd::Scene* s) { delete s; }
// `points` is borrowed until tpdh_engine_compile returns, like ent::group in the reference
uint32_t tpdh_scene_add_group(tpd::Scene* s, const void* points240, uint32_t count) {
    return static_cast<uint32_t>(s->add(tpd::EntityGroup<tpd::GaussianPoint>{ static_cast<const tpd::GaussianPoint*>(points240), count }));
}
uint32_t tpdh_scene_add_point(tpd::Scene* s, const void* point240) {
    tpd::GaussianPoint p;
    std::memcpy(&p, point240, sizeof(p));
    return static_cast<uint32_t>(s->add(std::move(p)));
}
uint32_t tpdh_scene_count_all(const tpd::Scene* s) { return s->countAll<tpd::GaussianPoint>(); }

// ---- engine ---------------------------------------------------------------------------------------
tpd::GaussianEngine* tpdh_engine_create(uint32_t w, uint32_t h, int device) {
    tpd::GaussianEngine* e = nullptr;
    guarded([&] { e = new tpd::GaussianEngine(w, h, device); });
    return e;
}
void tpdh_engine_destroy(tpd::GaussianEngine* e) { delete e; }
int tpdh_engine_compile(tpd::GaussianEngine* e, const tpd::Scene* s, uint32_t shDegree) {
    return guarded([&] { e->compile(*s, tpd::GaussianEngine::Settings{ shDegree }); });
}
int tpdh_engine_compile_device(tpd::GaussianEngine* e, const void* dRecs240, uint32_t count, uint32_t shDegree, void* stream) {
    return guarded([&] { e->compileDevice(dRecs240, count, tpd::GaussianEngine::Settings{ shDegree }, stream); });
}
int tpdh_engine_transform(tpd::GaussianEngine* e, uint32_t entity, const float m[16]) {
    return guarded([&] {
        tpd::mat4 t;
        std::memcpy(t.data_ptr(), m, sizeof(tpd::mat4));
        e->getTransformHost()->transform(static_cast<tpd::Entity>(entity), t);
    });
}
int tpdh_engine_raster_frame(tpd::GaussianEngine* e, const tpd::PerspectiveCamera* c, void* stream) {
    return guarded([&] { e->rasterFrame(*c, stream); });
}
int tpdh_engine_draw(tpd::GaussianEngine* e, void* hostRgba8, size_t pitch) {
    return guarded([&] { e->draw(hostRgba8, pitch); });
}
int tpdh_engine_draw_async(tpd::GaussianEngine* e, void* hostRgba8, size_t pitch, void* stream) {
    return guarded([&] { e->drawAsync(hostRgba8, pitch, stream); });
}
int tpdh_engine_resize(tpd::GaussianEngine* e, uint32_t w, uint32_t h) {
    return guarded([&] { e->resize(w, h); });
}
int tpdh_engine_wait_idle(tpd::GaussianEngine* e) {
    return guarded([&] { e->waitIdle(); });
}
tpdcu_ctx* tpdh_engine_handle(tpd::GaussianEngine* e) { return e->handle(); }

}  // extern "C"
