// TransformHost.hpp — per-entity model matrices (torpedo/rendering/include/torpedo/rendering/TransformHost.h:10-44,
// rendering/src/TransformHost.cpp:3-11). The reference memcpy's the matrix into a mapped, per-frame-in-flight
// uniform ring at entitySlot*64; here it is forwarded to tpdcu_set_transform and becomes effective with the
// next rasterFrame(). Unknown entities are ignored, as in the reference (:4-6).
#pragma once

#include "../tpdcu.h"
#include "Scene.hpp"
#include "math.hpp"

#include <map>
#include <stdexcept>
#include <string>

namespace tpd {

class TransformHost final {
public:
    explicit TransformHost(tpdcu_ctx* ctx) noexcept : _ctx{ ctx } {}
    TransformHost(const TransformHost&) = delete;
    TransformHost& operator=(const TransformHost&) = delete;

    void update(std::map<Entity, uint32_t>&& entityMap) noexcept { _entityMap = std::move(entityMap); }

    void transform(Entity entity, const mat4& transform) const {
        const auto it = _entityMap.find(entity);
        if (it == _entityMap.end()) [[unlikely]] return;
        if (tpdcu_set_transform(_ctx, it->second, transform.data_ptr()) != TPDCU_OK)
            throw std::runtime_error(std::string("TransformHost::transform: ") + tpdcu_last_error());
    }

private:
    tpdcu_ctx* _ctx;
    std::map<Entity, uint32_t> _entityMap{};
};

}  // namespace tpd
