// Scene.hpp — tpd::Scene with the slice of the reference interface GaussianEngine::compile consumes
// (torpedo/rendering/include/torpedo/rendering/Scene.h:17-163). The reference stores components in an EnTT
// registry; compile() only ever asks for "all groups first, then all singles" (Scene.h:130-146), the group
// sizes and an entity -> transform-slot map, so a pair of vectors is enough. Entities enumerate in INSERTION
// order here (EnTT v3.15's view order among several groups is an off-tree implementation detail; the
// reference demos use one group plus at most one single, for which both orders agree).
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <span>
#include <type_traits>
#include <typeindex>
#include <utility>
#include <vector>

namespace tpd {

enum class Entity : uint32_t {};

template <typename T>
using EntityGroup = std::span<const T>;

namespace ent {
template <typename T>
[[nodiscard]] EntityGroup<T> group(const std::vector<T>& elements) {
    return EntityGroup<T>{ elements.data(), elements.size() };
}
}  // namespace ent

class Scene final {
public:
    /// A group is BORROWED (a span), exactly like the reference: it must stay alive until compile() returns.
    template <typename T>
    Entity add(EntityGroup<T> elements) {
        auto& s = store<T>();
        const Entity e{ _next++ };
        s.groups.push_back({ e, elements });
        return e;
    }
    /// A single element is copied into the scene.
    template <typename T>
    Entity add(T&& element) {
        using U = std::remove_cvref_t<T>;
        auto& s = store<U>();
        const Entity e{ _next++ };
        s.singles.push_back({ e, std::forward<T>(element) });
        return e;
    }

    template <typename T>
    [[nodiscard]] uint32_t count() const noexcept { return static_cast<uint32_t>(store<T>().singles.size()); }
    template <typename T>
    [[nodiscard]] uint32_t countGroup() const noexcept { return static_cast<uint32_t>(store<T>().groups.size()); }
    template <typename T>
    [[nodiscard]] uint32_t countAll() const noexcept {
        uint32_t n = count<T>();
        for (const auto& g : store<T>().groups) n += static_cast<uint32_t>(g.second.size());
        return n;
    }
    template <typename T>
    [[nodiscard]] std::vector<uint32_t> groupSizes() const {
        std::vector<uint32_t> sizes;
        for (const auto& g : store<T>().groups) sizes.push_back(static_cast<uint32_t>(g.second.size()));
        return sizes;
    }
    /// Groups first, then singles (Scene.h:130-146).
    template <typename T>
    [[nodiscard]] std::vector<std::byte> dataAll() const {
        std::vector<std::byte> bytes(static_cast<std::size_t>(countAll<T>()) * sizeof(T));
        std::size_t off = 0;
        for (const auto& g : store<T>().groups) {
            std::memcpy(bytes.data() + off, g.second.data(), g.second.size_bytes());
            off += g.second.size_bytes();
        }
        for (const auto& s : store<T>().singles) {
            std::memcpy(bytes.data() + off, &s.second, sizeof(T));
            off += sizeof(T);
        }
        return bytes;
    }
    /// entity -> transform slot, groups first then singles (Scene.h:155-163).
    template <typename T>
    [[nodiscard]] std::map<Entity, uint32_t> buildEntityMap() const {
        std::map<Entity, uint32_t> map;
        uint32_t slot = 0;
        for (const auto& g : store<T>().groups) map.emplace(g.first, slot++);
        for (const auto& s : store<T>().singles) map.emplace(s.first, slot++);
        return map;
    }

private:
    template <typename T>
    struct Store {
        std::vector<std::pair<Entity, EntityGroup<T>>> groups;
        std::vector<std::pair<Entity, T>> singles;
    };
    template <typename T>
    Store<T>& store() const {
        auto& slot = _stores[std::type_index(typeid(T))];
        if (!slot) slot = std::make_shared<Store<T>>();
        return *static_cast<Store<T>*>(slot.get());
    }
    mutable std::map<std::type_index, std::shared_ptr<void>> _stores;
    uint32_t _next = 0;

public:
    Scene() = default;
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;
};

}  // namespace tpd
