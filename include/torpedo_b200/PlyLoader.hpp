// PlyLoader.hpp — tpd::GaussianPoint::fromModel: the 3DGS point-cloud (PLY) ingest of the reference
// (torpedo/volumetric/src/GaussianGeometry.cpp:59-127). The reference parses the file with the vendored miniply; this is a
// self-contained reader for what 3DGS trainers write — one `vertex` element of scalar properties, binary_little_endian or
// ascii — followed by the reference's field transforms (:110-117):
//     opacity    = 1 / (1 + exp(-raw))
//     quaternion = normalize(rot_1, rot_2, rot_3, rot_0)          (w LAST; compensated dot, math/vec4.h:248-265)
//     scale      = (exp(scale_0), exp(scale_1), exp(scale_2), 1)
//     sh         = f_dc_0..2 followed by the (count of f_rest_*) properties that FOLLOW f_dc_2 in the file (:88-93)
#pragma once

#include "GaussianGeometry.hpp"

#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace tpd {
namespace detail {

struct PlyProperty {
    std::string name;
    std::size_t size;   // bytes in the binary encoding
    char kind;          // 'f' float32, 'd' float64, 'i' signed int, 'u' unsigned int
};

inline std::size_t plyTypeSize(const std::string& t, char& kind) {
    if (t == "float" || t == "float32") { kind = 'f'; return 4; }
    if (t == "double" || t == "float64") { kind = 'd'; return 8; }
    if (t == "char" || t == "int8") { kind = 'i'; return 1; }
    if (t == "uchar" || t == "uint8") { kind = 'u'; return 1; }
    if (t == "short" || t == "int16") { kind = 'i'; return 2; }
    if (t == "ushort" || t == "uint16") { kind = 'u'; return 2; }
    if (t == "int" || t == "int32") { kind = 'i'; return 4; }
    if (t == "uint" || t == "uint32") { kind = 'u'; return 4; }
    throw std::runtime_error("PLY: unsupported property type " + t);
}

inline float plyToFloat(const unsigned char* p, const PlyProperty& prop) {
    switch (prop.kind) {
        case 'f': { float v; std::memcpy(&v, p, 4); return v; }
        case 'd': { double v; std::memcpy(&v, p, 8); return static_cast<float>(v); }
        case 'i':
            if (prop.size == 1) { int8_t v; std::memcpy(&v, p, 1); return static_cast<float>(v); }
            if (prop.size == 2) { int16_t v; std::memcpy(&v, p, 2); return static_cast<float>(v); }
            { int32_t v; std::memcpy(&v, p, 4); return static_cast<float>(v); }
        default:
            if (prop.size == 1) { uint8_t v; std::memcpy(&v, p, 1); return static_cast<float>(v); }
            if (prop.size == 2) { uint16_t v; std::memcpy(&v, p, 2); return static_cast<float>(v); }
            { uint32_t v; std::memcpy(&v, p, 4); return static_cast<float>(v); }
    }
}

}  // namespace detail

inline std::vector<GaussianPoint> GaussianPoint::fromModel(const std::filesystem::path& plyFile) {
    using namespace detail;
    std::ifstream in(plyFile, std::ios::binary);
    if (!in.is_open()) throw std::runtime_error("Failed to open file: " + plyFile.string());

    // ---- header -------------------------------------------------------------------------------------------------------
    std::string line;
    if (!std::getline(in, line) || line.rfind("ply", 0) != 0) throw std::runtime_error("Not a PLY file: " + plyFile.string());
    bool binary = false, ascii = false, inVertex = false, sawVertex = false;
    std::size_t count = 0;
    std::vector<PlyProperty> props;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ls(line);
        std::string word;
        ls >> word;
        if (word == "format") {
            std::string fmt;
            ls >> fmt;
            binary = fmt == "binary_little_endian";
            ascii = fmt == "ascii";
            if (!binary && !ascii) throw std::runtime_error("PLY: unsupported format " + fmt);
        } else if (word == "element") {
            std::string name;
            std::size_t n = 0;
            ls >> name >> n;
            if (sawVertex && inVertex && name != "vertex") { inVertex = false; }
            if (name == "vertex") {
                if (!props.empty()) throw std::runtime_error("PLY: more than one vertex element");
                inVertex = sawVertex = true;
                count = n;
            } else if (!sawVertex && n > 0) {
                throw std::runtime_error("PLY: elements before `vertex` are not supported");
            } else {
                inVertex = false;
            }
        } else if (word == "property" && inVertex) {
            std::string type, name;
            ls >> type;
            if (type == "list") throw std::runtime_error("PLY: list properties in the vertex element are not supported");
            ls >> name;
            PlyProperty p;
            p.name = name;
            p.size = plyTypeSize(type, p.kind);
            props.push_back(p);
        } else if (word == "end_header") {
            break;
        }
    }
    if (!sawVertex) throw std::runtime_error("Could NOT find vertices: " + plyFile.string());

    auto find = [&](const char* name) -> std::size_t {
        for (std::size_t i = 0; i < props.size(); ++i)
            if (props[i].name == name) return i;
        throw std::runtime_error(std::string("PLY: missing property ") + name);
    };
    const std::size_t iPos[3] = { find("x"), find("y"), find("z") };
    const std::size_t iRot[4] = { find("rot_0"), find("rot_1"), find("rot_2"), find("rot_3") };
    const std::size_t iScale[3] = { find("scale_0"), find("scale_1"), find("scale_2") };
    const std::size_t iOpacity = find("opacity");
    const std::size_t iDc[3] = { find("f_dc_0"), find("f_dc_1"), find("f_dc_2") };
    std::size_t featureCount = 3;  // countFeatures (:49-55)
    for (const auto& p : props)
        if (p.name.rfind("f_rest_", 0) == 0) ++featureCount;
    if (featureCount > MAX_SH_FLOATS) throw std::runtime_error("PLY: more than 45 f_rest_* properties");
    if (iDc[2] + (featureCount - 3) >= props.size() + (featureCount == 3 ? 1 : 0) && featureCount > 3)
        throw std::runtime_error("PLY: f_rest_* properties must follow f_dc_2");

    std::vector<std::size_t> offset(props.size());
    std::size_t stride = 0;
    for (std::size_t i = 0; i < props.size(); ++i) { offset[i] = stride; stride += props[i].size; }

    // ---- body ---------------------------------------------------------------------------------------------------------
    std::vector<GaussianPoint> points(count);
    std::vector<unsigned char> row(stride);
    std::vector<float> values(props.size());
    for (std::size_t n = 0; n < count; ++n) {
        if (binary) {
            in.read(reinterpret_cast<char*>(row.data()), static_cast<std::streamsize>(stride));
            if (!in) throw std::runtime_error("PLY: unexpected end of file");
            for (std::size_t i = 0; i < props.size(); ++i) values[i] = plyToFloat(row.data() + offset[i], props[i]);
        } else {
            for (std::size_t i = 0; i < props.size(); ++i) {
                double v;
                if (!(in >> v)) throw std::runtime_error("PLY: unexpected end of file");
                values[i] = static_cast<float>(v);
            }
        }
        GaussianPoint& p = points[n];
        p.position = { values[iPos[0]], values[iPos[1]], values[iPos[2]] };
        p.opacity = 1.f / (1.f + std::exp(-values[iOpacity]));
        const vec4 q{ values[iRot[1]], values[iRot[2]], values[iRot[3]], values[iRot[0]] };
        const float inv = 1.0f / std::sqrt(math::dot(q, q));
        p.quaternion = { q.x * inv, q.y * inv, q.z * inv, q.w * inv };
        p.scale = { std::exp(values[iScale[0]]), std::exp(values[iScale[1]]), std::exp(values[iScale[2]]), 1.0f };
        p.sh.fill(0.0f);
        for (std::size_t k = 0; k < 3; ++k) p.sh[k] = values[iDc[k]];
        for (std::size_t k = 3; k < featureCount; ++k) p.sh[k] = values[iDc[2] + (k - 2)];
    }
    return points;
}

}  // namespace tpd
