// synthetic.cpp — generator for the memory-bound benchmark scene ("config D",
// SURVEY §8d): the house meshes and suzanne instanced on a grid x grid field so that
// nodes + primitives far exceed L2. Not part of the reference; it only produces inputs
// for the same build path (pack_meshes + build_bvh) every other scene takes.
#include <cmath>

#include "scene.hpp"

namespace rt {
namespace {

// The path tracer's PCG hash (shader.wgsl:611-617), reused as the instance RNG.
struct Pcg {
    uint32_t state;
    uint32_t next() {
        state = state * 747796405u + 2891336453u;
        uint32_t r = ((state >> ((state >> 28) + 4)) ^ state) * 277803737u;
        return (r >> 22) ^ r;
    }
    float uniform() { return (float)next() * (1.0f / 4294967296.0f); }
};

}  // namespace

void build_synthetic_field(const std::string& assets_dir, uint32_t grid, uint32_t seed,
                           HostScene& out, bool with_bvh) {
    out = HostScene();
    auto mat = [&](float r, float g, float b, float rough, float metal) {
        out.materials.push_back(Material{v3(r, g, b), rough, metal, v3(0, 0, 0)});
        return (uint32_t)out.materials.size() - 1;
    };
    uint32_t m_ground = mat(0.95f, 0.95f, 0.95f, 1, 0);
    uint32_t m_wall = mat(0.95f, 0.95f, 0.95f, 1, 0);
    uint32_t m_roof = mat(0.67f, 0.55f, 0.52f, 1, 0);
    uint32_t m_red = mat(1.0f, 0.52f, 0.56f, 1, 0);
    uint32_t m_green = mat(0.56f, 1.0f, 0.52f, 1, 0);
    uint32_t m_blue = mat(0.52f, 0.56f, 1.0f, 1, 0);
    uint32_t m_metal = mat(0.8f, 0.8f, 0.8f, 0.5f, 1);

    struct Part {
        const char* file;
        uint32_t material;
        Vec3 offset;
    };
    const Part parts[] = {
        {"house_roof.obj", m_roof, v3(0, 0, 0)},       {"house_walls.obj", m_wall, v3(0, 0, 0)},
        {"house_red_wall.obj", m_red, v3(0, 0, 0)},    {"house_green_wall.obj", m_green, v3(0, 0, 0)},
        {"house_blue_wall.obj", m_blue, v3(0, 0, 0)},  {"suzanne.obj", m_metal, v3(2.4f, 1.0f, -1.0f)},
    };
    std::vector<Mesh> protos;
    for (const Part& p : parts) {
        std::string text;
        try {
            text = read_text_file(assets_dir + "/" + p.file);
        } catch (const std::exception& e) {
            throw SceneError(RT_ERR_IO, "Cannot open '" + assets_dir + "/" + p.file + "': " + e.what());
        }
        protos.push_back(load_obj_mesh(text, p.material));
    }

    const float spacing = 6.0f;
    const float half = 0.5f * spacing * (float)grid;
    Pcg rng{seed};
    rng.next();
    for (uint32_t j = 0; j < grid; j++)
        for (uint32_t i = 0; i < grid; i++) {
            float cx = ((float)i + 0.5f) * spacing - half;
            float cz = ((float)j + 0.5f) * spacing - half;
            float yaw = rng.uniform() * 6.2831853f;
            float lift = rng.uniform() * 0.05f;  // keeps centroids of different instances distinct
            float s = std::sin(yaw), c = std::cos(yaw);
            for (size_t k = 0; k < protos.size(); k++) {
                Mesh inst = protos[k];
                Vec3 off = parts[k].offset;
                for (Vec3& v : inst.vertices) {
                    Vec3 p = v + off;
                    v = v3(c * p.x + s * p.z + cx, p.y + lift, -s * p.x + c * p.z + cz);
                }
                for (Vec3& n : inst.normals) n = v3(c * n.x + s * n.z, n.y, -s * n.x + c * n.z);
                out.meshes.push_back(std::move(inst));
            }
        }
    out.planes.push_back(Plane{v3(-half - 10.f, 0, -half - 10.f), v3(0, 0, 2.f * half + 20.f),
                               v3(2.f * half + 20.f, 0, 0), m_ground});
    out.camera.pos = v3(0.f, 0.06f * half + 6.f, half + 4.f);
    out.camera.yaw = 0.f;
    out.camera.pitch = -0.30f;
    out.camera.fov_y = 1.2217305f;  // 70 degrees
    out.finalize(with_bvh);
}

}  // namespace rt
