// scene.hpp — host-side scene model of the headless renderer.
// Mirrors src/scene.rs (Scene, Sphere, Plane, Material, Bounds3), src/mesh.rs
// (Mesh, PackedMeshes) and src/bvh.rs (flattened BVH) of the reference.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "rt_types.h"

namespace rt {

// glam::Vec3 (scalar, 12 bytes) with the operation order glam 0.30.9 uses.
struct Vec3 {
    float x = 0, y = 0, z = 0;
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 v3(float x, float y, float z) { return Vec3{x, y, z}; }
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float fmin_rs(float a, float b) { return b < a ? b : a; }  // f32::min without NaNs
inline float fmax_rs(float a, float b) { return b > a ? b : a; }
inline Vec3 vmin(Vec3 a, Vec3 b) { return {fmin_rs(a.x, b.x), fmin_rs(a.y, b.y), fmin_rs(a.z, b.z)}; }
inline Vec3 vmax(Vec3 a, Vec3 b) { return {fmax_rs(a.x, b.x), fmax_rs(a.y, b.y), fmax_rs(a.z, b.z)}; }
inline float dot(Vec3 a, Vec3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
inline Vec3 cross(Vec3 a, Vec3 b) {
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}

// src/scene.rs:60-141
struct Bounds3 {
    Vec3 min, max;
    static Bounds3 identity();
    Bounds3 unite(const Bounds3& o) const { return {vmin(min, o.min), vmax(max, o.max)}; }
    Bounds3 extend(Vec3 p) const { return {vmin(min, p), vmax(max, p)}; }
    Vec3 center() const { return min * 0.5f + max * 0.5f; }  // scene.rs:102-104
    int max_axis() const;                                    // scene.rs:113-122
    float surface_area() const;                              // scene.rs:125-128
};

struct Material {
    Vec3 color;
    float roughness = 0, metallic = 0;
    Vec3 emission;
};
struct Sphere {
    Vec3 pos;
    float radius = 0;
    uint32_t material_id = 0;
};
// src/scene.rs:182-188 (the un-baked plane: bounds need forward/right).
struct Plane {
    Vec3 pos, forward, right;
    uint32_t material_id = 0;
};
struct MeshTriangle {
    uint32_t v[3];
    uint32_t n[3];
    uint32_t material_id;
};
// src/mesh.rs:16-22
struct Mesh {
    std::vector<Vec3> vertices, normals;
    std::vector<MeshTriangle> triangles;
};
struct Camera {
    Vec3 pos;
    float yaw = 0, pitch = 0, fov_y = 0;  // radians
};

struct SceneError : std::runtime_error {
    int status;
    SceneError(int status_, const std::string& msg) : std::runtime_error(msg), status(status_) {}
};

// Mesh::load (src/mesh.rs:29-81) over the wavefront_obj 11.0.0 parse.
Mesh load_obj_mesh(const std::string& text, uint32_t material_id);

struct HostScene {
    // Source description (src/scene.rs:224-231).
    std::vector<Material> materials;
    std::vector<Sphere> spheres;
    std::vector<Plane> planes;
    std::vector<Mesh> meshes;
    Camera camera;

    // Baked, reference-layout buffers (what State::new uploads, src/state.rs:394-458).
    bool finalized = false;
    std::vector<RtMaterial> u_materials;
    std::vector<RtSphere> u_spheres;
    std::vector<RtPlane> u_planes;
    std::vector<RtVec3> u_vertices, u_normals;
    std::vector<RtTriangle> u_triangles;
    std::vector<RtPrimitiveInfo> u_primitives;
    std::vector<RtBvhNode> u_nodes;
    uint32_t bvh_depth = 0;
    std::vector<float> plane_bounds;  // Plane::bounds (scene.rs:203-207), 6 floats per plane: input of a device BVH build

    // pack_meshes + to_uniform + build_bvh; with_bvh = false leaves the BVH to rt_bvh_build (device)
    void finalize(bool with_bvh = true);
};

// Scene::load_toml (src/scene.rs:235-252, 324-441). Throws SceneError with the
// reference's message text.
void load_scene_toml(const std::string& path, HostScene& out);

// build_bvh (src/bvh.rs:13-28): fills u_primitives / u_nodes from the packed arrays.
void build_bvh(HostScene& scene);

// Plane::to_uniform (src/scene.rs:190-201).
RtPlane plane_to_uniform(const Plane& p);

// Camera (src/camera.rs).
void camera_uniform(const Camera& c, RtCameraUniform* out);
std::string camera_serialize(const Camera& c);
Camera camera_deserialize(const std::string& encoded);  // throws SceneError

// Environments.
void hdr_decode(const uint8_t* bytes, size_t len, uint32_t* w, uint32_t* h, std::vector<float>& rgb);
std::vector<uint8_t> hdr_encode_flat(uint32_t w, uint32_t h, const float* rgb);
void hdri_synthetic(uint32_t w, uint32_t h, float* rgb);
uint32_t alias_build(uint32_t w, uint32_t h, const float* rgb, RtAliasEntry* out);
uint32_t alias_build_serial(uint32_t w, uint32_t h, const float* rgb, RtAliasEntry* out);  // literal form

// Config D generator.
void build_synthetic_field(const std::string& assets_dir, uint32_t grid, uint32_t seed,
                           HostScene& out, bool with_bvh = true);

std::string read_text_file(const std::string& path);  // throws std::runtime_error(strerror)

}  // namespace rt
