// scene.cpp — scene TOML loading and baking into the reference's buffer layouts.
// Follows src/scene.rs:235-252 (load_toml), 264-322 (descriptors), 324-441
// (build_scene), 190-201 (Plane::to_uniform) and src/mesh.rs:92-113 (pack_meshes).
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "scene.hpp"
#include "toml_lite.hpp"

namespace rt {

Bounds3 Bounds3::identity() {
    const float m = 3.40282347e+38f;  // Vec3::MAX / Vec3::MIN (scene.rs:67-72)
    return {v3(m, m, m), v3(-m, -m, -m)};
}

int Bounds3::max_axis() const {
    Vec3 d = max - min;
    if (d.z > d.x && d.z > d.y) return 2;
    if (d.y > d.x) return 1;
    return 0;
}

float Bounds3::surface_area() const {
    Vec3 d = max - min;
    return 2.f * (d.x * d.y + d.x * d.z + d.y * d.z);
}

namespace {

std::string dirname_of(const std::string& path) {
    size_t slash = path.find_last_of('/');
    if (slash == std::string::npos) return ".";
    if (slash == 0) return "/";
    return path.substr(0, slash);
}

// serde-style field access with messages that name the missing / mistyped field.
const toml::Value* need_ptr(const toml::Value& table, const char* key, const char* ctx) {
    const toml::Value* v = table.find(key);
    if (!v) throw std::runtime_error("missing field `" + std::string(key) + "`" + ctx);
    return v;
}
float need_f32(const toml::Value& table, const char* key, const char* ctx) {
    const toml::Value& v = *need_ptr(table, key, ctx);
    if (!v.is_number())
        throw std::runtime_error("invalid type: " + std::string(v.kind_name()) +
                                 ", expected f32 for `" + key + "`" + ctx);
    return (float)v.as_double();  // toml f64/i64 -> f32 cast (scene.rs:243)
}
Vec3 need_vec3(const toml::Value& table, const char* key, const char* ctx) {
    const toml::Value& v = *need_ptr(table, key, ctx);
    if (v.kind != toml::Value::Array || v.array.size() != 3 || !v.array[0]->is_number() ||
        !v.array[1]->is_number() || !v.array[2]->is_number())
        throw std::runtime_error("invalid value for `" + std::string(key) +
                                 "`: expected a sequence of 3 f32 values" + ctx);
    return v3((float)v.array[0]->as_double(), (float)v.array[1]->as_double(),
              (float)v.array[2]->as_double());
}
std::string need_string(const toml::Value& table, const char* key, const char* ctx) {
    const toml::Value& v = *need_ptr(table, key, ctx);
    if (v.kind != toml::Value::String)
        throw std::runtime_error("invalid type: " + std::string(v.kind_name()) +
                                 ", expected a string for `" + key + "`" + ctx);
    return v.str;
}

// cgmath 0.18 Deg<f32> -> Rad<f32>: deg * (PI / 180) in f32 (scene.rs:305-314).
float deg_to_rad(float deg) {
    const float k = (float)(3.14159265358979323846 / 180.0);
    return deg * k;
}

}  // namespace

RtPlane plane_to_uniform(const Plane& p) {
    // scene.rs:191: normal = forward.cross(right).normalize(); glam normalize = v * (1/len).
    Vec3 c = cross(p.forward, p.right);
    float len = std::sqrt(dot(c, c));
    Vec3 normal = c * (1.0f / len);

    // scene.rs:196: Mat3::from_cols(right, normal, forward).inverse() — glam 0.30.9:
    //   tmp0 = y.cross(z); tmp1 = z.cross(x); tmp2 = x.cross(y); det = z.dot(tmp2);
    //   inverse = from_cols(tmp0/det, tmp1/det, tmp2/det).transpose()   (times 1/det)
    Vec3 x = p.right, y = normal, z = p.forward;
    Vec3 tmp0 = cross(y, z), tmp1 = cross(z, x), tmp2 = cross(x, y);
    float det = dot(z, tmp2);
    float inv_det = 1.0f / det;
    Vec3 r0 = tmp0 * inv_det, r1 = tmp1 * inv_det, r2 = tmp2 * inv_det;
    // Rows of the inverse are r0, r1, r2; columns are their transposes.
    RtPlane u;
    std::memset(&u, 0, sizeof u);
    u.pos[0] = p.pos.x; u.pos[1] = p.pos.y; u.pos[2] = p.pos.z;
    u.normal[0] = normal.x; u.normal[1] = normal.y; u.normal[2] = normal.z;
    u.base_change[0][0] = r0.x; u.base_change[0][1] = r1.x; u.base_change[0][2] = r2.x;
    u.base_change[1][0] = r0.y; u.base_change[1][1] = r1.y; u.base_change[1][2] = r2.y;
    u.base_change[2][0] = r0.z; u.base_change[2][1] = r1.z; u.base_change[2][2] = r2.z;
    u.material_id = p.material_id;
    return u;
}

void HostScene::finalize(bool with_bvh) {
    u_materials.clear(); u_spheres.clear(); u_planes.clear();
    u_vertices.clear(); u_normals.clear(); u_triangles.clear();

    for (const Material& m : materials) {
        RtMaterial u;
        std::memset(&u, 0, sizeof u);
        u.color[0] = m.color.x; u.color[1] = m.color.y; u.color[2] = m.color.z;
        u.roughness = m.roughness;
        u.metallic = m.metallic;
        u.emission[0] = m.emission.x; u.emission[1] = m.emission.y; u.emission[2] = m.emission.z;
        u_materials.push_back(u);
    }
    for (const Sphere& s : spheres) {
        RtSphere u;
        std::memset(&u, 0, sizeof u);
        u.pos[0] = s.pos.x; u.pos[1] = s.pos.y; u.pos[2] = s.pos.z;
        u.radius = s.radius;
        u.material_id = s.material_id;
        u_spheres.push_back(u);
    }
    for (const Plane& p : planes) u_planes.push_back(plane_to_uniform(p));

    // PackedMeshes::pack_meshes (mesh.rs:92-113): concatenate with running offsets.
    for (const Mesh& mesh : meshes) {
        uint32_t vo = (uint32_t)u_vertices.size(), no = (uint32_t)u_normals.size();
        for (const MeshTriangle& t : mesh.triangles)
            u_triangles.push_back(RtTriangle{t.v[0] + vo, t.v[1] + vo, t.v[2] + vo, t.n[0] + no,
                                             t.n[1] + no, t.n[2] + no, t.material_id});
        for (Vec3 v : mesh.vertices) u_vertices.push_back(RtVec3{v.x, v.y, v.z, 0.f});
        for (Vec3 n : mesh.normals) u_normals.push_back(RtVec3{n.x, n.y, n.z, 0.f});
    }

    plane_bounds.clear();
    for (const Plane& p : planes) {
        Bounds3 b = Bounds3::identity().extend(p.pos).extend(p.pos + p.forward + p.right);
        const float six[6] = {b.min.x, b.min.y, b.min.z, b.max.x, b.max.y, b.max.z};
        plane_bounds.insert(plane_bounds.end(), six, six + 6);
    }
    if (!with_bvh) {
        u_primitives.clear();
        u_nodes.clear();
        bvh_depth = 0;
        finalized = true;
        return;
    }

    const bool timing = std::getenv("RT_HOST_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    build_bvh(*this);
    if (timing)
        std::fprintf(stderr, "[host] build_bvh: %.3f s for %zu primitives, %zu nodes\n",
                     std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(),
                     u_primitives.size(), u_nodes.size());
    finalized = true;
}

void load_scene_toml(const std::string& path, HostScene& out) {
    std::string content;
    try {
        content = read_text_file(path);
    } catch (const std::exception& e) {
        throw SceneError(RT_ERR_IO, "Couldn't open scene " + path + ":\n  " + e.what());
    }

    toml::ValuePtr doc;
    struct MaterialDescr {
        std::string name;
        Material m;
    };
    std::vector<MaterialDescr> material_descrs;
    struct ObjectDescr {
        std::string type, material, mesh_path;
        Vec3 pos, forward, right;
        float radius = 0;
    };
    std::vector<ObjectDescr> objects;
    Camera camera;

    try {
        doc = toml::parse(content);
        const toml::Value& mats = *need_ptr(*doc, "material", "");
        if (mats.kind != toml::Value::Array)
            throw std::runtime_error("invalid type: expected a sequence for `material`");
        for (auto& mv : mats.array) {
            if (mv->kind != toml::Value::Table)
                throw std::runtime_error("invalid type: expected a table in `material`");
            MaterialDescr d;
            d.name = need_string(*mv, "name", " in material");
            d.m.color = need_vec3(*mv, "color", " in material");
            d.m.roughness = need_f32(*mv, "roughness", " in material");
            d.m.metallic = need_f32(*mv, "metallic", " in material");
            d.m.emission = need_vec3(*mv, "emission", " in material");
            material_descrs.push_back(d);
        }

        const toml::Value& objs = *need_ptr(*doc, "object", "");
        if (objs.kind != toml::Value::Array)
            throw std::runtime_error("invalid type: expected a sequence for `object`");
        for (auto& ov : objs.array) {
            // Externally tagged enum: a table with exactly one variant key (scene.rs:273-294).
            if (ov->kind != toml::Value::Table || ov->table.size() != 1)
                throw std::runtime_error(
                    "wanted exactly 1 element, expected one of `Sphere`, `Plane`, `Mesh`");
            const std::string& tag = ov->table[0].first;
            const toml::Value& body = *ov->table[0].second;
            if (body.kind != toml::Value::Table)
                throw std::runtime_error("invalid type: expected a table for variant `" + tag + "`");
            ObjectDescr d;
            d.type = tag;
            if (tag == "Sphere") {
                d.material = need_string(body, "material", " in Sphere");
                d.pos = need_vec3(body, "pos", " in Sphere");
                d.radius = need_f32(body, "radius", " in Sphere");
            } else if (tag == "Plane") {
                d.material = need_string(body, "material", " in Plane");
                d.pos = need_vec3(body, "pos", " in Plane");
                d.forward = need_vec3(body, "forward", " in Plane");
                d.right = need_vec3(body, "right", " in Plane");
            } else if (tag == "Mesh") {
                d.material = need_string(body, "material", " in Mesh");
                d.mesh_path = need_string(body, "path", " in Mesh");
            } else {
                throw std::runtime_error("unknown variant `" + tag +
                                         "`, expected one of `Sphere`, `Plane`, `Mesh`");
            }
            objects.push_back(d);
        }

        const toml::Value& cam = *need_ptr(*doc, "camera", "");
        if (cam.kind != toml::Value::Table)
            throw std::runtime_error("invalid type: expected a table for `camera`");
        camera.pos = need_vec3(cam, "pos", " in camera");
        camera.yaw = deg_to_rad(need_f32(cam, "yaw", " in camera"));
        camera.pitch = deg_to_rad(need_f32(cam, "pitch", " in camera"));
        camera.fov_y = deg_to_rad(need_f32(cam, "fov_y", " in camera"));
    } catch (const SceneError&) {
        throw;
    } catch (const std::exception& e) {
        throw SceneError(RT_ERR_PARSE, "Couldn't parse scene " + path + ":\n  " + e.what());
    }

    // SceneDescriptor::build_scene (scene.rs:324-441).
    auto material_index = [&](const std::string& name) -> int {
        for (size_t i = 0; i < material_descrs.size(); i++)
            if (material_descrs[i].name == name) return (int)i;  // first match (scene.rs:326-332)
        return -1;
    };
    auto object_error = [&](size_t index, const std::string& type, const std::string& msg) {
        return SceneError(RT_ERR_SCENE, "Error in object " + std::to_string(index) + " (" + type +
                                            "): " + msg + "\n  --> " + path);
    };

    out = HostScene();
    for (auto& d : material_descrs) out.materials.push_back(d.m);
    for (size_t i = 0; i < objects.size(); i++) {
        const ObjectDescr& d = objects[i];
        int mat = material_index(d.material);
        if (mat < 0)
            throw object_error(i, d.type, "Material '" + d.material + "' does not exist.");
        if (d.type == "Sphere") {
            out.spheres.push_back(Sphere{d.pos, d.radius, (uint32_t)mat});
        } else if (d.type == "Plane") {
            out.planes.push_back(Plane{d.pos, d.forward, d.right, (uint32_t)mat});
        } else {
            std::string obj_text;
            try {
                obj_text = read_text_file(dirname_of(path) + "/" + d.mesh_path);
            } catch (const std::exception& e) {
                throw object_error(i, "Mesh", "Cannot open '" + d.mesh_path + "': " + e.what());
            }
            try {
                out.meshes.push_back(load_obj_mesh(obj_text, (uint32_t)mat));
            } catch (const std::exception& e) {
                throw object_error(i, "Mesh", e.what());
            }
        }
    }
    out.camera = camera;
    out.finalize();
}

}  // namespace rt
