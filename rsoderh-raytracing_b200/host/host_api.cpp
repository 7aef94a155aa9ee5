// host_api.cpp — extern "C" surface of librt_host.so (include/rt_host.h) plus the
// headless image writers (PNG via zlib, PFM).
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "rt_host.h"
#include "scene.hpp"

struct RtHostScene {
    rt::HostScene scene;
};

namespace {

void set_err(char* err, size_t err_len, const std::string& msg) {
    if (!err || err_len == 0) return;
    size_t n = msg.size() < err_len - 1 ? msg.size() : err_len - 1;
    std::memcpy(err, msg.data(), n);
    err[n] = '\0';
}

template <typename F>
int guarded(char* err, size_t err_len, F&& f) {
    try {
        f();
        return RT_OK;
    } catch (const rt::SceneError& e) {
        set_err(err, err_len, e.what());
        return e.status;
    } catch (const std::bad_alloc&) {
        set_err(err, err_len, "out of host memory");
        return RT_ERR_OUT_OF_MEMORY;
    } catch (const std::exception& e) {
        set_err(err, err_len, e.what());
        return RT_ERR_INVALID_ARGUMENT;
    }
}

rt::Vec3 v3p(const float* p) { return rt::v3(p[0], p[1], p[2]); }

void png_chunk(std::vector<uint8_t>& out, const char type[4], const uint8_t* data, size_t len) {
    auto be32 = [&](uint32_t v) {
        out.push_back((uint8_t)(v >> 24));
        out.push_back((uint8_t)(v >> 16));
        out.push_back((uint8_t)(v >> 8));
        out.push_back((uint8_t)v);
    };
    be32((uint32_t)len);
    size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (len) out.insert(out.end(), data, data + len);
    uint32_t crc = (uint32_t)crc32(0L, out.data() + start, (uInt)(len + 4));
    be32(crc);
}

}  // namespace

extern "C" {

int rt_scene_load_toml(const char* path, RtHostScene** out, char* err, size_t err_len) {
    if (!path || !out) return RT_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    RtHostScene* s = new RtHostScene();
    int st = guarded(err, err_len, [&] { rt::load_scene_toml(path, s->scene); });
    if (st != RT_OK) {
        delete s;
        return st;
    }
    *out = s;
    return RT_OK;
}

int rt_scene_create(RtHostScene** out) {
    if (!out) return RT_ERR_INVALID_ARGUMENT;
    *out = new RtHostScene();
    return RT_OK;
}

int rt_scene_add_material(RtHostScene* s, const float color[3], float roughness, float metallic,
                          const float emission[3]) {
    if (!s || !color || !emission) return RT_ERR_INVALID_ARGUMENT;
    s->scene.materials.push_back(rt::Material{v3p(color), roughness, metallic, v3p(emission)});
    s->scene.finalized = false;
    return RT_OK;
}

int rt_scene_add_sphere(RtHostScene* s, const float pos[3], float radius, uint32_t material_id) {
    if (!s || !pos) return RT_ERR_INVALID_ARGUMENT;
    s->scene.spheres.push_back(rt::Sphere{v3p(pos), radius, material_id});
    s->scene.finalized = false;
    return RT_OK;
}

int rt_scene_add_plane(RtHostScene* s, const float pos[3], const float forward[3],
                       const float right[3], uint32_t material_id) {
    if (!s || !pos || !forward || !right) return RT_ERR_INVALID_ARGUMENT;
    s->scene.planes.push_back(rt::Plane{v3p(pos), v3p(forward), v3p(right), material_id});
    s->scene.finalized = false;
    return RT_OK;
}

int rt_scene_add_mesh_obj(RtHostScene* s, const char* obj_text, uint32_t material_id, char* err,
                          size_t err_len) {
    if (!s || !obj_text) return RT_ERR_INVALID_ARGUMENT;
    return guarded(err, err_len, [&] {
        s->scene.meshes.push_back(rt::load_obj_mesh(obj_text, material_id));
        s->scene.finalized = false;
    });
}

int rt_scene_add_mesh_raw(RtHostScene* s, const float* vertices, uint32_t vertex_count,
                          const float* normals, uint32_t normal_count, const uint32_t* triangles,
                          uint32_t triangle_count, uint32_t material_id) {
    if (!s || (!vertices && vertex_count) || (!normals && normal_count) ||
        (!triangles && triangle_count))
        return RT_ERR_INVALID_ARGUMENT;
    rt::Mesh m;
    for (uint32_t i = 0; i < vertex_count; i++) m.vertices.push_back(v3p(vertices + 3 * i));
    for (uint32_t i = 0; i < normal_count; i++) m.normals.push_back(v3p(normals + 3 * i));
    for (uint32_t i = 0; i < triangle_count; i++) {
        const uint32_t* t = triangles + 6 * i;
        if (t[0] >= vertex_count || t[1] >= vertex_count || t[2] >= vertex_count ||
            t[3] >= normal_count || t[4] >= normal_count || t[5] >= normal_count)
            return RT_ERR_INVALID_ARGUMENT;
        m.triangles.push_back(rt::MeshTriangle{{t[0], t[1], t[2]}, {t[3], t[4], t[5]}, material_id});
    }
    s->scene.meshes.push_back(std::move(m));
    s->scene.finalized = false;
    return RT_OK;
}

int rt_scene_set_camera(RtHostScene* s, const float pos[3], float yaw, float pitch, float fov_y) {
    if (!s || !pos) return RT_ERR_INVALID_ARGUMENT;
    s->scene.camera = rt::Camera{v3p(pos), yaw, pitch, fov_y};
    return RT_OK;
}

int rt_scene_finalize(RtHostScene* s, char* err, size_t err_len) {
    if (!s) return RT_ERR_INVALID_ARGUMENT;
    return guarded(err, err_len, [&] {
        for (auto& sp : s->scene.spheres)
            if (sp.material_id >= s->scene.materials.size())
                throw rt::SceneError(RT_ERR_SCENE, "sphere material index out of range");
        for (auto& pl : s->scene.planes)
            if (pl.material_id >= s->scene.materials.size())
                throw rt::SceneError(RT_ERR_SCENE, "plane material index out of range");
        for (auto& m : s->scene.meshes)
            for (auto& t : m.triangles)
                if (t.material_id >= s->scene.materials.size())
                    throw rt::SceneError(RT_ERR_SCENE, "mesh material index out of range");
        s->scene.finalize();
    });
}

int rt_scene_synthetic_field(const char* assets_dir, uint32_t grid, uint32_t seed,
                             RtHostScene** out, char* err, size_t err_len) {
    if (!assets_dir || !out || grid == 0) return RT_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    RtHostScene* s = new RtHostScene();
    int st = guarded(err, err_len,
                     [&] { rt::build_synthetic_field(assets_dir, grid, seed, s->scene); });
    if (st != RT_OK) {
        delete s;
        return st;
    }
    *out = s;
    return RT_OK;
}

int rt_scene_synthetic_field_ex(const char* assets_dir, uint32_t grid, uint32_t seed, int build_bvh,
                                RtHostScene** out, char* err, size_t err_len) {
    if (!assets_dir || !out || grid == 0) return RT_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    RtHostScene* s = new RtHostScene();
    int st = guarded(err, err_len,
                     [&] { rt::build_synthetic_field(assets_dir, grid, seed, s->scene, build_bvh != 0); });
    if (st != RT_OK) {
        delete s;
        return st;
    }
    *out = s;
    return RT_OK;
}

int rt_scene_bvh_input(const RtHostScene* s, RtBvhBuildInput* out) {
    if (!s || !out) return RT_ERR_INVALID_ARGUMENT;
    if (!s->scene.finalized) return RT_ERR_STATE;
    const rt::HostScene& h = s->scene;
    out->spheres = h.u_spheres.data();          out->sphere_count = (uint32_t)h.u_spheres.size();
    out->plane_bounds = h.plane_bounds.data();  out->plane_count = (uint32_t)h.u_planes.size();
    out->vertices = h.u_vertices.data();        out->vertex_count = (uint32_t)h.u_vertices.size();
    out->triangles = h.u_triangles.data();      out->triangle_count = (uint32_t)h.u_triangles.size();
    return RT_OK;
}

int rt_scene_set_bvh(RtHostScene* s, const RtBvhNode* nodes, uint32_t node_count,
                     const RtPrimitiveInfo* primitives, uint32_t primitive_count, uint32_t depth) {
    if (!s || !nodes || !primitives || node_count == 0) return RT_ERR_INVALID_ARGUMENT;
    if (!s->scene.finalized) return RT_ERR_STATE;
    rt::HostScene& h = s->scene;
    if ((size_t)primitive_count != h.u_spheres.size() + h.u_planes.size() + h.u_triangles.size())
        return RT_ERR_INVALID_ARGUMENT;
    h.u_nodes.assign(nodes, nodes + node_count);
    h.u_primitives.assign(primitives, primitives + primitive_count);
    h.bvh_depth = depth;
    return RT_OK;
}

int rt_scene_bvh_storage(RtHostScene* s, uint32_t node_count, RtBvhNode** nodes, RtPrimitiveInfo** primitives) {
    if (!s || !nodes || !primitives || node_count == 0) return RT_ERR_INVALID_ARGUMENT;
    if (!s->scene.finalized) return RT_ERR_STATE;
    rt::HostScene& h = s->scene;
    const size_t prims = h.u_spheres.size() + h.u_planes.size() + h.u_triangles.size();
    // The vectors are sized here and filled by the caller (rt_bvh_build_fetch copies straight into
    // them): one pass over the memory instead of a download into a temporary plus a copy.
    h.u_nodes.resize(node_count);
    h.u_primitives.resize(prims);
    h.bvh_depth = 0;
    *nodes = h.u_nodes.data();
    *primitives = h.u_primitives.data();
    return RT_OK;
}

int rt_scene_bvh_commit(RtHostScene* s, uint32_t depth) {
    if (!s) return RT_ERR_INVALID_ARGUMENT;
    if (!s->scene.finalized || s->scene.u_nodes.empty()) return RT_ERR_STATE;
    s->scene.bvh_depth = depth;
    return RT_OK;
}

int rt_scene_build_bvh(RtHostScene* s, char* err, size_t err_len) {
    if (!s) return RT_ERR_INVALID_ARGUMENT;
    if (!s->scene.finalized) return RT_ERR_STATE;
    return guarded(err, err_len, [&] { rt::build_bvh(s->scene); });
}

void rt_scene_free(RtHostScene* s) { delete s; }

int rt_scene_desc(const RtHostScene* s, RtSceneDesc* out) {
    if (!s || !out) return RT_ERR_INVALID_ARGUMENT;
    if (!s->scene.finalized) return RT_ERR_STATE;
    const rt::HostScene& h = s->scene;
    out->materials = h.u_materials.data();   out->material_count = (uint32_t)h.u_materials.size();
    out->spheres = h.u_spheres.data();       out->sphere_count = (uint32_t)h.u_spheres.size();
    out->planes = h.u_planes.data();         out->plane_count = (uint32_t)h.u_planes.size();
    out->vertices = h.u_vertices.data();     out->vertex_count = (uint32_t)h.u_vertices.size();
    out->normals = h.u_normals.data();       out->normal_count = (uint32_t)h.u_normals.size();
    out->triangles = h.u_triangles.data();   out->triangle_count = (uint32_t)h.u_triangles.size();
    out->primitives = h.u_primitives.data(); out->primitive_count = (uint32_t)h.u_primitives.size();
    out->bvh_nodes = h.u_nodes.data();       out->bvh_node_count = (uint32_t)h.u_nodes.size();
    return RT_OK;
}

int rt_scene_camera(const RtHostScene* s, float pos[3], float* yaw, float* pitch, float* fov_y) {
    if (!s || !pos || !yaw || !pitch || !fov_y) return RT_ERR_INVALID_ARGUMENT;
    const rt::Camera& c = s->scene.camera;
    pos[0] = c.pos.x; pos[1] = c.pos.y; pos[2] = c.pos.z;
    *yaw = c.yaw; *pitch = c.pitch; *fov_y = c.fov_y;
    return RT_OK;
}

uint32_t rt_scene_bvh_depth(const RtHostScene* s) { return s ? s->scene.bvh_depth : 0; }

void rt_camera_uniform(const float pos[3], float yaw, float pitch, float fov_y,
                       RtCameraUniform* out) {
    rt::camera_uniform(rt::Camera{v3p(pos), yaw, pitch, fov_y}, out);
}

int rt_camera_serialize(const float pos[3], float yaw, float pitch, float fov_y, char* out,
                        size_t out_len) {
    if (!pos || !out || out_len < 33) return RT_ERR_INVALID_ARGUMENT;
    std::string s = rt::camera_serialize(rt::Camera{v3p(pos), yaw, pitch, fov_y});
    std::memcpy(out, s.c_str(), s.size() + 1);
    return RT_OK;
}

int rt_camera_deserialize(const char* encoded, float pos[3], float* yaw, float* pitch,
                          float* fov_y, char* err, size_t err_len) {
    if (!encoded || !pos || !yaw || !pitch || !fov_y) return RT_ERR_INVALID_ARGUMENT;
    return guarded(err, err_len, [&] {
        rt::Camera c = rt::camera_deserialize(encoded);
        pos[0] = c.pos.x; pos[1] = c.pos.y; pos[2] = c.pos.z;
        *yaw = c.yaw; *pitch = c.pitch; *fov_y = c.fov_y;
    });
}

int rt_hdr_decode(const uint8_t* bytes, size_t len, uint32_t* width, uint32_t* height,
                  float** rgb, char* err, size_t err_len) {
    if (!bytes || !width || !height || !rgb) return RT_ERR_INVALID_ARGUMENT;
    *rgb = nullptr;
    return guarded(err, err_len, [&] {
        std::vector<float> pixels;
        rt::hdr_decode(bytes, len, width, height, pixels);
        float* buf = (float*)std::malloc(pixels.size() * sizeof(float));
        if (!buf) throw std::bad_alloc();
        std::memcpy(buf, pixels.data(), pixels.size() * sizeof(float));
        *rgb = buf;
    });
}

int rt_hdr_load(const char* path, uint32_t* width, uint32_t* height, float** rgb, char* err,
                size_t err_len) {
    if (!path) return RT_ERR_INVALID_ARGUMENT;
    std::ifstream in(path, std::ios::binary);
    if (!in) {
        set_err(err, err_len, std::string("Cannot open '") + path + "'");
        return RT_ERR_IO;
    }
    std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    return rt_hdr_decode(bytes.data(), bytes.size(), width, height, rgb, err, err_len);
}

int rt_hdr_write(const char* path, uint32_t width, uint32_t height, const float* rgb) {
    if (!path || !rgb || !width || !height) return RT_ERR_INVALID_ARGUMENT;
    std::vector<uint8_t> bytes = rt::hdr_encode_flat(width, height, rgb);
    std::ofstream out(path, std::ios::binary);
    if (!out) return RT_ERR_IO;
    out.write((const char*)bytes.data(), (std::streamsize)bytes.size());
    return out ? RT_OK : RT_ERR_IO;
}

void rt_hdri_synthetic(uint32_t width, uint32_t height, float* rgb) {
    if (rgb && width && height) rt::hdri_synthetic(width, height, rgb);
}

int rt_alias_build(uint32_t width, uint32_t height, const float* rgb, RtAliasEntry* out,
                   uint32_t* leftover_count) {
    if (!rgb || !out || !width || !height) return RT_ERR_INVALID_ARGUMENT;
    uint32_t leftover = rt::alias_build(width, height, rgb, out);
    if (leftover_count) *leftover_count = leftover;
    return RT_OK;
}

int rt_alias_build_serial(uint32_t width, uint32_t height, const float* rgb, RtAliasEntry* out,
                          uint32_t* leftover_count) {
    if (!rgb || !out || !width || !height) return RT_ERR_INVALID_ARGUMENT;
    uint32_t leftover = rt::alias_build_serial(width, height, rgb, out);
    if (leftover_count) *leftover_count = leftover;
    return RT_OK;
}

void rt_rgb_to_rgba(uint32_t pixel_count, const float* rgb, float* rgba) {
    for (uint32_t i = 0; i < pixel_count; i++) {
        rgba[4 * (size_t)i + 0] = rgb[3 * (size_t)i + 0];
        rgba[4 * (size_t)i + 1] = rgb[3 * (size_t)i + 1];
        rgba[4 * (size_t)i + 2] = rgb[3 * (size_t)i + 2];
        rgba[4 * (size_t)i + 3] = 0.f;
    }
}

int rt_png_write_rgb8(const char* path, uint32_t width, uint32_t height, const uint8_t* rgb8) {
    if (!path || !rgb8 || !width || !height) return RT_ERR_INVALID_ARGUMENT;
    std::vector<uint8_t> raw((size_t)height * (1 + (size_t)width * 3));
    for (uint32_t y = 0; y < height; y++) {
        uint8_t* row = raw.data() + (size_t)y * (1 + (size_t)width * 3);
        row[0] = 0;  // filter: none
        std::memcpy(row + 1, rgb8 + (size_t)y * width * 3, (size_t)width * 3);
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return RT_ERR_IO;

    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    uint8_t ihdr[13] = {(uint8_t)(width >> 24),  (uint8_t)(width >> 16),  (uint8_t)(width >> 8),
                        (uint8_t)width,          (uint8_t)(height >> 24), (uint8_t)(height >> 16),
                        (uint8_t)(height >> 8),  (uint8_t)height,         8, 2, 0, 0, 0};
    png_chunk(out, "IHDR", ihdr, 13);
    png_chunk(out, "IDAT", z.data(), zlen);
    png_chunk(out, "IEND", nullptr, 0);
    std::ofstream f(path, std::ios::binary);
    if (!f) return RT_ERR_IO;
    f.write((const char*)out.data(), (std::streamsize)out.size());
    return f ? RT_OK : RT_ERR_IO;
}

int rt_pfm_write_rgb(const char* path, uint32_t width, uint32_t height, const float* data,
                     uint32_t channels_in) {
    if (!path || !data || !width || !height || channels_in < 3) return RT_ERR_INVALID_ARGUMENT;
    std::ofstream f(path, std::ios::binary);
    if (!f) return RT_ERR_IO;
    // (snprintf, not operator<<: formatted stream output needs the C++ locale machinery, which
    // is not reliably initialised when this library is loaded into a process that already
    // carries another copy of the C++ runtime, e.g. Python with numpy)
    char header[64];
    const int header_len = snprintf(header, sizeof header, "PF\n%u %u\n-1.0\n", width, height);
    f.write(header, header_len);
    std::vector<float> row((size_t)width * 3);
    for (uint32_t y = height; y-- > 0;) {  // PFM stores the bottom row first
        for (uint32_t x = 0; x < width; x++)
            for (int c = 0; c < 3; c++)
                row[(size_t)x * 3 + c] = data[((size_t)y * width + x) * channels_in + c];
        f.write((const char*)row.data(), (std::streamsize)(row.size() * sizeof(float)));
    }
    return f ? RT_OK : RT_ERR_IO;
}

void rt_free(void* p) { std::free(p); }

}  // extern "C"
