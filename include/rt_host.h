/* rt_host.h — C ABI of the host front-end (librt_host.so).
 *
 * Replaces the CPU side of the reference that feeds the per-pixel path:
 *   Scene::load_toml            src/scene.rs:235-252, 324-441
 *   Mesh::load / pack_meshes    src/mesh.rs:29-113          (+ wavefront_obj 11.0.0 semantics)
 *   build_bvh                   src/bvh.rs:13-28, 120-337
 *   AliasTable::build_by_luminance  src/environments.rs:97-186
 *   Camera / CameraUniform      src/camera.rs:23-119
 *   Radiance .hdr decode        image 0.25.9 (call site src/state.rs:123-130)
 *   RGB -> RGBA texel padding   src/texture.rs:96-118
 *
 * All functions return an RT_* status (rt_types.h). Functions that can fail for
 * user-visible reasons take (err, err_len) and write the reference's message text.
 * Nothing here touches a GPU.
 */
#ifndef RT_HOST_H
#define RT_HOST_H

#include <stddef.h>
#include <stdint.h>
#include "rt_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RtHostScene RtHostScene;

/* ---- scene loading: Scene::load_toml (src/scene.rs:235-252) ---- */
int rt_scene_load_toml(const char* path, RtHostScene** out, char* err, size_t err_len);

/* ---- programmatic construction (same build path as the TOML loader) ---- */
int rt_scene_create(RtHostScene** out);
int rt_scene_add_material(RtHostScene* s, const float color[3], float roughness, float metallic,
                          const float emission[3]);
int rt_scene_add_sphere(RtHostScene* s, const float pos[3], float radius, uint32_t material_id);
int rt_scene_add_plane(RtHostScene* s, const float pos[3], const float forward[3],
                       const float right[3], uint32_t material_id);
/* Mesh::load (src/mesh.rs:29-81) on OBJ text. */
int rt_scene_add_mesh_obj(RtHostScene* s, const char* obj_text, uint32_t material_id, char* err,
                          size_t err_len);
/* Raw triangle soup (already-parsed mesh): vertices/normals are xyz triples; triangles are
 * 6 indices each (v0,v1,v2,n0,n1,n2) local to this mesh. */
int rt_scene_add_mesh_raw(RtHostScene* s, const float* vertices, uint32_t vertex_count,
                          const float* normals, uint32_t normal_count, const uint32_t* triangles,
                          uint32_t triangle_count, uint32_t material_id);
int rt_scene_set_camera(RtHostScene* s, const float pos[3], float yaw_rad, float pitch_rad,
                        float fov_y_rad);
/* PackedMeshes::pack_meshes + Plane::to_uniform + build_bvh. Must be called once before
 * rt_scene_desc() on a scene made with rt_scene_create(). */
int rt_scene_finalize(RtHostScene* s, char* err, size_t err_len);

/* Synthetic "config D" scene: the house + suzanne triangles instanced on a
 * grid x grid field over one ground plane. assets_dir holds the reference OBJ files. */
int rt_scene_synthetic_field(const char* assets_dir, uint32_t grid, uint32_t seed,
                             RtHostScene** out, char* err, size_t err_len);

/* The same scene with or without its BVH: build_bvh = 0 leaves the node and primitive arrays empty
 * for a device build (rt_bvh_build, rt_backend.h) that rt_scene_set_bvh then installs. */
int rt_scene_synthetic_field_ex(const char* assets_dir, uint32_t grid, uint32_t seed, int build_bvh,
                                RtHostScene** out, char* err, size_t err_len);
/* Input of build_bvh (src/bvh.rs:13-28) as the device builder takes it: the packed primitive
 * arrays plus Plane::bounds (scene.rs:203-207) of every plane. Pointers stay valid until
 * rt_scene_free(). */
int rt_scene_bvh_input(const RtHostScene* s, RtBvhBuildInput* out);
/* Installs a BVH built elsewhere (rt_bvh_build): flattened nodes, leaf-ordered primitives, depth. */
int rt_scene_set_bvh(RtHostScene* s, const RtBvhNode* nodes, uint32_t node_count,
                     const RtPrimitiveInfo* primitives, uint32_t primitive_count, uint32_t depth);
/* The same without the intermediate copy: rt_scene_bvh_storage sizes the scene's own node and
 * primitive arrays (primitive count = spheres + planes + triangles) and returns where to write them
 * — pass the pointers to rt_bvh_build_fetch — and rt_scene_bvh_commit records the depth. */
int rt_scene_bvh_storage(RtHostScene* s, uint32_t node_count, RtBvhNode** nodes, RtPrimitiveInfo** primitives);
int rt_scene_bvh_commit(RtHostScene* s, uint32_t depth);
/* (Re)builds the BVH with the host builder (what rt_scene_finalize does). */
int rt_scene_build_bvh(RtHostScene* s, char* err, size_t err_len);

void rt_scene_free(RtHostScene* s);

/* Pointers stay valid until rt_scene_free(). */
int rt_scene_desc(const RtHostScene* s, RtSceneDesc* out);
/* Camera as loaded (radians): src/scene.rs:296-314. */
int rt_scene_camera(const RtHostScene* s, float pos[3], float* yaw, float* pitch, float* fov_y);
/* tree_depth of src/bvh.rs:124-134 (leaf = 0). */
uint32_t rt_scene_bvh_depth(const RtHostScene* s);

/* ---- camera: src/camera.rs ---- */
/* CameraUniform::new (camera.rs:111-119); rot = R_y(yaw) * R_x(pitch) (camera.rs:26-28). */
void rt_camera_uniform(const float pos[3], float yaw, float pitch, float fov_y,
                       RtCameraUniform* out);
/* Camera::serialize / deserialize (camera.rs:30-89): base64(24 bytes LE f32 pos.xyz,yaw,pitch,fov_y). */
int rt_camera_serialize(const float pos[3], float yaw, float pitch, float fov_y, char* out,
                        size_t out_len);
int rt_camera_deserialize(const char* encoded, float pos[3], float* yaw, float* pitch,
                          float* fov_y, char* err, size_t err_len);

/* ---- environments ---- */
/* Radiance RGBE decode -> w*h*3 floats (malloc'd; release with rt_free). */
int rt_hdr_load(const char* path, uint32_t* width, uint32_t* height, float** rgb, char* err,
                size_t err_len);
int rt_hdr_decode(const uint8_t* bytes, size_t len, uint32_t* width, uint32_t* height,
                  float** rgb, char* err, size_t err_len);
/* Flat (non-RLE) RGBE writer; used for the synthetic HDRI. */
int rt_hdr_write(const char* path, uint32_t width, uint32_t height, const float* rgb);
/* Deterministic synthetic sky (gradient + small bright sun + dim ground), already
 * RGBE-quantised so a write/read round trip returns the same floats. */
void rt_hdri_synthetic(uint32_t width, uint32_t height, float* rgb);
/* AliasTable::build_by_luminance (environments.rs:97-186). out has width*height entries. */
int rt_alias_build(uint32_t width, uint32_t height, const float* rgb, RtAliasEntry* out,
                   uint32_t* leftover_count);
/* The same table by the literal single-threaded loop of environments.rs:123-178 (two stacks, a
 * residual array, assigned flags): rt_alias_build is multi-threaded and stack-free and must equal
 * this byte for byte. */
int rt_alias_build_serial(uint32_t width, uint32_t height, const float* rgb, RtAliasEntry* out,
                          uint32_t* leftover_count);
/* Texture::from_rgb32float texel padding (texture.rs:104-118): rgba = (r,g,b,0). */
void rt_rgb_to_rgba(uint32_t pixel_count, const float* rgb, float* rgba);

/* ---- image output (headless replacement of the present pass) ---- */
int rt_png_write_rgb8(const char* path, uint32_t width, uint32_t height, const uint8_t* rgb8);
int rt_pfm_write_rgb(const char* path, uint32_t width, uint32_t height, const float* rgba,
                     uint32_t channels_in);

void rt_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* RT_HOST_H */
