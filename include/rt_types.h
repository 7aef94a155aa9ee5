/* rt_types.h — byte layouts shared by the host front-end (rt_host.h) and the CUDA
 * backend (rt_backend.h).
 *
 * Every struct below is byte-identical to the storage/uniform buffer the reference
 * uploads to wgpu through `encase` (WGSL std430-style layout), so a Rust caller can
 * hand `encase::StorageBuffer::content_of(..)` bytes straight to rt_upload_scene().
 * Reference citations are relative to /root/reference.
 */
#ifndef RT_TYPES_H
#define RT_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/scene.rs:16-23, src/shaders/shader.wgsl:86-91 — 48 B stride */
typedef struct RtMaterial {
    float color[3];     /* @0  */
    float roughness;    /* @12 */
    float metallic;     /* @16 */
    float _pad0[3];
    float emission[3];  /* @32 */
    float _pad1;
} RtMaterial;

/* src/scene.rs:166-171, shader.wgsl:93-97 — 32 B stride */
typedef struct RtSphere {
    float pos[3];          /* @0  */
    float radius;          /* @12 */
    uint32_t material_id;  /* @16 */
    uint32_t _pad[3];
} RtSphere;

/* src/scene.rs:210-222 (UniformPlane), shader.wgsl:99-110 — 96 B stride.
 * base_change holds the three COLUMNS of the mat3x3, each padded to 16 B. */
typedef struct RtPlane {
    float pos[3];             /* @0  */
    float _pad0;
    float normal[3];          /* @16 */
    float _pad1;
    float base_change[3][4];  /* @32, column-major, 16 B per column */
    uint32_t material_id;     /* @80 */
    uint32_t _pad2[3];
} RtPlane;

/* array<vec3<f32>> has a 16 B stride: shader.wgsl:217-221 */
typedef struct RtVec3 {
    float x, y, z, _pad;
} RtVec3;

/* src/mesh.rs:150-165, shader.wgsl:112-126 — 28 B stride */
typedef struct RtTriangle {
    uint32_t vertex_0, vertex_1, vertex_2;
    uint32_t normal_0, normal_1, normal_2;
    uint32_t material_id;
} RtTriangle;

/* src/bvh.rs:81-87, shader.wgsl:128-133 — 8 B stride.
 * primitive_type: 0 sphere, 1 plane, 2 triangle (src/bvh.rs:48,54,65). */
typedef struct RtPrimitiveInfo {
    uint32_t primitive_type;
    uint32_t index;
} RtPrimitiveInfo;

/* src/bvh.rs:89-99, shader.wgsl:135-146 — 48 B stride */
typedef struct RtBvhNode {
    float min[3];  /* @0  */
    float _pad0;
    float max[3];  /* @16 */
    float _pad1;
    uint32_t primitives_or_second_child_index; /* @32 */
    uint32_t primitives_len;                   /* @36; 0 for interior nodes */
    uint32_t split_axis;                       /* @40 */
    uint32_t _pad2;
} RtBvhNode;

/* src/environments.rs:200-213, shader.wgsl:158-170 — 16 B stride */
typedef struct RtAliasEntry {
    float probability;
    uint32_t alias_index;
    float pmf;
    uint32_t _pad;
} RtAliasEntry;

/* src/camera.rs:103-109, shader.wgsl:1-6 — 80 B uniform.
 * rot holds the three COLUMNS of rot_transform, 16 B each. */
typedef struct RtCameraUniform {
    float pos[3];     /* @0  */
    float _pad0;
    float rot[3][4];  /* @16 */
    float fov_y;      /* @64, radians */
    float _pad1[3];
} RtCameraUniform;

/* The scene bind group (group 2) of the reference: src/state.rs:394-458,
 * shader.wgsl:208-230. Arrays keep the reference's order and indices. */
typedef struct RtSceneDesc {
    const RtMaterial* materials;        uint32_t material_count;
    const RtSphere* spheres;            uint32_t sphere_count;
    const RtPlane* planes;              uint32_t plane_count;
    const RtVec3* vertices;             uint32_t vertex_count;
    const RtVec3* normals;              uint32_t normal_count;
    const RtTriangle* triangles;        uint32_t triangle_count;
    const RtPrimitiveInfo* primitives;  uint32_t primitive_count;
    const RtBvhNode* bvh_nodes;         uint32_t bvh_node_count;
} RtSceneDesc;

/* Input of build_bvh (src/bvh.rs:13-28): the primitive arrays PrimitiveInfo::build_array walks
 * (bvh.rs:40-72) — spheres, then planes, then triangles. A baked RtPlane no longer holds the
 * `forward` / `right` edges its bounds are made of (Plane::bounds, scene.rs:203-207: the box of
 * `pos` and `pos + forward + right`), so plane boxes come ready made: 6 floats per plane,
 * min xyz then max xyz (rt_scene_bvh_input fills them). */
typedef struct RtBvhBuildInput {
    const RtSphere* spheres;      uint32_t sphere_count;
    const float* plane_bounds;    uint32_t plane_count;
    const RtVec3* vertices;       uint32_t vertex_count;
    const RtTriangle* triangles;  uint32_t triangle_count;
} RtBvhBuildInput;

/* One environment: the rgba32float texture of src/texture.rs:96-149 (alpha = 0) plus its
 * alias table (src/environments.rs:97-186). alias has width*height entries. */
typedef struct RtEnvironmentDesc {
    uint32_t width, height;
    const float* rgba;          /* width*height*4 floats, row-major, row 0 = top */
    const RtAliasEntry* alias;  /* width*height entries */
} RtEnvironmentDesc;

/* Status codes returned by every rt_* function. */
enum {
    RT_OK = 0,
    RT_ERR_INVALID_ARGUMENT = 1,
    RT_ERR_CUDA = 2,
    RT_ERR_OUT_OF_MEMORY = 3,
    RT_ERR_NCCL = 4,
    RT_ERR_IO = 5,
    RT_ERR_PARSE = 6,
    RT_ERR_SCENE = 7,
    RT_ERR_STATE = 8
};

#define RT_MISS_SLOT 0xFFFFFFFFu

#ifdef __cplusplus
}
#endif
#endif /* RT_TYPES_H */
