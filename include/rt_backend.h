/* rt_backend.h — C ABI of the B200 (sm_100a) path-tracing backend (librt_backend.so).
 *
 * This is the drop-in boundary for the reference's render step. The reference has no
 * plugin API; its de-facto boundary is the wgpu resource set built by State::new and
 * driven by State::update / State::render (src/state.rs:60-649, 722-833). Each entry
 * point below names the piece of that file it replaces. A Rust host binds these with
 * `extern "C"` from build.rs.
 *
 * Conventions: every function returns an RT_* status (rt_types.h); rt_last_error()
 * returns the message of the last failure on that context. A context is bound to one
 * CUDA device and one stream and is not thread-safe; different contexts may be driven
 * from different host threads. Host pointers are only read/written during the call.
 * There is no CPU fallback: without a usable CUDA device rt_create() fails.
 */
#ifndef RT_BACKEND_H
#define RT_BACKEND_H

#include <stddef.h>
#include <stdint.h>
#include "rt_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RtContext RtContext;
typedef struct RtGroup RtGroup;

/* Work counters. rays / shadow_rays / paths are always maintained; the rest only when
 * detailed counting is enabled (rt_set_detailed_counters), at a throughput cost. The
 * layout is fixed so tests can compare the fields one by one. */
typedef struct RtCounters {
    uint64_t paths;           /* pixel-samples traced (shader.wgsl main invocations)        */
    uint64_t rays;            /* closest-hit cast_ray calls (shader.wgsl:1221)              */
    uint64_t shadow_rays;     /* NEE cast_ray_bvh calls actually traced (shader.wgsl:1249)  */
    uint64_t nodes_tested;    /* ray_intersects_bounds evaluations                          */
    uint64_t tris_tested;
    uint64_t spheres_tested;  /* includes brute-force fallback tests (shader.wgsl:583-590)  */
    uint64_t planes_tested;   /* includes brute-force fallback tests (shader.wgsl:591-598)  */
    uint64_t hits;            /* closest-hit rays that hit something                        */
    uint64_t tri_hits;        /* ... of which the BVH hit was a triangle                    */
    uint64_t shaded_bounces;  /* sample_environment calls                                   */
    uint64_t env_lookups;     /* bilinear environment fetches (sky_light + NEE)             */
    uint64_t fallback_rays;   /* cast_ray calls that fell through to the brute-force loop   */
} RtCounters;

const char* rt_backend_version(void);
int rt_device_count(int* count);

/* Device + queue creation: State::new, src/state.rs:71-95. */
int rt_create(int device_ordinal, RtContext** out);
void rt_destroy(RtContext* ctx);
/* ctx == NULL returns the message of the last failed rt_create / rt_group_create. */
const char* rt_last_error(const RtContext* ctx);

/* Scene bind group (group 2): src/state.rs:394-458 + build_bvh upload 446-458.
 * Arrays are in the reference's order and byte layout; the backend re-lays them out
 * privately (128-bit aligned nodes, leaf-ordered pre-gathered primitive records). */
int rt_upload_scene(RtContext* ctx, const RtSceneDesc* scene);

/* Texture bind group (group 0) environment part: EnvironmentMaps::new,
 * src/environments.rs:19-64, src/state.rs:119-142. Replaces all environments. */
int rt_upload_environments(RtContext* ctx, const RtEnvironmentDesc* envs, uint32_t count);

/* Camera uniform write: State::update, src/state.rs:727-735. */
int rt_set_camera(RtContext* ctx, const RtCameraUniform* camera);

/* Accumulation target: HdrPipeline textures (src/hdr.rs:203-241) and State::resize
 * (src/state.rs:651-700). Allocates W*H rgba32f and clears it. */
int rt_resize(RtContext* ctx, uint32_t width, uint32_t height);
/* Render into a caller-owned device buffer of W*H*4 floats instead (e.g. a tensor that
 * a collective will reduce). Pass NULL to return to the internal buffer. */
int rt_bind_accum(RtContext* ctx, void* device_rgba);
int rt_accum_device_ptr(RtContext* ctx, void** device_rgba);
/* Scene-hash change path of State::render (src/state.rs:775-789): zero the accumulation. */
int rt_clear(RtContext* ctx);

/* The compute pass of State::render (src/state.rs:808-824) for `sample_count`
 * consecutive frames whose `sample_count` uniform runs sample_begin .. sample_begin +
 * sample_count - 1 (these values seed the per-pixel RNG, shader.wgsl:1309-1312).
 * Accumulates into the accumulation buffer exactly like shader.wgsl:1367-1371.
 * max_bounces is the reference's MAX_BOUNCES (10, shader.wgsl:232) made a parameter.
 * Asynchronous: the work is enqueued on the context's stream, ordered with everything else
 * on it, and the call returns at once — the wave loop of the render is one CUDA-graph launch
 * (a WHILE node that runs until every path slot has retired). Use rt_sync / rt_read_accum to
 * wait for the result. A render makes no call into the kernel-mode driver unless its buffers
 * have to grow. (With rt_set_profiling on, or RT_WAVE_GRAPH=0, the loop is paced from the
 * calling thread instead and the call returns when the last waves have been queued.) */
int rt_render(RtContext* ctx, uint32_t sample_begin, uint32_t sample_count, uint32_t max_bounces,
              uint32_t environment_index);

/* Parity hook: primary rays only. slot_out[y*W+x] = index into `primitives` of the
 * closest hit of cast_ray for the camera ray of sample `sample_index`;
 * RT_MISS_SLOT on a miss; 0x80000000 | type << 28 | index for hits found only by the
 * brute-force fallback (shader.wgsl:583-598). t_out (optional) = hit distance, 0 on miss.
 * Host output buffers; synchronous. Frames of 2 Mi pixels and more are traced in four bands of
 * rows whose maps travel to the host on a second stream while the next band is traced (page-locked
 * output buffers, rt_host_alloc, let that overlap happen: 4K, 1.68 -> 1.35 ms per call). */
int rt_first_hit(RtContext* ctx, uint32_t sample_index, uint32_t* slot_out, float* t_out);
/* Device time (CUDA events) of the kernel of the last rt_first_hit call, without its copies. */
float rt_first_hit_kernel_ms(const RtContext* ctx);

/* Page-locked host memory for output buffers (rt_first_hit, rt_read_accum): device -> host
 * copies into it run at link speed; any other host memory works too, through the driver's
 * bounce buffer. Free with rt_host_free. */
int rt_host_alloc(size_t bytes, void** out);
void rt_host_free(void* p);

int rt_sync(RtContext* ctx);
/* Device -> host copy of the W*H*4 accumulation floats (alpha = 1 where rendered). */
int rt_read_accum(RtContext* ctx, float* rgba_out);
/* Host -> device copy into the accumulation buffer (resume a progressive render). */
int rt_write_accum(RtContext* ctx, const float* rgba_in);

/* Present pass: out = accum / sample_total rounded to f16 (out_texture is rgba16float,
 * src/hdr.rs:203-216), ACES fit with magenta for negatives (src/shaders/hdr.wgsl:3-22),
 * sRGB encode (surface format, src/hdr.rs:120). rgb8_out is W*H*3 host bytes. */
int rt_tonemap_rgb8(RtContext* ctx, uint32_t sample_total, uint8_t* rgb8_out);

/* The reference's dev_index views (keys 2 and 3, src/camera.rs:279-285; shader.wgsl:1314-1338):
 * the contents of out_texture (rgba16float, W*H*4 floats here) after one dispatch with that
 * dev_index, for the frame whose sample_count uniform is `sample_index`.
 *   2: where the environment's alias table sends samples — every pixel draws 20 texels and adds
 *      0.1 / 20 at their coordinates. The shader's read-modify-write races between invocations;
 *      this applies every draw (a race-free execution of the same code).
 *   3: the environment's texels at the pixel's own coordinates, saturated.
 * Any other dev_index renders normally in the reference and is refused here (use rt_render). */
int rt_debug_view(RtContext* ctx, uint32_t dev_index, uint32_t environment_index, uint32_t sample_index,
                  float* rgba_out);

/* Per-kernel device time of the wavefront pipeline (CUDA events around every launch on the
 * launching stream), accumulated over rt_render calls while profiling is enabled. Kernel
 * order: shade, finish, pre-pass, walk (rsoderh-raytracing_b200/csrc/rt_wave.cuh). */
typedef struct RtWaveProfile {
    uint64_t renders;       /* rt_render calls measured                      */
    uint64_t waves;         /* waves launched (including trailing empty ones) */
    uint64_t launches[4];   /* kernel launches per kind                       */
    double kernel_ms[4];    /* summed device time per kind                    */
    double gap_ms;          /* time between a wave's last kernel and the next wave's first */
    double total_ms;        /* first launch to last completion, summed over renders */
} RtWaveProfile;
int rt_set_profiling(RtContext* ctx, int enabled);   /* also clears the accumulated profile */
int rt_get_profile(RtContext* ctx, RtWaveProfile* out);

/* Device memory a context takes besides the scene and the environments (64 MiB per 2k
 * environment): the accumulation buffer (16 B per pixel), the path-slot arena of the render
 * pipeline (162 B per slot; default 8 Mi slots = 1.3 GB) and the sample-staging planes that
 * keep the per-pixel addition order of shader.wgsl:1367-1371 (16 B per pixel and staged
 * sample; default at most 8 GiB, and never more than 3/4 of what the device has free when
 * rt_render runs). rt_set_memory_budget lowers (or raises) the two knobs; 0 keeps a value.
 * Measured on house.toml 1080p, 256 spp: 8 GiB staging 7.90 Grays/s, 4 GiB 7.88, 2 GiB 7.82
 * (-1.0 %), 1 GiB 7.72 (-2.3 %), 512 MiB 7.52; 4 Mi slots 7.65, 2 Mi 7.28 — so 2 GiB + 8 Mi slots
 * (3.4 GB) or 1 GiB + 8 Mi slots (2.4 GB) are the settings for a GPU that is shared. Renders that
 * bring fewer than three paths per slot use a smaller part of the pool on their own.
 * rt_get_memory_usage reports what is allocated right now. */
int rt_set_memory_budget(RtContext* ctx, uint64_t staging_bytes, uint32_t path_slots);
int rt_get_memory_usage(RtContext* ctx, uint64_t* bytes);

/* build_bvh (src/bvh.rs:13-28; build_sah :215-337, flatten :155-178) on the device. Produces the
 * node array and leaf-ordered primitive list of rt_scene_finalize's host builder byte for byte —
 * same pre-order node indices, same slot order (the reference's sequential swap partition,
 * bvh.rs:304-315, has a closed form that is evaluated in parallel), same bounds down to the sign
 * of a zero — one level of the tree per pass (csrc/rt_bvh_build.cuh). rt_bvh_build uploads the
 * input, builds, and leaves the result on the device; rt_bvh_build_fetch copies it out (either
 * pointer may be NULL) and releases it. node_count_out <= 2 * primitives - 1. device_ms_out
 * (optional) is the time of the device work alone, uploads and downloads excluded.
 * Errors: RT_ERR_SCENE "scene has no primitives" (bvh.rs:222) and "bounds have non-zero surface
 * area" (a NaN SAH cost, bvh.rs:296), like the host builder. */
int rt_bvh_build(RtContext* ctx, const RtBvhBuildInput* input, uint32_t* node_count_out,
                 uint32_t* depth_out, float* device_ms_out);
int rt_bvh_build_fetch(RtContext* ctx, RtBvhNode* nodes_out, RtPrimitiveInfo* primitives_out);

int rt_set_detailed_counters(RtContext* ctx, int enabled);
int rt_reset_counters(RtContext* ctx);
int rt_get_counters(RtContext* ctx, RtCounters* out); /* synchronises */

/* Stream interop: enqueue on a caller-provided cudaStream_t. The handle is used as is, so
 * NULL means the CUDA legacy default stream. rt_reset_stream() returns to the context's
 * own non-blocking stream. */
int rt_set_stream(RtContext* ctx, void* cuda_stream);
int rt_reset_stream(RtContext* ctx);
/* CUDA-event timing on the context's stream. */
int rt_timer_start(RtContext* ctx);
int rt_timer_stop(RtContext* ctx, float* elapsed_ms); /* synchronises */
/* Number of kernels this context has launched since creation. */
uint64_t rt_launch_count(const RtContext* ctx);

/* ---- single-process multi-GPU: sample-range split + NCCL sum-reduce ---- */
int rt_group_create(int count, const int* device_ordinals, RtGroup** out);
void rt_group_destroy(RtGroup* group);
int rt_group_size(const RtGroup* group);
RtContext* rt_group_context(RtGroup* group, int rank);
/* Splits [sample_begin, sample_begin+sample_count) into contiguous per-GPU ranges. */
int rt_group_render(RtGroup* group, uint32_t sample_begin, uint32_t sample_count,
                    uint32_t max_bounces, uint32_t environment_index);
/* ncclReduce(sum, f32) of the accumulation buffers onto `root`. */
int rt_group_reduce(RtGroup* group, int root);
int rt_group_sync(RtGroup* group);

#ifdef __cplusplus
}
#endif
#endif /* RT_BACKEND_H */
