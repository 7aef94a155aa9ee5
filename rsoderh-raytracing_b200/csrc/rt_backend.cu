// rt_backend.cu — kernels and extern "C" surface of librt_backend.so (include/rt_backend.h).
//
// Replaces the reference's device/pipeline layer (src/state.rs) and its compute shader
// (src/shaders/shader.wgsl) with sm_100a CUDA. The render path is the wavefront pipeline of
// rt_wave.cuh (four dense kernels per wave for two overlapping slot groups, the wave loop one
// CUDA graph with a WHILE node: wave_graph_build / wave_render below). This file holds the C ABI,
// the uploads (two scene arenas, so a re-upload allocates nothing), the host side of the device
// BVH build (rt_bvh_build; kernels in rt_bvh_build.cuh), the NCCL group, and the kernels that are
// NOT the product path but cross-checks of it:
// `render_kernel` / `first_hit_kernel` below walk the reference-layout arrays in the
// reference's order, one thread per pixel, and count node / primitive tests exactly like the
// CPU restatement the tests compare with (rt_set_detailed_counters; RT_KERNEL=pixel) — the
// source of the algorithmic work figures bench.py reports. RT_KERNEL=local selects the
// persistent single-launch kernel of rt_local.cuh, RT_KERNEL=fast the thread-per-pixel
// megakernel of rt_fast.cuh; the tests check that all four agree bit for bit.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <time.h>
#include <unistd.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "rt_backend.h"
#include "rt_device.cuh"
#include "rt_fast.cuh"
#include "rt_fast_host.hpp"
#include "rt_wave.cuh"
#include "rt_local.cuh"
#include "rt_bvh_build.cuh"

using namespace rtd;

// =====================================================================================
// Kernels
// =====================================================================================

struct RenderParams {
    SceneView scene;
    EnvView env;
    CameraParams camera;
    float4* accum;                 // W*H rgba32f
    unsigned long long* counters;  // RtCounters as 12 x u64
    uint32_t width, height;
    uint32_t sample_begin, sample_count;
    uint32_t max_bounces;
};

constexpr int kTileW = 16, kTileH = 8, kBlockThreads = kTileW * kTileH;

// Pixel owned by this thread: a warp covers an 8x4 patch, a block a 16x8 tile.
RT_DEV void thread_pixel(uint32_t& px, uint32_t& py) {
    uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    px = blockIdx.x * kTileW + (warp & 1u) * 8u + (lane & 7u);
    py = blockIdx.y * kTileH + (warp >> 1) * 4u + (lane >> 3);
}

RT_DEV void flush_counts(const Counts& cn, unsigned long long* out, bool detailed) {
    const uint32_t vals[12] = {cn.paths, cn.rays,    cn.shadow_rays, cn.nodes,  cn.tris,        cn.spheres,
                               cn.planes, cn.hits,   cn.tri_hits,    cn.shaded, cn.env_lookups, cn.fallback};
    int n = detailed ? 12 : 3;
    for (int i = 0; i < n; i++) {
        uint32_t v = vals[i];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
        if ((threadIdx.x & 31u) == 0u && v) atomicAdd(out + i, (unsigned long long)v);
    }
}

#ifndef RT_MIN_BLOCKS
#define RT_MIN_BLOCKS 6
#endif
template <bool COUNT, int STACK_CAP>
__global__ void __launch_bounds__(kBlockThreads, RT_MIN_BLOCKS) render_kernel(const RenderParams p) {
    uint32_t px, py;
    thread_pixel(px, py);
    const bool active = px < p.width && py < p.height;
    const size_t pixel = (size_t)py * p.width + px;

    Counts cn;
    memset(&cn, 0, sizeof cn);
    uint32_t stack[STACK_CAP];

    const float sin_half_fov = rt_sin(p.camera.fov_y / 2.0f);
    const uint32_t sample_end = p.sample_begin + p.sample_count;

    float4 acc4 = active ? p.accum[pixel] : make_float4(0.f, 0.f, 0.f, 0.f);
    V3 acc = v3(acc4.x, acc4.y, acc4.z);

    uint32_t sample = (active && p.max_bounces > 0u) ? p.sample_begin : sample_end;

    // Path state.
    uint32_t rng = 0u, bounce = 0u;
    V3 o = v3s(0.f), d = v3s(0.f), T = v3s(1.f), L = v3s(0.f);
    float last_pdf = 1.0f;
    if (sample < sample_end) {
        camera_ray(p.camera, sin_half_fov, px, py, p.width, p.height, sample, rng, o, d);
        cn.paths++;
    }

    while (sample < sample_end) {
        bool done = false;
        Hit hit;
        cn.rays++;
        trace<COUNT>(p.scene, o, d, false, true, hit, stack, cn);

        if (hit.rec == 0xFFFFFFFFu) {
            // shader.wgsl:1222-1231 — escaped into the environment.
            float u, v;
            direction_to_uv(d, u, v);
            V3 sky = env_bilinear(p.env, u, v);
            if (COUNT) cn.env_lookups++;
            float pdf = env_direction_pdf(p.env, u, v);
            float w = power_heuristic(last_pdf, pdf);
            L = L + T * sky * w;
            done = true;
        } else {
            V3 point, n;
            uint32_t mat_id;
            reconstruct_hit(p.scene, o, d, hit, point, n, mat_id);
            if (COUNT) {
                cn.hits++;
                if (hit.rec < p.scene.slot_count &&
                    __float_as_int(ld4(p.scene.records + 4u * (size_t)hit.rec).w) == 2)
                    cn.tri_hits++;
            }
            Material m = load_material(p.scene, mat_id);
            L = L + T * m.emission;  // shader.wgsl:1236

            V3 wo_world = -d;
            Frame frame = make_frame(n);
            V3 wo = to_local(frame, wo_world);

            // ---- next-event estimation: sample_environment (shader.wgsl:782-820) ----
            {
                if (COUNT) cn.shaded++;
                uint32_t length = p.env.width * p.env.height;
                uint32_t index = min((uint32_t)(rng_uniform(rng) * (float)length), length - 1u);
                float4 entry = ld4(p.env.alias + index);
                uint32_t pick = rng_uniform(rng) < entry.x ? index : __float_as_uint(entry.y);
                uint32_t ex = pick % p.env.width, ey = pick / p.env.width;
                float jx = rng_uniform(rng), jy = rng_uniform(rng);
                float eu = ((float)ex + jx) / (float)p.env.width;
                float ev = ((float)ey + jy) / (float)p.env.height;
                float phi = (2.0f * eu - 1.0f) * RT_PI;
                float theta = RT_PI * ev;
                float st, ct, sph, cph;
                rt_sincos(theta, &st, &ct);
                rt_sincos(phi, &sph, &cph);
                V3 wi_world = v3(st * cph, ct, st * sph);
                float pmf = pick == index ? entry.z : ld4(p.env.alias + pick).z;
                // environment_pixel_solid_angle (shader.wgsl:739-749); sin(theta) is `st`.
                float d_phi = RT_TWO_PI / (float)p.env.width;
                float d_theta = RT_PI / (float)p.env.height;
                float pdf_env = pmf / (d_phi * d_theta * fmaxf(1.0e-6f, st));
                float cos_theta = fmaxf(0.0f, dot(n, wi_world));
                // The reference fetches the radiance unconditionally; it is only consumed
                // when the sample is unoccluded, so fetch it there.
                if (COUNT) cn.env_lookups++;
                if (cos_theta > 0.0f && pdf_env > 0.0f) {
                    Hit shadow;
                    cn.shadow_rays++;
                    trace<COUNT>(p.scene, point, wi_world, true, false, shadow, stack, cn);
                    if (shadow.rec == 0xFFFFFFFFu) {
                        V3 radiance = env_bilinear(p.env, eu, ev);
                        V3 wi = to_local(frame, wi_world);
                        V3 f;
                        float pdf_bsdf;
                        bsdf_eval_pdf(wo, wi, m, f, pdf_bsdf);
                        float w = power_heuristic(pdf_env, pdf_bsdf);
                        L = L + T * w * radiance * f * cos_theta / pdf_env;
                    }
                }
            }

            // ---- bsdf_sample (shader.wgsl:1116-1202) + path update (1270-1298) ----
            if (dot(n, wo_world) <= 0.0f) {
                L = v3(0.f, 0.f, 1.f);
                done = true;
            } else if (wo.z <= 0.0f) {
                L = v3(0.f, 1.f, 0.f);
                done = true;
            } else {
                float sample_u = rng_uniform(rng);
                V3 wi;
                bool blocked = false;
                if (sample_u < m.p_diff) {
                    float sx = sample_u / fmaxf(m.p_diff, 1.e-6f);
                    float sy = rng_uniform(rng);
                    // sample_cosine_hemisphere (shader.wgsl:892-901)
                    float r = sqrtf(sx);
                    float s, c;
                    rt_sincos(RT_TWO_PI * sy, &s, &c);
                    float x = r * c, y = r * s;
                    wi = v3(x, y, sqrtf(fmaxf(0.0f, 1.0f - x * x - y * y)));
                } else {
                    float sx = (sample_u - m.p_diff) / fmaxf(m.p_spec, 1.e-6f);
                    float sy = rng_uniform(rng);
                    V3 h = sample_ggx_vndf(sx, sy, wo, m.alpha);
                    wi = reflect(-wo, h);
                    blocked = wi.z <= 0.0f;
                }
                if (blocked) {
                    done = true;  // direction (1,0,0) with pdf 0: path ends, L kept
                } else {
                    V3 f;
                    float pdf;
                    bsdf_eval_pdf(wo, wi, m, f, pdf);
                    V3 wi_world = to_world(frame, wi);
                    if (dot(n, wi_world) < 0.0f) {
                        L = v3(0.f, 1.f, 0.f);
                        done = true;
                    } else if (pdf <= 0.0f) {
                        done = true;
                    } else {
                        float cos_theta = fmaxf(0.0f, dot(n, wi_world));
                        T = T * (f * (cos_theta / pdf));
                        if (length(T) < 0.001f) {
                            done = true;
                        } else {
                            last_pdf = pdf;
                            o = point;
                            d = wi_world;
                        }
                    }
                }
            }
        }

        bounce++;
        if (done || bounce >= p.max_bounces) {
            acc = acc + L;  // shader.wgsl:1367-1368
            sample++;
            if (sample < sample_end) {
                camera_ray(p.camera, sin_half_fov, px, py, p.width, p.height, sample, rng, o, d);
                cn.paths++;
                T = v3s(1.f);
                L = v3s(0.f);
                last_pdf = 1.0f;
                bounce = 0u;
            }
        }
    }

    if (active) {
        if (p.max_bounces == 0u) cn.paths += p.sample_count;
        p.accum[pixel] = make_float4(acc.x, acc.y, acc.z, 1.0f);  // shader.wgsl:1371
    }
    flush_counts(cn, p.counters, COUNT);
}

struct FirstHitParams {
    SceneView scene;
    CameraParams camera;
    uint32_t* slot_out;
    float* t_out;
    unsigned long long* counters;
    uint32_t width, height, sample;
};

template <bool COUNT, int STACK_CAP>
__global__ void __launch_bounds__(kBlockThreads) first_hit_kernel(const FirstHitParams p) {
    uint32_t px, py;
    thread_pixel(px, py);
    Counts cn;
    memset(&cn, 0, sizeof cn);
    if (px < p.width && py < p.height) {
        uint32_t stack[STACK_CAP];
        uint32_t rng;
        V3 o, d;
        camera_ray(p.camera, rt_sin(p.camera.fov_y / 2.0f), px, py, p.width, p.height, p.sample,
                   rng, o, d);
        Hit hit;
        cn.paths++;
        cn.rays++;
        trace<COUNT>(p.scene, o, d, false, true, hit, stack, cn);
        uint32_t slot = hit.rec;
        if (hit.rec != 0xFFFFFFFFu && hit.rec >= p.scene.slot_count) {
            uint32_t k = hit.rec - p.scene.slot_count;
            slot = k < p.scene.sphere_count ? (0x80000000u | k)
                                            : (0x90000000u | (k - p.scene.sphere_count));
        }
        size_t i = (size_t)py * p.width + px;
        p.slot_out[i] = slot;
        p.t_out[i] = hit.rec != 0xFFFFFFFFu ? hit.t : 0.0f;
    }
    flush_counts(cn, p.counters, COUNT);
}

__global__ void bake_materials_kernel(const float* src, float4* dst, uint32_t count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) bake_material(src + 12u * (size_t)i, dst + 3u * (size_t)i);
}

// max_bounces == 0: trace_ray's loop body never runs (shader.wgsl:1220); each sample adds 0.
__global__ void zero_bounce_kernel(float4* accum, size_t count, unsigned long long* counters,
                                   uint32_t samples) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) accum[i].w = 1.0f;
    if (i == 0) atomicAdd(counters, (unsigned long long)count * samples);
}

// After the sum-reduce of a group: every rank had written alpha = 1, so the root holds the group
// size there (shader.wgsl:1371 wants 1).
__global__ void set_alpha_kernel(float4* accum, size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) accum[i].w = 1.0f;
}

__global__ void clear_accum_kernel(float4* accum, size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) accum[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Present pass: src/shaders/hdr.wgsl:3-22 over out_texture = accum / sample_total (f16).
__global__ void tonemap_kernel(const float4* accum, uint8_t* rgb8, size_t count, float inv_unused,
                               float sample_total) {
    (void)inv_unused;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float4 a = accum[i];
    // pixel_color = total_light / f32(sample_count + 1), stored to an rgba16float texture.
    float h[3] = {__half2float(__float2half_rn(a.x / sample_total)),
                  __half2float(__float2half_rn(a.y / sample_total)),
                  __half2float(__float2half_rn(a.z / sample_total))};
    float out[3];
    if (h[0] < 0.0f || h[1] < 0.0f || h[2] < 0.0f) {
        out[0] = 1.0f; out[1] = 0.0f; out[2] = 1.0f;
    } else {
        // WGSL mat3x3 constructors take columns.
        float v[3] = {(0.59719f * h[0] + 0.35458f * h[1]) + 0.04823f * h[2],
                      (0.07600f * h[0] + 0.90834f * h[1]) + 0.01566f * h[2],
                      (0.02840f * h[0] + 0.13383f * h[1]) + 0.83777f * h[2]};
        float r[3];
        for (int k = 0; k < 3; k++) {
            float av = v[k] * (v[k] + 0.0245786f) - 0.000090537f;
            float bv = v[k] * (0.983729f * v[k] + 0.4329510f) + 0.238081f;
            r[k] = av / bv;
        }
        out[0] = (1.60475f * r[0] + -0.53108f * r[1]) + -0.07367f * r[2];
        out[1] = (-0.10208f * r[0] + 1.10813f * r[1]) + -0.00605f * r[2];
        out[2] = (-0.00327f * r[0] + -0.07276f * r[1]) + 1.07602f * r[2];
        for (int k = 0; k < 3; k++) out[k] = fminf(fmaxf(out[k], 0.0f), 1.0f);
    }
    for (int k = 0; k < 3; k++) {
        // linear -> sRGB, as the sRGB surface format does on store.
        float c = out[k];
        float s = c <= 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
        rgb8[i * 3 + k] = (uint8_t)(fminf(fmaxf(s, 0.0f), 1.0f) * 255.0f + 0.5f);
    }
}

// dev_index views of `main` (shader.wgsl:1314-1338): they write out_texture (rgba16float, cleared
// to zero every frame, state.rs:797-800) and return before any ray is cast.
//   2: every invocation draws 20 texels from the environment's alias table and adds 0.1 / 20 to
//      out_texture at the texel's coordinates. The shader's load + store pairs of different
//      invocations race; here every draw is applied (what a race-free execution gives): count the
//      draws per texel, then replay that many f16-rounded additions.
//   3: out_texture = saturate(environment texel at the pixel's own coordinates).
__global__ void view_distribution_count_kernel(EnvView env, uint32_t width, uint32_t height, uint32_t sample,
                                               uint32_t* counts) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= width || y >= height) return;
    uint32_t rng = rng_seed(y * width + x, sample);
    const uint32_t length = env.width * env.height;
    for (int i = 0; i < 20; i++) {
        // random_index_in_environment (shader.wgsl:689-706)
        uint32_t index = min((uint32_t)(rng_uniform(rng) * (float)length), length - 1u);
        float4 entry = env.alias[index];
        uint32_t pick = rng_uniform(rng) < entry.x ? index : f_as_u(entry.y);
        uint32_t ex = pick % env.width, ey = pick / env.width;
        if (ex < width && ey < height) atomicAdd(&counts[ey * width + ex], 1u);  // stores outside the texture are dropped
    }
}

__global__ void view_distribution_resolve_kernel(const uint32_t* counts, float4* out, size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float step = 0.1f / 20.0f;  // vec3(0.1) / f32(count)
    float c = 0.0f;
    for (uint32_t k = counts[i]; k > 0u; k--) {
        float next = __half2float(__float2half_rn(c + step));
        if (next == c) break;  // the f16 texel has stopped growing
        c = next;
    }
    out[i] = make_float4(c, c, c, 0.0f);
}

__global__ void view_environment_kernel(EnvView env, uint32_t width, uint32_t height, float4* out) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= width || y >= height) return;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);  // textureLoad outside the texture: zero
    if (x < env.width && y < env.height) t = env.texels[(size_t)y * env.width + x];
    out[(size_t)y * width + x] = make_float4(__half2float(__float2half_rn(saturate(t.x))),
                                             __half2float(__float2half_rn(saturate(t.y))),
                                             __half2float(__float2half_rn(saturate(t.z))), 0.0f);
}

// =====================================================================================
// Host side
// =====================================================================================

namespace {

std::mutex g_error_mutex;
std::string g_global_error;

}  // namespace

struct RtContext {
    int device = 0;
    // result of the last rt_bvh_build, until rt_bvh_build_fetch
    RtBvhNode* d_built_nodes = nullptr;
    RtPrimitiveInfo* d_built_prims = nullptr;
    uint32_t built_node_count = 0, built_prim_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    std::string error;

    // scene: every array below lives in one of two arenas. An upload fills the arena that is not in
    // use and the two swap on success, so re-uploading a scene of the same size allocates nothing
    // (no kernel-mode driver call on the per-frame path) and a failed upload leaves the old scene.
    cudaEvent_t first_hit_events[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // per band
    void* scene_arena[2] = {nullptr, nullptr};
    size_t scene_arena_bytes[2] = {0, 0};
    int scene_arena_cur = 0;
    float4* d_nodes = nullptr;
    float4* d_records = nullptr;
    float4* d_tri_normals = nullptr;
    float4* d_materials = nullptr;
    SceneView scene{};
    // fast layout (rt_fast.cuh)
    float4* d_n2 = nullptr;
    float4* d_tri = nullptr;
    float4* d_ana = nullptr;
    float4* d_pre = nullptr;
    float4* d_flat = nullptr;
    uint32_t* d_big_len = nullptr;
    uint32_t* d_ana_paths = nullptr;
    rtf::FastScene fast{};
    uint32_t bvh_depth = 0;
    uint32_t node_count = 0, record_count = 0, material_count = 0;
    bool has_scene = false;

    // environments
    struct Env {
        cudaArray_t array = nullptr;
        cudaTextureObject_t texture = 0;
        float4* d_alias = nullptr;   // one allocation: alias table, then the texels in linear memory
        float4* d_texels = nullptr;
        uint32_t width = 0, height = 0;
    };
    std::vector<Env> envs;

    CameraParams camera{};
    bool has_camera = false;

    uint32_t width = 0, height = 0;
    float4* d_accum_own = nullptr;
    float4* d_accum = nullptr;

    unsigned long long* d_counters = nullptr;
    int sm_count = 0;
    enum Kernel { KERNEL_LOCAL, KERNEL_WAVE, KERNEL_FAST, KERNEL_PIXEL };
    Kernel kernel = KERNEL_WAVE;
    // persistent CTA-local wavefront (rt_local.cuh)
    void* local_arena = nullptr;
    size_t local_arena_bytes = 0;
    uint32_t local_slots = 1024;          // path slots per CTA
    uint32_t local_window = 16;           // samples per window (power of two)
    uint32_t local_stage_paths = 8192;    // parked samples per CTA (ring of windows)
    bool local_flat_allowed = false;      // measured: the stack walk is faster (profiles/README.md, r2)
    size_t local_smem_set[2] = {0, 0};    // dynamic shared memory opted in per kernel form
    unsigned long long* d_local_profile = nullptr;  // 11 x u64 (LocalParams::profile)
    // wavefront pipeline state (rt_wave.cuh)
    static constexpr int kMaxGroups = 4;
    rtw::WaveBuffers wave[kMaxGroups]{};
    void* wave_arena = nullptr;
    size_t wave_arena_bytes = 0;
    uint32_t wave_slots = 0;          // per-group capacity the arena was allocated for
    int wave_arena_groups = 0;
    uint32_t* wave_next_path = nullptr;
    uint32_t wave_max_slots = 8192u * 1024u;  // both groups together
    int wave_groups = 2;
    int wave_blocks1[4] = {RT_WAVE_SHADE_MIN_BLOCKS, RT_WAVE_FINISH_MIN_BLOCKS, RT_WAVE_PREPASS_MIN_BLOCKS,
                           RT_WAVE_WALK_MIN_BLOCKS};  // blocks per SM of shade, finish, pre-pass, walk: one group
    int wave_blocks2[4] = {8, 8, 4, 5};               // ... two groups (measured best: full grids on both streams)
    // the wave loop as a CUDA graph: a WHILE node whose body is `wave_body_waves` waves of every group
    bool wave_graph_enabled = true;
    int wave_body_waves = 4;
    cudaGraph_t wave_graph = nullptr;
    cudaGraphExec_t wave_exec = nullptr;
    std::string wave_graph_key;       // bytes of everything the graph's nodes were built from
    std::string wave_graph_shape;     // what fixes the graph's topology and launch shapes
    struct WaveGraphNode {
        cudaGraphNode_t node;
        int group, kind, wave;        // kernel `kind` of slot group `group`, wave parity `wave`
    };
    std::vector<WaveGraphNode> wave_graph_nodes;  // the body's kernel nodes (not wave_continue_kernel)
    cudaGraphNode_t wave_graph_continue = nullptr;   // the body's last node: wave_continue_kernel
    cudaGraphConditionalHandle wave_graph_handle = 0;
    bool wave_graph_update_ok = true;  // cudaGraphExecKernelNodeSetParams works on body nodes (probed once)
    uint64_t wave_graph_builds = 0, wave_graph_updates = 0;
    unsigned long long* d_graph_launches = nullptr;  // kernels the graph bodies have run (device-side count)
    void* wave_staging = nullptr;     // per-sample radiance planes of the current batch
    size_t wave_staging_bytes = 0;
    size_t wave_staging_budget = (size_t)8192 << 20;  // upper bound; clamped to what the device has free
    uint32_t* wave_live_host = nullptr;  // pinned read-back of WaveCtrl::live: [2 chunks in flight][groups]
    cudaStream_t wave_streams[kMaxGroups] = {};
    cudaEvent_t wave_events[2][kMaxGroups] = {};
    cudaEvent_t wave_fork = nullptr, wave_phase[kMaxGroups] = {}, wave_join[kMaxGroups] = {};
    bool profiling = false;
    RtWaveProfile profile{};
    bool detailed = false;
    float first_hit_kernel_ms = 0.f;  // device time of the last rt_first_hit kernel
    void* d_first_hit = nullptr;      // slot map, then distance map, of rt_first_hit
    size_t first_hit_capacity = 0;    // pixels
    uint64_t launches = 0;
};

namespace {

int fail(RtContext* ctx, int status, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) {
        ctx->error = buf;
    } else {
        std::lock_guard<std::mutex> lock(g_error_mutex);
        g_global_error = buf;
    }
    return status;
}

#define RT_CUDA(ctx, expr)                                                                   \
    do {                                                                                     \
        cudaError_t e_ = (expr);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? RT_ERR_OUT_OF_MEMORY          \
                                                             : RT_ERR_CUDA,                  \
                        "%s failed: %s", #expr, cudaGetErrorString(e_));                     \
    } while (0)

void free_scene(RtContext* ctx) {
    for (int i = 0; i < 2; i++) {
        cudaFree(ctx->scene_arena[i]);
        ctx->scene_arena[i] = nullptr;
        ctx->scene_arena_bytes[i] = 0;
    }
    ctx->d_flat = nullptr;
    ctx->d_ana_paths = nullptr;
    ctx->d_n2 = ctx->d_tri = ctx->d_ana = ctx->d_pre = nullptr;
    ctx->d_big_len = nullptr;
    ctx->d_nodes = ctx->d_records = ctx->d_tri_normals = ctx->d_materials = nullptr;
    ctx->has_scene = false;
}

void free_envs(RtContext* ctx) {
    for (auto& e : ctx->envs) {
        if (e.texture) cudaDestroyTextureObject(e.texture);
        if (e.array) cudaFreeArray(e.array);
        cudaFree(e.d_alias);
    }
    ctx->envs.clear();
}

using rtf::bits;

// tree_depth (src/bvh.rs:124-134) of the flattened array, iteratively; also validates
// child / primitive ranges so a malformed upload cannot send a kernel out of bounds.
bool validate_bvh(const RtSceneDesc& s, uint32_t* depth_out, std::string* why) {
    struct Item {
        uint32_t node, depth;
    };
    std::vector<Item> todo;
    todo.push_back({0u, 0u});
    uint32_t max_depth = 0;
    uint64_t visited = 0;
    while (!todo.empty()) {
        Item it = todo.back();
        todo.pop_back();
        if (it.node >= s.bvh_node_count) {
            *why = "BVH child index out of range";
            return false;
        }
        if (++visited > (uint64_t)s.bvh_node_count) {
            *why = "BVH is not a tree";
            return false;
        }
        const RtBvhNode& n = s.bvh_nodes[it.node];
        if (it.depth > max_depth) max_depth = it.depth;
        if (n.primitives_len > 0) {
            if ((uint64_t)n.primitives_or_second_child_index + n.primitives_len > s.primitive_count) {
                *why = "BVH leaf primitive range out of range";
                return false;
            }
            if (n.primitives_len > 0x1FFFFFFFu) {
                *why = "BVH leaf too large";
                return false;
            }
        } else {
            if (n.split_axis > 2u) {
                *why = "BVH split axis out of range";
                return false;
            }
            todo.push_back({it.node + 1u, it.depth + 1u});
            todo.push_back({n.primitives_or_second_child_index, it.depth + 1u});
        }
    }
    *depth_out = max_depth;
    return true;
}

// ---- wavefront pipeline (rt_wave.cuh): buffers and the wave loop ----
int wave_reserve(RtContext* ctx, uint32_t slots, int groups, size_t staging_bytes) {
    if (!(ctx->wave_arena && ctx->wave_slots >= slots && ctx->wave_arena_groups == groups)) {
        RT_CUDA(ctx, cudaDeviceSynchronize());
        cudaFree(ctx->wave_arena);
        ctx->wave_arena = nullptr;
        ctx->wave_slots = 0;
        const size_t S = slots;
        // per group: 3 x 32-byte pair arrays + 2 x 16 + fb (8) + rayq (4 x 4) + occl (1) + stage (1) + ctrl
        const size_t per_group = (S * (3 * 32 + 2 * 16 + 8 + 16 + 2) + 1024 + 255) & ~(size_t)255;
        ctx->wave_arena_bytes = 0;
        RT_CUDA(ctx, cudaMalloc(&ctx->wave_arena, per_group * (size_t)groups + 256));
        ctx->wave_arena_bytes = per_group * (size_t)groups + 256;
        for (int g = 0; g < groups; g++) {
            char* q = (char*)ctx->wave_arena + per_group * (size_t)g;
            rtw::WaveBuffers& b = ctx->wave[g];
            b.rOD = (float4*)q;    q += S * 32;
            b.pTLPI = (float4*)q;  q += S * 64;
            b.sray_d = (float4*)q; q += S * 16;
            b.hit = (float4*)q;    q += S * 16;
            b.fb = (float2*)q;     q += S * 8;
            b.rayq = (uint32_t*)q; q += S * 16;
            b.occl = (uint8_t*)q;  q += S;
            b.stage = (uint8_t*)q; q += S;
            q = (char*)(((uintptr_t)q + 63) & ~(uintptr_t)63);
            b.ctrl = (rtw::WaveCtrl*)q;
        }
        ctx->wave_next_path = (uint32_t*)((char*)ctx->wave_arena + per_group * (size_t)groups);
        ctx->wave_slots = slots;
        ctx->wave_arena_groups = groups;
    }
    if (!(ctx->wave_staging && ctx->wave_staging_bytes >= staging_bytes)) {
        RT_CUDA(ctx, cudaDeviceSynchronize());
        cudaFree(ctx->wave_staging);
        ctx->wave_staging = nullptr;
        ctx->wave_staging_bytes = 0;
        RT_CUDA(ctx, cudaMalloc(&ctx->wave_staging, staging_bytes));
        ctx->wave_staging_bytes = staging_bytes;
    }
    return RT_OK;
}

// The wave loop of one batch as a CUDA graph: a WHILE node whose body runs `body_waves` waves of every
// slot group — one chain of kernels per group, the chains side by side, group g rotated g * lead
// kernels into its cycle (its first g * lead kernels ran before the loop) so that memory-bound and
// issue-bound kernels of different groups overlap — and ends with wave_continue_kernel, which keeps
// the loop going while any slot is live. One graph launch replaces ~1600 kernel launches and every
// host synchronisation of the loop: rt_render returns at once, and the loop stops within one body
// of the last path instead of running up to two 8-wave chunks empty.
// Upper bound on the bodies one batch can run: a slot ends a path within max_bounces + 3 waves and takes
// the next one at once, so after (paths per slot + 2) such rounds nothing can be live; twice that, plus
// slack, only ever stops a loop that would not have ended.
uint32_t wave_max_bodies(const rtw::WaveParams* p, int groups, int body_waves) {
    const uint64_t slots = (uint64_t)p[0].slots * (uint64_t)groups;
    const uint64_t rounds = (uint64_t)p[0].path_total / (slots ? slots : 1u) + 2u;
    const uint64_t waves = 2u * rounds * ((uint64_t)p[0].max_bounces + 3u) + 64u;
    const uint64_t bodies = waves / (uint64_t)body_waves + 2u;
    if (const char* k = getenv("RT_WAVE_MAX_BODIES"))  // test hook: provoke the cut-off
        if (atol(k) > 0) return (uint32_t)atol(k);
    return bodies > 0x7FFFFFFFull ? 0x7FFFFFFFu : (uint32_t)bodies;
}

int wave_graph_build(RtContext* ctx, const rtw::WaveParams* p, int groups, const unsigned* blocks, bool shallow) {
    const int body_waves = ctx->wave_body_waves;
    std::string key;
    key.append((const char*)p, sizeof(rtw::WaveParams) * (size_t)groups);
    key.append((const char*)blocks, sizeof(unsigned) * 4);
    key.push_back((char)groups);
    key.push_back((char)body_waves);
    key.push_back(shallow ? 1 : 0);
    const uint32_t max_bodies = wave_max_bodies(p, groups, body_waves);
    key.append((const char*)&max_bodies, sizeof max_bodies);
    if (ctx->wave_exec && key == ctx->wave_graph_key) return RT_OK;
    std::string shape;
    shape.append((const char*)blocks, sizeof(unsigned) * 4);
    shape.push_back((char)groups);
    shape.push_back((char)body_waves);
    shape.push_back(shallow ? 1 : 0);
    void* const funcs[4] = {(void*)rtw::wave_shade_kernel, (void*)rtw::wave_finish_kernel, (void*)rtw::wave_prepass_kernel,
                            shallow ? (void*)rtw::wave_walk_kernel<true> : (void*)rtw::wave_walk_kernel<false>};
    const unsigned threads[4] = {(unsigned)rtw::kShadeThreads, (unsigned)rtw::kFinishThreads, (unsigned)rtw::kPrepassThreads,
                                 (unsigned)rtw::kWalkThreads};
    if (ctx->wave_exec && ctx->wave_graph_update_ok && shape == ctx->wave_graph_shape) {
        // Same topology, new arguments (another scene upload, camera, sample range, buffer): the
        // instantiated graph keeps its nodes and only their kernel parameters change — a new frame of
        // an interactive caller costs 64 parameter updates, not a graph instantiation.
        cudaError_t ue = cudaSuccess;
        for (const RtContext::WaveGraphNode& n : ctx->wave_graph_nodes) {
            rtw::WaveParams arg = p[n.group];
            arg.wave = (uint32_t)n.wave;
            void* args[1] = {&arg};
            cudaKernelNodeParams kp;
            memset(&kp, 0, sizeof kp);
            kp.func = funcs[n.kind];
            kp.gridDim = dim3(blocks[n.kind]);
            kp.blockDim = dim3(threads[n.kind]);
            kp.kernelParams = args;
            ue = cudaGraphExecKernelNodeSetParams(ctx->wave_exec, n.node, &kp);
            if (ue != cudaSuccess) break;
        }
        if (ue == cudaSuccess) {
            // the loop condition reads the groups' control blocks, which move with the slot arena
            rtw::WaveLoopCtrl lc;
            memset(&lc, 0, sizeof lc);
            for (int g = 0; g < groups; g++) lc.group[g] = p[g].b.ctrl;
            lc.groups = (uint32_t)groups;
            lc.body_kernels = (uint32_t)(groups * 4 * body_waves + 1);
            lc.launches = ctx->d_graph_launches;
            lc.max_bodies = max_bodies;
            void* args[2] = {&ctx->wave_graph_handle, &lc};
            cudaKernelNodeParams kp;
            memset(&kp, 0, sizeof kp);
            kp.func = (void*)rtw::wave_continue_kernel;
            kp.gridDim = dim3(1);
            kp.blockDim = dim3(1);
            kp.kernelParams = args;
            ue = cudaGraphExecKernelNodeSetParams(ctx->wave_exec, ctx->wave_graph_continue, &kp);
        }
        if (ue == cudaSuccess) {
            ctx->wave_graph_key = key;
            ctx->wave_graph_updates++;
            return RT_OK;
        }
        cudaGetLastError();                 // not supported for nodes of a conditional body here:
        ctx->wave_graph_update_ok = false;  // rebuild from now on (some nodes may hold new arguments)
    }
    ctx->wave_graph_builds++;
    ctx->wave_graph_nodes.clear();
    if (ctx->wave_exec) cudaGraphExecDestroy(ctx->wave_exec);
    if (ctx->wave_graph) cudaGraphDestroy(ctx->wave_graph);
    ctx->wave_exec = nullptr;
    ctx->wave_graph = nullptr;
    ctx->wave_graph_key.clear();
    if (!ctx->d_graph_launches) {
        RT_CUDA(ctx, cudaMalloc(&ctx->d_graph_launches, 3 * sizeof(unsigned long long)));
        RT_CUDA(ctx, cudaMemset(ctx->d_graph_launches, 0, 3 * sizeof(unsigned long long)));
    }
    cudaGraph_t graph = nullptr;
    RT_CUDA(ctx, cudaGraphCreate(&graph, 0));
    auto bail = [&](const char* what, cudaError_t e) {
        cudaGraphDestroy(graph);
        return fail(ctx, RT_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
    };
    cudaGraphConditionalHandle handle;
    cudaError_t e = cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault);
    if (e != cudaSuccess) return bail("cudaGraphConditionalHandleCreate", e);
    cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = handle;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t loop;
    e = cudaGraphAddNode(&loop, graph, nullptr, 0, &np);
    if (e != cudaSuccess) return bail("cudaGraphAddNode(while)", e);
    cudaGraph_t body = np.conditional.phGraph_out[0];

    const int lead = 4 / groups > 0 ? 4 / groups : 1;
    std::vector<cudaGraphNode_t> tails;
    for (int g = 0; g < groups; g++) {
        cudaGraphNode_t prev = nullptr;
        for (int i = 0; i < 4 * body_waves; i++) {
            const int j = i + g * lead;  // position in the group's own cycle
            rtw::WaveParams arg = p[g];
            arg.wave = (uint32_t)(j / 4) & 1u;
            void* args[1] = {&arg};
            cudaKernelNodeParams kp;
            memset(&kp, 0, sizeof kp);
            kp.func = funcs[j % 4];
            kp.gridDim = dim3(blocks[j % 4]);
            kp.blockDim = dim3(threads[j % 4]);
            kp.kernelParams = args;
            cudaGraphNode_t node;
            e = cudaGraphAddKernelNode(&node, body, prev ? &prev : nullptr, prev ? 1 : 0, &kp);
            if (e != cudaSuccess) return bail("cudaGraphAddKernelNode", e);
            ctx->wave_graph_nodes.push_back({node, g, j % 4, (j / 4) & 1});
            prev = node;
        }
        tails.push_back(prev);
    }
    {
        rtw::WaveLoopCtrl lc;
        memset(&lc, 0, sizeof lc);
        for (int g = 0; g < groups; g++) lc.group[g] = p[g].b.ctrl;
        lc.groups = (uint32_t)groups;
        lc.body_kernels = (uint32_t)(groups * 4 * body_waves + 1);
        lc.launches = ctx->d_graph_launches;
        lc.max_bodies = max_bodies;
        void* args[2] = {&handle, &lc};
        cudaKernelNodeParams kp;
        memset(&kp, 0, sizeof kp);
        kp.func = (void*)rtw::wave_continue_kernel;
        kp.gridDim = dim3(1);
        kp.blockDim = dim3(1);
        kp.kernelParams = args;
        cudaGraphNode_t node;
        e = cudaGraphAddKernelNode(&node, body, tails.data(), tails.size(), &kp);
        if (e != cudaSuccess) return bail("cudaGraphAddKernelNode(continue)", e);
        ctx->wave_graph_continue = node;
        ctx->wave_graph_handle = handle;
    }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    if (e != cudaSuccess) return bail("cudaGraphInstantiate", e);
    ctx->wave_graph = graph;
    ctx->wave_exec = exec;
    ctx->wave_graph_key = key;
    ctx->wave_graph_shape = shape;
    return RT_OK;
}

// Renders the sample range in batches. Inside a batch the paths are independent work items
// handed to the slots as they free up; waves run until every slot has retired, then the
// staged samples are added to the accumulation buffer in sample order. The slots form two
// groups on two streams, started half a wave apart. The host only paces the launches: it
// reads the live-slot counts one chunk behind, so the streams never run dry.
int wave_render(RtContext* ctx, const RtContext::Env& e, uint32_t sample_begin, uint32_t sample_count,
                uint32_t max_bounces) {
    rtw::WaveParams base;
    memset(&base, 0, sizeof base);
    base.scene = ctx->fast;
    base.env.texture = e.texture;
    base.env.alias = e.d_alias;
    base.env.texels = e.d_texels;
    base.env.width = e.width;
    base.env.height = e.height;
    base.camera = ctx->camera;
    base.accum = ctx->d_accum;
    base.counters = ctx->d_counters;
    base.width = ctx->width;
    base.height = ctx->height;
    base.tiles_x = (ctx->width + 7u) / 8u;
    base.pixel_ids = base.tiles_x * ((ctx->height + 7u) / 8u) * 64u;
    base.max_bounces = max_bounces;

    // Batch size: as many samples as the staging budget holds, and path indices must fit 32 bits.
    const size_t pixels = (size_t)ctx->width * ctx->height;
    // The staging planes must fit what the device has left (the budget is only an upper bound).
    size_t budget = ctx->wave_staging_budget;
    // Asked only when the planes have to grow: cudaMemGetInfo goes through the kernel-mode driver,
    // and on a host whose other GPUs are busy that call was measured to block for 2 - 75 ms now and
    // then (tools/variance_probe.py) — inside what would otherwise be a purely user-mode enqueue.
    {
        uint64_t want = sample_count < 1u ? 1u : sample_count;
        if (want * pixels * sizeof(float4) > budget) want = budget / (pixels * sizeof(float4));
        if (want < 1u) want = 1u;
        if (want * pixels * sizeof(float4) <= ctx->wave_staging_bytes) {
            budget = ctx->wave_staging_bytes;  // what is allocated already holds the batch
        } else {
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
                const size_t usable = free_b + ctx->wave_staging_bytes;  // our own planes are reusable
                const size_t reserve = (size_t)2048 << 20;               // slot arena + headroom for the caller
                const size_t cap = usable > reserve ? (usable - reserve) / 4 * 3 : usable / 2;
                if (budget > cap) budget = cap;
            }
        }
    }
    uint64_t batch = budget / (pixels * sizeof(float4));
    // Every slot that retires after the paths ran out still bumps next_path once per wave it
    // waits: leave that much headroom below 2^32.
    const uint64_t index_cap = (0xFFFFFFFFull - 64ull * ctx->wave_max_slots) / base.pixel_ids;
    if (batch > index_cap) batch = index_cap;
    if (batch > sample_count) batch = sample_count;
    if (batch < 1) batch = 1;
    const uint64_t paths_per_batch = (uint64_t)base.pixel_ids * batch;
    // A slot should carry two to three paths or more over the render: with fewer, most of the render
    // is the pipeline filling and draining. Measured at 1080p with the full pool / this rule: 1 spp
    // 1.66 / 1.59 ms, 2 spp 2.81 / 2.72, 4 spp 5.11 / 4.81, 8 spp 9.31 / 9.08; 640x360 x 64 spp 5.97 /
    // 5.86. Large renders keep the full pool.
    uint64_t want_slots = paths_per_batch <= 12000000ull ? paths_per_batch / 2u : paths_per_batch / 3u;
    want_slots = (want_slots + 65535u) & ~(uint64_t)65535u;
    if (want_slots < 262144u) want_slots = paths_per_batch < 262144u ? paths_per_batch : 262144u;
    const uint32_t total_slots = (uint32_t)(want_slots < ctx->wave_max_slots ? want_slots : ctx->wave_max_slots);
    const int groups = total_slots >= 262144u ? ctx->wave_groups : 1;  // small frames: one group
    const uint32_t per_group = (total_slots + (uint32_t)groups - 1u) / (uint32_t)groups;
    int st = wave_reserve(ctx, per_group, groups, pixels * sizeof(float4) * (size_t)batch);
    if (st != RT_OK) return st;

    rtw::WaveParams p[RtContext::kMaxGroups];
    for (int g = 0; g < groups; g++) {
        p[g] = base;
        p[g].b = ctx->wave[g];
        p[g].slots = per_group;
        p[g].staging = (float4*)ctx->wave_staging;
        p[g].next_path = ctx->wave_next_path;
    }
    auto blocks_for = [&](uint32_t items, int threads, int per_sm) {
        unsigned need = (items + (unsigned)threads - 1u) / (unsigned)threads;
        unsigned cap = (unsigned)ctx->sm_count * (unsigned)per_sm;
        return need < cap ? (need ? need : 1u) : cap;
    };
    // Blocks per SM are tunable per group count (RT_WAVE_BLOCKS2); full-size grids on both
    // streams measured best, the block scheduler interleaves the two kernels as SMs free up.
    const int* per_sm = groups >= 2 ? ctx->wave_blocks2 : ctx->wave_blocks1;
    const unsigned shade_blocks = blocks_for(per_group, rtw::kShadeThreads, per_sm[0]);
    const unsigned finish_blocks = blocks_for(per_group, rtw::kFinishThreads, per_sm[1]);
    const unsigned prepass_blocks = blocks_for(2u * per_group, rtw::kPrepassThreads, per_sm[2]);
    const unsigned walk_blocks = blocks_for(2u * per_group, rtw::kWalkThreads, per_sm[3]);

    // RT_WAVE_TIMELINE=1 (diagnostics) or rt_set_profiling: CUDA events around every kernel.
    static const bool timeline_print = getenv("RT_WAVE_TIMELINE") != nullptr;
    const bool timeline = timeline_print || ctx->profiling;
    std::vector<cudaEvent_t> tl[RtContext::kMaxGroups];
    auto mark = [&](int g, cudaStream_t s) {
        if (!timeline) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, s);
        tl[g].push_back(ev);
    };
    // Shallow trees walk with their stack in shared memory (RT_WALK_SHALLOW=0: never).
    const bool shallow_allowed = !(getenv("RT_WALK_SHALLOW") && atoi(getenv("RT_WALK_SHALLOW")) == 0);
    const bool shallow = shallow_allowed && ctx->bvh_depth <= rtw::kShallowDepth;
    // One wave of group g: kernels first .. last-1 of (shade, finish, pre-pass, walk).
    auto launch = [&](int g, uint32_t wave, int first, int last) {
        cudaStream_t s = groups >= 2 ? ctx->wave_streams[g] : ctx->stream;
        p[g].wave = wave;
        for (int k = first; k < last; k++) {
            mark(g, s);
            if (k == 0) rtw::wave_shade_kernel<<<shade_blocks, rtw::kShadeThreads, 0, s>>>(p[g]);
            else if (k == 1) rtw::wave_finish_kernel<<<finish_blocks, rtw::kFinishThreads, 0, s>>>(p[g]);
            else if (k == 2) rtw::wave_prepass_kernel<<<prepass_blocks, rtw::kPrepassThreads, 0, s>>>(p[g]);
            else if (shallow) rtw::wave_walk_kernel<true><<<walk_blocks, rtw::kWalkThreads, 0, s>>>(p[g]);
            else rtw::wave_walk_kernel<false><<<walk_blocks, rtw::kWalkThreads, 0, s>>>(p[g]);
            mark(g, s);
            ctx->launches++;
        }
    };

    // The loop runs as one CUDA graph unless per-kernel timing was asked for (events around every
    // launch need the host-paced loop) or RT_WAVE_GRAPH=0.
    const bool use_graph = ctx->wave_graph_enabled && !timeline && groups <= 4;
    const unsigned blocks4[4] = {shade_blocks, finish_blocks, prepass_blocks, walk_blocks};

    int status = RT_OK;
    cudaEvent_t t_begin = nullptr, t_end = nullptr;
    if (timeline) {
        cudaEventCreate(&t_begin);
        cudaEventCreate(&t_end);
        cudaEventRecord(t_begin, ctx->stream);
    }
    for (uint32_t done = 0; done < sample_count && status == RT_OK; done += (uint32_t)batch) {
        for (int g = 0; g < groups; g++) {
            p[g].batch_begin = sample_begin + done;
            p[g].batch_count = (uint32_t)(sample_count - done < batch ? sample_count - done : batch);
            p[g].path_total = base.pixel_ids * p[g].batch_count;
        }
        cudaMemsetAsync(ctx->wave_next_path, 0, sizeof(uint32_t), ctx->stream);
        if (use_graph) {
            // everything on the context's stream: init, the head start of groups 1.., the loop, the sum
            const int lead = 4 / groups > 0 ? 4 / groups : 1;
            for (int g = 0; g < groups; g++) {
                rtw::wave_init_kernel<<<(per_group + 255) / 256, 256, 0, ctx->stream>>>(p[g]);
                ctx->launches++;
            }
            for (int g = 1; g < groups; g++) {
                p[g].wave = 0u;
                for (int k = 0; k < g * lead && k < 4; k++) {
                    if (k == 0) rtw::wave_shade_kernel<<<shade_blocks, rtw::kShadeThreads, 0, ctx->stream>>>(p[g]);
                    else if (k == 1) rtw::wave_finish_kernel<<<finish_blocks, rtw::kFinishThreads, 0, ctx->stream>>>(p[g]);
                    else if (k == 2) rtw::wave_prepass_kernel<<<prepass_blocks, rtw::kPrepassThreads, 0, ctx->stream>>>(p[g]);
                    ctx->launches++;
                }
            }
            cudaError_t ce = cudaGetLastError();
            if (ce != cudaSuccess) {
                status = fail(ctx, RT_ERR_CUDA, "wave launch failed: %s", cudaGetErrorString(ce));
                break;
            }
            status = wave_graph_build(ctx, p, groups, blocks4, shallow);
            if (status != RT_OK) break;
            // A caller's stream may be the legacy default stream, on which the graph's two chains
            // lose their overlap (measured: 4 % slower): the loop then runs on one of the context's
            // own non-blocking streams, ordered into the caller's stream with a pair of events.
            cudaStream_t loop_stream = ctx->stream;
            if (ctx->stream != ctx->own_stream) {
                loop_stream = ctx->wave_streams[0];
                ce = cudaEventRecord(ctx->wave_fork, ctx->stream);
                if (ce == cudaSuccess) ce = cudaStreamWaitEvent(loop_stream, ctx->wave_fork, 0);
            }
            if (ce == cudaSuccess) ce = cudaMemsetAsync(ctx->d_graph_launches + 1, 0, sizeof(unsigned long long), loop_stream);
            if (ce == cudaSuccess) ce = cudaGraphLaunch(ctx->wave_exec, loop_stream);
            if (ce == cudaSuccess && loop_stream != ctx->stream) {
                ce = cudaEventRecord(ctx->wave_join[0], loop_stream);
                if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->stream, ctx->wave_join[0], 0);
            }
            if (ce != cudaSuccess) {
                status = fail(ctx, RT_ERR_CUDA, "cudaGraphLaunch failed: %s", cudaGetErrorString(ce));
                break;
            }
            rtw::wave_accumulate_kernel<<<(unsigned)((pixels + 255) / 256), 256, 0, ctx->stream>>>(p[0]);
            ctx->launches++;
            ce = cudaGetLastError();
            if (ce != cudaSuccess) status = fail(ctx, RT_ERR_CUDA, "accumulate launch failed: %s", cudaGetErrorString(ce));
            continue;
        }
        if (groups >= 2) {
            // fork: the group streams start after everything queued on the context's stream
            cudaEventRecord(ctx->wave_fork, ctx->stream);
            for (int g = 0; g < groups; g++) cudaStreamWaitEvent(ctx->wave_streams[g], ctx->wave_fork, 0);
        }
        for (int g = 0; g < groups; g++) {
            cudaStream_t s = groups >= 2 ? ctx->wave_streams[g] : ctx->stream;
            rtw::wave_init_kernel<<<(per_group + 255) / 256, 256, 0, s>>>(p[g]);
            ctx->launches++;
        }

        const uint32_t chunk = 8;
        uint32_t wave = 0;
        for (uint32_t k = 0;; k++) {
            for (uint32_t i = 0; i < chunk; i++, wave++) {
                if (groups == 1 || wave > 0) {
                    for (int g = 0; g < groups; g++) launch(g, wave, 0, 4);
                } else {
                    // Stagger the groups: group g starts its first wave when group g - 1 is 4 / groups
                    // kernels into its own, so memory-bound and issue-bound kernels of different
                    // groups run side by side from then on.
                    const int lead = 4 / groups > 0 ? 4 / groups : 1;
                    for (int g = 0; g < groups; g++) {
                        if (g > 0) cudaStreamWaitEvent(ctx->wave_streams[g], ctx->wave_phase[g - 1], 0);
                        launch(g, 0, 0, lead);
                        cudaEventRecord(ctx->wave_phase[g], ctx->wave_streams[g]);
                        launch(g, 0, lead, 4);
                    }
                }
            }
            cudaError_t ce = cudaGetLastError();
            for (int g = 0; g < groups && ce == cudaSuccess; g++) {
                cudaStream_t s = groups >= 2 ? ctx->wave_streams[g] : ctx->stream;
                ce = cudaMemcpyAsync(&ctx->wave_live_host[(k & 1u) * RtContext::kMaxGroups + g], &p[g].b.ctrl->live, sizeof(uint32_t),
                                     cudaMemcpyDeviceToHost, s);
                if (ce == cudaSuccess) ce = cudaEventRecord(ctx->wave_events[k & 1u][g], s);
            }
            if (ce == cudaSuccess && k > 0) {
                uint32_t live = 0;
                for (int g = 0; g < groups && ce == cudaSuccess; g++) {
                    ce = cudaEventSynchronize(ctx->wave_events[(k - 1u) & 1u][g]);
                    live += ctx->wave_live_host[((k - 1u) & 1u) * RtContext::kMaxGroups + g];
                }
                if (ce == cudaSuccess && live == 0u) break;  // the last chunk ran empty
            }
            if (ce != cudaSuccess) {
                status = fail(ctx, RT_ERR_CUDA, "wave launch failed: %s", cudaGetErrorString(ce));
                break;
            }
            if (wave > 0x7FFFFF00u) {
                status = fail(ctx, RT_ERR_STATE, "wave loop did not terminate");
                break;
            }
        }
        if (groups >= 2) {
            // join: the context's stream waits for every group
            for (int g = 0; g < groups; g++) {
                cudaEventRecord(ctx->wave_join[g], ctx->wave_streams[g]);
                cudaStreamWaitEvent(ctx->stream, ctx->wave_join[g], 0);
            }
        }
        if (status == RT_OK) {
            rtw::wave_accumulate_kernel<<<(unsigned)((pixels + 255) / 256), 256, 0, ctx->stream>>>(p[0]);
            ctx->launches++;
            cudaError_t ce = cudaGetLastError();
            if (ce != cudaSuccess) status = fail(ctx, RT_ERR_CUDA, "accumulate launch failed: %s", cudaGetErrorString(ce));
        }
    }

    if (timeline) {
        cudaEventRecord(t_end, ctx->stream);
        cudaEventSynchronize(t_end);
        double sum[4] = {0, 0, 0, 0};
        uint64_t count[4] = {0, 0, 0, 0};
        size_t waves = 0;
        for (int g = 0; g < groups; g++) {
            // events come in (before, after) pairs; kernel kind cycles shade, finish, pre-pass, walk
            // (group 1 of a two-group render starts its cycle at shade as well)
            const size_t n = tl[g].size() / 2;
            for (size_t i = 0; i < n; i++) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, tl[g][2 * i], tl[g][2 * i + 1]);
                const int kind = (int)(i % 4);
                sum[kind] += ms;
                count[kind]++;
            }
            if (g == 0) waves = n / 4;
        }
        float total = 0.f;
        cudaEventElapsedTime(&total, t_begin, t_end);
        ctx->profile.renders++;
        ctx->profile.waves += waves;
        double busy = 0;
        for (int j = 0; j < 4; j++) {
            ctx->profile.launches[j] += count[j];
            ctx->profile.kernel_ms[j] += sum[j];
            busy += sum[j];
        }
        ctx->profile.gap_ms += groups == 1 ? total - busy : 0.0;
        ctx->profile.total_ms += total;
        if (timeline_print) {
            fprintf(stderr, "[wave timeline] groups=%d slots=%u x %d waves=%zu total=%.3f ms | shade %.3f finish %.3f prepass %.3f walk %.3f (sum %.3f) ms\n",
                    groups, per_group, groups, waves, total, sum[0], sum[1], sum[2], sum[3], busy);
            const size_t n = tl[0].size() / 8;
            for (size_t w = 0; w < n; w += (n / 10 ? n / 10 : 1)) {
                float t[4];
                for (int j = 0; j < 4; j++) cudaEventElapsedTime(&t[j], tl[0][8 * w + 2 * j], tl[0][8 * w + 2 * j + 1]);
                fprintf(stderr, "   wave %4zu: shade %7.1f finish %7.1f prepass %7.1f walk %7.1f us\n", w, t[0] * 1e3,
                        t[1] * 1e3, t[2] * 1e3, t[3] * 1e3);
            }
        }
        for (int g = 0; g < groups; g++)
            for (cudaEvent_t ev : tl[g]) cudaEventDestroy(ev);
        cudaEventDestroy(t_begin);
        cudaEventDestroy(t_end);
    }
    return status;
}

// After a synchronisation: did a graph loop stop at its safety bound? (Never seen; a render that did
// would be incomplete, so the caller is told.)
int wave_loop_check(RtContext* ctx) {
    if (!ctx->d_graph_launches) return RT_OK;
    unsigned long long cut = 0;
    RT_CUDA(ctx, cudaMemcpy(&cut, ctx->d_graph_launches + 2, sizeof cut, cudaMemcpyDeviceToHost));
    if (!cut) return RT_OK;
    RT_CUDA(ctx, cudaMemset(ctx->d_graph_launches + 2, 0, sizeof cut));
    return fail(ctx, RT_ERR_STATE, "wave loop did not terminate within its bound; the render is incomplete");
}

// ---- persistent CTA-local wavefront (rt_local.cuh): one asynchronous launch per render ----
uint32_t floor_pow2(uint32_t v) {
    uint32_t r = 1u;
    while (r * 2u <= v) r *= 2u;
    return r;
}

int local_render(RtContext* ctx, const RtContext::Env& e, uint32_t sample_begin, uint32_t sample_count,
                 uint32_t max_bounces) {
    rtl::LocalParams p;
    memset(&p, 0, sizeof p);
    p.scene = ctx->fast;
    p.env.texture = e.texture;
    p.env.alias = e.d_alias;
    p.env.texels = e.d_texels;
    p.env.width = e.width;
    p.env.height = e.height;
    p.camera = ctx->camera;
    p.accum = ctx->d_accum;
    p.counters = ctx->d_counters;
    p.width = ctx->width;
    p.height = ctx->height;
    p.tiles_x = (ctx->width + 7u) / 8u;
    p.tiles_total = p.tiles_x * ((ctx->height + 7u) / 8u);
    p.sample_begin = sample_begin;
    p.sample_count = sample_count;
    p.max_bounces = max_bounces;

    const bool flat = ctx->local_flat_allowed && ctx->fast.flat_leaves > 0u;
    const uint32_t window = floor_pow2(sample_count < ctx->local_window ? sample_count : ctx->local_window);
    uint32_t shift = 0;
    while ((1u << shift) < window) shift++;
    p.window_shift = shift;
    uint32_t ring = floor_pow2(ctx->local_stage_paths / (64u * window) > 0u ? ctx->local_stage_paths / (64u * window) : 1u);
    if (ring > rtl::kRing) ring = rtl::kRing;
    p.ring_entries = ring;
    // few samples per tile: take several tiles per atomic so that one grab feeds the slots once
    const uint64_t per_tile = 64ull * sample_count;
    uint32_t grab = (uint32_t)((ctx->local_slots + per_tile - 1) / per_tile);
    p.tile_grab = grab < 1u ? 1u : (grab > 64u ? 64u : grab);

    // no more CTAs than tiles, no more slots than a CTA's share of the paths needs
    const uint32_t grid_full = (uint32_t)ctx->sm_count * (uint32_t)rtl::kMinBlocks;
    const uint32_t grid = p.tiles_total < grid_full ? p.tiles_total : grid_full;
    p.slots = ctx->local_slots;

    const size_t T = (size_t)grid_full * p.slots;
    const size_t stage_f4 = (size_t)grid_full * ((size_t)ring * window * 64u);
    // rOD 32 + pTLPI 64 + sray_d 16 + hit 16 + fb 8 + occl 1 per slot, staging, tile counter
    const size_t bytes = T * (32 + 64 + 16 + 16 + 8 + 1) + 256 + stage_f4 * sizeof(float4) + 256;
    if (!ctx->local_arena || ctx->local_arena_bytes < bytes) {
        RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->local_arena);
        ctx->local_arena = nullptr;
        ctx->local_arena_bytes = 0;
        RT_CUDA(ctx, cudaMalloc(&ctx->local_arena, bytes));
        ctx->local_arena_bytes = bytes;
    }
    {
        char* q = (char*)ctx->local_arena;
        p.pTLPI = (float4*)q;   q += T * 64;
        p.rOD = (float4*)q;     q += T * 32;
        p.sray_d = (float4*)q;  q += T * 16;
        p.hit = (float4*)q;     q += T * 16;
        p.staging = (float4*)q; q += stage_f4 * sizeof(float4);
        p.fb = (float2*)q;      q += T * 8;
        p.next_tile = (uint32_t*)q; q += 256;
        p.occl = (uint8_t*)q;
    }
    const size_t smem = rtl::smem_layout(p.scene, p.slots, flat).bytes;
    if (ctx->local_smem_set[flat ? 1 : 0] < smem) {
        cudaError_t ce = flat ? cudaFuncSetAttribute(rtl::local_wave_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                              : cudaFuncSetAttribute(rtl::local_wave_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess)
            return fail(ctx, RT_ERR_CUDA, "%zu bytes of shared memory per CTA refused: %s", smem, cudaGetErrorString(ce));
        ctx->local_smem_set[flat ? 1 : 0] = smem;
    }
    RT_CUDA(ctx, cudaMemsetAsync(p.next_tile, 0, sizeof(uint32_t), ctx->stream));
    static const bool timeline_print = getenv("RT_LOCAL_PROFILE") != nullptr;
    const bool profiling = ctx->profiling || timeline_print;
    if (profiling) {
        if (!ctx->d_local_profile) RT_CUDA(ctx, cudaMalloc(&ctx->d_local_profile, 11 * sizeof(unsigned long long)));
        RT_CUDA(ctx, cudaMemsetAsync(ctx->d_local_profile, 0, 11 * sizeof(unsigned long long), ctx->stream));
        p.profile = ctx->d_local_profile;
        RT_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
    }
    if (flat) rtl::local_wave_kernel<true><<<grid, rtl::kThreads, smem, ctx->stream>>>(p);
    else rtl::local_wave_kernel<false><<<grid, rtl::kThreads, smem, ctx->stream>>>(p);
    ctx->launches++;
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return fail(ctx, RT_ERR_CUDA, "render kernel launch failed: %s", cudaGetErrorString(ce));
    if (profiling) {
        // Phase shares of the one launch (clock64 per CTA, thread 0, at the barriers) scaled to its
        // CUDA-event duration. Profiling makes the call synchronous.
        unsigned long long h[11];
        float ms = 0.f;
        RT_CUDA(ctx, cudaEventRecord(ctx->ev_stop, ctx->stream));
        RT_CUDA(ctx, cudaMemcpyAsync(h, ctx->d_local_profile, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        RT_CUDA(ctx, cudaEventSynchronize(ctx->ev_stop));
        RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        RT_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
        double cyc = 0;
        for (int i = 0; i < 5; i++) cyc += (double)h[i];
        if (cyc <= 0) cyc = 1;
        const double ctas = h[10] ? (double)h[10] : 1.0;
        // kernel order of RtWaveProfile: shade, finish, pre-pass, walk; bookkeeping goes to gap_ms
        const int order[4] = {0, 2, 3, 4};
        ctx->profile.renders++;
        ctx->profile.waves += (uint64_t)((double)h[9] / ctas + 0.5);
        for (int j = 0; j < 4; j++) {
            ctx->profile.launches[j] += 1;
            ctx->profile.kernel_ms[j] += ms * (double)h[order[j]] / cyc;
        }
        ctx->profile.gap_ms += ms * (double)h[1] / cyc;
        ctx->profile.total_ms += ms;
        if (timeline_print) {
            const double w = h[9] ? (double)h[9] : 1.0;
            fprintf(stderr, "[local profile] %s walk, %u CTAs x %u slots, window %u x ring %u: %.3f ms, %.0f waves per CTA, "
                    "%.1f us per wave | shade %.1f%% book %.1f%% finish %.1f%% prepass %.1f%% walk %.1f%% | items per wave: "
                    "shade %.0f finish %.0f rays %.0f walk %.0f\n",
                    flat ? "flat" : "stack", grid, p.slots, window, ring, ms, (double)h[9] / ctas, ms * 1e3 / ((double)h[9] / ctas),
                    100.0 * h[0] / cyc, 100.0 * h[1] / cyc, 100.0 * h[2] / cyc, 100.0 * h[3] / cyc, 100.0 * h[4] / cyc,
                    h[5] / w, h[6] / w, h[7] / w, h[8] / w);
        }
    }
    return RT_OK;
}

template <bool COUNT>
cudaError_t launch_render(const RenderParams& p, uint32_t depth, cudaStream_t stream) {
    dim3 grid((p.width + kTileW - 1) / kTileW, (p.height + kTileH - 1) / kTileH);
    dim3 block(kBlockThreads);
    if (depth <= 8) render_kernel<COUNT, 8><<<grid, block, 0, stream>>>(p);
    else if (depth <= 16) render_kernel<COUNT, 16><<<grid, block, 0, stream>>>(p);
    else if (depth <= 32) render_kernel<COUNT, 32><<<grid, block, 0, stream>>>(p);
    else render_kernel<COUNT, 64><<<grid, block, 0, stream>>>(p);
    return cudaGetLastError();
}

template <bool COUNT>
cudaError_t launch_first_hit(const FirstHitParams& p, uint32_t depth, cudaStream_t stream) {
    dim3 grid((p.width + kTileW - 1) / kTileW, (p.height + kTileH - 1) / kTileH);
    dim3 block(kBlockThreads);
    if (depth <= 8) first_hit_kernel<COUNT, 8><<<grid, block, 0, stream>>>(p);
    else if (depth <= 16) first_hit_kernel<COUNT, 16><<<grid, block, 0, stream>>>(p);
    else if (depth <= 32) first_hit_kernel<COUNT, 32><<<grid, block, 0, stream>>>(p);
    else first_hit_kernel<COUNT, 64><<<grid, block, 0, stream>>>(p);
    return cudaGetLastError();
}

int check_ready(RtContext* ctx, uint32_t env_index, bool need_env) {
    if (!ctx->has_scene) return fail(ctx, RT_ERR_STATE, "no scene uploaded (rt_upload_scene)");
    if (!ctx->has_camera) return fail(ctx, RT_ERR_STATE, "no camera set (rt_set_camera)");
    if (!ctx->width || !ctx->height || (!ctx->d_accum && need_env))
        return fail(ctx, RT_ERR_STATE, "no render target (rt_resize)");
    if (need_env && env_index >= ctx->envs.size())
        return fail(ctx, RT_ERR_INVALID_ARGUMENT, "environment index %u out of range (%zu loaded)",
                    env_index, ctx->envs.size());
    return RT_OK;
}

}  // namespace

extern "C" {

const char* rt_backend_version(void) { return "rt_backend 0.1 (sm_100a)"; }

int rt_device_count(int* count) {
    if (!count) return RT_ERR_INVALID_ARGUMENT;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(nullptr, RT_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    return RT_OK;
}

int rt_create(int device, RtContext** out) {
    if (!out) return RT_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, RT_ERR_CUDA, "no usable CUDA device (%s); this backend has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
    if (device < 0 || device >= count)
        return fail(nullptr, RT_ERR_INVALID_ARGUMENT, "device ordinal %d out of range (0..%d)",
                    device, count - 1);
    RtContext* ctx = new RtContext();
    ctx->device = device;
    auto bail = [&](const char* what, cudaError_t err) {
        int st = fail(nullptr, RT_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(err));
        delete ctx;
        return st;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail("cudaStreamCreate", e);
    ctx->stream = ctx->own_stream;
    if ((e = cudaEventCreate(&ctx->ev_start)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&ctx->ev_stop)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaMalloc(&ctx->d_counters, sizeof(RtCounters))) != cudaSuccess)
        return bail("cudaMalloc(counters)", e);
    if ((e = cudaMemset(ctx->d_counters, 0, sizeof(RtCounters))) != cudaSuccess)
        return bail("cudaMemset(counters)", e);
    if ((e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess)
        return bail("cudaDeviceGetAttribute", e);
    if (const char* k = getenv("RT_KERNEL")) {
        // Development switch: the reference-order kernels are kept for the detailed work counters.
        if (!strcmp(k, "pixel")) ctx->kernel = RtContext::KERNEL_PIXEL;
        else if (!strcmp(k, "fast")) ctx->kernel = RtContext::KERNEL_FAST;
        else if (!strcmp(k, "wave")) ctx->kernel = RtContext::KERNEL_WAVE;
        else if (!strcmp(k, "local")) ctx->kernel = RtContext::KERNEL_LOCAL;
    }
    if (const char* k = getenv("RT_LOCAL_SLOTS")) {
        long v = atol(k);
        if (v >= 64 && v <= (long)rtl::kMaxSlots) ctx->local_slots = (uint32_t)v;
    }
    if (const char* k = getenv("RT_LOCAL_WINDOW")) {
        long v = atol(k);
        if (v >= 1 && v <= 256) ctx->local_window = floor_pow2((uint32_t)v);
    }
    if (const char* k = getenv("RT_LOCAL_STAGE")) {
        long v = atol(k);
        if (v >= 64 && v <= (1l << 22)) ctx->local_stage_paths = (uint32_t)v;
    }
    if (const char* k = getenv("RT_LOCAL_FLAT")) ctx->local_flat_allowed = atoi(k) != 0;
    if (const char* k = getenv("RT_WAVE_SLOTS")) {
        long v = atol(k);
        if (v >= 1024 && v <= (1l << 24)) ctx->wave_max_slots = (uint32_t)v;
    }
    if (const char* k = getenv("RT_WAVE_STAGING_MB")) {
        long v = atol(k);
        if (v >= 1 && v <= (1l << 18)) ctx->wave_staging_budget = (size_t)v << 20;
    }
    if (const char* k = getenv("RT_WAVE_GRAPH")) ctx->wave_graph_enabled = atoi(k) != 0;
    if (const char* k = getenv("RT_WAVE_BODY")) {
        int v = atoi(k);
        if (v >= 2 && v <= 32) ctx->wave_body_waves = v & ~1;  // an even number of waves: bodies start on parity 0
    }
    if (const char* k = getenv("RT_WAVE_GROUPS")) {
        int v = atoi(k);
        ctx->wave_groups = v < 1 ? 1 : (v > RtContext::kMaxGroups ? RtContext::kMaxGroups : v);
    }
    if (const char* k = getenv("RT_WAVE_BLOCKS2")) {  // e.g. "4,4,2,2"
        int v[4];
        if (sscanf(k, "%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3]) == 4)
            for (int i = 0; i < 4; i++)
                if (v[i] >= 1 && v[i] <= 16) ctx->wave_blocks2[i] = v[i];
    }
    if ((e = cudaHostAlloc(&ctx->wave_live_host, 2 * RtContext::kMaxGroups * sizeof(uint32_t), cudaHostAllocDefault)) != cudaSuccess)
        return bail("cudaHostAlloc", e);
    for (int g = 0; g < RtContext::kMaxGroups; g++) {
        if ((e = cudaStreamCreateWithFlags(&ctx->wave_streams[g], cudaStreamNonBlocking)) != cudaSuccess)
            return bail("cudaStreamCreate", e);
        if ((e = cudaEventCreateWithFlags(&ctx->wave_join[g], cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&ctx->wave_phase[g], cudaEventDisableTiming)) != cudaSuccess)
            return bail("cudaEventCreate", e);
        for (int i = 0; i < 2; i++)
            if ((e = cudaEventCreateWithFlags(&ctx->wave_events[i][g], cudaEventDisableTiming)) != cudaSuccess)
                return bail("cudaEventCreate", e);
    }
    if ((e = cudaEventCreateWithFlags(&ctx->wave_fork, cudaEventDisableTiming)) != cudaSuccess)
        return bail("cudaEventCreate", e);
    *out = ctx;
    return RT_OK;
}

void rt_destroy(RtContext* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (getenv("RT_WAVE_GRAPH_DEBUG"))
        fprintf(stderr, "[wave graph] %llu instantiations, %llu parameter updates (node updates %s)\n",
                (unsigned long long)ctx->wave_graph_builds, (unsigned long long)ctx->wave_graph_updates,
                ctx->wave_graph_update_ok ? "work" : "refused");
    free_scene(ctx);
    free_envs(ctx);
    cudaFree(ctx->d_accum_own);
    cudaFree(ctx->d_counters);
    if (ctx->wave_exec) cudaGraphExecDestroy(ctx->wave_exec);
    if (ctx->wave_graph) cudaGraphDestroy(ctx->wave_graph);
    cudaFree(ctx->d_graph_launches);
    cudaFree(ctx->d_built_nodes);
    cudaFree(ctx->d_built_prims);
    for (cudaEvent_t ev : ctx->first_hit_events)
        if (ev) cudaEventDestroy(ev);
    cudaFree(ctx->wave_arena);
    cudaFree(ctx->local_arena);
    cudaFree(ctx->d_local_profile);
    cudaFree(ctx->d_first_hit);
    if (ctx->wave_live_host) cudaFreeHost(ctx->wave_live_host);
    cudaFree(ctx->wave_staging);
    for (int g = 0; g < RtContext::kMaxGroups; g++) {
        for (int i = 0; i < 2; i++)
            if (ctx->wave_events[i][g]) cudaEventDestroy(ctx->wave_events[i][g]);
        if (ctx->wave_join[g]) cudaEventDestroy(ctx->wave_join[g]);
        if (ctx->wave_phase[g]) cudaEventDestroy(ctx->wave_phase[g]);
        if (ctx->wave_streams[g]) cudaStreamSynchronize(ctx->wave_streams[g]), cudaStreamDestroy(ctx->wave_streams[g]);
    }
    if (ctx->wave_fork) cudaEventDestroy(ctx->wave_fork);
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->ev_stop) cudaEventDestroy(ctx->ev_stop);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* rt_last_error(const RtContext* ctx) {
    if (ctx) return ctx->error.c_str();
    std::lock_guard<std::mutex> lock(g_error_mutex);
    static thread_local std::string copy;
    copy = g_global_error;
    return copy.c_str();
}

int rt_upload_scene(RtContext* ctx, const RtSceneDesc* s) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    if (!s || !s->bvh_nodes || !s->bvh_node_count || !s->primitives || !s->primitive_count ||
        !s->materials || !s->material_count)
        return fail(ctx, RT_ERR_INVALID_ARGUMENT, "scene needs materials, primitives and BVH nodes");
    if ((s->sphere_count && !s->spheres) || (s->plane_count && !s->planes) ||
        (s->triangle_count && (!s->triangles || !s->vertices || !s->normals)))
        return fail(ctx, RT_ERR_INVALID_ARGUMENT, "scene array pointer is NULL with a non-zero count");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));

    uint32_t depth = 0;
    std::string why;
    if (!validate_bvh(*s, &depth, &why)) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "%s", why.c_str());
    if (depth > 64)
        return fail(ctx, RT_ERR_INVALID_ARGUMENT,
                    "BVH depth %u exceeds the reference's 64-entry traversal stack", depth);

    // ---- re-layout (per-node and per-slot loops on all host threads: 10 M triangles are 2 GB of records) ----
    rtf::RawFloat4Vec nodes;
    nodes.resize((size_t)s->bvh_node_count * 2);
    rtf::parallel_ranges(s->bvh_node_count, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            const RtBvhNode& n = s->bvh_nodes[i];
            uint32_t meta = n.split_axis << 30;
            if (n.primitives_len > 0) {
                meta = n.primitives_len;
                for (uint32_t k = 0; k < n.primitives_len; k++)
                    if (s->primitives[n.primitives_or_second_child_index + k].primitive_type != 2u)
                        meta |= 0x20000000u;  // leaf holds a sphere or plane
            }
            nodes[2 * i] = make_float4(n.min[0], n.min[1], n.min[2], bits(n.primitives_or_second_child_index));
            nodes[2 * i + 1] = make_float4(n.max[0], n.max[1], n.max[2], bits(meta));
        }
    });

    const uint32_t slots = s->primitive_count;
    const size_t rec_count = (size_t)slots + s->sphere_count + s->plane_count;
    rtf::RawFloat4Vec records, tri_normals;
    records.resize(rec_count * 4);
    tri_normals.resize((size_t)slots * 3);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    auto write_sphere = [&](size_t rec, const RtSphere& sp) {
        records[rec * 4 + 0] = make_float4(sp.pos[0], sp.pos[1], sp.pos[2], bits(0u));
        records[rec * 4 + 1] = make_float4(sp.radius, sp.radius * sp.radius, bits(sp.material_id), 0.f);
        records[rec * 4 + 2] = zero4;
        records[rec * 4 + 3] = zero4;
    };
    auto write_plane = [&](size_t rec, const RtPlane& pl) {
        records[rec * 4 + 0] = make_float4(pl.pos[0], pl.pos[1], pl.pos[2], bits(1u));
        records[rec * 4 + 1] = make_float4(pl.normal[0], pl.normal[1], pl.normal[2], bits(pl.material_id));
        records[rec * 4 + 2] = make_float4(pl.base_change[0][0], pl.base_change[1][0], pl.base_change[2][0], 0.f);
        records[rec * 4 + 3] = make_float4(pl.base_change[0][2], pl.base_change[1][2], pl.base_change[2][2], 0.f);
    };
    // The first invalid primitive (lowest slot) is the one reported, whatever thread met it.
    std::mutex bad_mutex;
    size_t bad_slot = (size_t)-1;
    std::string bad_text;
    auto report = [&](size_t slot, const char* fmt, uint32_t a, uint32_t b = 0) {
        char buf[160];
        snprintf(buf, sizeof buf, fmt, a, b);
        std::lock_guard<std::mutex> lock(bad_mutex);
        if (slot < bad_slot) {
            bad_slot = slot;
            bad_text = buf;
        }
    };
    rtf::parallel_ranges(slots, [&](size_t lo, size_t hi) {
        for (size_t slot = lo; slot < hi; slot++) {
            const RtPrimitiveInfo& info = s->primitives[slot];
            float4* nrm = &tri_normals[slot * 3];
            nrm[0] = nrm[1] = nrm[2] = zero4;
            if (info.primitive_type == 0) {
                if (info.index >= s->sphere_count) return report(slot, "primitive %u: sphere index out of range", (uint32_t)slot);
                if (s->spheres[info.index].material_id >= s->material_count)
                    return report(slot, "sphere %u: material out of range", info.index);
                write_sphere(slot, s->spheres[info.index]);
            } else if (info.primitive_type == 1) {
                if (info.index >= s->plane_count) return report(slot, "primitive %u: plane index out of range", (uint32_t)slot);
                if (s->planes[info.index].material_id >= s->material_count)
                    return report(slot, "plane %u: material out of range", info.index);
                write_plane(slot, s->planes[info.index]);
            } else if (info.primitive_type == 2) {
                if (info.index >= s->triangle_count)
                    return report(slot, "primitive %u: triangle index out of range", (uint32_t)slot);
                const RtTriangle& t = s->triangles[info.index];
                if (t.vertex_0 >= s->vertex_count || t.vertex_1 >= s->vertex_count || t.vertex_2 >= s->vertex_count ||
                    t.normal_0 >= s->normal_count || t.normal_1 >= s->normal_count || t.normal_2 >= s->normal_count ||
                    t.material_id >= s->material_count)
                    return report(slot, "triangle %u: index out of range", info.index);
                const RtVec3 &a = s->vertices[t.vertex_0], &b = s->vertices[t.vertex_1], &c = s->vertices[t.vertex_2];
                // edge_0 = b - a, edge_1 = c - a (shader.wgsl:415-416): single f32 subtractions,
                // so hoisting them to upload time leaves every bit unchanged.
                records[slot * 4 + 0] = make_float4(a.x, a.y, a.z, bits(2u));
                records[slot * 4 + 1] = make_float4(b.x - a.x, b.y - a.y, b.z - a.z, bits(t.material_id));
                records[slot * 4 + 2] = make_float4(c.x - a.x, c.y - a.y, c.z - a.z, 0.f);
                records[slot * 4 + 3] = zero4;
                const RtVec3 &n0 = s->normals[t.normal_0], &n1 = s->normals[t.normal_1], &n2 = s->normals[t.normal_2];
                nrm[0] = make_float4(n0.x, n0.y, n0.z, 0.f);
                nrm[1] = make_float4(n1.x, n1.y, n1.z, 0.f);
                nrm[2] = make_float4(n2.x, n2.y, n2.z, 0.f);
            } else {
                return report(slot, "primitive %u: unknown type %u", (uint32_t)slot, info.primitive_type);
            }
        }
    });
    if (bad_slot != (size_t)-1) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "%s", bad_text.c_str());
    for (uint32_t i = 0; i < s->sphere_count; i++) write_sphere((size_t)slots + i, s->spheres[i]);
    for (uint32_t i = 0; i < s->plane_count; i++)
        write_plane((size_t)slots + s->sphere_count + i, s->planes[i]);

    if (slots >= rtf::kLeafSlotMask)
        return fail(ctx, RT_ERR_INVALID_ARGUMENT, "too many primitives (%u) for the 26-bit leaf slot field", slots);
    rtf::FastLayout fl;
    rtf::build_fast_layout(*s, fl);

    // ---- upload: into the arena that is not in use; the scene in use is replaced only when all of
    // the new one is there ----
    struct Piece {
        const void* src;
        size_t bytes, alloc;
        size_t offset;
    };
    std::vector<Piece> pieces;
    size_t arena_need = 0;
    auto piece = [&](const void* src, size_t bytes, size_t alloc_bytes) -> size_t {
        arena_need = (arena_need + 255) & ~(size_t)255;
        pieces.push_back({src, bytes, alloc_bytes ? alloc_bytes : 16, arena_need});
        arena_need += alloc_bytes ? alloc_bytes : 16;
        return pieces.size() - 1;
    };
    auto piece4 = [&](const auto& v, size_t extra = 0) {
        return piece(v.data(), v.size() * sizeof(float4), (v.size() + extra) * sizeof(float4));
    };
    const size_t i_nodes = piece4(nodes), i_records = piece4(records), i_tri_normals = piece4(tri_normals);
    const size_t i_materials = piece(nullptr, 0, (size_t)s->material_count * 3 * sizeof(float4));
    const size_t i_raw_materials = piece(s->materials, (size_t)s->material_count * sizeof(RtMaterial),
                                         (size_t)s->material_count * sizeof(RtMaterial));
    const size_t i_n2 = piece4(fl.n2), i_tri = piece4(fl.tri, 1), i_ana = piece4(fl.ana), i_pre = piece4(fl.pre),
                 i_flat = piece4(fl.flat);
    const size_t i_big_len = fl.big_len.empty() ? (size_t)-1
                                                : piece(fl.big_len.data(), fl.big_len.size() * sizeof(uint32_t),
                                                        fl.big_len.size() * sizeof(uint32_t));
    const size_t i_ana_paths = piece(fl.ana_paths.data(), fl.ana_paths.size() * sizeof(uint32_t),
                                     fl.ana_paths.size() * sizeof(uint32_t));
    arena_need += 256;
    const int spare = ctx->has_scene ? ctx->scene_arena_cur ^ 1 : ctx->scene_arena_cur;
    cudaError_t up_err = cudaSuccess;
    if (ctx->scene_arena_bytes[spare] < arena_need) {
        // the spare arena is not referenced by any queued work: renders use the current one
        cudaFree(ctx->scene_arena[spare]);
        ctx->scene_arena[spare] = nullptr;
        ctx->scene_arena_bytes[spare] = 0;
        up_err = cudaMalloc(&ctx->scene_arena[spare], arena_need);
        if (up_err == cudaSuccess) ctx->scene_arena_bytes[spare] = arena_need;
    }
    char* arena = (char*)ctx->scene_arena[spare];
    auto at = [&](size_t i) -> void* { return i == (size_t)-1 ? nullptr : (void*)(arena + pieces[i].offset); };
    for (size_t i = 0; i < pieces.size() && up_err == cudaSuccess; i++)
        if (pieces[i].bytes)
            up_err = cudaMemcpyAsync(at(i), pieces[i].src, pieces[i].bytes, cudaMemcpyHostToDevice, ctx->stream);
    float4* n_nodes = (float4*)at(i_nodes);
    float4* n_records = (float4*)at(i_records);
    float4* n_tri_normals = (float4*)at(i_tri_normals);
    float4* n_materials = (float4*)at(i_materials);
    float4* n_n2 = (float4*)at(i_n2);
    float4* n_tri = (float4*)at(i_tri);
    float4* n_ana = (float4*)at(i_ana);
    float4* n_pre = (float4*)at(i_pre);
    float4* n_flat = (float4*)at(i_flat);
    uint32_t* n_big_len = (uint32_t*)at(i_big_len);
    uint32_t* n_ana_paths = (uint32_t*)at(i_ana_paths);
    if (up_err == cudaSuccess) {
        bake_materials_kernel<<<(s->material_count + 127) / 128, 128, 0, ctx->stream>>>(
            (const float*)at(i_raw_materials), n_materials, s->material_count);
        ctx->launches++;
        up_err = cudaGetLastError();
    }
    // The host vectors above are pageable: the copies have left them when the calls return, but the
    // stream is drained anyway so that a failure surfaces here, with the old scene still in place.
    if (up_err == cudaSuccess) up_err = cudaStreamSynchronize(ctx->stream);
    if (up_err != cudaSuccess)
        return fail(ctx, up_err == cudaErrorMemoryAllocation ? RT_ERR_OUT_OF_MEMORY : RT_ERR_CUDA,
                    "scene upload failed: %s (the previous scene is still in place)", cudaGetErrorString(up_err));
    ctx->scene_arena_cur = spare;
    ctx->d_nodes = n_nodes;
    ctx->d_records = n_records;
    ctx->d_tri_normals = n_tri_normals;
    ctx->d_materials = n_materials;
    ctx->d_n2 = n_n2;
    ctx->d_tri = n_tri;
    ctx->d_ana = n_ana;
    ctx->d_pre = n_pre;
    ctx->d_flat = n_flat;
    ctx->d_big_len = n_big_len;
    ctx->d_ana_paths = n_ana_paths;
    ctx->fast = fl.view;
    ctx->fast.n2 = ctx->d_n2;
    ctx->fast.tri = ctx->d_tri;
    ctx->fast.ana = ctx->d_ana;
    ctx->fast.pre = ctx->d_pre;
    ctx->fast.flat = ctx->d_flat;
    ctx->fast.big_len = ctx->d_big_len;
    ctx->fast.ana_paths = ctx->d_ana_paths;

    ctx->scene.nodes = ctx->d_nodes;
    ctx->scene.records = ctx->d_records;
    ctx->scene.tri_normals = ctx->d_tri_normals;
    ctx->scene.materials = ctx->d_materials;
    ctx->fast.tri_normals = ctx->d_tri_normals;
    ctx->fast.materials = ctx->d_materials;
    ctx->scene.slot_count = slots;
    ctx->scene.sphere_count = s->sphere_count;
    ctx->scene.plane_count = s->plane_count;
    ctx->scene.prune = rtf::scene_has_loose_plane(*s) ? 0u : 1u;
    ctx->bvh_depth = depth;
    ctx->node_count = s->bvh_node_count;
    ctx->record_count = (uint32_t)rec_count;
    ctx->material_count = s->material_count;
    ctx->has_scene = true;
    return RT_OK;
}

int rt_upload_environments(RtContext* ctx, const RtEnvironmentDesc* envs, uint32_t count) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    if (!envs || !count) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "no environments given");
    for (uint32_t i = 0; i < count; i++)
        if (!envs[i].width || !envs[i].height || !envs[i].rgba || !envs[i].alias ||
            (uint64_t)envs[i].width * envs[i].height > 0xFFFFFFFFull)
            return fail(ctx, RT_ERR_INVALID_ARGUMENT, "environment %u is incomplete", i);
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    free_envs(ctx);
    for (uint32_t i = 0; i < count; i++) {
        RtContext::Env e;
        e.width = envs[i].width;
        e.height = envs[i].height;
        size_t texels = (size_t)e.width * e.height;
        cudaChannelFormatDesc desc = cudaCreateChannelDesc<float4>();
        RT_CUDA(ctx, cudaMallocArray(&e.array, &desc, e.width, e.height));
        RT_CUDA(ctx, cudaMemcpy2DToArray(e.array, 0, 0, envs[i].rgba, (size_t)e.width * 16,
                                         (size_t)e.width * 16, e.height, cudaMemcpyHostToDevice));
        cudaResourceDesc res;
        memset(&res, 0, sizeof res);
        res.resType = cudaResourceTypeArray;
        res.res.array.array = e.array;
        cudaTextureDesc tex;
        memset(&tex, 0, sizeof tex);
        tex.addressMode[0] = tex.addressMode[1] = cudaAddressModeClamp;  // state.rs:135-137
        tex.filterMode = cudaFilterModePoint;  // bilinear blend is done in f32 in the kernel
        tex.readMode = cudaReadModeElementType;
        tex.normalizedCoords = 0;
        RT_CUDA(ctx, cudaCreateTextureObject(&e.texture, &res, &tex, nullptr));
        RT_CUDA(ctx, cudaMalloc(&e.d_alias, 2 * texels * sizeof(float4)));
        e.d_texels = e.d_alias + texels;
        RT_CUDA(ctx, cudaMemcpy(e.d_alias, envs[i].alias, texels * sizeof(RtAliasEntry),
                                cudaMemcpyHostToDevice));
        RT_CUDA(ctx, cudaMemcpy(e.d_texels, envs[i].rgba, texels * sizeof(float4), cudaMemcpyHostToDevice));
        ctx->envs.push_back(e);
    }
    return RT_OK;
}

int rt_set_camera(RtContext* ctx, const RtCameraUniform* cam) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    if (!cam) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "camera is NULL");
    for (int i = 0; i < 3; i++) ctx->camera.pos[i] = cam->pos[i];
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) ctx->camera.rot[c * 3 + r] = cam->rot[c][r];
    ctx->camera.fov_y = cam->fov_y;
    ctx->has_camera = true;
    return RT_OK;
}

int rt_clear(RtContext* ctx) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    if (!ctx->d_accum) return fail(ctx, RT_ERR_STATE, "no render target (rt_resize)");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t count = (size_t)ctx->width * ctx->height;
    clear_accum_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_accum, count);
    ctx->launches++;
    RT_CUDA(ctx, cudaGetLastError());
    return RT_OK;
}

int rt_resize(RtContext* ctx, uint32_t width, uint32_t height) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    // pixel ids are handed out per 8x8 tile, partial tiles included: they must fit 31 bits too
    const uint64_t tile_ids = (((uint64_t)width + 7u) / 8u) * (((uint64_t)height + 7u) / 8u) * 64u;
    if (!width || !height || (uint64_t)width * height > 0x7FFFFFFFull || tile_ids > 0x7FFFFFFFull)
        return fail(ctx, RT_ERR_INVALID_ARGUMENT, "invalid resolution %ux%u", width, height);
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    bool external = ctx->d_accum && ctx->d_accum != ctx->d_accum_own;
    if (external && (width != ctx->width || height != ctx->height))
        return fail(ctx, RT_ERR_STATE, "cannot resize while an external accumulation buffer is bound");
    if (!external) {
        cudaFree(ctx->d_accum_own);
        ctx->d_accum_own = nullptr;
        ctx->d_accum = nullptr;
        RT_CUDA(ctx, cudaMalloc(&ctx->d_accum_own, (size_t)width * height * sizeof(float4)));
        ctx->d_accum = ctx->d_accum_own;
    }
    ctx->width = width;
    ctx->height = height;
    return rt_clear(ctx);
}

int rt_bind_accum(RtContext* ctx, void* device_rgba) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    if (!ctx->width || !ctx->height) return fail(ctx, RT_ERR_STATE, "call rt_resize first");
    if (device_rgba && ((uintptr_t)device_rgba & 15u))
        return fail(ctx, RT_ERR_INVALID_ARGUMENT, "accumulation buffer must be 16-byte aligned");
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->d_accum = device_rgba ? (float4*)device_rgba : ctx->d_accum_own;
    return RT_OK;
}

int rt_accum_device_ptr(RtContext* ctx, void** out) {
    if (!ctx || !out) return RT_ERR_INVALID_ARGUMENT;
    *out = ctx->d_accum;
    return RT_OK;
}

int rt_render(RtContext* ctx, uint32_t sample_begin, uint32_t sample_count, uint32_t max_bounces,
              uint32_t env_index) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    int st = check_ready(ctx, env_index, true);
    if (st != RT_OK) return st;
    if (sample_count == 0) return RT_OK;
    if ((uint64_t)sample_begin + sample_count > 0xFFFFFFFFull)
        return fail(ctx, RT_ERR_INVALID_ARGUMENT, "sample range overflows u32");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    const RtContext::Env& e = ctx->envs[env_index];
    cudaError_t e2;
    if (max_bounces == 0u) {
        // No ray is ever cast: every sample adds zero radiance and writes alpha = 1.
        size_t count = (size_t)ctx->width * ctx->height;
        zero_bounce_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(
            ctx->d_accum, count, ctx->d_counters, sample_count);
        e2 = cudaGetLastError();
    } else if (ctx->kernel == RtContext::KERNEL_LOCAL && !ctx->detailed) {
        return local_render(ctx, e, sample_begin, sample_count, max_bounces);
    } else if (ctx->kernel == RtContext::KERNEL_WAVE && !ctx->detailed) {
        return wave_render(ctx, e, sample_begin, sample_count, max_bounces);
    } else if (ctx->kernel == RtContext::KERNEL_FAST && !ctx->detailed) {
        rtf::FastParams p;
        p.scene = ctx->fast;
        p.env.texture = e.texture;
        p.env.alias = e.d_alias;
        p.env.texels = e.d_texels;
    p.env.texels = e.d_texels;
        p.env.width = e.width;
        p.env.height = e.height;
        p.camera = ctx->camera;
        p.accum = ctx->d_accum;
        p.counters = ctx->d_counters;
        p.slot_out = nullptr;
        p.t_out = nullptr;
        p.width = ctx->width;
        p.height = ctx->height;
        p.sample_begin = sample_begin;
        p.sample_count = sample_count;
        p.max_bounces = max_bounces;
        p.sphere_count = ctx->scene.sphere_count;
        p.row_begin = 0u;
        dim3 grid((p.width + rtf::kFastTileW - 1) / rtf::kFastTileW, (p.height + rtf::kFastTileH - 1) / rtf::kFastTileH);
        rtf::render_fast_kernel<<<grid, rtf::kFastThreads, 0, ctx->stream>>>(p);
        e2 = cudaGetLastError();
    } else {
        RenderParams p;
        p.scene = ctx->scene;
        p.env.texture = e.texture;
        p.env.alias = e.d_alias;
        p.env.texels = e.d_texels;
    p.env.texels = e.d_texels;
        p.env.width = e.width;
        p.env.height = e.height;
        p.camera = ctx->camera;
        p.accum = ctx->d_accum;
        p.counters = ctx->d_counters;
        p.width = ctx->width;
        p.height = ctx->height;
        p.sample_begin = sample_begin;
        p.sample_count = sample_count;
        p.max_bounces = max_bounces;
        e2 = ctx->detailed ? launch_render<true>(p, ctx->bvh_depth, ctx->stream)
                           : launch_render<false>(p, ctx->bvh_depth, ctx->stream);
    }
    ctx->launches++;
    if (e2 != cudaSuccess)
        return fail(ctx, RT_ERR_CUDA, "render kernel launch failed: %s", cudaGetErrorString(e2));
    return RT_OK;
}

int rt_first_hit(RtContext* ctx, uint32_t sample_index, uint32_t* slot_out, float* t_out) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    if (!slot_out) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "slot_out is NULL");
    int st = check_ready(ctx, 0, false);
    if (st != RT_OK) return st;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t count = (size_t)ctx->width * ctx->height;
    // The two device maps live as long as the context (re-allocated when the frame grows).
    if (ctx->first_hit_capacity < count) {
        RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_first_hit);
        ctx->d_first_hit = nullptr;
        ctx->first_hit_capacity = 0;
        RT_CUDA(ctx, cudaMalloc(&ctx->d_first_hit, count * 8));
        ctx->first_hit_capacity = count;
    }
    uint32_t* d_slot = (uint32_t*)ctx->d_first_hit;
    float* d_t = (float*)ctx->d_first_hit + count;
    cudaError_t e = cudaSuccess;
    FirstHitParams p;
    p.scene = ctx->scene;
    p.camera = ctx->camera;
    p.slot_out = d_slot;
    p.t_out = d_t;
    p.counters = ctx->d_counters;
    p.width = ctx->width;
    p.height = ctx->height;
    p.sample = sample_index;
    cudaEventRecord(ctx->ev_start, ctx->stream);
    if (ctx->kernel != RtContext::KERNEL_PIXEL && !ctx->detailed) {
        rtf::FastParams fp;
        memset(&fp, 0, sizeof fp);
        fp.scene = ctx->fast;
        fp.camera = ctx->camera;
        fp.counters = ctx->d_counters;
        fp.slot_out = d_slot;
        fp.t_out = d_t;
        fp.width = ctx->width;
        fp.height = ctx->height;
        fp.sample_begin = sample_index;
        fp.sphere_count = ctx->scene.sphere_count;
        // Large frames go in four bands of rows: the device -> host copies of a band's maps run on a
        // second stream while the next band is traced (the call is copy-bound: 8 bytes per pixel
        // over PCIe against 0.05 ns of kernel time per pixel).
        const uint32_t bands = count >= ((size_t)1 << 21) ? 4u : 1u;
        const uint32_t band_rows = (((fp.height + bands - 1u) / bands) + (uint32_t)rtf::kFastTileH - 1u) /
                                   (uint32_t)rtf::kFastTileH * (uint32_t)rtf::kFastTileH;
        cudaStream_t copy_stream = ctx->wave_streams[0];
        for (int i = 0; i < 8 && bands > 1u && e == cudaSuccess; i++)
            if (!ctx->first_hit_events[i]) e = cudaEventCreate(&ctx->first_hit_events[i]);
        uint32_t bands_run = 0;
        for (uint32_t band = 0; band < bands && e == cudaSuccess; band++) {
            const uint32_t row0 = band * band_rows;
            if (row0 >= fp.height) break;
            const uint32_t rows = fp.height - row0 < band_rows ? fp.height - row0 : band_rows;
            fp.row_begin = row0;
            dim3 grid((fp.width + rtf::kFastTileW - 1) / rtf::kFastTileW, (rows + rtf::kFastTileH - 1) / rtf::kFastTileH);
            // every band's kernel between its own pair of events: the kernel time stays the kernels'
            // alone even when the copies go to pageable memory and block this thread in between
            if (bands > 1u) cudaEventRecord(ctx->first_hit_events[2u * band], ctx->stream);
            rtf::first_hit_fast_kernel<<<grid, rtf::kFastThreads, 0, ctx->stream>>>(fp);
            e = cudaGetLastError();
            ctx->launches++;
            bands_run++;
            if (bands > 1u && e == cudaSuccess) {
                const size_t at = (size_t)row0 * fp.width, n = (size_t)rows * fp.width;
                e = cudaEventRecord(ctx->first_hit_events[2u * band + 1u], ctx->stream);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(copy_stream, ctx->first_hit_events[2u * band + 1u], 0);
                if (e == cudaSuccess) e = cudaMemcpyAsync(slot_out + at, d_slot + at, n * 4, cudaMemcpyDeviceToHost, copy_stream);
                if (e == cudaSuccess && t_out) e = cudaMemcpyAsync(t_out + at, d_t + at, n * 4, cudaMemcpyDeviceToHost, copy_stream);
            }
        }
        ctx->launches--;  // counted once more below
        if (bands > 1u) {
            if (e == cudaSuccess) e = cudaStreamSynchronize(copy_stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            ctx->first_hit_kernel_ms = 0.0f;
            for (uint32_t band = 0; band < bands_run && e == cudaSuccess; band++) {
                float ms = 0.0f;
                cudaEventElapsedTime(&ms, ctx->first_hit_events[2u * band], ctx->first_hit_events[2u * band + 1u]);
                ctx->first_hit_kernel_ms += ms;
            }
            ctx->launches++;
            if (e != cudaSuccess) return fail(ctx, RT_ERR_CUDA, "first-hit pass failed: %s", cudaGetErrorString(e));
            return RT_OK;
        }
    } else {
        e = ctx->detailed ? launch_first_hit<true>(p, ctx->bvh_depth, ctx->stream)
                          : launch_first_hit<false>(p, ctx->bvh_depth, ctx->stream);
    }
    ctx->launches++;
    cudaEventRecord(ctx->ev_stop, ctx->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(slot_out, d_slot, count * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && t_out)
        e = cudaMemcpyAsync(t_out, d_t, count * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) cudaEventElapsedTime(&ctx->first_hit_kernel_ms, ctx->ev_start, ctx->ev_stop);
    if (e != cudaSuccess)
        return fail(ctx, RT_ERR_CUDA, "first-hit pass failed: %s", cudaGetErrorString(e));
    return RT_OK;
}

// Page-locked host memory for the caller's output buffers: the device -> host copies of
// rt_first_hit / rt_read_accum then run at link speed instead of through a bounce buffer.
int rt_host_alloc(size_t bytes, void** out) {
    if (!out || !bytes) return RT_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess)
        return fail(nullptr, RT_ERR_OUT_OF_MEMORY, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return RT_OK;
}

void rt_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int rt_sync(RtContext* ctx) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return wave_loop_check(ctx);
}

int rt_read_accum(RtContext* ctx, float* rgba_out) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    if (!rgba_out) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "output pointer is NULL");
    if (!ctx->d_accum) return fail(ctx, RT_ERR_STATE, "no render target (rt_resize)");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaMemcpyAsync(rgba_out, ctx->d_accum,
                                 (size_t)ctx->width * ctx->height * sizeof(float4),
                                 cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return wave_loop_check(ctx);
}

int rt_write_accum(RtContext* ctx, const float* rgba_in) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    if (!rgba_in) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "input pointer is NULL");
    if (!ctx->d_accum) return fail(ctx, RT_ERR_STATE, "no render target (rt_resize)");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaMemcpyAsync(ctx->d_accum, rgba_in,
                                 (size_t)ctx->width * ctx->height * sizeof(float4),
                                 cudaMemcpyHostToDevice, ctx->stream));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

int rt_tonemap_rgb8(RtContext* ctx, uint32_t sample_total, uint8_t* rgb8_out) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    if (!rgb8_out || !sample_total)
        return fail(ctx, RT_ERR_INVALID_ARGUMENT, "rgb8_out NULL or sample_total 0");
    if (!ctx->d_accum) return fail(ctx, RT_ERR_STATE, "no render target (rt_resize)");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t count = (size_t)ctx->width * ctx->height;
    uint8_t* d_rgb = nullptr;
    RT_CUDA(ctx, cudaMalloc(&d_rgb, count * 3));
    tonemap_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(
        ctx->d_accum, d_rgb, count, 0.f, (float)sample_total);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(rgb8_out, d_rgb, count * 3, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_rgb);
    if (e != cudaSuccess) return fail(ctx, RT_ERR_CUDA, "tone-map pass failed: %s", cudaGetErrorString(e));
    return RT_OK;
}

int rt_debug_view(RtContext* ctx, uint32_t dev_index, uint32_t environment_index, uint32_t sample_index,
                  float* rgba_out) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    if (!rgba_out) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "output pointer is NULL");
    if (dev_index != 2u && dev_index != 3u)
        return fail(ctx, RT_ERR_INVALID_ARGUMENT, "dev_index %u is not a view: only 2 (environment sampling "
                    "distribution) and 3 (environment texels) leave the path tracer", dev_index);
    if (!ctx->width || !ctx->height) return fail(ctx, RT_ERR_STATE, "no render target (rt_resize)");
    if (environment_index >= ctx->envs.size())
        return fail(ctx, RT_ERR_INVALID_ARGUMENT, "environment index %u out of range (%zu loaded)",
                    environment_index, ctx->envs.size());
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    const RtContext::Env& e = ctx->envs[environment_index];
    EnvView env;
    env.texture = e.texture;
    env.alias = e.d_alias;
    env.texels = e.d_texels;
    env.width = e.width;
    env.height = e.height;
    const size_t count = (size_t)ctx->width * ctx->height;
    float4* d_out = nullptr;
    uint32_t* d_counts = nullptr;
    RT_CUDA(ctx, cudaMalloc(&d_out, count * sizeof(float4)));
    cudaError_t ce = cudaSuccess;
    dim3 block(16, 16), grid((ctx->width + 15u) / 16u, (ctx->height + 15u) / 16u);  // the shader's workgroups
    if (dev_index == 2u) {
        ce = cudaMalloc(&d_counts, count * sizeof(uint32_t));
        if (ce == cudaSuccess) ce = cudaMemsetAsync(d_counts, 0, count * sizeof(uint32_t), ctx->stream);
        if (ce == cudaSuccess) {
            view_distribution_count_kernel<<<grid, block, 0, ctx->stream>>>(env, ctx->width, ctx->height, sample_index, d_counts);
            view_distribution_resolve_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(d_counts, d_out, count);
            ctx->launches += 2;
            ce = cudaGetLastError();
        }
    } else {
        view_environment_kernel<<<grid, block, 0, ctx->stream>>>(env, ctx->width, ctx->height, d_out);
        ctx->launches++;
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(rgba_out, d_out, count * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_out);
    cudaFree(d_counts);
    if (ce != cudaSuccess) return fail(ctx, RT_ERR_CUDA, "debug view failed: %s", cudaGetErrorString(ce));
    return RT_OK;
}

int rt_set_profiling(RtContext* ctx, int enabled) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    ctx->profiling = enabled != 0;
    memset(&ctx->profile, 0, sizeof ctx->profile);
    return RT_OK;
}

int rt_get_profile(RtContext* ctx, RtWaveProfile* out) {
    if (!ctx || !out) return RT_ERR_INVALID_ARGUMENT;
    *out = ctx->profile;
    return RT_OK;
}

int rt_set_memory_budget(RtContext* ctx, uint64_t staging_bytes, uint32_t path_slots) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    if (staging_bytes) {
        if (staging_bytes < ((uint64_t)1 << 20)) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "staging budget below 1 MiB");
        ctx->wave_staging_budget = (size_t)staging_bytes;
    }
    if (path_slots) {
        if (path_slots < 1024u || path_slots > (1u << 24))
            return fail(ctx, RT_ERR_INVALID_ARGUMENT, "path slot count %u outside [1024, 2^24]", path_slots);
        ctx->wave_max_slots = path_slots;
    }
    return RT_OK;
}

int rt_get_memory_usage(RtContext* ctx, uint64_t* bytes) {
    if (!ctx || !bytes) return RT_ERR_INVALID_ARGUMENT;
    *bytes = (uint64_t)ctx->wave_staging_bytes + (uint64_t)ctx->wave_arena_bytes + (uint64_t)ctx->local_arena_bytes +
             (uint64_t)ctx->first_hit_capacity * 8u +
             (ctx->d_accum_own ? (uint64_t)ctx->width * ctx->height * sizeof(float4) : 0u);
    return RT_OK;
}

// ---- device BVH build (rt_bvh_build.cuh) ----
}  // extern "C"
namespace {

struct ArenaCarver {
    char* base;
    size_t used = 0;
    explicit ArenaCarver(char* b) : base(b) {}
    template <typename T>
    T* take(size_t count) {
        used = (used + 255) & ~(size_t)255;
        T* p = base ? (T*)(base + used) : nullptr;
        used += count * sizeof(T);
        return p;
    }
};

// Lays the build's buffers out in one allocation (base == nullptr: only measures).
size_t bvh_carve(rtb::Build& b, char* base, uint32_t n, uint32_t s, uint32_t pl, uint32_t v, uint32_t t, RtSphere** d_spheres,
                 float** d_plane_bounds, RtVec3** d_vertices, RtTriangle** d_triangles) {
    ArenaCarver a(base);
    b.n = n;
    b.node_cap = 2u * n;
    b.acc_cap = n / 6u + 2u;
    for (int i = 0; i < 2; i++) {
        b.rmin[i] = a.take<float4>(n);
        b.rmax[i] = a.take<float4>(n);
        b.node_of[i] = a.take<uint32_t>(n);
        b.active[i] = a.take<uint32_t>(b.acc_cap);
    }
    b.bucket = a.take<uint8_t>(n);
    b.prefix = a.take<uint32_t>((size_t)n + 1);
    b.frpos = a.take<uint32_t>(n);
    b.bgpos = a.take<uint32_t>(n);
    b.blocksum = a.take<uint32_t>((size_t)n / rtb::kScanBlock + 4);
    b.acc = a.take<int>((size_t)rtb::kFields * rtb::kBuckets * b.acc_cap);
    b.counters = a.take<uint32_t>(4);
    rtb::Nodes& nd = b.nodes;
    const size_t c = b.node_cap;
    nd.begin = a.take<uint32_t>(c);
    nd.count = a.take<uint32_t>(c);
    nd.left = a.take<uint32_t>(c);
    nd.axis = a.take<uint32_t>(c);
    nd.split = a.take<uint32_t>(c);
    nd.good = a.take<uint32_t>(c);
    nd.slot = a.take<uint32_t>(c);
    nd.bmin = a.take<float4>(c);
    nd.bmax = a.take<float4>(c);
    nd.cmin = a.take<float4>(c);
    nd.cmax = a.take<float4>(c);
    nd.size = a.take<uint32_t>(c);
    nd.depth = a.take<uint32_t>(c);
    nd.pre = a.take<uint32_t>(c);
    *d_spheres = a.take<RtSphere>(s);
    *d_plane_bounds = a.take<float>((size_t)pl * 6);
    *d_vertices = a.take<RtVec3>(v);
    *d_triangles = a.take<RtTriangle>(t);
    return a.used + 256;
}

void bvh_release_result(RtContext* ctx) {
    cudaFree(ctx->d_built_nodes);
    cudaFree(ctx->d_built_prims);
    ctx->d_built_nodes = nullptr;
    ctx->d_built_prims = nullptr;
    ctx->built_node_count = ctx->built_prim_count = 0;
}

}  // namespace
extern "C" {

int rt_bvh_build(RtContext* ctx, const RtBvhBuildInput* in, uint32_t* node_count_out, uint32_t* depth_out,
                 float* device_ms_out) {
    if (!ctx || !in) return RT_ERR_INVALID_ARGUMENT;
    const uint64_t total = (uint64_t)in->sphere_count + in->plane_count + in->triangle_count;
    if (total == 0) return fail(ctx, RT_ERR_SCENE, "scene has no primitives");  // bvh.rs:222
    if (total > 0x3FFFFFFFull) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "too many primitives for a device build");
    if ((in->sphere_count && !in->spheres) || (in->plane_count && !in->plane_bounds) ||
        (in->triangle_count && (!in->triangles || !in->vertices)))
        return fail(ctx, RT_ERR_INVALID_ARGUMENT, "rt_bvh_build: missing input array");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    bvh_release_result(ctx);
    const uint32_t n = (uint32_t)total;
    cudaStream_t st = ctx->stream;

    rtb::Build b;
    memset(&b, 0, sizeof b);
    RtSphere* d_spheres = nullptr;
    float* d_plane_bounds = nullptr;
    RtVec3* d_vertices = nullptr;
    RtTriangle* d_triangles = nullptr;
    const size_t bytes = bvh_carve(b, nullptr, n, in->sphere_count, in->plane_count, in->vertex_count, in->triangle_count,
                                   &d_spheres, &d_plane_bounds, &d_vertices, &d_triangles);
    char* arena = nullptr;
    RT_CUDA(ctx, cudaMalloc(&arena, bytes));
    bvh_carve(b, arena, n, in->sphere_count, in->plane_count, in->vertex_count, in->triangle_count, &d_spheres,
              &d_plane_bounds, &d_vertices, &d_triangles);
    int status = RT_OK;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<uint32_t> level_start;
    uint32_t host_counters[4] = {0, 0, 0, 0};
    auto check = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && status == RT_OK)
            status = fail(ctx, e == cudaErrorMemoryAllocation ? RT_ERR_OUT_OF_MEMORY : RT_ERR_CUDA, "%s failed: %s", what,
                          cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    auto grid = [](size_t items, int threads) { return (unsigned)((items + (size_t)threads - 1) / (size_t)threads); };
    do {
        if (in->sphere_count && !check(cudaMemcpyAsync(d_spheres, in->spheres, (size_t)in->sphere_count * sizeof(RtSphere), cudaMemcpyHostToDevice, st), "sphere upload")) break;
        if (in->plane_count && !check(cudaMemcpyAsync(d_plane_bounds, in->plane_bounds, (size_t)in->plane_count * 6 * sizeof(float), cudaMemcpyHostToDevice, st), "plane upload")) break;
        if (in->triangle_count) {
            if (!check(cudaMemcpyAsync(d_vertices, in->vertices, (size_t)in->vertex_count * sizeof(RtVec3), cudaMemcpyHostToDevice, st), "vertex upload")) break;
            if (!check(cudaMemcpyAsync(d_triangles, in->triangles, (size_t)in->triangle_count * sizeof(RtTriangle), cudaMemcpyHostToDevice, st), "triangle upload")) break;
        }
        if (!check(cudaEventCreate(&ev0), "cudaEventCreate") || !check(cudaEventCreate(&ev1), "cudaEventCreate")) break;
        cudaEventRecord(ev0, st);
        cudaMemsetAsync(b.counters, 0, 4 * sizeof(uint32_t), st);
        rtb::init_records_kernel<<<grid(n, rtb::kThreads), rtb::kThreads, 0, st>>>(
            b, d_spheres, in->sphere_count, d_plane_bounds, in->plane_count, d_vertices, in->vertex_count, d_triangles,
            n <= rtb::kLeafMax ? 1u : 0u);
        rtb::acc_init_kernel<<<grid((size_t)rtb::kFields * rtb::kBuckets, 256), 256, 0, st>>>(b, 1u);
        rtb::bucket_kernel<true><<<grid(n, rtb::kThreads), rtb::kThreads, 0, st>>>(b, 0);
        rtb::root_kernel<<<1, 32, 0, st>>>(b);
        ctx->launches += 4;
        if (!check(cudaMemcpyAsync(host_counters, b.counters, sizeof host_counters, cudaMemcpyDeviceToHost, st), "counter read")) break;
        if (!check(cudaStreamSynchronize(st), "bvh root pass")) break;
        if (host_counters[2] & rtb::kErrBadIndex) {  // checked by init_records_kernel
            status = fail(ctx, RT_ERR_INVALID_ARGUMENT, "a triangle names a vertex out of range");
            break;
        }
        level_start.push_back(0u);
        level_start.push_back(1u);
        int cur = 0, act = 0;
        uint32_t active = host_counters[1];
        const unsigned scan_blocks = grid((size_t)n + 1, rtb::kScanBlock);
        while (active > 0u && status == RT_OK) {
            rtb::acc_init_kernel<<<grid((size_t)rtb::kFields * rtb::kBuckets * active, 256), 256, 0, st>>>(b, active);
            rtb::bucket_kernel<false><<<grid(n, rtb::kThreads), rtb::kThreads, 0, st>>>(b, cur);
            cudaMemsetAsync(b.counters + 1, 0, sizeof(uint32_t), st);
            rtb::split_kernel<<<grid(active, 128), 128, 0, st>>>(b, act, active);
            rtb::flag_count_kernel<<<scan_blocks, rtb::kThreads, 0, st>>>(b, cur);
            rtb::blocksum_scan_kernel<<<1, 1024, 0, st>>>(b, scan_blocks);
            rtb::prefix_kernel<<<scan_blocks, rtb::kThreads, 0, st>>>(b, cur);
            rtb::scatter_kernel<<<grid(n, rtb::kThreads), rtb::kThreads, 0, st>>>(b, cur);
            rtb::move_kernel<<<grid(n, rtb::kThreads), rtb::kThreads, 0, st>>>(b, cur);
            ctx->launches += 8;
            cur ^= 1;
            act ^= 1;
            if (!check(cudaMemcpyAsync(host_counters, b.counters, sizeof host_counters, cudaMemcpyDeviceToHost, st), "counter read")) break;
            if (!check(cudaStreamSynchronize(st), "bvh level pass")) break;
            if (host_counters[2] & rtb::kErrNanCost) {
                status = fail(ctx, RT_ERR_SCENE, "bounds have non-zero surface area");  // NotNan::new(..).expect, bvh.rs:296
                break;
            }
            if (host_counters[2]) {
                status = fail(ctx, RT_ERR_SCENE, host_counters[2] & rtb::kErrEmptySide ? "BVH partition produced an empty side"
                                                                                        : "BVH node array overflow");
                break;
            }
            level_start.push_back(host_counters[0]);
            active = host_counters[1];
            if (level_start.size() > 4096u) {
                status = fail(ctx, RT_ERR_STATE, "BVH build did not terminate");
                break;
            }
        }
        if (status != RT_OK) break;
        const uint32_t node_count = host_counters[0];
        if (!check(cudaMalloc(&ctx->d_built_nodes, (size_t)node_count * sizeof(RtBvhNode)), "cudaMalloc(nodes)")) break;
        if (!check(cudaMalloc(&ctx->d_built_prims, (size_t)n * sizeof(RtPrimitiveInfo)), "cudaMalloc(primitives)")) break;
        rtb::leaf_bounds_kernel<<<grid(node_count, 128), 128, 0, st>>>(b, cur, node_count);
        ctx->launches++;
        const size_t levels = level_start.size() - 1;  // level L holds nodes [level_start[L], level_start[L + 1])
        for (size_t l = levels; l-- > 0;) {
            const uint32_t first = level_start[l], last = level_start[l + 1];
            if (last > first) rtb::up_kernel<<<grid(last - first, 128), 128, 0, st>>>(b, first, last);
            ctx->launches++;
        }
        for (size_t l = 0; l < levels; l++) {
            const uint32_t first = level_start[l], last = level_start[l + 1];
            if (last > first) rtb::down_kernel<<<grid(last - first, 128), 128, 0, st>>>(b, first, last);
            ctx->launches++;
        }
        rtb::emit_nodes_kernel<<<grid(node_count, 128), 128, 0, st>>>(b, node_count, ctx->d_built_nodes);
        rtb::emit_primitives_kernel<<<grid(n, 256), 256, 0, st>>>(b, cur, ctx->d_built_prims);
        ctx->launches += 2;
        cudaEventRecord(ev1, st);
        uint32_t depth = 0;
        if (!check(cudaMemcpyAsync(&depth, b.nodes.depth, sizeof depth, cudaMemcpyDeviceToHost, st), "depth read")) break;
        if (!check(cudaStreamSynchronize(st), "bvh finish")) break;
        if (!check(cudaGetLastError(), "bvh kernels")) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev0, ev1);
        ctx->built_node_count = node_count;
        ctx->built_prim_count = n;
        if (node_count_out) *node_count_out = node_count;
        if (depth_out) *depth_out = depth;
        if (device_ms_out) *device_ms_out = ms;
    } while (false);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    cudaStreamSynchronize(st);
    cudaFree(arena);
    if (status != RT_OK) bvh_release_result(ctx);
    return status;
}

int rt_bvh_build_fetch(RtContext* ctx, RtBvhNode* nodes_out, RtPrimitiveInfo* primitives_out) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    if (!ctx->d_built_nodes) return fail(ctx, RT_ERR_STATE, "rt_bvh_build_fetch without a built BVH");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    if (nodes_out)
        RT_CUDA(ctx, cudaMemcpyAsync(nodes_out, ctx->d_built_nodes, (size_t)ctx->built_node_count * sizeof(RtBvhNode),
                                     cudaMemcpyDeviceToHost, ctx->stream));
    if (primitives_out)
        RT_CUDA(ctx, cudaMemcpyAsync(primitives_out, ctx->d_built_prims, (size_t)ctx->built_prim_count * sizeof(RtPrimitiveInfo),
                                     cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    bvh_release_result(ctx);
    return RT_OK;
}

int rt_set_detailed_counters(RtContext* ctx, int enabled) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    ctx->detailed = enabled != 0;
    return RT_OK;
}

int rt_reset_counters(RtContext* ctx) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(RtCounters), ctx->stream));
    return RT_OK;
}

int rt_get_counters(RtContext* ctx, RtCounters* out) {
    if (!ctx || !out) return RT_ERR_INVALID_ARGUMENT;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_counters, sizeof(RtCounters), cudaMemcpyDeviceToHost,
                                 ctx->stream));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

int rt_set_stream(RtContext* ctx, void* cuda_stream) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = (cudaStream_t)cuda_stream;
    return RT_OK;
}

int rt_reset_stream(RtContext* ctx) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = ctx->own_stream;
    return RT_OK;
}

int rt_timer_start(RtContext* ctx) {
    if (!ctx) return RT_ERR_INVALID_ARGUMENT;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
    return RT_OK;
}

int rt_timer_stop(RtContext* ctx, float* elapsed_ms) {
    if (!ctx || !elapsed_ms) return RT_ERR_INVALID_ARGUMENT;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaEventRecord(ctx->ev_stop, ctx->stream));
    RT_CUDA(ctx, cudaEventSynchronize(ctx->ev_stop));
    RT_CUDA(ctx, cudaEventElapsedTime(elapsed_ms, ctx->ev_start, ctx->ev_stop));
    return RT_OK;
}

uint64_t rt_launch_count(const RtContext* ctx) {
    if (!ctx) return 0;
    unsigned long long in_graphs = 0;  // kernels the wave loop's graph bodies ran: counted on the device
    if (ctx->d_graph_launches) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(&in_graphs, ctx->d_graph_launches, sizeof in_graphs, cudaMemcpyDeviceToHost);
    }
    return ctx->launches + in_graphs;
}

float rt_first_hit_kernel_ms(const RtContext* ctx) { return ctx ? ctx->first_hit_kernel_ms : 0.f; }

}  // extern "C"

// =====================================================================================
// Single-process multi-GPU group: sample split + NCCL sum-reduce. NCCL is bound lazily
// with dlopen so that single-GPU users do not need the library on their loader path.
// =====================================================================================

namespace {

struct NcclApi {
    void* handle = nullptr;
    std::string load_error;  // dlerror() text of the failed dlopen (dlerror clears itself when read)
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t,
                           cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!api.handle) {
            const char* e = dlerror();
            api.load_error = e ? e : "dlopen failed";
            return;
        }
        api.CommInitAll = (decltype(api.CommInitAll))dlsym(api.handle, "ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
        api.Reduce = (decltype(api.Reduce))dlsym(api.handle, "ncclReduce");
        api.GroupStart = (decltype(api.GroupStart))dlsym(api.handle, "ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.handle, "ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
        api.ok = api.CommInitAll && api.CommDestroy && api.Reduce && api.GroupStart &&
                 api.GroupEnd && api.GetErrorString;
        if (!api.ok) api.load_error = "missing symbols";
    });
    return api;
}

}  // namespace

struct RtGroup {
    std::vector<RtContext*> ctxs;
    std::vector<ncclComm_t> comms;
};

extern "C" {

int rt_group_create(int count, const int* devices, RtGroup** out) {
    if (!out || count <= 0 || !devices) return RT_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    RtGroup* g = new RtGroup();
    for (int i = 0; i < count; i++) {
        RtContext* c = nullptr;
        int st = rt_create(devices[i], &c);
        if (st != RT_OK) {
            rt_group_destroy(g);
            return st;
        }
        g->ctxs.push_back(c);
    }
    if (count > 1) {
        NcclApi& api = nccl_api();
        if (!api.ok) {
            rt_group_destroy(g);
            return fail(nullptr, RT_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", api.load_error.c_str());
        }
        g->comms.resize((size_t)count);
        ncclResult_t r = api.CommInitAll(g->comms.data(), count, devices);
        if (r != ncclSuccess) {
            g->comms.clear();
            rt_group_destroy(g);
            return fail(nullptr, RT_ERR_NCCL, "ncclCommInitAll failed: %s", api.GetErrorString(r));
        }
    }
    *out = g;
    return RT_OK;
}

void rt_group_destroy(RtGroup* g) {
    if (!g) return;
    for (ncclComm_t c : g->comms) nccl_api().CommDestroy(c);
    for (RtContext* c : g->ctxs) rt_destroy(c);
    delete g;
}

int rt_group_size(const RtGroup* g) { return g ? (int)g->ctxs.size() : 0; }

RtContext* rt_group_context(RtGroup* g, int rank) {
    if (!g || rank < 0 || rank >= (int)g->ctxs.size()) return nullptr;
    return g->ctxs[(size_t)rank];
}

int rt_group_render(RtGroup* g, uint32_t sample_begin, uint32_t sample_count, uint32_t max_bounces,
                    uint32_t env_index) {
    if (!g) return RT_ERR_INVALID_ARGUMENT;
    uint32_t n = (uint32_t)g->ctxs.size();
    // rt_render paces its waves from the calling thread, so every device gets its own.
    std::vector<int> status(n, RT_OK);
    std::vector<std::thread> workers;
    for (uint32_t r = 0; r < n; r++) {
        // Contiguous, disjoint sample ranges whose union is the single-GPU range.
        uint32_t lo = (uint32_t)((uint64_t)sample_count * r / n);
        uint32_t hi = (uint32_t)((uint64_t)sample_count * (r + 1) / n);
        workers.emplace_back([=, &status] {
            status[r] = rt_render(g->ctxs[r], sample_begin + lo, hi - lo, max_bounces, env_index);
        });
    }
    for (auto& w : workers) w.join();
    for (uint32_t r = 0; r < n; r++)
        if (status[r] != RT_OK) return status[r];
    return RT_OK;
}

int rt_group_reduce(RtGroup* g, int root) {
    if (!g || root < 0 || root >= (int)g->ctxs.size()) return RT_ERR_INVALID_ARGUMENT;
    if (g->ctxs.size() == 1) return RT_OK;
    NcclApi& api = nccl_api();
    RtContext* r0 = g->ctxs[(size_t)root];
    size_t floats = (size_t)r0->width * r0->height * 4;
    for (RtContext* c : g->ctxs)
        if (c->width != r0->width || c->height != r0->height || !c->d_accum)
            return fail(r0, RT_ERR_STATE, "group members have different render targets");
    ncclResult_t r = api.GroupStart();
    for (size_t i = 0; r == ncclSuccess && i < g->ctxs.size(); i++) {
        RtContext* c = g->ctxs[i];
        cudaSetDevice(c->device);
        r = api.Reduce(c->d_accum, c->d_accum, floats, ncclFloat32, ncclSum, root, g->comms[i],
                       c->stream);
    }
    ncclResult_t r2 = api.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess)
        return fail(r0, RT_ERR_NCCL, "ncclReduce failed: %s", api.GetErrorString(r));
    // Post-conditions (rt_backend.h): the root holds the sum with alpha = 1; every other rank's
    // buffer is zero again, so a further rt_group_render + rt_group_reduce round adds only its
    // own samples (progressive accumulation, one round per batch of frames).
    const size_t count = floats / 4;
    for (size_t i = 0; i < g->ctxs.size(); i++) {
        RtContext* c = g->ctxs[i];
        cudaSetDevice(c->device);
        if ((int)i == root) set_alpha_kernel<<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(c->d_accum, count);
        else clear_accum_kernel<<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(c->d_accum, count);
        c->launches++;
        cudaError_t ce = cudaGetLastError();
        if (ce != cudaSuccess) return fail(r0, RT_ERR_CUDA, "post-reduce kernel failed: %s", cudaGetErrorString(ce));
    }
    return RT_OK;
}

int rt_group_sync(RtGroup* g) {
    if (!g) return RT_ERR_INVALID_ARGUMENT;
    for (RtContext* c : g->ctxs) {
        int st = rt_sync(c);
        if (st != RT_OK) return st;
    }
    return RT_OK;
}

}  // extern "C"
