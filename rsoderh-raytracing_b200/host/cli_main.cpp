// cli_main.cpp — headless renderer: the reference's CLI (src/cli.rs:10-46) with the
// window replaced by render-to-file.
//
// Kept from the reference: --scene (required, repeatable, last wins), --state (base64
// camera), --movement-keys / --other-keys (accepted and validated with the reference's
// messages and exit code 2; they have no effect without a window).
// Added for headless use: --width --height --spp --bounces --env --hdri --out --pfm
// --first-hit-out --gpus --sample-begin --bvh.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rt_backend.h"
#include "rt_host.h"

namespace {

struct Args {
    std::string movement_keys = "wasdqe", other_keys = "cpe";
    std::string state, out, pfm, first_hit_out;
    std::vector<std::string> scenes, hdris;
    uint32_t width = 1280, height = 720, spp = 64, bounces = 10, env = 0, sample_begin = 0;
    int gpus = 1;
    std::string bvh = "host";  // "device": rebuild the BVH with rt_bvh_build (same bytes, for large scenes)
};

[[noreturn]] void usage(const char* msg) {
    if (msg) std::fprintf(stderr, "error: %s\n\n", msg);
    std::fprintf(stderr,
                 "Usage: rt_render --scene <SCENE> [--scene <SCENE> ...] [--state <BASE64>]\n"
                 "                 [--width W] [--height H] [--spp N] [--bounces N] [--env I]\n"
                 "                 [--hdri file.hdr ...] [--out image.png] [--pfm image.pfm]\n"
                 "                 [--first-hit-out slots.u32] [--gpus N] [--sample-begin S] [--bvh host|device]\n"
                 "                 [--movement-keys wasdqe] [--other-keys cpe]\n");
    std::exit(2);
}

bool take(int argc, char** argv, int& i, const char* name, std::string& value) {
    size_t n = std::strlen(name);
    if (std::strncmp(argv[i], name, n) != 0) return false;
    if (argv[i][n] == '=') {
        value = argv[i] + n + 1;
        return true;
    }
    if (argv[i][n] == '\0') {
        if (i + 1 >= argc) usage((std::string("a value is required for '") + name + "'").c_str());
        value = argv[++i];
        return true;
    }
    return false;
}

int die(const char* what, const char* detail) {
    std::fprintf(stderr, "Error: %s%s%s\n", what, detail && *detail ? ": " : "", detail ? detail : "");
    return 1;
}

}  // namespace

int main(int argc, char** argv) {
    Args a;
    for (int i = 1; i < argc; i++) {
        std::string v;
        if (take(argc, argv, i, "--scene", v)) a.scenes.push_back(v);
        else if (take(argc, argv, i, "--state", v)) a.state = v;
        else if (take(argc, argv, i, "--movement-keys", v)) a.movement_keys = v;
        else if (take(argc, argv, i, "--other-keys", v)) a.other_keys = v;
        else if (take(argc, argv, i, "--width", v)) a.width = (uint32_t)std::strtoul(v.c_str(), nullptr, 10);
        else if (take(argc, argv, i, "--height", v)) a.height = (uint32_t)std::strtoul(v.c_str(), nullptr, 10);
        else if (take(argc, argv, i, "--spp", v)) a.spp = (uint32_t)std::strtoul(v.c_str(), nullptr, 10);
        else if (take(argc, argv, i, "--bounces", v)) a.bounces = (uint32_t)std::strtoul(v.c_str(), nullptr, 10);
        else if (take(argc, argv, i, "--env", v)) a.env = (uint32_t)std::strtoul(v.c_str(), nullptr, 10);
        else if (take(argc, argv, i, "--sample-begin", v)) a.sample_begin = (uint32_t)std::strtoul(v.c_str(), nullptr, 10);
        else if (take(argc, argv, i, "--gpus", v)) a.gpus = std::atoi(v.c_str());
        else if (take(argc, argv, i, "--bvh", v)) a.bvh = v;
        else if (take(argc, argv, i, "--hdri", v)) a.hdris.push_back(v);
        else if (take(argc, argv, i, "--out", v)) a.out = v;
        else if (take(argc, argv, i, "--pfm", v)) a.pfm = v;
        else if (take(argc, argv, i, "--first-hit-out", v)) a.first_hit_out = v;
        else if (!std::strcmp(argv[i], "--help") || !std::strcmp(argv[i], "-h")) usage(nullptr);
        else usage((std::string("unexpected argument '") + argv[i] + "'").c_str());
    }
    if (a.scenes.empty())  // clap: `required = true` (cli.rs:24)
        usage("the following required arguments were not provided:\n  --scene <SCENE>");
    // KeyboardLayout::parse_config (camera.rs:139-181), exit code 2 (cli.rs:31-35).
    if (a.movement_keys.size() != 6) {
        std::fprintf(stderr, "Invalid keyboard config: Invalid keyboard config '%s': expected 6 characters.\n",
                     a.movement_keys.c_str());
        return 2;
    }
    if (a.other_keys.size() != 3) {
        std::fprintf(stderr, "Invalid keyboard config: Invalid mouse capture config '%s': expected 3 character.\n",
                     a.other_keys.c_str());
        return 2;
    }
    if (!a.width || !a.height || !a.spp || a.gpus < 1) usage("width, height, spp and gpus must be positive");
    if (a.bvh != "host" && a.bvh != "device") usage("--bvh takes 'host' or 'device'");

    char err[1024] = {0};
    RtHostScene* scene = nullptr;
    if (rt_scene_load_toml(a.scenes.back().c_str(), &scene, err, sizeof err) != RT_OK)  // last wins (cli.rs:37)
        return die(err, nullptr);

    float pos[3], yaw, pitch, fov;
    rt_scene_camera(scene, pos, &yaw, &pitch, &fov);
    if (!a.state.empty() &&
        rt_camera_deserialize(a.state.c_str(), pos, &yaw, &pitch, &fov, err, sizeof err) != RT_OK)
        return die(err, nullptr);  // cli.rs:39-43 unwraps
    RtCameraUniform camera;
    rt_camera_uniform(pos, yaw, pitch, fov, &camera);

    // Environment slots: explicit --hdri files, else the reference's two Poly Haven files
    // if present next to the scene assets, else the synthetic sky.
    struct EnvData {
        uint32_t w = 0, h = 0;
        std::vector<float> rgba;
        std::vector<RtAliasEntry> alias;
    };
    std::vector<EnvData> envs;
    auto add_env = [&](uint32_t w, uint32_t h, const float* rgb) {
        EnvData e;
        e.w = w;
        e.h = h;
        e.rgba.resize((size_t)w * h * 4);
        e.alias.resize((size_t)w * h);
        rt_rgb_to_rgba(w * h, rgb, e.rgba.data());
        uint32_t leftover = 0;
        rt_alias_build(w, h, rgb, e.alias.data(), &leftover);
        std::fprintf(stderr, "AliasTable: %u left over pixels out of %u\n", leftover, w * h);
        envs.push_back(std::move(e));
    };
    for (const std::string& path : a.hdris) {
        uint32_t w, h;
        float* rgb = nullptr;
        if (rt_hdr_load(path.c_str(), &w, &h, &rgb, err, sizeof err) != RT_OK) return die(err, nullptr);
        add_env(w, h, rgb);
        rt_free(rgb);
    }
    if (envs.empty()) {
        std::vector<float> rgb((size_t)2048 * 1024 * 3);
        rt_hdri_synthetic(2048, 1024, rgb.data());
        add_env(2048, 1024, rgb.data());
    }
    if (a.env >= envs.size()) return die("--env is out of range", nullptr);
    std::vector<RtEnvironmentDesc> env_descs;
    for (auto& e : envs) env_descs.push_back(RtEnvironmentDesc{e.w, e.h, e.rgba.data(), e.alias.data()});

    if (a.bvh == "device") {
        // build_bvh (bvh.rs:13-28) on GPU 0: the same node array and slot order as the host builder's
        RtContext* builder = nullptr;
        if (rt_create(0, &builder) != RT_OK) return die("GPU setup failed", rt_last_error(nullptr));
        RtBvhBuildInput input;
        uint32_t node_count = 0, depth = 0;
        float device_ms = 0.f;
        rt_scene_bvh_input(scene, &input);
        if (rt_bvh_build(builder, &input, &node_count, &depth, &device_ms) != RT_OK)
            return die("device BVH build failed", rt_last_error(builder));
        RtBvhNode* nodes = nullptr;
        RtPrimitiveInfo* primitives = nullptr;
        if (rt_scene_bvh_storage(scene, node_count, &nodes, &primitives) != RT_OK ||
            rt_bvh_build_fetch(builder, nodes, primitives) != RT_OK || rt_scene_bvh_commit(scene, depth) != RT_OK)
            return die("device BVH build failed", rt_last_error(builder));
        std::fprintf(stderr, "BVH built on the device in %.2f ms\n", device_ms);
        rt_destroy(builder);
    }
    RtSceneDesc desc;
    rt_scene_desc(scene, &desc);
    std::fprintf(stderr, "tree depth: %u\ntree node count: %u\n", rt_scene_bvh_depth(scene),
                 desc.bvh_node_count);

    std::vector<int> devices;
    for (int i = 0; i < a.gpus; i++) devices.push_back(i);
    RtGroup* group = nullptr;
    if (rt_group_create(a.gpus, devices.data(), &group) != RT_OK) return die("GPU setup failed", rt_last_error(nullptr));
    for (int r = 0; r < a.gpus; r++) {
        RtContext* ctx = rt_group_context(group, r);
        if (rt_upload_scene(ctx, &desc) != RT_OK || rt_upload_environments(ctx, env_descs.data(), (uint32_t)env_descs.size()) != RT_OK ||
            rt_set_camera(ctx, &camera) != RT_OK || rt_resize(ctx, a.width, a.height) != RT_OK)
            return die("GPU upload failed", rt_last_error(ctx));
    }
    RtContext* root = rt_group_context(group, 0);

    if (!a.first_hit_out.empty()) {
        std::vector<uint32_t> slots((size_t)a.width * a.height);
        if (rt_first_hit(root, a.sample_begin, slots.data(), nullptr) != RT_OK)
            return die("first-hit pass failed", rt_last_error(root));
        FILE* f = std::fopen(a.first_hit_out.c_str(), "wb");
        if (!f) return die("cannot write", a.first_hit_out.c_str());
        std::fwrite(slots.data(), 4, slots.size(), f);
        std::fclose(f);
    }

    rt_group_sync(group);
    auto t0 = std::chrono::steady_clock::now();
    if (rt_group_render(group, a.sample_begin, a.spp, a.bounces, a.env) != RT_OK ||
        rt_group_reduce(group, 0) != RT_OK || rt_group_sync(group) != RT_OK)
        return die("render failed", rt_last_error(root));
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    uint64_t rays = 0, paths = 0;
    for (int r = 0; r < a.gpus; r++) {
        RtCounters c;
        rt_get_counters(rt_group_context(group, r), &c);
        rays += c.rays + c.shadow_rays;
        paths += c.paths;
    }
    std::fprintf(stderr, "rendered %ux%u, %u spp, %u bounces on %d GPU(s) in %.3f s: %.1f Mrays/s, %.1f Msamples/s\n",
                 a.width, a.height, a.spp, a.bounces, a.gpus, secs, rays / secs / 1e6, paths / secs / 1e6);

    if (!a.out.empty()) {
        std::vector<uint8_t> rgb8((size_t)a.width * a.height * 3);
        if (rt_tonemap_rgb8(root, a.spp, rgb8.data()) != RT_OK) return die("tone-map failed", rt_last_error(root));
        if (rt_png_write_rgb8(a.out.c_str(), a.width, a.height, rgb8.data()) != RT_OK)
            return die("cannot write", a.out.c_str());
    }
    if (!a.pfm.empty()) {
        std::vector<float> accum((size_t)a.width * a.height * 4);
        if (rt_read_accum(root, accum.data()) != RT_OK) return die("read-back failed", rt_last_error(root));
        if (rt_pfm_write_rgb(a.pfm.c_str(), a.width, a.height, accum.data(), 4) != RT_OK)
            return die("cannot write", a.pfm.c_str());
    }
    rt_group_destroy(group);
    rt_scene_free(scene);
    return 0;
}
