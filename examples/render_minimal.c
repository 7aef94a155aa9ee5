/* render_minimal.c — the C ABI from plain C99: load a scene, upload it, render, write a PNG.
 *
 *   gcc -std=c99 -Iinclude examples/render_minimal.c -Lrsoderh-raytracing_b200 -lrt_host -lrt_backend \
 *       -Wl,-rpath,$PWD/rsoderh-raytracing_b200 -o render_minimal
 *   ./render_minimal assets/scenes/house.toml out.png 640 360 16
 *
 * This is the call sequence a host in any language makes (INTEGRATION.md shows the same in Rust):
 * rt_scene_load_toml -> rt_scene_desc -> rt_create -> rt_upload_scene / rt_upload_environments /
 * rt_set_camera / rt_resize -> rt_render -> rt_tonemap_rgb8 -> rt_png_write_rgb8.
 */
#include <stdio.h>
#include <stdlib.h>

#include "rt_backend.h"
#include "rt_host.h"

static int fail(const char* what, const char* detail) {
    fprintf(stderr, "%s: %s\n", what, detail ? detail : "");
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 3) return fail("usage", "render_minimal <scene.toml> <out.png> [width height spp]");
    const uint32_t width = argc > 3 ? (uint32_t)atoi(argv[3]) : 640u;
    const uint32_t height = argc > 4 ? (uint32_t)atoi(argv[4]) : 360u;
    const uint32_t spp = argc > 5 ? (uint32_t)atoi(argv[5]) : 16u;
    char err[512] = {0};

    RtHostScene* scene = NULL;
    if (rt_scene_load_toml(argv[1], &scene, err, sizeof err) != RT_OK) return fail("scene", err);
    RtSceneDesc desc;
    rt_scene_desc(scene, &desc);
    float pos[3], yaw, pitch, fov;
    rt_scene_camera(scene, pos, &yaw, &pitch, &fov);
    RtCameraUniform camera;
    rt_camera_uniform(pos, yaw, pitch, fov, &camera);

    /* one environment: the synthetic sky, its RGBA texels and its alias table */
    const uint32_t ew = 512, eh = 256;
    float* rgb = (float*)malloc((size_t)ew * eh * 3 * sizeof(float));
    float* rgba = (float*)malloc((size_t)ew * eh * 4 * sizeof(float));
    RtAliasEntry* alias = (RtAliasEntry*)malloc((size_t)ew * eh * sizeof(RtAliasEntry));
    uint32_t leftover = 0;
    rt_hdri_synthetic(ew, eh, rgb);
    rt_rgb_to_rgba(ew * eh, rgb, rgba);
    rt_alias_build(ew, eh, rgb, alias, &leftover);
    RtEnvironmentDesc env;
    env.width = ew;
    env.height = eh;
    env.rgba = rgba;
    env.alias = alias;

    RtContext* ctx = NULL;
    if (rt_create(0, &ctx) != RT_OK) return fail("rt_create", rt_last_error(NULL));
    if (rt_upload_scene(ctx, &desc) != RT_OK || rt_upload_environments(ctx, &env, 1) != RT_OK ||
        rt_set_camera(ctx, &camera) != RT_OK || rt_resize(ctx, width, height) != RT_OK)
        return fail("upload", rt_last_error(ctx));
    if (rt_render(ctx, 0, spp, 8, 0) != RT_OK || rt_sync(ctx) != RT_OK) return fail("render", rt_last_error(ctx));

    uint8_t* rgb8 = (uint8_t*)malloc((size_t)width * height * 3);
    if (rt_tonemap_rgb8(ctx, spp, rgb8) != RT_OK) return fail("tone-map", rt_last_error(ctx));
    if (rt_png_write_rgb8(argv[2], width, height, rgb8) != RT_OK) return fail("png", argv[2]);
    RtCounters counters;
    rt_get_counters(ctx, &counters);
    printf("%ux%u, %u spp: %llu rays\n", width, height, spp,
           (unsigned long long)(counters.rays + counters.shadow_rays));

    rt_destroy(ctx);
    rt_scene_free(scene);
    free(rgb8);
    free(alias);
    free(rgba);
    free(rgb);
    return 0;
}
