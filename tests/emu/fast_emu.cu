// fast_emu.cu — TEST INFRASTRUCTURE. Compiles the production traversal of
// rsoderh-raytracing_b200/csrc/rt_fast.cuh (trace_fast and the layout builder) for the
// HOST, so the CPU test suite can check its decisions — N2 walk, hoisted analytic
// pre-pass, tie-breaks, fallback — against the oracle on arbitrary rays without a GPU.
// Nothing in the product links this file.
#define RT_HOST_EMU 1
#include "../../rsoderh-raytracing_b200/csrc/rt_fast.cuh"
#include "../../rsoderh-raytracing_b200/csrc/rt_fast_host.hpp"

namespace {

struct EmuScene {
    rtf::FastLayout layout;
    std::vector<float4> tri_normals;
    uint32_t sphere_count = 0;
};

uint32_t encode(const rtf::FastScene& sc, uint32_t sphere_count, uint32_t rec) {
    using namespace rtf;
    if (rec == kNone) return kNone;
    if ((rec & kRecFallback) == kRecFallback) {
        uint32_t index = rec & kRecIndexMask;
        return index < sphere_count ? (0x80000000u | index) : (0x90000000u | (index - sphere_count));
    }
    return hit_slot(sc, rec);
}

}  // namespace

extern "C" {

void* emu_scene_create(const RtSceneDesc* s) {
    EmuScene* e = new EmuScene();
    rtf::build_fast_layout(*s, e->layout);
    rtf::FastScene& v = e->layout.view;
    v.n2 = e->layout.n2.data();
    v.tri = e->layout.tri.data();
    v.ana = e->layout.ana.data();
    v.pre = e->layout.pre.data();
    v.ana_paths = e->layout.ana_paths.data();
    v.big_len = e->layout.big_len.empty() ? nullptr : e->layout.big_len.data();
    v.tri_normals = nullptr;
    v.materials = nullptr;
    e->sphere_count = s->sphere_count;
    return e;
}

void emu_scene_free(void* h) { delete (EmuScene*)h; }

int emu_scene_hoisted(void* h) { return (int)((EmuScene*)h)->layout.view.hoisted; }

// Traces n rays (origins/directions are xyz triples). code_out uses rt_first_hit's encoding
// (slot, fallback code or RT_MISS_SLOT); for any_hit rays only "hit or not" is meaningful.
void emu_trace(void* h, uint32_t n, const float* origins, const float* dirs, int any_hit,
               uint32_t* code_out, float* t_out, float* uv_out) {
    EmuScene* e = (EmuScene*)h;
    const rtf::FastScene& sc = e->layout.view;
    uint2 stack[rtf::kStackCap];
    for (uint32_t i = 0; i < n; i++) {
        rtf::FHit hit;
        rtd::V3 o = rtd::v3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        rtd::V3 d = rtd::v3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
        rtf::trace_fast(sc, sc.pre, o, d, any_hit != 0, hit, stack);
        code_out[i] = encode(sc, e->sphere_count, hit.rec);
        t_out[i] = hit.rec != rtf::kNone ? hit.t : 0.0f;
        if (uv_out) {
            uv_out[2 * i] = hit.u;
            uv_out[2 * i + 1] = hit.v;
        }
    }
}

}  // extern "C"
