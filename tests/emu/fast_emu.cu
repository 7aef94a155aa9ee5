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
    v.flat = e->layout.flat.data();
    v.big_len = e->layout.big_len.empty() ? nullptr : e->layout.big_len.data();
    v.tri_normals = nullptr;
    v.materials = nullptr;
    e->sphere_count = s->sphere_count;
    return e;
}

void emu_scene_free(void* h) { delete (EmuScene*)h; }

int emu_scene_hoisted(void* h) { return (int)((EmuScene*)h)->layout.view.hoisted; }

// Leaves of the flat walk program (0: the scene has none and emu_trace_flat must not be used).
int emu_scene_flat_leaves(void* h) { return (int)((EmuScene*)h)->layout.view.flat_leaves; }

// Structural check of the effective tree: walks every N2 record from the root and verifies
// that each interior box encloses its children's boxes, that leaf refs are well formed and
// that every triangle slot is covered exactly once. out = {interior records, walked leaves,
// triangle slots covered, max depth, pre-pass leaves, boxes, prims, hoisted}. Returns 0 when
// consistent, otherwise a code naming the first violated property.
int emu_check_layout(void* h, const RtSceneDesc* s, uint32_t out[8]) {
    using namespace rtf;
    EmuScene* e = (EmuScene*)h;
    const FastScene& sc = e->layout.view;
    for (int i = 0; i < 8; i++) out[i] = 0;
    out[4] = sc.pre_leaves;
    out[5] = sc.pre_boxes;
    out[6] = sc.pre_prims;
    out[7] = sc.hoisted;
    std::vector<uint8_t> seen(s->primitive_count, 0);
    struct Item {
        uint32_t ref, depth;
        float mn[3], mx[3];
    };
    std::vector<Item> todo;
    if (sc.root_ref != kNone) {
        Item r;
        r.ref = sc.root_ref;
        r.depth = 0;
        for (int a = 0; a < 3; a++) {
            r.mn[a] = sc.root_min[a];
            r.mx[a] = sc.root_max[a];
        }
        todo.push_back(r);
    }
    while (!todo.empty()) {
        Item it = todo.back();
        todo.pop_back();
        if (it.depth > out[3]) out[3] = it.depth;
        if (it.ref & kLeafBit) {
            const uint32_t first = it.ref & kLeafSlotMask;
            uint32_t len = (it.ref >> kLeafLenShift) & 31u;
            if (len == kLeafLenEscape) {
                if (!sc.big_len) return 1;
                len = sc.big_len[first];
            }
            if (len == 0 || (uint64_t)first + len > s->primitive_count) return 2;
            out[1]++;
            for (uint32_t k = first; k < first + len; k++) {
                if (seen[k]) return 3;  // a slot reachable twice
                seen[k] = 1;
                if (s->primitives[k].primitive_type == 2u) out[2]++;
            }
            continue;
        }
        if ((size_t)it.ref * 4 + 3 >= e->layout.n2.size()) return 4;
        out[0]++;
        const float4* q = sc.n2 + 4u * (size_t)it.ref;
        const float4 boxes[2][2] = {{q[0], q[1]}, {q[2], q[3]}};
        const uint32_t refs[2] = {rtd::f_as_u(q[0].w), rtd::f_as_u(q[1].w)};
        if (rtd::f_as_u(q[3].w) > 2u) return 5;  // split axis
        for (int c = 0; c < 2; c++) {
            if (refs[c] == kNone) return 6;  // the effective tree has no empty children
            Item ch;
            ch.ref = refs[c];
            ch.depth = it.depth + 1;
            const float mn[3] = {boxes[c][0].x, boxes[c][0].y, boxes[c][0].z};
            const float mx[3] = {boxes[c][1].x, boxes[c][1].y, boxes[c][1].z};
            for (int a = 0; a < 3; a++) {
                if (mn[a] < it.mn[a] || mx[a] > it.mx[a]) return 7;  // child sticks out of its parent
                ch.mn[a] = mn[a];
                ch.mx[a] = mx[a];
            }
            todo.push_back(ch);
        }
    }
    // every triangle must be walked; analytic primitives are walked iff not hoisted
    for (uint32_t k = 0; k < s->primitive_count; k++) {
        const bool tri = s->primitives[k].primitive_type == 2u;
        if (tri && !seen[k]) return 8;
        if (!tri && !sc.hoisted && !seen[k]) return 9;
    }
    return 0;
}

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

// The same rays through the flat walk (trace_flat: leaf-box mask + per-octant visiting order).
void emu_trace_flat(void* h, uint32_t n, const float* origins, const float* dirs, int any_hit,
                    uint32_t* code_out, float* t_out, float* uv_out) {
    EmuScene* e = (EmuScene*)h;
    const rtf::FastScene& sc = e->layout.view;
    for (uint32_t i = 0; i < n; i++) {
        rtf::FHit hit;
        rtd::V3 o = rtd::v3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        rtd::V3 d = rtd::v3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
        rtf::trace_flat(sc, sc.pre, sc.flat, o, d, any_hit != 0, hit);
        code_out[i] = encode(sc, e->sphere_count, hit.rec);
        t_out[i] = hit.rec != rtf::kNone ? hit.t : 0.0f;
        if (uv_out) {
            uv_out[2 * i] = hit.u;
            uv_out[2 * i + 1] = hit.v;
        }
    }
}

}  // extern "C"
