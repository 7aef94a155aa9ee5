// env.cpp — environment (HDRI) inputs: Radiance RGBE decode/encode, the synthetic sky
// that stands in for the reference's missing Poly Haven files, and the luminance alias
// table of src/environments.rs:87-186.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "scene.hpp"

namespace rt {
namespace {

// image 0.25.9 `Rgbe8Pixel::to_hdr`: e == 0 -> black, else mantissa * 2^(e - 136).
inline void rgbe_to_float(const uint8_t p[4], float* out) {
    if (p[3] == 0) {
        out[0] = out[1] = out[2] = 0.f;
        return;
    }
    float scale = std::ldexp(1.0f, (int)p[3] - (128 + 8));
    out[0] = (float)p[0] * scale;
    out[1] = (float)p[1] * scale;
    out[2] = (float)p[2] * scale;
}

// Standard Radiance float -> RGBE (Ward): exponent from the largest component.
inline void float_to_rgbe(const float* rgb, uint8_t out[4]) {
    float v = rgb[0];
    if (rgb[1] > v) v = rgb[1];
    if (rgb[2] > v) v = rgb[2];
    if (v < 1e-32f) {
        out[0] = out[1] = out[2] = out[3] = 0;
        return;
    }
    int e = 0;
    float m = std::frexp(v, &e);  // v = m * 2^e, m in [0.5, 1)
    float scale = m * 256.0f / v;
    out[0] = (uint8_t)(rgb[0] * scale);
    out[1] = (uint8_t)(rgb[1] * scale);
    out[2] = (uint8_t)(rgb[2] * scale);
    out[3] = (uint8_t)(e + 128);
}

[[noreturn]] void hdr_fail(const std::string& msg) {
    throw SceneError(RT_ERR_PARSE, "Radiance HDR decode: " + msg);
}

}  // namespace

void hdr_decode(const uint8_t* bytes, size_t len, uint32_t* w_out, uint32_t* h_out,
                std::vector<float>& rgb) {
    size_t p = 0;
    auto read_line = [&](std::string& line) -> bool {
        line.clear();
        if (p >= len) return false;
        while (p < len && bytes[p] != '\n') line += (char)bytes[p++];
        if (p < len) p++;
        return true;
    };
    std::string line;
    if (!read_line(line) || (line.rfind("#?RADIANCE", 0) != 0 && line.rfind("#?RGBE", 0) != 0))
        hdr_fail("missing #?RADIANCE signature");
    bool format_ok = false;
    for (;;) {
        if (!read_line(line)) hdr_fail("unexpected end of header");
        if (line.empty() || line == "\r") break;
        if (line.rfind("FORMAT=", 0) == 0) {
            if (line.find("32-bit_rle_rgbe") == std::string::npos)
                hdr_fail("unsupported FORMAT (only 32-bit_rle_rgbe)");
            format_ok = true;
        }
    }
    (void)format_ok;  // FORMAT is optional in files found in the wild
    if (!read_line(line)) hdr_fail("missing resolution line");
    int h = 0, w = 0;
    if (std::sscanf(line.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0)
        hdr_fail("unsupported resolution line '" + line + "' (only -Y H +X W)");

    rgb.assign((size_t)w * h * 3, 0.f);
    std::vector<uint8_t> scan((size_t)w * 4);
    for (int y = 0; y < h; y++) {
        bool rle = false;
        if (w >= 8 && w < 32768 && p + 4 <= len && bytes[p] == 2 && bytes[p + 1] == 2 &&
            (bytes[p + 2] & 0x80) == 0) {
            int sw = (bytes[p + 2] << 8) | bytes[p + 3];
            if (sw == w) rle = true;
        }
        if (rle) {
            p += 4;
            // New-style RLE: the four channels of the scanline are stored separately.
            for (int ch = 0; ch < 4; ch++) {
                int x = 0;
                while (x < w) {
                    if (p >= len) hdr_fail("truncated RLE scanline");
                    int count = bytes[p++];
                    if (count > 128) {
                        count -= 128;
                        if (p >= len || x + count > w) hdr_fail("bad RLE run");
                        uint8_t v = bytes[p++];
                        for (int i = 0; i < count; i++) scan[(size_t)(x++) * 4 + ch] = v;
                    } else {
                        if (count == 0 || p + count > len || x + count > w)
                            hdr_fail("bad RLE literal run");
                        for (int i = 0; i < count; i++) scan[(size_t)(x++) * 4 + ch] = bytes[p++];
                    }
                }
            }
        } else {
            if (p + (size_t)w * 4 > len) hdr_fail("truncated flat scanline");
            std::memcpy(scan.data(), bytes + p, (size_t)w * 4);
            p += (size_t)w * 4;
        }
        float* row = rgb.data() + (size_t)y * w * 3;
        for (int x = 0; x < w; x++) rgbe_to_float(&scan[(size_t)x * 4], row + (size_t)x * 3);
    }
    *w_out = (uint32_t)w;
    *h_out = (uint32_t)h;
}

std::vector<uint8_t> hdr_encode_flat(uint32_t w, uint32_t h, const float* rgb) {
    char header[128];
    int n = std::snprintf(header, sizeof header,
                          "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %u +X %u\n", h, w);
    std::vector<uint8_t> out(header, header + n);
    out.resize((size_t)n + (size_t)w * h * 4);
    uint8_t* px = out.data() + n;
    for (size_t i = 0; i < (size_t)w * h; i++) float_to_rgbe(rgb + i * 3, px + i * 4);
    return out;
}

void hdri_synthetic(uint32_t w, uint32_t h, float* rgb) {
    // Pixel-space construction (no trig) so every platform produces the same bytes:
    //   upper half: blue-ish gradient brightening towards the horizon;
    //   lower half: dim ground;
    //   sun: a disc of radius h/100 px centred at (0.70 w, 0.22 h), radiance ~1.5e3, with a
    //        soft halo — small and very bright, to exercise the alias table and MIS.
    const float sun_cx = 0.70f * (float)w, sun_cy = 0.22f * (float)h;
    const float sun_r = (float)h / 100.0f;
    for (uint32_t y = 0; y < h; y++) {
        float v = ((float)y + 0.5f) / (float)h;
        for (uint32_t x = 0; x < w; x++) {
            float u = ((float)x + 0.5f) / (float)w;
            float c[3];
            if (v < 0.5f) {
                float t = v * 2.0f;  // 0 zenith .. 1 horizon
                c[0] = 0.25f + 0.55f * t;
                c[1] = 0.45f + 0.45f * t;
                c[2] = 0.95f + 0.05f * t;
                // Slow azimuthal variation so u matters for the diffuse sky too.
                float az = u < 0.5f ? u * 2.0f : (1.0f - u) * 2.0f;
                c[0] *= 0.8f + 0.4f * az;
                c[1] *= 0.85f + 0.3f * az;
            } else {
                float t = (v - 0.5f) * 2.0f;
                c[0] = 0.20f - 0.08f * t;
                c[1] = 0.18f - 0.07f * t;
                c[2] = 0.15f - 0.06f * t;
            }
            float dx = ((float)x + 0.5f) - sun_cx, dy = ((float)y + 0.5f) - sun_cy;
            float d2 = dx * dx + dy * dy;
            if (d2 < sun_r * sun_r) {
                c[0] += 1500.0f; c[1] += 1350.0f; c[2] += 1100.0f;
            } else if (d2 < 16.0f * sun_r * sun_r) {
                float f = sun_r * sun_r / d2;  // 1 .. 1/16
                c[0] += 3.0f * f; c[1] += 2.5f * f; c[2] += 1.75f * f;
            }
            // Quantise through RGBE so the data equals what a .hdr round trip yields.
            uint8_t q[4];
            float_to_rgbe(c, q);
            rgbe_to_float(q, rgb + ((size_t)y * w + x) * 3);
        }
    }
}

// Runs fn(t, begin, end) over [0, n) cut into one contiguous range per thread.
template <typename F>
static void parallel_ranges(size_t n, int threads, F fn) {
    if (threads <= 1 || n < (size_t)1 << 14) {
        fn(0, (size_t)0, n);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        const size_t lo = n * (size_t)t / (size_t)threads, hi = n * (size_t)(t + 1) / (size_t)threads;
        pool.emplace_back([=] { fn(t, lo, hi); });
    }
    for (auto& th : pool) th.join();
}

// AliasTable::build_by_luminance (environments.rs:97-186), byte-identical to the literal loop of
// alias_build_serial below (tests/test_abi_and_env.py), organised for speed:
//  * weights, normalisation and the small / large partition run on all host threads;
//  * the f32 sum stays one sequential chain (environments.rs:110 — its rounding is the contract);
//  * the pairing loop keeps no stacks. Both stacks start as ascending index lists and are popped
//    from the end, and the only element ever pushed is the large entry just popped — onto the top
//    of one of the two stacks, from where the next iteration takes it again. So a cursor into each
//    list, the current large entry with its remaining weight in a register, and one carried
//    "became small" entry reproduce the stack discipline exactly, with no per-entry residual array
//    and no `assigned` flags: what was never assigned is what is left of the two lists.
uint32_t alias_build(uint32_t w, uint32_t h, const float* rgb, RtAliasEntry* out) {
    const size_t length = (size_t)w * h;
    const float kPi = 3.14159274f;  // std::f32::consts::PI (environments.rs:1)
    int threads = (int)std::thread::hardware_concurrency();
    threads = threads < 1 ? 1 : (threads > 32 ? 32 : threads);
    if (const char* k = std::getenv("RT_HOST_THREADS")) threads = std::atoi(k) > 0 ? std::atoi(k) : 1;

    std::vector<float> row_sin(h);
    for (uint32_t y = 0; y < h; y++) {
        float angle = kPi * ((float)y + 0.5f) / (float)h;
        row_sin[y] = (float)std::sin((double)angle);
    }
    std::unique_ptr<float[]> prob(new float[length]);  // first touched by the threads that fill it
    parallel_ranges(h, threads, [&](int, size_t y0, size_t y1) {
        for (size_t y = y0; y < y1; y++)
            for (uint32_t x = 0; x < w; x++) {
                const float* c = rgb + (y * w + x) * 3;
                float lum = 0.2126f * c[0] + 0.7152f * c[1] + 0.0722f * c[2];  // environments.rs:88-90
                prob[y * w + x] = lum * row_sin[y];
            }
    });
    float weight_sum = 0.f;  // sequential f32 sum (environments.rs:110)
    for (size_t i = 0; i < length; i++) weight_sum += prob[i];
    const float length_f = (float)length;

    // normalise; count the entries below / not below 1 per range, then write the two ascending lists
    std::vector<size_t> n_small((size_t)threads + 1, 0), n_large((size_t)threads + 1, 0);
    parallel_ranges(length, threads, [&](int t, size_t lo, size_t hi) {
        size_t ns = 0;
        for (size_t i = lo; i < hi; i++) {
            prob[i] = prob[i] * length_f / weight_sum;
            ns += prob[i] < 1.f ? 1u : 0u;
            out[i].pmf = prob[i] / length_f;  // environments.rs:146 (entries never assigned get 1 / N below)
            out[i]._pad = 0;
        }
        n_small[(size_t)t + 1] = ns;
        n_large[(size_t)t + 1] = (hi - lo) - ns;
    });
    for (int t = 0; t < threads; t++) {
        n_small[(size_t)t + 1] += n_small[(size_t)t];
        n_large[(size_t)t + 1] += n_large[(size_t)t];
    }
    const size_t small_count = n_small[(size_t)threads], large_count = n_large[(size_t)threads];
    std::unique_ptr<uint32_t[]> small(new uint32_t[small_count + 1]), large(new uint32_t[large_count + 1]);
    parallel_ranges(length, threads, [&](int t, size_t lo, size_t hi) {
        size_t ks = n_small[(size_t)t], kl = n_large[(size_t)t];
        for (size_t i = lo; i < hi; i++) {
            if (prob[i] < 1.f) small[ks++] = (uint32_t)i;
            else large[kl++] = (uint32_t)i;
        }
    });

    // environments.rs:134-159
    size_t ks = small_count, kl = large_count;
    const uint32_t kNoEntry = 0xFFFFFFFFu;
    uint32_t carried = kNoEntry;  // a large entry that fell below 1: on top of the small stack
    float carried_weight = 0.f;
    uint32_t cur = kNoEntry;      // the large entry on top of the large stack, its weight in a register
    float cur_weight = 0.f;
    uint32_t stranded = kNoEntry;  // popped from the small stack when the large stack was empty
    for (;;) {
        uint32_t s;
        float s_weight;
        if (carried != kNoEntry) {
            s = carried;
            s_weight = carried_weight;
            carried = kNoEntry;
        } else {
            if (ks == 0) break;
            s = small[--ks];
            s_weight = prob[s];
        }
        if (cur == kNoEntry) {
            if (kl == 0) {
                stranded = s;  // environments.rs:136-141: popped before the large stack is checked
                break;
            }
            cur = large[--kl];
            cur_weight = prob[cur];
        }
        out[s].probability = s_weight;
        out[s].alias_index = cur;
        cur_weight -= 1.f - s_weight;
        if (cur_weight < 1.f) {
            carried = cur;
            carried_weight = cur_weight;
            cur = kNoEntry;
        }
    }
    // environments.rs:162-178: everything never assigned keeps itself
    auto keep = [&](uint32_t i) {
        out[i].probability = 1.f;
        out[i].alias_index = i;
        out[i].pmf = 1.f / length_f;
        out[i]._pad = 0;
    };
    uint32_t leftover = 0;
    for (size_t i = 0; i < ks; i++, leftover++) keep(small[i]);
    for (size_t i = 0; i < kl; i++, leftover++) keep(large[i]);
    if (cur != kNoEntry) keep(cur), leftover++;
    if (stranded != kNoEntry) keep(stranded), leftover++;
    return leftover;
}

// The literal statement of the same function (vectors as stacks, residual array, assigned flags):
// what alias_build is checked against.
uint32_t alias_build_serial(uint32_t w, uint32_t h, const float* rgb, RtAliasEntry* out) {
    const size_t length = (size_t)w * h;
    const float kPi = 3.14159274f;  // std::f32::consts::PI (environments.rs:1)
    // environments.rs:98-106: weight = luminance * sin(PI * (y + 0.5) / height).
    // The row sine is taken in f64 and rounded to f32 (i.e. a correctly rounded sinf),
    // which is the platform-independent reading of Rust's f32::sin here.
    std::vector<float> row_sin(h);
    for (uint32_t y = 0; y < h; y++) {
        float angle = kPi * ((float)y + 0.5f) / (float)h;
        row_sin[y] = (float)std::sin((double)angle);
    }
    std::vector<float> prob(length);
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            const float* c = rgb + ((size_t)y * w + x) * 3;
            float lum = 0.2126f * c[0] + 0.7152f * c[1] + 0.0722f * c[2];  // environments.rs:88-90
            prob[(size_t)y * w + x] = lum * row_sin[y];
        }
    float weight_sum = 0.f;  // sequential f32 sum (environments.rs:110)
    for (size_t i = 0; i < length; i++) weight_sum += prob[i];
    const float length_f = (float)length;
    for (size_t i = 0; i < length; i++) prob[i] = prob[i] * length_f / weight_sum;
    std::vector<float> alias_prob(prob);

    // environments.rs:123-132: ascending-index stacks, popped from the end.
    std::vector<uint32_t> small, large;
    for (size_t i = 0; i < length; i++)
        if (prob[i] < 1.f) small.push_back((uint32_t)i);
    for (size_t i = 0; i < length; i++)
        if (prob[i] >= 1.f) large.push_back((uint32_t)i);

    std::vector<uint8_t> assigned(length, 0);
    for (;;) {
        // environments.rs:136-141: the small index is popped BEFORE the large stack is
        // checked, so when the large stack runs dry that small index stays unassigned.
        if (small.empty()) break;
        uint32_t s = small.back();
        small.pop_back();
        if (large.empty()) break;
        uint32_t l = large.back();
        large.pop_back();

        out[s].probability = alias_prob[s];
        out[s].alias_index = l;
        out[s].pmf = prob[s] / length_f;
        out[s]._pad = 0;
        assigned[s] = 1;

        alias_prob[l] -= 1.f - alias_prob[s];
        if (alias_prob[l] < 1.f)
            small.push_back(l);
        else
            large.push_back(l);
    }
    uint32_t leftover = 0;
    for (size_t i = 0; i < length; i++)
        if (!assigned[i]) {  // environments.rs:162-178
            leftover++;
            out[i].probability = 1.f;
            out[i].alias_index = (uint32_t)i;
            out[i].pmf = 1.f / length_f;
            out[i]._pad = 0;
        }
    return leftover;
}

}  // namespace rt
