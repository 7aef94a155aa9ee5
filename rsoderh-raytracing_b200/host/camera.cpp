// camera.cpp — camera model of src/camera.rs:14-119.
#include <cmath>
#include <cstring>

#include "scene.hpp"

namespace rt {
namespace {

struct Mat3 {
    Vec3 x_axis, y_axis, z_axis;  // columns
};

// glam 0.30.9 Mat3::from_axis_angle for a unit axis.
Mat3 from_axis_angle(Vec3 axis, float angle) {
    float s = std::sin(angle), c = std::cos(angle);
    float xsin = axis.x * s, ysin = axis.y * s, zsin = axis.z * s;
    float x = axis.x, y = axis.y, z = axis.z;
    float x2 = x * x, y2 = y * y, z2 = z * z;
    float omc = 1.0f - c;
    float xyomc = x * y * omc, xzomc = x * z * omc, yzomc = y * z * omc;
    Mat3 m;
    m.x_axis = v3(x2 * omc + c, xyomc + zsin, xzomc - ysin);
    m.y_axis = v3(xyomc - zsin, y2 * omc + c, yzomc + xsin);
    m.z_axis = v3(xzomc + ysin, yzomc - xsin, z2 * omc + c);
    return m;
}

// glam Mat3 * Vec3: x_axis * v.x + y_axis * v.y + z_axis * v.z.
Vec3 mul(const Mat3& m, Vec3 v) { return (m.x_axis * v.x + m.y_axis * v.y) + m.z_axis * v.z; }
Mat3 mul(const Mat3& a, const Mat3& b) { return {mul(a, b.x_axis), mul(a, b.y_axis), mul(a, b.z_axis)}; }

const char kB64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";

}  // namespace

void camera_uniform(const Camera& c, RtCameraUniform* out) {
    // camera.rs:26-28: R_y(yaw) * R_x(pitch)
    Mat3 rot = mul(from_axis_angle(v3(0, 1, 0), c.yaw), from_axis_angle(v3(1, 0, 0), c.pitch));
    std::memset(out, 0, sizeof *out);
    out->pos[0] = c.pos.x; out->pos[1] = c.pos.y; out->pos[2] = c.pos.z;
    const Vec3 cols[3] = {rot.x_axis, rot.y_axis, rot.z_axis};
    for (int i = 0; i < 3; i++) {
        out->rot[i][0] = cols[i].x;
        out->rot[i][1] = cols[i].y;
        out->rot[i][2] = cols[i].z;
    }
    out->fov_y = c.fov_y;
}

std::string camera_serialize(const Camera& c) {
    // camera.rs:30-49: 24 bytes = pos.xyz, yaw, pitch, fov_y as little-endian f32.
    float f[6] = {c.pos.x, c.pos.y, c.pos.z, c.yaw, c.pitch, c.fov_y};
    uint8_t data[24];
    std::memcpy(data, f, 24);
    std::string out;
    for (int i = 0; i < 24; i += 3) {
        uint32_t n = (data[i] << 16) | (data[i + 1] << 8) | data[i + 2];
        out += kB64[(n >> 18) & 63];
        out += kB64[(n >> 12) & 63];
        out += kB64[(n >> 6) & 63];
        out += kB64[n & 63];
    }
    return out;
}

Camera camera_deserialize(const std::string& encoded) {
    // base64 STANDARD engine: padded, canonical.
    std::vector<uint8_t> data;
    uint32_t acc = 0;
    int bits = 0;
    size_t pad = 0;
    for (size_t i = 0; i < encoded.size(); i++) {
        char ch = encoded[i];
        if (ch == '=') {
            pad++;
            continue;
        }
        const char* pos = std::strchr(kB64, ch);
        if (pad || !pos || ch == '\0')
            throw SceneError(RT_ERR_PARSE, "Invalid symbol " + std::to_string((int)(uint8_t)ch) +
                                               ", offset " + std::to_string(i) + ".");
        acc = (acc << 6) | (uint32_t)(pos - kB64);
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            data.push_back((uint8_t)((acc >> bits) & 0xFF));
        }
    }
    if ((encoded.size() % 4) != 0 || pad > 2) throw SceneError(RT_ERR_PARSE, "Invalid padding");
    if (data.size() != 24)  // camera.rs:55-60
        throw SceneError(RT_ERR_PARSE, "Couldn't deserialize camera: binary data (" +
                                           std::to_string(data.size()) + " bytes) not 24 bytes");
    float f[6];
    std::memcpy(f, data.data(), 24);
    Camera c;
    c.pos = v3(f[0], f[1], f[2]);
    c.yaw = f[3];
    c.pitch = f[4];
    c.fov_y = f[5];
    return c;
}

}  // namespace rt
