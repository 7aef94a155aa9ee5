// obj.cpp — OBJ mesh loading with the behaviour the reference gets from
// wavefront_obj 11.0.0 + Mesh::load (src/mesh.rs:24-81):
//   * coordinates are read as f64 and cast to f32 (mesh.rs:24-26);
//   * `f` indices are 1-based and global to the file (wavefront_obj rebases them per
//     `o` object and mesh.rs:37-38,53-64 adds the per-object offsets back, so the net
//     effect is the global 0-based index);
//   * a face with n > 3 corners is fan-triangulated anchored at its LAST corner:
//     corners c0..c(n-1) give Triangle(c(n-1), c(i), c(i+1)), i = 0..n-3, so a plain
//     triangle `a b c` becomes (c, a, b). This restates the crate from memory — the
//     crate is not vendored in the reference tree, so the convention is unpinned
//     (see DESIGN.md "parity unpinned");
//   * points / lines (faces with < 3 corners, `p`, `l`) are not triangles and are
//     dropped by the filter_map at mesh.rs:49-72;
//   * a face corner without a normal index is an error (mesh.rs:60-64 panics with
//     "Object must include baked normals");
//   * vt, usemtl, mtllib, s, g are ignored.
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "scene.hpp"

namespace rt {
namespace {

struct Corner {
    long v = 0, t = 0, n = 0;
    bool has_t = false, has_n = false;
};

// Splits on ASCII whitespace.
void split_ws(const std::string& line, std::vector<std::string>& out) {
    out.clear();
    size_t i = 0, n = line.size();
    while (i < n) {
        while (i < n && (line[i] == ' ' || line[i] == '\t' || line[i] == '\r')) i++;
        size_t j = i;
        while (j < n && !(line[j] == ' ' || line[j] == '\t' || line[j] == '\r')) j++;
        if (j > i) out.emplace_back(line, i, j - i);
        i = j;
    }
}

double parse_f64(const std::string& tok, int line_no) {
    char* end = nullptr;
    errno = 0;
    double v = std::strtod(tok.c_str(), &end);
    if (end == tok.c_str() || *end != '\0')
        throw SceneError(RT_ERR_PARSE, "OBJ line " + std::to_string(line_no) +
                                           ": invalid number '" + tok + "'");
    return v;
}

long parse_index(const std::string& tok, int line_no) {
    char* end = nullptr;
    long v = std::strtol(tok.c_str(), &end, 10);
    if (end == tok.c_str() || *end != '\0' || v == 0)
        throw SceneError(RT_ERR_PARSE, "OBJ line " + std::to_string(line_no) +
                                           ": invalid index '" + tok + "'");
    return v;
}

Corner parse_corner(const std::string& tok, int line_no) {
    Corner c;
    size_t s1 = tok.find('/');
    if (s1 == std::string::npos) {
        c.v = parse_index(tok, line_no);
        return c;
    }
    c.v = parse_index(tok.substr(0, s1), line_no);
    size_t s2 = tok.find('/', s1 + 1);
    if (s2 == std::string::npos) {
        std::string t = tok.substr(s1 + 1);
        if (!t.empty()) {
            c.t = parse_index(t, line_no);
            c.has_t = true;
        }
        return c;
    }
    std::string t = tok.substr(s1 + 1, s2 - s1 - 1);
    if (!t.empty()) {
        c.t = parse_index(t, line_no);
        c.has_t = true;
    }
    std::string n = tok.substr(s2 + 1);
    if (!n.empty()) {
        c.n = parse_index(n, line_no);
        c.has_n = true;
    }
    return c;
}

uint32_t resolve(long idx, size_t count, const char* what, int line_no) {
    long zero_based = idx > 0 ? idx - 1 : (long)count + idx;  // negative = relative to end
    if (zero_based < 0 || (size_t)zero_based >= count)
        throw SceneError(RT_ERR_PARSE, "OBJ line " + std::to_string(line_no) + ": " + what +
                                           " index " + std::to_string(idx) + " out of range");
    return (uint32_t)zero_based;
}

}  // namespace

Mesh load_obj_mesh(const std::string& text, uint32_t material_id) {
    Mesh mesh;
    struct Face {
        std::vector<Corner> corners;
        int line_no;
    };
    std::vector<Face> faces;
    std::vector<std::string> tok;

    std::istringstream in(text);
    std::string line;
    int line_no = 0;
    while (std::getline(in, line)) {
        line_no++;
        size_t hash = line.find('#');
        if (hash != std::string::npos) line.erase(hash);
        split_ws(line, tok);
        if (tok.empty()) continue;
        const std::string& kw = tok[0];
        if (kw == "v" || kw == "vn") {
            if (tok.size() < 4)
                throw SceneError(RT_ERR_PARSE, "OBJ line " + std::to_string(line_no) + ": '" + kw +
                                                   "' needs three coordinates");
            Vec3 p = v3((float)parse_f64(tok[1], line_no), (float)parse_f64(tok[2], line_no),
                        (float)parse_f64(tok[3], line_no));
            (kw == "v" ? mesh.vertices : mesh.normals).push_back(p);
        } else if (kw == "f") {
            Face f;
            f.line_no = line_no;
            for (size_t i = 1; i < tok.size(); i++) f.corners.push_back(parse_corner(tok[i], line_no));
            faces.push_back(std::move(f));
        }
        // vt, o, g, s, usemtl, mtllib, l, p: nothing the render path consumes.
    }

    // Indices may legally refer to vertices declared later in the file, so resolve
    // after the whole file is read (wavefront_obj also resolves after parsing).
    for (const Face& f : faces) {
        size_t n = f.corners.size();
        if (n < 3) continue;  // Point / Line primitives are filtered out (mesh.rs:49-72)
        auto emit = [&](const Corner& a, const Corner& b, const Corner& c) {
            if (!a.has_n || !b.has_n || !c.has_n)
                throw SceneError(RT_ERR_SCENE, "Object must include baked normals");
            MeshTriangle t;
            t.v[0] = resolve(a.v, mesh.vertices.size(), "vertex", f.line_no);
            t.v[1] = resolve(b.v, mesh.vertices.size(), "vertex", f.line_no);
            t.v[2] = resolve(c.v, mesh.vertices.size(), "vertex", f.line_no);
            t.n[0] = resolve(a.n, mesh.normals.size(), "normal", f.line_no);
            t.n[1] = resolve(b.n, mesh.normals.size(), "normal", f.line_no);
            t.n[2] = resolve(c.n, mesh.normals.size(), "normal", f.line_no);
            t.material_id = material_id;
            mesh.triangles.push_back(t);
        };
        const Corner& last = f.corners[n - 1];
        for (size_t i = 0; i + 2 < n; i++) emit(last, f.corners[i], f.corners[i + 1]);
    }
    return mesh;
}

std::string read_text_file(const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error(std::string(std::strerror(errno)) + " (os error " +
                                      std::to_string(errno) + ")");
    std::ostringstream ss;
    ss << in.rdbuf();
    return ss.str();
}

}  // namespace rt
