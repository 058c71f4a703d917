// objparse.cpp -- host-side mesh ingestion ("next" row N2 of SURVEY section 8f).
//
// Replaces the readers of ekzhang/rpt src/io.rs:
//   * load_obj -> parse_obj_point / parse_obj_face (:27-73,151-200): `v`, `vn` and `f` records; faces
//     are fan-triangulated; `a/b/c` and `a//c` corner syntax; 1-based and negative (relative) indices;
//     a corner without a normal index makes the whole triangle flat-shaded (Triangle::from_vertices,
//     src/shape/mesh.rs:24-37); `vt`, `mtllib`, `usemtl` and unknown records are skipped.
//   * load_obj_with_mtl + load_mtl (:83-149,202-258): a `usemtl` that names a different material than
//     the previous one closes the current run of faces into one mesh carrying the previous material;
//     `newmtl` starts from Material::default() (specular red 0.5, src/material.rs:28-32) and
//     Kd / Ns / Ni / d are mapped best-effort (roughness = (2/(Ns+2))^(1/4), index = max(Ni, 1+1e-4),
//     d < 0.8 -> transparent).
//   * load_stl (:260-360): binary when `size == 84 + 50 n` with n the u32 at byte 80, otherwise ASCII
//     when the file starts with "solid "; every corner of a facet gets the facet normal as stored.
// The reference parses line by line through BufReader + split_ascii_whitespace + str::parse; this is
// a single pass over the byte buffer with strtod / strtol.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/rpt_b200.h"

namespace rptb {

namespace {

inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v'; }

// Walks a byte buffer one trimmed, NUL-terminated line at a time (strtod / strtol need a terminator).
struct LineReader {
    const char* p;
    const char* end;
    size_t lineno = 0;
    std::string line;
    LineReader(const char* text, size_t len) : p(text), end(text + len) {}
    // Returns a pointer to the first non-blank byte of the next line, or nullptr at end of input.
    char* next() {
        if (p >= end) return nullptr;
        const char* le = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!le) le = end;
        lineno++;
        line.assign(p, (size_t)(le - p));
        p = le + 1;
        while (!line.empty() && is_ws(line.back())) line.pop_back();
        char* q = &line[0];
        while (is_ws(*q)) q++;
        return q;
    }
    std::string where() const { return "line " + std::to_string(lineno) + ": "; }
};

// Splits off the first whitespace-delimited token of q; returns its length and advances q past it.
inline size_t take_token(char*& q, char*& tok) {
    while (is_ws(*q)) q++;
    tok = q;
    while (*q && !is_ws(*q)) q++;
    return (size_t)(q - tok);
}

inline bool tok_is(const char* tok, size_t tl, const char* name) { return strlen(name) == tl && memcmp(tok, name, tl) == 0; }

inline bool take_doubles(char*& q, double* out, int n) {
    for (int k = 0; k < n; k++) {
        char* e2 = nullptr;
        out[k] = std::strtod(q, &e2);
        if (e2 == q) return false;
        q = e2;
    }
    return true;
}

// The `v` / `vn` pools and the face triangulator shared by load_obj and load_obj_with_mtl.
struct ObjState {
    std::vector<double> verts, norms;
    std::vector<long> vi, ni;

    // parse_obj_face (src/io.rs:163-200): appends 18 doubles per fan triangle to `tris`.
    bool face(char* q, std::vector<double>& tris, std::string& err) {
        vi.clear();
        ni.clear();
        const long nv = (long)(verts.size() / 3), nn = (long)(norms.size() / 3);
        while (true) {
            while (is_ws(*q)) q++;
            if (!*q) break;
            // corner: a[/b[/c]]
            char* e2 = nullptr;
            const long ia = std::strtol(q, &e2, 10);
            if (e2 == q || (*e2 && *e2 != '/' && !is_ws(*e2))) {
                err = "Invalid vertex index";
                return false;
            }
            const long v0 = ia > 0 ? ia - 1 : nv + ia;  // parse_index, src/io.rs:10-18
            if (v0 < 0 || v0 >= nv) {
                err = "vertex index out of range";
                return false;
            }
            vi.push_back(v0);
            q = e2;
            long n0 = -1;
            if (*q == '/') {  // texture index (ignored)
                q++;
                while (*q && *q != '/' && !is_ws(*q)) q++;
                if (*q == '/') {  // normal index
                    q++;
                    const long ic = std::strtol(q, &e2, 10);
                    if (e2 != q && (!*e2 || is_ws(*e2) || *e2 == '/')) {
                        n0 = ic > 0 ? ic - 1 : nn + ic;
                        if (n0 < 0 || n0 >= nn) {
                            err = "normal index out of range";
                            return false;
                        }
                    }
                    q = e2;
                }
            }
            while (*q && !is_ws(*q)) q++;
            ni.push_back(n0);
        }
        for (size_t i = 1; i + 1 < vi.size(); i++) {  // fan: (0, i, i+1)
            const size_t idx[3] = {0, i, i + 1};
            const double* P[3];
            for (int k = 0; k < 3; k++) P[k] = &verts[3 * (size_t)vi[idx[k]]];
            for (int k = 0; k < 3; k++)
                for (int c = 0; c < 3; c++) tris.push_back(P[k][c]);
            if (ni[0] < 0 || ni[i] < 0 || ni[i + 1] < 0) {  // Triangle::from_vertices
                const double d0[3] = {P[1][0] - P[0][0], P[1][1] - P[0][1], P[1][2] - P[0][2]};
                const double d1[3] = {P[2][0] - P[0][0], P[2][1] - P[0][1], P[2][2] - P[0][2]};
                double n[3] = {d0[1] * d1[2] - d0[2] * d1[1], d0[2] * d1[0] - d0[0] * d1[2], d0[0] * d1[1] - d0[1] * d1[0]};
                const double l = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
                for (int k = 0; k < 3; k++)
                    for (int c = 0; c < 3; c++) tris.push_back(n[c] / l);
            } else {
                for (int k = 0; k < 3; k++) {
                    const double* N = &norms[3 * (size_t)ni[idx[k]]];
                    for (int c = 0; c < 3; c++) tris.push_back(N[c]);
                }
            }
        }
        return true;
    }

    // `v` / `vn` record (parse_obj_point, src/io.rs:151-161).
    bool point(char* q, bool normal, std::string& err) {
        double c[3];
        if (!take_doubles(q, c, 3)) {
            err = "Failed to parse vertex in .OBJ";
            return false;
        }
        std::vector<double>& dst = normal ? norms : verts;
        dst.insert(dst.end(), c, c + 3);
        return true;
    }
};

rptb_material default_material() {  // Material::default(), src/material.rs:28-32 + hex_color(0xff0000)
    rptb_material m;
    std::memset(&m, 0, sizeof(m));
    m.color[0] = 1.0;
    m.index = 1.5;
    m.roughness = 0.5;
    return m;
}

}  // namespace

// Returns 0 on success; on failure a negative value and `err` set.  `tris` receives 18 doubles per
// triangle: v1 v2 v3 n1 n2 n3.
int parse_obj_text(const char* text, size_t len, std::vector<double>& tris, std::string& err) {
    ObjState st;
    LineReader rd(text, len);
    tris.clear();
    while (char* q = rd.next()) {
        char* tok;
        const size_t tl = take_token(q, tok);
        if (tl == 0 || *tok == '#') continue;
        bool ok = true;
        if (tok_is(tok, tl, "v")) ok = st.point(q, false, err);
        else if (tok_is(tok, tl, "vn")) ok = st.point(q, true, err);
        else if (tok_is(tok, tl, "f")) ok = st.face(q, tris, err);
        if (!ok) {
            err = rd.where() + err;
            return -1;
        }
    }
    return 0;
}

// load_mtl (src/io.rs:202-258).  `names[i]` owns `mats[i]`; a repeated `newmtl` keeps editing the
// material it already has (HashMap::entry().or_default()).
int parse_mtl_text(const char* text, size_t len, std::vector<std::string>& names, std::vector<rptb_material>& mats,
                   std::string& err) {
    std::unordered_map<std::string, size_t> index;
    long current = -1;
    LineReader rd(text, len);
    names.clear();
    mats.clear();
    while (char* q = rd.next()) {
        char* tok;
        const size_t tl = take_token(q, tok);
        if (tl == 0 || *tok == '#') continue;
        if (tok_is(tok, tl, "newmtl")) {
            char* name;
            const size_t nl = take_token(q, name);
            if (nl == 0) {
                err = rd.where() + "`newmtl` without a name";
                return -1;
            }
            const std::string key(name, nl);
            auto it = index.find(key);
            if (it == index.end()) {
                it = index.emplace(key, names.size()).first;
                names.push_back(key);
                mats.push_back(default_material());
            }
            current = (long)it->second;
            continue;
        }
        if (current < 0) {
            err = rd.where() + "Material was not specified with `newmtl` before properties were added";
            return -1;
        }
        rptb_material& m = mats[(size_t)current];
        double v[3];
        if (tok_is(tok, tl, "Kd")) {
            if (!take_doubles(q, v, 3)) {
                err = rd.where() + "Failed to parse vertex in .OBJ";  // the reference reuses parse_obj_point
                return -1;
            }
            m.color[0] = v[0], m.color[1] = v[1], m.color[2] = v[2];
        } else if (tok_is(tok, tl, "Ns")) {
            if (!take_doubles(q, v, 1)) {
                err = rd.where() + "Could not parse Ns value";
                return -1;
            }
            m.roughness = std::sqrt(std::sqrt(2.0 / (v[0] + 2.0)));
        } else if (tok_is(tok, tl, "Ni")) {
            if (!take_doubles(q, v, 1)) {
                err = rd.where() + "Could not parse Ni value";
                return -1;
            }
            m.index = std::fmax(v[0], 1.0 + 1e-4);  // f64::max: a NaN Ni yields 1.0001
        } else if (tok_is(tok, tl, "d")) {
            if (!take_doubles(q, v, 1)) {
                err = rd.where() + "Could not parse d value";
                return -1;
            }
            if (v[0] < 0.8) m.transparent = 1;
        }
    }
    return 0;
}

struct ObjGroup {
    rptb_material material;
    uint64_t first_tri, ntris;
};

// load_obj_with_mtl (src/io.rs:83-149).  All triangles land in `tris` in file order; `groups` cuts
// them into the reference's Vec<Object>.
int parse_obj_mtl_text(const char* obj, size_t obj_len, const char* mtl, size_t mtl_len, std::vector<double>& tris,
                       std::vector<ObjGroup>& groups, std::string& err) {
    std::vector<std::string> names;
    std::vector<rptb_material> mats;
    if (parse_mtl_text(mtl, mtl_len, names, mats, err) != 0) {
        err = ".mtl " + err;
        return -1;
    }
    ObjState st;
    LineReader rd(obj, obj_len);
    tris.clear();
    groups.clear();
    rptb_material current = default_material();
    std::string last_usemtl;
    bool have_usemtl = false;
    uint64_t first = 0;
    auto flush = [&]() {
        const uint64_t n = tris.size() / 18;
        if (n > first) groups.push_back(ObjGroup{current, first, n - first});
        first = n;
    };
    while (char* q = rd.next()) {
        char* tok;
        const size_t tl = take_token(q, tok);
        if (tl == 0 || *tok == '#') continue;
        bool ok = true;
        if (tok_is(tok, tl, "v")) ok = st.point(q, false, err);
        else if (tok_is(tok, tl, "vn")) ok = st.point(q, true, err);
        else if (tok_is(tok, tl, "f")) ok = st.face(q, tris, err);
        else if (tok_is(tok, tl, "usemtl")) {
            char* name;
            const size_t nl = take_token(q, name);
            if (nl == 0) {
                err = "`usemtl` without a name";
                ok = false;
            } else {
                const std::string key(name, nl);
                if (!have_usemtl || key != last_usemtl) {
                    flush();
                    size_t i = 0;
                    while (i < names.size() && names[i] != key) i++;
                    if (i == names.size()) {
                        err = "Could not found `usemtl " + key + "` in library";
                        ok = false;
                    } else {
                        current = mats[i];
                        last_usemtl = key;
                        have_usemtl = true;
                    }
                }
            }
        }
        if (!ok) {
            err = rd.where() + err;
            return -1;
        }
    }
    flush();
    return 0;
}

namespace {

void push_facet(std::vector<double>& tris, const double* vn, const double (*vs)[3]) {
    for (int k = 0; k < 3; k++) tris.insert(tris.end(), vs[k], vs[k] + 3);
    for (int k = 0; k < 3; k++) tris.insert(tris.end(), vn, vn + 3);
}

// load_stl_ascii (src/io.rs:289-329).  Deviation, on purpose: the reference's loop demands
// `facet normal` on every line after the header, so it rejects the closing `endsolid` of every
// well-formed file; here `endsolid` (and blank lines between facets) end / are skipped.
int parse_stl_ascii(const char* data, size_t len, std::vector<double>& tris, std::string& err) {
    LineReader rd(data, len);
    rd.next();  // "solid <name>"
    auto expect_prefix = [&](char*& q, const char* prefix) {
        const size_t pl = strlen(prefix);
        if (strncmp(q, prefix, pl) != 0) return false;
        q += pl;
        return true;
    };
    while (char* q = rd.next()) {
        if (!*q) continue;
        if (strncmp(q, "endsolid", 8) == 0) break;
        double vn[3], vs[3][3];
        if (!expect_prefix(q, "facet normal ")) {
            err = rd.where() + "Malformed STL file: expected `facet normal`";
            return -1;
        }
        if (!take_doubles(q, vn, 3)) {
            err = rd.where() + "Invalid facet normal";
            return -1;
        }
        if (!rd.next()) {  // "outer loop"
            err = "Malformed STL file: truncated facet";
            return -1;
        }
        for (int k = 0; k < 3; k++) {
            q = rd.next();
            if (!q || !expect_prefix(q, "vertex ")) {
                err = rd.where() + "Malformed STL file: expected `vertex`";
                return -1;
            }
            if (!take_doubles(q, vs[k], 3)) {
                err = rd.where() + "Invalid vertex";
                return -1;
            }
        }
        if (!rd.next() || !rd.next()) {  // "endloop", "endfacet"
            err = "Malformed STL file: truncated facet";
            return -1;
        }
        push_facet(tris, vn, vs);
    }
    return 0;
}

// load_stl_binary (src/io.rs:331-360): per facet 12 little-endian f32 (normal, v1, v2, v3) + 2 bytes.
void parse_stl_binary(const unsigned char* data, uint64_t ntris, std::vector<double>& tris) {
    tris.reserve((size_t)ntris * 18);
    const unsigned char* p = data + 84;
    for (uint64_t t = 0; t < ntris; t++, p += 50) {
        double f[12];
        for (int k = 0; k < 12; k++) {
            const unsigned char* b = p + 4 * k;
            const uint32_t u = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
            float x;
            std::memcpy(&x, &u, 4);
            f[k] = (double)x;
        }
        const double vs[3][3] = {{f[3], f[4], f[5]}, {f[6], f[7], f[8]}, {f[9], f[10], f[11]}};
        push_facet(tris, f, vs);
    }
}

}  // namespace

// load_stl (src/io.rs:260-287).
int parse_stl_bytes(const void* data, size_t len, std::vector<double>& tris, std::string& err) {
    const unsigned char* b = (const unsigned char*)data;
    tris.clear();
    if (len < 15) {
        err = "Loaded .STL file is too short";
        return -1;
    }
    if (len >= 84) {
        const uint64_t n = (uint64_t)b[80] | ((uint64_t)b[81] << 8) | ((uint64_t)b[82] << 16) | ((uint64_t)b[83] << 24);
        if ((uint64_t)len == 84 + n * 50) {  // very likely binary
            parse_stl_binary(b, n, tris);
            return 0;
        }
    }
    if (memcmp(b, "solid ", 6) == 0) return parse_stl_ascii((const char*)data, len, tris, err);
    err = "Loaded .STL file, but could not determine format";
    return -1;
}

}  // namespace rptb
