// objparse.cpp -- host-side Wavefront .OBJ ingestion ("next" row N2 of SURVEY section 8f).
//
// Replaces load_obj -> parse_obj_point / parse_obj_face (ekzhang/rpt src/io.rs:27-73,151-200):
// `v`, `vn` and `f` records; faces are fan-triangulated; `a/b/c` and `a//c` corner syntax; 1-based and
// negative (relative) indices; a corner without a normal index makes the whole triangle flat-shaded
// (Triangle::from_vertices, src/shape/mesh.rs:24-37); `vt`, `mtllib`, `usemtl` and unknown records
// are skipped.  The reference parses line by line through BufReader + split_ascii_whitespace +
// str::parse; this is a single pass over the byte buffer with strtod / strtol.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace rptb {

// Returns 0 on success; on failure a negative value and `err` set.  `tris` receives 18 doubles per
// triangle: v1 v2 v3 n1 n2 n3.
int parse_obj_text(const char* text, size_t len, std::vector<double>& tris, std::string& err) {
    std::vector<double> verts, norms;
    std::vector<long> vi, ni;
    std::string line;  // one NUL-terminated line at a time (strtod / strtol need a terminator)
    tris.clear();
    const char* p = text;
    const char* end = text + len;
    size_t lineno = 0;
    auto is_ws = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v'; };
    while (p < end) {
        const char* le = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!le) le = end;
        lineno++;
        line.assign(p, (size_t)(le - p));
        p = le + 1;
        char* q = &line[0];
        while (is_ws(*q)) q++;
        char* tok = q;
        while (*q && !is_ws(*q)) q++;
        const size_t tl = (size_t)(q - tok);
        if (tl == 0 || *tok == '#') continue;
        const bool is_v = tl == 1 && tok[0] == 'v';
        const bool is_vn = tl == 2 && tok[0] == 'v' && tok[1] == 'n';
        const bool is_f = tl == 1 && tok[0] == 'f';
        if (is_v || is_vn) {
            double c[3];
            for (int k = 0; k < 3; k++) {
                char* e2 = nullptr;
                c[k] = std::strtod(q, &e2);
                if (e2 == q) {
                    err = "line " + std::to_string(lineno) + ": Failed to parse vertex in .OBJ";
                    return -1;
                }
                q = e2;
            }
            std::vector<double>& dst = is_v ? verts : norms;
            dst.push_back(c[0]);
            dst.push_back(c[1]);
            dst.push_back(c[2]);
        } else if (is_f) {
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
                    err = "line " + std::to_string(lineno) + ": Invalid vertex index";
                    return -1;
                }
                const long v0 = ia > 0 ? ia - 1 : nv + ia;  // parse_index, src/io.rs:10-18
                if (v0 < 0 || v0 >= nv) {
                    err = "line " + std::to_string(lineno) + ": vertex index out of range";
                    return -1;
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
                                err = "line " + std::to_string(lineno) + ": normal index out of range";
                                return -1;
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
        }
    }
    return 0;
}

}  // namespace rptb
