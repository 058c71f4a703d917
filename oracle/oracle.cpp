// oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
//
// A CPU restatement, in double precision, of the one hot path of ekzhang/rpt
// (@815b21c): Renderer::sample -> get_color -> trace_ray and everything under
// it.  Each function cites the reference file:line it follows.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may load this library; the product (rpt_b200/) never does.
//
// PARITY UNPINNED: the reference holds no golden vector, known-answer test or
// fixture for this path (its three unit tests cover hex_color/color_bytes, a
// monomial surface and an RK4 step).  Only `colors_work` (src/color.rs:26-39)
// pins anything here -- oracle_hex_color/oracle_color_bytes are checked against
// it in tests/test_oracle.py.  rustc/cargo are absent, so the reference itself
// cannot be run.  The oracle is therefore validated by closed-form cases,
// kd-tree == brute force, pdf normalisation and BSDF identities (tests/).
//
// The single deliberate deviation: the reference seeds a ChaCha12 StdRng from
// OS entropy per image row (src/renderer.rs:121) and is not reproducible; the
// oracle draws from Philox4x32-10 keyed by (seed, pixel, sample).  The
// *distributions* drawn (rand 0.8 / rand_distr 0.4 semantics, restated from
// their documented algorithms) and their order are the reference's.
//
// Third-party arithmetic restated here because it is not under the reference
// tree (Cargo.toml:13-18, no Cargo.lock): nalgebra-glm 0.10 vector/matrix ops,
// rand 0.8 (gen, gen_range, gen_bool, Uniform), rand_distr 0.4 (UnitDisc,
// UnitCircle).
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/rpt_b200.h"

namespace {

const double INF = std::numeric_limits<double>::infinity();
const double PI = 3.14159265358979323846264338327950288;

// ---------------------------------------------------------------- glm ------
struct V3 {
    double x, y, z;
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    double& at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 v3(double x, double y, double z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(double s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
inline V3 cmul(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }  // component_mul
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double length2(V3 a) { return dot(a, a); }
inline double length(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) { return a / length(a); }  // nalgebra: self / self.norm()
inline V3 vmin(V3 a, V3 b) { return {std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)}; }
inline V3 vmax(V3 a, V3 b) { return {std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)}; }
inline V3 mix(V3 a, V3 b, double t) { return a * (1.0 - t) + b * t; }  // glm::mix / lerp
// Rust f64::powi -> compiler-rt __powidf2 (square and multiply)
inline double powi(double a, int b) {
    const bool recip = b < 0;
    double r = 1.0;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0 / r : r;
}
// Rust f64::signum: 1.0 for +0.0, -1.0 for -0.0, NaN for NaN
inline double signum(double x) { return std::isnan(x) ? x : std::copysign(1.0, x); }

struct M4 {  // column-major, m[c*4+r]
    double m[16];
};
struct M3 {  // column-major, m[c*3+r]
    double m[9];
};
inline V3 mul_point(const M4& a, V3 p) {  // (M * (p,1)).xyz
    return {a.m[0] * p.x + a.m[4] * p.y + a.m[8] * p.z + a.m[12],
            a.m[1] * p.x + a.m[5] * p.y + a.m[9] * p.z + a.m[13],
            a.m[2] * p.x + a.m[6] * p.y + a.m[10] * p.z + a.m[14]};
}
inline V3 mul_dir(const M4& a, V3 d) {  // (M * (d,0)).xyz
    return {a.m[0] * d.x + a.m[4] * d.y + a.m[8] * d.z, a.m[1] * d.x + a.m[5] * d.y + a.m[9] * d.z,
            a.m[2] * d.x + a.m[6] * d.y + a.m[10] * d.z};
}
inline V3 mul(const M3& a, V3 d) {
    return {a.m[0] * d.x + a.m[3] * d.y + a.m[6] * d.z, a.m[1] * d.x + a.m[4] * d.y + a.m[7] * d.z,
            a.m[2] * d.x + a.m[5] * d.y + a.m[8] * d.z};
}
inline M3 mat4_to_mat3(const M4& a) {
    return M3{{a.m[0], a.m[1], a.m[2], a.m[4], a.m[5], a.m[6], a.m[8], a.m[9], a.m[10]}};
}
inline double det3(const M3& a) {
    return a.m[0] * (a.m[4] * a.m[8] - a.m[7] * a.m[5]) - a.m[3] * (a.m[1] * a.m[8] - a.m[7] * a.m[2]) +
           a.m[6] * (a.m[1] * a.m[5] - a.m[4] * a.m[2]);
}
inline M3 inverse_transpose3(const M3& a) {  // glm::inverse_transpose
    const double d = det3(a);
    M3 r;
    // cofactor matrix / det == (A^-1)^T
    r.m[0] = (a.m[4] * a.m[8] - a.m[7] * a.m[5]) / d;
    r.m[1] = -(a.m[3] * a.m[8] - a.m[6] * a.m[5]) / d;
    r.m[2] = (a.m[3] * a.m[7] - a.m[6] * a.m[4]) / d;
    r.m[3] = -(a.m[1] * a.m[8] - a.m[7] * a.m[2]) / d;
    r.m[4] = (a.m[0] * a.m[8] - a.m[6] * a.m[2]) / d;
    r.m[5] = -(a.m[0] * a.m[7] - a.m[6] * a.m[1]) / d;
    r.m[6] = (a.m[1] * a.m[5] - a.m[4] * a.m[2]) / d;
    r.m[7] = -(a.m[0] * a.m[5] - a.m[3] * a.m[2]) / d;
    r.m[8] = (a.m[0] * a.m[4] - a.m[3] * a.m[1]) / d;
    return r;
}
// glm::inverse for a 4x4: Gauss-Jordan with partial pivoting (double).
inline M4 inverse4(const M4& a) {
    double w[4][8];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            w[r][c] = a.m[c * 4 + r];
            w[r][c + 4] = (r == c) ? 1.0 : 0.0;
        }
    for (int i = 0; i < 4; i++) {
        int p = i;
        for (int r = i + 1; r < 4; r++)
            if (std::fabs(w[r][i]) > std::fabs(w[p][i])) p = r;
        if (p != i)
            for (int c = 0; c < 8; c++) std::swap(w[i][c], w[p][c]);
        const double piv = w[i][i];
        for (int c = 0; c < 8; c++) w[i][c] /= piv;
        for (int r = 0; r < 4; r++)
            if (r != i) {
                const double f = w[r][i];
                if (f != 0.0)
                    for (int c = 0; c < 8; c++) w[r][c] -= f * w[i][c];
            }
    }
    M4 out;
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) out.m[c * 4 + r] = w[r][c + 4];
    return out;
}

// ------------------------------------------------------------ Philox -------
// Philox4x32-10 (Salmon et al., SC'11).  key = (seed_lo, seed_hi),
// counter = (block, pixel, sample_lo, sample_hi).  A stream of u64 draws is two
// Philox streams side by side: draw 4b + w = (word w of block b) << 32 | (word w
// of block b | 2^31) -- the product's f32 kernels, which keep only the high half
// of a draw, then need the first stream alone (rpt_b200/csrc/rng.cuh).
struct Philox {
    uint32_t key[2];
    uint32_t ctr[4];
    uint64_t out[4];
    int have;  // number of unread u64 in out (0..4)
    static inline void round(uint32_t c[4], const uint32_t k[2]) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
        const uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    static inline void block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
        uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
        uint32_t k[2] = {key[0], key[1]};
        for (int i = 0; i < 10; i++) {
            round(c, k);
            k[0] += 0x9E3779B9u;
            k[1] += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
    Philox(uint64_t seed, uint32_t pixel, uint64_t sample) {
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
        ctr[0] = 0; ctr[1] = pixel; ctr[2] = (uint32_t)sample; ctr[3] = (uint32_t)(sample >> 32);
        have = 0;
    }
    uint64_t next_u64() {
        if (have == 0) {
            uint32_t hi[4], lo[4];
            const uint32_t low_ctr[4] = {ctr[0] | 0x80000000u, ctr[1], ctr[2], ctr[3]};
            block(ctr, key, hi);
            block(low_ctr, key, lo);
            for (int w = 0; w < 4; w++) out[w] = ((uint64_t)hi[w] << 32) | lo[w];
            ctr[0]++;
            have = 4;
        }
        return out[4 - have--];
    }
};

// rand 0.8 / rand_distr 0.4 semantics over the Philox stream [recall].
struct Rng {
    Philox p;
    Rng(uint64_t seed, uint32_t pixel, uint64_t sample) : p(seed, pixel, sample) {}
    // Standard f64: 53 bits, [0,1)
    double gen_f64() { return (double)(p.next_u64() >> 11) * (1.0 / 9007199254740992.0); }
    // UniformFloat: 52-bit mantissa value in [0,1)
    double u52() { return (double)(p.next_u64() >> 12) * (1.0 / 4503599627370496.0); }
    // Rng::gen_range(lo..hi) -> UniformFloat::sample_single
    double gen_range(double lo, double hi) {
        const double scale = hi - lo;
        while (true) {
            const double v12 = 1.0 + u52();
            const double res = v12 * scale + (lo - scale);
            if (res < hi) return res;
        }
    }
    // Uniform::new(-1., 1.).sample
    double uniform_pm1() { return u52() * 2.0 + -1.0; }
    // Rng::gen_bool(p) -> Bernoulli::new(p).unwrap().sample; p outside [0,1] panics in the reference
    bool gen_bool(double prob) {
        if (prob >= 1.0) { p.next_u64(); return true; }  // ALWAYS_TRUE; see note below
        const uint64_t p_int = (uint64_t)(prob * 18446744073709551616.0);
        return p.next_u64() < p_int;
    }
    // Rng::gen::<bool>() -> Standard: the sign bit of one 32-bit word; here the top bit of one u64 draw
    bool gen_bool_std() { return (p.next_u64() >> 63) != 0; }
    // Uniform::from(0..n) for usize (widening-multiply rejection)
    uint64_t uniform_usize(uint64_t n) {
        const uint64_t ints_to_reject = (UINT64_MAX - n + 1) % n;
        const uint64_t zone = UINT64_MAX - ints_to_reject;
        while (true) {
            const uint64_t v = p.next_u64();
            const unsigned __int128 m = (unsigned __int128)v * n;
            const uint64_t lo = (uint64_t)m;
            if (lo <= zone) return (uint64_t)(m >> 64);
        }
    }
    void unit_disc(double& x, double& y) {  // rand_distr::UnitDisc
        while (true) {
            x = uniform_pm1();
            y = uniform_pm1();
            if (x * x + y * y <= 1.0) return;
        }
    }
    void unit_circle(double& x, double& y) {  // rand_distr::UnitCircle (von Neumann)
        double x1, x2, sum;
        while (true) {
            x1 = uniform_pm1();
            x2 = uniform_pm1();
            sum = x1 * x1 + x2 * x2;
            if (sum < 1.0) break;
        }
        const double diff = x1 * x1 - x2 * x2;
        x = diff / sum;
        y = 2.0 * x1 * x2 / sum;
    }
};
// Note on gen_bool(1.0): rand's Bernoulli stores p_int = u64::MAX sentinel and
// returns true *without* drawing.  The oracle (and the GPU kernel) DO consume
// one u64 in that case so that the draw count per vertex does not depend on a
// material parameter; the distribution is identical.

// ------------------------------------------------------------- counters ----
struct Counters {
    uint64_t segments = 0, rays = 0, node_visits = 0, tri_tests = 0, mesh_hits = 0, env_lookups = 0, object_tests = 0;
    void add(const Counters& o) {
        segments += o.segments; rays += o.rays; node_visits += o.node_visits;
        tri_tests += o.tri_tests; mesh_hits += o.mesh_hits; env_lookups += o.env_lookups; object_tests += o.object_tests;
    }
};

// --------------------------------------------------- src/shape.rs:48-97 ----
struct Ray {
    V3 origin, dir;
    V3 at(double t) const { return origin + t * dir; }  // :59-61
    Ray apply_transform(const M4& m) const {            // :64-72 (dir not renormalised)
        return Ray{mul_point(m, origin), mul_dir(m, dir)};
    }
};
struct HitRecord {
    double time = INF;  // :83-90
    V3 normal = {0, 0, 0};
    bool on_mesh = false;  // oracle-only bookkeeping for counters
};

// ------------------------------------------------- src/kdtree.rs:27-87 -----
struct BoundingBox {
    V3 p_min = {INF, INF, INF};
    V3 p_max = {-INF, -INF, -INF};
    BoundingBox merge(const BoundingBox& o) const { return BoundingBox{vmin(p_min, o.p_min), vmax(p_max, o.p_max)}; }
    void intersect(const Ray& ray, double& t0, double& t1) const {  // :54-68
        double x1 = (p_min.x - ray.origin.x) / ray.dir.x;
        double x2 = (p_max.x - ray.origin.x) / ray.dir.x;
        double a = std::fmin(x1, x2), b = std::fmax(x1, x2);
        x1 = a; x2 = b;
        double y1 = (p_min.y - ray.origin.y) / ray.dir.y;
        double y2 = (p_max.y - ray.origin.y) / ray.dir.y;
        a = std::fmin(y1, y2); b = std::fmax(y1, y2);
        y1 = a; y2 = b;
        double z1 = (p_min.z - ray.origin.z) / ray.dir.z;
        double z2 = (p_max.z - ray.origin.z) / ray.dir.z;
        a = std::fmin(z1, z2); b = std::fmax(z1, z2);
        z1 = a; z2 = b;
        t0 = std::fmax(std::fmax(x1, y1), z1);
        t1 = std::fmin(std::fmin(x2, y2), z2);
    }
    void split(int axis, double value, BoundingBox& l, BoundingBox& r) const {  // :71-86
        V3 mid_max = p_max; mid_max.at(axis) = value;
        V3 mid_min = p_min; mid_min.at(axis) = value;
        l = BoundingBox{p_min, mid_max};
        r = BoundingBox{mid_min, p_max};
    }
};

struct Shape {  // src/shape.rs:18-25
    virtual ~Shape() {}
    virtual bool intersect(const Ray& ray, double t_min, HitRecord& rec, Counters& c) const = 0;
    virtual void sample(const V3& target, Rng& rng, V3& v, V3& n, double& p) const = 0;
    // trait Bounded (src/kdtree.rs:8-12); false = the shape does not implement it (Plane)
    virtual bool bounding_box(BoundingBox&) const { return false; }
};

// ------------------------------------------- src/shape/sphere.rs:13-64 -----
struct Sphere : Shape {
    bool intersect(const Ray& ray, double t_min, HitRecord& rec, Counters&) const override {
        const double a = length2(ray.dir);
        const double b = dot(ray.dir, ray.origin);
        const double c = length2(ray.origin) - 1.0;
        double d = b * b - a * c;
        if (std::signbit(d)) return false;  // is_sign_negative
        d = std::sqrt(d);
        double t;
        const double t_minus = (-b - d) / a;
        if (t_minus < t_min) {
            const double t_plus = (-b + d) / a;
            if (t_plus < t_min) return false;
            t = t_plus;
        } else {
            t = t_minus;
        }
        if (t < rec.time) {
            rec.time = t;
            rec.normal = normalize(ray.at(t));
            rec.on_mesh = false;
            return true;
        }
        return false;
    }
    void sample(const V3& target, Rng& rng, V3& v, V3& nn, double& p) const override {  // :52-64
        double x, y;
        rng.unit_disc(x, y);
        const double z = std::sqrt(1.0 - x * x - y * y);
        const V3 n = normalize(target);
        const V3 n1 = std::isnormal(n.x) ? normalize(v3(n.y, -n.x, 0.0)) : normalize(v3(0.0, -n.z, n.y));
        const V3 n2 = cross(n1, n);
        const V3 pt = x * n1 + y * n2 + z * n;
        v = pt;
        nn = pt;
        p = z * (1.0 / PI);  // z * FRAC_1_PI
    }
    bool bounding_box(BoundingBox& b) const override {  // :67-74
        b = BoundingBox{v3(-1.0, -1.0, -1.0), v3(1.0, 1.0, 1.0)};
        return true;
    }
};

// -------------------------------------------- src/shape/plane.rs:17-36 -----
struct Plane : Shape {
    V3 normal;
    double value;
    bool intersect(const Ray& ray, double t_min, HitRecord& rec, Counters&) const override {
        const double cosine = dot(normal, ray.dir);
        if (std::fabs(cosine) < 1e-8) return false;
        const double time = (value - dot(normal, ray.origin)) / cosine;
        if (time >= t_min && time < rec.time) {
            rec.time = time;
            rec.normal = -normalize(normal) * signum(cosine);
            rec.on_mesh = false;
            return true;
        }
        return false;
    }
    void sample(const V3&, Rng&, V3& v, V3& n, double& p) const override {
        // unimplemented!() in the reference (:34-36): a plane cannot be a light.
        v = n = v3(NAN, NAN, NAN);
        p = NAN;
    }
};

// --------------------------------------------- src/shape/cube.rs:20-87 -----
struct Cube : Shape {
    bool intersect(const Ray& ray, double t_min, HitRecord& rec, Counters&) const override {
        double lo[3], hi[3];
        V3 lon[3], hin[3];
        for (int dim = 0; dim < 3; dim++) {  // compute_interval
            double x1 = (-0.5 - ray.origin[dim]) / ray.dir[dim];
            double x2 = (0.5 - ray.origin[dim]) / ray.dir[dim];
            V3 x1n = {0, 0, 0}, x2n = {0, 0, 0};
            x1n.at(dim) = -1.0;
            x2n.at(dim) = 1.0;
            if (x1 > x2) {
                std::swap(x1, x2);
                std::swap(x1n, x2n);
            }
            lo[dim] = x1; hi[dim] = x2; lon[dim] = x1n; hin[dim] = x2n;
        }
        double start, end;
        V3 start_normal, end_normal;
        if (lo[0] > lo[1] && lo[0] > lo[2]) { start = lo[0]; start_normal = lon[0]; }
        else if (lo[1] > lo[2]) { start = lo[1]; start_normal = lon[1]; }
        else { start = lo[2]; start_normal = lon[2]; }
        if (hi[0] < hi[1] && hi[0] < hi[2]) { end = hi[0]; end_normal = hin[0]; }
        else if (hi[1] < hi[2]) { end = hi[1]; end_normal = hin[1]; }
        else { end = hi[2]; end_normal = hin[2]; }
        if (start > end || end < t_min) return false;
        double time;
        V3 normal;
        if (start < t_min) { time = end; normal = end_normal; }
        else { time = start; normal = start_normal; }
        if (time < rec.time) {
            rec.time = time;
            rec.normal = normal;
            rec.on_mesh = false;
            return true;
        }
        return false;
    }
    void sample(const V3&, Rng& rng, V3& v, V3& n, double& p) const override {  // :74-87
        const double a = rng.gen_f64() - 0.5;
        const double b = rng.gen_f64() - 0.5;
        switch (rng.uniform_usize(6)) {
            case 0: v = v3(a, b, 0.5); n = v3(0, 0, 1); break;
            case 1: v = v3(a, b, -0.5); n = v3(0, 0, -1); break;
            case 2: v = v3(a, 0.5, b); n = v3(0, 1, 0); break;
            case 3: v = v3(a, -0.5, b); n = v3(0, -1, 0); break;
            case 4: v = v3(0.5, a, b); n = v3(1, 0, 0); break;
            default: v = v3(-0.5, a, b); n = v3(-1, 0, 0); break;
        }
        p = 1.0 / 6.0;
    }
    bool bounding_box(BoundingBox& b) const override {  // :10-17
        b = BoundingBox{v3(-0.5, -0.5, -0.5), v3(0.5, 0.5, 0.5)};
        return true;
    }
};

// ------------------------------------ src/shape/monomial_surface.rs:13-187 --
// y = height * (x^2 + z^2)^(exp/2) over the unit disc; intersect and the normals hard-code exp = 4,
// as the reference does ("Normals and other things ... work only for exp=4 for now", :11).
struct MonomialSurface : Shape {
    double height = 1.0, exp = 4.0;
    bool bounding_box(BoundingBox& b) const override {  // :179-186
        b = BoundingBox{v3(-1.0, 0.0, -1.0), v3(1.0, height, 1.0)};
        return true;
    }
    bool intersect(const Ray& ray, double t_min, HitRecord& rec, Counters&) const override {  // :21-105
        BoundingBox bb;
        bounding_box(bb);
        double b_min, b_max;
        bb.intersect(ray, b_min, b_max);
        if (std::fmax(b_min, t_min) > std::fmin(b_max, rec.time)) return false;
        auto dist = [&](double t) {
            const double x = ray.origin.x + t * ray.dir.x;
            const double y = ray.origin.y + t * ray.dir.y;
            const double z = ray.origin.z + t * ray.dir.z;
            return y - height * powi(x * x + z * z, 2);
        };
        const double coef0 = powi(ray.origin.x, 2) + powi(ray.origin.z, 2);
        const double coef1 = 2. * (ray.origin.x * ray.dir.x + ray.origin.z * ray.dir.z);
        const double coef2 = powi(ray.dir.x, 2) + powi(ray.dir.z, 2);
        auto deriv = [&](double t) {
            const double dy = 2. * coef0 * coef1 + 2. * t * (coef1 * coef1 + 2. * coef0 * coef2) +
                              3. * powi(t, 2) * 2. * coef1 * coef2 + 4. * powi(t, 3) * coef2 * coef2;
            return ray.dir.y - height * dy;
        };
        auto deriv2 = [&](double t) {
            const double dy = 2. * (coef1 * coef1 + 2. * coef0 * coef2) + 3. * 2. * t * 2. * coef1 * coef2 +
                              4. * 3. * powi(t, 2) * coef2 * coef2;
            return -height * dy;
        };
        double t_max;
        const bool maximize = dist(t_min) < 0.0;
        if (maximize) {  // below the surface: Newton towards the maximum of dist along the ray
            double cur_x = (b_min + b_max) / 2.;
            for (int it = 0; it < 10; it++) {
                const double f = dist(cur_x);
                if (f > 0.) break;
                const double der = deriv(cur_x);
                const double der2 = deriv2(cur_x);
                cur_x -= der / der2;
            }
            // (:62-64 prints a diagnostic here; no effect on the result)
            t_max = cur_x;
            if (t_max < t_min) return false;
        } else {
            t_max = 10000.;
        }
        if ((dist(t_min) < 0.0) == (dist(t_max) < 0.0)) return false;
        double l = t_min, r = t_max;
        for (int it = 0; it < 60; it++) {  // bisection :75-84
            const double m = (l + r) / 2.0;
            if ((dist(m) >= 0.0) == maximize) r = m;
            else l = m;
        }
        if (r > rec.time) return false;
        const V3 pos = ray.at(r);
        if (pos.x * pos.x + pos.z * pos.z > 1.0) return false;  // outside the rim
        rec.time = r;
        rec.normal = normalize(v3(height * 4.0 * pos.x * (pos.x * pos.x + pos.z * pos.z), -1.0,
                                  height * 4.0 * pos.z * (pos.x * pos.x + pos.z * pos.z)));
        if (dot(rec.normal, ray.dir) > 0.0) rec.normal = -rec.normal;  // two-sided
        rec.on_mesh = false;
        return true;
    }
    void sample(const V3&, Rng& rng, V3& v, V3& n, double& p) const override {  // :107-122
        // NB: UnitCircle -- the reference samples the RIM of the surface only (a quirk, kept)
        double x, z;
        rng.unit_circle(x, z);
        const V3 pos = v3(x, height * std::pow(x * x + z * z, exp / 2.), z);
        V3 normal = normalize(v3(height * 4. * pos.x * (pos.x * pos.x + pos.z * pos.z), -1.,
                                 height * 4. * pos.z * (pos.x * pos.x + pos.z * pos.z)));
        const double AREA = 6.3406654362;
        if (rng.gen_bool_std()) normal = -normal;
        v = pos;
        n = normal;
        p = 1. / (2. * AREA);
    }
};

// ---------------------------------------------- src/shape/mesh.rs:7-98 -----
struct Triangle {
    V3 v1, v2, v3_, n1, n2, n3;
    BoundingBox bounding_box() const {  // :39-46
        return BoundingBox{vmin(vmin(v1, v2), v3_), vmax(vmax(v1, v2), v3_)};
    }
    bool intersect(const Ray& ray, double t_min, HitRecord& rec, Counters& c) const {  // :49-82
        c.tri_tests++;
        const V3 d0 = v2 - v1, d1 = v3_ - v1;
        const V3 plane_normal = normalize(cross(d0, d1));
        const double cosine = dot(plane_normal, ray.dir);
        if (std::fabs(cosine) < 1e-8) return false;
        const double time = dot(plane_normal, v1 - ray.origin) / cosine;
        if (time < t_min || time >= rec.time) return false;
        const V3 d2 = ray.at(time) - v1;
        const double d00 = dot(d0, d0), d01 = dot(d0, d1), d11 = dot(d1, d1);
        const double d20 = dot(d2, d0), d21 = dot(d2, d1);
        const double denom = d00 * d11 - d01 * d01;
        const double v = (d11 * d20 - d01 * d21) / denom;
        const double w = (d00 * d21 - d01 * d20) / denom;
        const double u = 1.0 - v - w;
        if (u >= 0.0 && v >= 0.0 && w >= 0.0) {
            rec.time = time;
            rec.normal = normalize(u * n1 + v * n2 + w * n3);
            rec.on_mesh = true;
            return true;
        }
        return false;
    }
    void sample(Rng& rng, V3& pt, V3& n, double& p) const {  // :84-98
        double u = rng.gen_f64();
        double v = rng.gen_f64();
        while (u + v > 1.0) {
            u = rng.gen_f64();
            v = rng.gen_f64();
        }
        const double w = 1.0 - u - v;
        const double area = 0.5 * length(cross(v2 - v1, v3_ - v1));
        pt = u * v1 + v * v2 + w * v3_;
        n = normalize(u * n1 + v * n2 + w * n3);
        p = 1.0 / area;
    }
};

// --------------------------------------------- src/kdtree.rs:99-355 --------
struct KdNode {  // :226-233
    int kind = 3;  // 0,1,2 = SplitX/Y/Z, 3 = Leaf
    double value = 0;
    std::unique_ptr<KdNode> left, right;
    std::vector<size_t> indices;
};

double median(const std::vector<double>& s) {  // :347-355
    assert(!s.empty());
    if (s.size() % 2 == 1) return s[s.size() / 2];
    const size_t mid = s.size() / 2;
    return (s[mid] + s[mid - 1]) / 2.0;
}

// `boxes[i]` = objects[i].bounding_box(): the reference recomputes it at every node, the value is the same.
std::unique_ptr<KdNode> construct(const std::vector<BoundingBox>& boxes, std::vector<size_t> indices) {  // :235-345
    auto node = std::make_unique<KdNode>();
    if (indices.size() < 16) {
        node->indices = std::move(indices);
        return node;
    }
    std::vector<double> xs, ys, zs;
    std::vector<BoundingBox> bboxs;
    for (size_t index : indices) {
        const BoundingBox bb = boxes[index];
        xs.push_back(bb.p_min.x); xs.push_back(bb.p_max.x);
        ys.push_back(bb.p_min.y); ys.push_back(bb.p_max.y);
        zs.push_back(bb.p_min.z); zs.push_back(bb.p_max.z);
        bboxs.push_back(bb);
    }
    std::sort(xs.begin(), xs.end());
    std::sort(ys.begin(), ys.end());
    std::sort(zs.begin(), zs.end());
    const double m[3] = {median(xs), median(ys), median(zs)};
    auto partition_score = [&](int dim, double value) {
        size_t left = 0, right = 0;
        for (const BoundingBox& bb : bboxs) {
            if (bb.p_min[dim] <= value) left++;
            if (bb.p_max[dim] >= value) right++;
        }
        return std::max(left, right);
    };
    const size_t sx = partition_score(0, m[0]), sy = partition_score(1, m[1]), sz = partition_score(2, m[2]);
    const size_t threshold = (size_t)((double)indices.size() * 0.85);  // SCORE_THRESHOLD :6,286
    if (std::min(std::min(sx, sy), sz) >= threshold) {
        node->indices = std::move(indices);
        return node;
    }
    int split_dir = -1;
    BoundingBox bounds;
    for (const BoundingBox& bb : bboxs) bounds = bounds.merge(bb);
    const V3 extent = bounds.p_max - bounds.p_min;
    if (extent.x > extent.y && extent.x > extent.z) {
        if (sx < threshold) split_dir = 0;
    } else if (extent.y > extent.z) {
        if (sy < threshold) split_dir = 1;
    } else if (sz < threshold) {
        split_dir = 2;
    }
    if (split_dir == -1) {
        if (sx < sy && sx < sz) split_dir = 0;
        else if (sy < sz) split_dir = 1;
        else split_dir = 2;
    }
    std::vector<size_t> left, right;
    for (size_t i = 0; i < indices.size(); i++) {  // partition :270-281
        if (bboxs[i].p_min[split_dir] <= m[split_dir]) left.push_back(indices[i]);
        if (bboxs[i].p_max[split_dir] >= m[split_dir]) right.push_back(indices[i]);
    }
    node->kind = split_dir;
    node->value = m[split_dir];
    node->left = construct(boxes, std::move(left));
    node->right = construct(boxes, std::move(right));
    return node;
}

// What KdTree<T> needs from T (T: Bounded, src/kdtree.rs:106): Triangle, or any Bounded shape
// behind a pointer (Box<dyn Bounded> / Arc<Mesh>, src/kdtree.rs:14-24).
typedef std::shared_ptr<const Shape> ShapePtr;
inline BoundingBox elem_box(const Triangle& t) { return t.bounding_box(); }
inline bool elem_intersect(const Triangle& t, const Ray& ray, double t_min, HitRecord& rec, Counters& c) {
    return t.intersect(ray, t_min, rec, c);
}
inline void elem_sample(const Triangle& t, const V3&, Rng& rng, V3& v, V3& n, double& p) { t.sample(rng, v, n, p); }
inline BoundingBox elem_box(const ShapePtr& s) {
    BoundingBox b;
    const bool bounded = s->bounding_box(b);
    assert(bounded);
    (void)bounded;
    return b;
}
inline bool elem_intersect(const ShapePtr& s, const Ray& ray, double t_min, HitRecord& rec, Counters& c) {
    c.object_tests++;
    return s->intersect(ray, t_min, rec, c);
}
inline void elem_sample(const ShapePtr& s, const V3& target, Rng& rng, V3& v, V3& n, double& p) {
    s->sample(target, rng, v, n, p);
}

template <class T>
struct KdTreeT : Shape {  // KdTree<T>; KdTree<Triangle> = Mesh
    std::unique_ptr<KdNode> root;
    std::vector<T> objects;
    BoundingBox bounds;
    bool brute_force = false;  // oracle self-check mode: ignore the tree

    std::vector<BoundingBox> boxes() const {
        std::vector<BoundingBox> b;
        b.reserve(objects.size());
        for (const T& t : objects) b.push_back(elem_box(t));
        return b;
    }
    void init_bounds() {  // :108-119
        bounds = BoundingBox();
        for (const T& t : objects) bounds = bounds.merge(elem_box(t));
    }
    void build() {  // KdTree::new
        std::vector<size_t> idx(objects.size());
        for (size_t i = 0; i < idx.size(); i++) idx[i] = i;
        root = construct(boxes(), std::move(idx));
    }
    bool bounding_box(BoundingBox& b) const override {  // :122-126
        b = bounds;
        return true;
    }
    bool intersect(const Ray& ray, double t_min, HitRecord& rec, Counters& c) const override {  // :129-136
        if (brute_force) {
            bool result = false;
            for (const T& t : objects)
                if (elem_intersect(t, ray, t_min, rec, c)) result = true;
            return result;
        }
        double b_min, b_max;
        bounds.intersect(ray, b_min, b_max);
        if (std::fmax(b_min, t_min) > std::fmin(b_max, rec.time)) return false;
        return intersect_subtree(*root, bounds, ray, t_min, rec, c);
    }
    void sample(const V3& target, Rng& rng, V3& v, V3& n, double& p) const override {  // :138-143
        const size_t num = objects.size();
        const size_t index = (size_t)rng.uniform_usize(num);
        elem_sample(objects[index], target, rng, v, n, p);
        p = p / (double)num;
    }
    bool intersect_subtree(const KdNode& node, const BoundingBox& bbox, const Ray& ray, double t_min,
                           HitRecord& rec, Counters& c) const {  // :151-223
        c.node_visits++;
        double b_min, b_max;
        bbox.intersect(ray, b_min, b_max);
        if (node.kind == 3) {
            bool result = false;
            for (size_t index : node.indices)
                if (elem_intersect(objects[index], ray, t_min, rec, c)) result = true;
            return result;
        }
        const int ax = node.kind;
        const double value = node.value;
        const double t_split = (value - ray.origin[ax]) / ray.dir[ax];
        const bool left_first = (ray.origin[ax] < value) || (ray.origin[ax] == value && ray.dir[ax] <= 0.0);
        BoundingBox bl, br;
        bbox.split(ax, value, bl, br);
        const KdNode* first = left_first ? node.left.get() : node.right.get();
        const KdNode* second = left_first ? node.right.get() : node.left.get();
        const BoundingBox& b0 = left_first ? bl : br;
        const BoundingBox& b1 = left_first ? br : bl;
        if (t_split > std::fmin(b_max, rec.time) || t_split <= 0.0) {
            return intersect_subtree(*first, b0, ray, t_min, rec, c);
        } else if (t_split < std::fmax(b_min, t_min)) {
            return intersect_subtree(*second, b1, ray, t_min, rec, c);
        } else {
            const bool h1 = intersect_subtree(*first, b0, ray, t_min, rec, c);
            if (h1 && rec.time < t_split) return true;
            const bool h2 = intersect_subtree(*second, b1, ray, t_split, rec, c);
            return h1 || h2;
        }
    }
};
typedef KdTreeT<Triangle> KdTree;      // Mesh (src/shape/mesh.rs:102)
typedef KdTreeT<ShapePtr> ShapeTree;   // KdTree<Box<dyn Bounded>> (examples/fractal_*.rs)

// ----------------------------------------------- src/shape.rs:99-150 -------
struct Transformed : Shape {
    ShapePtr shape;
    M4 transform, inverse_transform;
    M3 linear, normal_transform;
    double scale;
    Transformed(ShapePtr s, const M4& t) : shape(std::move(s)), transform(t) {  // :111-124
        inverse_transform = inverse4(t);
        linear = mat4_to_mat3(t);
        scale = det3(linear);
        normal_transform = inverse_transpose3(linear);
    }
    bool intersect(const Ray& ray, double t_min, HitRecord& rec, Counters& c) const override {  // :128-137
        const Ray local = ray.apply_transform(inverse_transform);
        if (shape->intersect(local, t_min, rec, c)) {
            rec.normal = normalize(mul(normal_transform, rec.normal));
            return true;
        }
        return false;
    }
    void sample(const V3& target, Rng& rng, V3& v, V3& n, double& p) const override {  // :139-150
        const V3 t = mul_point(inverse_transform, target);
        V3 lv, ln;
        double lp;
        shape->sample(t, rng, lv, ln, lp);
        const V3 new_normal = normalize(mul(normal_transform, ln));
        const double parallelepiped_height = dot(mul(linear, ln), new_normal);
        const double parallelepiped_base = scale / parallelepiped_height;
        v = mul_point(transform, lv);
        n = new_normal;
        p = lp / parallelepiped_base;
    }
    bool bounding_box(BoundingBox& out) const override {  // :153-175: box of the 8 transformed corners
        BoundingBox b;
        if (!shape->bounding_box(b)) return false;
        out = BoundingBox();
        for (int i = 0; i < 8; i++) {
            const V3 corner = v3((i & 4) ? b.p_max.x : b.p_min.x, (i & 2) ? b.p_max.y : b.p_min.y, (i & 1) ? b.p_max.z : b.p_min.z);
            const V3 v = mul_point(transform, corner);
            out = BoundingBox{vmin(out.p_min, v), vmax(out.p_max, v)};
        }
        return true;
    }
};

// ------------------------------------------------ src/material.rs ----------
struct Material {
    V3 color;
    double index, roughness, metallic, emittance;
    bool transparent;
};

V3 bsdf(const Material& m, const V3& n, const V3& wo, const V3& wi) {  // :125-210
    const double n_dot_wi = dot(n, wi);
    const double n_dot_wo = dot(n, wo);
    const bool wi_outside = !std::signbit(n_dot_wi);
    const bool wo_outside = !std::signbit(n_dot_wo);
    if (!m.transparent && (!wi_outside || !wo_outside)) return v3(0, 0, 0);
    const V3 one = v3(1, 1, 1);
    if (wi_outside == wo_outside) {
        const V3 h = normalize(wi + wo);
        const double wo_dot_h = dot(wo, h);
        const double n_dot_h = dot(n, h);
        const double nh2 = powi(n_dot_h, 2);
        const double m2 = m.roughness * m.roughness;
        const double d = std::exp((nh2 - 1.0) / (m2 * nh2)) / (m2 * PI * nh2 * nh2);
        V3 f;
        if (!wi_outside && std::sqrt(1.0 - wo_dot_h * wo_dot_h) * m.index > 1.0) {
            f = one;
        } else {
            const double f0s = powi((m.index - 1.0) / (m.index + 1.0), 2);
            const V3 f0 = mix(v3(f0s, f0s, f0s), m.color, m.metallic);
            f = f0 + (one - f0) * powi(1.0 - wo_dot_h, 5);
        }
        double g = std::fmin(n_dot_wi * n_dot_h, n_dot_wo * n_dot_h);
        g = (2.0 * g) / wo_dot_h;
        g = std::fmin(g, 1.0);
        const V3 specular = d * f * g / (4.0 * n_dot_wo * n_dot_wi);
        if (m.transparent) return specular;
        const V3 diffuse = cmul(one - f, m.color) / PI;
        return specular + diffuse;
    } else {
        const double eta_t = wo_outside ? m.index : 1.0 / m.index;
        const V3 h = normalize(wi * eta_t + wo);
        const double wi_dot_h = dot(wi, h);
        const double wo_dot_h = dot(wo, h);
        const double n_dot_h = dot(n, h);
        const double nh2 = powi(n_dot_h, 2);
        const double m2 = m.roughness * m.roughness;
        const double d = std::exp((nh2 - 1.0) / (m2 * nh2)) / (m2 * PI * nh2 * nh2);
        const double f0s = powi((m.index - 1.0) / (m.index + 1.0), 2);
        const V3 f0 = mix(v3(f0s, f0s, f0s), m.color, m.metallic);
        const V3 f = f0 + (one - f0) * powi(1.0 - std::fabs(wi_dot_h), 5);
        double g = std::fmin(std::fabs(n_dot_wi * n_dot_h), std::fabs(n_dot_wo * n_dot_h));
        g = (2.0 * g) / std::fabs(wo_dot_h);
        g = std::fmin(g, 1.0);
        const V3 btdf = std::fabs(wi_dot_h * wo_dot_h / (n_dot_wi * n_dot_wo)) *
                        (d * (one - f) * g / powi(eta_t * wi_dot_h + wo_dot_h, 2));
        return cmul(btdf, m.color);
    }
}

struct Frame {  // local_to_world :316-324: columns (ns, nss, n)
    V3 ns, nss, n;
    V3 apply(V3 h) const { return ns * h.x + nss * h.y + n * h.z; }
};
Frame local_to_world(const V3& n) {
    const V3 ns = std::isnormal(n.x) ? normalize(v3(n.y, -n.x, 0.0)) : normalize(v3(0.0, -n.z, n.y));
    const V3 nss = cross(n, ns);
    return Frame{ns, nss, n};
}

bool sample_f(const Material& m, const V3& n, const V3& wo, Rng& rng, V3& wi_out, double& pdf_out) {  // :224-313
    const double m2 = m.roughness * m.roughness;
    const double f0 = powi((m.index - 1.0) / (m.index + 1.0), 2);
    double f = (1.0 - m.metallic) * f0 + m.metallic * ((m.color.x + m.color.y + m.color.z) / 3.0);
    f = f * (1.0 - 0.2) + 1.0 * 0.2;  // glm::mix_scalar(f, 1.0, 0.2)
    const double eta_t = dot(wo, n) > 0.0 ? m.index : 1.0 / m.index;

    auto beckmann = [&](Rng& r) {
        const double theta = std::atan(std::sqrt(m2 * -std::log(r.gen_f64())));
        const double sin_t = std::sin(theta), cos_t = std::cos(theta);
        double x, y;
        r.unit_circle(x, y);
        const V3 h = v3(x * sin_t, y * sin_t, cos_t);
        return local_to_world(n).apply(h);
    };
    auto beckmann_pdf = [&](const V3& h) {
        const double cos_t = std::fabs(dot(h, n));
        const double sin_t = std::sqrt(1.0 - cos_t * cos_t);
        return (1.0 / (PI * m2 * powi(cos_t, 3))) * std::exp(-powi(sin_t / cos_t, 2) / m2);
    };

    V3 wi;
    if (rng.gen_bool(f)) {
        const V3 h = beckmann(rng);
        wi = -(wo - 2.0 * dot(h, wo) * h);  // -glm::reflect_vec(wo, h)
    } else if (!m.transparent) {
        double x, y;
        rng.unit_disc(x, y);
        const double z = std::sqrt(1.0 - x * x - y * y);
        wi = local_to_world(n).apply(v3(x, y, z));
    } else {
        const V3 h = beckmann(rng);
        const double cos_to = dot(h, wo);
        const V3 wo_perp = wo - h * cos_to;
        const V3 wi_perp = -wo_perp / eta_t;
        const double sin2_ti = length2(wi_perp);
        if (sin2_ti > 1.0) return false;  // TIR -> None: the path ends
        const double cos_ti = std::sqrt(1.0 - sin2_ti);
        wi = -signum(cos_to) * cos_ti * h + wi_perp;
    }

    double p = 0.0;
    {
        const V3 h = normalize(wi + wo);
        const double p_h = beckmann_pdf(h);
        p += f * p_h / (4.0 * std::fabs(dot(h, wo)));
    }
    if (!m.transparent) {
        p += (1.0 - f) * std::fmax(dot(wi, n), 0.0) * (1.0 / PI);
    } else if (!std::signbit(dot(wo, n)) != !std::signbit(dot(wi, n))) {
        const V3 h = normalize(wi * eta_t + wo);
        const double p_h = beckmann_pdf(h);
        const double h_dot_wo = dot(h, wo);
        const double h_dot_wi = dot(h, wi);
        const double jacobian = std::fabs(h_dot_wo) / powi(eta_t * h_dot_wi + h_dot_wo, 2);
        p += (1.0 - f) * p_h * jacobian;
    } else {
        p += 0.0;
    }
    wi_out = wi;
    pdf_out = p;
    return true;
}

// ---------------------------------------------- src/environment.rs ---------
struct Environment {
    int kind = 0;
    V3 color = {0, 0, 0};
    uint32_t width = 0, height = 0;
    const double* buf = nullptr;
    V3 texel(uint32_t x, uint32_t y) const {
        // The reference indexes x0+1 / y0+1 unclamped (:39-51): out-of-bounds panic at
        // polar = pi, row wrap at azimuth = 2pi.  Clamped here (and on the GPU) -- a
        // measure-zero deviation, SURVEY Appendix A #15.
        x = std::min(x, width - 1);
        y = std::min(y, height - 1);
        const double* p = buf + 3 * ((size_t)y * width + x);
        return v3(p[0], p[1], p[2]);
    }
    V3 get_color(const V3& dir_in, Counters& c) const {  // :25-52,72-77
        if (kind == 0) return color;
        c.env_lookups++;
        const V3 dir = normalize(dir_in);
        const double azimuth = std::atan2(dir.z, dir.x) + PI;
        const double polar = std::acos(dir.y);
        const double x = azimuth / (2.0 * PI) * (double)(width - 1);
        const double y = polar / PI * (double)(height - 1);
        const uint32_t x0 = std::min((uint32_t)x, width - 1);
        const uint32_t y0 = std::min((uint32_t)y, height - 1);
        const double ax = x - (double)x0;
        const double ay = y - (double)y0;
        return mix(mix(texel(x0, y0), texel(x0 + 1, y0), ax), mix(texel(x0, y0 + 1), texel(x0 + 1, y0 + 1), ax), ay);
    }
};

// ----------------------------------------- src/object.rs, src/light.rs -----
struct Object {
    ShapePtr shape;
    Material material;
};
struct Light {
    int kind;
    V3 color, vec;
    Object object;
    // illuminate :23-47 -> (intensity, wi, dist)
    void illuminate(const V3& world_pos, Rng& rng, V3& intensity, V3& wi, double& dist) const {
        switch (kind) {
            case RPTB_LIGHT_AMBIENT:
                intensity = color; wi = v3(0, 0, 0); dist = 0.0;
                return;
            case RPTB_LIGHT_POINT: {
                const V3 disp = vec - world_pos;
                const double len = length(disp);
                intensity = color / (len * len); wi = disp / len; dist = len;
                return;
            }
            case RPTB_LIGHT_DIRECTIONAL:
                intensity = color; wi = -normalize(vec); dist = INF;
                return;
            default: {
                V3 v, n;
                double p;
                object.shape->sample(world_pos, rng, v, n, p);
                const V3 disp = v - world_pos;
                const double len = length(disp);
                const double cosine = std::fmax(-dot(disp, n), 0.0) / len;
                const double surface_area = std::fmax(cosine, 0.0) / (len * len);
                intensity = object.material.color * object.material.emittance * surface_area / p;
                wi = disp / len;
                dist = len;
                return;
            }
        }
    }
};

struct Scene {
    std::vector<Object> objects;
    std::vector<Light> lights;
    Environment environment;
};

// --------------------------------------------------- src/camera.rs ---------
struct Camera {
    V3 eye, direction, up;
    double fov, aperture, focal_distance;
    Ray cast_ray(double x, double y, Rng& rng) const {  // :64-81
        const double d = 1.0 / std::tan(fov / 2.0);
        const V3 right = normalize(cross(direction, up));
        V3 origin = eye;
        V3 new_dir = d * direction + x * right + y * up;
        if (aperture > 0.0) {
            const V3 focal_point = origin + normalize(new_dir) * focal_distance;
            double ax, ay;
            rng.unit_disc(ax, ay);
            origin += (ax * right + ay * up) * aperture;
            new_dir = focal_point - origin;
        }
        return Ray{origin, normalize(new_dir)};
    }
};

// ------------------------------------------------- src/renderer.rs ---------
const double EPSILON = 1e-12;        // :14
const double FIREFLY_CLAMP = 100.0;  // :15

struct Renderer {
    const Scene* scene;
    Camera camera;
    uint32_t width, height, max_bounces;
    double exposure_value;

    // get_closest_hit :211-220
    const Object* get_closest_hit(const Ray& ray, HitRecord& h, Counters& c) const {
        c.rays++;
        h = HitRecord();
        const Object* hit = nullptr;
        for (const Object& object : scene->objects) {
            c.object_tests++;
            if (object.shape->intersect(ray, EPSILON, h, c)) hit = &object;
        }
        return hit;
    }
    // sample_lights :177-204
    V3 sample_lights(const Material& material, const V3& pos, const V3& n, const V3& wo, Rng& rng, Counters& c) const {
        V3 color = v3(0, 0, 0);
        for (const Light& light : scene->lights) {
            if (light.kind == RPTB_LIGHT_AMBIENT) {
                color += cmul(light.color, material.color);
            } else {
                V3 intensity, wi;
                double dist_to_light;
                light.illuminate(pos, rng, intensity, wi, dist_to_light);
                HitRecord r;
                const Object* o = get_closest_hit(Ray{pos, wi}, r, c);
                if (o == nullptr || r.time > dist_to_light) {
                    const V3 f = bsdf(material, n, wo, wi);
                    color += cmul(f, intensity) * dot(wi, n);
                }
            }
        }
        return color;
    }
    // trace_ray :145-174
    V3 trace_ray(const Ray& ray, uint32_t num_bounces, Rng& rng, Counters& c) const {
        c.segments++;
        HitRecord h;
        const Object* object = get_closest_hit(ray, h, c);
        if (object == nullptr) return scene->environment.get_color(ray.dir, c);
        if (h.on_mesh) c.mesh_hits++;
        const V3 world_pos = ray.at(h.time);
        const Material& material = object->material;
        const V3 wo = -normalize(ray.dir);
        V3 color = material.emittance * material.color;
        color += sample_lights(material, world_pos, h.normal, wo, rng, c);
        if (num_bounces < max_bounces) {
            V3 wi;
            double pdf;
            if (sample_f(material, h.normal, wo, rng, wi, pdf)) {
                const V3 f = bsdf(material, h.normal, wo, wi);
                const Ray next{world_pos, wi};
                const V3 indirect = 1.0 / pdf * cmul(f, trace_ray(next, num_bounces + 1, rng, c)) * std::fabs(dot(wi, h.normal));
                // f64::min drops NaN in favour of the other operand
                color.x += std::fmin(indirect.x, FIREFLY_CLAMP);
                color.y += std::fmin(indirect.y, FIREFLY_CLAMP);
                color.z += std::fmin(indirect.z, FIREFLY_CLAMP);
            }
        }
        return color;
    }
    // get_color :131-142 -- one Philox stream per (pixel, sample) instead of one StdRng per row
    V3 get_color(uint32_t x, uint32_t y, uint32_t iterations, uint64_t seed, uint64_t first_sample, Counters& c) const {
        const double dim = (double)std::max(width, height);
        const double xn = ((double)(2 * x + 1) - (double)width) / dim;
        const double yn = ((double)(2 * (height - y) - 1) - (double)height) / dim;
        V3 color = v3(0, 0, 0);
        for (uint32_t i = 0; i < iterations; i++) {
            Rng rng(seed, y * width + x, first_sample + i);
            const double dx = rng.gen_range(-1.0 / dim, 1.0 / dim);
            const double dy = rng.gen_range(-1.0 / dim, 1.0 / dim);
            color += trace_ray(camera.cast_ray(xn + dx, yn + dy, rng), 0, rng, c);
        }
        return color / (double)iterations * std::pow(2.0, exposure_value);
    }
};

// ------------------------------------------------ desc -> oracle scene -----
Material to_material(const rptb_material& m) {
    return Material{v3(m.color[0], m.color[1], m.color[2]), m.index, m.roughness, m.metallic, m.emittance, m.transparent != 0};
}

void flatten_tree(const KdNode& n, std::vector<rptb_kdnode>& nodes, std::vector<uint32_t>& refs, uint32_t depth,
                  uint32_t& max_depth, uint32_t& max_leaf) {
    const size_t me = nodes.size();
    nodes.push_back(rptb_kdnode{});
    max_depth = std::max(max_depth, depth);
    if (n.kind == 3) {
        nodes[me].kind = 3;
        nodes[me].first_ref = (uint32_t)refs.size();
        nodes[me].num_refs = (uint32_t)n.indices.size();
        max_leaf = std::max(max_leaf, (uint32_t)n.indices.size());
        for (size_t i : n.indices) refs.push_back((uint32_t)i);
        return;
    }
    nodes[me].kind = (uint32_t)n.kind;
    nodes[me].split = n.value;
    nodes[me].left = (uint32_t)nodes.size();
    flatten_tree(*n.left, nodes, refs, depth + 1, max_depth, max_leaf);
    nodes[me].right = (uint32_t)nodes.size();
    flatten_tree(*n.right, nodes, refs, depth + 1, max_depth, max_leaf);
}

std::unique_ptr<KdNode> unflatten_tree(const rptb_kdnode* nodes, const uint32_t* refs, uint32_t idx) {
    auto n = std::make_unique<KdNode>();
    const rptb_kdnode& s = nodes[idx];
    n->kind = (int)s.kind;
    if (s.kind == 3) {
        for (uint32_t i = 0; i < s.num_refs; i++) n->indices.push_back(refs[s.first_ref + i]);
    } else {
        n->value = s.split;
        n->left = unflatten_tree(nodes, refs, s.left);
        n->right = unflatten_tree(nodes, refs, s.right);
    }
    return n;
}

std::vector<Triangle> to_triangles(const double* tris, uint64_t n) {
    std::vector<Triangle> out(n);
    for (uint64_t i = 0; i < n; i++) {
        const double* t = tris + 18 * i;
        out[i] = Triangle{v3(t[0], t[1], t[2]), v3(t[3], t[4], t[5]), v3(t[6], t[7], t[8]),
                          v3(t[9], t[10], t[11]), v3(t[12], t[13], t[14]), v3(t[15], t[16], t[17])};
    }
    return out;
}

// Meshes are built once per scene and shared (Arc<Mesh>, examples/fractal_teapots.rs:48-49).
struct ShapeFactory {
    const rptb_scene_desc& d;
    bool brute_force;
    std::vector<ShapePtr> meshes;
    ShapeFactory(const rptb_scene_desc& desc, bool bf) : d(desc), brute_force(bf), meshes(desc.nmeshes) {}

    ShapePtr mesh(uint32_t index) {
        if (meshes[index]) return meshes[index];
        const rptb_mesh& m = d.meshes[index];
        auto k = std::make_shared<KdTree>();
        k->objects = to_triangles(m.tris, m.ntris);
        k->init_bounds();
        k->brute_force = brute_force;
        if (!brute_force) {
            if (m.nodes != nullptr) k->root = unflatten_tree(m.nodes, m.refs, 0);
            else k->build();
        }
        meshes[index] = k;
        return k;
    }
    ShapePtr group(uint32_t index) {  // KdTree::new(Vec<Box<dyn Bounded>>)
        const rptb_group& g = d.groups[index];
        auto k = std::make_shared<ShapeTree>();
        for (uint64_t i = 0; i < g.nchildren; i++) k->objects.push_back(shape(g.children[i]));
        k->init_bounds();
        k->brute_force = brute_force;
        if (!brute_force) {
            if (g.nodes != nullptr) k->root = unflatten_tree(g.nodes, g.refs, 0);
            else k->build();
        }
        return k;
    }
    ShapePtr shape(const rptb_object& o) {
        ShapePtr s;
        switch (o.kind) {
            case RPTB_SHAPE_SPHERE: s = std::make_shared<Sphere>(); break;
            case RPTB_SHAPE_PLANE: {
                auto p = std::make_shared<Plane>();
                p->normal = v3(o.plane_normal[0], o.plane_normal[1], o.plane_normal[2]);
                p->value = o.plane_value;
                s = p;
                break;
            }
            case RPTB_SHAPE_CUBE: s = std::make_shared<Cube>(); break;
            case RPTB_SHAPE_MONOMIAL: {
                auto p = std::make_shared<MonomialSurface>();
                p->height = o.monomial_height;
                p->exp = o.monomial_exp;
                s = p;
                break;
            }
            case RPTB_SHAPE_GROUP: s = group(o.mesh); break;
            default: s = mesh(o.mesh);
        }
        if (o.has_transform) {
            M4 t;
            std::memcpy(t.m, o.transform, sizeof(t.m));
            s = std::make_shared<Transformed>(s, t);
        }
        return s;
    }
};

std::unique_ptr<Scene> to_scene(const rptb_scene_desc& d, bool brute_force) {
    auto sc = std::make_unique<Scene>();
    ShapeFactory factory(d, brute_force);
    for (uint32_t i = 0; i < d.nobjects; i++) {
        Object o;
        o.shape = factory.shape(d.objects[i]);
        o.material = to_material(d.materials[d.objects[i].material]);
        sc->objects.push_back(std::move(o));
    }
    for (uint32_t i = 0; i < d.nlights; i++) {
        const rptb_light& l = d.lights[i];
        Light li;
        li.kind = (int)l.kind;
        li.color = v3(l.color[0], l.color[1], l.color[2]);
        li.vec = v3(l.vec[0], l.vec[1], l.vec[2]);
        if (l.kind == RPTB_LIGHT_OBJECT) {
            li.object.shape = factory.shape(l.object);
            li.object.material = to_material(d.materials[l.object.material]);
        }
        sc->lights.push_back(std::move(li));
    }
    sc->environment.kind = (int)d.environment.kind;
    sc->environment.color = v3(d.environment.color[0], d.environment.color[1], d.environment.color[2]);
    sc->environment.width = d.environment.width;
    sc->environment.height = d.environment.height;
    sc->environment.buf = d.environment.texels;
    return sc;
}

void put_stats(const Counters& c, rptb_stats* s) {
    if (!s) return;
    std::memset(s, 0, sizeof(*s));
    s->segments = c.segments; s->rays = c.rays; s->node_visits = c.node_visits;
    s->tri_tests = c.tri_tests; s->mesh_hits = c.mesh_hits; s->env_lookups = c.env_lookups;
    s->object_tests = c.object_tests;
}

}  // namespace

// ============================================================== C API ======
extern "C" {

struct oracle_scene {
    std::unique_ptr<Scene> scene;
};

// Build once, render many times (the kd-tree build of a large mesh is slow).
oracle_scene* oracle_scene_create(const rptb_scene_desc* desc, int brute_force) {
    auto* s = new oracle_scene();
    s->scene = to_scene(*desc, brute_force != 0);
    return s;
}
void oracle_scene_destroy(oracle_scene* s) { delete s; }

int oracle_hardware_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// Renderer::sample (src/renderer.rs:117-129): rows in parallel (OpenMP stands for rayon).
int oracle_render(const oracle_scene* s, const rptb_camera* cam, const rptb_render_params* p, double* out_rgb,
                  rptb_stats* stats, int nthreads) {
    Renderer r;
    r.scene = s->scene.get();
    r.camera = Camera{v3(cam->eye[0], cam->eye[1], cam->eye[2]), v3(cam->direction[0], cam->direction[1], cam->direction[2]),
                      v3(cam->up[0], cam->up[1], cam->up[2]), cam->fov, cam->aperture, cam->focal_distance};
    r.width = p->width; r.height = p->height; r.max_bounces = p->max_bounces; r.exposure_value = p->exposure_value;
    Counters total;
    // Pixel tiles (16x8, round-robin) of other shards are left zero, mirroring the multi-GPU
    // partition of the product (rpt_b200/distributed.py); shard_count <= 1 renders everything.
    const uint32_t shard_count = p->shard_count ? p->shard_count : 1;
    const uint32_t tiles_x = (p->width + 15) / 16;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        Counters local;
#pragma omp for schedule(dynamic, 1)
        for (int64_t y = 0; y < (int64_t)p->height; y++) {
            for (uint32_t x = 0; x < p->width; x++) {
                double* o = out_rgb + 3 * ((size_t)y * p->width + x);
                const uint32_t tile = ((uint32_t)y / 8) * tiles_x + x / 16;
                if (tile % shard_count != p->shard_index) {
                    o[0] = o[1] = o[2] = 0.0;
                    continue;
                }
                const V3 c = r.get_color(x, (uint32_t)y, p->iterations, p->seed, p->first_sample, local);
                o[0] = c.x; o[1] = c.y; o[2] = c.z;
            }
        }
#pragma omp critical
        total.add(local);
    }
    put_stats(total, stats);
    return 0;
}

// Renderer::get_closest_hit (src/renderer.rs:211-220) for n world rays.
int oracle_closest_hit(const oracle_scene* s, const double* rays, uint64_t n, double t_min, double* out_t,
                       int32_t* out_object, double* out_normal, rptb_stats* stats) {
    Counters total;
#pragma omp parallel
    {
        Counters c;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < (int64_t)n; i++) {
            const Ray ray{v3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), v3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5])};
            HitRecord h;
            int32_t hit = -1;
            c.rays++;
            for (size_t k = 0; k < s->scene->objects.size(); k++)
                if (s->scene->objects[k].shape->intersect(ray, t_min, h, c)) hit = (int32_t)k;
            out_t[i] = hit >= 0 ? h.time : INF;
            out_object[i] = hit;
            if (out_normal) {
                out_normal[3 * i] = h.normal.x; out_normal[3 * i + 1] = h.normal.y; out_normal[3 * i + 2] = h.normal.z;
            }
        }
#pragma omp critical
        total.add(c);
    }
    put_stats(total, stats);
    return 0;
}

// Material::bsdf (src/material.rs:125-210): dirs = n x (n, wo, wi).
void oracle_bsdf(const rptb_material* m, const double* dirs, uint64_t n, double* out) {
    const Material mat = to_material(*m);
    for (uint64_t i = 0; i < n; i++) {
        const double* d = dirs + 9 * i;
        const V3 f = bsdf(mat, v3(d[0], d[1], d[2]), v3(d[3], d[4], d[5]), v3(d[6], d[7], d[8]));
        out[3 * i] = f.x; out[3 * i + 1] = f.y; out[3 * i + 2] = f.z;
    }
}

// Material::sample_f (src/material.rs:224-313): dirs = n x (n, wo); stream i = Philox(seed, pixel=i, sample=0).
void oracle_sample_f(const rptb_material* m, const double* dirs, uint64_t n, uint64_t seed, double* out_wi, double* out_pdf) {
    const Material mat = to_material(*m);
    for (uint64_t i = 0; i < n; i++) {
        const double* d = dirs + 6 * i;
        Rng rng(seed, (uint32_t)i, 0);
        V3 wi = v3(0, 0, 0);
        double pdf = -1.0;
        if (!sample_f(mat, v3(d[0], d[1], d[2]), v3(d[3], d[4], d[5]), rng, wi, pdf)) {
            wi = v3(0, 0, 0);
            pdf = -1.0;
        }
        out_wi[3 * i] = wi.x; out_wi[3 * i + 1] = wi.y; out_wi[3 * i + 2] = wi.z;
        out_pdf[i] = pdf;
    }
}

// Light::illuminate (src/light.rs:23-47) of light `index` at n positions; stream i = Philox(seed, i, 0).
void oracle_illuminate(const oracle_scene* s, uint32_t index, const double* pos, uint64_t n, uint64_t seed,
                       double* out_intensity, double* out_wi, double* out_dist) {
    const Light& l = s->scene->lights[index];
    for (uint64_t i = 0; i < n; i++) {
        Rng rng(seed, (uint32_t)i, 0);
        V3 I, wi;
        double dist;
        l.illuminate(v3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), rng, I, wi, dist);
        out_intensity[3 * i] = I.x; out_intensity[3 * i + 1] = I.y; out_intensity[3 * i + 2] = I.z;
        out_wi[3 * i] = wi.x; out_wi[3 * i + 1] = wi.y; out_wi[3 * i + 2] = wi.z;
        out_dist[i] = dist;
    }
}

// KdTree::new -> construct (src/kdtree.rs:108-119,235-355), serialised in DFS pre-order.
static int export_tree(const std::vector<BoundingBox>& boxes, rptb_kdtree_out* out);

int oracle_build_kdtree(const double* tris, uint64_t ntris, rptb_kdtree_out* out) {
    const std::vector<Triangle> objects = to_triangles(tris, ntris);
    std::vector<BoundingBox> boxes;
    for (const Triangle& t : objects) boxes.push_back(t.bounding_box());
    return export_tree(boxes, out);
}

// construct over caller-supplied boxes (6 doubles each: p_min, p_max)
int oracle_build_kdtree_boxes(const double* in, uint64_t n, rptb_kdtree_out* out) {
    std::vector<BoundingBox> boxes(n);
    for (uint64_t i = 0; i < n; i++)
        boxes[i] = BoundingBox{v3(in[6 * i], in[6 * i + 1], in[6 * i + 2]), v3(in[6 * i + 3], in[6 * i + 4], in[6 * i + 5])};
    return export_tree(boxes, out);
}

// Bounded::bounding_box of a described shape (out: p_min, p_max); returns 0 if it has none (Plane)
int oracle_shape_bounds(const rptb_scene_desc* desc, const rptb_object* o, double* out) {
    ShapeFactory factory(*desc, true);
    BoundingBox b;
    if (!factory.shape(*o)->bounding_box(b)) return 0;
    out[0] = b.p_min.x; out[1] = b.p_min.y; out[2] = b.p_min.z;
    out[3] = b.p_max.x; out[4] = b.p_max.y; out[5] = b.p_max.z;
    return 1;
}

static int export_tree(const std::vector<BoundingBox>& boxes, rptb_kdtree_out* out) {
    std::vector<size_t> idx(boxes.size());
    for (size_t i = 0; i < idx.size(); i++) idx[i] = i;
    const std::unique_ptr<KdNode> root = construct(boxes, std::move(idx));
    std::vector<rptb_kdnode> nodes;
    std::vector<uint32_t> refs;
    uint32_t depth = 0, max_leaf = 0;
    flatten_tree(*root, nodes, refs, 0, depth, max_leaf);
    out->nnodes = nodes.size();
    out->nrefs = refs.size();
    out->depth = depth;
    out->max_leaf = max_leaf;
    out->nodes = (rptb_kdnode*)std::malloc(sizeof(rptb_kdnode) * std::max<size_t>(nodes.size(), 1));
    out->refs = (uint32_t*)std::malloc(sizeof(uint32_t) * std::max<size_t>(refs.size(), 1));
    std::memcpy(out->nodes, nodes.data(), sizeof(rptb_kdnode) * nodes.size());
    std::memcpy(out->refs, refs.data(), sizeof(uint32_t) * refs.size());
    return 0;
}
void oracle_free_kdtree(rptb_kdtree_out* out) {
    std::free(out->nodes);
    std::free(out->refs);
    out->nodes = nullptr;
    out->refs = nullptr;
}

// hex_color / color_bytes (src/color.rs:10-23)
void oracle_hex_color(uint32_t x, double* out) {
    const double r = (double)((x >> 16) & 0xff) / 255.0;
    const double g = (double)((x >> 8) & 0xff) / 255.0;
    const double b = (double)(x & 0xff) / 255.0;
    out[0] = std::pow(r, 2.2); out[1] = std::pow(g, 2.2); out[2] = std::pow(b, 2.2);
}
void oracle_color_bytes(const double* c, uint8_t* out) {
    for (int i = 0; i < 3; i++)
        out[i] = (uint8_t)(std::pow(std::fmin(std::fmax(c[i], 0.0), 1.0), 1.0 / 2.2) * 255.0);  // `as u8` truncates
}

// Buffer::image (src/buffer.rs:43-56,75-93) for a buffer of nbatches entries per pixel whose
// per-pixel sums are `sums` (every pixel holds the same number of entries after add_samples).
void oracle_film_resolve(const double* sums, uint32_t nbatches, uint32_t width, uint32_t height, uint32_t radius, uint8_t* out) {
    for (uint32_t y = 0; y < height; y++)
        for (uint32_t x = 0; x < width; x++) {
            double color[3] = {0, 0, 0};
            uint64_t count = 0;
            const uint32_t i0 = x >= radius ? x - radius : 0, j0 = y >= radius ? y - radius : 0;  // saturating_sub
            for (uint32_t i = i0; i <= x + radius; i++)
                for (uint32_t j = j0; j <= y + radius; j++)
                    if (i < width && j < height) {
                        const double* p = sums + 3 * ((size_t)j * width + i);
                        color[0] += p[0]; color[1] += p[1]; color[2] += p[2];
                        count += nbatches;
                    }
            const double c[3] = {color[0] / (double)count, color[1] / (double)count, color[2] / (double)count};
            oracle_color_bytes(c, out + 3 * ((size_t)y * width + x));
        }
}

// Buffer::variance (src/buffer.rs:59-73): batches = nbatches x (width*height*3) entries.
double oracle_variance(const double* batches, uint32_t nbatches, uint64_t npixels) {
    double variance = 0.0, count = 0.0;
    for (uint64_t p = 0; p < npixels; p++) {
        double mean[3] = {0, 0, 0};
        for (uint32_t b = 0; b < nbatches; b++)
            for (int k = 0; k < 3; k++) mean[k] += batches[((size_t)b * npixels + p) * 3 + k];
        for (int k = 0; k < 3; k++) mean[k] /= (double)nbatches;
        double ss = 0.0;
        for (uint32_t b = 0; b < nbatches; b++)
            for (int k = 0; k < 3; k++) {
                const double d = batches[((size_t)b * npixels + p) * 3 + k] - mean[k];
                ss += d * d;
            }
        variance += ss / ((double)nbatches - 1.0);
        count += 1.0;
    }
    return variance / count;
}

// Raw Philox blocks, for checking the device generator bit for bit.
void oracle_philox(uint64_t seed, uint32_t pixel, uint64_t sample, uint32_t first_block, uint32_t nblocks, uint32_t* out) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (uint32_t b = 0; b < nblocks; b++) {
        const uint32_t ctr[4] = {first_block + b, pixel, (uint32_t)sample, (uint32_t)(sample >> 32)};
        Philox::block(ctr, key, out + 4 * b);
    }
}

// The distributions of rand/rand_distr over the stream, for the RNG-parity test:
// kind 0 gen_f64, 1 gen_range(-1,1), 2 unit_disc (2 values), 3 unit_circle (2 values), 4 uniform_usize(n=param)
void oracle_draws(uint64_t seed, uint32_t pixel, uint32_t kind, uint64_t param, uint32_t count, double* out) {
    Rng rng(seed, pixel, 0);
    for (uint32_t i = 0; i < count; i++) {
        switch (kind) {
            case 0: out[i] = rng.gen_f64(); break;
            case 1: out[i] = rng.gen_range(-1.0, 1.0); break;
            case 2: rng.unit_disc(out[2 * i], out[2 * i + 1]); break;
            case 3: rng.unit_circle(out[2 * i], out[2 * i + 1]); break;
            default: out[i] = (double)rng.uniform_usize(param); break;
        }
    }
}

}  // extern "C"
