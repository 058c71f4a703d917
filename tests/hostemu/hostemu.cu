// hostemu.cu -- TEST INFRASTRUCTURE, NOT PRODUCT.  Never linked into librpt_b200.so, never loaded by
// rpt_b200/*: only tests/test_hostemu.py builds and loads it.
//
// The geometry / light-sampling device functions of rpt_b200/csrc (geometry.cuh, shading.cuh: the code
// the CUDA kernels inline) compiled for the HOST with RPTB_HOST_EMU, over the very arrays
// rptb_scene_create would upload (flatten.h), so that their control flow -- kd traversal, the kd-tree of
// shapes, MonomialSurface, Transformed, finalize_hit, shape_sample -- can be checked against the oracle
// in the build container, which has no GPU.  It proves nothing about the kernels' scheduling, fast-math
// or memory behaviour; the `-m gpu` parity tests through the C ABI stay the gate.
//
// Built by `make hostemu` with nvcc (host pass) and -ffp-contract=off so Real = double rounds like the
// f64 parity kernels (-fmad=false).
#ifndef RPTB_HOST_EMU
#error "compile with -DRPTB_HOST_EMU"
#endif
#include <cstring>
#include <new>
#include <string>

#include "../../rpt_b200/csrc/flatten.h"
#include "../../rpt_b200/csrc/shading.cuh"

using namespace rptb;

namespace {

// bind_scene's "uploader": the arrays stay where they are
struct InPlace {
    uint64_t total = 0;
    template <class T>
    bool operator()(std::vector<T>& host, const T** where) {
        *where = host.empty() ? nullptr : host.data();
        total += host.size() * sizeof(T);
        return true;
    }
    uint64_t bytes() const { return total; }
};

template <class R>
void closest_hits(const SceneView<R>& sv, const double* rays, uint64_t n, double tmin, double* out_t, int32_t* out_obj,
                  double* out_n, rptb_stats* stats) {
    unsigned long long nv = 0, tt = 0, ot = 0;
#pragma omp parallel for schedule(static) reduction(+ : nv, tt, ot)
    for (int64_t i = 0; i < (int64_t)n; i++) {  // body of closest_hit_kernel (integrator.cuh)
        const double* r = rays + 6 * i;
        const Vec3<R> o = {(R)r[0], (R)r[1], (R)r[2]};
        const Vec3<R> d = {(R)r[3], (R)r[4], (R)r[5]};
        TravStats ts = {0, 0, 0};
        Hit<R> h;
        h.t = M<R>::inf();
        closest_hit<R, true, F_EVERY>(sv, o, d, (R)tmin, false, h, ts);
        out_obj[i] = h.obj;
        out_t[i] = h.obj >= 0 ? (double)h.t : (double)INFINITY;
        if (out_n) {
            Vec3<R> nn = {(R)0, (R)0, (R)0};
            if (h.obj >= 0) nn = finalize_hit<R, F_EVERY>(sv, sv.objects[h.obj], o, d, h).n;
            out_n[3 * i] = (double)nn.x;
            out_n[3 * i + 1] = (double)nn.y;
            out_n[3 * i + 2] = (double)nn.z;
        }
        nv += ts.node_visits;
        tt += ts.tri_tests;
        ot += ts.object_tests;
    }
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->rays = n;
        stats->node_visits = nv;
        stats->tri_tests = tt;
        stats->object_tests = ot;
    }
}

template <class R>
void illuminations(const SceneView<R>& sv, uint32_t light, const double* pos, uint64_t n, uint64_t seed, double* out_i,
                   double* out_wi, double* out_dist) {
    for (uint64_t i = 0; i < n; i++) {  // same stream as oracle_illuminate: key (seed), pixel i, sample 0
        Rng<R> rng;
        rng.init(seed, (uint32_t)i, 0);
        Vec3<R> I, wi;
        R dist;
        illuminate<R, F_EVERY>(sv, sv.lights[light], mk((R)pos[3 * i], (R)pos[3 * i + 1], (R)pos[3 * i + 2]), rng, I, wi, dist);
        out_i[3 * i] = (double)I.x; out_i[3 * i + 1] = (double)I.y; out_i[3 * i + 2] = (double)I.z;
        out_wi[3 * i] = (double)wi.x; out_wi[3 * i + 1] = (double)wi.y; out_wi[3 * i + 2] = (double)wi.z;
        out_dist[i] = (double)dist;
    }
}

}  // namespace

extern "C" {

struct hostemu_scene {
    HostScene hs;
    SceneView<float> v32;
    SceneView<double> v64;
    int features;
};

// Returns NULL and writes the flattener's message to `err` on a bad description.
hostemu_scene* hostemu_scene_create(const rptb_scene_desc* desc, char* err, size_t errlen) {
    hostemu_scene* s = new (std::nothrow) hostemu_scene();
    if (!s) return nullptr;
    std::memset(&s->v32, 0, sizeof(s->v32));
    std::memset(&s->v64, 0, sizeof(s->v64));
    std::string msg;
    if (flatten_scene(desc, s->hs, msg) != RPTB_OK) {
        if (err && errlen) {
            std::strncpy(err, msg.c_str(), errlen - 1);
            err[errlen - 1] = 0;
        }
        delete s;
        return nullptr;
    }
    InPlace put;
    uint64_t f32_bytes = 0;
    bind_scene(s->hs, put, false, s->v32, s->v64, f32_bytes);
    s->features = s->hs.features;
    return s;
}

void hostemu_scene_destroy(hostemu_scene* s) { delete s; }

int hostemu_scene_features(const hostemu_scene* s) { return s->features; }

// precision: 0 = Real float, 1 = Real double (rptb_precision)
int hostemu_closest_hit(const hostemu_scene* s, const double* rays, uint64_t n, double t_min, uint32_t precision,
                        double* out_t, int32_t* out_object, double* out_normal, rptb_stats* stats) {
    if (precision == RPTB_PRECISION_F64) closest_hits<double>(s->v64, rays, n, t_min, out_t, out_object, out_normal, stats);
    else closest_hits<float>(s->v32, rays, n, t_min, out_t, out_object, out_normal, stats);
    return 0;
}

int hostemu_illuminate(const hostemu_scene* s, uint32_t light, const double* pos, uint64_t n, uint64_t seed, uint32_t precision,
                       double* out_intensity, double* out_wi, double* out_dist) {
    if (precision == RPTB_PRECISION_F64) illuminations<double>(s->v64, light, pos, n, seed, out_intensity, out_wi, out_dist);
    else illuminations<float>(s->v32, light, pos, n, seed, out_intensity, out_wi, out_dist);
    return 0;
}

}  // extern "C"
