// vec.cuh -- small vector algebra + per-precision math policy for the device code.
//
// `Real = float` is the product path; `Real = double` is the parity gate that
// keeps the reference's literal f64 formulas (ekzhang/rpt is f64 throughout,
// src/color.rs:2).  Everything is header-only and __host__ __device__ where that
// is free, so the host flattener can share the types.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define RPTB_HD __host__ __device__ __forceinline__
#ifdef RPTB_HOST_EMU
// tests/hostemu only (test infrastructure, never part of librpt_b200.so): the device functions are
// also compiled for the host, so their control flow can be checked against the oracle without a GPU.
#define RPTB_D __host__ __device__ __forceinline__
#else
#define RPTB_D __device__ __forceinline__
#endif

namespace rptb {

template <class R>
struct Vec3 {
    R x, y, z;
};
template <class R>
RPTB_HD Vec3<R> mk(R x, R y, R z) {
    return Vec3<R>{x, y, z};
}
template <class R>
RPTB_HD Vec3<R> operator+(Vec3<R> a, Vec3<R> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class R>
RPTB_HD Vec3<R> operator-(Vec3<R> a, Vec3<R> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class R>
RPTB_HD Vec3<R> operator-(Vec3<R> a) { return {-a.x, -a.y, -a.z}; }
template <class R>
RPTB_HD Vec3<R> operator*(Vec3<R> a, R s) { return {a.x * s, a.y * s, a.z * s}; }
template <class R>
RPTB_HD Vec3<R> operator*(R s, Vec3<R> a) { return {a.x * s, a.y * s, a.z * s}; }
template <class R>
RPTB_HD Vec3<R> operator/(Vec3<R> a, R s) { return {a.x / s, a.y / s, a.z / s}; }
template <class R>
RPTB_HD Vec3<R> cmul(Vec3<R> a, Vec3<R> b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
template <class R>
RPTB_HD R dot(Vec3<R> a, Vec3<R> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class R>
RPTB_HD Vec3<R> cross(Vec3<R> a, Vec3<R> b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <class R>
RPTB_HD R length2(Vec3<R> a) { return dot(a, a); }
template <class R>
RPTB_HD R comp(const Vec3<R>& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// ---- per-precision math ---------------------------------------------------------
template <class R>
struct M;

template <>
struct M<double> {
    static constexpr bool literal = true;  // follow the reference's formulas verbatim
    static RPTB_HD double inf() { return (double)INFINITY; }
    static RPTB_HD double pi() { return 3.14159265358979323846264338327950288; }
    static RPTB_HD double sqrt(double x) { return ::sqrt(x); }
    static RPTB_HD double abs(double x) { return ::fabs(x); }
    static RPTB_HD double min(double a, double b) { return ::fmin(a, b); }  // drops NaN like f64::min
    static RPTB_HD double max(double a, double b) { return ::fmax(a, b); }
    static RPTB_HD double exp(double x) { return ::exp(x); }
    static RPTB_HD double log(double x) { return ::log(x); }
    static RPTB_HD double pow(double x, double y) { return ::pow(x, y); }
    static RPTB_HD double div(double a, double b) { return a / b; }
    static RPTB_HD double rcp(double a) { return 1.0 / a; }
    static RPTB_HD bool signbit(double x) { return ::signbit(x); }
    static RPTB_HD double copysign(double a, double b) { return ::copysign(a, b); }
    static RPTB_HD bool isnormal(double x) { return ::fabs(x) >= 2.2250738585072014e-308 && ::fabs(x) < (double)INFINITY; }
    static RPTB_HD Vec3<double> normalize(Vec3<double> a) { return a / ::sqrt(dot(a, a)); }
    static RPTB_HD double next_up(double x) { return ::nextafter(x, (double)INFINITY); }
};

template <>
struct M<float> {
    static constexpr bool literal = false;
    static RPTB_HD float inf() { return INFINITY; }
    static RPTB_HD float pi() { return 3.14159265358979323846f; }
    static RPTB_HD float sqrt(float x) { return ::sqrtf(x); }
    static RPTB_HD float abs(float x) { return ::fabsf(x); }
    static RPTB_HD float min(float a, float b) { return ::fminf(a, b); }
    static RPTB_HD float max(float a, float b) { return ::fmaxf(a, b); }
#ifdef __CUDA_ARCH__
    static RPTB_D float exp(float x) { return __expf(x); }
    static RPTB_D float log(float x) { return __logf(x); }
    static RPTB_D float div(float a, float b) { return __fdividef(a, b); }
    static RPTB_D float rcp(float a) { return __frcp_rn(a); }
    static RPTB_D Vec3<float> normalize(Vec3<float> a) { return a * rsqrtf(dot(a, a)); }
#else
    static float exp(float x) { return ::expf(x); }
    static float log(float x) { return ::logf(x); }
    static float div(float a, float b) { return a / b; }
    static float rcp(float a) { return 1.0f / a; }
    static Vec3<float> normalize(Vec3<float> a) { return a * (1.0f / ::sqrtf(dot(a, a))); }
#endif
    static RPTB_HD float pow(float x, float y) { return ::powf(x, y); }
    static RPTB_HD bool signbit(float x) { return ::signbit(x); }
    static RPTB_HD float copysign(float a, float b) { return ::copysignf(a, b); }
    static RPTB_HD bool isnormal(float x) { return ::fabsf(x) >= 1.17549435e-38f && ::fabsf(x) < INFINITY; }
    static RPTB_HD float next_up(float x) { return ::nextafterf(x, INFINITY); }
};

// read-only global load / fast divide (plain equivalents in the host emulation build)
template <class T>
RPTB_D T ldg(const T* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
RPTB_D float fdividef(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdividef(a, b);
#else
    return a / b;
#endif
}

// Rust f64::signum: +1 for +0.0, -1 for -0.0
template <class R>
RPTB_HD R signum(R x) { return M<R>::copysign((R)1, x); }

// Rust f64::powi(5) -> x * (x^2)^2 (square-and-multiply), powi(2) -> x*x, powi(3) -> x * x^2
template <class R>
RPTB_HD R pow5(R x) { const R x2 = x * x; return x * (x2 * x2); }

}  // namespace rptb
