// rng.cuh -- per-path Philox4x32-10 stream and the rand 0.8 / rand_distr 0.4
// distributions the reference draws from:
//   gen::<f64>()      src/material.rs:247, src/shape/mesh.rs:85-90, src/shape/cube.rs:75-76
//   gen_range(a..b)   src/renderer.rs:137-138
//   gen_bool(p)       src/material.rs:264
//   Uniform(0..n)     src/kdtree.rs:140, src/shape/cube.rs:77
//   UnitDisc          src/camera.rs:73, src/shape/sphere.rs:53, src/material.rs:271
//   UnitCircle        src/material.rs:251
// The reference seeds ChaCha12 from OS entropy per row (src/renderer.rs:121); here
// key = (seed_lo, seed_hi), counter = (block, pixel, sample_lo, sample_hi).
// Each block yields two 64-bit draws: (x0 | x1<<32), then (x2 | x3<<32).
#pragma once
#include "vec.cuh"

namespace rptb {

struct Philox {
    uint32_t key0, key1;
    uint32_t block, pixel, samp_lo, samp_hi;
    uint32_t spare_lo, spare_hi;
    bool have_spare;

    RPTB_HD void init(uint64_t seed, uint32_t pix, uint64_t sample) {
        key0 = (uint32_t)seed;
        key1 = (uint32_t)(seed >> 32);
        block = 0;
        pixel = pix;
        samp_lo = (uint32_t)sample;
        samp_hi = (uint32_t)(sample >> 32);
        have_spare = false;
        spare_lo = spare_hi = 0;
    }

    static RPTB_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
        lo = a * b;
        hi = __umulhi(a, b);
#else
        const uint64_t p = (uint64_t)a * b;
        lo = (uint32_t)p;
        hi = (uint32_t)(p >> 32);
#endif
    }

    static RPTB_HD void block10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                uint32_t out[4]) {
#pragma unroll
        for (int i = 0; i < 10; i++) {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(0xD2511F53u, c0, hi0, lo0);
            mulhilo(0xCD9E8D57u, c2, hi1, lo1);
            const uint32_t n0 = hi1 ^ c1 ^ k0;
            const uint32_t n2 = hi0 ^ c3 ^ k1;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }

    // One Philox block costs ~60 instructions; the integrator draws from ~20 call sites, so
    // the block function is kept out of line (a pure function of register arguments: one
    // copy in the instruction cache instead of 20, generator state stays in registers).
#ifdef __CUDA_ARCH__
    static __device__ __noinline__ uint4 block_call(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                    uint32_t k1) {
        uint32_t o[4];
        block10(c0, c1, c2, c3, k0, k1, o);
        return make_uint4(o[0], o[1], o[2], o[3]);
    }
#endif
    RPTB_HD uint64_t refill() {
#ifdef __CUDA_ARCH__
        const uint4 v = block_call(block, pixel, samp_lo, samp_hi, key0, key1);
        const uint32_t o[4] = {v.x, v.y, v.z, v.w};
#else
        uint32_t o[4];
        block10(block, pixel, samp_lo, samp_hi, key0, key1, o);
#endif
        block++;
        spare_lo = o[2];
        spare_hi = o[3];
        have_spare = true;
        return (uint64_t)o[0] | ((uint64_t)o[1] << 32);
    }

    RPTB_HD uint64_t next_u64() {
        if (have_spare) {
            have_spare = false;
            return (uint64_t)spare_lo | ((uint64_t)spare_hi << 32);
        }
        return refill();
    }
};

template <class R>
struct Rng;

// f64: bit-for-bit the oracle's conversions of the 64-bit draws.
template <>
struct Rng<double> {
    Philox p;
    RPTB_HD void init(uint64_t seed, uint32_t pix, uint64_t sample) { p.init(seed, pix, sample); }
    RPTB_HD void bind(uint32_t*, uint32_t) {}
    RPTB_HD void ensure() {}
    template <class W>
    RPTB_HD void ensure(unsigned, uint32_t) {}
    RPTB_HD double gen() { return (double)(p.next_u64() >> 11) * (1.0 / 9007199254740992.0); }
    RPTB_HD double u52() { return (double)(p.next_u64() >> 12) * (1.0 / 4503599627370496.0); }
    // Rng::gen_bool(p): Bernoulli -> u64 < p * 2^64
    RPTB_HD bool bernoulli(double prob) {
        const uint64_t v = p.next_u64();
        if (prob >= 1.0) return true;
        return v < (uint64_t)(prob * 18446744073709551616.0);
    }
    // Rng::gen::<bool>(): the top bit of one draw
    RPTB_HD bool coin() { return (p.next_u64() >> 63) != 0; }
    // Uniform::from(0..n) for usize: widening multiply with rejection zone
    RPTB_HD uint64_t below(uint64_t n) {
        const uint64_t ints_to_reject = (0xFFFFFFFFFFFFFFFFull - n + 1) % n;
        const uint64_t zone = 0xFFFFFFFFFFFFFFFFull - ints_to_reject;
        while (true) {
            const uint64_t v = p.next_u64();
            const uint64_t lo = v * n;
#ifdef __CUDA_ARCH__
            const uint64_t hi = __umul64hi(v, n);
#else
            const uint64_t hi = (uint64_t)(((unsigned __int128)v * n) >> 64);
#endif
            if (lo <= zone) return hi;
        }
    }
};

// f32: only the HIGH 32-bit word of each 64-bit draw is kept -- its top 24 bits are the f64
// value truncated to a float in [0,1), so both precisions see the same stream.  The words
// are produced a block (two draws) at a time into a 4-entry register FIFO by ensure(),
// which the integrator calls where the whole warp is converged: the ~60-instruction
// Philox block then runs once for 32 lanes instead of once per lane per call site.
// gen_bool / Uniform(0..n) decide on the high word alone; the decision differs from the
// 64-bit one with probability <= n * 2^-32 per draw (f32 mode only).
template <>
struct Rng<float> {
    uint32_t key0, key1, block, pixel, samp_lo, samp_hi;
    uint32_t q0, q1, q2, q3;
    uint32_t avail;
    RPTB_HD void init(uint64_t seed, uint32_t pix, uint64_t sample) {
        key0 = (uint32_t)seed;
        key1 = (uint32_t)(seed >> 32);
        block = 0;
        pixel = pix;
        samp_lo = (uint32_t)sample;
        samp_hi = (uint32_t)(sample >> 32);
        q0 = q1 = q2 = q3 = 0;
        avail = 0;
    }
    RPTB_HD void push_block() {  // requires avail <= 2
#ifdef __CUDA_ARCH__
        const uint4 v = Philox::block_call(block, pixel, samp_lo, samp_hi, key0, key1);
        const uint32_t a = v.y, b = v.w;
#else
        uint32_t o[4];
        Philox::block10(block, pixel, samp_lo, samp_hi, key0, key1, o);
        const uint32_t a = o[1], b = o[3];
#endif
        block++;
        if (avail == 0) { q0 = a; q1 = b; }
        else if (avail == 1) { q1 = a; q2 = b; }
        else { q2 = a; q3 = b; }
        avail += 2;
    }
    RPTB_HD void ensure() {
        if (avail <= 2) push_block();
        if (avail <= 2) push_block();
    }
    template <class W>
    RPTB_HD void ensure(unsigned, uint32_t) { ensure(); }  // (a 6-entry FIFO refilled to the slot's expected draws was measured too: Cornell 4 841 vs 5 673)
    RPTB_HD void bind(uint32_t*, uint32_t) {}
    RPTB_HD uint32_t next32() {
        if (avail == 0) push_block();  // rare: a slot consumed more than the FIFO held
        const uint32_t v = q0;
        q0 = q1; q1 = q2; q2 = q3;
        avail--;
        return v;
    }
    RPTB_HD float gen() { return (float)(next32() >> 8) * (1.0f / 16777216.0f); }
    RPTB_HD float u52() { return gen(); }
    RPTB_HD bool bernoulli(float prob) {
        const uint32_t v = next32();
        if (prob >= 1.0f) return true;
        return v < (uint32_t)((uint64_t)((double)prob * 18446744073709551616.0) >> 32);
    }
    RPTB_HD bool coin() { return (next32() >> 31) != 0; }
    RPTB_HD uint64_t below(uint64_t n) {
#ifdef __CUDA_ARCH__
        return (uint64_t)__umulhi(next32(), (uint32_t)n);
#else
        return ((uint64_t)next32() * (uint32_t)n) >> 32;
#endif
    }
};

// The megakernel's f32 generator: the same stream as Rng<float> (high word of every 64-bit draw), buffered in an
// 8-entry ring per thread that lives in SHARED memory on the device (entry i of thread t at ring[i * stride + t]:
// one bank per lane, no conflicts; a ring in registers would need shuffling moves per draw and four more live
// registers in a kernel that already spills).  What this buys (ncu, Cornell, round 2): with the 4-entry register FIFO
// a slot that drew more than the FIFO held refilled it inside whatever rejection loop it was in -- the ~60-instruction
// Philox block ran at 7.4 of 32 lanes and was 27 % of the kernel's warp instructions.  ensure() tops a lane up to
// what it is about to draw where the warp is converged, and all lanes that need a block compute it together.
constexpr uint32_t RNG_RING = 8;
#ifndef RPTB_RNG_FILL
#define RPTB_RNG_FILL 0
#endif
#ifndef RPTB_RNG_FIFO
// 1 = the slot engine keeps the 4-entry register FIFO (Rng<float>), 0 = it draws from the ring below.  Measured on one
// B200 with everything else equal (gpurun r02h, Msamples/s, FIFO / ring topped up to 6 / ring topped up on demand):
// cornell 5 676 / 5 252 / 4 919, glass 20 553 / 14 676 / 15 590, sphere 9 284 / 9 638 / 8 908 -- the ring does raise the
// lanes per Philox instruction (7.4 -> 11.2, ncu) but its shared-memory traffic and vote loops cost more than that saves,
// so the FIFO is the default and the ring is what the vertex-at-once engine (integrator_vx.cuh) uses.
#define RPTB_RNG_FIFO 1
#endif
struct RngRing {
    uint32_t key0, key1, block, pixel, samp_lo, samp_hi;
    uint32_t head, avail;
    uint32_t* ring;     // device: shared memory; host emulation: `own`
    uint32_t stride;
#ifdef __CUDA_ARCH__
    // (a local object of device code only: it never crosses the host/device boundary, so the layouts may differ)
    __device__ void bind(uint32_t* shared_ring, uint32_t shared_stride) { ring = shared_ring; stride = shared_stride; }
#else
    uint32_t own[RNG_RING];
    void bind(uint32_t* shared_ring, uint32_t shared_stride) {
        if (shared_ring) { ring = shared_ring; stride = shared_stride; }
        else { ring = own; stride = 1; }
    }
#endif
    RPTB_HD void init(uint64_t seed, uint32_t pix, uint64_t sample) {
        key0 = (uint32_t)seed;
        key1 = (uint32_t)(seed >> 32);
        block = 0;
        pixel = pix;
        samp_lo = (uint32_t)sample;
        samp_hi = (uint32_t)(sample >> 32);
        head = 0;
        avail = 0;
    }
    RPTB_HD uint32_t& at(uint32_t i) { return ring[(i & (RNG_RING - 1u)) * stride]; }
    RPTB_HD void push_block() {  // requires avail <= RNG_RING - 2
#ifdef __CUDA_ARCH__
        const uint4 v = Philox::block_call(block, pixel, samp_lo, samp_hi, key0, key1);
        const uint32_t a = v.y, b = v.w;
#else
        uint32_t o[4];
        Philox::block10(block, pixel, samp_lo, samp_hi, key0, key1, o);
        const uint32_t a = o[1], b = o[3];
#endif
        block++;
        at(head + avail) = a;
        at(head + avail + 1u) = b;
        avail += 2;
    }
    // W = the warp policy of integrator.cuh (real votes on the device, a single lane in host emulation).  `need` = how
    // many draws this lane expects to take before the next converged point (0 for a lane that will draw nothing): the
    // lanes that are short compute their blocks together, nobody generates draws on speculation -- a path that ends
    // after two draws (a camera ray that leaves the scene) costs one Philox block, not a ring full.
    template <class W>
    RPTB_HD void ensure(unsigned mask, uint32_t need) {
        // (a block is two entries: a lane can take one while avail <= RNG_RING - 2)
#if RPTB_RNG_FILL
        // A/B switch: top every lane up to >= 6 entries whatever it is about to draw
        (void)need;
        while (W::any(mask, avail <= 5u)) {
            if (avail <= RNG_RING - 2u) push_block();
        }
#else
        while (W::any(mask, avail < need && avail <= RNG_RING - 2u)) {
            if (avail < need && avail <= RNG_RING - 2u) push_block();
        }
#endif
    }
    RPTB_HD uint32_t next32() {
        if (avail == 0) push_block();  // rare: a slot consumed more than six draws
        const uint32_t v = at(head);
        head++;
        avail--;
        return v;
    }
    RPTB_HD float gen() { return (float)(next32() >> 8) * (1.0f / 16777216.0f); }
    RPTB_HD float u52() { return gen(); }
    RPTB_HD bool bernoulli(float prob) {
        const uint32_t v = next32();
        if (prob >= 1.0f) return true;
        return v < (uint32_t)((uint64_t)((double)prob * 18446744073709551616.0) >> 32);
    }
    RPTB_HD bool coin() { return (next32() >> 31) != 0; }
    RPTB_HD uint64_t below(uint64_t n) {
#ifdef __CUDA_ARCH__
        return (uint64_t)__umulhi(next32(), (uint32_t)n);
#else
        return ((uint64_t)next32() * (uint32_t)n) >> 32;
#endif
    }
};

// Rng::gen_range(lo..hi) -> UniformFloat::sample_single
template <class R, class RNG>
RPTB_HD R gen_range(RNG& r, R lo, R hi) {
    const R scale = hi - lo;
    while (true) {
        const R v12 = (R)1 + r.u52();
        const R res = v12 * scale + (lo - scale);
        if (res < hi) return res;
    }
}
// Uniform::new(-1, 1).sample
template <class R, class RNG>
RPTB_HD R uniform_pm1(RNG& r) { return r.u52() * (R)2 + (R)(-1); }

// Rng::gen_bool(p).  p >= 1 returns true (the reference's ALWAYS_TRUE case) but still
// consumes one draw so that the number of draws per vertex is material-independent.
template <class R, class RNG>
RPTB_HD bool gen_bool(RNG& r, R prob) { return r.bernoulli(prob); }

// Uniform::from(0..n) for usize
template <class RNG>
RPTB_HD uint64_t uniform_usize(RNG& r, uint64_t n) { return r.below(n); }

// rand_distr::UnitDisc: rejection from the square, boundary inclusive
template <class R, class RNG>
RPTB_HD void unit_disc(RNG& r, R& x, R& y) {
    while (true) {
        x = uniform_pm1<R>(r);
        y = uniform_pm1<R>(r);
        if (x * x + y * y <= (R)1) return;
    }
}
// rand_distr::UnitCircle: von Neumann's method
template <class R, class RNG>
RPTB_HD void unit_circle(RNG& r, R& x, R& y) {
    R x1, x2, sum;
    while (true) {
        x1 = uniform_pm1<R>(r);
        x2 = uniform_pm1<R>(r);
        sum = x1 * x1 + x2 * x2;
        if (sum < (R)1) break;
    }
    const R diff = x1 * x1 - x2 * x2;
    x = diff / sum;
    y = (R)2 * x1 * x2 / sum;
}

}  // namespace rptb
