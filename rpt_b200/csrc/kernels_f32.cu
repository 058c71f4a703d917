// kernels_f32.cu -- the product path: every kernel instantiated for Real = float.
#include <cstdlib>

#include "launch_impl.cuh"
#include <cub/device/device_radix_sort.cuh>

#include "wavefront.cuh"

namespace rptb {
RPTB_DEFINE_LAUNCHERS(f32, float)

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t wavefront_struct_size() { return sizeof(WfBuffers); }

static size_t sort_temp_bytes(size_t slots) {
    size_t bytes = 0;
    cub::DoubleBuffer<uint32_t> k(nullptr, nullptr), v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, (int)slots, 0, 31);
    return bytes;
}

// paths in flight per owned pixel slot: enough to keep ~2M paths alive, never more than there are chunks
uint32_t wavefront_groups(uint32_t npix, uint32_t nchunks) {
    if (npix == 0) return 1;
    uint32_t g = (2000000u + npix - 1u) / npix;
    if (g > nchunks) g = nchunks;
    return g < 1u ? 1u : g;
}

size_t wavefront_bytes(uint32_t npaths, uint32_t Ks, uint32_t maxd) {
    const size_t n = npaths, slots = n * (Ks + 1);
    return align256(n * sizeof(WfPath)) + align256(n * Ks * 3 * sizeof(float) + 16) + align256(n * maxd * 6 * sizeof(float) + 16) +
           align256(slots * sizeof(WfRay)) + align256(slots * sizeof(WfHit)) + 4 * align256(slots * sizeof(uint32_t)) +
           align256(sort_temp_bytes(slots)) + 256;
}

void wavefront_carve(void* mem, uint32_t npix, uint32_t G, uint32_t Ks, uint32_t maxd, const float* bounds_lo,
                     const float* bounds_inv_extent, WfBuffers* out) {
    const uint32_t npaths = npix * G;
    out->npix = npix;
    out->G = G;
    char* p = (char*)mem;
    const size_t n = npaths, slots = n * (Ks + 1);
    out->paths = (WfPath*)p; p += align256(n * sizeof(WfPath));
    out->contrib = (float*)p; p += align256(n * Ks * 3 * sizeof(float) + 16);
    out->levels = (float*)p; p += align256(n * maxd * 6 * sizeof(float) + 16);
    out->rays = (WfRay*)p; p += align256(slots * sizeof(WfRay));
    out->hits = (WfHit*)p; p += align256(slots * sizeof(WfHit));
    out->list = (uint32_t*)p; p += align256(slots * sizeof(uint32_t));
    out->list_alt = (uint32_t*)p; p += align256(slots * sizeof(uint32_t));
    out->keys = (uint32_t*)p; p += align256(slots * sizeof(uint32_t));
    out->keys_alt = (uint32_t*)p; p += align256(slots * sizeof(uint32_t));
    out->sort_tmp = p;
    out->sort_tmp_bytes = sort_temp_bytes(slots);
    p += align256(out->sort_tmp_bytes);
    out->count = (uint32_t*)p;
    for (int i = 0; i < 3; i++) {
        out->bounds_lo[i] = bounds_lo[i];
        out->bounds_inv[i] = bounds_inv_extent[i];
    }
    out->npaths = npaths;
    out->Ks = Ks;
    out->maxd = maxd;
}

cudaError_t run_wavefront_f32(const SceneView<float>& sv, const RenderArgs<float>& args, const WfBuffers* bufs,
                              bool stats, bool use_bvh, cudaStream_t stream, uint32_t* pinned, uint32_t* launches) {
    const WfBuffers b = *bufs;
    const uint32_t capacity = b.npaths * (b.Ks + 1);
    // Optional coherence sort of the ray list (Morton key of origin + direction octant).  Measured on
    // the dragon proxy: 142.9 vs 141.7 Msamples/s -- the compact list is already in pixel-tile order, so
    // the sort buys nothing there and is off unless RPTB_WF_SORT is set.
    const bool sort_rays = getenv("RPTB_WF_SORT") != nullptr;
    uint32_t nl = 0;
    cudaError_t e;
    const size_t nvals = (size_t)args.width * args.height * 3;
    if (args.shard_count > 1 && !args.compact) {
        clear_kernel<float><<<(unsigned)((nvals + 255) / 256), 256, 0, stream>>>(args.out, nvals);
        nl++;
    }
    if (b.npaths == 0) {
        if (launches) *launches = nl;
        return cudaGetLastError();
    }
    const unsigned pgrid = (b.npaths + WF_THREADS - 1) / WF_THREADS;
    wf_init_kernel<<<pgrid, WF_THREADS, 0, stream>>>(args, b);
    nl++;
    // persistent trace grid: as many CTAs as are resident at once
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // with counters the trace kernel always walks the reference-shaped kd-trees (their counts are the algorithmic work)
    const bool bvh = !stats && sv.nmeshes > 0 && use_bvh;
    if (stats) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wf_trace_kernel<true, false>, WF_THREADS, 0);
    else if (bvh) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wf_trace_kernel<false, true>, WF_THREADS, 0);
    else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wf_trace_kernel<false, false>, WF_THREADS, 0);
    if (e != cudaSuccess) return e;
    const unsigned tgrid = (unsigned)(sms * (per_sm > 0 ? per_sm : 1));
    // a sample with n <= max_bounces + 1 segments takes n + 1 steps (camera ray, one per vertex,
    // and the step that consumes the last vertex's shadow rays and emits the next camera ray); a path
    // runs ceil(nchunks / G) chunks of `chunk` samples
    const unsigned long long per_path = (unsigned long long)((args.nchunks + b.G - 1) / b.G) * args.chunk;
    const unsigned long long max_steps = (per_path < args.iterations ? per_path : args.iterations) * (args.max_bounces + 2ull) + 2ull;
    for (unsigned long long step = 0; step < max_steps; step++) {
        e = cudaMemsetAsync(b.count, 0, 2 * sizeof(uint32_t), stream);
        if (e != cudaSuccess) return e;
        if (stats) wf_shade_kernel<true><<<pgrid, WF_THREADS, 0, stream>>>(sv, args, b);
        else wf_shade_kernel<false><<<pgrid, WF_THREADS, 0, stream>>>(sv, args, b);
        const uint32_t* list = b.list;
        if (sort_rays) {
            wf_key_kernel<<<(capacity + 255) / 256, 256, 0, stream>>>(b, capacity);
            cub::DoubleBuffer<uint32_t> dk(b.keys, b.keys_alt), dv(b.list, b.list_alt);
            size_t tmp = b.sort_tmp_bytes;
            e = cub::DeviceRadixSort::SortPairs(b.sort_tmp, tmp, dk, dv, (int)capacity, 0, 31, stream);
            if (e != cudaSuccess) return e;
            list = dv.Current();
            nl += 2;
        }
        if (stats) wf_trace_kernel<true, false><<<tgrid, WF_THREADS, 0, stream>>>(sv, b, list, args.counters);
        else if (bvh) wf_trace_kernel<false, true><<<tgrid, WF_THREADS, 0, stream>>>(sv, b, list, nullptr);
        else wf_trace_kernel<false, false><<<tgrid, WF_THREADS, 0, stream>>>(sv, b, list, nullptr);
        nl += 2;
        if ((step & 3ull) == 3ull || step + 1 == max_steps) {
            e = cudaMemcpyAsync(pinned, b.count, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
            if (e != cudaSuccess) return e;
            e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess) return e;
            if (*pinned == 0) break;
        }
    }
    if (args.nchunks > 1) resolve_chunks_kernel<float><<<args.ntiles_mine, RENDER_THREADS, 0, stream>>>(args);
    else wf_finish_kernel<<<pgrid, WF_THREADS, 0, stream>>>(args, b);
    nl++;
    if (launches) *launches = nl;
    return cudaGetLastError();
}
}  // namespace rptb
