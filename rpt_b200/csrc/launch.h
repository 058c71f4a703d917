// launch.h -- host-callable launchers, one set per precision (each set lives in its
// own translation unit so the f64 parity gate can be compiled with -fmad=false).
#pragma once
#include "scene_dev.cuh"

namespace rptb {

#define RPTB_DECLARE_LAUNCHERS(SUFFIX, R)                                                                          \
    cudaError_t launch_render_##SUFFIX(const SceneView<R>& sv, const RenderArgs<R>& args, int stats,               \
                                       int features, cudaStream_t stream, uint32_t* launches);                     \
    cudaError_t launch_closest_hit_##SUFFIX(const SceneView<R>& sv, const double* rays, uint64_t n, double tmin,   \
                                            double* out_t, int32_t* out_obj, double* out_n,                        \
                                            DeviceCounters* counters, int stats, int features,                     \
                                            cudaStream_t stream);                                                  \
    cudaError_t launch_bsdf_##SUFFIX(const MaterialRec<R>& m, const double* dirs, uint64_t n, double* out,         \
                                     cudaStream_t stream);                                                         \
    cudaError_t launch_sample_f_##SUFFIX(const MaterialRec<R>& m, const double* dirs, uint64_t n, uint64_t seed,   \
                                         double* out_wi, double* out_pdf, cudaStream_t stream);                    \
    cudaError_t launch_illuminate_##SUFFIX(const SceneView<R>& sv, uint32_t light, const double* pos, uint64_t n,  \
                                           uint64_t seed, double* out_i, double* out_wi, double* out_dist,         \
                                           cudaStream_t stream);

RPTB_DECLARE_LAUNCHERS(f32, float)
RPTB_DECLARE_LAUNCHERS(f64, double)

// ---- wavefront engine (f32 only; wavefront.cuh) -------------------------------------------
struct WfBuffers;
// bytes of device memory the engine needs for (npaths, Ks sampled lights, maxd levels)
size_t wavefront_bytes(uint32_t npaths, uint32_t Ks, uint32_t maxd);
// carve `mem` (wavefront_bytes big, 256-byte aligned) into the engine's arrays
void wavefront_carve(void* mem, uint32_t npix, uint32_t G, uint32_t Ks, uint32_t maxd, WfBuffers* out);
uint32_t wavefront_groups(uint32_t npix, uint32_t nchunks);
size_t wavefront_struct_size();
// run Renderer::sample with the wavefront schedule: everything is enqueued on `stream` (the step loop is a CUDA graph
// WHILE node re-armed on the device); the call does not wait
cudaError_t run_wavefront_f32(const SceneView<float>& sv, const RenderArgs<float>& args, const WfBuffers* bufs,
                              bool stats, bool use_bvh, cudaStream_t stream, uint32_t* launches);

// ---- the vertex-at-once f32 megakernel (integrator_vx.cuh; its own translation unit, kernels_vx.cu) -----------------
// stats as in launch_render_f32.  args.ks must be <= VX_MAX_SHADOW (vx_supported).
cudaError_t launch_render_vx_f32(const SceneView<float>& sv, const RenderArgs<float>& args, int stats, int features,
                                 cudaStream_t stream, uint32_t* launches);
bool vx_supported(uint32_t sampled_lights);

// max_bounces the render kernels are instantiated for
constexpr uint32_t MAX_BOUNCES_SUPPORTED = 64;

}  // namespace rptb
