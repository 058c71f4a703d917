// film.cu -- film resolve on the device ("next" row N1 of SURVEY section 8f).
//
// Replaces Buffer::image -> get_filtered_color -> color_bytes
// (ekzhang/rpt src/buffer.rs:43-56,75-93, src/color.rs:17-23): box filter of radius r
// over the per-pixel sample sums (every pixel holds `nbatches` equally weighted
// entries), then clamp, gamma 1/2.2 and a truncating cast to u8.  Computed in f64 and
// summed in the reference's order (x outer, y inner), so the bytes match the CPU
// restatement exactly.
#include <cuda_runtime.h>
#include <stdint.h>

namespace rptb {

__global__ void film_resolve_kernel(const double* __restrict__ sums, uint32_t nbatches, uint32_t width,
                                    uint32_t height, uint32_t radius, uint8_t* __restrict__ out) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= width || y >= height) return;
    double c0 = 0.0, c1 = 0.0, c2 = 0.0;
    unsigned long long count = 0;
    const uint32_t i0 = x >= radius ? x - radius : 0u;  // saturating_sub
    const uint32_t j0 = y >= radius ? y - radius : 0u;
    const uint64_t i1 = (uint64_t)x + radius, j1 = (uint64_t)y + radius;
    for (uint64_t i = i0; i <= i1 && i < width; i++)
        for (uint64_t j = j0; j <= j1 && j < height; j++) {
            const double* p = sums + 3 * (j * width + i);
            c0 += p[0];
            c1 += p[1];
            c2 += p[2];
            count += nbatches;
        }
    const double n = (double)count;
    const double c[3] = {c0 / n, c1 / n, c2 / n};
    uint8_t* o = out + 3 * ((size_t)y * width + x);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double v = fmin(fmax(c[k], 0.0), 1.0);
        o[k] = (uint8_t)(pow(v, 1.0 / 2.2) * 255.0);  // `as u8` truncates
    }
}

__global__ void convert_f64_f32_kernel(const double* __restrict__ in, float* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

// Buffer::variance (src/buffer.rs:59-73): per pixel, the sample variance (n - 1) of its `nbatches`
// entries summed over the three channels; block-reduced and added to *out_sum (the caller divides by
// the pixel count).  The block partials are added with atomics, so the last bits depend on the order.
__global__ void film_variance_kernel(const double* __restrict__ batches, uint32_t nbatches, uint64_t npixels,
                                     double* __restrict__ out_sum) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0.0;
    if (p < npixels) {
        double mean[3] = {0.0, 0.0, 0.0};
        for (uint32_t b = 0; b < nbatches; b++)
            for (int k = 0; k < 3; k++) mean[k] += batches[((size_t)b * npixels + p) * 3 + k];
        for (int k = 0; k < 3; k++) mean[k] /= (double)nbatches;
        double ss = 0.0;
        for (uint32_t b = 0; b < nbatches; b++)
            for (int k = 0; k < 3; k++) {
                const double d = batches[((size_t)b * npixels + p) * 3 + k] - mean[k];
                ss += d * d;
            }
        v = ss / ((double)nbatches - 1.0);
    }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __shared__ double warp_sum[8];
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (unsigned w = 0; w < blockDim.x / 32; w++) s += warp_sum[w];
        atomicAdd(out_sum, s);
    }
}

cudaError_t launch_film_variance(const double* batches, uint32_t nbatches, uint64_t npixels, double* out_sum,
                                 cudaStream_t stream) {
    film_variance_kernel<<<(unsigned)((npixels + 255) / 256), 256, 0, stream>>>(batches, nbatches, npixels, out_sum);
    return cudaGetLastError();
}

cudaError_t launch_film_resolve(const double* sums, uint32_t nbatches, uint32_t width, uint32_t height,
                                uint32_t radius, uint8_t* out, cudaStream_t stream) {
    const dim3 block(32, 8), grid((width + 31) / 32, (height + 7) / 8);
    film_resolve_kernel<<<grid, block, 0, stream>>>(sums, nbatches, width, height, radius, out);
    return cudaGetLastError();
}

cudaError_t launch_convert_f64_to_f32(const double* in, float* out, size_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    convert_f64_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(in, out, n);
    return cudaGetLastError();
}

}  // namespace rptb
