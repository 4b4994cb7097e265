// loss.cu -- fused training-loss block: kernels and C ABI.
//
// Reference: utils/loss_utils.py:17-64 and train.py:113-136 -- ~60 PyTorch kernels per iteration (sky composite,
// |.|.mean(), five 11x11 grouped convolutions, the SSIM map and its autograd, two regulariser means).  Here:
//   forward   loss_photometric_fwd_kernel  (composite + L1 + SSIM map + the 3 maps the backward convolves)
//             loss_regulariser_fwd_kernel  (normal-consistency and distortion sums)
//             loss_finalize_kernel         (deterministic double-precision second stage of the means)
//   backward  loss_photometric_bwd_kernel  (d_render, d_alpha, d_sky in one pass)
//             loss_regulariser_bwd_kernel
// The tile bodies live in loss_tile.cuh (shared with the host emulation used by the CPU tests).
#include <cmath>

#include "../../include/surfel_rasterizer.h"
#include "kernels.h"
#include "loss_tile.cuh"

namespace surfel {

namespace {

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// CTA-wide sums of two values; result valid in thread 0.  red = [2][warps] floats of shared memory.
__device__ __forceinline__ void block_sum2(float &a, float &b, float (*red)[LOSS_THREADS / 32])
{
    a = warp_sum(a);
    b = warp_sum(b);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        red[0][warp] = a;
        red[1][warp] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = 0.f;
        b = 0.f;
        for (int w = 0; w < LOSS_THREADS / 32; w++) {
            a += red[0][w];
            b += red[1][w];
        }
    }
}

}  // namespace

// How a backward kernel forms its two scalar weights from the upstream gradient vector on the device (no host
// read-back): weight_k = sum_{j<n} c[k][j] * g[j].  The standalone entry points use the identity on 2 values; the fused
// training loss mixes the 5 gradients of (loss, l1, ssim, Lnormal, Ldist) with the lambdas of train.py:117-136.
struct UpstreamMix {
    const float *g;
    int n;
    float c[2][5];
    __device__ __forceinline__ float weight(const int k) const
    {
        float w = 0.f;
        for (int j = 0; j < n; j++) w += c[k][j] * g[j];
        return w;
    }
};

__global__ void __launch_bounds__(LOSS_THREADS)
loss_photometric_fwd_kernel(const LossImages im, const LossWindow win, float *__restrict__ deriv,
                            float *__restrict__ partials)
{
    __shared__ LossFwdSmem s;
    float acc_l1 = 0.f, acc_ssim = 0.f;
    loss_fwd_tile(s, im, win, blockIdx.x * LT, blockIdx.y * LT, threadIdx.x, LOSS_THREADS, deriv, acc_l1, acc_ssim);
    block_sum2(acc_l1, acc_ssim, s.red);
    if (threadIdx.x == 0) {
        const int nblocks = gridDim.x * gridDim.y, b = blockIdx.y * gridDim.x + blockIdx.x;
        partials[b] = acc_l1;
        partials[nblocks + b] = acc_ssim;
    }
}

__global__ void __launch_bounds__(LOSS_THREADS)
loss_photometric_bwd_kernel(const LossImages im, const LossWindow win, const float *__restrict__ deriv,
                            const UpstreamMix upstream, float *__restrict__ d_render,
                            float *__restrict__ d_alpha, float *__restrict__ d_sky)
{
    __shared__ LossBwdSmem s;
    const float inv_n = 1.0f / (3.0f * (float)im.W * (float)im.H);
    const float wl = upstream.weight(0) * inv_n, ws = upstream.weight(1) * inv_n;
    loss_bwd_tile(s, im, win, blockIdx.x * LT, blockIdx.y * LT, threadIdx.x, LOSS_THREADS, deriv, wl, ws, d_render,
                  d_alpha, d_sky);
}

__global__ void __launch_bounds__(LOSS_THREADS)
loss_regulariser_fwd_kernel(const float *__restrict__ rn, const float *__restrict__ sn, const float *__restrict__ dist,
                            const size_t HW, float *__restrict__ partials)
{
    __shared__ float red[2][LOSS_THREADS / 32];
    float acc_n = 0.f, acc_d = 0.f;
    regulariser_sums(rn, sn, dist, HW, (size_t)blockIdx.x * LOSS_THREADS + threadIdx.x, (size_t)gridDim.x * LOSS_THREADS,
                     acc_n, acc_d);
    block_sum2(acc_n, acc_d, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = acc_n;
        partials[gridDim.x + blockIdx.x] = acc_d;
    }
}

__global__ void __launch_bounds__(LOSS_THREADS)
loss_regulariser_bwd_kernel(const float *__restrict__ rn, const float *__restrict__ sn, const size_t HW,
                            const UpstreamMix upstream, float *__restrict__ d_rn, float *__restrict__ d_sn,
                            float *__restrict__ d_dist)
{
    const float wn = upstream.weight(0) / (float)HW, wd = upstream.weight(1) / (float)HW;
    regulariser_grads(rn, sn, HW, (size_t)blockIdx.x * LOSS_THREADS + threadIdx.x, (size_t)gridDim.x * LOSS_THREADS, wn, wd,
                      d_rn, d_sn, d_dist);
}

// one CTA; out[k] = scale * sum of column k of partials
__global__ void __launch_bounds__(LOSS_THREADS)
loss_finalize_kernel(const float *__restrict__ partials, const int nblocks, const double scale, float *__restrict__ out)
{
    __shared__ double red[2][LOSS_THREADS];
    for (int k = 0; k < 2; k++) red[k][threadIdx.x] = partial_column_sum(partials, nblocks, k, threadIdx.x, LOSS_THREADS);
    __syncthreads();
    for (int o = LOSS_THREADS / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            red[0][threadIdx.x] += red[0][threadIdx.x + o];
            red[1][threadIdx.x] += red[1][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x < 2) out[threadIdx.x] = (float)(red[threadIdx.x][0] * scale);
}

// one CTA: both pairs of means and the combination of train.py:117-136 -> out5 = (loss, l1, ssim, Lnormal, Ldist)
__global__ void __launch_bounds__(LOSS_THREADS)
loss_training_finalize_kernel(const float *__restrict__ photo_partials, const int photo_blocks, const double photo_scale,
                              const float *__restrict__ reg_partials, const int reg_blocks, const double reg_scale,
                              const float lambda_dssim, const float lambda_normal, const float lambda_dist,
                              float *__restrict__ out5)
{
    __shared__ double red[4][LOSS_THREADS];
    for (int k = 0; k < 2; k++) {
        red[k][threadIdx.x] = partial_column_sum(photo_partials, photo_blocks, k, threadIdx.x, LOSS_THREADS);
        red[2 + k][threadIdx.x] = partial_column_sum(reg_partials, reg_blocks, k, threadIdx.x, LOSS_THREADS);
    }
    __syncthreads();
    for (int o = LOSS_THREADS / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o)
            for (int k = 0; k < 4; k++) red[k][threadIdx.x] += red[k][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float l1 = (float)(red[0][0] * photo_scale), ssim = (float)(red[1][0] * photo_scale);
        const float ln = lambda_normal * (float)(red[2][0] * reg_scale), ld = lambda_dist * (float)(red[3][0] * reg_scale);
        out5[0] = (1.0f - lambda_dssim) * l1 + lambda_dssim * (1.0f - ssim) + ln + ld;
        out5[1] = l1;
        out5[2] = ssim;
        out5[3] = ln;
        out5[4] = ld;
    }
}

constexpr int REG_BLOCKS = 148 * 4;   // regulariser kernels: grid-stride, a multiple of the SM count

}  // namespace surfel

using namespace surfel;

static UpstreamMix identity_mix(const float *g)
{
    UpstreamMix m{g, 2, {{1.f, 0.f, 0.f, 0.f, 0.f}, {0.f, 1.f, 0.f, 0.f, 0.f}}};
    return m;
}

extern "C" {

size_t surfel_loss_scratch_bytes(int width, int height)
{
    if (width <= 0 || height <= 0) {
        surfel_internal_fail("surfel_loss_scratch_bytes", "bad image size");
        return 0;
    }
    const size_t tiles = (size_t)((width + LT - 1) / LT) * ((height + LT - 1) / LT);
    return 2 * (tiles + (size_t)REG_BLOCKS) * sizeof(float) + 256;   // both partial arrays, each 128-byte aligned
}

int surfel_loss_photometric_forward(int width, int height, const float *render, const float *rend_alpha,
                                    const float *sky, const float *gt, float *deriv, char *scratch, float *out_means,
                                    void *stream)
{
    const char *where = "surfel_loss_photometric_forward";
    if (width <= 0 || height <= 0 || !render || !gt || !scratch || !out_means) return surfel_internal_fail(where, "bad arguments");
    if ((sky != nullptr) != (rend_alpha != nullptr)) return surfel_internal_fail(where, "sky and rend_alpha go together");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const LossImages im{render, rend_alpha, sky, gt, width, height};
    const dim3 grid((width + LT - 1) / LT, (height + LT - 1) / LT);
    char *p = scratch;
    float *partials = carve<float>(p, 2 * (size_t)grid.x * grid.y);
    loss_photometric_fwd_kernel<<<grid, LOSS_THREADS, 0, st>>>(im, make_loss_window(), deriv, partials);
    loss_finalize_kernel<<<1, LOSS_THREADS, 0, st>>>(partials, (int)(grid.x * grid.y), 1.0 / (3.0 * (double)width * (double)height),
                                                     out_means);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : surfel_internal_fail(where, cudaGetErrorString(e));
}

int surfel_loss_photometric_backward(int width, int height, const float *render, const float *rend_alpha,
                                     const float *sky, const float *gt, const float *deriv, const float *upstream,
                                     float *d_render, float *d_rend_alpha, float *d_sky, void *stream)
{
    const char *where = "surfel_loss_photometric_backward";
    if (width <= 0 || height <= 0 || !render || !gt || !deriv || !upstream || !d_render)
        return surfel_internal_fail(where, "bad arguments");
    if ((sky != nullptr) != (rend_alpha != nullptr)) return surfel_internal_fail(where, "sky and rend_alpha go together");
    const LossImages im{render, rend_alpha, sky, gt, width, height};
    const dim3 grid((width + LT - 1) / LT, (height + LT - 1) / LT);
    loss_photometric_bwd_kernel<<<grid, LOSS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        im, make_loss_window(), deriv, identity_mix(upstream), d_render, d_rend_alpha, d_sky);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : surfel_internal_fail(where, cudaGetErrorString(e));
}

int surfel_loss_regulariser_forward(int width, int height, const float *rend_normal, const float *surf_normal,
                                    const float *rend_dist, char *scratch, float *out_means, void *stream)
{
    const char *where = "surfel_loss_regulariser_forward";
    if (width <= 0 || height <= 0 || !rend_normal || !surf_normal || !rend_dist || !scratch || !out_means)
        return surfel_internal_fail(where, "bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t HW = (size_t)width * height;
    char *p = scratch;
    float *partials = carve<float>(p, 2 * (size_t)REG_BLOCKS);
    loss_regulariser_fwd_kernel<<<REG_BLOCKS, LOSS_THREADS, 0, st>>>(rend_normal, surf_normal, rend_dist, HW, partials);
    loss_finalize_kernel<<<1, LOSS_THREADS, 0, st>>>(partials, REG_BLOCKS, 1.0 / (double)HW, out_means);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : surfel_internal_fail(where, cudaGetErrorString(e));
}

int surfel_loss_regulariser_backward(int width, int height, const float *rend_normal, const float *surf_normal,
                                     const float *upstream, float *d_rend_normal, float *d_surf_normal,
                                     float *d_rend_dist, void *stream)
{
    const char *where = "surfel_loss_regulariser_backward";
    if (width <= 0 || height <= 0 || !rend_normal || !surf_normal || !upstream || !d_rend_normal || !d_surf_normal ||
        !d_rend_dist)
        return surfel_internal_fail(where, "bad arguments");
    loss_regulariser_bwd_kernel<<<REG_BLOCKS, LOSS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        rend_normal, surf_normal, (size_t)width * height, identity_mix(upstream), d_rend_normal, d_surf_normal, d_rend_dist);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : surfel_internal_fail(where, cudaGetErrorString(e));
}

int surfel_loss_training_forward(int width, int height, const float *render, const float *rend_alpha, const float *sky,
                                 const float *gt, const float *rend_normal, const float *surf_normal,
                                 const float *rend_dist, float lambda_dssim, float lambda_normal, float lambda_dist,
                                 float *deriv, char *scratch, float *out5, void *stream)
{
    const char *where = "surfel_loss_training_forward";
    if (width <= 0 || height <= 0 || !render || !gt || !rend_normal || !surf_normal || !rend_dist || !scratch || !out5)
        return surfel_internal_fail(where, "bad arguments");
    if ((sky != nullptr) != (rend_alpha != nullptr)) return surfel_internal_fail(where, "sky and rend_alpha go together");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const LossImages im{render, rend_alpha, sky, gt, width, height};
    const dim3 grid((width + LT - 1) / LT, (height + LT - 1) / LT);
    const int tiles = (int)(grid.x * grid.y);
    const size_t HW = (size_t)width * height;
    char *p = scratch;
    float *photo_partials = carve<float>(p, 2 * (size_t)tiles);
    float *reg_partials = carve<float>(p, 2 * (size_t)REG_BLOCKS);
    loss_photometric_fwd_kernel<<<grid, LOSS_THREADS, 0, st>>>(im, make_loss_window(), deriv, photo_partials);
    loss_regulariser_fwd_kernel<<<REG_BLOCKS, LOSS_THREADS, 0, st>>>(rend_normal, surf_normal, rend_dist, HW, reg_partials);
    loss_training_finalize_kernel<<<1, LOSS_THREADS, 0, st>>>(photo_partials, tiles, 1.0 / (3.0 * (double)HW), reg_partials,
                                                              REG_BLOCKS, 1.0 / (double)HW, lambda_dssim, lambda_normal,
                                                              lambda_dist, out5);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : surfel_internal_fail(where, cudaGetErrorString(e));
}

int surfel_loss_training_backward(int width, int height, const float *render, const float *rend_alpha, const float *sky,
                                  const float *gt, const float *rend_normal, const float *surf_normal, const float *deriv,
                                  const float *g_out5, float lambda_dssim, float lambda_normal, float lambda_dist,
                                  float *d_render, float *d_rend_alpha, float *d_sky, float *d_rend_normal,
                                  float *d_surf_normal, float *d_rend_dist, void *stream)
{
    const char *where = "surfel_loss_training_backward";
    if (width <= 0 || height <= 0 || !render || !gt || !rend_normal || !surf_normal || !deriv || !g_out5 || !d_render ||
        !d_rend_normal || !d_surf_normal || !d_rend_dist)
        return surfel_internal_fail(where, "bad arguments");
    if ((sky != nullptr) != (rend_alpha != nullptr)) return surfel_internal_fail(where, "sky and rend_alpha go together");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const LossImages im{render, rend_alpha, sky, gt, width, height};
    const dim3 grid((width + LT - 1) / LT, (height + LT - 1) / LT);
    // d loss / d l1 = 1 - lambda_dssim, d loss / d ssim = -lambda_dssim (train.py:117); Lnormal = lambda_normal * mean, ...
    const UpstreamMix photo{g_out5, 5, {{1.0f - lambda_dssim, 1.f, 0.f, 0.f, 0.f}, {-lambda_dssim, 0.f, 1.f, 0.f, 0.f}}};
    const UpstreamMix reg{g_out5, 5, {{lambda_normal, 0.f, 0.f, lambda_normal, 0.f}, {lambda_dist, 0.f, 0.f, 0.f, lambda_dist}}};
    loss_photometric_bwd_kernel<<<grid, LOSS_THREADS, 0, st>>>(im, make_loss_window(), deriv, photo, d_render, d_rend_alpha, d_sky);
    loss_regulariser_bwd_kernel<<<REG_BLOCKS, LOSS_THREADS, 0, st>>>(rend_normal, surf_normal, (size_t)width * height, reg,
                                                                      d_rend_normal, d_surf_normal, d_rend_dist);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : surfel_internal_fail(where, cudaGetErrorString(e));
}

}  // extern "C"
