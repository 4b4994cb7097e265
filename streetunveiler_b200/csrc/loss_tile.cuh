// loss_tile.cuh -- per-tile bodies of the fused training-loss kernels.
//
// Behavioural specification: utils/loss_utils.py:17-64 (l1_loss, gaussian window, ssim) and train.py:113-136
// (sky composite, L1 + D-SSIM mix, normal-consistency and distortion means) of the reference.
//
// The bodies are __host__ __device__ functions of (tid, nthreads): on the GPU a 256-thread CTA runs them
// with __syncthreads() between phases; tests/emul/ compiles the SAME functions for the host (tid = 0,
// nthreads = 1, phases run back to back) so that indexing, halo and zero-padding logic is checked against the
// oracle on machines without a GPU.  Nothing per-thread lives across a phase boundary for that reason.
//
// One CTA = one 32x32 output tile, all three channels in turn.  The 11x11 Gaussian window is applied in
// separable form: a horizontal pass over the 42x42 halo tile into shared memory, then a vertical pass in
// which every thread produces 4 vertically adjacent outputs from 14 loads per map (register blocking).
#pragma once
#include <math.h>

#include "common.cuh"

namespace surfel {

constexpr int LT = 32;               // output tile edge
constexpr int LR = 5;                // window radius (window_size // 2, loss_utils.py:45)
constexpr int LWIN = 2 * LR + 1;     // 11
constexpr int LIN = LT + 2 * LR;     // 42: tile + halo
constexpr int LPITCH = LIN + 1;
constexpr int LOSS_THREADS = 256;
constexpr int LROWS = 4;             // outputs per thread in the vertical pass
constexpr float SSIM_C1 = 0.01f * 0.01f;   // loss_utils.py:56-57
constexpr float SSIM_C2 = 0.03f * 0.03f;

struct LossWindow {
    float w[LWIN];
};

// loss_utils.py:23-25: python-double exp -> float32, float32 sum, float32 division
inline LossWindow make_loss_window()
{
    LossWindow w;
    float sum = 0.f;
    for (int x = 0; x < LWIN; x++) {
        w.w[x] = (float)exp(-(double)((x - LR) * (x - LR)) / (2.0 * 1.5 * 1.5));
        sum += w.w[x];
    }
    for (int x = 0; x < LWIN; x++) w.w[x] = w.w[x] / sum;
    return w;
}

#ifdef __CUDA_ARCH__
#define LOSS_SYNC() __syncthreads()
#else
#define LOSS_SYNC() ((void)0)
#endif

struct LossImages {          // all [C,H,W] fp32, device pointers
    const float *render;     // [3,H,W]
    const float *alpha;      // [1,H,W] or nullptr (with sky == nullptr: img1 = render)
    const float *sky;        // [3,H,W] or nullptr
    const float *gt;         // [3,H,W]
    int W, H;
};

// train.py:113 with the three separately rounded fp32 operations of the reference's PyTorch ops
// (no FMA contraction: where the composite equals gt bit for bit, |.| sits on its kink and sign(0) = 0).
__host__ __device__ inline float composite_px(const LossImages &im, const int ch, const size_t p, const size_t HW)
{
    const float r = im.render[ch * HW + p];
    if (!im.sky) return r;
#ifdef __CUDA_ARCH__
    return __fadd_rn(r, __fmul_rn(im.sky[ch * HW + p], __fsub_rn(1.0f, im.alpha[p])));
#else
    volatile float one_m = 1.0f - im.alpha[p];
    volatile float prod = im.sky[ch * HW + p] * one_m;
    return r + prod;
#endif
}

struct LossFwdSmem {
    float x[LIN][LPITCH];    // img1 (composite), zero outside the image = conv2d's zero padding
    float y[LIN][LPITCH];    // img2 (gt)
    float h[5][LIN][LT];     // horizontally filtered x, y, x^2, y^2, xy
    float red[2][LOSS_THREADS / 32];
};

// Forward tile: accumulates this thread's share of sum|x - y| and sum(ssim_map) into (acc_l1, acc_ssim) and,
// if deriv != nullptr, stores per channel the three maps the backward convolves:
//   deriv[3 ch + 0] = d ssim_map / d mu1 (total: direct + through sigma1_sq and sigma12)
//   deriv[3 ch + 1] = d ssim_map / d E[x^2],  deriv[3 ch + 2] = d ssim_map / d E[xy]
__host__ __device__ inline void loss_fwd_tile(LossFwdSmem &s, const LossImages &im, const LossWindow &win, const int tx0,
                                              const int ty0, const int tid, const int nthreads, float *deriv,
                                              float &acc_l1, float &acc_ssim)
{
    const int W = im.W, H = im.H;
    const size_t HW = (size_t)W * H;
    for (int ch = 0; ch < 3; ch++) {
        for (int i = tid; i < LIN * LIN; i += nthreads) {
            const int r = i / LIN, c = i - r * LIN;
            const int gy = ty0 - LR + r, gx = tx0 - LR + c;
            float xv = 0.f, yv = 0.f;
            if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
                const size_t p = (size_t)gy * W + gx;
                xv = composite_px(im, ch, p, HW);
                yv = im.gt[ch * HW + p];
            }
            s.x[r][c] = xv;
            s.y[r][c] = yv;
        }
        LOSS_SYNC();
        for (int i = tid; i < LIN * LT; i += nthreads) {
            const int r = i / LT, c = i - r * LT;
            float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f;
#pragma unroll
            for (int k = 0; k < LWIN; k++) {
                const float xv = s.x[r][c + k], yv = s.y[r][c + k], wk = win.w[k];
                m0 += wk * xv;
                m1 += wk * yv;
                m2 += wk * (xv * xv);
                m3 += wk * (yv * yv);
                m4 += wk * (xv * yv);
            }
            s.h[0][r][c] = m0;
            s.h[1][r][c] = m1;
            s.h[2][r][c] = m2;
            s.h[3][r][c] = m3;
            s.h[4][r][c] = m4;
        }
        LOSS_SYNC();
        for (int t = tid; t < LT * (LT / LROWS); t += nthreads) {
            const int c = t % LT, r0 = (t / LT) * LROWS;
            float o[5][LROWS];
#pragma unroll
            for (int m = 0; m < 5; m++) {
                float v[LROWS + LWIN - 1];
#pragma unroll
                for (int j = 0; j < LROWS + LWIN - 1; j++) v[j] = s.h[m][r0 + j][c];
#pragma unroll
                for (int q = 0; q < LROWS; q++) {
                    float a = 0.f;
#pragma unroll
                    for (int k = 0; k < LWIN; k++) a += win.w[k] * v[q + k];
                    o[m][q] = a;
                }
            }
#pragma unroll
            for (int q = 0; q < LROWS; q++) {
                const int gy = ty0 + r0 + q, gx = tx0 + c;
                if (gx >= W || gy >= H) continue;
                const float mu1 = o[0][q], mu2 = o[1][q];
                const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;   // loss_utils.py:48-50
                const float s1 = o[2][q] - mu1_sq, s2 = o[3][q] - mu2_sq, s12 = o[4][q] - mu12;   // :52-54
                const float a = 2.f * mu12 + SSIM_C1, b = 2.f * s12 + SSIM_C2;
                const float cc = mu1_sq + mu2_sq + SSIM_C1, d = s1 + s2 + SSIM_C2;
                const float inv_cd = 1.f / (cc * d);
                const float f = (a * b) * inv_cd;                                       // :59
                acc_ssim += f;
                acc_l1 += fabsf(s.x[r0 + q + LR][c + LR] - s.y[r0 + q + LR][c + LR]);   // :18
                if (deriv) {
                    const float df_ds1 = -f / d;
                    const float df_ds12 = 2.f * a * inv_cd;
                    const float df_dmu1 = 2.f * mu2 * b * inv_cd - 2.f * mu1 * f / cc;
                    const size_t p = (size_t)gy * W + gx;
                    deriv[(3 * ch + 0) * HW + p] = df_dmu1 - 2.f * mu1 * df_ds1 - mu2 * df_ds12;
                    deriv[(3 * ch + 1) * HW + p] = df_ds1;
                    deriv[(3 * ch + 2) * HW + p] = df_ds12;
                }
            }
        }
        LOSS_SYNC();
    }
}

struct LossBwdSmem {
    float a[3][LIN][LPITCH];   // the three derivative maps with halo, zero outside the image
    float h[3][LIN][LT];
    float dalpha[LT][LT];
};

// Backward tile: d_render = ws (K*A + 2 x K*B + y K*C) + wl sign(x - y)  (K* = the same zero-padded window: the
// adjoint of a symmetric convolution), then the composite's chain rule: d_sky = d_render (1 - alpha),
// d_alpha = -sum_ch d_render sky.  wl = dL/dLl1 / (3HW), ws = dL/dLssim / (3HW).
__host__ __device__ inline void loss_bwd_tile(LossBwdSmem &s, const LossImages &im, const LossWindow &win, const int tx0,
                                              const int ty0, const int tid, const int nthreads, const float *deriv,
                                              const float wl, const float ws, float *d_render, float *d_alpha,
                                              float *d_sky)
{
    const int W = im.W, H = im.H;
    const size_t HW = (size_t)W * H;
    for (int ch = 0; ch < 3; ch++) {
        for (int i = tid; i < LIN * LIN; i += nthreads) {
            const int r = i / LIN, c = i - r * LIN;
            const int gy = ty0 - LR + r, gx = tx0 - LR + c;
            float v0 = 0.f, v1 = 0.f, v2 = 0.f;
            if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
                const size_t p = (size_t)gy * W + gx;
                v0 = deriv[(3 * ch + 0) * HW + p];
                v1 = deriv[(3 * ch + 1) * HW + p];
                v2 = deriv[(3 * ch + 2) * HW + p];
            }
            s.a[0][r][c] = v0;
            s.a[1][r][c] = v1;
            s.a[2][r][c] = v2;
        }
        LOSS_SYNC();
        for (int i = tid; i < LIN * LT; i += nthreads) {
            const int r = i / LT, c = i - r * LT;
            float m0 = 0.f, m1 = 0.f, m2 = 0.f;
#pragma unroll
            for (int k = 0; k < LWIN; k++) {
                const float wk = win.w[k];
                m0 += wk * s.a[0][r][c + k];
                m1 += wk * s.a[1][r][c + k];
                m2 += wk * s.a[2][r][c + k];
            }
            s.h[0][r][c] = m0;
            s.h[1][r][c] = m1;
            s.h[2][r][c] = m2;
        }
        LOSS_SYNC();
        for (int t = tid; t < LT * (LT / LROWS); t += nthreads) {
            const int c = t % LT, r0 = (t / LT) * LROWS;
            float o[3][LROWS];
#pragma unroll
            for (int m = 0; m < 3; m++) {
                float v[LROWS + LWIN - 1];
#pragma unroll
                for (int j = 0; j < LROWS + LWIN - 1; j++) v[j] = s.h[m][r0 + j][c];
#pragma unroll
                for (int q = 0; q < LROWS; q++) {
                    float a = 0.f;
#pragma unroll
                    for (int k = 0; k < LWIN; k++) a += win.w[k] * v[q + k];
                    o[m][q] = a;
                }
            }
#pragma unroll
            for (int q = 0; q < LROWS; q++) {
                const int gy = ty0 + r0 + q, gx = tx0 + c;
                if (gx >= W || gy >= H) continue;
                const size_t p = (size_t)gy * W + gx;
                const float x = composite_px(im, ch, p, HW), y = im.gt[ch * HW + p];
                const float diff = x - y;
                const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);   // torch: d|u|/du = sgn(u), sgn(0) = 0
                const float dx = ws * (o[0][q] + 2.f * x * o[1][q] + y * o[2][q]) + wl * sgn;
                d_render[ch * HW + p] = dx;
                if (im.sky) {
                    const float skyv = im.sky[ch * HW + p];
                    if (d_sky) d_sky[ch * HW + p] = dx * (1.0f - im.alpha[p]);
                    const float prev = ch == 0 ? 0.f : s.dalpha[r0 + q][c];   // same thread owns the pixel in every channel
                    s.dalpha[r0 + q][c] = prev - dx * skyv;
                    if (ch == 2 && d_alpha) d_alpha[p] = s.dalpha[r0 + q][c];
                }
            }
        }
        LOSS_SYNC();
    }
}

// Regulariser sums over a strided range of pixels (train.py:125-126,134): sum(1 - <rend_normal, surf_normal>), sum(rend_dist)
__host__ __device__ inline void regulariser_sums(const float *rn, const float *sn, const float *dist, const size_t HW,
                                                 const size_t first, const size_t stride, float &acc_n, float &acc_d)
{
    for (size_t p = first; p < HW; p += stride) {
        const float dot = rn[p] * sn[p] + rn[HW + p] * sn[HW + p] + rn[2 * HW + p] * sn[2 * HW + p];
        acc_n += 1.0f - dot;
        acc_d += dist[p];
    }
}

__host__ __device__ inline void regulariser_grads(const float *rn, const float *sn, const size_t HW, const size_t first,
                                                  const size_t stride, const float wn, const float wd, float *d_rn,
                                                  float *d_sn, float *d_dist)
{
    for (size_t p = first; p < HW; p += stride) {
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            d_rn[ch * HW + p] = -wn * sn[ch * HW + p];
            d_sn[ch * HW + p] = -wn * rn[ch * HW + p];
        }
        d_dist[p] = wd;
    }
}

// Deterministic second stage of the means: out[k] = scale * sum_b partials[k * nblocks + b] (double accumulation,
// fixed order per thread) -- `tid`/`nthreads` as above; on the device the caller reduces `acc` over the CTA.
__host__ __device__ inline double partial_column_sum(const float *partials, const int nblocks, const int k, const int tid,
                                                     const int nthreads)
{
    double acc = 0.0;
    for (int b = tid; b < nblocks; b += nthreads) acc += (double)partials[(size_t)k * nblocks + b];
    return acc;
}

}  // namespace surfel
