// activate_tile.cuh -- bodies of the fused parameter-activation kernels (the caller-side row next to K1 / K8).
//
// Behavioural specification: scene/gaussian_model.py:31-39,101-127 of the reference -- the properties every render()
// call evaluates before it reaches the rasterizer:
//     get_scaling  = torch.exp(_scaling)                       [P,2]
//     get_rotation = torch.nn.functional.normalize(_rotation)  [P,4]   (x / max(||x||_2, 1e-12))
//     get_opacity  = torch.sigmoid(_opacity)                   [P,1]
//     get_features = torch.cat((_features_dc, _features_rest), dim=1)   [P,1+R,3]
// and what autograd runs for them backwards (ExpBackward, the norm / clamp_min / div chain, SigmoidBackward, the
// split of the concatenation into two freshly allocated gradient tensors): ~12 PyTorch kernels forward and ~15
// backward, two of them full copies of the 192 B/Gaussian SH block.  Here: one launch each way.
//
// As in loss_tile.cuh / adam_tile.cuh the bodies are __host__ __device__ so tests/emul/ can run them on the CPU.
#pragma once
#include <math.h>

#include "common.cuh"

namespace surfel {

constexpr int ACT_THREADS = 256;
constexpr float NORMALIZE_EPS = 1e-12f;   // torch.nn.functional.normalize default

struct ActivateArgs {
    int P;
    int F;   // floats per packed SH row = 3 * (1 + R)
    const float *scaling_raw, *rotation_raw, *opacity_raw, *features_dc, *features_rest;
    float *scaling, *rotation, *opacity, *features;
};

struct ActivateGradArgs {
    int P;
    int F;
    const float *rotation_raw, *scaling, *opacity;                      // raw quaternion, ACTIVATED scaling / opacity
    const float *g_scaling, *g_rotation, *g_opacity, *g_features;        // upstream gradients (g_features may be null)
    float *d_scaling_raw, *d_rotation_raw, *d_opacity_raw, *d_features_dc, *d_features_rest;
};

__host__ __device__ inline void activate_one(const ActivateArgs &a, const long long i)
{
    a.scaling[2 * i] = expf(a.scaling_raw[2 * i]);
    a.scaling[2 * i + 1] = expf(a.scaling_raw[2 * i + 1]);
    const float o = a.opacity_raw[i];
    a.opacity[i] = 1.0f / (1.0f + expf(-o));                             // ATen sigmoid: one / (one + exp(-a))
    const float qw = a.rotation_raw[4 * i], qx = a.rotation_raw[4 * i + 1], qy = a.rotation_raw[4 * i + 2],
                qz = a.rotation_raw[4 * i + 3];
    const float n = sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
    const float d = fmaxf(n, NORMALIZE_EPS);
    a.rotation[4 * i] = qw / d;
    a.rotation[4 * i + 1] = qx / d;
    a.rotation[4 * i + 2] = qy / d;
    a.rotation[4 * i + 3] = qz / d;
}

__host__ __device__ inline void activate_grad_one(const ActivateGradArgs &a, const long long i)
{
    // ExpBackward: grad * result
    a.d_scaling_raw[2 * i] = a.g_scaling[2 * i] * a.scaling[2 * i];
    a.d_scaling_raw[2 * i + 1] = a.g_scaling[2 * i + 1] * a.scaling[2 * i + 1];
    // sigmoid_backward: grad * (1 - y) * y
    const float y = a.opacity[i];
    a.d_opacity_raw[i] = a.g_opacity[i] * (1.0f - y) * y;
    // y = x / d, d = max(||x||, eps):  dx = g / d - x (g . x) / (d^2 ||x||)   (second term only where ||x|| >= eps)
    const float x[4] = {a.rotation_raw[4 * i], a.rotation_raw[4 * i + 1], a.rotation_raw[4 * i + 2], a.rotation_raw[4 * i + 3]};
    const float g[4] = {a.g_rotation[4 * i], a.g_rotation[4 * i + 1], a.g_rotation[4 * i + 2], a.g_rotation[4 * i + 3]};
    const float n = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
    const float d = fmaxf(n, NORMALIZE_EPS);
    const float gx = g[0] * x[0] + g[1] * x[1] + g[2] * x[2] + g[3] * x[3];
    const float k = (n >= NORMALIZE_EPS && n > 0.f) ? gx / (d * d * n) : 0.f;
#pragma unroll
    for (int c = 0; c < 4; c++) a.d_rotation_raw[4 * i + c] = g[c] / d - x[c] * k;
}

// The packed [P, F] SH block is moved as 128-bit words: gather from dc / rest forward, scatter the gradient backward.
// ACT_UNROLL words per thread, `stride` words apart (coalesced across the warp), all loads issued before the first
// store so that several 128-bit transactions per thread are in flight.
constexpr int ACT_UNROLL = 4;

__host__ __device__ inline void pack_features_words(const ActivateArgs &a, const long long t, const long long stride)
{
    const long long total = (long long)a.P * a.F, R3 = a.F - 3;
    float v[ACT_UNROLL][4];
#pragma unroll
    for (int w = 0; w < ACT_UNROLL; w++) {
        const long long e0 = 4 * (t + w * stride);
        long long i = e0 / a.F;
        int j = (int)(e0 - i * a.F);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            v[w][u] = (e0 + u < total) ? (j < 3 ? a.features_dc[3 * i + j] : a.features_rest[R3 * i + (j - 3)]) : 0.f;
            if (++j == a.F) { j = 0; i++; }
        }
    }
#pragma unroll
    for (int w = 0; w < ACT_UNROLL; w++) {
        const long long e0 = 4 * (t + w * stride);
        if (e0 + 4 <= total) {
            *reinterpret_cast<float4 *>(a.features + e0) = make_float4(v[w][0], v[w][1], v[w][2], v[w][3]);
        } else {   // last, partial word when P * F is not a multiple of 4 (or nothing at all past the end)
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (e0 + u < total) a.features[e0 + u] = v[w][u];
        }
    }
}

__host__ __device__ inline void unpack_feature_grad_words(const ActivateGradArgs &a, const long long t, const long long stride)
{
    const long long total = (long long)a.P * a.F, R3 = a.F - 3;
    float v[ACT_UNROLL][4];
#pragma unroll
    for (int w = 0; w < ACT_UNROLL; w++) {
        const long long e0 = 4 * (t + w * stride);
        v[w][0] = v[w][1] = v[w][2] = v[w][3] = 0.f;
        if (e0 + 4 <= total) {
            const float4 g = *reinterpret_cast<const float4 *>(a.g_features + e0);
            v[w][0] = g.x; v[w][1] = g.y; v[w][2] = g.z; v[w][3] = g.w;
        } else {
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (e0 + u < total) v[w][u] = a.g_features[e0 + u];
        }
    }
#pragma unroll
    for (int w = 0; w < ACT_UNROLL; w++) {
        const long long e0 = 4 * (t + w * stride);
        long long i = e0 / a.F;
        int j = (int)(e0 - i * a.F);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (e0 + u < total) {
                if (j < 3) a.d_features_dc[3 * i + j] = v[w][u];
                else a.d_features_rest[R3 * i + (j - 3)] = v[w][u];
            }
            if (++j == a.F) { j = 0; i++; }
        }
    }
}

}  // namespace surfel
