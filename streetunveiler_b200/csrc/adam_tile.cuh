// adam_tile.cuh -- bodies of the fused parameter-update kernels.
//
// Behavioural specification: train.py:168,197 and scene/gaussian_model.py:171-180,555-557 of the reference, which
// run torch.optim.Adam (PyTorch's multi-tensor implementation: ~12 foreach kernels over 6 parameter tensors) and
// five indexed PyTorch statements for the densification statistics.  Here: ONE launch updates every parameter group
// (each element read and written once: 28 B in+out per parameter), one more updates the three statistics arrays.
//
// As in loss_tile.cuh the bodies are __host__ __device__ so that tests/emul/ can run the same code on the CPU.
#pragma once
#include "common.cuh"

namespace surfel {

constexpr int ADAM_MAX_GROUPS = 8;
constexpr int ADAM_THREADS = 256;
constexpr int ADAM_VEC = 4;                                    // floats per 128-bit access
constexpr int ADAM_UNROLL = 4;                                 // independent 128-bit accesses in flight per thread
constexpr int ADAM_CHUNK = ADAM_THREADS * ADAM_VEC * ADAM_UNROLL;   // elements per CTA

struct AdamGroup {
    float *param;
    const float *grad;
    float *exp_avg;
    float *exp_avg_sq;
    long long n;            // elements
    long long first_chunk;  // exclusive prefix of chunk counts over the groups
    float step_size;        // -(lr / (1 - beta1^t))     torch/optim/adam.py: step_size, already negated
    float bc2_sqrt;         // sqrt(1 - beta2^t)
    int vec_ok;             // all four pointers 16-byte aligned
};

struct AdamLaunch {
    AdamGroup g[ADAM_MAX_GROUPS];
    int n_groups;
    float w1;               // 1 - beta1 (lerp weight)
    float beta2;
    float w2;               // 1 - beta2
    float eps;
};

// IEEE-rounded sqrtf / division that never leave the hardware fast path.  nvcc's sqrtf(x) branches to a software
// routine for x == 0 and x < 2^-101, and a / b does for zero, denormal or extreme-exponent operands.  In this workload
// those are the COMMON case, not the corner case: a Gaussian outside the current view has an exactly zero gradient,
// its second moment is zero or decays towards the denormals, and its first moment shrinks by 0.9 per step -- measured
// inside a training iteration the update kernel ran 2x slower than on dense random gradients because of it.  Scaling
// by an even power of two is exact and commutes with round-to-nearest as long as the result is a normal number, so
// the values are bit-identical to the plain expressions (a denormal quotient may differ in its last bit).
__host__ __device__ inline float sqrt_fast_path(const float v)   // v >= 0
{
    const bool tiny = v < 1.0e-30f;
    float s = tiny ? v * 18446744073709551616.0f : v;   // 2^64
    const bool zero = s == 0.f;
    s = zero ? 1.0f : s;
    float r = sqrtf(s);
    r = tiny ? r * 2.3283064365386963e-10f : r;         // 2^-32
    return zero ? 0.f : r;
}

__host__ __device__ inline float div_fast_path(const float a, const float b)   // b: a normal number of moderate size
{
    const bool tiny = fabsf(a) < 1.0e-18f;
    float s = tiny ? a * 18446744073709551616.0f : a;
    const bool zero = s == 0.f;
    s = zero ? 1.0f : s;
    float q = s / b;
    q = tiny ? q * 5.421010862427522e-20f : q;          // 2^-64
    return zero ? 0.f : q;
}

// torch's arithmetic, in its order: lerp (|w| < 0.5 branch), mul, addcmul, sqrt / bc2_sqrt + eps, addcdiv
__host__ __device__ inline void adam_element(float &p, const float g, float &m, float &v, const AdamLaunch &L,
                                             const AdamGroup &G)
{
    m = m + L.w1 * (g - m);
    v = v * L.beta2;
    v = v + (L.w2 * g) * g;
    const float denom = div_fast_path(sqrt_fast_path(v), G.bc2_sqrt) + L.eps;
    p = p + G.step_size * div_fast_path(m, denom);
}

// which group does chunk `c` belong to (groups are few: linear scan)
__host__ __device__ inline int adam_group_of_chunk(const AdamLaunch &L, const long long c)
{
    int gi = 0;
    while (gi + 1 < L.n_groups && c >= L.g[gi + 1].first_chunk) gi++;
    return gi;
}

// One CTA's chunk; `tid`/`nthreads` as in loss_tile.cuh.
__host__ __device__ inline void adam_chunk(const AdamLaunch &L, const long long chunk, const int tid, const int nthreads)
{
    const AdamGroup &G = L.g[adam_group_of_chunk(L, chunk)];
    const long long base = (chunk - G.first_chunk) * ADAM_CHUNK;
    const long long n_here = (G.n - base) < ADAM_CHUNK ? (G.n - base) : ADAM_CHUNK;
    if (G.vec_ok) {
        const long long nvec = n_here / ADAM_VEC;
        float4 *p4 = reinterpret_cast<float4 *>(G.param + base);
        const float4 *g4 = reinterpret_cast<const float4 *>(G.grad + base);
        float4 *m4 = reinterpret_cast<float4 *>(G.exp_avg + base);
        float4 *v4 = reinterpret_cast<float4 *>(G.exp_avg_sq + base);
        for (long long i0 = tid; i0 < nvec; i0 += (long long)nthreads * ADAM_UNROLL) {
            float4 p[ADAM_UNROLL], g[ADAM_UNROLL], m[ADAM_UNROLL], v[ADAM_UNROLL];
#pragma unroll
            for (int u = 0; u < ADAM_UNROLL; u++) {
                const long long i = i0 + (long long)u * nthreads;
                if (i < nvec) {
                    p[u] = p4[i];
                    g[u] = g4[i];
                    m[u] = m4[i];
                    v[u] = v4[i];
                }
            }
#pragma unroll
            for (int u = 0; u < ADAM_UNROLL; u++) {
                const long long i = i0 + (long long)u * nthreads;
                if (i < nvec) {
                    adam_element(p[u].x, g[u].x, m[u].x, v[u].x, L, G);
                    adam_element(p[u].y, g[u].y, m[u].y, v[u].y, L, G);
                    adam_element(p[u].z, g[u].z, m[u].z, v[u].z, L, G);
                    adam_element(p[u].w, g[u].w, m[u].w, v[u].w, L, G);
                    p4[i] = p[u];
                    m4[i] = m[u];
                    v4[i] = v[u];
                }
            }
        }
        for (long long i = nvec * ADAM_VEC + tid; i < n_here; i += nthreads)   // at most 3 tail elements of the group
            adam_element(G.param[base + i], G.grad[base + i], G.exp_avg[base + i], G.exp_avg_sq[base + i], L, G);
    } else {
        for (long long i = tid; i < n_here; i += nthreads)
            adam_element(G.param[base + i], G.grad[base + i], G.exp_avg[base + i], G.exp_avg_sq[base + i], L, G);
    }
}

// train.py:168 + scene/gaussian_model.py:555-557 for Gaussian i (visibility_filter = radii > 0)
__host__ __device__ inline void densification_stats_one(const long long i, const int *radii, const float *vgrad,
                                                        float *max_radii2D, float *xyz_gradient_accum, float *denom)
{
    const int r = radii[i];
    if (r <= 0) return;
    max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);
    const float x = vgrad[3 * i], y = vgrad[3 * i + 1], z = vgrad[3 * i + 2];
    xyz_gradient_accum[i] += sqrtf(x * x + y * y + z * z);
    denom[i] += 1.0f;
}

}  // namespace surfel
