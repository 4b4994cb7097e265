// activate.cu -- fused parameter activations and SH packing (caller-side row next to K1 / K8): kernels and C ABI.
// Bodies: activate_tile.cuh.  One launch each way: the first blocks run the per-Gaussian activations, the rest move
// the SH block as 128-bit words.
#include "../../include/surfel_rasterizer.h"
#include "activate_tile.cuh"
#include "kernels.h"

namespace surfel {

__global__ void __launch_bounds__(ACT_THREADS) activate_fwd_kernel(const ActivateArgs a, const int act_blocks)
{
    if ((int)blockIdx.x < act_blocks) {
        const long long i = (long long)blockIdx.x * ACT_THREADS + threadIdx.x;
        if (i < a.P) activate_one(a, i);
        return;
    }
    // block b of the packing part covers words [b * ACT_THREADS * ACT_UNROLL, ...): thread x takes x, x + 256, x + 512, ...
    const long long t = (long long)(blockIdx.x - act_blocks) * (ACT_THREADS * ACT_UNROLL) + threadIdx.x;
    pack_features_words(a, t, ACT_THREADS);
}

__global__ void __launch_bounds__(ACT_THREADS) activate_bwd_kernel(const ActivateGradArgs a, const int act_blocks)
{
    if ((int)blockIdx.x < act_blocks) {
        const long long i = (long long)blockIdx.x * ACT_THREADS + threadIdx.x;
        if (i < a.P) activate_grad_one(a, i);
        return;
    }
    const long long t = (long long)(blockIdx.x - act_blocks) * (ACT_THREADS * ACT_UNROLL) + threadIdx.x;
    unpack_feature_grad_words(a, t, ACT_THREADS);
}

}  // namespace surfel

using namespace surfel;

namespace {
bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
}  // namespace

extern "C" {

int surfel_activate_forward(int P, int sh_rest, const float *scaling_raw, const float *rotation_raw,
                            const float *opacity_raw, const float *features_dc, const float *features_rest,
                            float *scaling, float *rotation, float *opacity, float *features, void *stream)
{
    const char *where = "surfel_activate_forward";
    if (P < 0 || sh_rest < 0) return surfel_internal_fail(where, "bad sizes");
    if (P == 0) return 0;
    if (!scaling_raw || !rotation_raw || !opacity_raw || !scaling || !rotation || !opacity)
        return surfel_internal_fail(where, "NULL required pointer");
    const bool pack = features != nullptr;
    if (pack && (!features_dc || (sh_rest > 0 && !features_rest))) return surfel_internal_fail(where, "features without sources");
    const int F = 3 * (1 + sh_rest);
    if (pack && !aligned16(features)) return surfel_internal_fail(where, "the packed SH block must be 16-byte aligned");
    const ActivateArgs a{P, F, scaling_raw, rotation_raw, opacity_raw, features_dc, features_rest, scaling, rotation, opacity, features};
    const int act_blocks = (P + ACT_THREADS - 1) / ACT_THREADS;
    const long long words = pack ? ((long long)P * F + 3) / 4 : 0;
    const long long pack_blocks = (words + ACT_THREADS * ACT_UNROLL - 1) / (ACT_THREADS * ACT_UNROLL);
    if (act_blocks + pack_blocks > 0x7fffffffLL) return surfel_internal_fail(where, "too many elements for one launch");
    activate_fwd_kernel<<<(unsigned)(act_blocks + pack_blocks), ACT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a, act_blocks);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : surfel_internal_fail(where, cudaGetErrorString(e));
}

int surfel_activate_backward(int P, int sh_rest, const float *rotation_raw, const float *scaling, const float *opacity,
                             const float *g_scaling, const float *g_rotation, const float *g_opacity,
                             const float *g_features, float *d_scaling_raw, float *d_rotation_raw, float *d_opacity_raw,
                             float *d_features_dc, float *d_features_rest, void *stream)
{
    const char *where = "surfel_activate_backward";
    if (P < 0 || sh_rest < 0) return surfel_internal_fail(where, "bad sizes");
    if (P == 0) return 0;
    if (!rotation_raw || !scaling || !opacity || !g_scaling || !g_rotation || !g_opacity || !d_scaling_raw ||
        !d_rotation_raw || !d_opacity_raw)
        return surfel_internal_fail(where, "NULL required pointer");
    const bool unpack = g_features != nullptr;
    if (unpack && (!d_features_dc || (sh_rest > 0 && !d_features_rest))) return surfel_internal_fail(where, "g_features without destinations");
    const int F = 3 * (1 + sh_rest);
    if (unpack && !aligned16(g_features)) return surfel_internal_fail(where, "the packed SH gradient must be 16-byte aligned");
    const ActivateGradArgs a{P, F, rotation_raw, scaling, opacity, g_scaling, g_rotation, g_opacity, g_features,
                             d_scaling_raw, d_rotation_raw, d_opacity_raw, d_features_dc, d_features_rest};
    const int act_blocks = (P + ACT_THREADS - 1) / ACT_THREADS;
    const long long words = unpack ? ((long long)P * F + 3) / 4 : 0;
    const long long pack_blocks = (words + ACT_THREADS * ACT_UNROLL - 1) / (ACT_THREADS * ACT_UNROLL);
    if (act_blocks + pack_blocks > 0x7fffffffLL) return surfel_internal_fail(where, "too many elements for one launch");
    activate_bwd_kernel<<<(unsigned)(act_blocks + pack_blocks), ACT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a, act_blocks);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : surfel_internal_fail(where, cudaGetErrorString(e));
}

}  // extern "C"
