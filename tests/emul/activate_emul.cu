// activate_emul.cu -- TEST INFRASTRUCTURE: runs the __host__ __device__ bodies of the fused parameter-activation kernels
// (streetunveiler_b200/csrc/activate_tile.cuh) on the CPU (see loss_emul.cu).  All pointers are HOST pointers.
#include "../../streetunveiler_b200/csrc/activate_tile.cuh"

using namespace surfel;

extern "C" {

__attribute__((visibility("default"))) void emul_activate_forward(int P, int sh_rest, const float *scaling_raw, const float *rotation_raw,
                                      const float *opacity_raw, const float *dc, const float *rest, float *scaling,
                                      float *rotation, float *opacity, float *features)
{
    const ActivateArgs a{P, 3 * (1 + sh_rest), scaling_raw, rotation_raw, opacity_raw, dc, rest, scaling, rotation, opacity, features};
    for (long long i = 0; i < P; i++) activate_one(a, i);
    // the launch geometry of csrc/activate.cu: blocks of ACT_THREADS threads, ACT_UNROLL words per thread, ACT_THREADS apart
    const long long words = ((long long)P * a.F + 3) / 4, per_block = (long long)ACT_THREADS * ACT_UNROLL;
    for (long long b = 0; b * per_block < words; b++)
        for (int x = 0; x < ACT_THREADS; x++) pack_features_words(a, b * per_block + x, ACT_THREADS);
}

__attribute__((visibility("default"))) void emul_activate_backward(int P, int sh_rest, const float *rotation_raw, const float *scaling, const float *opacity,
                                       const float *g_scaling, const float *g_rotation, const float *g_opacity,
                                       const float *g_features, float *d_scaling_raw, float *d_rotation_raw,
                                       float *d_opacity_raw, float *d_dc, float *d_rest)
{
    const ActivateGradArgs a{P, 3 * (1 + sh_rest), rotation_raw, scaling, opacity, g_scaling, g_rotation, g_opacity, g_features,
                             d_scaling_raw, d_rotation_raw, d_opacity_raw, d_dc, d_rest};
    for (long long i = 0; i < P; i++) activate_grad_one(a, i);
    const long long words = ((long long)P * a.F + 3) / 4, per_block = (long long)ACT_THREADS * ACT_UNROLL;
    for (long long b = 0; b * per_block < words; b++)
        for (int x = 0; x < ACT_THREADS; x++) unpack_feature_grad_words(a, b * per_block + x, ACT_THREADS);
}

}  // extern "C"
