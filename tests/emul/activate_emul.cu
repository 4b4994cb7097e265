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
    for (long long t = 0; 4 * t < (long long)P * a.F; t++) pack_features_word(a, t);
}

__attribute__((visibility("default"))) void emul_activate_backward(int P, int sh_rest, const float *rotation_raw, const float *scaling, const float *opacity,
                                       const float *g_scaling, const float *g_rotation, const float *g_opacity,
                                       const float *g_features, float *d_scaling_raw, float *d_rotation_raw,
                                       float *d_opacity_raw, float *d_dc, float *d_rest)
{
    const ActivateGradArgs a{P, 3 * (1 + sh_rest), rotation_raw, scaling, opacity, g_scaling, g_rotation, g_opacity, g_features,
                             d_scaling_raw, d_rotation_raw, d_opacity_raw, d_dc, d_rest};
    for (long long i = 0; i < P; i++) activate_grad_one(a, i);
    for (long long t = 0; 4 * t < (long long)P * a.F; t++) unpack_feature_grad_word(a, t);
}

}  // extern "C"
