// adam_emul.cu -- TEST INFRASTRUCTURE: runs the __host__ __device__ bodies of the fused parameter-update kernels
// (streetunveiler_b200/csrc/adam_tile.cuh) on the CPU, chunk by chunk with tid = 0 / nthreads = 1 (see loss_emul.cu).
// All pointers are HOST pointers.  The launch descriptor is built exactly as csrc/adam.cu builds it.
#include <cmath>
#include <cstdint>

#include "../../include/surfel_rasterizer.h"
#include "../../streetunveiler_b200/csrc/adam_tile.cuh"

using namespace surfel;

extern "C" {

__attribute__((visibility("default"))) int emul_adam_step(int n_groups, const surfel_adam_group *groups, double beta1, double beta2, double eps)
{
    AdamLaunch L;
    L.n_groups = 0;
    L.w1 = (float)(1.0 - beta1);
    L.beta2 = (float)beta2;
    L.w2 = (float)(1.0 - beta2);
    L.eps = (float)eps;
    long long chunks = 0;
    for (int i = 0; i < n_groups; i++) {
        const surfel_adam_group &s = groups[i];
        if (s.n == 0) continue;
        AdamGroup &g = L.g[L.n_groups++];
        g.param = s.param; g.grad = s.grad; g.exp_avg = s.exp_avg; g.exp_avg_sq = s.exp_avg_sq;
        g.n = s.n;
        g.first_chunk = chunks;
        g.step_size = (float)(-(s.lr / (1.0 - std::pow(beta1, (double)s.step))));
        g.bc2_sqrt = (float)std::sqrt(1.0 - std::pow(beta2, (double)s.step));
        g.vec_ok = ((reinterpret_cast<uintptr_t>(s.param) | reinterpret_cast<uintptr_t>(s.grad) |
                     reinterpret_cast<uintptr_t>(s.exp_avg) | reinterpret_cast<uintptr_t>(s.exp_avg_sq)) & 15u) == 0;
        chunks += (s.n + ADAM_CHUNK - 1) / ADAM_CHUNK;
    }
    for (long long c = 0; c < chunks; c++) adam_chunk(L, c, 0, 1);
    return (int)chunks;
}

__attribute__((visibility("default"))) void emul_densification_stats(int P, const int *radii, const float *vgrad, float *max_radii2D, float *accum,
                                         float *denom)
{
    for (long long i = 0; i < P; i++) densification_stats_one(i, radii, vgrad, max_radii2D, accum, denom);
}

}  // extern "C"
