// adam.cu -- fused parameter update: kernels and C ABI.  Bodies: adam_tile.cuh.
#include <cmath>
#include <cstdint>

#include "../../include/surfel_rasterizer.h"
#include "adam_tile.cuh"
#include "kernels.h"

namespace surfel {

__global__ void __launch_bounds__(ADAM_THREADS) adam_multi_kernel(const __grid_constant__ AdamLaunch L)
{
    adam_chunk(L, (long long)blockIdx.x, threadIdx.x, ADAM_THREADS);
}

__global__ void __launch_bounds__(ADAM_THREADS)
densification_stats_kernel(const int P, const int *__restrict__ radii, const float *__restrict__ vgrad,
                           float *__restrict__ max_radii2D, float *__restrict__ xyz_gradient_accum,
                           float *__restrict__ denom)
{
    const long long i = (long long)blockIdx.x * ADAM_THREADS + threadIdx.x;
    if (i < P) densification_stats_one(i, radii, vgrad, max_radii2D, xyz_gradient_accum, denom);
}

}  // namespace surfel

using namespace surfel;

extern "C" {

int surfel_adam_step(int n_groups, const surfel_adam_group *groups, double beta1, double beta2, double eps, void *stream)
{
    const char *where = "surfel_adam_step";
    if (n_groups < 0 || n_groups > ADAM_MAX_GROUPS || (n_groups > 0 && !groups))
        return surfel_internal_fail(where, "between 0 and 8 parameter groups per call");
    AdamLaunch L;
    L.n_groups = 0;
    L.w1 = (float)(1.0 - beta1);
    L.beta2 = (float)beta2;
    L.w2 = (float)(1.0 - beta2);
    L.eps = (float)eps;
    long long chunks = 0;
    for (int i = 0; i < n_groups; i++) {
        const surfel_adam_group &s = groups[i];
        if (s.n < 0 || s.step < 1) return surfel_internal_fail(where, "group with negative size or step < 1");
        if (s.n == 0) continue;
        if (!s.param || !s.grad || !s.exp_avg || !s.exp_avg_sq) return surfel_internal_fail(where, "null pointer in a group");
        AdamGroup &g = L.g[L.n_groups++];
        g.param = s.param;
        g.grad = s.grad;
        g.exp_avg = s.exp_avg;
        g.exp_avg_sq = s.exp_avg_sq;
        g.n = s.n;
        g.first_chunk = chunks;
        // torch/optim/adam.py: bias corrections in double on the host, applied as fp32 scalars on the device
        const double bc1 = 1.0 - std::pow(beta1, (double)s.step);
        const double bc2 = 1.0 - std::pow(beta2, (double)s.step);
        g.step_size = (float)(-(s.lr / bc1));
        g.bc2_sqrt = (float)std::sqrt(bc2);
        g.vec_ok = ((reinterpret_cast<uintptr_t>(s.param) | reinterpret_cast<uintptr_t>(s.grad) |
                     reinterpret_cast<uintptr_t>(s.exp_avg) | reinterpret_cast<uintptr_t>(s.exp_avg_sq)) & 15u) == 0;
        chunks += (s.n + ADAM_CHUNK - 1) / ADAM_CHUNK;
    }
    if (chunks == 0) return 0;
    if (chunks > 0x7fffffffLL) return surfel_internal_fail(where, "too many elements for one launch");
    adam_multi_kernel<<<(unsigned)chunks, ADAM_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(L);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : surfel_internal_fail(where, cudaGetErrorString(e));
}

int surfel_densification_stats(int P, const int *radii, const float *viewspace_grad, float *max_radii2D,
                               float *xyz_gradient_accum, float *denom, void *stream)
{
    const char *where = "surfel_densification_stats";
    if (P < 0 || (P > 0 && (!radii || !viewspace_grad || !max_radii2D || !xyz_gradient_accum || !denom)))
        return surfel_internal_fail(where, "bad arguments");
    if (P == 0) return 0;
    densification_stats_kernel<<<(P + ADAM_THREADS - 1) / ADAM_THREADS, ADAM_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        P, radii, viewspace_grad, max_radii2D, xyz_gradient_accum, denom);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : surfel_internal_fail(where, cudaGetErrorString(e));
}

}  // extern "C"
