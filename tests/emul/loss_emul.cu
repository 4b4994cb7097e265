// loss_emul.cu -- TEST INFRASTRUCTURE: runs the __host__ __device__ tile bodies of the fused loss kernels
// (streetunveiler_b200/csrc/loss_tile.cuh) on the CPU, tile by tile with tid = 0 / nthreads = 1, so that the
// CPU test-suite can check the CUDA code's indexing, halo and zero-padding logic against oracle/loss_oracle.py
// without a GPU.  Built on demand by tests/test_loss_emulation.py into tests/emul/libloss_emul.so; never loaded
// by the product (streetunveiler_b200/ only ever opens libsurfel_b200.so).  All pointers are HOST pointers.
#include <vector>

#include "../../streetunveiler_b200/csrc/loss_tile.cuh"

using namespace surfel;

extern "C" {

__attribute__((visibility("default"))) void emul_loss_photometric_forward(int W, int H, const float *render, const float *alpha, const float *sky,
                                              const float *gt, float *deriv, float *out_means)
{
    const LossImages im{render, alpha, sky, gt, W, H};
    const LossWindow win = make_loss_window();
    const int gx = (W + LT - 1) / LT, gy = (H + LT - 1) / LT, nblocks = gx * gy;
    std::vector<float> partials(2 * (size_t)nblocks);
    LossFwdSmem *s = new LossFwdSmem;
    for (int by = 0; by < gy; by++)
        for (int bx = 0; bx < gx; bx++) {
            float a = 0.f, b = 0.f;
            loss_fwd_tile(*s, im, win, bx * LT, by * LT, 0, 1, deriv, a, b);
            partials[by * gx + bx] = a;
            partials[nblocks + by * gx + bx] = b;
        }
    delete s;
    const double scale = 1.0 / (3.0 * (double)W * (double)H);
    for (int k = 0; k < 2; k++) out_means[k] = (float)(partial_column_sum(partials.data(), nblocks, k, 0, 1) * scale);
}

__attribute__((visibility("default"))) void emul_loss_photometric_backward(int W, int H, const float *render, const float *alpha, const float *sky,
                                               const float *gt, const float *deriv, const float *upstream,
                                               float *d_render, float *d_alpha, float *d_sky)
{
    const LossImages im{render, alpha, sky, gt, W, H};
    const LossWindow win = make_loss_window();
    const int gx = (W + LT - 1) / LT, gy = (H + LT - 1) / LT;
    const float inv_n = 1.0f / (3.0f * (float)W * (float)H);
    LossBwdSmem *s = new LossBwdSmem;
    for (int by = 0; by < gy; by++)
        for (int bx = 0; bx < gx; bx++)
            loss_bwd_tile(*s, im, win, bx * LT, by * LT, 0, 1, deriv, upstream[0] * inv_n, upstream[1] * inv_n, d_render,
                          d_alpha, d_sky);
    delete s;
}

__attribute__((visibility("default"))) void emul_loss_regulariser(int W, int H, const float *rn, const float *sn, const float *dist, const float *upstream,
                                      float *out_means, float *d_rn, float *d_sn, float *d_dist)
{
    const size_t HW = (size_t)W * H;
    const int nblocks = 7;   // any partition of the pixels must give the same sums
    std::vector<float> partials(2 * nblocks, 0.f);
    for (int b = 0; b < nblocks; b++) regulariser_sums(rn, sn, dist, HW, b, nblocks, partials[b], partials[nblocks + b]);
    for (int k = 0; k < 2; k++) out_means[k] = (float)(partial_column_sum(partials.data(), nblocks, k, 0, 1) / (double)HW);
    regulariser_grads(rn, sn, HW, 0, 1, upstream[0] / (float)HW, upstream[1] / (float)HW, d_rn, d_sn, d_dist);
}

}  // extern "C"
