// epilogue.cu -- fused render() epilogue ("next" row 1).
//
// Behavioural specification (reference, Python): gaussian_renderer/__init__.py:148-186 and
// utils/point_utils.py:8-37 -- view->world rotation of the rendered normals, expected depth
// D/alpha and median depth through nan_to_num, surf_depth = mix by depth_ratio, unprojection to world
// points, surface normal from central differences (cross product, F.normalize), weighted by the
// detached alpha.  The reference runs this as ~15 separate PyTorch kernels forward and as many
// backward on full-resolution planes; here it is one kernel each way, bound by HBM:
//   forward : reads the 7 allmap planes once, writes 10 planes (+2 optional pass-through copies)
//   backward: reads 3 allmap + 3 point + 10-12 gradient planes, writes 7 (23-25 planes, ~92 B / pixel)
// The 5-point (forward) and 13-point (backward) stencils re-read neighbours through L1/L2.
#include "../../include/surfel_rasterizer.h"
#include "common.cuh"
#include "kernels.h"

namespace surfel {

struct EpilogueCam {
    float W3[9];    // world_view_transform[:3,:3] (row-major, transposed-view layout)
    float Rc[9];    // c2w rotation
    float o[3];     // camera origin (c2w translation)
    float fx, fy, cx, cy;
};

// c2w = inverse(world_view_transform^T) for an affine matrix (last row 0 0 0 1), point_utils.py:9
__device__ __forceinline__ void load_camera(const float *__restrict__ vm, const float fx, const float fy, const int W,
                                            const int H, EpilogueCam &c)
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) c.W3[3 * i + j] = vm[4 * i + j];
    // w2c = vm^T: A[i][j] = vm[4 j + i], t[i] = vm[12 + i]
    float A[3][3], t[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        t[i] = vm[12 + i];
#pragma unroll
        for (int j = 0; j < 3; j++) A[i][j] = vm[4 * j + i];
    }
    const float c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1], c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2],
                c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
    const float det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
    const float id = 1.0f / det;
    c.Rc[0] = c00 * id;
    c.Rc[1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * id;
    c.Rc[2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * id;
    c.Rc[3] = c01 * id;
    c.Rc[4] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * id;
    c.Rc[5] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
    c.Rc[6] = c02 * id;
    c.Rc[7] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * id;
    c.Rc[8] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
#pragma unroll
    for (int i = 0; i < 3; i++) c.o[i] = -(c.Rc[3 * i] * t[0] + c.Rc[3 * i + 1] * t[1] + c.Rc[3 * i + 2] * t[2]);
    c.fx = fx; c.fy = fy; c.cx = W / 2.0f; c.cy = H / 2.0f;
}

// torch.nan_to_num(x, 0, 0): nan -> 0, +inf -> 0, -inf -> lowest finite
__device__ __forceinline__ float nan_to_num00(const float x)
{
    if (x != x) return 0.f;
    if (x == INFINITY) return 0.f;
    if (x == -INFINITY) return -3.402823466e+38f;
    return x;
}

__device__ __forceinline__ float3 ray_dir(const EpilogueCam &c, const int x, const int y)
{
    const float dx = ((float)x - c.cx) / c.fx, dy = ((float)y - c.cy) / c.fy;
    return make_float3(c.Rc[0] * dx + c.Rc[1] * dy + c.Rc[2], c.Rc[3] * dx + c.Rc[4] * dy + c.Rc[5],
                       c.Rc[6] * dx + c.Rc[7] * dy + c.Rc[8]);
}

__device__ __forceinline__ float surf_depth_at(const float *__restrict__ allmap, const size_t HW, const size_t p,
                                               const float ratio)
{
    const float expected = nan_to_num00(allmap[p] / allmap[HW + p]);
    const float median = nan_to_num00(allmap[5 * HW + p]);
    return expected * (1 - ratio) + ratio * median;
}

__device__ __forceinline__ float3 point_at(const EpilogueCam &c, const float *__restrict__ allmap, const size_t HW,
                                           const int W, const int x, const int y, const float ratio)
{
    const float d = surf_depth_at(allmap, HW, (size_t)y * W + x, ratio);
    const float3 r = ray_dir(c, x, y);
    return make_float3(d * r.x + c.o[0], d * r.y + c.o[1], d * r.z + c.o[2]);
}

__device__ __forceinline__ float3 sub3(const float3 a, const float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 cross3(const float3 a, const float3 b)
{
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

__global__ void __launch_bounds__(256)
epilogue_fwd_kernel(const int W, const int H, const float *__restrict__ allmap, const float *__restrict__ viewmatrix,
                    const float fx, const float fy, const float ratio, float *__restrict__ rend_normal,
                    float *__restrict__ surf_depth, float *__restrict__ surf_normal, float *__restrict__ surf_point,
                    float *__restrict__ rend_alpha, float *__restrict__ rend_dist)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    EpilogueCam c;
    load_camera(viewmatrix, fx, fy, W, H, c);
    const size_t HW = (size_t)W * H, p = (size_t)y * W + x;
    const float alpha = allmap[HW + p];
    if (rend_alpha) rend_alpha[p] = alpha;                      // optional copies of the two pass-through planes
    if (rend_dist) rend_dist[p] = allmap[6 * HW + p];
    const float n0 = allmap[2 * HW + p], n1 = allmap[3 * HW + p], n2 = allmap[4 * HW + p];
    rend_normal[p] = c.W3[0] * n0 + c.W3[1] * n1 + c.W3[2] * n2;
    rend_normal[HW + p] = c.W3[3] * n0 + c.W3[4] * n1 + c.W3[5] * n2;
    rend_normal[2 * HW + p] = c.W3[6] * n0 + c.W3[7] * n1 + c.W3[8] * n2;
    const float d = surf_depth_at(allmap, HW, p, ratio);
    const float3 r = ray_dir(c, x, y);
    surf_depth[p] = d;
    surf_point[p] = d * r.x + c.o[0];
    surf_point[HW + p] = d * r.y + c.o[1];
    surf_point[2 * HW + p] = d * r.z + c.o[2];
    float3 n = make_float3(0.f, 0.f, 0.f);
    if (x >= 1 && x < W - 1 && y >= 1 && y < H - 1) {
        const float3 dx = sub3(point_at(c, allmap, HW, W, x, y + 1, ratio), point_at(c, allmap, HW, W, x, y - 1, ratio));
        const float3 dy = sub3(point_at(c, allmap, HW, W, x + 1, y, ratio), point_at(c, allmap, HW, W, x - 1, y, ratio));
        const float3 cr = cross3(dx, dy);
        const float inv = 1.0f / fmaxf(sqrtf(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z), 1e-12f);
        n = make_float3(cr.x * inv, cr.y * inv, cr.z * inv);
    }
    surf_normal[p] = n.x * alpha;
    surf_normal[HW + p] = n.y * alpha;
    surf_normal[2 * HW + p] = n.z * alpha;
}

__device__ __forceinline__ float3 load3(const float *__restrict__ planes, const size_t HW, const size_t p)
{
    return make_float3(planes[p], planes[HW + p], planes[2 * HW + p]);
}

// Gradients w.r.t. the two finite differences of the normal centred at (cx, cy); zero outside the interior.
__device__ __forceinline__ void center_grads(const int W, const int H, const size_t HW, const int cx, const int cy,
                                             const float *__restrict__ pts, const float *__restrict__ g_sn,
                                             const float *__restrict__ allmap, float3 &g_dx, float3 &g_dy)
{
    g_dx = make_float3(0.f, 0.f, 0.f);
    g_dy = g_dx;
    if (cx < 1 || cx >= W - 1 || cy < 1 || cy >= H - 1) return;
    const size_t pc = (size_t)cy * W + cx;
    const float3 dx = sub3(load3(pts, HW, pc + W), load3(pts, HW, pc - W));
    const float3 dy = sub3(load3(pts, HW, pc + 1), load3(pts, HW, pc - 1));
    const float3 n = cross3(dx, dy);
    const float len = sqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
    const float alpha = allmap[HW + pc];
    float3 g = load3(g_sn, HW, pc);
    g = make_float3(g.x * alpha, g.y * alpha, g.z * alpha);     // surf_normal = normalize(n) * alpha.detach()
    float3 gn;
    if (len > 1e-12f) {
        const float inv = 1.0f / len;
        const float3 nh = make_float3(n.x * inv, n.y * inv, n.z * inv);
        const float dot = nh.x * g.x + nh.y * g.y + nh.z * g.z;
        gn = make_float3((g.x - nh.x * dot) * inv, (g.y - nh.y * dot) * inv, (g.z - nh.z * dot) * inv);
    } else {
        gn = make_float3(g.x * 1e12f, g.y * 1e12f, g.z * 1e12f);
    }
    g_dx = cross3(dy, gn);   // n = dx x dy
    g_dy = cross3(gn, dx);
}

__global__ void __launch_bounds__(256)
epilogue_bwd_kernel(const int W, const int H, const float *__restrict__ allmap, const float *__restrict__ pts,
                    const float *__restrict__ viewmatrix, const float fx, const float fy, const float ratio,
                    const float *__restrict__ g_rn, const float *__restrict__ g_sd, const float *__restrict__ g_sn,
                    const float *__restrict__ g_sp, const float *__restrict__ g_alpha, const float *__restrict__ g_dist,
                    float *__restrict__ dL_dallmap)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    EpilogueCam c;
    load_camera(viewmatrix, fx, fy, W, H, c);
    const size_t HW = (size_t)W * H, p = (size_t)y * W + x;
    // rotation of the rendered normals
    const float3 gr = load3(g_rn, HW, p);
    dL_dallmap[2 * HW + p] = c.W3[0] * gr.x + c.W3[3] * gr.y + c.W3[6] * gr.z;
    dL_dallmap[3 * HW + p] = c.W3[1] * gr.x + c.W3[4] * gr.y + c.W3[7] * gr.z;
    dL_dallmap[4 * HW + p] = c.W3[2] * gr.x + c.W3[5] * gr.y + c.W3[8] * gr.z;
    dL_dallmap[6 * HW + p] = g_dist ? g_dist[p] : 0.f;
    // gradient w.r.t. this pixel's world point: own term + the four normals whose stencil contains it
    float3 gp = load3(g_sp, HW, p);
    float3 a, b;
    center_grads(W, H, HW, x, y - 1, pts, g_sn, allmap, a, b);   // P is the y+1 sample of centre (x, y-1)
    gp = make_float3(gp.x + a.x, gp.y + a.y, gp.z + a.z);
    center_grads(W, H, HW, x, y + 1, pts, g_sn, allmap, a, b);   // ... the y-1 sample of centre (x, y+1)
    gp = make_float3(gp.x - a.x, gp.y - a.y, gp.z - a.z);
    center_grads(W, H, HW, x - 1, y, pts, g_sn, allmap, a, b);   // ... the x+1 sample of centre (x-1, y)
    gp = make_float3(gp.x + b.x, gp.y + b.y, gp.z + b.z);
    center_grads(W, H, HW, x + 1, y, pts, g_sn, allmap, a, b);   // ... the x-1 sample of centre (x+1, y)
    gp = make_float3(gp.x - b.x, gp.y - b.y, gp.z - b.z);
    const float3 r = ray_dir(c, x, y);
    const float gd = g_sd[p] + gp.x * r.x + gp.y * r.y + gp.z * r.z;
    // surf_depth = nan_to_num(D / alpha) (1 - ratio) + ratio nan_to_num(median); nan_to_num passes the
    // gradient only where its input is finite, and the division's backward divides by alpha again
    // (0/0 = NaN at alpha == 0, exactly what autograd produces in the reference)
    const float D = allmap[p], alpha = allmap[HW + p], med = allmap[5 * HW + p];
    const float e = D / alpha;
    const bool fin_e = (e == e) && (fabsf(e) != INFINITY);
    const bool fin_m = (med == med) && (fabsf(med) != INFINITY);
    const float ge = fin_e ? gd * (1 - ratio) : 0.f;
    dL_dallmap[p] = ge / alpha;
    dL_dallmap[HW + p] = -ge * D / (alpha * alpha) + (g_alpha ? g_alpha[p] : 0.f);
    dL_dallmap[5 * HW + p] = fin_m ? gd * ratio : 0.f;
}

}  // namespace surfel

using namespace surfel;

extern "C" {

int surfel_epilogue_forward(int width, int height, const float *allmap, const float *viewmatrix, float fx, float fy,
                            float depth_ratio, float *rend_normal, float *surf_depth, float *surf_normal,
                            float *surf_point, float *rend_alpha, float *rend_dist, void *stream)
{
    if (width <= 0 || height <= 0 || !allmap || !viewmatrix || !rend_normal || !surf_depth || !surf_normal || !surf_point)
        return surfel_internal_fail("surfel_epilogue_forward", "bad arguments");
    const dim3 grid((width + 31) / 32, (height + 7) / 8);
    epilogue_fwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(width, height, allmap, viewmatrix, fx, fy,
                                                                             depth_ratio, rend_normal, surf_depth,
                                                                             surf_normal, surf_point, rend_alpha,
                                                                             rend_dist);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : surfel_internal_fail("surfel_epilogue_forward", cudaGetErrorString(e));
}

int surfel_epilogue_backward(int width, int height, const float *allmap, const float *surf_point,
                             const float *viewmatrix, float fx, float fy, float depth_ratio, const float *g_rend_normal,
                             const float *g_surf_depth, const float *g_surf_normal, const float *g_surf_point,
                             const float *g_rend_alpha, const float *g_rend_dist, float *dL_dallmap, void *stream)
{
    if (width <= 0 || height <= 0 || !allmap || !surf_point || !viewmatrix || !g_rend_normal || !g_surf_depth ||
        !g_surf_normal || !g_surf_point || !dL_dallmap)
        return surfel_internal_fail("surfel_epilogue_backward", "bad arguments");
    const dim3 grid((width + 31) / 32, (height + 7) / 8);
    epilogue_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        width, height, allmap, surf_point, viewmatrix, fx, fy, depth_ratio, g_rend_normal, g_surf_depth, g_surf_normal,
        g_surf_point, g_rend_alpha, g_rend_dist, dL_dallmap);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : surfel_internal_fail("surfel_epilogue_backward", cudaGetErrorString(e));
}

}  // extern "C"
