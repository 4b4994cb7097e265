// preprocess.cu -- per-Gaussian forward projection (K1) and per-Gaussian backward (K8).
//
// Behavioural specification (reference, "RAST/" = submodules/diff-surfel-rasterization/):
//   forward : RAST/cuda_rasterizer/forward.cu:148-251 (preprocessCUDA), :75-115 (compute_transmat),
//             :119-145 (compute_aabb), :20-71 (computeColorFromSH), auxiliary.h:67-77,185-235
//   backward: RAST/cuda_rasterizer/backward.cu:581-636 (preprocessCUDA), :449-579
//             (compute_transmat_aabb), :20-139 (SH VJP), auxiliary.h:128-138,238-282
//
// The arithmetic keeps the reference's operation order (including glm's column-major product
// order) because radii / tile rects / sort keys must come out bit-identical; the memory side is
// new: one thread per Gaussian, 128-bit loads of rotation / SH rows, and ONE packed 96-B record
// per Gaussian (common.cuh) instead of six separate arrays, including a conservative alpha>=1/255
// footprint (ellipse + low-pass disc) used by the render kernels for sub-tile culling.
#include <cstdio>

#include "async_copy.cuh"
#include "common.cuh"
#include "kernels.h"

namespace surfel {

__device__ __constant__ float kSH_C0 = 0.28209479177387814f;
__device__ __constant__ float kSH_C1 = 0.4886025119029199f;
__device__ __constant__ float kSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                           -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float kSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                           0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                           -0.5900435899266435f};

// Rotation matrix columns from a (w,x,y,z) quaternion, normalised in-kernel (auxiliary.h:213-235).
__device__ __forceinline__ void quat_columns(const float4 q, float R[3][3])
{
    const float s = rsqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    const float w = q.x * s, x = q.y * s, y = q.z * s, z = q.w * s;
    R[0][0] = 1.f - 2.f * (y * y + z * z);
    R[0][1] = 2.f * (x * y + w * z);
    R[0][2] = 2.f * (x * z - w * y);
    R[1][0] = 2.f * (x * y - w * z);
    R[1][1] = 1.f - 2.f * (x * x + z * z);
    R[1][2] = 2.f * (y * z + w * x);
    R[2][0] = 2.f * (x * z + w * y);
    R[2][1] = 2.f * (y * z - w * x);
    R[2][2] = 1.f - 2.f * (x * x + y * y);
}

// L = R * diag(sx, sy, 1) in glm's accumulation order (type_mat3x3.inl operator*).
__device__ __forceinline__ void scaled_frame(const float R[3][3], float sx, float sy, float L[3][3])
{
    const float S[3][3] = {{sx, 0.f, 0.f}, {0.f, sy, 0.f}, {0.f, 0.f, 1.f}};
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++) L[c][r] = R[0][r] * S[c][0] + R[1][r] * S[c][1] + R[2][r] * S[c][2];
}

__device__ __forceinline__ float3 xform_point(const float3 p, const float *m)
{
    return make_float3(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}
__device__ __forceinline__ float3 xform_vec(const float3 p, const float *m)
{
    return make_float3(m[0] * p.x + m[4] * p.y + m[8] * p.z, m[1] * p.x + m[5] * p.y + m[9] * p.z,
                       m[2] * p.x + m[6] * p.y + m[10] * p.z);
}
__device__ __forceinline__ float3 xform_vec_transposed(const float3 p, const float *m)
{
    return make_float3(m[0] * p.x + m[1] * p.y + m[2] * p.z, m[4] * p.x + m[5] * p.y + m[6] * p.z,
                       m[8] * p.x + m[9] * p.y + m[10] * p.z);
}

__device__ __forceinline__ float dot3(const float a0, const float a1, const float a2, const float b0, const float b1,
                                      const float b2)
{
    const float t0 = a0 * b0, t1 = a1 * b1, t2 = a2 * b2;
    return t0 + t1 + t2;
}

// Tile rectangle of a splat (auxiliary.h:67-77): int truncation of float quotients, clamped.
__device__ __forceinline__ void tile_rect(const float px, const float py, const int max_radius, const int gx,
                                          const int gy, int &x0, int &y0, int &x1, int &y1)
{
    x0 = min(gx, max(0, (int)((px - max_radius) / TILE_X)));
    y0 = min(gy, max(0, (int)((py - max_radius) / TILE_Y)));
    x1 = min(gx, max(0, (int)((px + max_radius + TILE_X - 1) / TILE_X)));
    y1 = min(gy, max(0, (int)((py + max_radius + TILE_Y - 1) / TILE_Y)));
}

// Conservative footprint of the region where this splat can reach alpha >= 1/255, i.e.
// rho = min(rho3d, rho2d) <= tau = 2 ln(255 o):
//   * rho2d <= tau is the disc |p - mean2D|^2 <= tau / 2 (the render kernels recompute tau from opacity),
//   * rho3d <= tau is a conic in pixel space.  Two bounds of it are stored:
//     (a) its axis-aligned bounding box from the same closed form the reference uses for the 3-sigma
//         extent (forward.cu:119-145 with cutoff^2 = tau) -- numerically robust (no near-singular
//         division), with explicit slack for the fp32 cancellation in c^2 - q -- united with the disc's
//         box and quantised outward to 8x8-pixel units (4 x u8 in one word);
//     (b) when well conditioned, the ellipse itself (X-e)^T M (X-e) <= 1: with pixel offsets (X,Y) from
//         the projected centre c the intersection point is p = a0 X + a1 Y + a2, a0 = Tv' x Tw,
//         a1 = Tw x Tu', a2 = Tu' x Tv' (Tu' = Tu - c.x Tw, Tv' = Tv - c.y Tw), and
//         rho3d <= tau <=> p.x^2 + p.y^2 - tau p.z^2 <= 0.  Thin diagonal needles make the 2x2 system
//         for the centre ill-conditioned (in fp32: errors of pixels against a sub-pixel minor axis), so
//         it is solved in fp64 and the render kernels add an evaluation-error term to their threshold.
// tau is inflated (x1.002 + 0.01) to absorb fp32 rounding of the per-pixel evaluation.
// out: word0 = bbox (x0 | x1<<8 | y0<<16 | y1<<24, 8-px units, x1<x0: never contributes),
//      (e.x, e.y, M00, M01, M11); M00 == 0: no ellipse bound.
__device__ __forceinline__ void cull_footprint(const float Tm[3][3], const float mx, const float my,
                                               const float opacity, const int W, const int H, uint32_t &bbox,
                                               float out[5])
{
    const uint32_t BOX_ALL = 0u | (255u << 8) | (0u << 16) | (255u << 24), BOX_NONE = 1u | (0u << 8) | (1u << 16);
    bbox = BOX_ALL;
    out[0] = mx; out[1] = my; out[2] = 0.f; out[3] = 0.f; out[4] = 0.f;
    const float o255 = 255.0f * opacity;
    if (!(o255 >= 1.0f)) {
        if (o255 < 0.999f) bbox = BOX_NONE;  // o exp(<=0) < 1/255 everywhere (1e-3 guard for the product's rounding)
        return;
    }
    const float tau = 2.0f * __logf(o255) * 1.002f + 0.01f;
    const float Tw[3] = {Tm[2][0], Tm[2][1], Tm[2][2]};

    // ---- (a) bounding box ----
    const float rl = sqrtf(0.5f * tau) + 0.5f;  // low-pass disc
    float lx = mx - rl, hx = mx + rl, ly = my - rl, hy = my + rl;
    const float tz2 = Tw[2] * Tw[2];
    const float dd = tau * (Tw[0] * Tw[0] + Tw[1] * Tw[1]) - tz2;
    if (!(dd < -1e-3f * tz2)) return;  // conic not safely bounded: evaluate everywhere
    {
        const float inv = 1.0f / dd;
        const float f0 = tau * inv, f2 = -inv;
        const float cx = f0 * (Tm[0][0] * Tw[0] + Tm[0][1] * Tw[1]) + f2 * Tm[0][2] * Tw[2];
        const float cy = f0 * (Tm[1][0] * Tw[0] + Tm[1][1] * Tw[1]) + f2 * Tm[1][2] * Tw[2];
        const float qx = f0 * (Tm[0][0] * Tm[0][0] + Tm[0][1] * Tm[0][1]) + f2 * Tm[0][2] * Tm[0][2];
        const float qy = f0 * (Tm[1][0] * Tm[1][0] + Tm[1][1] * Tm[1][1]) + f2 * Tm[1][2] * Tm[1][2];
        const float hx2 = cx * cx - qx, hy2 = cy * cy - qy;
        if (!(hx2 == hx2 && hy2 == hy2 && fabsf(cx) < 1e7f && fabsf(cy) < 1e7f)) return;
        // fp32 cancellation in c^2 - q: |error| <= ~1e-6 (c^2 + |q|); plus relative and absolute slack
        const float sx2 = 1e-6f * (cx * cx + fabsf(qx)), sy2 = 1e-6f * (cy * cy + fabsf(qy));
        const float ex = sqrtf(fmaxf(hx2 + sx2, 0.f)) * 1.002f + 0.75f;
        const float ey = sqrtf(fmaxf(hy2 + sy2, 0.f)) * 1.002f + 0.75f;
        lx = fminf(lx, cx - ex); hx = fmaxf(hx, cx + ex);
        ly = fminf(ly, cy - ey); hy = fmaxf(hy, cy + ey);
    }
    {
        const float x0 = fmaxf(floorf(lx), 0.f), y0 = fmaxf(floorf(ly), 0.f);
        const float x1 = fminf(ceilf(hx), (float)(W - 1)), y1 = fminf(ceilf(hy), (float)(H - 1));
        if (x1 < x0 || y1 < y0) { bbox = BOX_NONE; return; }
        const uint32_t bx0 = min((uint32_t)x0 >> 3, 255u), bx1 = min((uint32_t)x1 >> 3, 255u);
        const uint32_t by0 = min((uint32_t)y0 >> 3, 255u), by1 = min((uint32_t)y1 >> 3, 255u);
        bbox = bx0 | (bx1 << 8) | (by0 << 16) | (by1 << 24);  // a clamped 255 means "to the image border"
    }

    // ---- (b) ellipse: solved in double precision (a thin diagonal needle makes the 2x2 system for the
    //      centre lose ~1/aspect^2 of the digits; B200 has full-rate fp64 to spare in this kernel) ----
    double Tu[3], Tv[3];
    const double Twd[3] = {(double)Tw[0], (double)Tw[1], (double)Tw[2]};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        Tu[i] = (double)Tm[0][i] - (double)mx * Twd[i];
        Tv[i] = (double)Tm[1][i] - (double)my * Twd[i];
    }
    const double td = (double)tau;
    const double a0[3] = {Tv[1] * Twd[2] - Tv[2] * Twd[1], Tv[2] * Twd[0] - Tv[0] * Twd[2], Tv[0] * Twd[1] - Tv[1] * Twd[0]};
    const double a1[3] = {Twd[1] * Tu[2] - Twd[2] * Tu[1], Twd[2] * Tu[0] - Twd[0] * Tu[2], Twd[0] * Tu[1] - Twd[1] * Tu[0]};
    const double a2[3] = {Tu[1] * Tv[2] - Tu[2] * Tv[1], Tu[2] * Tv[0] - Tu[0] * Tv[2], Tu[0] * Tv[1] - Tu[1] * Tv[0]};
    const double q00 = a0[0] * a0[0] + a0[1] * a0[1] - td * a0[2] * a0[2];
    const double q01 = a0[0] * a1[0] + a0[1] * a1[1] - td * a0[2] * a1[2];
    const double q11 = a1[0] * a1[0] + a1[1] * a1[1] - td * a1[2] * a1[2];
    const double q02 = a0[0] * a2[0] + a0[1] * a2[1] - td * a0[2] * a2[2];
    const double q12 = a1[0] * a2[0] + a1[1] * a2[1] - td * a1[2] * a2[2];
    const double q22 = a2[0] * a2[0] + a2[1] * a2[1] - td * a2[2] * a2[2];
    const double det = q00 * q11 - q01 * q01;
    if (!(q00 > 0.0 && q11 > 0.0 && det > 1e-7 * q00 * q11)) return;
    const double inv = 1.0 / det;
    const double ex = (q01 * q12 - q11 * q02) * inv;
    const double ey = (q01 * q02 - q00 * q12) * inv;
    const double fmin = q22 + q02 * ex + q12 * ey;
    if (!(fabs(ex) < 1e5 && fabs(ey) < 1e5) || !(fmin < 0.0)) return;  // (empty or odd conic: box only)
    const double sc = -1.0 / fmin;
    const float m00 = (float)(q00 * sc), m01 = (float)(q01 * sc), m11 = (float)(q11 * sc);
    if (!(m00 > 0.f && m11 > 0.f && m00 < 1e12f && m11 < 1e12f)) return;
    out[0] = (float)((double)mx + ex); out[1] = (float)((double)my + ey);
    out[2] = m00; out[3] = m01; out[4] = m11;
}

// =============================================================================================
// K1: forward preprocess
// =============================================================================================
__global__ void __launch_bounds__(256)
preprocess_fwd_kernel(const int P, const int D, const int M, const float *__restrict__ means3D,
                      const float2 *__restrict__ scales, const float scale_modifier,
                      const float4 *__restrict__ rotations, const float *__restrict__ opacities,
                      const float *__restrict__ shs, const float *__restrict__ transMat_precomp,
                      const float *__restrict__ colors_precomp, const float *__restrict__ viewmatrix,
                      const float *__restrict__ projmatrix, const float *__restrict__ cam_pos, const int W,
                      const int H, const int gx, const int gy, int *__restrict__ radii, float *__restrict__ rec,
                      uint32_t *__restrict__ tiles_touched, uint32_t *__restrict__ depth_key,
                      uint32_t *__restrict__ idx_in, uint8_t *__restrict__ clamped, const int prefiltered)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;

    int radius_out = 0;
    uint32_t tiles_out = 0;
    uint32_t key_out = 0xFFFFFFFFu;

    const float3 p_orig = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
    const float3 p_view = xform_point(p_orig, viewmatrix);

    do {
        if (p_view.z <= 0.2f) {  // auxiliary.h:199
            if (prefiltered) {
                printf("Point is filtered although prefiltered is set. This shouldn't happen!");
                __trap();
            }
            break;
        }
        float Tm[3][3];  // Tm[0] = Tu, Tm[1] = Tv, Tm[2] = Tw
        float3 normal;
        if (transMat_precomp == nullptr) {
            float R[3][3], L[3][3];
            const float2 sc = scales[idx];
            quat_columns(rotations[idx], R);
            scaled_frame(R, scale_modifier * sc.x, scale_modifier * sc.y, L);
            // A = transpose(splat2world): columns (L0[k], L1[k], p[k]) and (0,0,1)
            const float A[4][3] = {{L[0][0], L[1][0], p_orig.x}, {L[0][1], L[1][1], p_orig.y},
                                   {L[0][2], L[1][2], p_orig.z}, {0.f, 0.f, 1.f}};
            float B[4][3];  // A * world2ndc, world2ndc[c][k] = projmatrix[c + 4k]
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int r = 0; r < 3; r++)
                    B[c][r] = A[0][r] * projmatrix[c] + A[1][r] * projmatrix[c + 4] + A[2][r] * projmatrix[c + 8] +
                              A[3][r] * projmatrix[c + 12];
            const float N[3][4] = {{float(W) / 2.0f, 0.f, 0.f, float(W - 1) / 2.0f},
                                   {0.f, float(H) / 2.0f, 0.f, float(H - 1) / 2.0f},
                                   {0.f, 0.f, 0.f, 1.f}};
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int r = 0; r < 3; r++)
                    Tm[j][r] = B[0][r] * N[j][0] + B[1][r] * N[j][1] + B[2][r] * N[j][2] + B[3][r] * N[j][3];
            normal = xform_vec(make_float3(L[2][0], L[2][1], L[2][2]), viewmatrix);
        } else {
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int r = 0; r < 3; r++) Tm[j][r] = transMat_precomp[9 * idx + 3 * j + r];
            normal = make_float3(0.f, 0.f, 1.f);
        }

        // dual-visible flip (forward.cu:209-214)
        const float cosv = -(p_view.x * normal.x + p_view.y * normal.y + p_view.z * normal.z);
        if (cosv == 0) break;
        const float mult = cosv > 0 ? 1.f : -1.f;
        normal = make_float3(mult * normal.x, mult * normal.y, mult * normal.z);

        // centre + extent at cutoff 3 (forward.cu:119-145)
        const float t0 = 3.0f * 3.0f, t1 = 3.0f * 3.0f, t2 = -1.0f;
        const float d = dot3(t0, t1, t2, Tm[2][0] * Tm[2][0], Tm[2][1] * Tm[2][1], Tm[2][2] * Tm[2][2]);
        if (d == 0.0f) break;
        const float invd = 1 / d;
        const float f0 = invd * t0, f1 = invd * t1, f2 = invd * t2;
        const float cx = dot3(f0, f1, f2, Tm[0][0] * Tm[2][0], Tm[0][1] * Tm[2][1], Tm[0][2] * Tm[2][2]);
        const float cy = dot3(f0, f1, f2, Tm[1][0] * Tm[2][0], Tm[1][1] * Tm[2][1], Tm[1][2] * Tm[2][2]);
        const float h0x = cx * cx - dot3(f0, f1, f2, Tm[0][0] * Tm[0][0], Tm[0][1] * Tm[0][1], Tm[0][2] * Tm[0][2]);
        const float h0y = cy * cy - dot3(f0, f1, f2, Tm[1][0] * Tm[1][0], Tm[1][1] * Tm[1][1], Tm[1][2] * Tm[1][2]);
        const float ex = sqrtf((1e-4f < h0x) ? h0x : 1e-4f);
        const float ey = sqrtf((1e-4f < h0y) ? h0y : 1e-4f);
        const float radius = ceilf(fmaxf(fmaxf(ex, ey), 3.0f * FILTER_SIZE));

        int rx0, ry0, rx1, ry1;
        tile_rect(cx, cy, (int)radius, gx, gy, rx0, ry0, rx1, ry1);
        if ((rx1 - rx0) * (ry1 - ry0) == 0) break;

        // colour (forward.cu:20-71)
        float rgb[3];
        uint32_t clamp_bits = 0;
        if (colors_precomp == nullptr) {
            const float3 cam = make_float3(cam_pos[0], cam_pos[1], cam_pos[2]);
            float dx = p_orig.x - cam.x, dy = p_orig.y - cam.y, dz = p_orig.z - cam.z;
            const float len = sqrtf(dot3(dx, dy, dz, dx, dy, dz));
            const float x = dx / len, y = dy / len, z = dz / len;
            // the SH row of one Gaussian is 12*M bytes, 16-B aligned whenever M is a multiple of 4
            float sh[48];
            const float *row = shs + (size_t)idx * M * 3;
            const int ncoef = (D + 1) * (D + 1);
            if ((M & 3) == 0 && (reinterpret_cast<uintptr_t>(shs) & 15) == 0) {   // rows 16-B aligned: 128-bit loads
                const float4 *row4 = reinterpret_cast<const float4 *>(row);
#pragma unroll
                for (int i = 0; i < 12; i++)
                    if (i * 4 < ncoef * 3) {
                        const float4 v = __ldg(row4 + i);
                        sh[4 * i] = v.x; sh[4 * i + 1] = v.y; sh[4 * i + 2] = v.z; sh[4 * i + 3] = v.w;
                    }
            } else {
#pragma unroll
                for (int i = 0; i < 48; i++)
                    if (i < ncoef * 3) sh[i] = __ldg(row + i);
            }
#pragma unroll
            for (int c = 0; c < 3; c++) {
#define SHC(k) sh[3 * (k) + c]
                float res = kSH_C0 * SHC(0);
                if (D > 0) {
                    res = res - kSH_C1 * y * SHC(1) + kSH_C1 * z * SHC(2) - kSH_C1 * x * SHC(3);
                    if (D > 1) {
                        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                        res = res + kSH_C2[0] * xy * SHC(4) + kSH_C2[1] * yz * SHC(5) +
                              kSH_C2[2] * (2.0f * zz - xx - yy) * SHC(6) + kSH_C2[3] * xz * SHC(7) +
                              kSH_C2[4] * (xx - yy) * SHC(8);
                        if (D > 2) {
                            res = res + kSH_C3[0] * y * (3.0f * xx - yy) * SHC(9) + kSH_C3[1] * xy * z * SHC(10) +
                                  kSH_C3[2] * y * (4.0f * zz - xx - yy) * SHC(11) +
                                  kSH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SHC(12) +
                                  kSH_C3[4] * x * (4.0f * zz - xx - yy) * SHC(13) + kSH_C3[5] * z * (xx - yy) * SHC(14) +
                                  kSH_C3[6] * x * (xx - 3.0f * yy) * SHC(15);
                        }
                    }
                }
#undef SHC
                res += 0.5f;
                if (res < 0) clamp_bits |= (1u << c);
                rgb[c] = fmaxf(res, 0.0f);
            }
        } else {
            rgb[0] = colors_precomp[3 * idx];
            rgb[1] = colors_precomp[3 * idx + 1];
            rgb[2] = colors_precomp[3 * idx + 2];
        }

        const float opacity = opacities[idx];
        float fp[5];
        uint32_t bbox;
        cull_footprint(Tm, cx, cy, opacity, W, H, bbox, fp);

        float4 *r4 = reinterpret_cast<float4 *>(rec + (size_t)idx * REC_FLOATS);
        r4[0] = make_float4(Tm[0][0], Tm[0][1], Tm[0][2], Tm[1][0]);
        r4[1] = make_float4(Tm[1][1], Tm[1][2], Tm[2][0], Tm[2][1]);
        r4[2] = make_float4(Tm[2][2], cx, cy, opacity);
        r4[3] = make_float4(normal.x, normal.y, normal.z, rgb[0]);
        r4[4] = make_float4(rgb[1], rgb[2], __uint_as_float(bbox), fp[0]);
        r4[5] = make_float4(fp[1], fp[2], fp[3], fp[4]);
        clamped[idx] = (uint8_t)clamp_bits;

        radius_out = (int)radius;
        tiles_out = (uint32_t)((ry1 - ry0) * (rx1 - rx0));
        key_out = __float_as_uint(p_view.z);
    } while (false);

    radii[idx] = radius_out;
    if (tiles_touched) tiles_touched[idx] = tiles_out;   // NULL in the sharded path (counted per tile-row window later)
    depth_key[idx] = key_out;
    if (idx_in) idx_in[idx] = (uint32_t)idx;
}

void launch_preprocess_fwd(const PreprocessFwdArgs &a, cudaStream_t stream)
{
    if (a.P == 0) return;
    preprocess_fwd_kernel<<<(a.P + 255) / 256, 256, 0, stream>>>(
        a.P, a.D, a.M, a.means3D, reinterpret_cast<const float2 *>(a.scales), a.scale_modifier,
        reinterpret_cast<const float4 *>(a.rotations), a.opacities, a.shs, a.transMat_precomp, a.colors_precomp,
        a.viewmatrix, a.projmatrix, a.cam_pos, a.W, a.H, a.gx, a.gy, a.radii, a.rec, a.tiles_touched, a.depth_key,
        a.idx_in, a.clamped, a.prefiltered);
}

// =============================================================================================
// K0: markVisible (rasterizer_impl.cu:54-66)
// =============================================================================================
__global__ void __launch_bounds__(256)
mark_visible_kernel(const int P, const float *__restrict__ means3D, const float *__restrict__ viewmatrix,
                    unsigned char *__restrict__ present)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float3 p = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
    const float3 pv = xform_point(p, viewmatrix);
    present[idx] = pv.z > 0.2f ? 1 : 0;
}

void launch_mark_visible(int P, const float *means3D, const float *viewmatrix, unsigned char *present,
                         cudaStream_t stream)
{
    if (P == 0) return;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, viewmatrix, present);
}

// ---------------------------------------------------------------------------------------------
// Colour passes over one geometry state: the blend kernels read a splat's colour
// from words 15..17 of its packed record (common.cuh), so another pass with other precomputed colours
// only has to rewrite those 12 bytes; its colour gradient is moved out of (and cleared in) words 12..14
// of the gradient accumulator so that the geometry terms of all passes can keep summing there.
// ---------------------------------------------------------------------------------------------
__global__ void set_record_colors_kernel(const int P, const float *__restrict__ colors, float *__restrict__ rec)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    float *r = rec + (size_t)idx * REC_FLOATS;
    r[15] = colors[3 * idx];
    r[16] = colors[3 * idx + 1];
    r[17] = colors[3 * idx + 2];
}

__global__ void take_color_grad_kernel(const int P, float *__restrict__ gacc, float *__restrict__ dL_dcolor)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    float *g = gacc + (size_t)idx * GACC_FLOATS;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        if (dL_dcolor) dL_dcolor[3 * idx + c] = g[12 + c];
        g[12 + c] = 0.f;
    }
}

// class-probability pass: word 15 carries the class label (int bits; negative = contributes to no class)
__global__ void set_record_labels_kernel(const int P, const int *__restrict__ labels, float *__restrict__ rec)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    rec[(size_t)idx * REC_FLOATS + 15] = __int_as_float(labels[idx]);
}

void launch_set_record_labels(int P, const int *labels, float *rec, cudaStream_t stream)
{
    if (P == 0) return;
    set_record_labels_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, labels, rec);
}

void launch_set_record_colors(int P, const float *colors, float *rec, cudaStream_t stream)
{
    if (P == 0) return;
    set_record_colors_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, colors, rec);
}

void launch_take_color_grad(int P, float *gacc, float *dL_dcolor, cudaStream_t stream)
{
    if (P == 0) return;
    take_color_grad_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, gacc, dL_dcolor);
}

// =============================================================================================
// K8: backward preprocess.  One thread per Gaussian; consumes the gradient accumulator record
// written by the backward render kernel and writes EVERY output element (zeros for culled
// Gaussians), so callers never pre-zero 304 B/Gaussian as the reference binding does.
// =============================================================================================
__device__ __forceinline__ float4 quat_vjp(const float4 q, const float vR[3][3])
{  // auxiliary.h:238-282, gradient w.r.t. the normalised quaternion taken as free (quirk 1)
    const float s = rsqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    const float w = q.x * s, x = q.y * s, y = q.z * s, z = q.w * s;
    float4 v;
    v.x = 2.f * (x * (vR[1][2] - vR[2][1]) + y * (vR[2][0] - vR[0][2]) + z * (vR[0][1] - vR[1][0]));
    v.y = 2.f * (-2.f * x * (vR[1][1] + vR[2][2]) + y * (vR[0][1] + vR[1][0]) + z * (vR[0][2] + vR[2][0]) +
                 w * (vR[1][2] - vR[2][1]));
    v.z = 2.f * (x * (vR[0][1] + vR[1][0]) - 2.f * y * (vR[0][0] + vR[2][2]) + z * (vR[1][2] + vR[2][1]) +
                 w * (vR[2][0] - vR[0][2]));
    v.w = 2.f * (x * (vR[0][2] + vR[2][0]) + y * (vR[1][2] + vR[2][1]) - 2.f * z * (vR[0][0] + vR[1][1]) +
                 w * (vR[0][1] - vR[1][0]));
    return v;
}

// STAGED (SH rows 16-byte aligned, M <= 16): the three row-shaped streams of a Gaussian -- its 80-B gradient
// accumulator row, its SH row in and its SH-gradient row out (192 B each at degree 3: two thirds of the kernel's bytes)
// -- travel through shared memory with the TMA engine: every lane issues ONE bulk copy per row instead of 5 + 12 + 12
// 128-bit accesses whose 32 lanes touch 32 different cache lines each (ncu r02: the kernel was bound by L1 wavefronts,
// 23 % issue-active at 56 % of the HBM peak).  Rows are padded by 16 B in shared memory so that the per-lane 128-bit
// accesses of a quarter warp fall into distinct banks.  The arithmetic is untouched.
template <bool STAGED>
__global__ void __launch_bounds__(256, STAGED ? 3 : 1)
preprocess_bwd_kernel(const int P, const int D, const int M, const float *__restrict__ means3D,
                      const int *__restrict__ radii, const float *__restrict__ shs,
                      const uint8_t *__restrict__ clamped, const float2 *__restrict__ scales,
                      const float4 *__restrict__ rotations, const float *__restrict__ transMat_precomp,
                      const float *__restrict__ viewmatrix, const float *__restrict__ projmatrix,
                      const float focal_x, const float focal_y, const float tan_fovx, const float tan_fovy,
                      const float *__restrict__ cam_pos, const float *__restrict__ rec,
                      const float *__restrict__ gacc, const uint32_t *__restrict__ gacc_slot, float *__restrict__ dL_dmean2D,
                      float *__restrict__ dL_dnormal,
                      float *__restrict__ dL_dopacity, float *__restrict__ dL_dcolor, float *__restrict__ dL_dmean3D,
                      float *__restrict__ dL_dtransMat, float *__restrict__ dL_dsh, float2 *__restrict__ dL_dscale,
                      float4 *__restrict__ dL_drot)
{
    extern __shared__ __align__(128) unsigned char k8_smem[];
    const int idx_raw = blockIdx.x * blockDim.x + threadIdx.x;
    if (!STAGED && idx_raw >= P) return;
    const bool in_range = idx_raw < P;
    const int idx = in_range ? idx_raw : P - 1;   // STAGED: lanes past the end keep the warp's barrier protocol, store nothing
    float *sh_row = nullptr, *g_row = nullptr;
    // STAGED: the Gaussian's small parameters are fetched while the bulk copies are in flight (one exposed memory
    // latency instead of a chain of them: the kernel is latency-bound at 2-3 CTAs per SM)
    float3 pf_p = make_float3(0.f, 0.f, 0.f);
    float2 pf_sc = make_float2(0.f, 0.f);
    float4 pf_q = make_float4(0.f, 0.f, 0.f, 1.f);
    float pf_depth = 0.f;
    uint32_t pf_cb = 0;
    if constexpr (STAGED) {
        const int sh_stride = 3 * M + 4;
        float *sh_stage = reinterpret_cast<float *>(k8_smem);
        float *g_stage = sh_stage + 256 * sh_stride;
        uint64_t *bar = reinterpret_cast<uint64_t *>(g_stage + 256 * GACC_FLOATS) + (threadIdx.x >> 5);
        sh_row = sh_stage + threadIdx.x * sh_stride;
        g_row = g_stage + threadIdx.x * GACC_FLOATS;
        if ((threadIdx.x & 31) == 0) {
            mbar_init(bar, 32);
            mbar_fence_init();
        }
        __syncwarp();
        if (in_range && radii[idx] > 0) {
            const size_t row = gacc_slot ? (size_t)gacc_slot[idx] : (size_t)idx;
            bulk_g2s(g_row, gacc + row * GACC_FLOATS, GACC_FLOATS * 4, bar);
            bulk_g2s(sh_row, shs + (size_t)idx * M * 3, (uint32_t)(12 * M), bar);
            mbar_arrive_expect_tx(bar, (uint32_t)(GACC_FLOATS * 4 + 12 * M));
            pf_p = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
            if (scales != nullptr) {
                pf_sc = scales[idx];
                pf_q = rotations[idx];
            }
            pf_depth = rec[(size_t)idx * REC_FLOATS + 8];
            pf_cb = clamped[idx];
        } else {
            mbar_arrive_plain(bar);
        }
        mbar_wait(bar, 0);
    }

    float o_m2x = 0.f, o_m2y = 0.f, o_op = 0.f;
    float o_col[3] = {0.f, 0.f, 0.f}, o_nrm[3] = {0.f, 0.f, 0.f}, o_m3[3] = {0.f, 0.f, 0.f};
    float o_T[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float2 o_sc = make_float2(0.f, 0.f);
    float4 o_rot = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool have_sh = (shs != nullptr) && M > 0;
    const bool visible = radii[idx] > 0;
    float dRGB[3] = {0.f, 0.f, 0.f};
    float dirx = 0.f, diry = 0.f, dirz = 0.f;

    if (visible) {
        const size_t row = gacc_slot ? (size_t)gacc_slot[idx] : (size_t)idx;  // compact rows in the sharded path
        const float4 *g4 = reinterpret_cast<const float4 *>(STAGED ? g_row : gacc + row * GACC_FLOATS);
        const float4 a0 = g4[0], a1 = g4[1], a2 = g4[2], a3 = g4[3], a4 = g4[4];
        float dT[3][3] = {{a0.x, a0.y, a0.z}, {a0.w, a1.x, a1.y}, {a1.z, a1.w, a2.x}};
        float dm2x = a2.y, dm2y = a2.z;
        o_op = a2.w;
        o_col[0] = a3.x; o_col[1] = a3.y; o_col[2] = a3.z;
        o_nrm[0] = a3.w; o_nrm[1] = a4.x; o_nrm[2] = a4.y;
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int r = 0; r < 3; r++) o_T[3 * j + r] = dT[j][r];

        const int Wb = int(focal_x * tan_fovx * 2);  // backward.cu:613-614 (fp32 round trip, quirk 3)
        const int Hb = int(focal_y * tan_fovy * 2);
        const bool precomp = (scales == nullptr);     // backward.cu:615
        const float3 p_orig = STAGED ? pf_p : make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);

        float Tm[3][3], Pm[3][4], R[3][3];
        float3 normal = make_float3(0.f, 0.f, 0.f);
        float2 sc = make_float2(0.f, 0.f);
        float4 q = make_float4(0.f, 0.f, 0.f, 1.f);
        if (precomp) {
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int r = 0; r < 3; r++) Tm[j][r] = transMat_precomp[9 * idx + 3 * j + r];
        } else {
            float L[3][3];
            sc = STAGED ? pf_sc : scales[idx];
            q = STAGED ? pf_q : rotations[idx];
            quat_columns(q, R);
            scaled_frame(R, 1.0f * sc.x, 1.0f * sc.y, L);  // scale_modifier ignored (quirk 2)
            const float A[4][3] = {{L[0][0], L[1][0], p_orig.x}, {L[0][1], L[1][1], p_orig.y},
                                   {L[0][2], L[1][2], p_orig.z}, {0.f, 0.f, 1.f}};
            const float N[3][4] = {{float(Wb) / 2.0f, 0.f, 0.f, float(Wb - 1) / 2.0f},
                                   {0.f, float(Hb) / 2.0f, 0.f, float(Hb - 1) / 2.0f},
                                   {0.f, 0.f, 0.f, 1.f}};
            // P = world2ndc * ndc2pix, world2ndc[c][r] = projmatrix[c + 4r]
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int r = 0; r < 4; r++)
                    Pm[j][r] = projmatrix[0 + 4 * r] * N[j][0] + projmatrix[1 + 4 * r] * N[j][1] +
                               projmatrix[2 + 4 * r] * N[j][2] + projmatrix[3 + 4 * r] * N[j][3];
            // T = transpose(M) * P
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int r = 0; r < 3; r++)
                    Tm[j][r] = A[0][r] * Pm[j][0] + A[1][r] * Pm[j][1] + A[2][r] * Pm[j][2] + A[3][r] * Pm[j][3];
            normal = xform_vec(make_float3(L[2][0], L[2][1], L[2][2]), viewmatrix);
        }

        bool early = false;
        if (dm2x != 0 || dm2y != 0) {  // backward.cu:519-549: fold dL/dmean2D through the centre formula
            const float tv[3] = {9.0f, 9.0f, -1.0f};
            const float d = dot3(tv[0], tv[1], tv[2], Tm[2][0] * Tm[2][0], Tm[2][1] * Tm[2][1], Tm[2][2] * Tm[2][2]);
            const float invd = 1.0f / d;
            float f[3], dT3[3], df[3];
#pragma unroll
            for (int r = 0; r < 3; r++) f[r] = tv[r] * invd;
#pragma unroll
            for (int r = 0; r < 3; r++) {
                dT[0][r] += dm2x * f[r] * Tm[2][r];
                dT[1][r] += dm2y * f[r] * Tm[2][r];
                dT3[r] = dm2x * f[r] * Tm[0][r] + dm2y * f[r] * Tm[1][r];
                df[r] = dm2x * Tm[0][r] * Tm[2][r] + dm2y * Tm[1][r] * Tm[2][r];
            }
            const float dL_dd = dot3(df[0], df[1], df[2], f[0], f[1], f[2]) * (-1.0 / d);
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const float dd_dT3 = tv[r] * Tm[2][r] * 2.0f;
                dT3[r] += dL_dd * dd_dT3;
                dT[2][r] += dT3[r];
            }
            if (precomp) {
#pragma unroll
                for (int j = 0; j < 3; j++)
#pragma unroll
                    for (int r = 0; r < 3; r++) o_T[3 * j + r] = dT[j][r];
                early = true;
            }
        }
        if (!precomp && !early) {
            // dL_dM = P * transpose(dL_dT)
            float dM[3][4];
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int r = 0; r < 4; r++) dM[j][r] = Pm[0][r] * dT[0][j] + Pm[1][r] * dT[1][j] + Pm[2][r] * dT[2][j];
            float3 dtn = xform_vec_transposed(make_float3(o_nrm[0], o_nrm[1], o_nrm[2]), viewmatrix);
            const float3 p_view = xform_point(p_orig, viewmatrix);
            const float cosv = -(p_view.x * normal.x + p_view.y * normal.y + p_view.z * normal.z);
            const float mult = cosv > 0 ? 1.f : -1.f;
            dtn = make_float3(mult * dtn.x, mult * dtn.y, mult * dtn.z);
            const float dRS[3][3] = {{dM[0][0], dM[0][1], dM[0][2]}, {dM[1][0], dM[1][1], dM[1][2]}, {dtn.x, dtn.y, dtn.z}};
            float dR[3][3];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                dR[0][r] = dRS[0][r] * sc.x;
                dR[1][r] = dRS[1][r] * sc.y;
                dR[2][r] = dRS[2][r];
            }
            o_rot = quat_vjp(q, dR);
            o_sc.x = dot3(dRS[0][0], dRS[0][1], dRS[0][2], R[0][0], R[0][1], R[0][2]);
            o_sc.y = dot3(dRS[1][0], dRS[1][1], dRS[1][2], R[1][0], R[1][1], R[1][2]);
            o_m3[0] = dM[2][0]; o_m3[1] = dM[2][1]; o_m3[2] = dM[2][2];
        }

        if (have_sh) {
            const float3 cam = make_float3(cam_pos[0], cam_pos[1], cam_pos[2]);
            dirx = p_orig.x - cam.x; diry = p_orig.y - cam.y; dirz = p_orig.z - cam.z;
            const uint32_t cb = STAGED ? pf_cb : (uint32_t)clamped[idx];
#pragma unroll
            for (int c = 0; c < 3; c++) dRGB[c] = o_col[c] * (((cb >> c) & 1u) ? 0.f : 1.f);
        }

        // densification proxy (backward.cu:631-635, quirk 4): uses forward's Tw.z and the
        // (possibly folded, precomp path only) dL_dtransMat
        const float depth = STAGED ? pf_depth : rec[(size_t)idx * REC_FLOATS + 8];
        o_m2x = o_T[2] * depth * 0.5 * float(Wb);
        o_m2y = o_T[5] * depth * 0.5 * float(Hb);
    }

    // ---- SH VJP (backward.cu:20-139); also writes the zero rows of culled Gaussians ----
    // Streaming formulation: the row is visited once in memory order, four floats (one 128-bit access
    // when rows are 16-byte aligned) at a time; element i = 3k + c needs only the k-th basis value and
    // its gradient, which are compile-time selected polynomials of the view direction.  No 48-float
    // staging arrays -> ~half the registers of the direct transcription.
    if (have_sh) {
        float *dsh = dL_dsh + (size_t)idx * M * 3;
        const float *sh = shs + (size_t)idx * M * 3;
        const bool vec_ok = ((M & 3) == 0) && ((reinterpret_cast<uintptr_t>(dL_dsh) & 15) == 0) &&
                            ((reinterpret_cast<uintptr_t>(shs) & 15) == 0);
        const int nflt = 3 * M;
        if (!visible) {
            if (STAGED) {
                float4 *d4 = reinterpret_cast<float4 *>(sh_row);
                for (int i = 0; i < nflt / 4; i++) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            } else if (vec_ok) {
                float4 *d4 = reinterpret_cast<float4 *>(dsh);
                for (int i = 0; i < nflt / 4; i++) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                for (int i = 0; i < nflt; i++) dsh[i] = 0.f;
            }
        } else {
            const float len = sqrtf(dot3(dirx, diry, dirz, dirx, diry, dirz));
            const float x = dirx / len, y = diry / len, z = dirz / len;
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            const int nact = 3 * (D + 1) * (D + 1);  // floats of the active coefficients
            float gx[3] = {0.f, 0.f, 0.f}, gy[3] = {0.f, 0.f, 0.f}, gz[3] = {0.f, 0.f, 0.f};  // dRGB/d(dir)
#pragma unroll
            for (int j = 0; j < 12; j++) {
                if (4 * j < nflt) {
                    float v[4] = {0.f, 0.f, 0.f, 0.f}, o[4];
                    if (4 * j < nact) {
                        if (STAGED) {
                            const float4 t = reinterpret_cast<const float4 *>(sh_row)[j];
                            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                        } else if (vec_ok) {
                            const float4 t = __ldg(reinterpret_cast<const float4 *>(sh) + j);
                            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; e++)
                                if (4 * j + e < nact) v[e] = __ldg(sh + 4 * j + e);
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        constexpr int dummy = 0; (void)dummy;
                        const int i = 4 * j + e, k = i / 3, c = i % 3;
                        float bv = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
                        switch (k) {  // folded at compile time (j, e are unrolled)
                            case 0: bv = kSH_C0; break;
                            case 1: bv = -kSH_C1 * y; by = -kSH_C1; break;
                            case 2: bv = kSH_C1 * z; bz = kSH_C1; break;
                            case 3: bv = -kSH_C1 * x; bx = -kSH_C1; break;
                            case 4: bv = kSH_C2[0] * xy; bx = kSH_C2[0] * y; by = kSH_C2[0] * x; break;
                            case 5: bv = kSH_C2[1] * yz; by = kSH_C2[1] * z; bz = kSH_C2[1] * y; break;
                            case 6: bv = kSH_C2[2] * (2.f * zz - xx - yy); bx = kSH_C2[2] * -2.f * x; by = kSH_C2[2] * -2.f * y;
                                    bz = kSH_C2[2] * 4.f * z; break;
                            case 7: bv = kSH_C2[3] * xz; bx = kSH_C2[3] * z; bz = kSH_C2[3] * x; break;
                            case 8: bv = kSH_C2[4] * (xx - yy); bx = kSH_C2[4] * 2.f * x; by = kSH_C2[4] * -2.f * y; break;
                            case 9: bv = kSH_C3[0] * y * (3.f * xx - yy); bx = kSH_C3[0] * 6.f * xy; by = kSH_C3[0] * 3.f * (xx - yy); break;
                            case 10: bv = kSH_C3[1] * xy * z; bx = kSH_C3[1] * yz; by = kSH_C3[1] * xz; bz = kSH_C3[1] * xy; break;
                            case 11: bv = kSH_C3[2] * y * (4.f * zz - xx - yy); bx = kSH_C3[2] * -2.f * xy;
                                     by = kSH_C3[2] * (-3.f * yy + 4.f * zz - xx); bz = kSH_C3[2] * 8.f * yz; break;
                            case 12: bv = kSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy); bx = kSH_C3[3] * -6.f * xz;
                                     by = kSH_C3[3] * -6.f * yz; bz = kSH_C3[3] * 3.f * (2.f * zz - xx - yy); break;
                            case 13: bv = kSH_C3[4] * x * (4.f * zz - xx - yy); bx = kSH_C3[4] * (-3.f * xx + 4.f * zz - yy);
                                     by = kSH_C3[4] * -2.f * xy; bz = kSH_C3[4] * 8.f * xz; break;
                            case 14: bv = kSH_C3[5] * z * (xx - yy); bx = kSH_C3[5] * 2.f * xz; by = kSH_C3[5] * -2.f * yz;
                                     bz = kSH_C3[5] * (xx - yy); break;
                            case 15: bv = kSH_C3[6] * x * (xx - 3.f * yy); bx = kSH_C3[6] * 3.f * (xx - yy);
                                     by = kSH_C3[6] * -6.f * xy; break;
                            default: break;
                        }
                        const bool act = i < nact;  // coefficients above the active degree: exact zeros
                        o[e] = act ? bv * dRGB[c] : 0.f;
                        if (act) {
                            gx[c] += bx * v[e];
                            gy[c] += by * v[e];
                            gz[c] += bz * v[e];
                        }
                    }
                    if (STAGED) {
                        reinterpret_cast<float4 *>(sh_row)[j] = make_float4(o[0], o[1], o[2], o[3]);   // in place: row j was read above
                    } else if (vec_ok) {
                        reinterpret_cast<float4 *>(dsh)[j] = make_float4(o[0], o[1], o[2], o[3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; e++)
                            if (4 * j + e < nflt) dsh[4 * j + e] = o[e];
                    }
                }
            }
            for (int i = 48; i < nflt; i++) dsh[i] = 0.f;  // M > 16 (never produced by the reference's callers)
            const float ddx = dot3(gx[0], gx[1], gx[2], dRGB[0], dRGB[1], dRGB[2]);
            const float ddy = dot3(gy[0], gy[1], gy[2], dRGB[0], dRGB[1], dRGB[2]);
            const float ddz = dot3(gz[0], gz[1], gz[2], dRGB[0], dRGB[1], dRGB[2]);
            // normalisation Jacobian (auxiliary.h:128-138); "+=" on top of the geometric term (quirk 6)
            const float sum2 = dirx * dirx + diry * diry + dirz * dirz;
            const float invsum32 = 1.0f / sqrt(sum2 * sum2 * sum2);
            o_m3[0] += ((+sum2 - dirx * dirx) * ddx - diry * dirx * ddy - dirz * dirx * ddz) * invsum32;
            o_m3[1] += (-dirx * diry * ddx + (sum2 - diry * diry) * ddy - dirz * diry * ddz) * invsum32;
            o_m3[2] += (-dirx * dirz * ddx - diry * dirz * ddy + (sum2 - dirz * dirz) * ddz) * invsum32;
        }
        if (STAGED && in_range) bulk_s2g(dsh, sh_row, (uint32_t)(12 * M));   // the whole gradient row, one request
    }
    if (STAGED && !in_range) {
        bulk_wait_read_all();
        return;
    }

    dL_dmean2D[3 * idx] = o_m2x;
    dL_dmean2D[3 * idx + 1] = o_m2y;
    dL_dmean2D[3 * idx + 2] = 0.f;
    dL_dopacity[idx] = o_op;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        dL_dcolor[3 * idx + c] = o_col[c];
        dL_dmean3D[3 * idx + c] = o_m3[c];
        if (dL_dnormal) dL_dnormal[3 * idx + c] = o_nrm[c];
    }
#pragma unroll
    for (int i = 0; i < 9; i++) dL_dtransMat[9 * idx + i] = o_T[i];
    dL_dscale[idx] = o_sc;
    dL_drot[idx] = o_rot;
    if (STAGED) bulk_wait_read_all();   // the staging row must outlive the bulk store's read of it
}

void launch_preprocess_bwd(const PreprocessBwdArgs &a, cudaStream_t stream)
{
    if (a.P == 0) return;
    const bool staged = a.shs != nullptr && a.dL_dsh != nullptr && a.M > 0 && a.M <= 16 && (a.M & 3) == 0 &&
                        (reinterpret_cast<uintptr_t>(a.shs) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.dL_dsh) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(a.gacc) & 15) == 0;
    if (staged) {
        const size_t smem = (size_t)256 * (3 * a.M + 4) * 4 + (size_t)256 * GACC_FLOATS * 4 + 8 * sizeof(uint64_t);
        cudaFuncSetAttribute(preprocess_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        preprocess_bwd_kernel<true><<<(a.P + 255) / 256, 256, smem, stream>>>(
            a.P, a.D, a.M, a.means3D, a.radii, a.shs, a.clamped, reinterpret_cast<const float2 *>(a.scales),
            reinterpret_cast<const float4 *>(a.rotations), a.transMat_precomp, a.viewmatrix, a.projmatrix, a.focal_x,
            a.focal_y, a.tan_fovx, a.tan_fovy, a.cam_pos, a.rec, a.gacc, a.gacc_slot, a.dL_dmean2D, a.dL_dnormal,
            a.dL_dopacity, a.dL_dcolor, a.dL_dmean3D, a.dL_dtransMat, a.dL_dsh, reinterpret_cast<float2 *>(a.dL_dscale),
            reinterpret_cast<float4 *>(a.dL_drot));
        return;
    }
    preprocess_bwd_kernel<false><<<(a.P + 255) / 256, 256, 0, stream>>>(
        a.P, a.D, a.M, a.means3D, a.radii, a.shs, a.clamped, reinterpret_cast<const float2 *>(a.scales),
        reinterpret_cast<const float4 *>(a.rotations), a.transMat_precomp, a.viewmatrix, a.projmatrix, a.focal_x,
        a.focal_y, a.tan_fovx, a.tan_fovy, a.cam_pos, a.rec, a.gacc, a.gacc_slot, a.dL_dmean2D, a.dL_dnormal, a.dL_dopacity,
        a.dL_dcolor, a.dL_dmean3D, a.dL_dtransMat, a.dL_dsh, reinterpret_cast<float2 *>(a.dL_dscale),
        reinterpret_cast<float4 *>(a.dL_drot));
}

}  // namespace surfel
