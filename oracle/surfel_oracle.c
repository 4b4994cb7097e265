/*
 * surfel_oracle.c -- CPU restatement of the reference 2D-Gaussian-surfel rasterizer.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (streetunveiler_b200/) never imports, links or calls anything in oracle/.
 *
 * Parity pin: the reference has no tests or golden vectors for this path
 * (SURVEY.md section 4 / 8c).  This restatement is pinned against outputs of the
 * UNMODIFIED reference CUDA extension (oracle/_ref, built by oracle/build_ref.py)
 * captured on a B200 and committed under tests/golden/ (see tests/golden/make_golden.py).
 *
 * All file:line citations are relative to
 *   /root/reference/submodules/diff-surfel-rasterization/   ("RAST/")
 *
 * Arithmetic is fp32 in the reference's operation order (without FMA contraction,
 * so agreement with the GPU is to rounding, not bit-exact).  Per-Gaussian gradient
 * scatter is accumulated in fp64 so the oracle is independent of atomic ordering.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16                       /* RAST/cuda_rasterizer/config.h:16-17 */
#define NEAR_N 0.2f                   /* RAST/cuda_rasterizer/auxiliary.h:37 */
#define FAR_N 100.0f                  /* auxiliary.h:38 */
#define FILTER_SIZE 0.707106f         /* auxiliary.h:39 */
#define FILTER_INV_SQUARE 2.0f        /* auxiliary.h:40 */

/* auxiliary.h:43-60 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

typedef struct {
    int P, W, H, tiles_x, tiles_y, D, M;
    int64_t R;
    /* GeometryState (rasterizer_impl.h:31-45) */
    float *depths;        /* [P]   */
    uint8_t *clamped;     /* [3P]  */
    float *means2D;       /* [2P]  */
    float *transMat;      /* [9P]  */
    float *normal_opacity;/* [4P]  */
    float *rgb;           /* [3P]  */
    uint32_t *tiles_touched; /* [P] */
    int *radii;           /* [P]   */
    /* BinningState (rasterizer_impl.h:57-66) */
    uint32_t *point_list; /* [R]   */
    uint64_t *point_keys; /* [R]   */
    /* ImageState (rasterizer_impl.h:47-55) */
    uint32_t *ranges;     /* [2*tiles] */
    float *final_T;       /* [3*HW] : T, M1, M2 */
    uint32_t *n_contrib;  /* [2*HW] : last, median */
} OracleState;

/* ---- float->int conversion with the GPU's saturating cvt.rzi semantics ---- */
static inline int f2i(float v)
{
    if (!(v == v)) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (-2147483647 - 1);
    return (int)v;
}
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* auxiliary.h:67-77 getRect */
static void get_rect(float px, float py, int max_radius, int gx, int gy, int rmin[2], int rmax[2])
{
    rmin[0] = imin(gx, imax(0, f2i((px - max_radius) / TILE)));
    rmin[1] = imin(gy, imax(0, f2i((py - max_radius) / TILE)));
    rmax[0] = imin(gx, imax(0, f2i((px + max_radius + TILE - 1) / TILE)));
    rmax[1] = imin(gy, imax(0, f2i((py + max_radius + TILE - 1) / TILE)));
}

/* auxiliary.h:213-235 quat_to_rotmat; R is column-major R[c][r] like glm */
static void quat_to_rotmat(const float q[4], float R[3][3])
{
    float s = 1.0f / sqrtf(q[3] * q[3] + q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    float w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
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

/* auxiliary.h:238-282 quat_to_rotmat_vjp (no normalisation Jacobian) */
static void quat_to_rotmat_vjp(const float q[4], float vR[3][3], float vq[4])
{
    float s = 1.0f / sqrtf(q[3] * q[3] + q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    float w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
    vq[0] = 2.f * (x * (vR[1][2] - vR[2][1]) + y * (vR[2][0] - vR[0][2]) + z * (vR[0][1] - vR[1][0]));
    vq[1] = 2.f * (-2.f * x * (vR[1][1] + vR[2][2]) + y * (vR[0][1] + vR[1][0]) +
                   z * (vR[0][2] + vR[2][0]) + w * (vR[1][2] - vR[2][1]));
    vq[2] = 2.f * (x * (vR[0][1] + vR[1][0]) - 2.f * y * (vR[0][0] + vR[2][2]) +
                   z * (vR[1][2] + vR[2][1]) + w * (vR[2][0] - vR[0][2]));
    vq[3] = 2.f * (x * (vR[0][2] + vR[2][0]) + y * (vR[1][2] + vR[2][1]) -
                   2.f * z * (vR[0][0] + vR[1][1]) + w * (vR[0][1] - vR[1][0]));
}

/* auxiliary.h:89-98,100-118 */
static void transform_point4x3(const float p[3], const float *m, float o[3])
{
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
static void transform_vec4x3(const float p[3], const float *m, float o[3])
{
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2];
}
static void transform_vec4x3_transpose(const float p[3], const float *m, float o[3])
{
    o[0] = m[0] * p[0] + m[1] * p[1] + m[2] * p[2];
    o[1] = m[4] * p[0] + m[5] * p[1] + m[6] * p[2];
    o[2] = m[8] * p[0] + m[9] * p[1] + m[10] * p[2];
}

/*
 * forward.cu:75-115 compute_transmat:  T = (transpose(splat2world) * world2ndc) * ndc2pix
 * with glm's column-major products (type_mat4x3.inl).  A = transpose(splat2world) has
 * columns A[k] = (L0[k], L1[k], p[k]) for k<3 and A[3] = (0,0,1); world2ndc[c][k] = proj[c+4k].
 * T[j][r], j = u/v/w row of the pixel-space homogeneous map, r = (tangent-u, tangent-v, centre).
 */
static void compute_transmat_fwd(const float p[3], const float sc[2], float mod, const float q[4],
                                 const float *proj, const float *view, int W, int H,
                                 float T[3][3], float normal[3])
{
    float R[3][3];
    quat_to_rotmat(q, R);
    float L0[3], L1[3], L2[3];
    float s0 = mod * sc[0], s1 = mod * sc[1];
    for (int r = 0; r < 3; r++) {
        L0[r] = R[0][r] * s0;
        L1[r] = R[1][r] * s1;
        L2[r] = R[2][r];
    }
    float A[4][3] = {{L0[0], L1[0], p[0]}, {L0[1], L1[1], p[1]}, {L0[2], L1[2], p[2]}, {0.f, 0.f, 1.f}};
    float B[4][3];
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 3; r++)
            B[c][r] = A[0][r] * proj[c] + A[1][r] * proj[c + 4] + A[2][r] * proj[c + 8] + A[3][r] * proj[c + 12];
    float hw = (float)W / 2.0f, hw1 = (float)(W - 1) / 2.0f;
    float hh = (float)H / 2.0f, hh1 = (float)(H - 1) / 2.0f;
    for (int r = 0; r < 3; r++) {
        T[0][r] = B[0][r] * hw + B[3][r] * hw1;
        T[1][r] = B[1][r] * hh + B[3][r] * hh1;
        T[2][r] = B[3][r];
    }
    transform_vec4x3(L2, view, normal);
}

/* forward.cu:119-145 compute_aabb (cutoff = 3) */
static int compute_aabb(float T[3][3], float cutoff, float pt[2], float ext[2])
{
    float t[3] = {cutoff * cutoff, cutoff * cutoff, -1.0f};
    float d = t[0] * (T[2][0] * T[2][0]) + t[1] * (T[2][1] * T[2][1]) + t[2] * (T[2][2] * T[2][2]);
    if (d == 0.0f) return 0;
    float inv = 1 / d;
    float f[3] = {inv * t[0], inv * t[1], inv * t[2]};
    float px = f[0] * (T[0][0] * T[2][0]) + f[1] * (T[0][1] * T[2][1]) + f[2] * (T[0][2] * T[2][2]);
    float py = f[0] * (T[1][0] * T[2][0]) + f[1] * (T[1][1] * T[2][1]) + f[2] * (T[1][2] * T[2][2]);
    float qx = f[0] * (T[0][0] * T[0][0]) + f[1] * (T[0][1] * T[0][1]) + f[2] * (T[0][2] * T[0][2]);
    float qy = f[0] * (T[1][0] * T[1][0]) + f[1] * (T[1][1] * T[1][1]) + f[2] * (T[1][2] * T[1][2]);
    float h0x = px * px - qx, h0y = py * py - qy;
    ext[0] = sqrtf(fmaxf(1e-4f, h0x));
    ext[1] = sqrtf(fmaxf(1e-4f, h0y));
    pt[0] = px;
    pt[1] = py;
    return 1;
}

/* forward.cu:20-71 computeColorFromSH */
static void sh_to_rgb(int deg, int max_coeffs, const float p[3], const float cam[3], const float *sh_all,
                      int idx, uint8_t *clamped, float out[3])
{
    float dx = p[0] - cam[0], dy = p[1] - cam[1], dz = p[2] - cam[2];
    float len = sqrtf(dx * dx + dy * dy + dz * dz);
    float x = dx / len, y = dy / len, z = dz / len;
    const float *sh = sh_all + (size_t)idx * max_coeffs * 3;
    for (int c = 0; c < 3; c++) {
#define S(k) sh[3 * (k) + c]
        float res = SH_C0 * S(0);
        if (deg > 0) {
            res = res - SH_C1 * y * S(1) + SH_C1 * z * S(2) - SH_C1 * x * S(3);
            if (deg > 1) {
                float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                res = res + SH_C2[0] * xy * S(4) + SH_C2[1] * yz * S(5) + SH_C2[2] * (2.0f * zz - xx - yy) * S(6) +
                      SH_C2[3] * xz * S(7) + SH_C2[4] * (xx - yy) * S(8);
                if (deg > 2) {
                    res = res + SH_C3[0] * y * (3.0f * xx - yy) * S(9) + SH_C3[1] * xy * z * S(10) +
                          SH_C3[2] * y * (4.0f * zz - xx - yy) * S(11) +
                          SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * S(12) +
                          SH_C3[4] * x * (4.0f * zz - xx - yy) * S(13) + SH_C3[5] * z * (xx - yy) * S(14) +
                          SH_C3[6] * x * (xx - 3.0f * yy) * S(15);
                }
            }
        }
#undef S
        res += 0.5f;
        clamped[3 * idx + c] = (res < 0);
        out[c] = fmaxf(res, 0.0f);
    }
}

typedef struct { uint64_t key; uint32_t val; } KV;
/* stable bottom-up merge sort: same order as cub's stable radix sort on the same keys
 * (rasterizer_impl.cu:304-309) */
static void stable_sort_kv(KV *a, int64_t n)
{
    if (n < 2) return;
    KV *tmp = (KV *)malloc(sizeof(KV) * n);
    KV *src = a, *dst = tmp;
    for (int64_t w = 1; w < n; w *= 2) {
        for (int64_t i = 0; i < n; i += 2 * w) {
            int64_t m = i + w < n ? i + w : n, r = i + 2 * w < n ? i + 2 * w : n;
            int64_t a0 = i, b0 = m, k = i;
            while (a0 < m && b0 < r) dst[k++] = (src[b0].key < src[a0].key) ? src[b0++] : src[a0++];
            while (a0 < m) dst[k++] = src[a0++];
            while (b0 < r) dst[k++] = src[b0++];
        }
        KV *s = src; src = dst; dst = s;
    }
    if (src != a) memcpy(a, src, sizeof(KV) * n);
    free(tmp);
}

void oracle_free(OracleState *s)
{
    if (!s) return;
    free(s->depths); free(s->clamped); free(s->means2D); free(s->transMat); free(s->normal_opacity);
    free(s->rgb); free(s->tiles_touched); free(s->radii); free(s->point_list); free(s->point_keys);
    free(s->ranges); free(s->final_T); free(s->n_contrib);
    free(s);
}

/*
 * Forward: rasterizer_impl.cu:198-342 (orchestration), forward.cu:148-251 (preprocess),
 * rasterizer_impl.cu:70-138 (keys, ranges), forward.cu:256-448 (render).
 * out_color[3,H,W], out_others[7,H,W], radii[P] are written; returns the saved state.
 * Null pointers select the alternative input paths exactly as the reference does
 * (shs xor colors_precomp; scales+rotations xor transMat_precomp).
 */
OracleState *oracle_forward(int P, int D, int M, const float *bg, int W, int H, const float *means3D,
                            const float *shs, const float *colors_precomp, const float *opacities,
                            const float *scales, float scale_modifier, const float *rotations,
                            const float *transMat_precomp, const float *view, const float *proj,
                            const float *campos, float tan_fovx, float tan_fovy, float *out_color,
                            float *out_others, int *radii_out, int64_t *num_rendered)
{
    (void)tan_fovx; (void)tan_fovy;
    OracleState *s = (OracleState *)calloc(1, sizeof(OracleState));
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const size_t HW = (size_t)W * H;
    s->P = P; s->W = W; s->H = H; s->tiles_x = gx; s->tiles_y = gy; s->D = D; s->M = M;
    s->depths = (float *)calloc(P + 1, 4);
    s->clamped = (uint8_t *)calloc(3 * (size_t)P + 1, 1);
    s->means2D = (float *)calloc(2 * (size_t)P + 1, 4);
    s->transMat = (float *)calloc(9 * (size_t)P + 1, 4);
    s->normal_opacity = (float *)calloc(4 * (size_t)P + 1, 4);
    s->rgb = (float *)calloc(3 * (size_t)P + 1, 4);
    s->tiles_touched = (uint32_t *)calloc(P + 1, 4);
    s->radii = (int *)calloc(P + 1, 4);
    s->ranges = (uint32_t *)calloc(2 * (size_t)gx * gy, 4);
    s->final_T = (float *)calloc(3 * HW, 4);
    s->n_contrib = (uint32_t *)calloc(2 * HW, 4);

    /* ---------------- preprocess, forward.cu:148-251 ---------------- */
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; idx++) {
        const float *p = means3D + 3 * (size_t)idx;
        float pv[3];
        transform_point4x3(p, view, pv);
        if (pv[2] <= 0.2f) continue; /* auxiliary.h:199 */
        float T[3][3], normal[3];
        if (transMat_precomp == NULL) {
            compute_transmat_fwd(p, scales + 2 * (size_t)idx, scale_modifier, rotations + 4 * (size_t)idx,
                                 proj, view, W, H, T, normal);
            for (int j = 0; j < 3; j++)
                for (int r = 0; r < 3; r++) s->transMat[9 * (size_t)idx + 3 * j + r] = T[j][r];
        } else {
            for (int j = 0; j < 3; j++)
                for (int r = 0; r < 3; r++) T[j][r] = transMat_precomp[9 * (size_t)idx + 3 * j + r];
            normal[0] = 0.f; normal[1] = 0.f; normal[2] = 1.f;
        }
        /* forward.cu:209-214 DUAL_VISIABLE */
        float cosv = -((pv[0] * normal[0] + pv[1] * normal[1]) + pv[2] * normal[2]);
        if (cosv == 0) continue;
        float mult = cosv > 0 ? 1.f : -1.f;
        normal[0] *= mult; normal[1] *= mult; normal[2] *= mult;
        float pt[2], ext[2];
        if (!compute_aabb(T, 3.0f, pt, ext)) continue;
        float radius = ceilf(fmaxf(fmaxf(ext[0], ext[1]), 3.0f * FILTER_SIZE));
        int rmin[2], rmax[2];
        get_rect(pt[0], pt[1], f2i(radius), gx, gy, rmin, rmax);
        if ((rmax[0] - rmin[0]) * (rmax[1] - rmin[1]) == 0) continue;
        if (colors_precomp == NULL) {
            float c[3];
            sh_to_rgb(D, M, p, campos, shs, idx, s->clamped, c);
            s->rgb[3 * (size_t)idx + 0] = c[0];
            s->rgb[3 * (size_t)idx + 1] = c[1];
            s->rgb[3 * (size_t)idx + 2] = c[2];
        }
        s->depths[idx] = pv[2];
        s->radii[idx] = f2i(radius);
        s->means2D[2 * (size_t)idx] = pt[0];
        s->means2D[2 * (size_t)idx + 1] = pt[1];
        s->normal_opacity[4 * (size_t)idx + 0] = normal[0];
        s->normal_opacity[4 * (size_t)idx + 1] = normal[1];
        s->normal_opacity[4 * (size_t)idx + 2] = normal[2];
        s->normal_opacity[4 * (size_t)idx + 3] = opacities[idx];
        s->tiles_touched[idx] = (uint32_t)((rmax[1] - rmin[1]) * (rmax[0] - rmin[0]));
    }
    if (radii_out) memcpy(radii_out, s->radii, sizeof(int) * (size_t)P);

    /* ---------------- binning, rasterizer_impl.cu:278-318 ---------------- */
    int64_t R = 0;
    for (int i = 0; i < P; i++) R += s->tiles_touched[i];
    s->R = R;
    *num_rendered = R;
    KV *kv = (KV *)malloc(sizeof(KV) * (size_t)(R > 0 ? R : 1));
    {
        int64_t off = 0;
        for (int idx = 0; idx < P; idx++) { /* duplicateWithKeys, rasterizer_impl.cu:70-111 */
            if (s->radii[idx] <= 0) continue;
            int rmin[2], rmax[2];
            get_rect(s->means2D[2 * (size_t)idx], s->means2D[2 * (size_t)idx + 1], s->radii[idx], gx, gy, rmin, rmax);
            uint32_t dbits;
            memcpy(&dbits, &s->depths[idx], 4);
            for (int y = rmin[1]; y < rmax[1]; y++)
                for (int x = rmin[0]; x < rmax[0]; x++) {
                    kv[off].key = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
                    kv[off].val = (uint32_t)idx;
                    off++;
                }
        }
    }
    stable_sort_kv(kv, R);
    s->point_list = (uint32_t *)malloc(4 * (size_t)(R > 0 ? R : 1));
    s->point_keys = (uint64_t *)malloc(8 * (size_t)(R > 0 ? R : 1));
    for (int64_t i = 0; i < R; i++) { s->point_list[i] = kv[i].val; s->point_keys[i] = kv[i].key; }
    free(kv);
    for (int64_t i = 0; i < R; i++) { /* identifyTileRanges, rasterizer_impl.cu:116-138 */
        uint32_t cur = (uint32_t)(s->point_keys[i] >> 32);
        if (i == 0) s->ranges[2 * cur] = 0;
        else {
            uint32_t prev = (uint32_t)(s->point_keys[i - 1] >> 32);
            if (cur != prev) { s->ranges[2 * prev + 1] = (uint32_t)i; s->ranges[2 * cur] = (uint32_t)i; }
        }
        if (i == R - 1) s->ranges[2 * cur + 1] = (uint32_t)R;
    }

    /* ---------------- render, forward.cu:256-448 ---------------- */
    const float *feat = colors_precomp ? colors_precomp : s->rgb;
    const float *tm = transMat_precomp ? transMat_precomp : s->transMat;
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < gx * gy; tile++) {
        int tx = tile % gx, ty = tile / gx;
        uint32_t r0 = s->ranges[2 * tile], r1 = s->ranges[2 * tile + 1];
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                int pxi = tx * TILE + lx, pyi = ty * TILE + ly;
                if (pxi >= W || pyi >= H) continue;
                size_t pix_id = (size_t)W * pyi + pxi;
                float pxf = (float)pxi, pyf = (float)pyi;
                float T = 1.0f, C[3] = {0, 0, 0}, N[3] = {0, 0, 0}, Dp = 0, M1 = 0, M2 = 0, dist = 0, med = 0;
                float median_contributor = -1; /* forward.cu:317, stored into a uint32 (quirk 8) */
                uint32_t contributor = 0, last_contributor = 0;
                for (uint32_t i = r0; i < r1; i++) {
                    contributor++;
                    uint32_t g = s->point_list[i];
                    const float *Tu = tm + 9 * (size_t)g, *Tv = Tu + 3, *Tw = Tu + 6;
                    float k[3] = {pxf * Tw[0] - Tu[0], pxf * Tw[1] - Tu[1], pxf * Tw[2] - Tu[2]};
                    float l[3] = {pyf * Tw[0] - Tv[0], pyf * Tw[1] - Tv[1], pyf * Tw[2] - Tv[2]};
                    float p[3] = {k[1] * l[2] - k[2] * l[1], k[2] * l[0] - k[0] * l[2], k[0] * l[1] - k[1] * l[0]};
                    if (p[2] == 0.0f) continue;
                    float sx = p[0] / p[2], sy = p[1] / p[2];
                    float rho3d = sx * sx + sy * sy;
                    float dx = s->means2D[2 * (size_t)g] - pxf, dy = s->means2D[2 * (size_t)g + 1] - pyf;
                    float rho2d = FILTER_INV_SQUARE * (dx * dx + dy * dy);
                    float rho = fminf(rho3d, rho2d);
                    float depth = (sx * Tw[0] + sy * Tw[1]) + Tw[2];
                    if (depth < NEAR_N) continue;
                    const float *no = s->normal_opacity + 4 * (size_t)g;
                    float power = -0.5f * rho;
                    if (power > 0.0f) continue;
                    float alpha = fminf(0.99f, no[3] * expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    float test_T = T * (1 - alpha);
                    if (test_T < 0.0001f) break; /* done=true; nothing later is visited */
                    float w = alpha * T;
                    float A = 1 - T;
                    float m = FAR_N / (FAR_N - NEAR_N) * (1 - NEAR_N / depth);
                    dist += (m * m * A + M2 - 2 * m * M1) * w;
                    Dp += depth * w;
                    M1 += m * w;
                    M2 += m * m * w;
                    if (T > 0.5f) { med = depth; median_contributor = (float)contributor; }
                    for (int ch = 0; ch < 3; ch++) N[ch] += no[ch] * w;
                    for (int ch = 0; ch < 3; ch++) C[ch] += feat[3 * (size_t)g + ch] * w;
                    T = test_T;
                    last_contributor = contributor;
                }
                s->final_T[pix_id] = T;
                s->final_T[pix_id + HW] = M1;
                s->final_T[pix_id + 2 * HW] = M2;
                s->n_contrib[pix_id] = last_contributor;
                /* float -1 -> uint32: cvt.rzi.u32.f32 saturates to 0 on the GPU */
                s->n_contrib[pix_id + HW] = median_contributor < 0 ? 0u : (uint32_t)median_contributor;
                for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pix_id] = C[ch] + T * bg[ch];
                out_others[0 * HW + pix_id] = Dp;
                out_others[1 * HW + pix_id] = 1 - T;
                out_others[2 * HW + pix_id] = N[0];
                out_others[3 * HW + pix_id] = N[1];
                out_others[4 * HW + pix_id] = N[2];
                out_others[5 * HW + pix_id] = med;
                out_others[6 * HW + pix_id] = dist;
            }
    }
    return s;
}

/* state accessors for the debug / exact-equality tests */
int64_t oracle_num_rendered(const OracleState *s) { return s->R; }
const uint32_t *oracle_point_list(const OracleState *s) { return s->point_list; }
const uint32_t *oracle_ranges(const OracleState *s) { return s->ranges; }
const float *oracle_transmat(const OracleState *s) { return s->transMat; }
const float *oracle_means2d(const OracleState *s) { return s->means2D; }
const float *oracle_depths(const OracleState *s) { return s->depths; }
const float *oracle_rgb(const OracleState *s) { return s->rgb; }
const float *oracle_normal_opacity(const OracleState *s) { return s->normal_opacity; }
const uint32_t *oracle_tiles_touched(const OracleState *s) { return s->tiles_touched; }
const float *oracle_final_T(const OracleState *s) { return s->final_T; }
const uint32_t *oracle_n_contrib(const OracleState *s) { return s->n_contrib; }

static inline void atomic_add_d(double *p, double v)
{
#pragma omp atomic
    *p += v;
}

/*
 * Backward: rasterizer_impl.cu:346-448; render backward.cu:143-446; per-Gaussian
 * backward.cu:449-636 (+ SH VJP backward.cu:20-139).  Outputs (all fp32, caller-allocated, any
 * content -- they are overwritten): dL_dmean2D[P,3], dL_dcolors[P,3], dL_dopacity[P],
 * dL_dmean3D[P,3], dL_dtransMat[P,9], dL_dsh[P,M,3], dL_dscale[P,2], dL_drot[P,4],
 * and (internal in the reference, exposed here for tests) dL_dnormal[P,3].
 */
void oracle_backward(const OracleState *s, const float *bg, const float *means3D, const float *shs,
                     const float *colors_precomp, const float *scales, float scale_modifier,
                     const float *rotations, const float *transMat_precomp, const float *view,
                     const float *proj, const float *campos, float tan_fovx, float tan_fovy,
                     const float *dL_dpix, const float *dL_dothers, float *dL_dmean2D, float *dL_dcolors,
                     float *dL_dopacity, float *dL_dmean3D, float *dL_dtransMat, float *dL_dsh,
                     float *dL_dscale, float *dL_drot, float *dL_dnormal_out)
{
    (void)scale_modifier;
    const int P = s->P, W = s->W, H = s->H, gx = s->tiles_x, gy = s->tiles_y, D = s->D, M = s->M;
    const size_t HW = (size_t)W * H;
    const float focal_y = H / (2.0f * tan_fovy), focal_x = W / (2.0f * tan_fovx); /* rasterizer_impl.cu:388-389 */
    const float *feat = colors_precomp ? colors_precomp : s->rgb;
    const float *tm = transMat_precomp ? transMat_precomp : s->transMat;

    double *aT = (double *)calloc(9 * (size_t)P + 1, 8);   /* dL_dtransMat */
    double *aM2 = (double *)calloc(2 * (size_t)P + 1, 8);  /* dL_dmean2D.xy */
    double *aN = (double *)calloc(3 * (size_t)P + 1, 8);   /* dL_dnormal3D */
    double *aO = (double *)calloc((size_t)P + 1, 8);       /* dL_dopacity */
    double *aC = (double *)calloc(3 * (size_t)P + 1, 8);   /* dL_dcolors */

    /* ---------------- render backward, backward.cu:143-446 ---------------- */
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < gx * gy; tile++) {
        int tx = tile % gx, ty = tile / gx;
        uint32_t r0 = s->ranges[2 * tile], r1 = s->ranges[2 * tile + 1];
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                int pxi = tx * TILE + lx, pyi = ty * TILE + ly;
                if (pxi >= W || pyi >= H) continue;
                size_t pix_id = (size_t)W * pyi + pxi;
                float pxf = (float)pxi, pyf = (float)pyi;
                const float T_final = s->final_T[pix_id];
                float T = T_final;
                uint32_t contributor = r1 - r0;
                const int last_contributor = (int)s->n_contrib[pix_id];
                const int median_contributor = (int)s->n_contrib[pix_id + HW];
                float accum_rec[3] = {0, 0, 0}, dL_dpixel[3];
                for (int c = 0; c < 3; c++) dL_dpixel[c] = dL_dpix[c * HW + pix_id];
                float dL_ddepth = dL_dothers[0 * HW + pix_id];
                float dL_daccum = dL_dothers[1 * HW + pix_id];
                float dL_dreg = dL_dothers[6 * HW + pix_id];
                float dL_dnormal2D[3] = {dL_dothers[2 * HW + pix_id], dL_dothers[3 * HW + pix_id], dL_dothers[4 * HW + pix_id]};
                float dL_dmedian_depth = dL_dothers[5 * HW + pix_id];
                float last_depth = 0, last_normal[3] = {0, 0, 0}, accum_depth_rec = 0, accum_alpha_rec = 0;
                float accum_normal_rec[3] = {0, 0, 0};
                const float final_D = s->final_T[pix_id + HW], final_D2 = s->final_T[pix_id + 2 * HW];
                const float final_A = 1 - T_final;
                float last_dL_dT = 0, last_alpha = 0, last_color[3] = {0, 0, 0};
                for (uint32_t ii = r1; ii > r0; ii--) {
                    uint32_t i = ii - 1;
                    contributor--;
                    /* backward.cu:279: uint32 >= int promotes to unsigned */
                    if (contributor >= (uint32_t)last_contributor) continue;
                    uint32_t g = s->point_list[i];
                    const float *Tu = tm + 9 * (size_t)g, *Tv = Tu + 3, *Tw = Tu + 6;
                    float k[3] = {pxf * Tw[0] - Tu[0], pxf * Tw[1] - Tu[1], pxf * Tw[2] - Tu[2]};
                    float l[3] = {pyf * Tw[0] - Tv[0], pyf * Tw[1] - Tv[1], pyf * Tw[2] - Tv[2]};
                    float p[3] = {k[1] * l[2] - k[2] * l[1], k[2] * l[0] - k[0] * l[2], k[0] * l[1] - k[1] * l[0]};
                    if (p[2] == 0.0f) continue;
                    float sx = p[0] / p[2], sy = p[1] / p[2];
                    float rho3d = sx * sx + sy * sy;
                    float dx = s->means2D[2 * (size_t)g] - pxf, dy = s->means2D[2 * (size_t)g + 1] - pyf;
                    float rho2d = FILTER_INV_SQUARE * (dx * dx + dy * dy);
                    float rho = fminf(rho3d, rho2d);
                    float c_d = (sx * Tw[0] + sy * Tw[1]) + Tw[2];
                    if (c_d < NEAR_N) continue;
                    const float *no = s->normal_opacity + 4 * (size_t)g;
                    float power = -0.5f * rho;
                    if (power > 0.0f) continue;
                    const float G = expf(power);
                    const float alpha = fminf(0.99f, no[3] * G);
                    if (alpha < 1.0f / 255.0f) continue;
                    T = T / (1.f - alpha);
                    const float w = alpha * T;
                    float dL_dalpha = 0.0f;
                    for (int ch = 0; ch < 3; ch++) {
                        const float c = feat[3 * (size_t)g + ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                        last_color[ch] = c;
                        dL_dalpha += (c - accum_rec[ch]) * dL_dpixel[ch];
                        atomic_add_d(&aC[3 * (size_t)g + ch], w * dL_dpixel[ch]);
                    }
                    float dL_dz = 0.0f, dL_dweight = 0;
                    const float m_d = FAR_N / (FAR_N - NEAR_N) * (1 - NEAR_N / c_d);
                    const float dmd_dd = (FAR_N * NEAR_N) / ((FAR_N - NEAR_N) * c_d * c_d);
                    /* backward.cu:347: uint32 == int(-1 ...) in unsigned arithmetic (quirk 8) */
                    if (contributor == (uint32_t)(median_contributor - 1)) dL_dz += dL_dmedian_depth;
                    dL_dweight += (final_D2 + m_d * m_d * final_A - 2 * m_d * final_D) * dL_dreg;
                    dL_dalpha += dL_dweight - last_dL_dT;
                    last_dL_dT = dL_dweight * alpha + (1 - alpha) * last_dL_dT;
                    const float dL_dmd = 2.0f * (T * alpha) * (m_d * final_A - final_D) * dL_dreg;
                    dL_dz += dL_dmd * dmd_dd;
                    accum_depth_rec = last_alpha * last_depth + (1.f - last_alpha) * accum_depth_rec;
                    last_depth = c_d;
                    dL_dalpha += (c_d - accum_depth_rec) * dL_ddepth;
                    accum_alpha_rec = last_alpha * 1.0f + (1.f - last_alpha) * accum_alpha_rec;
                    dL_dalpha += (1 - accum_alpha_rec) * dL_daccum;
                    for (int ch = 0; ch < 3; ch++) {
                        accum_normal_rec[ch] = last_alpha * last_normal[ch] + (1.f - last_alpha) * accum_normal_rec[ch];
                        last_normal[ch] = no[ch];
                        dL_dalpha += (no[ch] - accum_normal_rec[ch]) * dL_dnormal2D[ch];
                        atomic_add_d(&aN[3 * (size_t)g + ch], alpha * T * dL_dnormal2D[ch]);
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    float bg_dot = 0;
                    for (int c = 0; c < 3; c++) bg_dot += bg[c] * dL_dpixel[c];
                    dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                    const float dL_dG = no[3] * dL_dalpha;
                    dL_dz += alpha * T * dL_ddepth;
                    if (rho3d <= rho2d) {
                        float dsx = dL_dG * -G * sx + dL_dz * Tw[0];
                        float dsy = dL_dG * -G * sy + dL_dz * Tw[1];
                        float dsx_pz = dsx / p[2], dsy_pz = dsy / p[2];
                        float dp[3] = {dsx_pz, dsy_pz, -(dsx_pz * sx + dsy_pz * sy)};
                        float dk[3] = {l[1] * dp[2] - l[2] * dp[1], l[2] * dp[0] - l[0] * dp[2], l[0] * dp[1] - l[1] * dp[0]};
                        float dl[3] = {dp[1] * k[2] - dp[2] * k[1], dp[2] * k[0] - dp[0] * k[2], dp[0] * k[1] - dp[1] * k[0]};
                        float dz_dTw[3] = {sx, sy, 1.0f};
                        for (int c = 0; c < 3; c++) {
                            atomic_add_d(&aT[9 * (size_t)g + c], -dk[c]);
                            atomic_add_d(&aT[9 * (size_t)g + 3 + c], -dl[c]);
                            atomic_add_d(&aT[9 * (size_t)g + 6 + c], pxf * dk[c] + pyf * dl[c] + dL_dz * dz_dTw[c]);
                        }
                    } else {
                        const float dG_ddelx = -G * FILTER_INV_SQUARE * dx;
                        const float dG_ddely = -G * FILTER_INV_SQUARE * dy;
                        atomic_add_d(&aM2[2 * (size_t)g], dL_dG * dG_ddelx);
                        atomic_add_d(&aM2[2 * (size_t)g + 1], dL_dG * dG_ddely);
                        atomic_add_d(&aT[9 * (size_t)g + 6], sx * dL_dz);
                        atomic_add_d(&aT[9 * (size_t)g + 7], sy * dL_dz);
                        atomic_add_d(&aT[9 * (size_t)g + 8], dL_dz);
                    }
                    atomic_add_d(&aO[g], G * dL_dalpha);
                }
            }
    }
    for (size_t i = 0; i < 9 * (size_t)P; i++) dL_dtransMat[i] = (float)aT[i];
    for (size_t i = 0; i < (size_t)P; i++) {
        dL_dmean2D[3 * i] = (float)aM2[2 * i];
        dL_dmean2D[3 * i + 1] = (float)aM2[2 * i + 1];
        dL_dmean2D[3 * i + 2] = 0.f;
        dL_dopacity[i] = (float)aO[i];
    }
    float *dL_dnormal = (float *)malloc(4 * (3 * (size_t)P + 1));
    for (size_t i = 0; i < 3 * (size_t)P; i++) { dL_dnormal[i] = (float)aN[i]; dL_dcolors[i] = (float)aC[i]; }
    free(aT); free(aM2); free(aN); free(aO); free(aC);
    memset(dL_dmean3D, 0, 12 * (size_t)P);
    if (dL_dsh && M > 0) memset(dL_dsh, 0, 12 * (size_t)P * M);
    if (dL_dscale) memset(dL_dscale, 0, 8 * (size_t)P);
    if (dL_drot) memset(dL_drot, 0, 16 * (size_t)P);

    /* ---------------- per-Gaussian backward, backward.cu:581-636 ---------------- */
    const int Wb = f2i(focal_x * tan_fovx * 2); /* backward.cu:613-614 (quirk 3) */
    const int Hb = f2i(focal_y * tan_fovy * 2);
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; idx++) {
        if (!(s->radii[idx] > 0)) continue;
        /* ---- compute_transmat_aabb, backward.cu:449-579 ---- */
        float T[3][3], Pm[3][4], R[3][3], normal[3] = {0, 0, 0};
        const float *p = means3D + 3 * (size_t)idx;
        float sc[2] = {0, 0};
        const int precomp = (scales == NULL); /* backward.cu:615 */
        if (precomp) {
            for (int j = 0; j < 3; j++)
                for (int r = 0; r < 3; r++) T[j][r] = tm[9 * (size_t)idx + 3 * j + r];
        } else {
            sc[0] = scales[2 * (size_t)idx]; sc[1] = scales[2 * (size_t)idx + 1];
            quat_to_rotmat(rotations + 4 * (size_t)idx, R);
            float L0[3], L1[3], L2[3];
            for (int r = 0; r < 3; r++) { L0[r] = R[0][r] * (1.0f * sc[0]); L1[r] = R[1][r] * (1.0f * sc[1]); L2[r] = R[2][r]; }
            float A[4][3] = {{L0[0], L1[0], p[0]}, {L0[1], L1[1], p[1]}, {L0[2], L1[2], p[2]}, {0.f, 0.f, 1.f}};
            float hw = (float)Wb / 2.0f, hw1 = (float)(Wb - 1) / 2.0f, hh = (float)Hb / 2.0f, hh1 = (float)(Hb - 1) / 2.0f;
            /* P = world2ndc * ndc2pix (mat4 * mat3x4): P[j][r] = sum_c W[c][r] N[j][c], W[c][r] = proj[c+4r] */
            for (int r = 0; r < 4; r++) {
                Pm[0][r] = proj[0 + 4 * r] * hw + proj[3 + 4 * r] * hw1;
                Pm[1][r] = proj[1 + 4 * r] * hh + proj[3 + 4 * r] * hh1;
                Pm[2][r] = proj[3 + 4 * r];
            }
            /* T = transpose(M) * P : T[j][r] = sum_k A[k][r] P[j][k] */
            for (int j = 0; j < 3; j++)
                for (int r = 0; r < 3; r++)
                    T[j][r] = A[0][r] * Pm[j][0] + A[1][r] * Pm[j][1] + A[2][r] * Pm[j][2] + A[3][r] * Pm[j][3];
            transform_vec4x3(L2, view, normal);
        }
        float dT[3][3];
        for (int j = 0; j < 3; j++)
            for (int r = 0; r < 3; r++) dT[j][r] = dL_dtransMat[9 * (size_t)idx + 3 * j + r];
        float dm2x = dL_dmean2D[3 * (size_t)idx], dm2y = dL_dmean2D[3 * (size_t)idx + 1];
        int early_return = 0;
        if (dm2x != 0 || dm2y != 0) { /* backward.cu:519-549 */
            float t[3] = {9.0f, 9.0f, -1.0f};
            float d = t[0] * (T[2][0] * T[2][0]) + t[1] * (T[2][1] * T[2][1]) + t[2] * (T[2][2] * T[2][2]);
            float invd = 1.0f / d;
            float f[3] = {t[0] * invd, t[1] * invd, t[2] * invd};
            float dT0[3], dT1[3], dT3[3], df[3];
            for (int r = 0; r < 3; r++) {
                dT0[r] = dm2x * f[r] * T[2][r];
                dT1[r] = dm2y * f[r] * T[2][r];
                dT3[r] = dm2x * f[r] * T[0][r] + dm2y * f[r] * T[1][r];
                df[r] = dm2x * T[0][r] * T[2][r] + dm2y * T[1][r] * T[2][r];
            }
            float dL_dd = (float)((double)((df[0] * f[0] + df[1] * f[1]) + df[2] * f[2]) * (-1.0 / (double)d));
            for (int r = 0; r < 3; r++) {
                float dd_dT3 = t[r] * T[2][r] * 2.0f;
                dT3[r] += dL_dd * dd_dT3;
                dT[0][r] += dT0[r];
                dT[1][r] += dT1[r];
                dT[2][r] += dT3[r];
            }
            if (precomp) {
                for (int j = 0; j < 3; j++)
                    for (int r = 0; r < 3; r++) dL_dtransMat[9 * (size_t)idx + 3 * j + r] = dT[j][r];
                early_return = 1;
            }
        }
        if (!precomp && !early_return) {
            /* dL_dM = P * transpose(dL_dT): dM[j][r] = sum_k P[k][r] dT[k][j] */
            float dM[3][4];
            for (int j = 0; j < 3; j++)
                for (int r = 0; r < 4; r++) dM[j][r] = Pm[0][r] * dT[0][j] + Pm[1][r] * dT[1][j] + Pm[2][r] * dT[2][j];
            float dtn[3];
            transform_vec4x3_transpose(dL_dnormal + 3 * (size_t)idx, view, dtn);
            float pv[3];
            transform_point4x3(p, view, pv);
            float cosv = -((pv[0] * normal[0] + pv[1] * normal[1]) + pv[2] * normal[2]);
            float mult = cosv > 0 ? 1.f : -1.f;
            dtn[0] *= mult; dtn[1] *= mult; dtn[2] *= mult;
            float dRS[3][3] = {{dM[0][0], dM[0][1], dM[0][2]}, {dM[1][0], dM[1][1], dM[1][2]}, {dtn[0], dtn[1], dtn[2]}};
            float dR[3][3];
            for (int r = 0; r < 3; r++) { dR[0][r] = dRS[0][r] * sc[0]; dR[1][r] = dRS[1][r] * sc[1]; dR[2][r] = dRS[2][r]; }
            quat_to_rotmat_vjp(rotations + 4 * (size_t)idx, dR, dL_drot + 4 * (size_t)idx);
            dL_dscale[2 * (size_t)idx] = (dRS[0][0] * R[0][0] + dRS[0][1] * R[0][1]) + dRS[0][2] * R[0][2];
            dL_dscale[2 * (size_t)idx + 1] = (dRS[1][0] * R[1][0] + dRS[1][1] * R[1][1]) + dRS[1][2] * R[1][2];
            dL_dmean3D[3 * (size_t)idx] = dM[2][0];
            dL_dmean3D[3 * (size_t)idx + 1] = dM[2][1];
            dL_dmean3D[3 * (size_t)idx + 2] = dM[2][2];
        }
        /* ---- SH VJP, backward.cu:20-139 ---- */
        if (shs) {
            float dox = p[0] - campos[0], doy = p[1] - campos[1], doz = p[2] - campos[2];
            float len = sqrtf(dox * dox + doy * doy + doz * doz);
            float x = dox / len, y = doy / len, z = doz / len;
            const float *sh = shs + (size_t)idx * M * 3;
            float *dsh = dL_dsh + (size_t)idx * M * 3;
            float dRGB[3];
            for (int c = 0; c < 3; c++) dRGB[c] = dL_dcolors[3 * (size_t)idx + c] * (s->clamped[3 * (size_t)idx + c] ? 0.f : 1.f);
            float dRGBdx[3] = {0, 0, 0}, dRGBdy[3] = {0, 0, 0}, dRGBdz[3] = {0, 0, 0};
#define SH(k, c) sh[3 * (k) + (c)]
#define DSH(k, v) for (int c = 0; c < 3; c++) dsh[3 * (k) + c] = (v) * dRGB[c]
            DSH(0, SH_C0);
            if (D > 0) {
                DSH(1, -SH_C1 * y); DSH(2, SH_C1 * z); DSH(3, -SH_C1 * x);
                for (int c = 0; c < 3; c++) { dRGBdx[c] = -SH_C1 * SH(3, c); dRGBdy[c] = -SH_C1 * SH(1, c); dRGBdz[c] = SH_C1 * SH(2, c); }
                if (D > 1) {
                    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    DSH(4, SH_C2[0] * xy); DSH(5, SH_C2[1] * yz); DSH(6, SH_C2[2] * (2.f * zz - xx - yy));
                    DSH(7, SH_C2[3] * xz); DSH(8, SH_C2[4] * (xx - yy));
                    for (int c = 0; c < 3; c++) {
                        dRGBdx[c] += SH_C2[0] * y * SH(4, c) + SH_C2[2] * 2.f * -x * SH(6, c) + SH_C2[3] * z * SH(7, c) + SH_C2[4] * 2.f * x * SH(8, c);
                        dRGBdy[c] += SH_C2[0] * x * SH(4, c) + SH_C2[1] * z * SH(5, c) + SH_C2[2] * 2.f * -y * SH(6, c) + SH_C2[4] * 2.f * -y * SH(8, c);
                        dRGBdz[c] += SH_C2[1] * y * SH(5, c) + SH_C2[2] * 2.f * 2.f * z * SH(6, c) + SH_C2[3] * x * SH(7, c);
                    }
                    if (D > 2) {
                        DSH(9, SH_C3[0] * y * (3.f * xx - yy)); DSH(10, SH_C3[1] * xy * z);
                        DSH(11, SH_C3[2] * y * (4.f * zz - xx - yy)); DSH(12, SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
                        DSH(13, SH_C3[4] * x * (4.f * zz - xx - yy)); DSH(14, SH_C3[5] * z * (xx - yy));
                        DSH(15, SH_C3[6] * x * (xx - 3.f * yy));
                        for (int c = 0; c < 3; c++) {
                            dRGBdx[c] += (SH_C3[0] * SH(9, c) * 3.f * 2.f * xy + SH_C3[1] * SH(10, c) * yz + SH_C3[2] * SH(11, c) * -2.f * xy +
                                          SH_C3[3] * SH(12, c) * -3.f * 2.f * xz + SH_C3[4] * SH(13, c) * (-3.f * xx + 4.f * zz - yy) +
                                          SH_C3[5] * SH(14, c) * 2.f * xz + SH_C3[6] * SH(15, c) * 3.f * (xx - yy));
                            dRGBdy[c] += (SH_C3[0] * SH(9, c) * 3.f * (xx - yy) + SH_C3[1] * SH(10, c) * xz + SH_C3[2] * SH(11, c) * (-3.f * yy + 4.f * zz - xx) +
                                          SH_C3[3] * SH(12, c) * -3.f * 2.f * yz + SH_C3[4] * SH(13, c) * -2.f * xy + SH_C3[5] * SH(14, c) * -2.f * yz +
                                          SH_C3[6] * SH(15, c) * -3.f * 2.f * xy);
                            dRGBdz[c] += (SH_C3[1] * SH(10, c) * xy + SH_C3[2] * SH(11, c) * 4.f * 2.f * yz + SH_C3[3] * SH(12, c) * 3.f * (2.f * zz - xx - yy) +
                                          SH_C3[4] * SH(13, c) * 4.f * 2.f * xz + SH_C3[5] * SH(14, c) * (xx - yy));
                        }
                    }
                }
            }
#undef SH
#undef DSH
            float ddir[3] = {(dRGBdx[0] * dRGB[0] + dRGBdx[1] * dRGB[1]) + dRGBdx[2] * dRGB[2],
                             (dRGBdy[0] * dRGB[0] + dRGBdy[1] * dRGB[1]) + dRGBdy[2] * dRGB[2],
                             (dRGBdz[0] * dRGB[0] + dRGBdz[1] * dRGB[1]) + dRGBdz[2] * dRGB[2]};
            /* auxiliary.h:128-138 dnormvdv */
            float sum2 = dox * dox + doy * doy + doz * doz;
            float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
            float gx_ = ((+sum2 - dox * dox) * ddir[0] - doy * dox * ddir[1] - doz * dox * ddir[2]) * invsum32;
            float gy_ = (-dox * doy * ddir[0] + (sum2 - doy * doy) * ddir[1] - doz * doy * ddir[2]) * invsum32;
            float gz_ = (-dox * doz * ddir[0] - doy * doz * ddir[1] + (sum2 - doz * doz) * ddir[2]) * invsum32;
            dL_dmean3D[3 * (size_t)idx] += gx_;
            dL_dmean3D[3 * (size_t)idx + 1] += gy_;
            dL_dmean3D[3 * (size_t)idx + 2] += gz_;
        }
        /* ---- densification proxy, backward.cu:631-635 (quirk 4) ---- */
        float depth = tm[9 * (size_t)idx + 8];
        dL_dmean2D[3 * (size_t)idx] = (float)((double)(dL_dtransMat[9 * (size_t)idx + 2] * depth) * 0.5 * (double)(float)Wb);
        dL_dmean2D[3 * (size_t)idx + 1] = (float)((double)(dL_dtransMat[9 * (size_t)idx + 5] * depth) * 0.5 * (double)(float)Hb);
    }
    if (dL_dnormal_out) memcpy(dL_dnormal_out, dL_dnormal, 12 * (size_t)P);
    free(dL_dnormal);
}

/* markVisible: rasterizer_impl.cu:54-66,141-153 */
void oracle_mark_visible(int P, const float *means3D, const float *view, const float *proj, uint8_t *present)
{
    (void)proj;
    for (int i = 0; i < P; i++) {
        float pv[3];
        transform_point4x3(means3D + 3 * (size_t)i, view, pv);
        present[i] = pv[2] > 0.2f;
    }
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
