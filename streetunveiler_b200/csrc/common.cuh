// common.cuh -- shared definitions of the sm_100a surfel rasterizer kernels.
//
// Reference semantics: RAST/cuda_rasterizer/{auxiliary.h,config.h} ("RAST/" =
// /root/reference/submodules/diff-surfel-rasterization/).  Nothing here is shared with the
// reference sources; constants are restated from auxiliary.h:21-60 and config.h:15-17.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace surfel {

constexpr int TILE_X = 16;        // config.h:16
constexpr int TILE_Y = 16;        // config.h:17
constexpr int TILE_PIX = TILE_X * TILE_Y;
constexpr float NEAR_N = 0.2f;    // auxiliary.h:37
constexpr float FAR_N = 100.0f;   // auxiliary.h:38
constexpr float FILTER_SIZE = 0.707106f;      // auxiliary.h:39
constexpr float FILTER_INV_SQUARE = 2.0f;     // auxiliary.h:40
constexpr float ALPHA_MIN = 1.0f / 255.0f;    // forward.cu:386
constexpr float T_EPS = 0.0001f;              // forward.cu:389

// ---------------------------------------------------------------------------------------------
// Per-Gaussian projected record: the ONLY thing the two render kernels read per instance.
// 96 B = 6 x 16 B, 16-B aligned, so one record is one bulk async copy (cp.async.bulk) into smem.
//   q0 = (Tu.x, Tu.y, Tu.z, Tv.x)
//   q1 = (Tv.y, Tv.z, Tw.x, Tw.y)
//   q2 = (Tw.z, mean2D.x, mean2D.y, opacity)
//   q3 = (n.x, n.y, n.z, r)
//   q4 = (g, b, bbox, e.x)
//   q5 = (e.y, M00, M01, M11)
// (Tu,Tv,Tw) = rows of the splat->pixel homogeneous map (reference geomState.transMat,
// rasterizer_impl.h:38), n = view-space normal flipped towards the camera, rgb = SH colour or
// colors_precomp.  (bbox, e, M) is the conservative footprint of {alpha >= 1/255} used for sub-tile
// culling: bbox = 4 x u8 bounds in 8-pixel units; inside it the pixel p can only contribute if
// (p-e)^T M (p-e) <= 1 (perspective-correct ellipse rho3d <= tau) or |p - mean2D|^2 <= tau/2 (low-pass
// disc rho2d <= tau), tau = 2 ln(255 opacity) slightly inflated.  M = 0: no ellipse bound (ill-conditioned).
// ---------------------------------------------------------------------------------------------
constexpr int REC_FLOATS = 24;
constexpr int REC_BYTES = REC_FLOATS * 4;

// Per-Gaussian gradient accumulator written by the backward render kernel and consumed by the
// backward preprocess kernel.  20 floats (80 B):
//   [0..8] dL_dtransMat, [9..10] dL_dmean2D.xy, [11] dL_dopacity, [12..14] dL_dcolor,
//   [15..17] dL_dnormal, [18..19] pad
constexpr int GACC_FLOATS = 20;

// Multi-GPU exchange (exchange.cu): a projected record travels as a 112-B row = record (24 floats) + depth key bits +
// radius (int bits) + 2 pad words; at most MAX_RANKS ranks share one scene.
constexpr int XROW_FLOATS = 28;
constexpr int MAX_RANKS = 16;

// Class-probability pass (render_fwd.cu / render_bwd.cu, CLASSES = true): channels rendered in one traversal.
constexpr int MAX_CLASSES = 8;

template <typename T>
__host__ __device__ inline T *carve(char *&p, size_t count)
{
    uintptr_t a = (reinterpret_cast<uintptr_t>(p) + 127) & ~uintptr_t(127);
    T *r = reinterpret_cast<T *>(a);
    p = reinterpret_cast<char *>(r + count);
    return r;
}

// Private layout of the geometry scratch (forward -> backward state, P-indexed).
struct GeomView {
    float *rec;              // [P][REC_FLOATS]
    uint32_t *tiles_touched; // [P]
    uint32_t *depth_key;     // [P]  fp32 bits of view-space z, 0xFFFFFFFF for culled
    uint32_t *depth_key_sorted; // [P]
    uint32_t *idx_in;        // [P]  iota
    uint32_t *idx_sorted;    // [P]  Gaussian ids by (depth, id)
    uint32_t *offsets;       // [P]  inclusive scan of tiles_touched in sorted order
    uint8_t *clamped;        // [P]  bit c set <=> SH colour channel c was clamped at 0
    char *cub_temp;
    size_t cub_temp_bytes;
};

struct ImageView {
    uint2 *ranges;       // [tiles]
    float *final_T;      // [3][HW]  T, M1, M2
    uint32_t *n_contrib; // [2][HW]  last contributor, median contributor
    uint32_t *tile_max_contrib; // [tiles] max over the tile's pixels of last contributor
    uint32_t *tile_order;       // [tiles] launch order (longest lists first)
    int *aux_flag;              // one word: "some depth/normal/distortion gradient is non-zero" (render_bwd.cu)
};

struct BinView {
    uint32_t *keys_unsorted; // [R] tile id
    uint32_t *vals_unsorted; // [R] Gaussian id
    uint32_t *keys_sorted;   // [R]
    uint32_t *point_list;    // [R]
    char *cub_temp;
    size_t cub_temp_bytes;
};

struct float3x3 {  // three rows/columns as plain floats
    float m[3][3];
};

__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }

}  // namespace surfel
