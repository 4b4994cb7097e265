// render_fwd.cu -- K6: per-tile front-to-back blend of colour, depth, normal, distortion moments.
//
// Behavioural specification: RAST/cuda_rasterizer/forward.cu:256-448 (renderCUDA) -- the per-pixel
// recurrence is restated in SURVEY.md Appendix A and kept in the reference's operation order so
// the integer decisions (alpha < 1/255, T(1-alpha) < 1e-4, T > 0.5) fall identically (the forward
// outputs are bit-identical to the reference extension on the same GPU; tests check it).
//
// What is different from the reference kernel:
//   * one packed 96-B record per instance, gathered into shared memory by the TMA engine
//     (cp.async.bulk + mbarrier transaction bytes) by a dedicated producer warp through a
//     multi-stage ring (tile_pipeline.cuh); colour lives in the record (the reference re-reads it from
//     global memory inside the inner loop, forward.cu:418); no __syncthreads per batch;
//   * each consumer warp owns an 8x4 pixel block of the 16x16 tile and first *compacts* each chunk:
//     32 instances are tested in parallel (one per lane) against the warp's block using the
//     conservative alpha >= 1/255 footprint from preprocess (ellipse + low-pass disc); only
//     surviving instances are evaluated by the 32 pixels.  Skipped instances can never pass the
//     reference's alpha test, so results are unchanged (tests compare culling on/off bit-for-bit);
//   * warp-granular early-out: a warp whose 32 pixels are saturated stops evaluating, and the
//     producer stops streaming once all eight warps are.
#include "kernels.h"
#include "tile_pipeline.cuh"

namespace surfel {

// CLASSES = false: the operator's colour pass (3 colour channels + the 7 allmap channels).
// CLASSES = true:  class-probability pass: word 15 of the record holds a class label instead of
//   red, the pass accumulates w into the label's channel -- bit for bit what the reference gets from one-hot
//   "colours" (1 * w and 0 * w are exact) -- for up to MAX_CLASSES channels in ONE traversal of the lists, and writes
//   only the class images plus the per-pixel state the backward needs (T, last contributor).
// PEERS = true (multi-GPU path, exchange fused into the blend): the ten output planes of this rank's tiles are stored
// straight into the image buffers of ALL ranks over NVLink -- either one `multimem.st` per value through the NVSwitch
// multicast address of the symmetric buffer (the switch replicates it to every GPU), or one plain store per peer.  The
// stores are fire-and-forget and overlap with the blend of the other tiles; a cross-rank barrier after the kernel
// replaces the all-reduce of ten mostly-zero planes (streetunveiler_b200/sharded.py).
struct PeerPlanes {
    float *base[MAX_RANKS];   // [10,H,W] planes of every destination (colour 0..2, allmap 3..9)
    int n;                    // number of destinations (1 with multicast)
    int multicast;            // base[0] is a multicast address: use multimem.st
};

__device__ __forceinline__ void peer_store(float *p, const float v, const int multicast)
{
    if (multicast)
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
    else
        *p = v;
}

#ifdef SURFEL_FWD_MAXREG   // tuning experiments only (tools/gpu_r02_o.sh); the default lets the compiler pick 56 (4 CTAs/SM)
#define SURFEL_FWD_BOUNDS __maxnreg__(SURFEL_FWD_MAXREG)
#else
#define SURFEL_FWD_BOUNDS __launch_bounds__(TILE_THREADS)
#endif
template <bool CULL, bool CLASSES, bool PEERS = false>
__global__ void SURFEL_FWD_BOUNDS
render_fwd_kernel(const PeerPlanes peers, const int n_classes,
                  const int W, const int H, const int gx, const uint32_t *__restrict__ tile_order,
                  const uint2 *__restrict__ ranges,
                  const uint32_t *__restrict__ point_list, const float *__restrict__ rec,
                  const float *__restrict__ bg, float *__restrict__ final_T, uint32_t *__restrict__ n_contrib,
                  uint32_t *__restrict__ tile_max_contrib, float *__restrict__ out_color,
                  float *__restrict__ out_others)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileRing<FWD_STAGES> &ring = *reinterpret_cast<TileRing<FWD_STAGES> *>(smem_raw);
    __shared__ uint32_t s_max_contrib;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    // blockIdx.x enumerates the tiles of this call's window, longest instance lists first
    const int tile = (int)tile_order[blockIdx.x];
    const uint2 range = ranges[tile];
    const int total = (int)(range.y - range.x);
    const int nchunks = (total + CHUNK - 1) / CHUNK;

    if (tid == 0) s_max_contrib = 0;
    ring_init(ring, tid);

    if (warp == CONSUMER_WARPS) {
        // ------------------------------ producer warp ------------------------------
        const uint32_t base = range.x;
        ring_produce<false, FWD_STAGES>(ring, lane, total, point_list, rec, [base](int i) { return base + (uint32_t)i; });
        __syncthreads();
        return;
    }

    // ------------------------------ consumer warps ------------------------------
    const int tile_x = tile % gx, tile_y = tile / gx;
    const int bx0 = tile_x * TILE_X + (warp & 1) * 8, by0 = tile_y * TILE_Y + (warp >> 1) * 4;
    const uint32_t pix_x = bx0 + (lane & 7), pix_y = by0 + (lane >> 3);
    const bool inside = pix_x < (uint32_t)W && pix_y < (uint32_t)H;
    const float2 pixf = make_float2((float)pix_x, (float)pix_y);

    bool done = !inside;
    float T = 1.0f;
    constexpr int NC = CLASSES ? MAX_CLASSES : 3;
    float C[NC], N[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < NC; k++) C[k] = 0.f;
    float Dacc = 0.f, M1 = 0.f, M2 = 0.f, distortion = 0.f, median_depth = 0.f;
    uint32_t last_contributor = 0, median_contributor = 0;  // float -1 -> u32 saturates to 0 (quirk 8)

    bool warp_done = __all_sync(0xffffffffu, done);
    if (warp_done && lane == 0) atomicAdd(const_cast<int *>(&ring.done_warps), 1);

    int stage = 0;
    uint32_t phase = 0;
    for (int c = 0; c < nchunks; c++) {
        if (!warp_done) {
            mbar_wait(&ring.full[stage], phase);
            const int n = min(CHUNK, total - c * CHUNK);
            const float *sb = ring.rec[stage];
#pragma unroll 1
            for (int h = 0; h < n; h += 32) {
                uint32_t mask;
                if (CULL) {
                    const int j = h + lane;
                    const bool hit = (j < n) && block_may_contribute(sb + j * REC_FLOATS, bx0, by0);
                    mask = __ballot_sync(0xffffffffu, hit);
                } else {
                    mask = (n - h >= 32) ? 0xffffffffu : ((1u << (n - h)) - 1u);
                }
                while (mask) {
                    const int jj = h + __ffs(mask) - 1;
                    mask &= mask - 1;
                    if (!done) {
                        // (plain generic loads on purpose: with explicit ld.shared asm the compiler contracted the
                        // intersection arithmetic into different FMAs and the forward stopped being bit-identical
                        // to the reference -- measured 1.7e-7, which the caller's depth-difference normals amplify)
                        const float4 *r4 = reinterpret_cast<const float4 *>(sb + jj * REC_FLOATS);
                        const float4 q0 = r4[0], q1 = r4[1], q2 = r4[2];
                        const float3 Tu = make_float3(q0.x, q0.y, q0.z);
                        const float3 Tv = make_float3(q0.w, q1.x, q1.y);
                        const float3 Tw = make_float3(q1.z, q1.w, q2.x);
                        // two homogeneous planes through the pixel, intersected with the splat plane
                        const float3 k = make_float3(pixf.x * Tw.x - Tu.x, pixf.x * Tw.y - Tu.y, pixf.x * Tw.z - Tu.z);
                        const float3 l = make_float3(pixf.y * Tw.x - Tv.x, pixf.y * Tw.y - Tv.y, pixf.y * Tw.z - Tv.z);
                        const float3 p = make_float3(k.y * l.z - k.z * l.y, k.z * l.x - k.x * l.z, k.x * l.y - k.y * l.x);
                        if (p.z != 0.0f) {
                            const float2 s = make_float2(p.x / p.z, p.y / p.z);
                            const float rho3d = (s.x * s.x + s.y * s.y);
                            const float2 d = make_float2(q2.y - pixf.x, q2.z - pixf.y);
                            const float rho2d = FILTER_INV_SQUARE * (d.x * d.x + d.y * d.y);
                            const float rho = fminf(rho3d, rho2d);
                            const float depth = (s.x * Tw.x + s.y * Tw.y) + Tw.z;
                            const float power = -0.5f * rho;
                            if (!(depth < NEAR_N) && !(power > 0.0f)) {
                                const float alpha = fminf(0.99f, q2.w * expf(power));
                                if (!(alpha < ALPHA_MIN)) {
                                    const float test_T = T * (1 - alpha);
                                    if (test_T < T_EPS) {
                                        done = true;
                                    } else {
                                        const float4 q3 = r4[3];
                                        const uint32_t contributor = (uint32_t)(c * CHUNK + jj + 1);
                                        const float w = alpha * T;
                                        if constexpr (CLASSES) {
                                            const int label = __float_as_int(q3.w);   // warp-uniform
#pragma unroll
                                            for (int k = 0; k < NC; k++) C[k] += (label == k) ? w : 0.f;
                                        } else {
                                            const float2 q4 = *reinterpret_cast<const float2 *>(r4 + 4);
                                            const float A = 1 - T;
                                            const float m = FAR_N / (FAR_N - NEAR_N) * (1 - NEAR_N / depth);
                                            distortion += (m * m * A + M2 - 2 * m * M1) * w;
                                            Dacc += depth * w;
                                            M1 += m * w;
                                            M2 += m * m * w;
                                            if (T > 0.5f) {
                                                median_depth = depth;
                                                median_contributor = contributor;
                                            }
                                            N[0] += q3.x * w;
                                            N[1] += q3.y * w;
                                            N[2] += q3.z * w;
                                            C[0] += q3.w * w;
                                            C[1] += q4.x * w;
                                            C[2] += q4.y * w;
                                        }
                                        T = test_T;
                                        last_contributor = contributor;
                                    }
                                }
                            }
                        }
                    }
                }
                if (__all_sync(0xffffffffu, done)) {
                    warp_done = true;
                    break;
                }
            }
            if (warp_done && lane == 0) atomicAdd(const_cast<int *>(&ring.done_warps), 1);
        } else if (!ring_wait_or_quit(ring, lane, stage, phase, c)) {
            break;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ring.empty[stage]);
        if (++stage == FWD_STAGES) { stage = 0; phase ^= 1u; }
    }

    // per-tile maximum of last_contributor: lets the backward pass skip the untouched list tail
    uint32_t mx = inside ? last_contributor : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) atomicMax(&s_max_contrib, mx);
    __syncthreads();
    if (tid == 0) tile_max_contrib[tile] = s_max_contrib;

    if (inside) {
        const size_t HW = (size_t)W * H;
        const size_t pix_id = (size_t)W * pix_y + pix_x;
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
        if constexpr (CLASSES) {
#pragma unroll
            for (int k = 0; k < NC; k++)
                if (k < n_classes) out_color[pix_id + k * HW] = C[k] + T * bg[k];
            return;
        }
        final_T[pix_id + HW] = M1;
        final_T[pix_id + 2 * HW] = M2;
        n_contrib[pix_id + HW] = median_contributor;
        if constexpr (PEERS) {
            const float v[10] = {C[0] + T * bg[0], C[1] + T * bg[1], C[2] + T * bg[2], Dacc, 1 - T, N[0], N[1], N[2],
                                 median_depth, distortion};
            for (int q = 0; q < peers.n; q++) {
                float *o = peers.base[q] + pix_id;
#pragma unroll
                for (int k = 0; k < 10; k++) peer_store(o + k * HW, v[k], peers.multicast);
            }
            return;
        }
        out_color[pix_id] = C[0] + T * bg[0];
        out_color[pix_id + HW] = C[1] + T * bg[1];
        out_color[pix_id + 2 * HW] = C[2] + T * bg[2];
        out_others[pix_id] = Dacc;
        out_others[pix_id + HW] = 1 - T;
        out_others[pix_id + 2 * HW] = N[0];
        out_others[pix_id + 3 * HW] = N[1];
        out_others[pix_id + 4 * HW] = N[2];
        out_others[pix_id + 5 * HW] = median_depth;
        out_others[pix_id + 6 * HW] = distortion;
    }
}

template <bool CULL, bool CLASSES, bool PEERS = false>
static void launch_fwd(const RenderFwdArgs &a, const int tiles, cudaStream_t stream)
{
    auto k = render_fwd_kernel<CULL, CLASSES, PEERS>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileRing<FWD_STAGES>));  // per device, cheap
    PeerPlanes peers{};
    if (PEERS) {
        peers.n = a.n_peers;
        peers.multicast = a.peer_multicast;
        for (int q = 0; q < a.n_peers && q < MAX_RANKS; q++) peers.base[q] = a.peer_planes[q];
    }
    k<<<tiles, TILE_THREADS, sizeof(TileRing<FWD_STAGES>), stream>>>(peers, a.n_classes, a.W, a.H, a.gx, a.tile_order, a.ranges, a.point_list,
                                                                      a.rec, a.bg, a.final_T, a.n_contrib, a.tile_max_contrib,
                                                                      a.out_color, a.out_others);
}

void launch_render_fwd(const RenderFwdArgs &a, cudaStream_t stream)
{
    const int tiles = (a.tile_hi < 0 ? a.gx * a.gy : a.tile_hi) - a.tile_lo;
    if (tiles <= 0) return;
    if (a.n_classes > 0)
        launch_fwd<true, true>(a, tiles, stream);
    else if (a.n_peers > 0)
        launch_fwd<true, false, true>(a, tiles, stream);
    else if (a.subtile_cull)
        launch_fwd<true, false>(a, tiles, stream);
    else
        launch_fwd<false, false>(a, tiles, stream);
}

}  // namespace surfel
