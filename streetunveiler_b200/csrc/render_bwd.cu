// render_bwd.cu -- K7: back-to-front gradient pass of the tile blend.
//
// Behavioural specification: RAST/cuda_rasterizer/backward.cu:143-446 (renderCUDA); the per-pixel
// recurrences are restated in SURVEY.md Appendix A ("Backward, per pixel") and reproduced with
// the reference's conventions (its quirks 5, 8: clamp ignored, tie goes to the 3-D branch,
// mixed-sign median comparison).
//
// What is different from the reference kernel (which issues up to 16 scalar float atomicAdds per
// contributing pixel-instance, backward.cu:339-443):
//   * the same packed 96-B records / producer-warp TMA ring / 8x4-pixel warp blocks / per-warp chunk
//     compaction as the forward kernel (tile_pipeline.cuh), plus a per-tile bound (max last
//     contributor, saved by the forward) so the list tail nobody blended is never even loaded;
//   * the 18 per-instance gradient values are summed over the warp's 32 pixels with a
//     multi-value butterfly (20 shuffles for all 18 values, not 18 x 5), leaving one value per
//     lane, and flushed with ONE reduction instruction per (warp, instance) into an 80-B
//     per-Gaussian accumulator record that the backward-preprocess kernel consumes.
#include "kernels.h"
#include "tile_pipeline.cuh"

namespace surfel {

// Fire-and-forget float reduction (RED, no return value): unlike ATOMG it does not hold a scoreboard
// until the L2 round trip completes, so the warp moves on to the next instance immediately.
__device__ __forceinline__ void red_add_f32(float *addr, const float v)
{
    asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ float fast_rcp(const float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

constexpr int NV = 18;  // gradient values per instance

// Multi-value warp reduction: N values per lane in, one fully reduced value per lane out.
// Each xor step halves the number of live values; lanes with the step's bit set keep the upper half.
template <int N, int BIT>
struct Butterfly {
    static constexpr int H = (N + 1) / 2;
    __device__ __forceinline__ static void run(float (&v)[NV], const int lane)
    {
        const bool hi = (lane >> BIT) & 1;
#pragma unroll
        for (int i = 0; i < H; i++) {
            const float upper = (i + H < N) ? v[i + H] : 0.f;
            const float send = hi ? v[i] : upper;
            const float keep = hi ? upper : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1 << BIT);
        }
        if constexpr (BIT > 0) Butterfly<H, BIT - 1>::run(v, lane);
    }
    // which of the original N values ends up in this lane's v[0] (-1: padding)
    __device__ __forceinline__ static int slot(const int lane, const int base, const int count)
    {
        // count = number of real values in [0, N) at this level for this lane's branch
        const bool hi = (lane >> BIT) & 1;
        const int nb = hi ? base + H : base;
        const int nc = hi ? count - H : (count < H ? count : H);
        if (nc <= 0) return -1;
        if constexpr (BIT > 0) return Butterfly<H, BIT - 1>::slot(lane, nb, nc);
        return nb;
    }
};

// Depth / normal / median-depth / distortion upstream gradients are exactly zero for whole frames in practice
// (train.py enables those losses late; BASELINE configs 2, 3, 5): aux_zero_scan_kernel finds out on the device and
// the launcher queues BOTH specialisations of the blend kernel -- AUX = true (all recurrences) and AUX = false
// (colour + alpha only: 14 fewer live values per pixel) -- each of which returns at once unless the flag selects it.
// No host round trip; the cost is one pass over six gradient planes + the contributor plane and one empty launch.
// Only pixels that blended at least one splat count (n_contrib > 0): the blend kernel never reads the gradients of the
// others, and the reference's caller puts NaN = 0/0 into dL/dallmap[0] exactly there (alpha == 0 pixels of
// gaussian_renderer/__init__.py:158, e.g. sky), which must not force the full specialisation on every street frame.
__global__ void __launch_bounds__(256) aux_zero_scan_kernel(const float *__restrict__ dL_dothers,
                                                            const uint32_t *__restrict__ n_contrib, const size_t HW,
                                                            int *__restrict__ flag)
{
    bool any = false;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
    if ((HW & 3) == 0 && (reinterpret_cast<uintptr_t>(dL_dothers) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(n_contrib) & 15) == 0) {
        // four pixels per thread and step: one 128-bit word of n_contrib and of channels 0, 2..6 (1 = alpha is always
        // consumed) -- seven independent loads in flight
        const float4 *v = reinterpret_cast<const float4 *>(dL_dothers);
        const uint4 *nc = reinterpret_cast<const uint4 *>(n_contrib);
        const size_t q = HW / 4;
        for (size_t i = tid; i < q; i += nthreads) {
            const uint4 c = nc[i];
            float4 x[6];
#pragma unroll
            for (int u = 0; u < 6; u++) x[u] = v[i + (u == 0 ? 0 : (u + 1) * q)];
#pragma unroll
            for (int u = 0; u < 6; u++)
                any |= (c.x != 0u && x[u].x != 0.f) || (c.y != 0u && x[u].y != 0.f) || (c.z != 0u && x[u].z != 0.f) ||
                       (c.w != 0u && x[u].w != 0.f);
        }
    } else {
        for (size_t p = tid; p < HW; p += nthreads)
            any |= n_contrib[p] != 0u &&
                   (dL_dothers[p] != 0.f || dL_dothers[2 * HW + p] != 0.f || dL_dothers[3 * HW + p] != 0.f ||
                    dL_dothers[4 * HW + p] != 0.f || dL_dothers[5 * HW + p] != 0.f || dL_dothers[6 * HW + p] != 0.f);
    }
    if (__syncthreads_or(any) && threadIdx.x == 0) *flag = 1;   // benign race: every writer stores 1
}

// CULL: sub-tile culling on.  AUX: see above.  REGS: register cap -- 72 registers put three 9-warp CTAs on an SM instead
// of two (the kernel is issue-bound with few resident warps: measured -20 % for AUX = false despite ~70 B of spills).
// (Tried and dropped: direct per-lane reductions when only 1-2 pixels of a warp contribute -- slower than the butterfly.)
// CLASSES: backward of the class-probability pass (render_fwd.cu).  With one-hot "colours" the three colour recurrences
// collapse into ONE scalar: sum_k (c_k - accum_rec_k) dL/dpixel_k = dL/dpixel_label - S with
// S <- alpha dL/dpixel_label + (1 - alpha) S, for any number of channels; labels are not trainable, so the colour
// gradient words are not produced (12 values per instance through the butterfly instead of 15).  Requires AUX = false.
template <bool CULL, bool AUX, int REGS, bool CLASSES>
__global__ void __maxnreg__(REGS)
render_bwd_kernel(const int *__restrict__ aux_flag, const int n_classes,
                  const int W, const int H, const int gx, const uint32_t *__restrict__ tile_order,
                  const uint2 *__restrict__ ranges,
                  const uint32_t *__restrict__ point_list, const float *__restrict__ rec,
                  const float *__restrict__ bg, const float *__restrict__ final_Ts,
                  const uint32_t *__restrict__ n_contrib, const uint32_t *__restrict__ tile_max_contrib,
                  const float *__restrict__ dL_dpixels, const float *__restrict__ dL_dothers,
                  float *__restrict__ gacc)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileRing<BWD_STAGES> &ring = *reinterpret_cast<TileRing<BWD_STAGES> *>(smem_raw);

    if (aux_flag && (*aux_flag != 0) != AUX) return;   // the other specialisation handles this frame

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    // blockIdx.x enumerates the tiles of this call's window, longest instance lists first
    const int tile = (int)tile_order[blockIdx.x];
    const uint2 range = ranges[tile];
    // only list positions [0, total) can have been blended by some pixel of this tile
    const int total = min((int)(range.y - range.x), (int)tile_max_contrib[tile]);
    const int nchunks = (total + CHUNK - 1) / CHUNK;
    if (nchunks == 0) return;

    ring_init(ring, tid);

    if (warp == CONSUMER_WARPS) {
        // producer: stream slot i <-> list position total-1-i (back to front)
        const uint32_t last = range.x + (uint32_t)(total - 1);
        ring_produce<true, BWD_STAGES>(ring, lane, total, point_list, rec, [last](int i) { return last - (uint32_t)i; });
        return;
    }

    const int tile_x = tile % gx, tile_y = tile / gx;
    const int bx0 = tile_x * TILE_X + (warp & 1) * 8, by0 = tile_y * TILE_Y + (warp >> 1) * 4;
    const uint32_t pix_x = bx0 + (lane & 7), pix_y = by0 + (lane >> 3);
    const bool inside = pix_x < (uint32_t)W && pix_y < (uint32_t)H;
    const float2 pixf = make_float2((float)pix_x, (float)pix_y);
    const size_t HW = (size_t)W * H;
    const size_t pix_id = (size_t)W * pix_y + pix_x;

    const float T_final = inside ? final_Ts[pix_id] : 0;
    float T = T_final;
    const uint32_t last_contributor = inside ? n_contrib[pix_id] : 0;
    const int median_contributor = inside ? (int)n_contrib[pix_id + HW] : 0;
    const uint32_t median_match = (uint32_t)(median_contributor - 1);  // backward.cu:347, unsigned compare

    static_assert(!(CLASSES && AUX), "the class-probability pass has no depth/normal/distortion outputs");
    constexpr int NC = CLASSES ? MAX_CLASSES : 3;
    float accum_rec[3] = {0.f, 0.f, 0.f}, dL_dpixel[NC];   // CLASSES: accum_rec[0] is the scalar S
#pragma unroll
    for (int k = 0; k < NC; k++) dL_dpixel[k] = 0.f;
    float dL_dreg = 0.f, dL_ddepth = 0.f, dL_daccum = 0.f, dL_dmedian_depth = 0.f;
    float dL_dnormal2D[3] = {0.f, 0.f, 0.f};
    if (inside) {
        if (!CLASSES) dL_daccum = dL_dothers[HW + pix_id];
        if (AUX) {
            dL_ddepth = dL_dothers[pix_id];
            dL_dnormal2D[0] = dL_dothers[2 * HW + pix_id];
            dL_dnormal2D[1] = dL_dothers[3 * HW + pix_id];
            dL_dnormal2D[2] = dL_dothers[4 * HW + pix_id];
            dL_dmedian_depth = dL_dothers[5 * HW + pix_id];
            dL_dreg = dL_dothers[6 * HW + pix_id];
        }
#pragma unroll
        for (int k = 0; k < NC; k++)
            if (!CLASSES || k < n_classes) dL_dpixel[k] = dL_dpixels[k * HW + pix_id];
    }
    float accum_depth_rec = 0.f, accum_alpha_rec = 0.f, last_dL_dT = 0.f;
    float accum_normal_rec[3] = {0.f, 0.f, 0.f};
    const float final_D = inside ? final_Ts[pix_id + HW] : 0;
    const float final_D2 = inside ? final_Ts[pix_id + 2 * HW] : 0;
    const float final_A = 1 - T_final;
    float bg_dot_dpixel = 0;
#pragma unroll
    for (int i = 0; i < NC; i++)
        if (!CLASSES || i < n_classes) bg_dot_dpixel += bg[i] * dL_dpixel[i];

    // highest list position any pixel of this warp blended, +1
    uint32_t warp_last = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, o));

    const int my_slot = Butterfly<NV, 4>::slot(lane, 0, NV);
    const int my_slot15 = Butterfly<15, 4>::slot(lane, 0, 15);  // without the three normal gradients
    const int my_slot12 = Butterfly<12, 4>::slot(lane, 0, 12);  // CLASSES: without the colour gradients either

    // ... and within a frame that has some, skip their recurrences for warps none of whose pixels has any.
    const bool aux_any = AUX && __any_sync(0xffffffffu, dL_ddepth != 0.f || dL_dreg != 0.f || dL_dmedian_depth != 0.f ||
                                                            dL_dnormal2D[0] != 0.f || dL_dnormal2D[1] != 0.f ||
                                                            dL_dnormal2D[2] != 0.f);

    int stage = 0;
    uint32_t phase = 0;
    for (int cb = 0; cb < nchunks; cb++) {
        const int n = min(CHUNK, total - cb * CHUNK);
        // chunk slot s holds list position pos0 - s (descending); skip chunks entirely behind this warp's pixels
        const int pos0 = total - 1 - cb * CHUNK;
        const bool chunk_live = (uint32_t)(pos0 - (n - 1)) < warp_last;
        // Always wait for the stage, even when skipping it: arriving on `empty` without having observed
        // `full` would let this warp lap the ring and double-count in a phase other warps still read.
        mbar_wait(&ring.full[stage], phase);
        const float *sb = ring.rec[stage];
        for (int c = 0; chunk_live && c < n; c += 32) {
            const int pos_first = pos0 - c;
            const int pos_min = pos_first - (min(32, n - c) - 1);
            if ((uint32_t)pos_min >= warp_last) continue;  // entirely behind everything this warp blended
            uint32_t mask;
            {
                const int j = c + lane;
                bool hit = false;
                if (j < n) {
                    const int pos = pos_first - lane;
                    hit = (uint32_t)pos < warp_last;
                    if (CULL && hit) hit = block_may_contribute(sb + j * REC_FLOATS, bx0, by0);
                }
                mask = __ballot_sync(0xffffffffu, hit);
            }
            while (mask) {
                const int sl = __ffs(mask) - 1;
                mask &= mask - 1;
                const int jj = c + sl;
                const uint32_t contributor = (uint32_t)(pos_first - sl);  // reference's `contributor` after --

                float v[NV];
#pragma unroll
                for (int i = 0; i < NV; i++) v[i] = 0.f;
                bool valid = false;

                if (contributor < last_contributor) {
                    const float4 *r4 = reinterpret_cast<const float4 *>(sb + jj * REC_FLOATS);
                    const float4 q0 = r4[0], q1 = r4[1], q2 = r4[2];
                    const float3 Tu = make_float3(q0.x, q0.y, q0.z);
                    const float3 Tv = make_float3(q0.w, q1.x, q1.y);
                    const float3 Tw = make_float3(q1.z, q1.w, q2.x);
                    const float3 k = make_float3(pixf.x * Tw.x - Tu.x, pixf.x * Tw.y - Tu.y, pixf.x * Tw.z - Tu.z);
                    const float3 l = make_float3(pixf.y * Tw.x - Tv.x, pixf.y * Tw.y - Tv.y, pixf.y * Tw.z - Tv.z);
                    const float3 p = make_float3(k.y * l.z - k.z * l.y, k.z * l.x - k.x * l.z, k.x * l.y - k.y * l.x);
                    if (p.z != 0.0f) {
                        const float2 s = make_float2(p.x / p.z, p.y / p.z);
                        const float rho3d = (s.x * s.x + s.y * s.y);
                        const float2 d = make_float2(q2.y - pixf.x, q2.z - pixf.y);
                        const float rho2d = FILTER_INV_SQUARE * (d.x * d.x + d.y * d.y);
                        const float rho = fminf(rho3d, rho2d);
                        const float c_d = (s.x * Tw.x + s.y * Tw.y) + Tw.z;
                        const float power = -0.5f * rho;
                        if (!(c_d < NEAR_N) && !(power > 0.0f)) {
                            const float G = expf(power);
                            const float opa = q2.w;
                            const float alpha = fminf(0.99f, opa * G);
                            if (!(alpha < ALPHA_MIN)) {
                                valid = true;
                                const float4 q3 = r4[3];
                                const float normal[3] = {q3.x, q3.y, q3.z};

                                // The reference keeps (last_alpha, last_color, ...) and folds them into the
                                // suffix accumulators at the NEXT contributor (backward.cu:329,365-374); folding
                                // them right after use is the same arithmetic on the same operands, with 8 fewer
                                // live registers.  Divisions that only feed gradients (not the alpha / skip
                                // decisions above) use the approximate reciprocal (rcp.approx.ftz, <= 1 ulp; the
                                // reference divides IEEE-exactly): far inside the 1e-4 tolerance and below the
                                // reference's own atomic-order noise.  Stated in include/surfel_rasterizer.h.
                                const float one_m_alpha = 1.f - alpha;
                                // 1 / (1 - alpha) is needed twice (T recovery, background term); one approximate
                                // reciprocal (<= 1 ulp) serves both.  T only feeds gradients here (the forward's T is
                                // not recomputed), and the product T * inv carries ~1.5 ulp per step instead of the
                                // 0.5 ulp of an IEEE division: a random walk of ~1e-6 over a pixel's list, below the
                                // reference's own atomic-order noise (the live test prints the margin to 1e-4).
                                const float inv_oma = fast_rcp(one_m_alpha);
                                T = T * inv_oma;
                                const float w = alpha * T;
                                float dL_dalpha = 0.0f;
                                if constexpr (CLASSES) {
                                    const int label = __float_as_int(q3.w);   // warp-uniform
                                    float g_label = 0.f;
#pragma unroll
                                    for (int k = 0; k < NC; k++) g_label = (label == k) ? dL_dpixel[k] : g_label;
                                    dL_dalpha = g_label - accum_rec[0];
                                    accum_rec[0] = alpha * g_label + one_m_alpha * accum_rec[0];
                                } else {
                                    const float2 q4 = *reinterpret_cast<const float2 *>(r4 + 4);
                                    const float col[3] = {q3.w, q4.x, q4.y};
#pragma unroll
                                    for (int ch = 0; ch < 3; ch++) {
                                        const float cc = col[ch];
                                        dL_dalpha += (cc - accum_rec[ch]) * dL_dpixel[ch];
                                        accum_rec[ch] = alpha * cc + one_m_alpha * accum_rec[ch];
                                        v[12 + ch] = w * dL_dpixel[ch];
                                    }
                                }
                                float dL_dz = 0.0f;
                                if (aux_any) {  // warp-uniform: some pixel of this warp has depth/normal/distortion gradients
                                    const float inv_cd = fast_rcp(c_d);
                                    const float m_d = FAR_N / (FAR_N - NEAR_N) * (1 - NEAR_N * inv_cd);
                                    const float dmd_dd = (FAR_N * NEAR_N / (FAR_N - NEAR_N)) * inv_cd * inv_cd;
                                    if (contributor == median_match) dL_dz += dL_dmedian_depth;
                                    const float dL_dweight = (final_D2 + m_d * m_d * final_A - 2 * m_d * final_D) * dL_dreg;
                                    dL_dalpha += dL_dweight - last_dL_dT;
                                    last_dL_dT = dL_dweight * alpha + one_m_alpha * last_dL_dT;
                                    const float dL_dmd = 2.0f * w * (m_d * final_A - final_D) * dL_dreg;
                                    dL_dz += dL_dmd * dmd_dd;
                                    dL_dalpha += (c_d - accum_depth_rec) * dL_ddepth;
                                    accum_depth_rec = alpha * c_d + one_m_alpha * accum_depth_rec;
#pragma unroll
                                    for (int ch = 0; ch < 3; ch++) {
                                        dL_dalpha += (normal[ch] - accum_normal_rec[ch]) * dL_dnormal2D[ch];
                                        accum_normal_rec[ch] = alpha * normal[ch] + one_m_alpha * accum_normal_rec[ch];
                                        v[15 + ch] = w * dL_dnormal2D[ch];
                                    }
                                    dL_dz += w * dL_ddepth;
                                }
                                if (!CLASSES) {   // the class pass has no alpha-channel output
                                    dL_dalpha += (1 - accum_alpha_rec) * dL_daccum;
                                    accum_alpha_rec = alpha + one_m_alpha * accum_alpha_rec;
                                }
                                dL_dalpha *= T;
                                dL_dalpha += (-T_final * inv_oma) * bg_dot_dpixel;

                                const float dL_dG = opa * dL_dalpha;

                                if (rho3d <= rho2d) {
                                    const float inv_pz = fast_rcp(p.z);
                                    const float nG = -G * dL_dG;
                                    const float dsx_pz = (nG * s.x + dL_dz * Tw.x) * inv_pz;
                                    const float dsy_pz = (nG * s.y + dL_dz * Tw.y) * inv_pz;
                                    const float3 dL_dp = make_float3(dsx_pz, dsy_pz, -(dsx_pz * s.x + dsy_pz * s.y));
                                    const float3 dL_dk = make_float3(l.y * dL_dp.z - l.z * dL_dp.y, l.z * dL_dp.x - l.x * dL_dp.z,
                                                                     l.x * dL_dp.y - l.y * dL_dp.x);
                                    const float3 dL_dl = make_float3(dL_dp.y * k.z - dL_dp.z * k.y, dL_dp.z * k.x - dL_dp.x * k.z,
                                                                     dL_dp.x * k.y - dL_dp.y * k.x);
                                    v[0] = -dL_dk.x; v[1] = -dL_dk.y; v[2] = -dL_dk.z;
                                    v[3] = -dL_dl.x; v[4] = -dL_dl.y; v[5] = -dL_dl.z;
                                    v[6] = pixf.x * dL_dk.x + pixf.y * dL_dl.x + dL_dz * s.x;
                                    v[7] = pixf.x * dL_dk.y + pixf.y * dL_dl.y + dL_dz * s.y;
                                    v[8] = pixf.x * dL_dk.z + pixf.y * dL_dl.z + dL_dz;
                                } else {
                                    const float nG2 = -G * FILTER_INV_SQUARE * dL_dG;
                                    v[9] = nG2 * d.x;
                                    v[10] = nG2 * d.y;
                                    v[6] = s.x * dL_dz;
                                    v[7] = s.y * dL_dz;
                                    v[8] = dL_dz;
                                }
                                v[11] = G * dL_dalpha;
                            }
                        }
                    }
                }
                if (__any_sync(0xffffffffu, valid)) {
                    float *dst = gacc + (size_t)ring.id[stage][jj] * GACC_FLOATS;
                    if constexpr (CLASSES) {
                        Butterfly<12, 4>::run(v, lane);
                        if (my_slot12 >= 0) red_add_f32(dst + my_slot12, v[0]);
                    } else if (aux_any) {
                        Butterfly<NV, 4>::run(v, lane);
                        if (my_slot >= 0) red_add_f32(dst + my_slot, v[0]);
                    } else {  // v[15..17] (normal gradients) are identically zero: 16 shuffles instead of 20
                        Butterfly<15, 4>::run(v, lane);
                        if (my_slot15 >= 0) red_add_f32(dst + my_slot15, v[0]);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ring.empty[stage]);
        if (++stage == BWD_STAGES) { stage = 0; phase ^= 1u; }
    }
}

template <bool CULL, bool AUX, int REGS, bool CLASSES = false>
static void launch_one(const RenderBwdArgs &a, const int tiles, const int *flag, cudaStream_t stream)
{
    auto k = render_bwd_kernel<CULL, AUX, REGS, CLASSES>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileRing<BWD_STAGES>));  // per device, cheap
    k<<<tiles, TILE_THREADS, sizeof(TileRing<BWD_STAGES>), stream>>>(flag, a.n_classes, a.W, a.H, a.gx, a.tile_order, a.ranges, a.point_list,
                                                                      a.rec, a.bg, a.final_T, a.n_contrib, a.tile_max_contrib,
                                                                      a.dL_dpix, a.dL_dothers, a.gacc);
}

void launch_render_bwd(const RenderBwdArgs &a, cudaStream_t stream)
{
    const int tiles = (a.tile_hi < 0 ? a.gx * a.gy : a.tile_hi) - a.tile_lo;
    if (tiles <= 0) return;
    if (a.n_classes > 0) {   // class-probability pass: no aux outputs exist
        launch_one<true, false, 72, true>(a, tiles, nullptr, stream);
        return;
    }
    if (!a.subtile_cull) {   // debugging aid only
        launch_one<false, true, 96>(a, tiles, nullptr, stream);
        return;
    }
    if (a.aux_flag == nullptr || a.variant == 0) {   // caller without a scratch word for the flag, or variant 0
        launch_one<true, true, 96>(a, tiles, nullptr, stream);
        return;
    }
    cudaMemsetAsync(a.aux_flag, 0, sizeof(int), stream);
    aux_zero_scan_kernel<<<148 * 8, 256, 0, stream>>>(a.dL_dothers, a.n_contrib, (size_t)a.W * a.H, a.aux_flag);
    // 2 is the default (api.cu).  Measured at 2M surfels (tools/bench_variants.py, ms for colour+alpha | all gradients):
    // v0 3.14 | 3.49, v1 2.64 | 3.51, v2 2.63 | 3.42, v3 2.98 | 3.51, v4 3.12 | 3.85.
    switch (a.variant) {
    case 1:  launch_one<true, false, 72>(a, tiles, a.aux_flag, stream); launch_one<true, true, 96>(a, tiles, a.aux_flag, stream); break;
    case 3:  launch_one<true, false, 56>(a, tiles, a.aux_flag, stream); launch_one<true, true, 96>(a, tiles, a.aux_flag, stream); break;
    case 4:  launch_one<true, false, 96>(a, tiles, a.aux_flag, stream); launch_one<true, true, 80>(a, tiles, a.aux_flag, stream); break;
    default: launch_one<true, false, 72>(a, tiles, a.aux_flag, stream); launch_one<true, true, 72>(a, tiles, a.aux_flag, stream); break;
    }
}

}  // namespace surfel
