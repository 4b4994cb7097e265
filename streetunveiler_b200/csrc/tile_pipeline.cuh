// tile_pipeline.cuh -- shared machinery of the two tile-blend kernels (render_fwd.cu / render_bwd.cu).
//
// One CTA per 16x16 tile: 8 consumer warps (each owns an 8x4 pixel block) + 1 producer warp.
// The tile's depth-ordered instance list is streamed through a ring of NSTAGE shared-memory stages
// of CHUNK records each.  The producer warp gathers the 96-B records with the TMA engine
// (cp.async.bulk, completion counted in transaction bytes on the stage's `full` mbarrier); each
// consumer warp releases a stage by arriving on its `empty` mbarrier.  There is no CTA-wide barrier
// in the steady state, so a warp whose pixel block is covered by few splats runs ahead of a busy
// one by up to NSTAGE*CHUNK instances instead of idling at a __syncthreads per batch.
#pragma once
#include "async_copy.cuh"
#include "common.cuh"

namespace surfel {

constexpr int CHUNK = 64;            // records per stage (64 * 96 B = 6 KB)
#ifndef SURFEL_FWD_STAGES
#define SURFEL_FWD_STAGES 6
#endif
constexpr int FWD_STAGES = SURFEL_FWD_STAGES;   // forward: 4 CTAs/SM x 38 KB (deeper rings measured no faster)
constexpr int BWD_STAGES = 8;        // backward: 2 CTAs/SM (register bound) x 50 KB
constexpr int CONSUMER_WARPS = 8;
constexpr int TILE_THREADS = (CONSUMER_WARPS + 1) * 32;

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

template <int NSTAGE>
struct TileRing {
    static constexpr int kStages = NSTAGE;
    float rec[NSTAGE][CHUNK * REC_FLOATS];
    uint32_t id[NSTAGE][CHUNK];
    uint64_t full[NSTAGE];
    uint64_t empty[NSTAGE];
    volatile int done_warps;  // consumer warps whose 32 pixels are all saturated (forward early-out)
    volatile int limit;       // chunks the producer will ever issue (lowered when every warp is done)
};

template <int NSTAGE>
__device__ __forceinline__ void ring_init(TileRing<NSTAGE> &r, const int tid)
{
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; s++) {
            mbar_init(&r.full[s], 32);               // every producer lane arrives (lane 0 with expect_tx)
            mbar_init(&r.empty[s], CONSUMER_WARPS);  // one arrival per consumer warp
        }
        r.done_warps = 0;
        r.limit = 0x7fffffff;
        mbar_fence_init();
    }
    __syncthreads();
}

// Producer warp body.  position(i) maps the i-th streamed slot to an index into point_list
// (front-to-back for the forward, back-to-front for the backward).
template <bool STORE_IDS, int NSTAGE, typename PosFn>
__device__ __forceinline__ void ring_produce(TileRing<NSTAGE> &r, const int lane, const int total, const uint32_t *__restrict__ point_list,
                                             const float *__restrict__ rec, PosFn position)
{
    const int nchunks = (total + CHUNK - 1) / CHUNK;
    constexpr int PER_LANE = CHUNK / 32;
    int stage = 0;
    uint32_t ephase = 1;  // parity of the *previous* phase of `empty`: passes immediately on the first lap
    int issued = 0;
    // First lap: every stage is free, so fetch the ids of all NSTAGE chunks with one batch of
    // independent loads (one exposed memory latency instead of NSTAGE of them) and fire the copies.
    {
        uint32_t ids[NSTAGE * PER_LANE];
        const int first_n = min(total, NSTAGE * CHUNK);
#pragma unroll
        for (int q = 0; q < NSTAGE * PER_LANE; q++) {
            const int slot = (q / PER_LANE) * CHUNK + (q % PER_LANE) * 32 + lane;
            ids[q] = (slot < first_n) ? point_list[position(slot)] : 0u;
        }
#pragma unroll
        for (int c = 0; c < NSTAGE; c++) {
            if (c < nchunks) {
                const int n = min(CHUNK, total - c * CHUNK);
#pragma unroll
                for (int k = 0; k < PER_LANE; k++) {
                    const int s = k * 32 + lane;
                    if (s < n) {
                        const uint32_t g = ids[c * PER_LANE + k];
                        if (STORE_IDS) r.id[c][s] = g;
                        bulk_g2s(&r.rec[c][s * REC_FLOATS], rec + (size_t)g * REC_FLOATS, REC_BYTES, &r.full[c]);
                    }
                }
                if (lane == 0)
                    mbar_arrive_expect_tx(&r.full[c], (uint32_t)n * REC_BYTES);
                else
                    mbar_arrive(&r.full[c]);
                issued = c + 1;
            }
        }
    }
    ephase = 0;
    // Steady state: the ids of chunk c are fetched BEFORE waiting for its stage to drain, so their
    // latency hides behind the wait.
    for (int c = NSTAGE; c < nchunks; c++) {
        int stop = (lane == 0 && r.done_warps >= CONSUMER_WARPS) ? 1 : 0;
        stop = __shfl_sync(0xffffffffu, stop, 0);
        if (stop) break;
        const int n = min(CHUNK, total - c * CHUNK);
        uint32_t ids[PER_LANE];
#pragma unroll
        for (int k = 0; k < PER_LANE; k++) {
            const int s = k * 32 + lane;
            ids[k] = (s < n) ? point_list[position(c * CHUNK + s)] : 0u;
        }
        mbar_wait(&r.empty[stage], ephase);
#pragma unroll
        for (int k = 0; k < PER_LANE; k++) {
            const int s = k * 32 + lane;
            if (s < n) {
                if (STORE_IDS) r.id[stage][s] = ids[k];
                bulk_g2s(&r.rec[stage][s * REC_FLOATS], rec + (size_t)ids[k] * REC_FLOATS, REC_BYTES, &r.full[stage]);
            }
        }
        if (lane == 0)
            mbar_arrive_expect_tx(&r.full[stage], (uint32_t)n * REC_BYTES);
        else
            mbar_arrive(&r.full[stage]);
        issued = c + 1;
        if (++stage == NSTAGE) { stage = 0; ephase ^= 1u; }
    }
    if (lane == 0) {
        r.limit = issued;
        __threadfence_block();
    }
    // bulk copies still in flight must land before the CTA may exit (its smem could be re-assigned)
    const int first = issued > NSTAGE ? issued - NSTAGE : 0;
    for (int c = first; c < issued; c++) mbar_wait(&r.full[c % NSTAGE], (uint32_t)((c / NSTAGE) & 1));
}

// Consumer side: wait until stage data landed.  A warp that is already done only recycles stages so
// the producer can keep feeding the others; it leaves as soon as the producer announces the end.
template <int NSTAGE>
__device__ __forceinline__ bool ring_wait_or_quit(TileRing<NSTAGE> &r, const int lane, const int stage, const uint32_t phase, const int c)
{
    while (true) {
        int st = 0;
        if (lane == 0) st = mbar_try_wait(&r.full[stage], phase) ? 1 : (r.limit <= c ? 2 : 0);
        st = __shfl_sync(0xffffffffu, st, 0);
        if (st == 1) return true;
        if (st == 2) return false;
    }
}

// Conservative test: can any pixel of the 8x4 block at (bx0, by0) reach alpha >= 1/255 for the splat
// whose record is at `rp`?  See preprocess.cu cull_footprint: 8-px bounding box first, then the low-pass
// disc and (when present) the ellipse against the rectangle spanned by the block's pixel centres, with a
// threshold that grows with the fp32 evaluation error of the quadratic (large for distant needles).
__device__ __forceinline__ bool block_may_contribute(const float *rp, const int bx0, const int by0)
{
    const uint32_t bb = __float_as_uint(rp[18]);
    const uint32_t ux = (uint32_t)bx0 >> 3, uy = (uint32_t)by0 >> 3;
    const uint32_t x0 = bb & 255u, x1 = (bb >> 8) & 255u, y0 = (bb >> 16) & 255u, y1 = bb >> 24;
    // a stored 255 is a clamp: it stands for every unit >= 255
    if (ux < x0 || (ux > x1 && x1 != 255u) || uy < y0 || (uy > y1 && y1 != 255u)) return false;
    const float4 m = *reinterpret_cast<const float4 *>(rp + 20);  // e.y, M00, M01, M11
    if (m.y == 0.f) return true;                                   // no ellipse bound: the box decides
    // The alpha test is evaluated at pixel CENTRES, which are the integer coordinates (pixf = (float)pix, the ndc2pix map
    // carries the half-pixel shift): the block's 32 centres span [bx0, bx0+7] x [by0, by0+3], and the minimum of a function
    // over that rectangle bounds its minimum over the centres.  (Round 1 inflated the rectangle by half a pixel on every
    // side; measured on the 2M scene that alone produced half of the non-contributing evaluations: 1.63 -> 1.46 evaluated
    // (warp, instance) pairs per tile instance against 1.30 contributing ones, no missed pair.)  0.01 px of slack remain.
    const float rx0 = (float)bx0 - 0.01f, rx1 = (float)bx0 + 7.01f, ry0 = (float)by0 - 0.01f, ry1 = (float)by0 + 3.01f;
    // low-pass disc |p - mean|^2 <= tau / 2 <= ln(255) (+ the same inflation as preprocess): 5.6 bounds it
    const float mx = rp[9], my = rp[10];
    const float dx = fmaxf(fmaxf(rx0 - mx, mx - rx1), 0.f);
    const float dy = fmaxf(fmaxf(ry0 - my, my - ry1), 0.f);
    if (dx * dx + dy * dy <= 5.6f) return true;
    // ellipse (X-e)^T M (X-e) <= 1: minimum of the convex quadratic over the rectangle
    const float ex = rp[19], ey = m.x, m00 = m.y, m01 = m.z, m11 = m.w;
    const float X0 = rx0 - ex, X1 = rx1 - ex, Y0 = ry0 - ey, Y1 = ry1 - ey;
    if (X0 <= 0.f && X1 >= 0.f && Y0 <= 0.f && Y1 >= 0.f) return true;
    // (approximate reciprocals: the clamped minimiser moves by ~1 ulp, g changes in second order -- far inside the
    // 0.02 + evaluation-error margin of the threshold below; the IEEE version cost two slow-path-checked calls per test)
    float r11, r00;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r11) : "f"(m11));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r00) : "f"(m00));
    const float ky = -m01 * r11, kx = -m01 * r00;
    float gmin;
    {
        const float Ya = fminf(fmaxf(ky * X0, Y0), Y1), Yb = fminf(fmaxf(ky * X1, Y0), Y1);
        const float ga = m00 * X0 * X0 + (2.f * m01 * X0 + m11 * Ya) * Ya;
        const float gb = m00 * X1 * X1 + (2.f * m01 * X1 + m11 * Yb) * Yb;
        gmin = fminf(ga, gb);
    }
    {
        const float Xa = fminf(fmaxf(kx * Y0, X0), X1), Xb = fminf(fmaxf(kx * Y1, X0), X1);
        const float ga = m11 * Y0 * Y0 + (2.f * m01 * Y0 + m00 * Xa) * Xa;
        const float gb = m11 * Y1 * Y1 + (2.f * m01 * Y1 + m00 * Xb) * Xb;
        gmin = fminf(gmin, fminf(ga, gb));
    }
    // fp32 evaluation error of g: a few ulps of its largest term (2|M01 X Y| <= M00 X^2 + M11 Y^2)
    const float ax = fmaxf(fabsf(X0), fabsf(X1)), ay = fmaxf(fabsf(Y0), fabsf(Y1));
    return gmin <= 1.02f + 4e-6f * (m00 * ax * ax + m11 * ay * ay);
}

}  // namespace surfel
