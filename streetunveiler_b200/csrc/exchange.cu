// exchange.cu -- device side of the multi-GPU path (DESIGN.md "Multi-GPU"): Gaussians are sharded by index, the
// SCREEN is partitioned into G contiguous, cost-balanced ranges of row-major tile ids, and every projected record
// travels only to the ranks whose range its tile rect touches.
//
// The reference is single-GPU (/root/reference/utils/general_utils.py:133); what must be reproduced is the result of
// its ONE binning + blend (rasterizer_impl.cu:278-341): per tile the same list, in the same (depth, id) order.  That
// fixes the rules here: destinations receive records in global index order (rank-major, then local index -- the
// scatter below is a stable compaction per destination), so that the receiver's stable depth sort resolves equal
// depth keys exactly like the single-GPU sort does.
//
//   tile_hist      per-tile instance counts of the local shard          (all-reduced by the caller -> global histogram)
//   partition      prefix sum of (count + cost_base) over the tiles, cut into G equal-cost contiguous ranges;
//                  also the exact number of instances of every range (= that rank's num_rendered, no read-back later)
//   route_count    per Gaussian the set of ranks its rect touches; per-destination stable compaction positions
//                  (per-CTA counts -> scan) and the send counts
//   route_scatter  packed 112-B rows (record + depth key + radius) into per-destination segments of the send buffer
//   unpack         received rows -> 96-B records (16-B aligned, one cp.async.bulk each), keys, radii
//   grad_accumulate  returned 80-B gradient rows summed into the owner's per-Gaussian accumulator
#include "async_copy.cuh"
#include "common.cuh"
#include "kernels.h"

namespace surfel {

constexpr int XR_THREADS = 256;

__device__ __forceinline__ bool splat_rect(const float *__restrict__ rec, const int *__restrict__ radii, const int g,
                                           const int gx, const int gy, int &x0, int &y0, int &x1, int &y1)
{
    const int r = radii[g];
    if (r <= 0) return false;
    const float px = rec[(size_t)g * REC_FLOATS + 9], py = rec[(size_t)g * REC_FLOATS + 10];
    x0 = min(gx, max(0, (int)((px - r) / TILE_X)));   // same rect as preprocess (auxiliary.h:67-77)
    y0 = min(gy, max(0, (int)((py - r) / TILE_Y)));
    x1 = min(gx, max(0, (int)((px + r + TILE_X - 1) / TILE_X)));
    y1 = min(gy, max(0, (int)((py + r + TILE_Y - 1) / TILE_Y)));
    return x1 > x0 && y1 > y0;
}

// Global-atomics version: only for tile grids too large for shared memory (> ~56k tiles).
__global__ void __launch_bounds__(XR_THREADS)
tile_hist_kernel(const int P, const int gx, const int gy, const float *__restrict__ rec, const int *__restrict__ radii,
                 uint32_t *__restrict__ hist)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P) return;
    int x0, y0, x1, y1;
    if (!splat_rect(rec, radii, g, gx, gy, x0, y0, x1, y1)) return;
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) atomicAdd(&hist[y * gx + x], 1u);
}

// Persistent CTAs, each with a private histogram of the whole tile grid in shared memory (38 KB at 1920x1280):
// Gaussians arrive in random screen order, so global atomics would serialise on the few dense tiles (measured 2.3 ms
// for 4M surfels); shared-memory atomics spread over many CTAs do not, and the flush adds each CTA's non-zero bins once.
__global__ void __launch_bounds__(XR_THREADS)
tile_hist_smem_kernel(const int P, const int gx, const int gy, const float *__restrict__ rec,
                      const int *__restrict__ radii, uint32_t *__restrict__ hist)
{
    extern __shared__ uint32_t sh[];
    const int ntiles = gx * gy;
    for (int t = threadIdx.x; t < ntiles; t += blockDim.x) sh[t] = 0;
    __syncthreads();
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < P; g += gridDim.x * blockDim.x) {
        int x0, y0, x1, y1;
        if (!splat_rect(rec, radii, g, gx, gy, x0, y0, x1, y1)) continue;
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) atomicAdd(&sh[y * gx + x], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < ntiles; t += blockDim.x) {
        const uint32_t v = sh[t];
        if (v) atomicAdd(&hist[t], v);
    }
}

void launch_tile_hist(int P, int gx, int gy, const float *rec, const int *radii, uint32_t *hist, cudaStream_t stream)
{
    cudaMemsetAsync(hist, 0, sizeof(uint32_t) * (size_t)gx * gy, stream);
    if (P <= 0) return;
    const size_t smem = sizeof(uint32_t) * (size_t)gx * gy;
    if (smem <= 200 * 1024) {
        cudaFuncSetAttribute(tile_hist_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int per_sm = smem > 100 * 1024 ? 1 : (smem > 64 * 1024 ? 2 : 4);
        const int grid = min((P + XR_THREADS - 1) / XR_THREADS, 148 * per_sm);
        tile_hist_smem_kernel<<<grid, XR_THREADS, smem, stream>>>(P, gx, gy, rec, radii, hist);
    } else {
        tile_hist_kernel<<<(P + XR_THREADS - 1) / XR_THREADS, XR_THREADS, 0, stream>>>(P, gx, gy, rec, radii, hist);
    }
}

// One CTA.  cost(t) = hist[t] + cost_base; rank k is meant to receive the share shares[k] of the total cost
// (targets.t[k] = round(2^32 * sum_{j<k} shares[j]); equal shares when the caller passes none).  Tile t goes to rank
// #{k >= 1 : E(t) / total >= targets[k]}, E = exclusive prefix of cost: monotone in t, hence contiguous ranges.
// cuts[k] = first tile of rank k (cuts[0] = 0, cuts[G] = ntiles); window_R[k] = instances inside range k (the exact
// num_rendered of rank k's window: no read-back later).
// scratch: ntiles ints (rank of every tile) + ntiles u64 (exclusive prefix of the instance counts).
struct PartitionTargets {
    unsigned long long t[MAX_RANKS + 1];   // fixed point, 2^32 = everything
};

__global__ void __launch_bounds__(1024)
partition_kernel(const int ntiles, const int G, const uint32_t *__restrict__ hist, const uint32_t cost_base,
                 const PartitionTargets targets, int *__restrict__ rank_of, unsigned long long *__restrict__ inst_ex,
                 int *__restrict__ cuts, long long *__restrict__ window_R)
{
    __shared__ unsigned long long wc[32], wi[32];
    __shared__ unsigned long long carry_cost, carry_inst, total_cost;
    __shared__ unsigned long long inst_at_cut[MAX_RANKS + 1];
    __shared__ unsigned long long bound[MAX_RANKS + 1];   // rank >= k  <=>  E >= bound[k]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long local = 0;
    for (int t = threadIdx.x; t < ntiles; t += blockDim.x) local += (unsigned long long)hist[t] + cost_base;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (lane == 0) wc[warp] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 32; w++) t += wc[w];
        total_cost = t > 0 ? t : 1;
        carry_cost = 0;
        carry_inst = 0;
        // bound[k] = ceil(total * targets[k] / 2^32), in 128-bit arithmetic (total < 2^45 for any real frame)
        for (int k = 0; k <= G; k++) {
            const unsigned long long hi = __umul64hi(total_cost, targets.t[k]), lo = total_cost * targets.t[k];
            bound[k] = (hi << 32) | (lo >> 32);
            if (lo & 0xffffffffull) bound[k] += 1;
        }
    }
    __syncthreads();
    for (int t0 = 0; t0 < ntiles; t0 += blockDim.x) {   // blocked exclusive scans of cost and of the instance counts
        const int t = t0 + threadIdx.x;
        const unsigned long long h = t < ntiles ? hist[t] : 0ull;
        const unsigned long long c = t < ntiles ? h + cost_base : 0ull;
        unsigned long long xc = c, xi = h;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long yc = __shfl_up_sync(0xffffffffu, xc, o), yi = __shfl_up_sync(0xffffffffu, xi, o);
            if (lane >= o) { xc += yc; xi += yi; }
        }
        if (lane == 31) { wc[warp] = xc; wi[warp] = xi; }
        __syncthreads();
        unsigned long long bc = carry_cost, bi = carry_inst;
        for (int w = 0; w < warp; w++) { bc += wc[w]; bi += wi[w]; }
        if (t < ntiles) {
            const unsigned long long Ec = bc + xc - c;
            int rk = 0;
            for (int k = 1; k < G; k++) rk += (Ec >= bound[k]) ? 1 : 0;   // bounds are non-decreasing in k
            rank_of[t] = rk;
            inst_ex[t] = bi + xi - h;
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) { carry_cost = bc + xc; carry_inst = bi + xi; }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < ntiles; t += blockDim.x) {
        const int rk = rank_of[t], rp = t > 0 ? rank_of[t - 1] : 0;
        for (int k = rp + 1; k <= rk; k++) { cuts[k] = t; inst_at_cut[k] = inst_ex[t]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int first_k = ntiles > 0 ? rank_of[0] : 0, last_k = ntiles > 0 ? rank_of[ntiles - 1] : 0;
        for (int k = 0; k <= first_k; k++) { cuts[k] = 0; inst_at_cut[k] = 0; }   // leading ranks with a zero share
        for (int k = last_k + 1; k <= G; k++) { cuts[k] = ntiles; inst_at_cut[k] = carry_inst; }   // ranks without tiles
        for (int k = 0; k < G; k++) window_R[k] = (long long)(inst_at_cut[k + 1] - inst_at_cut[k]);
    }
}

size_t partition_temp_bytes(int ntiles) { return (size_t)(ntiles > 0 ? ntiles : 1) * 12 + 512; }

void launch_partition(int ntiles, int G, const uint32_t *hist, uint32_t cost_base, const float *shares_host,
                      char *temp, int *cuts, long long *window_R, cudaStream_t stream)
{
    char *p = temp;
    unsigned long long *inst_ex = carve<unsigned long long>(p, (size_t)(ntiles > 0 ? ntiles : 1));
    int *rank_of = carve<int>(p, (size_t)(ntiles > 0 ? ntiles : 1));
    PartitionTargets tg;
    double total = 0.0, acc = 0.0;
    for (int k = 0; k < G; k++) total += shares_host ? (double)(shares_host[k] > 0.f ? shares_host[k] : 0.f) : 1.0;
    if (!(total > 0.0)) { shares_host = nullptr; total = (double)G; }
    for (int k = 0; k <= MAX_RANKS; k++) tg.t[k] = 1ull << 32;
    for (int k = 0; k < G; k++) {
        tg.t[k] = (unsigned long long)(acc / total * 4294967296.0 + 0.5);
        acc += shares_host ? (double)(shares_host[k] > 0.f ? shares_host[k] : 0.f) : 1.0;
    }
    tg.t[0] = 0;
    partition_kernel<<<1, 1024, 0, stream>>>(ntiles, G, hist, cost_base, tg, rank_of, inst_ex, cuts, window_R);
}

// ---------------------------------------------------------------------------------------------------------------
// routing
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int rank_of_tile(const int t, const int *cuts, const int G)
{
    int r = 0;
    while (r + 1 < G && t >= cuts[r + 1]) r++;
    return r;
}

// bit d set <=> some tile of the rect lies in [cuts[d], cuts[d+1])
__device__ __forceinline__ uint32_t dest_mask(const int x0, const int y0, const int x1, const int y1, const int gx,
                                              const int *cuts, const int G)
{
    uint32_t m = 0;
    for (int y = y0; y < y1; y++) {
        const int a = rank_of_tile(y * gx + x0, cuts, G), b = rank_of_tile(y * gx + x1 - 1, cuts, G);
        for (int d = a; d <= b; d++)
            if (cuts[d + 1] > cuts[d]) m |= 1u << d;   // ranks in between own a piece of this row segment by contiguity
    }
    return m;
}

// mask[g] and per-CTA per-destination counts: block_counts[d * nblocks + b]
__global__ void __launch_bounds__(XR_THREADS)
route_count_kernel(const int P, const int gx, const int gy, const int G, const float *__restrict__ rec,
                   const int *__restrict__ radii, const int *__restrict__ cuts_g, uint32_t *__restrict__ mask,
                   uint32_t *__restrict__ block_counts, const int nblocks)
{
    __shared__ int cuts[MAX_RANKS + 1];
    __shared__ uint32_t cnt[MAX_RANKS];
    if (threadIdx.x <= G) cuts[threadIdx.x] = cuts_g[threadIdx.x];
    if (threadIdx.x < MAX_RANKS) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t m = 0;
    if (g < P) {
        int x0, y0, x1, y1;
        if (splat_rect(rec, radii, g, gx, gy, x0, y0, x1, y1)) m = dest_mask(x0, y0, x1, y1, gx, cuts, G);
        mask[g] = m;
    }
    for (int d = 0; d < G; d++) {
        const uint32_t b = __ballot_sync(0xffffffffu, (m >> d) & 1u);
        if ((threadIdx.x & 31) == 0 && b) atomicAdd(&cnt[d], (uint32_t)__popc(b));
    }
    __syncthreads();
    if (threadIdx.x < G) block_counts[(size_t)threadIdx.x * nblocks + blockIdx.x] = cnt[threadIdx.x];
}

// Row d (one CTA): exclusive scan of block_counts[d][*] in place; row total -> send_counts[d].
__global__ void __launch_bounds__(256)
route_scan_kernel(const int nblocks, uint32_t *__restrict__ block_counts, int *__restrict__ send_counts, const int G,
                  const int extra)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) send_counts[G] = extra;   // a caller-defined word that travels with the counts
    __shared__ uint32_t warp_sum[8];
    __shared__ uint32_t carry_s;
    uint32_t *row = block_counts + (size_t)blockIdx.x * nblocks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nblocks; b0 += 256) {
        const int i = b0 + threadIdx.x;
        const uint32_t v = i < nblocks ? row[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        uint32_t wbase = 0;
        for (int w = 0; w < warp; w++) wbase += warp_sum[w];
        const uint32_t carry = carry_s;
        if (i < nblocks) row[i] = carry + wbase + x - v;
        __syncthreads();
        if (threadIdx.x == 255) carry_s = carry + wbase + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) send_counts[blockIdx.x] = (int)carry_s;
}

// Stable scatter: Gaussian g goes to row seg[d] + block_base[d][cta] + (its rank among the CTA's lower-indexed
// Gaussians bound for d) of the send buffer, for every destination d in its mask.
__global__ void __launch_bounds__(XR_THREADS)
route_scatter_kernel(const int P, const int G, const float *__restrict__ rec, const int *__restrict__ radii,
                     const uint32_t *__restrict__ keys, const uint32_t *__restrict__ mask,
                     const uint32_t *__restrict__ block_base, const int nblocks, const int *__restrict__ send_counts,
                     float *__restrict__ send_rows, uint32_t *__restrict__ send_src)
{
    __shared__ uint32_t wcnt[MAX_RANKS][XR_THREADS / 32];
    __shared__ uint32_t seg[MAX_RANKS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t m = g < P ? mask[g] : 0u;
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int d = 0; d < G; d++) { seg[d] = acc; acc += (uint32_t)send_counts[d]; }
    }
    uint32_t before[MAX_RANKS];
#pragma unroll
    for (int d = 0; d < MAX_RANKS; d++) {
        before[d] = 0;
        if (d < G) {
            const uint32_t b = __ballot_sync(0xffffffffu, (m >> d) & 1u);
            before[d] = (uint32_t)__popc(b & ((1u << lane) - 1u));
            if (lane == 0) wcnt[d][warp] = (uint32_t)__popc(b);
        }
    }
    __syncthreads();
    if (m == 0) return;
    float4 q[6];
    const float4 *src = reinterpret_cast<const float4 *>(rec + (size_t)g * REC_FLOATS);
#pragma unroll
    for (int i = 0; i < 6; i++) q[i] = src[i];
    const float4 tail = make_float4(__uint_as_float(keys[g]), __int_as_float(radii[g]), 0.f, 0.f);
#pragma unroll
    for (int d = 0; d < MAX_RANKS; d++) {
        if (d < G && ((m >> d) & 1u)) {
            uint32_t pos = seg[d] + block_base[(size_t)d * nblocks + blockIdx.x] + before[d];
            for (int w = 0; w < warp; w++) pos += wcnt[d][w];
            float4 *dst = reinterpret_cast<float4 *>(send_rows + (size_t)pos * XROW_FLOATS);
#pragma unroll
            for (int i = 0; i < 6; i++) dst[i] = q[i];
            dst[6] = tail;
            send_src[pos] = (uint32_t)g;
        }
    }
}

static void route_carve(char *temp, int P, int G, uint32_t *&mask, uint32_t *&block_counts, int &nblocks);

// Same routing, exchange fused in: rows are stored STRAIGHT into the receive buffers of their destination ranks over
// NVLink (symmetric memory), already split into the three arrays the receiver works on -- no send buffer, no
// all-to-all, no unpack pass.  Destination d receives this rank's rows at [row0[d], row0[d] + send_counts[d]) of its
// buffers (row0 = rows of lower ranks bound for d, from the all-gathered count matrix), so every receiver ends up with
// rank-major, index-minor = global index order, exactly like the all-to-all.  A 96-B record is six 16-B stores to
// consecutive addresses; rows of one CTA land next to each other.
struct PeerRows {
    float *rec[MAX_RANKS];
    uint32_t *keys[MAX_RANKS];
    int *radii[MAX_RANKS];
    long long row0[MAX_RANKS];
};

__global__ void __launch_bounds__(XR_THREADS)
route_scatter_peers_kernel(const int P, const int G, const float *__restrict__ rec, const int *__restrict__ radii,
                           const uint32_t *__restrict__ keys, const uint32_t *__restrict__ mask,
                           const uint32_t *__restrict__ block_base, const int nblocks,
                           const int *__restrict__ send_counts, const PeerRows dst, uint32_t *__restrict__ send_src)
{
    // The CTA's rows for one destination are contiguous there, so they are staged in shared memory in destination
    // order and written out by consecutive threads on consecutive 16-B words: stores that leave the GPU over NVLink are
    // not merged by a local L2, a 96-B-strided pattern would travel as 16-B packets.
    __shared__ __align__(128) float4 s_rec[XR_THREADS * 6];
    __shared__ uint32_t s_key[XR_THREADS];
    __shared__ int s_rad[XR_THREADS];
    __shared__ uint32_t wcnt[MAX_RANKS][XR_THREADS / 32];
    __shared__ uint32_t seg[MAX_RANKS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t m = g < P ? mask[g] : 0u;
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int d = 0; d < G; d++) { seg[d] = acc; acc += (uint32_t)send_counts[d]; }
    }
    uint32_t before[MAX_RANKS];
#pragma unroll
    for (int d = 0; d < MAX_RANKS; d++) {
        before[d] = 0;
        if (d < G) {
            const uint32_t b = __ballot_sync(0xffffffffu, (m >> d) & 1u);
            before[d] = (uint32_t)__popc(b & ((1u << lane) - 1u));
            if (lane == 0) wcnt[d][warp] = (uint32_t)__popc(b);
        }
    }
    __syncthreads();
    float4 q[6];
    uint32_t key = 0;
    int radius = 0;
    if (m) {
        const float4 *src = reinterpret_cast<const float4 *>(rec + (size_t)g * REC_FLOATS);
#pragma unroll
        for (int i = 0; i < 6; i++) q[i] = src[i];
        key = keys[g];
        radius = radii[g];
    }
#pragma unroll
    for (int d = 0; d < MAX_RANKS; d++) {          // unrolled: before[] stays in registers; d < G is CTA-uniform
        if (d >= G) break;
        uint32_t total = 0, local = before[d];
        for (int w = 0; w < XR_THREADS / 32; w++) {
            const uint32_t c = wcnt[d][w];
            if (w < warp) local += c;
            total += c;
        }
        if (total == 0) continue;                       // CTA-uniform
        const uint32_t base = block_base[(size_t)d * nblocks + blockIdx.x];
        if ((m >> d) & 1u) {
#pragma unroll
            for (int i = 0; i < 6; i++) s_rec[local * 6 + i] = q[i];
            s_key[local] = key;
            s_rad[local] = radius;
            send_src[seg[d] + base + local] = (uint32_t)g;
        }
        __syncthreads();
        const size_t row = (size_t)dst.row0[d] + base;
        float *o = dst.rec[d] + row * REC_FLOATS;
        if (threadIdx.x == 0) {
            // one bulk store (TMA engine, shared -> global) moves the CTA's whole block for this destination: a few KB
            // per request on the NVLink instead of 16-B stores
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(o), "r"(smem_addr(s_rec)),
                         "r"(total * (uint32_t)REC_BYTES)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        for (uint32_t w = threadIdx.x; w < total; w += XR_THREADS) {
            dst.keys[d][row + w] = s_key[w];
            dst.radii[d][row + w] = s_rad[w];
        }
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging buffer reusable
        __syncthreads();
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores performed before exit
}

cudaError_t run_route_scatter_peers(int P, int G, const float *rec, const int *radii, const uint32_t *keys, char *temp,
                                    const int *send_counts, float *const *dst_rec, uint32_t *const *dst_keys,
                                    int *const *dst_radii, const long long *dst_row0, uint32_t *send_src,
                                    cudaStream_t stream)
{
    if (P <= 0) return cudaSuccess;
    uint32_t *mask, *block_counts;
    int nblocks;
    route_carve(temp, P, G, mask, block_counts, nblocks);
    PeerRows dst{};
    for (int d = 0; d < G; d++) {
        dst.rec[d] = dst_rec[d]; dst.keys[d] = dst_keys[d]; dst.radii[d] = dst_radii[d]; dst.row0[d] = dst_row0[d];
    }
    route_scatter_peers_kernel<<<nblocks, XR_THREADS, 0, stream>>>(P, G, rec, radii, keys, mask, block_counts, nblocks,
                                                                   send_counts, dst, send_src);
    return cudaGetLastError();
}

// Backward counterpart: the gradient rows of the records this rank received go back to their owners.  Rows
// [seg_start[s], seg_start[s] + seg_count[s]) came from rank s and are stored at row dst_row0[s] + (offset in the
// segment) of rank s's return buffer = the position of that record in rank s's send order.  One 16-B word per thread,
// consecutive threads on consecutive words: full-width NVLink stores.
struct PeerGradRows {
    float *dst[MAX_RANKS];
    long long seg_start[MAX_RANKS + 1];
    long long row0[MAX_RANKS];
};

__global__ void __launch_bounds__(XR_THREADS)
push_grad_rows_kernel(const long long n_words, const int G, const float *__restrict__ rows, const PeerGradRows p)
{
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const long long row = w / 5;
    const int j = (int)(w - row * 5);
    int s = 0;
    while (s + 1 < G && row >= p.seg_start[s + 1]) s++;
    const float4 v = reinterpret_cast<const float4 *>(rows)[w];
    reinterpret_cast<float4 *>(p.dst[s] + (size_t)(p.row0[s] + (row - p.seg_start[s])) * GACC_FLOATS)[j] = v;
}

cudaError_t run_push_grad_rows(long long n_rows, const float *rows, int G, const long long *seg_count,
                               float *const *dst, const long long *dst_row0, cudaStream_t stream)
{
    if (n_rows <= 0) return cudaSuccess;
    PeerGradRows p{};
    long long acc = 0;
    for (int s = 0; s < G; s++) {
        p.dst[s] = dst[s]; p.row0[s] = dst_row0[s]; p.seg_start[s] = acc;
        acc += seg_count[s];
    }
    for (int s = G; s <= MAX_RANKS; s++) p.seg_start[s] = acc;
    if (acc != n_rows) return cudaErrorInvalidValue;
    const long long words = n_rows * 5;
    push_grad_rows_kernel<<<(unsigned)((words + XR_THREADS - 1) / XR_THREADS), XR_THREADS, 0, stream>>>(words, G, rows, p);
    return cudaGetLastError();
}

size_t route_temp_bytes(int P, int G)
{
    const size_t nblocks = (size_t)(P + XR_THREADS - 1) / XR_THREADS + 1;
    return (size_t)(P > 0 ? P : 1) * sizeof(uint32_t) + nblocks * (size_t)G * sizeof(uint32_t) + 512;
}

static void route_carve(char *temp, int P, int G, uint32_t *&mask, uint32_t *&block_counts, int &nblocks)
{
    nblocks = (P + XR_THREADS - 1) / XR_THREADS;
    char *p = temp;
    mask = carve<uint32_t>(p, (size_t)(P > 0 ? P : 1));
    block_counts = carve<uint32_t>(p, (size_t)(nblocks + 1) * G);
}

__global__ void route_empty_kernel(int *send_counts, const int G, const int extra)
{
    for (int d = threadIdx.x; d < G; d += blockDim.x) send_counts[d] = 0;
    if (threadIdx.x == 0) send_counts[G] = extra;
}

cudaError_t run_route_count(int P, int gx, int gy, int G, const float *rec, const int *radii, const int *cuts,
                            char *temp, int *send_counts, int extra, cudaStream_t stream)
{
    if (P <= 0) {
        route_empty_kernel<<<1, 32, 0, stream>>>(send_counts, G, extra);
        return cudaGetLastError();
    }
    uint32_t *mask, *block_counts;
    int nblocks;
    route_carve(temp, P, G, mask, block_counts, nblocks);
    route_count_kernel<<<nblocks, XR_THREADS, 0, stream>>>(P, gx, gy, G, rec, radii, cuts, mask, block_counts, nblocks);
    route_scan_kernel<<<G, 256, 0, stream>>>(nblocks, block_counts, send_counts, G, extra);
    return cudaGetLastError();
}

cudaError_t run_route_scatter(int P, int G, const float *rec, const int *radii, const uint32_t *keys, char *temp,
                              const int *send_counts, float *send_rows, uint32_t *send_src, cudaStream_t stream)
{
    if (P <= 0) return cudaSuccess;
    uint32_t *mask, *block_counts;
    int nblocks;
    route_carve(temp, P, G, mask, block_counts, nblocks);
    route_scatter_kernel<<<nblocks, XR_THREADS, 0, stream>>>(P, G, rec, radii, keys, mask, block_counts, nblocks,
                                                             send_counts, send_rows, send_src);
    return cudaGetLastError();
}

// received 112-B rows -> records / keys / radii (rows past n: none)
__global__ void __launch_bounds__(XR_THREADS)
unpack_rows_kernel(const int n, const float *__restrict__ rows, float *__restrict__ rec, uint32_t *__restrict__ keys,
                   int *__restrict__ radii)
{
    // 7 threads per row: thread j < 6 copies one 16-B word of the record, thread 6 the tail
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long row = tid / 7;
    const int j = (int)(tid - row * 7);
    if (row >= n) return;
    const float4 v = reinterpret_cast<const float4 *>(rows + (size_t)row * XROW_FLOATS)[j];
    if (j < 6) {
        reinterpret_cast<float4 *>(rec + (size_t)row * REC_FLOATS)[j] = v;
    } else {
        keys[row] = __float_as_uint(v.x);
        radii[row] = __float_as_int(v.y);
    }
}

void launch_unpack_rows(int n, const float *rows, float *rec, uint32_t *keys, int *radii, cudaStream_t stream)
{
    if (n <= 0) return;
    const long long threads = (long long)n * 7;
    unpack_rows_kernel<<<(unsigned)((threads + XR_THREADS - 1) / XR_THREADS), XR_THREADS, 0, stream>>>(n, rows, rec, keys, radii);
}

// gacc[src[j]] += rows[j] (five 16-B words per 80-B row; vector RED, fire-and-forget).  A Gaussian has one row per
// destination it was sent to: usually one (plain store semantics), sometimes two (float addition commutes, so the
// result does not depend on the order), rarely more.
__global__ void __launch_bounds__(XR_THREADS)
grad_accumulate_kernel(const long long n, const float *__restrict__ rows, const uint32_t *__restrict__ src,
                       float *__restrict__ gacc)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long row = tid / 5;
    const int j = (int)(tid - row * 5);
    if (row >= n) return;
    const float4 v = reinterpret_cast<const float4 *>(rows + (size_t)row * GACC_FLOATS)[j];
    float *dst = gacc + (size_t)src[row] * GACC_FLOATS + 4 * j;
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

cudaError_t run_grad_accumulate(int P, long long n_rows, const float *rows, const uint32_t *src, float *gacc,
                                cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(gacc, 0, (size_t)(P > 0 ? P : 1) * GACC_FLOATS * sizeof(float), stream);
    if (e != cudaSuccess || n_rows <= 0) return e;
    const long long threads = n_rows * 5;
    grad_accumulate_kernel<<<(unsigned)((threads + XR_THREADS - 1) / XR_THREADS), XR_THREADS, 0, stream>>>(n_rows, rows, src, gacc);
    return cudaGetLastError();
}

}  // namespace surfel
