// radix_sort.cu -- hand-written stable LSD radix sort of (uint32 key, uint32 value) pairs and an
// exclusive/inclusive scan, replacing the two cub calls of the reference
// (cub::DeviceRadixSort::SortPairs, rasterizer_impl.cu:304-309; cub::DeviceScan::InclusiveSum, :278).
//
// 8-bit digits.  One pass = three launches:
//   digit_histogram : per-CTA digit counts of a 2048-item tile          -> hist[digit][cta]
//   digit_row_scan  : exclusive scan of each digit's row of hist, row totals (the scatter kernel
//                     scans the 256 totals itself)
//   digit_scatter   : re-reads the tile, ranks every item among equal digits in ORIGINAL order
//                     (warp: __match_any_sync + popc; CTA: per-warp digit counters in shared memory) and
//                     writes it to hist[digit][cta] + rank.
// Order among equal digits = (cta, warp, round, lane) = input order, so every pass is stable and the
// composition over passes is a stable sort on the selected key bits (what the reference relies on for
// equal (tile, depth) keys).
#include <atomic>

#include "common.cuh"
#include "kernels.h"

namespace surfel {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 8;                                  // items per thread
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;               // 2048 items per CTA
constexpr int RS_WARP_SPAN = 32 * RS_ROUNDS;                  // contiguous items owned by one warp

__global__ void __launch_bounds__(RS_THREADS)
digit_histogram_kernel(const int64_t n, const uint32_t *__restrict__ keys, const int shift, const int nblocks,
                       uint32_t *__restrict__ hist)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const int64_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// Row d of hist (nblocks entries) -> exclusive scan within the row; row total -> totals[d].
__global__ void __launch_bounds__(256)
digit_row_scan_kernel(const int nblocks, uint32_t *__restrict__ hist, uint32_t *__restrict__ totals)
{
    __shared__ uint32_t warp_sum[8];
    __shared__ uint32_t carry_s;
    uint32_t *row = hist + (size_t)blockIdx.x * nblocks;
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
    if (threadIdx.x == 0) totals[blockIdx.x] = carry_s;
}

__global__ void __launch_bounds__(RS_THREADS)
digit_scatter_kernel(const int64_t n, const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                     uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, const int shift,
                     const int nblocks, const uint32_t *__restrict__ hist, const uint32_t *__restrict__ totals)
{
    __shared__ uint32_t wcnt[RS_WARPS][256];   // per-warp running digit counts -> exclusive bases across warps
    __shared__ uint32_t gbase[256];            // global offset of (digit, this CTA)
    __shared__ uint32_t wsum[RS_WARPS];
    __shared__ uint32_t dstart[256];           // start of each digit's run in CTA-sorted order
    __shared__ uint32_t s_k[RS_TILE], s_v[RS_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) wcnt[w][threadIdx.x] = 0;
    {   // exclusive scan of the 256 digit totals (thread = digit): where each digit's run starts
        const uint32_t t = totals[threadIdx.x];
        uint32_t x = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        uint32_t wb = 0;
        for (int w = 0; w < warp; w++) wb += wsum[w];
        gbase[threadIdx.x] = wb + x - t + hist[(size_t)threadIdx.x * nblocks + blockIdx.x];
    }
    __syncthreads();

    // warp w owns the contiguous span [w*256, w*256+256) of the tile, visited 32 items per round
    const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * RS_WARP_SPAN;
    uint32_t k[RS_ROUNDS], v[RS_ROUNDS], rank[RS_ROUNDS];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const int64_t i = wbase + r * 32 + lane;
        const bool ok = i < n;
        k[r] = ok ? keys_in[i] : 0xffffffffu;
        v[r] = ok ? vals_in[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const int64_t i = wbase + r * 32 + lane;
        const bool ok = i < n;
        const uint32_t d = (k[r] >> shift) & 255u;
        // lanes past the end get a private pseudo-digit so they never match a real item
        const uint32_t peers = __match_any_sync(0xffffffffu, ok ? d : (256u + lane));
        const uint32_t before = wcnt[warp][d];
        rank[r] = before + __popc(peers & lt);
        __syncwarp();
        if (ok && (peers & lt) == 0u) wcnt[warp][d] = before + __popc(peers);  // lowest peer lane updates
        __syncwarp();
    }
    __syncthreads();
    // exclusive prefix of the per-warp counts over the warps, per digit (thread = digit); the digit's CTA
    // total then gets an exclusive scan over the 256 digits = start of the digit's run in CTA-sorted order
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            const uint32_t c = wcnt[w][threadIdx.x];
            wcnt[w][threadIdx.x] = run;
            run += c;
        }
        uint32_t x = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        uint32_t wb = 0;
        for (int w = 0; w < warp; w++) wb += wsum[w];
        dstart[threadIdx.x] = wb + x - run;
    }
    __syncthreads();
    // stage the tile in CTA-sorted order in shared memory, then write it out with consecutive threads on
    // consecutive output slots (runs of equal digits are contiguous in global memory -> coalesced stores)
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const int64_t i = wbase + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (k[r] >> shift) & 255u;
            const uint32_t lpos = dstart[d] + wcnt[warp][d] + rank[r];
            s_k[lpos] = k[r];
            s_v[lpos] = v[r];
        }
    }
    __syncthreads();
    const int64_t tile_base = (int64_t)blockIdx.x * RS_TILE;
    const int nvalid = (int)((n - tile_base) < (int64_t)RS_TILE ? (n - tile_base) : (int64_t)RS_TILE);
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const int lp = r * RS_THREADS + threadIdx.x;
        if (lp < nvalid) {
            const uint32_t kk = s_k[lp];
            const uint32_t d = (kk >> shift) & 255u;
            const uint32_t pos = gbase[d] + ((uint32_t)lp - dstart[d]);
            keys_out[pos] = kk;
            vals_out[pos] = s_v[lp];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// One-sweep pass (round 2): the three launches of a pass collapse into one.  A prepass reads the keys ONCE and builds
// the global digit histogram of every pass (the multiset of keys does not change between passes).  The pass kernel
// then ranks its tile exactly like digit_scatter_kernel and obtains "items with my digit in all earlier tiles" by
// decoupled look-back over single-word tile descriptors (flag in the top two bits, count below: one relaxed load sees
// both), instead of reading a [digit][tile] table that a histogram kernel and a row-scan kernel had to produce first.
// Tiles are handed out by an atomic ticket, so a tile only ever waits for tiles that already run: no deadlock, and
// tile order = input order, hence the same stable result.  n < 2^30.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t OS_AGG = 1u << 30, OS_PREFIX = 2u << 30, OS_MASK = (1u << 30) - 1u;

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t *p, const uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(RS_THREADS)
digit_histogram_all_kernel(const int64_t n, const uint32_t *__restrict__ keys, const int passes, uint32_t *__restrict__ ghist)
{
    __shared__ uint32_t h[4][256];
    for (int i = threadIdx.x; i < 4 * 256; i += RS_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * RS_THREADS) {
        const uint32_t k = keys[i];
        for (int p = 0; p < passes; p++) atomicAdd(&h[p][(k >> (8 * p)) & 255u], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += RS_THREADS) {
        const uint32_t v = (&h[0][0])[i];
        if (v) atomicAdd(&ghist[i], v);
    }
}

__global__ void __launch_bounds__(RS_THREADS)
onesweep_pass_kernel(const int64_t n, const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                     uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, const int shift,
                     const uint32_t *__restrict__ ghist, uint32_t *__restrict__ desc, uint32_t *__restrict__ ticket)
{
    __shared__ uint32_t wcnt[RS_WARPS][256];   // per-warp running digit counts -> exclusive bases across warps
    __shared__ uint32_t gbase[256];            // global offset of (digit, this tile)
    __shared__ uint32_t wsum[RS_WARPS];
    __shared__ uint32_t dstart[256];           // start of each digit's run in tile-sorted order
    __shared__ uint32_t s_k[RS_TILE], s_v[RS_TILE];
    __shared__ uint32_t s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) wcnt[w][threadIdx.x] = 0;
    uint32_t digit_start;
    {   // exclusive scan of the 256 global digit totals (thread = digit): where each digit's run starts
        const uint32_t t = ghist[threadIdx.x];
        uint32_t x = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        uint32_t wb = 0;
        for (int w = 0; w < warp; w++) wb += wsum[w];
        digit_start = wb + x - t;
    }
    __syncthreads();
    const uint32_t tile = s_tile;

    const int64_t wbase = (int64_t)tile * RS_TILE + (int64_t)warp * RS_WARP_SPAN;
    uint32_t k[RS_ROUNDS], v[RS_ROUNDS], rank[RS_ROUNDS];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const int64_t i = wbase + r * 32 + lane;
        const bool ok = i < n;
        k[r] = ok ? keys_in[i] : 0xffffffffu;
        v[r] = ok ? vals_in[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const int64_t i = wbase + r * 32 + lane;
        const bool ok = i < n;
        const uint32_t d = (k[r] >> shift) & 255u;
        const uint32_t peers = __match_any_sync(0xffffffffu, ok ? d : (256u + lane));
        const uint32_t before = wcnt[warp][d];
        rank[r] = before + __popc(peers & lt);
        __syncwarp();
        if (ok && (peers & lt) == 0u) wcnt[warp][d] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            const uint32_t c = wcnt[w][threadIdx.x];
            wcnt[w][threadIdx.x] = run;
            run += c;
        }
        // publish this tile's count of digit `threadIdx.x`, then look back over the earlier tiles
        uint32_t *mine = desc + (size_t)tile * 256 + threadIdx.x;
        st_relaxed(mine, (tile == 0 ? OS_PREFIX : OS_AGG) | run);
        uint32_t excl = 0;
        if (tile > 0) {
            // eight predecessors per round trip (independent loads); a descriptor that is not published yet ends the
            // batch and is asked for again
            int64_t p = (int64_t)tile - 1;
            bool done = false;
            while (!done) {
                uint32_t d[8];
#pragma unroll
                for (int u = 0; u < 8; u++)
                    d[u] = (p - u >= 0) ? ld_relaxed(desc + (size_t)(p - u) * 256 + threadIdx.x) : OS_PREFIX;
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    if (done) break;
                    if ((d[u] >> 30) == 0u) break;          // not there yet: retry from this tile
                    excl += d[u] & OS_MASK;
                    p--;
                    if ((d[u] >> 30) == 2u) done = true;
                }
            }
            st_relaxed(mine, OS_PREFIX | (excl + run));
        }
        gbase[threadIdx.x] = digit_start + excl;
        uint32_t x = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        uint32_t wb = 0;
        for (int w = 0; w < warp; w++) wb += wsum[w];
        dstart[threadIdx.x] = wb + x - run;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const int64_t i = wbase + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (k[r] >> shift) & 255u;
            const uint32_t lpos = dstart[d] + wcnt[warp][d] + rank[r];
            s_k[lpos] = k[r];
            s_v[lpos] = v[r];
        }
    }
    __syncthreads();
    const int64_t tile_base = (int64_t)tile * RS_TILE;
    const int nvalid = (int)((n - tile_base) < (int64_t)RS_TILE ? (n - tile_base) : (int64_t)RS_TILE);
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const int lp = r * RS_THREADS + threadIdx.x;
        if (lp < nvalid) {
            const uint32_t kk = s_k[lp];
            const uint32_t d = (kk >> shift) & 255u;
            const uint32_t pos = gbase[d] + ((uint32_t)lp - dstart[d]);
            keys_out[pos] = kk;
            vals_out[pos] = s_v[lp];
        }
    }
}

std::atomic<int> g_radix_onesweep{1};   // surfel_set_option("radix_onesweep", 0): three-launch passes (debugging / tests)
void set_radix_onesweep(int v) { g_radix_onesweep.store(v, std::memory_order_relaxed); }

size_t radix_sort_temp_bytes(int64_t n)
{
    const int64_t nblocks = (n + RS_TILE - 1) / RS_TILE;
    // hist [256][nblocks] (or tile descriptors [nblocks][256]) + totals [256] (or the ticket) + global histograms
    // [4][256] + ping-pong keys/vals for the intermediate passes
    return (size_t)(256 * (nblocks > 0 ? nblocks : 1) + 256 + 4 * 256) * sizeof(uint32_t) +
           2 * (size_t)(n > 0 ? n : 1) * sizeof(uint32_t) + 1024;
}

// Stable sort of n pairs on key bits [0, end_bit).  Result in (keys_out, vals_out).  The inputs are
// preserved when the number of passes is 1; otherwise intermediate passes ping-pong between the
// outputs and scratch inside `temp`.
cudaError_t radix_sort_pairs(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out,
                             int64_t n, int end_bit, char *temp, size_t temp_bytes, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    if (temp_bytes < radix_sort_temp_bytes(n)) return cudaErrorInvalidValue;
    const int nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
    const int passes = (end_bit + 7) / 8;
    char *p = temp;
    uint32_t *hist = carve<uint32_t>(p, (size_t)256 * nblocks);
    uint32_t *totals = carve<uint32_t>(p, 256);
    uint32_t *ghist = carve<uint32_t>(p, 4 * 256);
    uint32_t *keys_tmp = carve<uint32_t>(p, (size_t)n);
    uint32_t *vals_tmp = carve<uint32_t>(p, (size_t)n);
    const bool onesweep = g_radix_onesweep.load(std::memory_order_relaxed) != 0 && n < (int64_t(1) << 30) && passes <= 4;
    if (onesweep) {
        cudaError_t e = cudaMemsetAsync(ghist, 0, 4 * 256 * sizeof(uint32_t), stream);
        if (e != cudaSuccess) return e;
        const int grid = nblocks < 148 * 4 ? nblocks : 148 * 4;
        digit_histogram_all_kernel<<<grid, RS_THREADS, 0, stream>>>(n, keys_in, passes, ghist);
    }
    // choose the ping-pong so that the LAST pass writes (keys_out, vals_out)
    const uint32_t *src_k = keys_in, *src_v = vals_in;
    for (int pass = 0; pass < passes; pass++) {
        const bool to_out = ((passes - 1 - pass) % 2) == 0;
        uint32_t *dst_k = to_out ? keys_out : keys_tmp, *dst_v = to_out ? vals_out : vals_tmp;
        const int shift = 8 * pass;
        if (onesweep) {
            // descriptors [nblocks][256] and the ticket (first word of `totals`, adjacent) are cleared per pass
            cudaError_t e = cudaMemsetAsync(hist, 0, (size_t)((char *)(totals + 256) - (char *)hist), stream);
            if (e != cudaSuccess) return e;
            onesweep_pass_kernel<<<nblocks, RS_THREADS, 0, stream>>>(n, src_k, src_v, dst_k, dst_v, shift, ghist + 256 * pass,
                                                                     hist, totals);
        } else {
            digit_histogram_kernel<<<nblocks, RS_THREADS, 0, stream>>>(n, src_k, shift, nblocks, hist);
            digit_row_scan_kernel<<<256, 256, 0, stream>>>(nblocks, hist, totals);
            digit_scatter_kernel<<<nblocks, RS_THREADS, 0, stream>>>(n, src_k, src_v, dst_k, dst_v, shift, nblocks, hist, totals);
        }
        src_k = dst_k;
        src_v = dst_v;
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Inclusive scan of tiles_touched[idx_sorted[i]] (three launches: CTA sums, scan of sums, CTA scan)
// ---------------------------------------------------------------------------------------------
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(const uint32_t v, uint32_t *warp_sum, uint32_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    uint32_t wbase = 0, t = 0;
#pragma unroll
    for (int w = 0; w < SC_THREADS / 32; w++) {
        const uint32_t s = warp_sum[w];
        if (w < warp) wbase += s;
        t += s;
    }
    total = t;
    __syncthreads();
    return wbase + x - v;
}

__global__ void __launch_bounds__(SC_THREADS)
scan_block_sums_kernel(const int n, const uint32_t *__restrict__ tiles_touched, const uint32_t *__restrict__ idx_sorted,
                       uint32_t *__restrict__ block_sums)
{
    __shared__ uint32_t warp_sum[SC_THREADS / 32];
    const int base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < SC_ITEMS; j++)
        if (base + j < n) s += tiles_touched[idx_sorted[base + j]];
    uint32_t total;
    block_exclusive_scan(s, warp_sum, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SC_THREADS)
scan_of_sums_kernel(const int nblocks, uint32_t *__restrict__ block_sums)
{
    __shared__ uint32_t warp_sum[SC_THREADS / 32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nblocks; b0 += SC_THREADS) {
        const int i = b0 + threadIdx.x;
        const uint32_t v = i < nblocks ? block_sums[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, warp_sum, total);
        const uint32_t carry = carry_s;
        if (i < nblocks) block_sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SC_THREADS)
scan_final_kernel(const int n, const uint32_t *__restrict__ tiles_touched, const uint32_t *__restrict__ idx_sorted,
                  const uint32_t *__restrict__ block_sums, uint32_t *__restrict__ offsets)
{
    __shared__ uint32_t warp_sum[SC_THREADS / 32];
    const int base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
    uint32_t v[SC_ITEMS], s = 0;
#pragma unroll
    for (int j = 0; j < SC_ITEMS; j++) {
        v[j] = (base + j < n) ? tiles_touched[idx_sorted[base + j]] : 0u;
        s += v[j];
    }
    uint32_t total;
    uint32_t run = block_sums[blockIdx.x] + block_exclusive_scan(s, warp_sum, total);
#pragma unroll
    for (int j = 0; j < SC_ITEMS; j++) {
        run += v[j];
        if (base + j < n) offsets[base + j] = run;  // inclusive
    }
}

size_t scan_temp_bytes(int n) { return (size_t)((n + SC_TILE - 1) / SC_TILE + 1) * sizeof(uint32_t) + 256; }

cudaError_t inclusive_scan_gathered(int n, const uint32_t *tiles_touched, const uint32_t *idx_sorted, uint32_t *offsets,
                                    char *temp, size_t temp_bytes, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    if (temp_bytes < scan_temp_bytes(n)) return cudaErrorInvalidValue;
    const int nblocks = (n + SC_TILE - 1) / SC_TILE;
    char *p = temp;
    uint32_t *block_sums = carve<uint32_t>(p, (size_t)nblocks);
    scan_block_sums_kernel<<<nblocks, SC_THREADS, 0, stream>>>(n, tiles_touched, idx_sorted, block_sums);
    scan_of_sums_kernel<<<1, SC_THREADS, 0, stream>>>(nblocks, block_sums);
    scan_final_kernel<<<nblocks, SC_THREADS, 0, stream>>>(n, tiles_touched, idx_sorted, block_sums, offsets);
    return cudaGetLastError();
}

}  // namespace surfel
