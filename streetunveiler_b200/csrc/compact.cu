// compact.cu -- order-preserving compaction of a shard's visible Gaussians (sharded path only).
//
// Only Gaussians that survive culling (radius > 0) have to be exchanged between the ranks: their
// 96-B records, radii and depth keys on the way out, their 80-B gradient records on the way back.
// With the index order preserved, ties between equal depth keys resolve exactly as in the dense
// layout, so the images stay bit-identical to the single-GPU path.
//
//   K_c1  per-block count of visible Gaussians        (1024 per block)
//   K_c2  exclusive scan of the block counts          (one block)
//   K_c3  in-block ranks + gather of the rows, slot[i] = compact row of Gaussian i (or 0xFFFFFFFF);
//         rows [count, P) of radii / keys are filled with the "culled" pattern (0 / 0xFFFFFFFF)
#include "common.cuh"
#include "kernels.h"

namespace surfel {

constexpr int CP_THREADS = 256;
constexpr int CP_ITEMS = 4;
constexpr int CP_TILE = CP_THREADS * CP_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums, uint32_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < CP_THREADS / 32; ++w) {
        const uint32_t s = warp_sums[w];
        if (w < warp) base += s;
        tot += s;
    }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(CP_THREADS)
compact_count_kernel(const int P, const int *__restrict__ radii, uint32_t *__restrict__ block_counts)
{
    __shared__ uint32_t warp_sums[CP_THREADS / 32];
    const int base = blockIdx.x * CP_TILE + threadIdx.x * CP_ITEMS;
    uint32_t n = 0;
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k)
        if (base + k < P && radii[base + k] > 0) ++n;
    uint32_t total;
    block_exclusive_scan(n, warp_sums, total);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(CP_THREADS)
compact_scan_counts_kernel(const int nblocks, uint32_t *__restrict__ block_counts, int *__restrict__ count_out)
{
    __shared__ uint32_t warp_sums[CP_THREADS / 32];
    uint32_t carry = 0;
    for (int start = 0; start < nblocks; start += CP_THREADS) {
        const int i = start + threadIdx.x;
        const uint32_t v = i < nblocks ? block_counts[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, warp_sums, total);
        if (i < nblocks) block_counts[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *count_out = (int)carry;
}

__global__ void __launch_bounds__(CP_THREADS)
compact_gather_kernel(const int P, const int *__restrict__ radii, const float *__restrict__ rec,
                      const uint32_t *__restrict__ keys, const uint32_t *__restrict__ block_base,
                      const int *__restrict__ count, float *__restrict__ rec_c, int *__restrict__ radii_c,
                      uint32_t *__restrict__ keys_c, uint32_t *__restrict__ slot)
{
    __shared__ uint32_t warp_sums[CP_THREADS / 32];
    __shared__ uint32_t src_of[CP_TILE];  // source Gaussian of the block's j-th visible one
    const int first = blockIdx.x * CP_TILE + threadIdx.x * CP_ITEMS;
    bool vis[CP_ITEMS];
    uint32_t n = 0;
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k) {
        vis[k] = first + k < P && radii[first + k] > 0;
        n += vis[k];
    }
    uint32_t total;
    uint32_t r = block_exclusive_scan(n, warp_sums, total);
    const uint32_t base = block_base[blockIdx.x];
    const int cnt = *count;
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k) {
        const int i = first + k;
        if (i >= P) break;
        if (vis[k]) {
            src_of[r] = (uint32_t)i;
            slot[i] = base + r;
            radii_c[base + r] = radii[i];
            keys_c[base + r] = keys[i];
            ++r;
        } else {
            slot[i] = 0xFFFFFFFFu;
        }
        if (i >= cnt) {  // tail rows of the compact arrays: "culled" pattern, never a compaction target
            radii_c[i] = 0;
            keys_c[i] = 0xFFFFFFFFu;
        }
    }
    __syncthreads();
    // coalesced copy of the records: six float4 per row, consecutive threads on consecutive float4
    const float4 *src = reinterpret_cast<const float4 *>(rec);
    float4 *dst = reinterpret_cast<float4 *>(rec_c);
    constexpr int Q = REC_FLOATS / 4;
    for (uint32_t t = threadIdx.x; t < total * Q; t += CP_THREADS) {
        const uint32_t j = t / Q, q = t % Q;
        dst[(size_t)(base + j) * Q + q] = src[(size_t)src_of[j] * Q + q];
    }
}

size_t compact_temp_bytes(int P) { return (size_t)((P + CP_TILE - 1) / CP_TILE + 1) * sizeof(uint32_t) + 256; }

cudaError_t run_compact_visible(int P, const int *radii, const float *rec, const uint32_t *keys, float *rec_c,
                                int *radii_c, uint32_t *keys_c, uint32_t *slot, int *count_dev, char *temp,
                                size_t temp_bytes, cudaStream_t stream)
{
    if (P <= 0) return cudaSuccess;
    if (temp_bytes < compact_temp_bytes(P)) return cudaErrorInvalidValue;
    const int nblocks = (P + CP_TILE - 1) / CP_TILE;
    char *p = temp;
    uint32_t *block_counts = carve<uint32_t>(p, (size_t)nblocks);
    compact_count_kernel<<<nblocks, CP_THREADS, 0, stream>>>(P, radii, block_counts);
    compact_scan_counts_kernel<<<1, CP_THREADS, 0, stream>>>(nblocks, block_counts, count_dev);
    compact_gather_kernel<<<nblocks, CP_THREADS, 0, stream>>>(P, radii, rec, keys, block_counts, count_dev, rec_c,
                                                              radii_c, keys_c, slot);
    return cudaGetLastError();
}

}  // namespace surfel
