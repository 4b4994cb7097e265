// binning.cu -- screen-tile binning: depth ordering, instance emission, tile sort, tile ranges.
//
// Reference behaviour (RAST/cuda_rasterizer/rasterizer_impl.cu): InclusiveSum of tiles_touched
// (:278), duplicateWithKeys (:70-111) emitting 64-bit keys (tile << 32 | fp32 depth bits), one
// stable 46-bit radix sort of R (key,value) pairs (:301-309), identifyTileRanges (:116-138).
//
// Design here (same resulting order, far less traffic): the (tile, depth, id) order is produced
// by TWO stable LSD sorts instead of one wide one --
//   1. sort the P Gaussians by the 32 depth bits (P keys, not R),
//   2. emit instances in that order with 32-bit tile keys, then stable-sort R pairs by only
//      ceil(log2(tiles)) bits (14 bits at 1920x1280 -> 2 digit passes over 8 B/instance instead of
//      6 passes over 12 B/instance).
// Stability of both sorts keeps ties (equal depth bits) in ascending Gaussian id, exactly the
// order cub's stable sort gives the reference.  The sort and scan themselves are
// hand-written (radix_sort.cu); no library call remains on the path.
#include "common.cuh"
#include "kernels.h"

namespace surfel {

size_t depth_sort_temp_bytes(int P)
{
    const size_t a = radix_sort_temp_bytes(P), b = scan_temp_bytes(P);
    return (a > b ? a : b) + 256;
}

size_t tile_sort_temp_bytes(int64_t R) { return radix_sort_temp_bytes(R) + 256; }

cudaError_t run_depth_order(int P, const uint32_t *depth_key, uint32_t *depth_key_sorted, const uint32_t *idx_in,
                            uint32_t *idx_sorted, const uint32_t *tiles_touched, uint32_t *offsets,
                            int64_t *num_rendered_dev, char *temp, size_t temp_bytes, cudaStream_t stream)
{
    (void)num_rendered_dev;
    cudaError_t e = radix_sort_pairs(depth_key, idx_in, depth_key_sorted, idx_sorted, P, 32, temp, temp_bytes, stream);
    if (e != cudaSuccess) return e;
    return inclusive_scan_gathered(P, tiles_touched, idx_sorted, offsets, temp, temp_bytes, stream);
}

// A window = the tiles [tile_lo, tile_hi) in row-major order (tile id = y * gx + x) that one call bins and blends.
// Single GPU: [0, gx * gy).  Sharded path: the contiguous, cost-balanced tile range of this rank (exchange.cu).
// Row y of a splat's rect [x0, x1) x [y0, y1) contributes its tiles [max(x0, tile_lo - y gx), min(x1, tile_hi - y gx)).
__device__ __forceinline__ void window_row_span(const int y, const int gx, const int x0, const int x1, const int tile_lo,
                                                const int tile_hi, int &xs, int &xe)
{
    xs = max(x0, tile_lo - y * gx);
    xe = min(x1, tile_hi - y * gx);
}

// One thread per depth-ordered Gaussian: write its (tile id, Gaussian id) instances.
__global__ void __launch_bounds__(256)
emit_instances_kernel(const int P, const int gx, const int gy, const int tile_lo, const int tile_hi,
                      const float *__restrict__ rec, const int *__restrict__ radii,
                      const uint32_t *__restrict__ idx_sorted, const uint32_t *__restrict__ offsets,
                      uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint32_t end = offsets[i];
    uint32_t off = (i == 0) ? 0u : offsets[i - 1];
    if (end == off) return;
    const uint32_t g = idx_sorted[i];
    const int r = radii[g];
    const float px = rec[(size_t)g * REC_FLOATS + 9], py = rec[(size_t)g * REC_FLOATS + 10];
    // same rect as preprocess (auxiliary.h:67-77)
    const int x0 = min(gx, max(0, (int)((px - r) / TILE_X)));
    const int y0 = min(gy, max(0, (int)((py - r) / TILE_Y)));
    const int x1 = min(gx, max(0, (int)((px + r + TILE_X - 1) / TILE_X)));
    const int y1 = min(gy, max(0, (int)((py + r + TILE_Y - 1) / TILE_Y)));
    for (int y = y0; y < y1; y++) {
        int xs, xe;
        window_row_span(y, gx, x0, x1, tile_lo, tile_hi, xs, xe);
        for (int x = xs; x < xe; x++) {
            keys[off] = (uint32_t)(y * gx + x);
            vals[off] = g;
            off++;
        }
    }
}

// Tiles of each Gaussian's rect that fall into the window (sharded path; the single-GPU path gets the
// full count from preprocess).
__global__ void __launch_bounds__(256)
count_window_tiles_kernel(const int P, const int gx, const int gy, const int tile_lo, const int tile_hi,
                          const float *__restrict__ rec, const int *__restrict__ radii,
                          uint32_t *__restrict__ tiles_touched, uint32_t *__restrict__ idx_in)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P) return;
    idx_in[g] = (uint32_t)g;
    const int r = radii[g];
    uint32_t n = 0;
    if (r > 0) {
        const float px = rec[(size_t)g * REC_FLOATS + 9], py = rec[(size_t)g * REC_FLOATS + 10];
        const int x0 = min(gx, max(0, (int)((px - r) / TILE_X)));
        const int y0 = min(gy, max(0, (int)((py - r) / TILE_Y)));
        const int x1 = min(gx, max(0, (int)((px + r + TILE_X - 1) / TILE_X)));
        const int y1 = min(gy, max(0, (int)((py + r + TILE_Y - 1) / TILE_Y)));
        for (int y = y0; y < y1; y++) {
            int xs, xe;
            window_row_span(y, gx, x0, x1, tile_lo, tile_hi, xs, xe);
            n += (uint32_t)max(xe - xs, 0);
        }
    }
    tiles_touched[g] = n;
}

void launch_count_window_tiles(int P, int gx, int gy, int tile_lo, int tile_hi, const float *rec,
                               const int *radii, uint32_t *tiles_touched, uint32_t *idx_in, cudaStream_t stream)
{
    if (P == 0) return;
    count_window_tiles_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, gx, gy, tile_lo, tile_hi, rec, radii,
                                                                   tiles_touched, idx_in);
}

__global__ void __launch_bounds__(256)
tile_ranges_kernel(const int64_t R, const uint32_t *__restrict__ keys_sorted, uint2 *__restrict__ ranges)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const uint32_t cur = keys_sorted[i];
    if (i == 0) {
        ranges[cur].x = 0;
    } else {
        const uint32_t prev = keys_sorted[i - 1];
        if (cur != prev) {
            ranges[prev].y = (uint32_t)i;
            ranges[cur].x = (uint32_t)i;
        }
    }
    if (i == R - 1) ranges[cur].y = (uint32_t)R;
}

// Launch order of the tiles of a window: longest instance lists first (buckets of floor(log2(len))),
// so that the few very long lists start at t = 0 instead of forming the tail of the blend kernels.
__global__ void __launch_bounds__(1024)
order_tiles_kernel(const int tile_lo, const int n, const uint2 *__restrict__ ranges, uint32_t *__restrict__ order)
{
    __shared__ uint32_t hist[33], base[33];
    if (threadIdx.x < 33) hist[threadIdx.x] = 0;
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const uint2 r = ranges[tile_lo + j];
        atomicAdd(&hist[32 - __clz(r.y - r.x)], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int b = 32; b >= 0; b--) { base[b] = acc; acc += hist[b]; }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const uint2 r = ranges[tile_lo + j];
        order[atomicAdd(&base[32 - __clz(r.y - r.x)], 1u)] = (uint32_t)(tile_lo + j);
    }
}

cudaError_t run_tile_binning(int P, int64_t R, int gx, int gy, int tile_lo, int tile_hi, const float *rec,
                             const int *radii,
                             const uint32_t *idx_sorted, const uint32_t *offsets, uint32_t *keys_unsorted,
                             uint32_t *vals_unsorted, uint32_t *keys_sorted, uint32_t *point_list, uint2 *ranges,
                             uint32_t *tile_order, char *temp, size_t temp_bytes, cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)gx * gy, stream);
    const int ntiles = tile_hi > tile_lo ? tile_hi - tile_lo : 0;
    if (e != cudaSuccess) return e;
    if (R == 0 || P == 0) {
        if (ntiles > 0) order_tiles_kernel<<<1, 1024, 0, stream>>>(tile_lo, ntiles, ranges, tile_order);
        return cudaGetLastError();
    }
    emit_instances_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, gx, gy, tile_lo, tile_hi, rec, radii,
                                                               idx_sorted, offsets, keys_unsorted, vals_unsorted);
    int bits = 1;
    while ((1u << bits) < (uint32_t)(gx * gy)) bits++;
    e = radix_sort_pairs(keys_unsorted, vals_unsorted, keys_sorted, point_list, R, bits, temp, temp_bytes, stream);
    if (e != cudaSuccess) return e;
    tile_ranges_kernel<<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(R, keys_sorted, ranges);
    if (ntiles > 0) order_tiles_kernel<<<1, 1024, 0, stream>>>(tile_lo, ntiles, ranges, tile_order);
    return cudaGetLastError();
}

}  // namespace surfel
