// kernels.h -- host-side launch interfaces between api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

// sets the thread's last-error message (api.cu) and returns a non-zero code
int surfel_internal_fail(const char *where, const char *what);

namespace surfel {

struct PreprocessFwdArgs {
    int P, D, M, W, H, gx, gy, prefiltered;
    float scale_modifier;
    const float *means3D, *scales, *rotations, *opacities, *shs, *transMat_precomp, *colors_precomp;
    const float *viewmatrix, *projmatrix, *cam_pos;
    int *radii;
    float *rec;
    uint32_t *tiles_touched, *depth_key, *idx_in;
    uint8_t *clamped;
};
void launch_preprocess_fwd(const PreprocessFwdArgs &a, cudaStream_t stream);
// colour passes over one geometry state: rewrite the rgb words of the packed records / move the colour gradient out
void launch_set_record_colors(int P, const float *colors, float *rec, cudaStream_t stream);
void launch_take_color_grad(int P, float *gacc, float *dL_dcolor, cudaStream_t stream);
void launch_set_record_labels(int P, const int *labels, float *rec, cudaStream_t stream);
void launch_mark_visible(int P, const float *means3D, const float *viewmatrix, unsigned char *present,
                         cudaStream_t stream);

struct PreprocessBwdArgs {
    int P, D, M;
    float focal_x, focal_y, tan_fovx, tan_fovy;
    const float *means3D, *shs, *scales, *rotations, *transMat_precomp, *viewmatrix, *projmatrix, *cam_pos;
    const int *radii;
    const uint8_t *clamped;
    const float *rec, *gacc;
    const uint32_t *gacc_slot = nullptr;  // optional: row of gacc that belongs to Gaussian i (compact layout)
    float *dL_dmean2D, *dL_dnormal, *dL_dopacity, *dL_dcolor, *dL_dmean3D, *dL_dtransMat, *dL_dsh, *dL_dscale,
        *dL_drot;
};
void launch_preprocess_bwd(const PreprocessBwdArgs &a, cudaStream_t stream);

// ---- radix sort / scan (radix_sort.cu) ----
size_t radix_sort_temp_bytes(int64_t n);
void set_radix_onesweep(int v);
cudaError_t radix_sort_pairs(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out,
                             int64_t n, int end_bit, char *temp, size_t temp_bytes, cudaStream_t stream);
size_t scan_temp_bytes(int n);
cudaError_t inclusive_scan_gathered(int n, const uint32_t *tiles_touched, const uint32_t *idx_sorted, uint32_t *offsets,
                                    char *temp, size_t temp_bytes, cudaStream_t stream);

// ---- visible-row compaction for the sharded path (compact.cu) ----
size_t compact_temp_bytes(int P);
cudaError_t run_compact_visible(int P, const int *radii, const float *rec, const uint32_t *keys, float *rec_c,
                                int *radii_c, uint32_t *keys_c, uint32_t *slot, int *count_dev, char *temp,
                                size_t temp_bytes, cudaStream_t stream);

// ---- multi-GPU exchange (exchange.cu) ----
void launch_tile_hist(int P, int gx, int gy, const float *rec, const int *radii, uint32_t *hist, cudaStream_t stream);
size_t partition_temp_bytes(int ntiles);
void launch_partition(int ntiles, int G, const uint32_t *hist, uint32_t cost_base, const float *shares_host,
                      char *temp, int *cuts, long long *window_R, cudaStream_t stream);
size_t route_temp_bytes(int P, int G);
cudaError_t run_route_count(int P, int gx, int gy, int G, const float *rec, const int *radii, const int *cuts,
                            char *temp, int *send_counts, int extra, cudaStream_t stream);
cudaError_t run_route_scatter(int P, int G, const float *rec, const int *radii, const uint32_t *keys, char *temp,
                              const int *send_counts, float *send_rows, uint32_t *send_src, cudaStream_t stream);
cudaError_t run_route_scatter_peers(int P, int G, const float *rec, const int *radii, const uint32_t *keys, char *temp,
                                    const int *send_counts, float *const *dst_rec, uint32_t *const *dst_keys,
                                    int *const *dst_radii, const long long *dst_row0, uint32_t *send_src,
                                    cudaStream_t stream);
cudaError_t run_push_grad_rows(long long n_rows, const float *rows, int G, const long long *seg_count,
                               float *const *dst, const long long *dst_row0, cudaStream_t stream);
void launch_unpack_rows(int n, const float *rows, float *rec, uint32_t *keys, int *radii, cudaStream_t stream);
cudaError_t run_grad_accumulate(int P, long long n_rows, const float *rows, const uint32_t *src, float *gacc,
                                cudaStream_t stream);

// ---- binning (binning.cu) ----
size_t depth_sort_temp_bytes(int P);
size_t tile_sort_temp_bytes(int64_t R);
// depth-order the Gaussians, scan tiles_touched in that order, leave R in *num_rendered_dev
cudaError_t run_depth_order(int P, const uint32_t *depth_key, uint32_t *depth_key_sorted, const uint32_t *idx_in,
                            uint32_t *idx_sorted, const uint32_t *tiles_touched, uint32_t *offsets,
                            int64_t *num_rendered_dev, char *temp, size_t temp_bytes, cudaStream_t stream);
// emit (tile, id) pairs in depth order, stable-sort by tile, find per-tile ranges
void launch_count_window_tiles(int P, int gx, int gy, int tile_lo, int tile_hi, const float *rec,
                               const int *radii, uint32_t *tiles_touched, uint32_t *idx_in, cudaStream_t stream);
cudaError_t run_tile_binning(int P, int64_t R, int gx, int gy, int tile_lo, int tile_hi, const float *rec,
                             const int *radii,
                             const uint32_t *idx_sorted, const uint32_t *offsets, uint32_t *keys_unsorted,
                             uint32_t *vals_unsorted, uint32_t *keys_sorted, uint32_t *point_list, uint2 *ranges,
                             uint32_t *tile_order, char *temp, size_t temp_bytes, cudaStream_t stream);

// ---- render (render_fwd.cu / render_bwd.cu) ----
struct RenderFwdArgs {
    int W, H, gx, gy;
    int tile_lo = 0, tile_hi = -1;       // window [tile_lo, tile_hi) of row-major tile ids rendered by this call (-1: all)
    const uint2 *ranges;
    const uint32_t *tile_order;  // launch order of the window's tiles (longest lists first); entries are tile ids
    const uint32_t *point_list;
    const float *rec;
    const float *bg;
    float *final_T;
    uint32_t *n_contrib;
    uint32_t *tile_max_contrib;
    float *out_color, *out_others;
    int subtile_cull;
    int n_classes = 0;  // > 0: class-probability pass (out_color = [n_classes,H,W], bg = [n_classes], out_others unused)
    // multi-GPU: store the window's pixels into these [10,H,W] plane buffers (one per rank, or ONE multicast address)
    // instead of out_color / out_others
    int n_peers = 0, peer_multicast = 0;
    float *peer_planes[16] = {};
};
void launch_render_fwd(const RenderFwdArgs &a, cudaStream_t stream);

struct RenderBwdArgs {
    int W, H, gx, gy;
    int tile_lo = 0, tile_hi = -1;
    const uint2 *ranges;
    const uint32_t *tile_order;
    const uint32_t *point_list;
    const float *rec;
    const float *bg;
    const float *final_T;
    const uint32_t *n_contrib;
    const uint32_t *tile_max_contrib;
    const float *dL_dpix, *dL_dothers;
    float *gacc;  // [P][GACC_FLOATS], zero-initialised by the caller of the launch
    int subtile_cull;
    int *aux_flag = nullptr;  // one scratch word for the "any depth/normal/distortion gradient?" flag, or nullptr
    int variant = 0;          // kernel selection, see launch_render_bwd ("bwd_variant" option)
    int n_classes = 0;        // > 0: backward of the class-probability pass (dL_dpix = [n_classes,H,W], bg = [n_classes])
};
void launch_render_bwd(const RenderBwdArgs &a, cudaStream_t stream);

}  // namespace surfel
