/*
 * surfel_rasterizer.h -- C ABI of the B200-native differentiable 2D-Gaussian surfel rasterizer.
 *
 * This is the drop-in boundary for the hot path of DavidXu-JJ/StreetUnveiler: it replaces the
 * raw-pointer layer `CudaRasterizer::Rasterizer::{forward,backward,markVisible}`
 * (RAST/cuda_rasterizer/rasterizer.h:24-86, "RAST/" = submodules/diff-surfel-rasterization/) that
 * the reference's torch binding (RAST/rasterize_points.cu:39,136,235; RAST/ext.cpp:15-19) calls.
 *
 * Differences from the reference interface, all forced by a C ABI (no std::function, no torch):
 *   - The caller owns ALL device memory.  The three `std::function<char*(size_t)>` resize
 *     callbacks (rasterizer.h:32-34) become size queries + a two-stage forward: stage A
 *     (`surfel_forward_prepare`) returns num_rendered, the caller allocates the binning scratch of
 *     `surfel_binning_bytes(num_rendered)`, stage B (`surfel_forward_render`) finishes.
 *   - Every entry point takes the CUDA stream to launch on (the reference uses the legacy
 *     default stream everywhere) and returns 0 on success / non-zero on failure with the message
 *     available from `surfel_last_error()` (the reference throws std::runtime_error only when
 *     `debug` is set, auxiliary.h:296-303; here `debug` != 0 synchronises and checks per stage).
 *   - Gradient outputs do not need to be zero-filled by the caller (the reference binding
 *     zero-fills 304 B/Gaussian per backward, rasterize_points.cu:187-195): every element of
 *     every output is written by the library.
 *
 * All pointers are DEVICE pointers unless marked host.  All arrays are dense, row-major fp32
 * (int32 for radii), with exactly the reference's shapes and meaning.  Matrices use the
 * reference's layout: `viewmatrix` / `projmatrix` are the transposed (row-vector convention)
 * world->view and full projection matrices (scene/cameras.py:61-71).
 * A NULL pointer selects the alternative input path exactly like the reference's empty tensors:
 * `shs` xor `colors_precomp`; (`scales`,`rotations`) xor `transMat_precomp`.
 *
 * The library keeps no state between calls: forward -> backward state lives in the three
 * caller-owned scratch buffers (geometry / binning / image), whose layout is private.
 *
 * Numerics: forward outputs keep the reference's operation order, IEEE divisions and expf, and are bit-identical
 * to the reference extension on the same GPU.  The backward blend departs in one documented way: divisions whose
 * results only feed gradients (1/(1-alpha) for the transmittance recovery and the background term, 1/p.z and 1/depth
 * in the ray-splat and distortion gradients) use `rcp.approx.ftz.f32` (<= 1 ulp) instead of IEEE division; measured
 * gradient differences to the reference (2e-7 ... 5e-6 of max|ref|) equal its own run-to-run float-atomic noise, the
 * bar is 1e-4 (tests/test_parity_gpu.py prints both per tensor).
 *
 * Threading: every entry point may be called concurrently from several host threads (e.g. one per
 * GPU or per stream) as long as no two calls share a scratch or output buffer.  The error message
 * (`surfel_last_error`) and the pinned num_rendered read-back word are per host thread; the options
 * of `surfel_set_option` are process-wide atomics (each call reads an option once); the optional
 * stage clocks ("time_stages") are accumulated under a mutex and make every stage synchronise.
 */
#ifndef SURFEL_RASTERIZER_H_INCLUDED
#define SURFEL_RASTERIZER_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SURFEL_ABI_VERSION 1

#if defined(__GNUC__)
#define SURFEL_API __attribute__((visibility("default")))
#else
#define SURFEL_API
#endif

/* ABI version of the loaded library (== SURFEL_ABI_VERSION it was built with). */
SURFEL_API int surfel_abi_version(void);

/* Message of the last failing call on this thread ("" if none). Never NULL. */
SURFEL_API const char *surfel_last_error(void);

/*
 * Scratch sizes.  Replace `required<GeometryState>(P)`, `required<ImageState>(W*H)`,
 * `required<BinningState>(R)` (RAST/cuda_rasterizer/rasterizer_impl.h:67-73, used at
 * rasterizer_impl.cu:226,239,284).  Return 0 and set the error message on failure.
 */
SURFEL_API size_t surfel_geometry_bytes(int P);
SURFEL_API size_t surfel_image_bytes(int width, int height);
SURFEL_API size_t surfel_binning_bytes(int64_t num_rendered);

/*
 * Forward, stage A  (replaces Rasterizer::forward up to the num_rendered read-back,
 * rasterizer_impl.cu:198-282): per-Gaussian projection / tangent-frame transform / SH colour /
 * tile-rect count, depth ordering and prefix sum.
 *   radii          out int32[P]      (0 = culled), same values as the reference
 *   num_rendered   out HOST int64    number of (tile, Gaussian) instances R
 * Blocks the calling thread until R is known (the reference does the same, :282).
 */
SURFEL_API int surfel_forward_prepare(
    int P, int D, int M,
    int width, int height,
    const float *means3D,          /* [P,3] */
    const float *shs,              /* [P,M,3] or NULL */
    const float *colors_precomp,   /* [P,3]   or NULL */
    const float *opacities,        /* [P] */
    const float *scales,           /* [P,2]   or NULL */
    float scale_modifier,
    const float *rotations,        /* [P,4]   or NULL */
    const float *transMat_precomp, /* [P,9]   or NULL */
    const float *viewmatrix,       /* [16] */
    const float *projmatrix,       /* [16] */
    const float *cam_pos,          /* [3] */
    float tan_fovx, float tan_fovy,
    int prefiltered,
    int *radii,
    char *geometry_buffer,         /* surfel_geometry_bytes(P) */
    int64_t *num_rendered,         /* host */
    void *stream, int debug);

/*
 * Forward, stage B  (replaces rasterizer_impl.cu:284-342): (tile, depth) ordered instance list,
 * per-tile ranges and the front-to-back blend.
 *   out_color   out fp32[3,H,W]
 *   out_others  out fp32[7,H,W]  (depth, alpha, normal xyz, median depth, distortion)
 * Neither output needs to be initialised.
 */
SURFEL_API int surfel_forward_render(
    int P, int width, int height, int64_t num_rendered,
    const float *background,       /* [3] */
    const int *radii,
    char *geometry_buffer, char *binning_buffer, char *image_buffer,
    float *out_color, float *out_others,
    void *stream, int debug);

/*
 * Backward  (replaces Rasterizer::backward, rasterizer_impl.cu:346-448, with the reference's
 * gradient conventions -- including its documented non-true gradients).
 * Outputs (all fully written, no pre-zeroing needed):
 *   dL_dmean2D [P,3], dL_dopacity [P], dL_dcolor [P,3], dL_dmean3D [P,3], dL_dtransMat [P,9],
 *   dL_dsh [P,M,3] (ignored if M == 0 / shs NULL), dL_dscale [P,2], dL_drot [P,4]
 * `dL_dnormal` (internal in the reference, rasterize_points.cu:190) is optional (may be NULL).
 * `grad_scratch`: surfel_grad_scratch_bytes(P) bytes of device scratch.
 */
SURFEL_API size_t surfel_grad_scratch_bytes(int P);
SURFEL_API int surfel_backward(
    int P, int D, int M, int64_t num_rendered,
    const float *background,
    int width, int height,
    const float *means3D, const float *shs, const float *colors_precomp,
    const float *scales, float scale_modifier, const float *rotations,
    const float *transMat_precomp,
    const float *viewmatrix, const float *projmatrix, const float *cam_pos,
    float tan_fovx, float tan_fovy,
    const int *radii,
    char *geometry_buffer, char *binning_buffer, char *image_buffer,
    const float *dL_dpix,          /* [3,H,W] */
    const float *dL_dothers,       /* [7,H,W] */
    float *dL_dmean2D, float *dL_dnormal, float *dL_dopacity, float *dL_dcolor,
    float *dL_dmean3D, float *dL_dtransMat, float *dL_dsh, float *dL_dscale, float *dL_drot,
    char *grad_scratch,
    void *stream, int debug);

/* Replaces Rasterizer::markVisible (rasterizer_impl.cu:141-153): present[i] = view_z > 0.2 */
SURFEL_API int surfel_mark_visible(int P, const float *means3D, const float *viewmatrix, const float *projmatrix,
                        unsigned char *present, void *stream);

/*
 * Debug views into the private scratch layout (used only by the exact-equality tests):
 * copies the per-tile ranges [tiles,2] (uint32) and the sorted instance list [R] (uint32 Gaussian
 * ids) into caller-provided DEVICE buffers.
 */
SURFEL_API int surfel_debug_copy_binning(int width, int height, int64_t num_rendered,
                              const char *binning_buffer, const char *image_buffer,
                              uint32_t *ranges_out, uint32_t *point_list_out, void *stream);

/*
 * ---------------------------------------------------------------------------------------------
 * Sharded (multi-GPU) path -- not part of the reference, which is single-GPU only
 * (/root/reference/utils/general_utils.py:133).  Gaussians are sharded by index across G ranks; the SCREEN is
 * partitioned into G contiguous ranges of row-major tile ids ("window" = tiles [tile_lo, tile_hi)), cut so that
 * every range holds the same share of the frame's (tile, Gaussian) instances, and a projected record travels only
 * to the ranks whose range its tile rect touches.  Call sequence per rank (host side and the collectives:
 * streetunveiler_b200/sharded.py):
 *   surfel_shard_preprocess     own shard -> projected records [P,24] fp32, radii, depth keys, clamp bytes
 *   surfel_shard_tile_hist      instances per tile of the own shard        (all-reduce(sum) -> global histogram)
 *   surfel_shard_partition      global histogram -> cuts[G+1] (tile ranges) and the exact instance count of every range
 *   surfel_shard_route_count    per-destination send counts                (all-gather -> counts matrix; ONE host sync)
 *   surfel_shard_route_scatter  112-B rows (record, depth key, radius) into per-destination segments, index order kept
 *   (all-to-all of the rows)
 *   surfel_window_unpack        received rows -> records / radii / depth keys
 *   surfel_window_prepare       depth order of the received Gaussians, instance offsets of the own window
 *   surfel_window_render        binning + blend of the own tiles into full-size image planes (other tiles untouched)
 *   (all-reduce(sum) of the image planes: every pixel has exactly one non-zero summand)
 *   surfel_window_backward      gradient rows [n_received,20] of the own tiles (zeroed by the call)
 *   (all-to-all of the gradient rows back along the same routes)
 *   surfel_shard_grad_accumulate  returned rows summed into the per-Gaussian accumulator [P,20]
 *   surfel_shard_backward       own shard: accumulator row -> parameter gradients
 * Every per-pixel result is identical to the single-GPU path (same lists, same order); upstream gradients must be
 * the same on every rank (each rank back-propagates only its own tiles).
 * surfel_shard_compact / the (records, radii, keys) all-gather scheme of round 1 remain available.
 * ---------------------------------------------------------------------------------------------
 */
SURFEL_API int surfel_shard_preprocess(
    int P, int D, int M, int width, int height,
    const float *means3D, const float *shs, const float *colors_precomp, const float *opacities,
    const float *scales, float scale_modifier, const float *rotations, const float *transMat_precomp,
    const float *viewmatrix, const float *projmatrix, const float *cam_pos,
    float tan_fovx, float tan_fovy, int prefiltered,
    int *radii, float *records, uint32_t *depth_keys, unsigned char *clamped, void *stream);
/* Compaction: outputs have room for P rows; rows [count, P) of radii_c / depth_keys_c are set to the
 * "culled" pattern (0 / 0xFFFFFFFF) so that any prefix of >= count rows can be exchanged as is.
 * *count_dev (device int) receives the number of visible Gaussians; temp: surfel_shard_compact_bytes(P). */
SURFEL_API size_t surfel_shard_compact_bytes(int P);
SURFEL_API int surfel_shard_compact(
    int P, const int *radii, const float *records, const uint32_t *depth_keys,
    float *records_c, int *radii_c, uint32_t *depth_keys_c, uint32_t *slot, int *count_dev,
    char *temp, void *stream);
/* hist: uint32[tiles] (tiles = ceil(W/16) * ceil(H/16)), overwritten. */
SURFEL_API int surfel_shard_tile_hist(int P, int width, int height, const float *records, const int *radii,
                                      uint32_t *hist, void *stream);
/* hist: the all-reduced histogram.  cost of a tile = its instance count + cost_base; rank k receives the share
 * shares[k] / sum(shares) of the total cost (shares: HOST float[G] or NULL = equal shares; the host side adapts them
 * from the ranks' measured blend times).  cuts: DEVICE int32[G+1], window_num_rendered: DEVICE int64[G];
 * temp: surfel_shard_partition_bytes(width, height).  1 <= G <= 16. */
SURFEL_API size_t surfel_shard_partition_bytes(int width, int height);
SURFEL_API int surfel_shard_partition(int width, int height, int G, const uint32_t *hist, int cost_base,
                                      const float *shares, char *temp, int *cuts, int64_t *window_num_rendered,
                                      void *stream);
/* temp: surfel_shard_route_bytes(P, G), shared by the two calls.  send_counts: DEVICE int32[G + 1]; word G receives
 * `extra`, a caller-defined value that travels with the counts through the all-gather (the host side puts the
 * rank's last measured blend time there).
 * send_rows: [sum(send_counts), 28] fp32, segment d = rows for rank d; send_src[row] = local Gaussian index. */
SURFEL_API size_t surfel_shard_route_bytes(int P, int G);
SURFEL_API int surfel_shard_route_count(int P, int width, int height, int G, const float *records, const int *radii,
                                        const int *cuts, char *temp, int *send_counts, int extra, void *stream);
SURFEL_API int surfel_shard_route_scatter(int P, int G, const float *records, const int *radii,
                                          const uint32_t *depth_keys, char *temp, const int *send_counts,
                                          float *send_rows, uint32_t *send_src, void *stream);
/* The same scatter with the exchange fused in: rows are stored straight into the destination ranks' receive arrays
 * (HOST arrays of G DEVICE addresses this GPU can store to -- symmetric memory over NVLink; entry d = rank d's records
 * [cap,24] fp32 / depth keys [cap] / radii [cap]) starting at row dst_row0[d] (HOST int64[G]: rows that lower ranks
 * send to d).  No send buffer, no all-to-all, no unpack; the caller synchronises the ranks before anyone reads. */
SURFEL_API int surfel_shard_route_scatter_peers(int P, int G, const float *records, const int *radii,
                                                const uint32_t *depth_keys, char *temp, const int *send_counts,
                                                float *const *dst_records, uint32_t *const *dst_depth_keys,
                                                int *const *dst_radii, const int64_t *dst_row0, uint32_t *send_src,
                                                void *stream);
/* Backward counterpart: gradient rows [n_rows,20] of the received records, in receive order (segment s = the
 * seg_count[s] rows that came from rank s), are stored into rank s's return buffer dst_rows[s] starting at row
 * dst_row0[s] -- the position those records have in rank s's send order. */
SURFEL_API int surfel_window_push_grad_rows(int64_t n_rows, const float *grad_rows, int G, const int64_t *seg_count,
                                            float *const *dst_rows, const int64_t *dst_row0, void *stream);
SURFEL_API int surfel_window_unpack(int n, const float *rows, float *records, int *radii, uint32_t *depth_keys,
                                    void *stream);
/* grad_records [P,20] is zeroed, then grad_records[send_src[j]] += grad_rows[j] for j < n_rows. */
SURFEL_API int surfel_shard_grad_accumulate(int P, int64_t n_rows, const float *grad_rows, const uint32_t *send_src,
                                            float *grad_records, void *stream);
SURFEL_API size_t surfel_window_bytes(int P_total);
/* num_rendered (HOST) may be NULL when the caller already knows the window's instance count (surfel_shard_partition):
 * then the call does not synchronise. */
SURFEL_API int surfel_window_prepare(
    int P_total, int width, int height, int tile_lo, int tile_hi,
    const float *records, const int *radii, const uint32_t *depth_keys,
    char *window_buffer, int64_t *num_rendered, void *stream, int debug);
SURFEL_API int surfel_window_render(
    int P_total, int width, int height, int tile_lo, int tile_hi, int64_t num_rendered,
    const float *background, const float *records, const int *radii,
    char *window_buffer, char *binning_buffer, char *image_buffer,
    float *out_color, float *out_others, void *stream, int debug);
/* Same, with the image exchange fused into the blend: the ten planes of the window's pixels are stored into the
 * [10,H,W] fp32 buffers peer_planes[0..n_peers) (HOST array of DEVICE addresses that this GPU can store to: its own
 * buffer and its peers' over NVLink; or, with multicast != 0, ONE NVSwitch multicast address that reaches all of
 * them, written with multimem.st).  Pixels outside the window are not touched.  The caller synchronises the ranks
 * (a barrier) before anyone reads. */
SURFEL_API int surfel_window_render_peers(
    int P_total, int width, int height, int tile_lo, int tile_hi, int64_t num_rendered,
    const float *background, const float *records, const int *radii,
    char *window_buffer, char *binning_buffer, char *image_buffer,
    int n_peers, float *const *peer_planes, int multicast, void *stream, int debug);
SURFEL_API int surfel_window_backward(
    int P_total, int width, int height, int tile_lo, int tile_hi, int64_t num_rendered,
    const float *background, const float *records, char *binning_buffer, char *image_buffer,
    const float *dL_dpix, const float *dL_dothers, float *grad_records, void *stream, int debug);
SURFEL_API int surfel_shard_backward(
    int P, int D, int M, int width, int height,
    const float *means3D, const float *shs, const float *scales, const float *rotations,
    const float *transMat_precomp, const float *viewmatrix, const float *projmatrix, const float *cam_pos,
    float tan_fovx, float tan_fovy,
    const int *radii, const float *records, const unsigned char *clamped, const float *grad_records,
    const uint32_t *grad_slot /* NULL: grad_records row i belongs to Gaussian i */,
    float *dL_dmean2D, float *dL_dnormal, float *dL_dopacity, float *dL_dcolor,
    float *dL_dmean3D, float *dL_dtransMat, float *dL_dsh, float *dL_dscale, float *dL_drot, void *stream);

/*
 * ---------------------------------------------------------------------------------------------
 * Fused render() epilogue (reference: gaussian_renderer/__init__.py:148-186 +
 * utils/point_utils.py:8-37, ~15 PyTorch kernels each way).  One kernel forward, one backward.
 *   allmap [7,H,W], viewmatrix [16] = world_view_transform (device), fx = W / (2 tan(FoVx/2)), fy likewise
 *   forward  -> rend_normal [3,H,W] (world space), surf_depth [1,H,W], surf_normal [3,H,W] (alpha-weighted),
 *               surf_point [3,H,W]; optionally copies of the pass-through planes rend_alpha / rend_dist
 *   backward -> dL_dallmap [7,H,W] from the gradients of those outputs (every element written;
 *               alpha == 0 pixels get 0/0 = NaN in channels 0,1 exactly like autograd in the reference)
 * ---------------------------------------------------------------------------------------------
 */
SURFEL_API int surfel_epilogue_forward(int width, int height, const float *allmap, const float *viewmatrix,
                                       float fx, float fy, float depth_ratio, float *rend_normal,
                                       float *surf_depth, float *surf_normal, float *surf_point,
                                       float *rend_alpha /* [1,H,W] copy of channel 1, or NULL */,
                                       float *rend_dist /* [1,H,W] copy of channel 6, or NULL */, void *stream);
SURFEL_API int surfel_epilogue_backward(int width, int height, const float *allmap, const float *surf_point,
                                        const float *viewmatrix, float fx, float fy, float depth_ratio,
                                        const float *g_rend_normal, const float *g_surf_depth,
                                        const float *g_surf_normal, const float *g_surf_point,
                                        const float *g_rend_alpha /* or NULL */, const float *g_rend_dist /* or NULL */,
                                        float *dL_dallmap, void *stream);

/*
 * ---------------------------------------------------------------------------------------------
 * Colour passes over ONE geometry / binning state.  The reference's render_semantic
 * (gaussian_renderer/__init__.py:327-460) calls the rasterizer once per 3 semantic classes with identical geometry
 * and different `colors_precomp` / `bg`, repeating projection, both sorts and binning every time, and autograd then
 * runs one full backward per call.  Here the first pass is surfel_forward_prepare + surfel_forward_render with the
 * first colour set; every further pass is
 *     surfel_pass_set_colors  (rewrites the 12 colour bytes of each packed record)
 *     surfel_pass_render      (blend only: same lists, same order -> out_color differs, out_others is identical)
 * and the backward is, per pass in any order,
 *     surfel_pass_set_colors; surfel_pass_backward_blend (clear = 1 for the first call only);
 *     surfel_pass_take_color_grad (that pass's dL/dcolors_precomp, or NULL to discard; clears the colour words)
 * followed by ONE surfel_pass_backward_geometry for the summed geometry gradients.  dL_dothers of a pass that
 * contributed no allmap gradient is an all-zero [7,H,W] array.  Buffers are the ones of the forward calls.
 * ---------------------------------------------------------------------------------------------
 */
SURFEL_API int surfel_pass_set_colors(int P, const float *colors_precomp, char *geometry_buffer, void *stream);
SURFEL_API int surfel_pass_render(int P, int width, int height, int64_t num_rendered, const float *background,
                                  char *geometry_buffer, char *binning_buffer, char *image_buffer, float *out_color,
                                  float *out_others, void *stream, int debug);
SURFEL_API int surfel_pass_backward_blend(int P, int width, int height, int64_t num_rendered, const float *background,
                                          char *geometry_buffer, char *binning_buffer, char *image_buffer,
                                          const float *dL_dpix, const float *dL_dothers, char *grad_scratch, int clear,
                                          void *stream, int debug);
SURFEL_API int surfel_pass_take_color_grad(int P, char *grad_scratch, float *dL_dcolor, void *stream);
SURFEL_API int surfel_pass_backward_geometry(int P, int width, int height, const float *means3D, const float *scales,
                                             const float *rotations, const float *transMat_precomp,
                                             const float *viewmatrix, const float *projmatrix, const float *cam_pos,
                                             float tan_fovx, float tan_fovy, const int *radii, char *geometry_buffer,
                                             char *grad_scratch, float *dL_dmean2D, float *dL_dnormal,
                                             float *dL_dopacity, float *dL_dcolor, float *dL_dmean3D,
                                             float *dL_dtransMat, float *dL_dscale, float *dL_drot, void *stream,
                                             int debug);

/*
 * Class-probability pass: ALL classes of render_semantic in one traversal of the lists.  The reference renders class
 * probabilities as one-hot "colours", three classes per rasterizer call (gaussian_renderer/__init__.py:417-446).  With a
 * per-Gaussian label instead (labels [P] int32; a label outside [0, n_classes) contributes to no class) one blend pass
 * accumulates up to 8 class channels -- bit-identical to the one-hot formulation, since 1 * w and 0 * w are exact --
 * and one backward blend returns the geometry gradients of all of them.  Sequence:
 *     surfel_forward_prepare (any colour source; colours are not used)  ->  surfel_forward_bin
 *     surfel_classes_set_labels  ->  surfel_classes_render (background [n_classes] -> out_probs [n_classes,H,W])
 *     backward: surfel_classes_set_labels; surfel_classes_backward_blend (dL_dprobs [n_classes,H,W]);
 *               surfel_pass_backward_geometry
 */
SURFEL_API int surfel_forward_bin(int P, int width, int height, int64_t num_rendered, const int *radii,
                                  char *geometry_buffer, char *binning_buffer, char *image_buffer, void *stream,
                                  int debug);
SURFEL_API int surfel_classes_set_labels(int P, const int *labels, char *geometry_buffer, void *stream);
SURFEL_API int surfel_classes_render(int P, int width, int height, int64_t num_rendered, int n_classes,
                                     const float *background, char *geometry_buffer, char *binning_buffer,
                                     char *image_buffer, float *out_probs, void *stream, int debug);
SURFEL_API int surfel_classes_backward_blend(int P, int width, int height, int64_t num_rendered, int n_classes,
                                             const float *background, char *geometry_buffer, char *binning_buffer,
                                             char *image_buffer, const float *dL_dprobs, char *grad_scratch, int clear,
                                             void *stream, int debug);

/*
 * ---------------------------------------------------------------------------------------------
 * Fused training-loss block (reference: utils/loss_utils.py:17-64 l1_loss / ssim and
 * train.py:113-136, ~60 PyTorch kernels per iteration).  All images are dense fp32 [C,H,W] device arrays.
 *
 *   photometric:  img1 = render + sky * (1 - rend_alpha)   (train.py:113; sky == rend_alpha == NULL: img1 = render)
 *                 out_means[0] = mean |img1 - gt|           (loss_utils.py:17-18)
 *                 out_means[1] = mean ssim_map(img1, gt)    (loss_utils.py:33-64: window 11, sigma 1.5, zero padding,
 *                                                            C1 = 0.01^2, C2 = 0.03^2, size_average = True)
 *     forward   also writes deriv [9,H,W] (three maps per channel that the backward convolves; pass NULL when no
 *               gradient is needed); scratch = surfel_loss_scratch_bytes(W, H) bytes; out_means = 2 floats (device).
 *     backward  upstream = 2 floats (device): dL/d out_means[0], dL/d out_means[1]  -> d_render [3,H,W] and, with a
 *               sky, d_rend_alpha [1,H,W] and d_sky [3,H,W] (either may be NULL).  Gradients w.r.t. gt are not produced.
 *   regulariser:  out_means[0] = mean_p (1 - sum_c rend_normal * surf_normal)  (train.py:125-126)
 *                 out_means[1] = mean rend_dist                                (train.py:134)
 *     backward  upstream = 2 floats (device) -> d_rend_normal, d_surf_normal [3,H,W], d_rend_dist [1,H,W].
 * The means are reduced in two deterministic stages (per-CTA fp32 partials, then one CTA in fp64): the same
 * inputs give bit-identical results run to run.  No call synchronises the stream.
 * ---------------------------------------------------------------------------------------------
 */
SURFEL_API size_t surfel_loss_scratch_bytes(int width, int height);
SURFEL_API int surfel_loss_photometric_forward(int width, int height, const float *render, const float *rend_alpha,
                                               const float *sky, const float *gt, float *deriv, char *scratch,
                                               float *out_means, void *stream);
SURFEL_API int surfel_loss_photometric_backward(int width, int height, const float *render, const float *rend_alpha,
                                                const float *sky, const float *gt, const float *deriv,
                                                const float *upstream, float *d_render, float *d_rend_alpha,
                                                float *d_sky, void *stream);
SURFEL_API int surfel_loss_regulariser_forward(int width, int height, const float *rend_normal,
                                               const float *surf_normal, const float *rend_dist, char *scratch,
                                               float *out_means, void *stream);
SURFEL_API int surfel_loss_regulariser_backward(int width, int height, const float *rend_normal,
                                                const float *surf_normal, const float *upstream,
                                                float *d_rend_normal, float *d_surf_normal, float *d_rend_dist,
                                                void *stream);
/* Both halves and the combination of train.py:117-136 in three launches forward and two backward:
 *   out5 = (loss, l1, ssim, Lnormal, Ldist),  loss = (1 - lambda_dssim) l1 + lambda_dssim (1 - ssim) + Lnormal + Ldist,
 *   Lnormal = lambda_normal * mean(1 - <rend_normal, surf_normal>),  Ldist = lambda_dist * mean(rend_dist).
 * g_out5 = 5 floats (device): the gradient of whatever the caller does with out5 (d/d loss = 1 for plain training). */
SURFEL_API int surfel_loss_training_forward(int width, int height, const float *render, const float *rend_alpha,
                                            const float *sky, const float *gt, const float *rend_normal,
                                            const float *surf_normal, const float *rend_dist, float lambda_dssim,
                                            float lambda_normal, float lambda_dist, float *deriv, char *scratch,
                                            float *out5, void *stream);
SURFEL_API int surfel_loss_training_backward(int width, int height, const float *render, const float *rend_alpha,
                                             const float *sky, const float *gt, const float *rend_normal,
                                             const float *surf_normal, const float *deriv, const float *g_out5,
                                             float lambda_dssim, float lambda_normal, float lambda_dist,
                                             float *d_render, float *d_rend_alpha, float *d_sky,
                                             float *d_rend_normal, float *d_surf_normal, float *d_rend_dist,
                                             void *stream);

/*
 * ---------------------------------------------------------------------------------------------
 * Fused per-iteration parameter update.
 *
 * surfel_adam_step: what `gaussians.optimizer.step()` (train.py:197) does for the optimiser built at
 * scene/gaussian_model.py:171-180 -- torch.optim.Adam over up to 8 single-tensor groups, each with its own
 * learning rate and step count (no weight decay, no amsgrad) -- in ONE kernel launch: every parameter, gradient
 * and moment element is read once and written once.  `groups` is a HOST array; the pointers inside are device
 * pointers to dense fp32 arrays of n elements; `step` is the 1-based count of this update (torch's state["step"]
 * after its increment).  The arithmetic follows torch/optim/adam.py (PyTorch 2.11) operation by operation.
 *
 * surfel_densification_stats: train.py:168 + scene/gaussian_model.py:555-557 in one launch, for the Gaussians with
 * radii > 0:  max_radii2D = max(max_radii2D, radii);  xyz_gradient_accum += ||viewspace_grad||_2;  denom += 1.
 * radii [P] int32, viewspace_grad [P,3], max_radii2D [P], xyz_gradient_accum [P,1], denom [P,1] (fp32).
 * ---------------------------------------------------------------------------------------------
 */
typedef struct surfel_adam_group {
    float *param;
    const float *grad;
    float *exp_avg;
    float *exp_avg_sq;
    int64_t n;
    double lr;   /* hyper-parameters are doubles, as in Python: 1 - beta2 must not be formed from a rounded fp32 beta2 */
    int step;
} surfel_adam_group;
SURFEL_API int surfel_adam_step(int n_groups, const surfel_adam_group *groups, double beta1, double beta2, double eps,
                                void *stream);
SURFEL_API int surfel_densification_stats(int P, const int *radii, const float *viewspace_grad, float *max_radii2D,
                                          float *xyz_gradient_accum, float *denom, void *stream);

/*
 * ---------------------------------------------------------------------------------------------
 * Fused parameter activations and SH packing -- the caller-side row in front of the projection kernel.
 * Reference: the GaussianModel properties every render() evaluates (scene/gaussian_model.py:31-39,101-127):
 *   get_scaling = exp(_scaling) [P,2];  get_rotation = normalize(_rotation) [P,4] (x / max(||x||, 1e-12));
 *   get_opacity = sigmoid(_opacity) [P,1];  get_features = cat((_features_dc [P,1,3], _features_rest [P,R,3]), 1)
 * and their autograd (~12 PyTorch kernels forward, ~15 backward, two of them copies of the whole SH block).
 * One launch each way.  sh_rest = R; features / g_features may be NULL (activations only); when given, the packed block
 * must be 16-byte aligned.  The backward takes the ACTIVATED scaling / opacity
 * (the forward's outputs) and the raw quaternion.
 * ---------------------------------------------------------------------------------------------
 */
SURFEL_API int surfel_activate_forward(int P, int sh_rest, const float *scaling_raw, const float *rotation_raw,
                                       const float *opacity_raw, const float *features_dc,
                                       const float *features_rest, float *scaling, float *rotation, float *opacity,
                                       float *features, void *stream);
SURFEL_API int surfel_activate_backward(int P, int sh_rest, const float *rotation_raw, const float *scaling,
                                        const float *opacity, const float *g_scaling, const float *g_rotation,
                                        const float *g_opacity, const float *g_features, float *d_scaling_raw,
                                        float *d_rotation_raw, float *d_opacity_raw, float *d_features_dc,
                                        float *d_features_rest, void *stream);

/* Test hook for the hand-written stable LSD radix sort used by the binning stage: sorts n
 * (uint32 key, uint32 value) pairs on key bits [0, end_bit) into the *_out arrays (device pointers). */
SURFEL_API int surfel_debug_sort_pairs(int64_t n, int end_bit, const uint32_t *keys_in, const uint32_t *vals_in,
                                       uint32_t *keys_out, uint32_t *vals_out, void *stream);

/* Debug view of the geometry scratch: per-Gaussian tile counts [P], depth-ordered ids [P], inclusive
 * offsets [P] (all uint32) and the packed 96-byte projected records [P,24] (fp32 words). Any output
 * may be NULL. */
SURFEL_API int surfel_debug_copy_geometry(int P, const char *geometry_buffer, uint32_t *tiles_touched_out,
                                          uint32_t *idx_sorted_out, uint32_t *offsets_out, float *records_out,
                                          void *stream);

/* Debug view of the backward's device-side decision (render_bwd.cu: aux_zero_scan_kernel): *flag_host (HOST int) = 1
 * if the last surfel_backward / surfel_pass_backward_blend on this grad_scratch ran the full blend specialisation
 * (some depth / normal / median-depth / distortion gradient was non-zero at a pixel that blended a splat), 0 if the
 * colour+alpha one.  Synchronises the stream. */
SURFEL_API int surfel_debug_aux_flag(int P, const char *grad_scratch, int *flag_host, void *stream);

/* Tuning / debug knobs: "radix_onesweep" (default 1; 0 = three-launch radix passes), "bwd_variant", "subtile_cull" (default 1), "time_stages" (default 0; setting it clears the
 * stage clocks).  Returns 0 if the option exists. */
SURFEL_API int surfel_set_option(const char *name, int value);

/*
 * Per-stage device timing for roofline measurement (bench.py).  With "time_stages" = 1 every
 * stage is bracketed by CUDA events on the launching stream (and the call synchronises on them):
 * stages 0..surfel_stage_count()-1 = preprocess_fwd, depth_order, tile_binning, render_fwd,
 * render_bwd, preprocess_bwd.  surfel_stage_time returns the accumulated milliseconds and the
 * number of timed launches since the option was last set.
 */
SURFEL_API int surfel_stage_count(void);
SURFEL_API const char *surfel_stage_name(int stage);
SURFEL_API int surfel_stage_time(int stage, double *total_ms, int *calls);

#ifdef __cplusplus
}
#endif
#endif /* SURFEL_RASTERIZER_H_INCLUDED */
