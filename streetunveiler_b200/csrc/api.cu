// api.cu -- extern "C" boundary (include/surfel_rasterizer.h) and host orchestration.
//
// Mirrors the raw-pointer layer CudaRasterizer::Rasterizer::{forward,backward,markVisible}
// (RAST/cuda_rasterizer/rasterizer.h:24-86; orchestration rasterizer_impl.cu:198-342,346-448)
// with caller-owned memory, an explicit stream and int error codes.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/surfel_rasterizer.h"
#include "common.cuh"
#include "kernels.h"

using namespace surfel;

namespace {

// Threading contract (stated in include/surfel_rasterizer.h): every entry point may be called concurrently from
// several host threads (one per GPU / stream).  Error text and the read-back slot are per thread; the options are
// process-wide atomics (a call reads each of them once); the optional stage clocks are guarded by a mutex.
thread_local std::string g_err;
std::atomic<int> g_subtile_cull{1};
std::atomic<int> g_bwd_variant{2};  // see launch_render_bwd

// ---- optional per-stage device timing (bench.py roofline measurement) ----
constexpr int N_STAGES = 6;
const char *const kStageNames[N_STAGES] = {"preprocess_fwd", "depth_order", "tile_binning",
                                           "render_fwd",     "render_bwd",  "preprocess_bwd"};
std::atomic<int> g_time_stages{0};
std::mutex g_stage_mutex;
double g_stage_ms[N_STAGES] = {0, 0, 0, 0, 0, 0};
int g_stage_calls[N_STAGES] = {0, 0, 0, 0, 0, 0};

struct StageClock {  // brackets a stage with two events on the launching stream
    cudaStream_t st;
    int stage;
    cudaEvent_t a = nullptr, b = nullptr;
    StageClock(cudaStream_t s, int id) : st(s), stage(id)
    {
        if (!g_time_stages.load(std::memory_order_relaxed)) return;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, st);
    }
    void stop()
    {
        if (!a) return;
        cudaEventRecord(b, st);
        cudaEventSynchronize(b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, a, b) == cudaSuccess) {
            std::lock_guard<std::mutex> lock(g_stage_mutex);
            g_stage_ms[stage] += ms;
            g_stage_calls[stage] += 1;
        }
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        a = b = nullptr;
    }
    ~StageClock() { stop(); }
};

int fail(const char *where, const char *what)
{
    g_err = std::string(where) + ": " + what;
    return 1;
}
int fail_cuda(const char *where, cudaError_t e)
{
    g_err = std::string(where) + ": CUDA error: " + cudaGetErrorString(e);
    return 2;
}

#define CK(where, expr)                                    \
    do {                                                   \
        cudaError_t _e = (expr);                           \
        if (_e != cudaSuccess) return fail_cuda(where, _e); \
    } while (0)

// debug: synchronise + surface asynchronous errors after every stage (reference CHECK_CUDA, auxiliary.h:296-303)
#define STAGE(where)                                                   \
    do {                                                               \
        cudaError_t _e = cudaGetLastError();                           \
        if (_e == cudaSuccess && debug) _e = cudaStreamSynchronize(st); \
        if (_e != cudaSuccess) return fail_cuda(where, _e);            \
    } while (0)

GeomView carve_geom(char *base, int P, size_t cub_bytes)
{
    GeomView g;
    char *p = base;
    const size_t n = (size_t)(P > 0 ? P : 1);
    g.rec = carve<float>(p, n * REC_FLOATS);
    g.tiles_touched = carve<uint32_t>(p, n);
    g.depth_key = carve<uint32_t>(p, n);
    g.depth_key_sorted = carve<uint32_t>(p, n);
    g.idx_in = carve<uint32_t>(p, n);
    g.idx_sorted = carve<uint32_t>(p, n);
    g.offsets = carve<uint32_t>(p, n);
    g.clamped = carve<uint8_t>(p, n);
    g.cub_temp = carve<char>(p, cub_bytes);
    g.cub_temp_bytes = cub_bytes;
    return g;
}
size_t geom_end(int P, size_t cub_bytes)
{
    GeomView g = carve_geom(nullptr, P, cub_bytes);
    return (size_t)(g.cub_temp + cub_bytes) + 128;
}

ImageView carve_image(char *base, int W, int H)
{
    ImageView v;
    char *p = base;
    const size_t HW = (size_t)W * H;
    const size_t tiles = (size_t)((W + TILE_X - 1) / TILE_X) * ((H + TILE_Y - 1) / TILE_Y);
    v.ranges = carve<uint2>(p, tiles);
    v.final_T = carve<float>(p, 3 * HW);
    v.n_contrib = carve<uint32_t>(p, 2 * HW);
    v.tile_max_contrib = carve<uint32_t>(p, tiles);
    v.tile_order = carve<uint32_t>(p, tiles);
    v.aux_flag = carve<int>(p, 1);
    return v;
}

BinView carve_bin(char *base, int64_t R, size_t cub_bytes)
{
    BinView b;
    char *p = base;
    const size_t n = (size_t)(R > 0 ? R : 1);
    b.keys_unsorted = carve<uint32_t>(p, n);
    b.vals_unsorted = carve<uint32_t>(p, n);
    b.keys_sorted = carve<uint32_t>(p, n);
    b.point_list = carve<uint32_t>(p, n);
    b.cub_temp = carve<char>(p, cub_bytes);
    b.cub_temp_bytes = cub_bytes;
    return b;
}

// Scratch of the sharded path's per-rank window pass (P = all gathered Gaussians).
struct WindowView {
    uint32_t *tiles_touched, *idx_in, *idx_sorted, *depth_key_sorted, *offsets;
    char *cub_temp;
    size_t cub_temp_bytes;
};
WindowView carve_window(char *base, int P, size_t cub_bytes)
{
    WindowView w;
    char *p = base;
    const size_t n = (size_t)(P > 0 ? P : 1);
    w.tiles_touched = carve<uint32_t>(p, n);
    w.idx_in = carve<uint32_t>(p, n);
    w.idx_sorted = carve<uint32_t>(p, n);
    w.depth_key_sorted = carve<uint32_t>(p, n);
    w.offsets = carve<uint32_t>(p, n);
    w.cub_temp = carve<char>(p, cub_bytes);
    w.cub_temp_bytes = cub_bytes;
    return w;
}

// ---- num_rendered read-back without the copy engine -------------------------------------------------
// The reference reads R back with a blocking cudaMemcpy into pageable memory (rasterizer_impl.cu:282).  That 4-byte copy
// is queued on the device-to-host DMA engine BEHIND whatever bulk transfer another stream has in flight -- in a pipelined
// training loop the whole forward then waits for the previous step's results to finish leaving the device (measured: the
// step time becomes compute + transfer instead of max(compute, transfer)).  Here a one-thread kernel stores the value
// straight into a pinned, mapped host word (one lazily allocated slot per host thread, never freed), so the read-back
// only waits for the kernels in front of it on its own stream.
__global__ void publish_u32_kernel(const uint32_t *__restrict__ src, volatile uint32_t *__restrict__ dst_mapped)
{
    *dst_mapped = *src;
    __threadfence_system();
}

cudaError_t read_back_u32(const uint32_t *dev_value, uint32_t *out, cudaStream_t st)
{
    thread_local uint32_t *slot = nullptr;
    if (!slot) {
        void *h = nullptr;
        cudaError_t e = cudaHostAlloc(&h, 64, cudaHostAllocPortable | cudaHostAllocMapped);
        if (e != cudaSuccess) return e;
        slot = static_cast<uint32_t *>(h);
    }
    uint32_t *dptr = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(reinterpret_cast<void **>(&dptr), slot, 0);
    if (e != cudaSuccess) return e;
    publish_u32_kernel<<<1, 1, 0, st>>>(dev_value, dptr);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    *out = *static_cast<volatile uint32_t *>(slot);
    return cudaSuccess;
}

bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

bool have_device()
{
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
}


// K7 into the per-Gaussian gradient accumulator (optionally cleared first) ...
int backward_blend(int P, int width, int height, int64_t num_rendered, const float *background, const GeomView &g,
                   const ImageView &iv, const BinView &bv, const float *dL_dpix, const float *dL_dothers,
                   char *grad_scratch, bool clear, cudaStream_t st, int debug, int n_classes = 0)
{
    char *gp = grad_scratch;
    float *gacc = carve<float>(gp, (size_t)P * GACC_FLOATS);
    if (clear) CK("grad scratch clear", cudaMemsetAsync(gacc, 0, (size_t)P * GACC_FLOATS * sizeof(float), st));
    if (num_rendered > 0) {
        RenderBwdArgs r;
        r.W = width; r.H = height;
        r.gx = (width + TILE_X - 1) / TILE_X; r.gy = (height + TILE_Y - 1) / TILE_Y;
        r.ranges = iv.ranges; r.tile_order = iv.tile_order; r.point_list = bv.point_list; r.rec = g.rec; r.bg = background;
        r.final_T = iv.final_T; r.n_contrib = iv.n_contrib; r.tile_max_contrib = iv.tile_max_contrib;
        r.dL_dpix = dL_dpix; r.dL_dothers = dL_dothers; r.gacc = gacc; r.subtile_cull = g_subtile_cull;
        r.aux_flag = carve<int>(gp, 1);   // the 128 spare bytes behind the accumulator (surfel_grad_scratch_bytes)
        r.variant = g_bwd_variant;
        r.n_classes = n_classes;
        {
            StageClock clk(st, 4);
            launch_render_bwd(r, st);
        }
        STAGE("render backward");
    }
    return 0;
}

// ... and K8 from the accumulator to the parameter gradients.
int backward_geometry(int P, int D, int M, int width, int height, const float *means3D, const float *shs,
                      const float *scales, const float *rotations, const float *transMat_precomp,
                      const float *viewmatrix, const float *projmatrix, const float *cam_pos, float tan_fovx,
                      float tan_fovy, const int *radii, const GeomView &g, char *grad_scratch, float *dL_dmean2D,
                      float *dL_dnormal, float *dL_dopacity, float *dL_dcolor, float *dL_dmean3D, float *dL_dtransMat,
                      float *dL_dsh, float *dL_dscale, float *dL_drot, cudaStream_t st, int debug)
{
    char *gp = grad_scratch;
    float *gacc = carve<float>(gp, (size_t)P * GACC_FLOATS);
    PreprocessBwdArgs a;
    a.P = P; a.D = D; a.M = M;
    a.focal_y = height / (2.0f * tan_fovy);  // rasterizer_impl.cu:388-389
    a.focal_x = width / (2.0f * tan_fovx);
    a.tan_fovx = tan_fovx; a.tan_fovy = tan_fovy;
    a.means3D = means3D; a.shs = shs; a.scales = scales; a.rotations = rotations;
    a.transMat_precomp = transMat_precomp; a.viewmatrix = viewmatrix; a.projmatrix = projmatrix; a.cam_pos = cam_pos;
    a.radii = radii; a.clamped = g.clamped; a.rec = g.rec; a.gacc = gacc;
    a.dL_dmean2D = dL_dmean2D; a.dL_dnormal = dL_dnormal; a.dL_dopacity = dL_dopacity; a.dL_dcolor = dL_dcolor;
    a.dL_dmean3D = dL_dmean3D; a.dL_dtransMat = dL_dtransMat; a.dL_dsh = dL_dsh; a.dL_dscale = dL_dscale;
    a.dL_drot = dL_drot;
    {
        StageClock clk(st, 5);
        launch_preprocess_bwd(a, st);
    }
    STAGE("preprocess backward");
    return 0;
}

// K3-K5: emit (tile, id) pairs in depth order, stable-sort by tile, per-tile ranges, launch order
int bin_stage(int P, int width, int height, int64_t num_rendered, const int *radii, const GeomView &g, const ImageView &iv,
              const BinView &bv, cudaStream_t st, int debug)
{
    const int gx = (width + TILE_X - 1) / TILE_X, gy = (height + TILE_Y - 1) / TILE_Y;
    StageClock clk_bin(st, 2);
    CK("tile binning", run_tile_binning(P, num_rendered, gx, gy, 0, gx * gy, g.rec, radii, g.idx_sorted, g.offsets,
                                        bv.keys_unsorted, bv.vals_unsorted, bv.keys_sorted, bv.point_list, iv.ranges,
                                        iv.tile_order, bv.cub_temp, bv.cub_temp_bytes, st));
    clk_bin.stop();
    STAGE("tile binning");
    return 0;
}

// K6 over an existing binning state
int render_pass(int width, int height, const float *background, const GeomView &g, const ImageView &iv,
                const BinView &bv, float *out_color, float *out_others, cudaStream_t st, int debug, int n_classes = 0)
{
    RenderFwdArgs r;
    r.n_classes = n_classes;
    r.W = width; r.H = height;
    r.gx = (width + TILE_X - 1) / TILE_X; r.gy = (height + TILE_Y - 1) / TILE_Y;
    r.ranges = iv.ranges; r.tile_order = iv.tile_order; r.point_list = bv.point_list; r.rec = g.rec; r.bg = background;
    r.final_T = iv.final_T; r.n_contrib = iv.n_contrib; r.tile_max_contrib = iv.tile_max_contrib;
    r.out_color = out_color; r.out_others = out_others; r.subtile_cull = g_subtile_cull;
    {
        StageClock clk(st, 3);
        launch_render_fwd(r, st);
    }
    STAGE("render forward");
    return 0;
}

}  // namespace

int surfel_internal_fail(const char *where, const char *what) { return fail(where, what); }

extern "C" {

int surfel_abi_version(void) { return SURFEL_ABI_VERSION; }

int surfel_stage_count(void) { return N_STAGES; }
const char *surfel_stage_name(int stage) { return (stage >= 0 && stage < N_STAGES) ? kStageNames[stage] : ""; }
int surfel_stage_time(int stage, double *total_ms, int *calls)
{
    if (stage < 0 || stage >= N_STAGES || !total_ms || !calls) return fail("surfel_stage_time", "bad arguments");
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    *total_ms = g_stage_ms[stage];
    *calls = g_stage_calls[stage];
    return 0;
}

const char *surfel_last_error(void) { return g_err.c_str(); }

int surfel_set_option(const char *name, int value)
{
    if (name && std::strcmp(name, "subtile_cull") == 0) {
        g_subtile_cull = value;
        return 0;
    }
    if (name && std::strcmp(name, "bwd_variant") == 0) {
        g_bwd_variant = value;
        return 0;
    }
    if (name && std::strcmp(name, "radix_onesweep") == 0) {   // 1 (default): one-launch passes with decoupled look-back
        set_radix_onesweep(value);
        return 0;
    }
    if (name && std::strcmp(name, "time_stages") == 0) {  // (re)arms and clears the stage clocks
        std::lock_guard<std::mutex> lock(g_stage_mutex);
        g_time_stages = value;
        for (int i = 0; i < N_STAGES; i++) { g_stage_ms[i] = 0; g_stage_calls[i] = 0; }
        return 0;
    }
    return fail("surfel_set_option", "unknown option");
}

size_t surfel_geometry_bytes(int P)
{
    if (P < 0) { fail("surfel_geometry_bytes", "negative P"); return 0; }
    if (!have_device()) { fail("surfel_geometry_bytes", "no CUDA device (this library has no CPU path)"); return 0; }
    return geom_end(P, depth_sort_temp_bytes(P));
}

size_t surfel_image_bytes(int width, int height)
{
    if (width <= 0 || height <= 0) { fail("surfel_image_bytes", "bad image size"); return 0; }
    ImageView v = carve_image(nullptr, width, height);
    return (size_t)(v.aux_flag + 1) + 128;
}

size_t surfel_binning_bytes(int64_t num_rendered)
{
    if (num_rendered < 0) { fail("surfel_binning_bytes", "negative num_rendered"); return 0; }
    if (!have_device()) { fail("surfel_binning_bytes", "no CUDA device (this library has no CPU path)"); return 0; }
    const size_t cub = tile_sort_temp_bytes(num_rendered);
    BinView b = carve_bin(nullptr, num_rendered, cub);
    return (size_t)(b.cub_temp + cub) + 128;
}

size_t surfel_grad_scratch_bytes(int P)
{
    if (P < 0) { fail("surfel_grad_scratch_bytes", "negative P"); return 0; }
    return (size_t)(P > 0 ? P : 1) * GACC_FLOATS * sizeof(float) + 256;   // accumulator + one aligned flag word
}

int surfel_forward_prepare(int P, int D, int M, int width, int height, const float *means3D, const float *shs,
                           const float *colors_precomp, const float *opacities, const float *scales,
                           float scale_modifier, const float *rotations, const float *transMat_precomp,
                           const float *viewmatrix, const float *projmatrix, const float *cam_pos, float tan_fovx,
                           float tan_fovy, int prefiltered, int *radii, char *geometry_buffer,
                           int64_t *num_rendered, void *stream, int debug)
{
    (void)tan_fovx; (void)tan_fovy;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!num_rendered) return fail("surfel_forward_prepare", "num_rendered is NULL");
    *num_rendered = 0;
    if (P < 0 || width <= 0 || height <= 0) return fail("surfel_forward_prepare", "bad sizes");
    if (P == 0) return 0;
    if (!means3D || !opacities || !viewmatrix || !projmatrix || !cam_pos || !radii || !geometry_buffer)
        return fail("surfel_forward_prepare", "NULL required pointer");
    if ((shs == nullptr) == (colors_precomp == nullptr))
        return fail("surfel_forward_prepare", "provide exactly one of shs / colors_precomp");
    if (shs && (M <= 0 || D < 0 || D > 3 || (D + 1) * (D + 1) > M))
        return fail("surfel_forward_prepare", "SH degree / coefficient count mismatch");
    const bool have_sr = scales != nullptr && rotations != nullptr;
    if (have_sr == (transMat_precomp != nullptr) || ((scales == nullptr) != (rotations == nullptr)))
        return fail("surfel_forward_prepare", "provide exactly one of (scales, rotations) / transMat_precomp");
    if (have_sr && (!aligned(scales, 8) || !aligned(rotations, 16)))
        return fail("surfel_forward_prepare", "scales must be 8-byte and rotations 16-byte aligned");
    if (!aligned(geometry_buffer, 16)) return fail("surfel_forward_prepare", "geometry_buffer must be 16-byte aligned");

    const size_t cub = depth_sort_temp_bytes(P);
    GeomView g = carve_geom(geometry_buffer, P, cub);

    PreprocessFwdArgs a;
    a.P = P; a.D = D; a.M = M; a.W = width; a.H = height;
    a.gx = (width + TILE_X - 1) / TILE_X; a.gy = (height + TILE_Y - 1) / TILE_Y;
    a.prefiltered = prefiltered; a.scale_modifier = scale_modifier;
    a.means3D = means3D; a.scales = scales; a.rotations = rotations; a.opacities = opacities; a.shs = shs;
    a.transMat_precomp = transMat_precomp; a.colors_precomp = colors_precomp;
    a.viewmatrix = viewmatrix; a.projmatrix = projmatrix; a.cam_pos = cam_pos;
    a.radii = radii; a.rec = g.rec; a.tiles_touched = g.tiles_touched; a.depth_key = g.depth_key;
    a.idx_in = g.idx_in; a.clamped = g.clamped;
    {
        StageClock clk(st, 0);
        launch_preprocess_fwd(a, st);
    }
    STAGE("preprocess");

    StageClock clk_depth(st, 1);
    CK("depth order", run_depth_order(P, g.depth_key, g.depth_key_sorted, g.idx_in, g.idx_sorted, g.tiles_touched,
                                      g.offsets, nullptr, g.cub_temp, g.cub_temp_bytes, st));
    clk_depth.stop();
    STAGE("depth order");

    uint32_t r32 = 0;
    CK("num_rendered readback", read_back_u32(g.offsets + (P - 1), &r32, st));
    *num_rendered = (int64_t)r32;
    return 0;
}

int surfel_forward_render(int P, int width, int height, int64_t num_rendered, const float *background,
                          const int *radii, char *geometry_buffer, char *binning_buffer, char *image_buffer,
                          float *out_color, float *out_others, void *stream, int debug)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P < 0 || width <= 0 || height <= 0 || num_rendered < 0) return fail("surfel_forward_render", "bad sizes");
    if (!background || !image_buffer || !out_color || !out_others)
        return fail("surfel_forward_render", "NULL required pointer");
    if (P > 0 && (!radii || !geometry_buffer)) return fail("surfel_forward_render", "NULL geometry");
    if (num_rendered > 0 && !binning_buffer) return fail("surfel_forward_render", "NULL binning_buffer");

    ImageView iv = carve_image(image_buffer, width, height);
    GeomView g{};
    BinView bv{};
    if (P > 0) g = carve_geom(geometry_buffer, P, depth_sort_temp_bytes(P));
    if (num_rendered > 0) {
        const size_t cub = tile_sort_temp_bytes(num_rendered);
        bv = carve_bin(binning_buffer, num_rendered, cub);
    }
    if (int rc = bin_stage(P, width, height, num_rendered, radii, g, iv, bv, st, debug)) return rc;
    return render_pass(width, height, background, g, iv, bv, out_color, out_others, st, debug);
}

int surfel_backward(int P, int D, int M, int64_t num_rendered, const float *background, int width, int height,
                    const float *means3D, const float *shs, const float *colors_precomp, const float *scales,
                    float scale_modifier, const float *rotations, const float *transMat_precomp,
                    const float *viewmatrix, const float *projmatrix, const float *cam_pos, float tan_fovx,
                    float tan_fovy, const int *radii, char *geometry_buffer, char *binning_buffer,
                    char *image_buffer, const float *dL_dpix, const float *dL_dothers, float *dL_dmean2D,
                    float *dL_dnormal, float *dL_dopacity, float *dL_dcolor, float *dL_dmean3D, float *dL_dtransMat,
                    float *dL_dsh, float *dL_dscale, float *dL_drot, char *grad_scratch, void *stream, int debug)
{
    (void)scale_modifier; (void)colors_precomp;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P < 0 || width <= 0 || height <= 0 || num_rendered < 0) return fail("surfel_backward", "bad sizes");
    if (P == 0) return 0;
    if (!background || !means3D || !viewmatrix || !projmatrix || !cam_pos || !radii || !geometry_buffer ||
        !image_buffer || !dL_dpix || !dL_dothers || !dL_dmean2D || !dL_dopacity || !dL_dcolor || !dL_dmean3D ||
        !dL_dtransMat || !dL_dscale || !dL_drot || !grad_scratch)
        return fail("surfel_backward", "NULL required pointer");
    if (shs && M > 0 && !dL_dsh) return fail("surfel_backward", "dL_dsh is NULL but shs given");
    if (num_rendered > 0 && !binning_buffer) return fail("surfel_backward", "NULL binning_buffer");
    if ((scales == nullptr) != (rotations == nullptr) || ((scales == nullptr) && !transMat_precomp))
        return fail("surfel_backward", "provide (scales, rotations) or transMat_precomp");
    if (!aligned(dL_dscale, 8) || !aligned(dL_drot, 16) || (scales && (!aligned(scales, 8) || !aligned(rotations, 16))))
        return fail("surfel_backward", "scale/rotation buffers must be 8/16-byte aligned");

    GeomView g = carve_geom(geometry_buffer, P, depth_sort_temp_bytes(P));
    ImageView iv = carve_image(image_buffer, width, height);
    BinView bv{};
    if (num_rendered > 0) bv = carve_bin(binning_buffer, num_rendered, tile_sort_temp_bytes(num_rendered));

    if (int rc = backward_blend(P, width, height, num_rendered, background, g, iv, bv, dL_dpix, dL_dothers, grad_scratch,
                                /*clear=*/true, st, debug))
        return rc;
    return backward_geometry(P, D, M, width, height, means3D, shs, scales, rotations, transMat_precomp, viewmatrix,
                             projmatrix, cam_pos, tan_fovx, tan_fovy, radii, g, grad_scratch, dL_dmean2D, dL_dnormal,
                             dL_dopacity, dL_dcolor, dL_dmean3D, dL_dtransMat, dL_dsh, dL_dscale, dL_drot, st, debug);
}

// ---------------------------------------------------------------------------------------------
// Colour passes over one geometry / binning state (see the header)
// ---------------------------------------------------------------------------------------------
int surfel_pass_set_colors(int P, const float *colors, char *geometry_buffer, void *stream)
{
    if (P < 0) return fail("surfel_pass_set_colors", "bad sizes");
    if (P == 0) return 0;
    if (!colors || !geometry_buffer) return fail("surfel_pass_set_colors", "NULL required pointer");
    GeomView g = carve_geom(geometry_buffer, P, depth_sort_temp_bytes(P));
    launch_set_record_colors(P, colors, g.rec, static_cast<cudaStream_t>(stream));
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail_cuda("surfel_pass_set_colors", e);
}

int surfel_pass_render(int P, int width, int height, int64_t num_rendered, const float *background,
                       char *geometry_buffer, char *binning_buffer, char *image_buffer, float *out_color,
                       float *out_others, void *stream, int debug)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P < 0 || width <= 0 || height <= 0 || num_rendered < 0) return fail("surfel_pass_render", "bad sizes");
    if (!background || !image_buffer || !out_color || !out_others) return fail("surfel_pass_render", "NULL required pointer");
    if (P > 0 && !geometry_buffer) return fail("surfel_pass_render", "NULL geometry");
    if (num_rendered > 0 && !binning_buffer) return fail("surfel_pass_render", "NULL binning_buffer");
    ImageView iv = carve_image(image_buffer, width, height);
    GeomView g{};
    BinView bv{};
    if (P > 0) g = carve_geom(geometry_buffer, P, depth_sort_temp_bytes(P));
    if (num_rendered > 0) bv = carve_bin(binning_buffer, num_rendered, tile_sort_temp_bytes(num_rendered));
    return render_pass(width, height, background, g, iv, bv, out_color, out_others, st, debug);
}

int surfel_pass_backward_blend(int P, int width, int height, int64_t num_rendered, const float *background,
                               char *geometry_buffer, char *binning_buffer, char *image_buffer, const float *dL_dpix,
                               const float *dL_dothers, char *grad_scratch, int clear, void *stream, int debug)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P < 0 || width <= 0 || height <= 0 || num_rendered < 0) return fail("surfel_pass_backward_blend", "bad sizes");
    if (P == 0) return 0;
    if (!background || !geometry_buffer || !image_buffer || !dL_dpix || !dL_dothers || !grad_scratch)
        return fail("surfel_pass_backward_blend", "NULL required pointer");
    if (num_rendered > 0 && !binning_buffer) return fail("surfel_pass_backward_blend", "NULL binning_buffer");
    GeomView g = carve_geom(geometry_buffer, P, depth_sort_temp_bytes(P));
    ImageView iv = carve_image(image_buffer, width, height);
    BinView bv{};
    if (num_rendered > 0) bv = carve_bin(binning_buffer, num_rendered, tile_sort_temp_bytes(num_rendered));
    return backward_blend(P, width, height, num_rendered, background, g, iv, bv, dL_dpix, dL_dothers, grad_scratch,
                          clear != 0, st, debug);
}

int surfel_pass_take_color_grad(int P, char *grad_scratch, float *dL_dcolor, void *stream)
{
    if (P < 0) return fail("surfel_pass_take_color_grad", "bad sizes");
    if (P == 0) return 0;
    if (!grad_scratch) return fail("surfel_pass_take_color_grad", "NULL required pointer");
    char *gp = grad_scratch;
    float *gacc = carve<float>(gp, (size_t)P * GACC_FLOATS);
    launch_take_color_grad(P, gacc, dL_dcolor, static_cast<cudaStream_t>(stream));
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail_cuda("surfel_pass_take_color_grad", e);
}

int surfel_pass_backward_geometry(int P, int width, int height, const float *means3D, const float *scales,
                                  const float *rotations, const float *transMat_precomp, const float *viewmatrix,
                                  const float *projmatrix, const float *cam_pos, float tan_fovx, float tan_fovy,
                                  const int *radii, char *geometry_buffer, char *grad_scratch, float *dL_dmean2D,
                                  float *dL_dnormal, float *dL_dopacity, float *dL_dcolor, float *dL_dmean3D,
                                  float *dL_dtransMat, float *dL_dscale, float *dL_drot, void *stream, int debug)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const char *where = "surfel_pass_backward_geometry";
    if (P < 0 || width <= 0 || height <= 0) return fail(where, "bad sizes");
    if (P == 0) return 0;
    if (!means3D || !viewmatrix || !projmatrix || !cam_pos || !radii || !geometry_buffer || !grad_scratch ||
        !dL_dmean2D || !dL_dopacity || !dL_dcolor || !dL_dmean3D || !dL_dtransMat || !dL_dscale || !dL_drot)
        return fail(where, "NULL required pointer");
    if ((scales == nullptr) != (rotations == nullptr) || ((scales == nullptr) && !transMat_precomp))
        return fail(where, "provide (scales, rotations) or transMat_precomp");
    if (!aligned(dL_dscale, 8) || !aligned(dL_drot, 16) || (scales && (!aligned(scales, 8) || !aligned(rotations, 16))))
        return fail(where, "scale/rotation buffers must be 8/16-byte aligned");
    GeomView g = carve_geom(geometry_buffer, P, depth_sort_temp_bytes(P));
    return backward_geometry(P, 0, 0, width, height, means3D, nullptr, scales, rotations, transMat_precomp, viewmatrix,
                             projmatrix, cam_pos, tan_fovx, tan_fovy, radii, g, grad_scratch, dL_dmean2D, dL_dnormal,
                             dL_dopacity, dL_dcolor, dL_dmean3D, dL_dtransMat, nullptr, dL_dscale, dL_drot, st, debug);
}

// ---------------------------------------------------------------------------------------------
// Class-probability pass (see the header)
// ---------------------------------------------------------------------------------------------
int surfel_forward_bin(int P, int width, int height, int64_t num_rendered, const int *radii, char *geometry_buffer,
                       char *binning_buffer, char *image_buffer, void *stream, int debug)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P < 0 || width <= 0 || height <= 0 || num_rendered < 0) return fail("surfel_forward_bin", "bad sizes");
    if (!image_buffer) return fail("surfel_forward_bin", "NULL required pointer");
    if (P > 0 && (!radii || !geometry_buffer)) return fail("surfel_forward_bin", "NULL geometry");
    if (num_rendered > 0 && !binning_buffer) return fail("surfel_forward_bin", "NULL binning_buffer");
    ImageView iv = carve_image(image_buffer, width, height);
    GeomView g{};
    BinView bv{};
    if (P > 0) g = carve_geom(geometry_buffer, P, depth_sort_temp_bytes(P));
    if (num_rendered > 0) bv = carve_bin(binning_buffer, num_rendered, tile_sort_temp_bytes(num_rendered));
    return bin_stage(P, width, height, num_rendered, radii, g, iv, bv, st, debug);
}

int surfel_classes_set_labels(int P, const int *labels, char *geometry_buffer, void *stream)
{
    if (P < 0) return fail("surfel_classes_set_labels", "bad sizes");
    if (P == 0) return 0;
    if (!labels || !geometry_buffer) return fail("surfel_classes_set_labels", "NULL required pointer");
    GeomView g = carve_geom(geometry_buffer, P, depth_sort_temp_bytes(P));
    launch_set_record_labels(P, labels, g.rec, static_cast<cudaStream_t>(stream));
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail_cuda("surfel_classes_set_labels", e);
}

int surfel_classes_render(int P, int width, int height, int64_t num_rendered, int n_classes, const float *background,
                          char *geometry_buffer, char *binning_buffer, char *image_buffer, float *out_probs,
                          void *stream, int debug)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P < 0 || width <= 0 || height <= 0 || num_rendered < 0) return fail("surfel_classes_render", "bad sizes");
    if (n_classes < 1 || n_classes > MAX_CLASSES) return fail("surfel_classes_render", "between 1 and 8 classes per pass");
    if (!background || !image_buffer || !out_probs) return fail("surfel_classes_render", "NULL required pointer");
    if (P > 0 && !geometry_buffer) return fail("surfel_classes_render", "NULL geometry");
    if (num_rendered > 0 && !binning_buffer) return fail("surfel_classes_render", "NULL binning_buffer");
    ImageView iv = carve_image(image_buffer, width, height);
    GeomView g{};
    BinView bv{};
    if (P > 0) g = carve_geom(geometry_buffer, P, depth_sort_temp_bytes(P));
    if (num_rendered > 0) bv = carve_bin(binning_buffer, num_rendered, tile_sort_temp_bytes(num_rendered));
    return render_pass(width, height, background, g, iv, bv, out_probs, nullptr, st, debug, n_classes);
}

int surfel_classes_backward_blend(int P, int width, int height, int64_t num_rendered, int n_classes,
                                  const float *background, char *geometry_buffer, char *binning_buffer,
                                  char *image_buffer, const float *dL_dprobs, char *grad_scratch, int clear,
                                  void *stream, int debug)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P < 0 || width <= 0 || height <= 0 || num_rendered < 0) return fail("surfel_classes_backward_blend", "bad sizes");
    if (n_classes < 1 || n_classes > MAX_CLASSES) return fail("surfel_classes_backward_blend", "between 1 and 8 classes per pass");
    if (P == 0) return 0;
    if (!background || !geometry_buffer || !image_buffer || !dL_dprobs || !grad_scratch)
        return fail("surfel_classes_backward_blend", "NULL required pointer");
    if (num_rendered > 0 && !binning_buffer) return fail("surfel_classes_backward_blend", "NULL binning_buffer");
    GeomView g = carve_geom(geometry_buffer, P, depth_sort_temp_bytes(P));
    ImageView iv = carve_image(image_buffer, width, height);
    BinView bv{};
    if (num_rendered > 0) bv = carve_bin(binning_buffer, num_rendered, tile_sort_temp_bytes(num_rendered));
    return backward_blend(P, width, height, num_rendered, background, g, iv, bv, dL_dprobs, nullptr, grad_scratch,
                          clear != 0, st, debug, n_classes);
}

int surfel_mark_visible(int P, const float *means3D, const float *viewmatrix, const float *projmatrix,
                        unsigned char *present, void *stream)
{
    (void)projmatrix;
    if (P < 0) return fail("surfel_mark_visible", "negative P");
    if (P == 0) return 0;
    if (!means3D || !viewmatrix || !present) return fail("surfel_mark_visible", "NULL required pointer");
    launch_mark_visible(P, means3D, viewmatrix, present, static_cast<cudaStream_t>(stream));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda("surfel_mark_visible", e);
    return 0;
}

int surfel_debug_copy_binning(int width, int height, int64_t num_rendered, const char *binning_buffer,
                              const char *image_buffer, uint32_t *ranges_out, uint32_t *point_list_out, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (width <= 0 || height <= 0 || num_rendered < 0 || !image_buffer)
        return fail("surfel_debug_copy_binning", "bad arguments");
    const size_t tiles = (size_t)((width + TILE_X - 1) / TILE_X) * ((height + TILE_Y - 1) / TILE_Y);
    ImageView iv = carve_image(const_cast<char *>(image_buffer), width, height);
    if (ranges_out)
        CK("copy ranges", cudaMemcpyAsync(ranges_out, iv.ranges, tiles * sizeof(uint2), cudaMemcpyDeviceToDevice, st));
    if (point_list_out && num_rendered > 0) {
        if (!binning_buffer) return fail("surfel_debug_copy_binning", "NULL binning_buffer");
        BinView bv = carve_bin(const_cast<char *>(binning_buffer), num_rendered, tile_sort_temp_bytes(num_rendered));
        CK("copy point list", cudaMemcpyAsync(point_list_out, bv.point_list, (size_t)num_rendered * sizeof(uint32_t),
                                              cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

// =====================================================================================================
// Sharded path (DESIGN.md "Multi-GPU"): Gaussians sharded by index, screen partitioned by tile rows.
// =====================================================================================================
int surfel_shard_preprocess(int P, int D, int M, int width, int height, const float *means3D, const float *shs,
                            const float *colors_precomp, const float *opacities, const float *scales,
                            float scale_modifier, const float *rotations, const float *transMat_precomp,
                            const float *viewmatrix, const float *projmatrix, const float *cam_pos, float tan_fovx,
                            float tan_fovy, int prefiltered, int *radii, float *records, uint32_t *depth_keys,
                            unsigned char *clamped, void *stream)
{
    (void)tan_fovx; (void)tan_fovy;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P < 0 || width <= 0 || height <= 0) return fail("surfel_shard_preprocess", "bad sizes");
    if (P == 0) return 0;
    if (!means3D || !opacities || !viewmatrix || !projmatrix || !cam_pos || !radii || !records || !depth_keys || !clamped)
        return fail("surfel_shard_preprocess", "NULL required pointer");
    if ((shs == nullptr) == (colors_precomp == nullptr))
        return fail("surfel_shard_preprocess", "provide exactly one of shs / colors_precomp");
    const bool have_sr = scales != nullptr && rotations != nullptr;
    if (have_sr == (transMat_precomp != nullptr))
        return fail("surfel_shard_preprocess", "provide exactly one of (scales, rotations) / transMat_precomp");
    if (!aligned(records, 16) || (have_sr && (!aligned(scales, 8) || !aligned(rotations, 16))))
        return fail("surfel_shard_preprocess", "alignment");
    PreprocessFwdArgs a;
    a.P = P; a.D = D; a.M = M; a.W = width; a.H = height;
    a.gx = (width + TILE_X - 1) / TILE_X; a.gy = (height + TILE_Y - 1) / TILE_Y;
    a.prefiltered = prefiltered; a.scale_modifier = scale_modifier;
    a.means3D = means3D; a.scales = scales; a.rotations = rotations; a.opacities = opacities; a.shs = shs;
    a.transMat_precomp = transMat_precomp; a.colors_precomp = colors_precomp;
    a.viewmatrix = viewmatrix; a.projmatrix = projmatrix; a.cam_pos = cam_pos;
    a.radii = radii; a.rec = records; a.tiles_touched = nullptr; a.depth_key = depth_keys; a.idx_in = nullptr;
    a.clamped = clamped;
    launch_preprocess_fwd(a, st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda("surfel_shard_preprocess", e);
    return 0;
}

size_t surfel_shard_compact_bytes(int P)
{
    if (P < 0) { fail("surfel_shard_compact_bytes", "negative P"); return 0; }
    return compact_temp_bytes(P);
}

int surfel_shard_compact(int P, const int *radii, const float *records, const uint32_t *depth_keys, float *records_c,
                         int *radii_c, uint32_t *depth_keys_c, uint32_t *slot, int *count_dev, char *temp,
                         void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P < 0) return fail("surfel_shard_compact", "bad sizes");
    if (!count_dev) return fail("surfel_shard_compact", "count_dev is NULL");
    if (P == 0) {
        cudaError_t e0 = cudaMemsetAsync(count_dev, 0, sizeof(int), st);
        if (e0 != cudaSuccess) return fail_cuda("surfel_shard_compact", e0);
        return 0;
    }
    if (!radii || !records || !depth_keys || !records_c || !radii_c || !depth_keys_c || !slot || !temp)
        return fail("surfel_shard_compact", "NULL required pointer");
    if (!aligned(records, 16) || !aligned(records_c, 16)) return fail("surfel_shard_compact", "alignment");
    cudaError_t e = run_compact_visible(P, radii, records, depth_keys, records_c, radii_c, depth_keys_c, slot,
                                        count_dev, temp, compact_temp_bytes(P), st);
    if (e != cudaSuccess) return fail_cuda("surfel_shard_compact", e);
    return 0;
}

int surfel_shard_tile_hist(int P, int width, int height, const float *records, const int *radii, uint32_t *hist,
                           void *stream)
{
    if (P < 0 || width <= 0 || height <= 0 || !hist) return fail("surfel_shard_tile_hist", "bad arguments");
    if (P > 0 && (!records || !radii)) return fail("surfel_shard_tile_hist", "NULL required pointer");
    const int gx = (width + TILE_X - 1) / TILE_X, gy = (height + TILE_Y - 1) / TILE_Y;
    launch_tile_hist(P, gx, gy, records, radii, hist, static_cast<cudaStream_t>(stream));
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail_cuda("surfel_shard_tile_hist", e);
}

size_t surfel_shard_partition_bytes(int width, int height)
{
    if (width <= 0 || height <= 0) { fail("surfel_shard_partition_bytes", "bad image size"); return 0; }
    return partition_temp_bytes(((width + TILE_X - 1) / TILE_X) * ((height + TILE_Y - 1) / TILE_Y));
}

int surfel_shard_partition(int width, int height, int G, const uint32_t *hist, int cost_base, const float *shares,
                           char *temp, int *cuts, int64_t *window_num_rendered, void *stream)
{
    if (width <= 0 || height <= 0 || G < 1 || G > MAX_RANKS || cost_base < 0)
        return fail("surfel_shard_partition", "bad sizes (1 <= G <= 16)");
    if (!hist || !temp || !cuts || !window_num_rendered) return fail("surfel_shard_partition", "NULL required pointer");
    const int ntiles = ((width + TILE_X - 1) / TILE_X) * ((height + TILE_Y - 1) / TILE_Y);
    launch_partition(ntiles, G, hist, (uint32_t)cost_base, shares, temp, cuts, reinterpret_cast<long long *>(window_num_rendered),
                     static_cast<cudaStream_t>(stream));
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail_cuda("surfel_shard_partition", e);
}

size_t surfel_shard_route_bytes(int P, int G)
{
    if (P < 0 || G < 1 || G > MAX_RANKS) { fail("surfel_shard_route_bytes", "bad sizes"); return 0; }
    return route_temp_bytes(P, G);
}

int surfel_shard_route_count(int P, int width, int height, int G, const float *records, const int *radii,
                             const int *cuts, char *temp, int *send_counts, int extra, void *stream)
{
    if (P < 0 || width <= 0 || height <= 0 || G < 1 || G > MAX_RANKS) return fail("surfel_shard_route_count", "bad sizes");
    if (!cuts || !send_counts || (P > 0 && (!records || !radii || !temp)))
        return fail("surfel_shard_route_count", "NULL required pointer");
    const int gx = (width + TILE_X - 1) / TILE_X, gy = (height + TILE_Y - 1) / TILE_Y;
    const cudaError_t e = run_route_count(P, gx, gy, G, records, radii, cuts, temp, send_counts, extra,
                                          static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda("surfel_shard_route_count", e);
}

int surfel_shard_route_scatter(int P, int G, const float *records, const int *radii, const uint32_t *depth_keys,
                               char *temp, const int *send_counts, float *send_rows, uint32_t *send_src, void *stream)
{
    if (P < 0 || G < 1 || G > MAX_RANKS) return fail("surfel_shard_route_scatter", "bad sizes");
    if (P == 0) return 0;
    if (!records || !radii || !depth_keys || !temp || !send_counts) return fail("surfel_shard_route_scatter", "NULL required pointer");
    if (!aligned(records, 16) || (send_rows && !aligned(send_rows, 16))) return fail("surfel_shard_route_scatter", "alignment");
    const cudaError_t e = run_route_scatter(P, G, records, radii, depth_keys, temp, send_counts, send_rows, send_src,
                                            static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda("surfel_shard_route_scatter", e);
}

int surfel_shard_route_scatter_peers(int P, int G, const float *records, const int *radii, const uint32_t *depth_keys,
                                     char *temp, const int *send_counts, float *const *dst_records,
                                     uint32_t *const *dst_depth_keys, int *const *dst_radii, const int64_t *dst_row0,
                                     uint32_t *send_src, void *stream)
{
    const char *where = "surfel_shard_route_scatter_peers";
    if (P < 0 || G < 1 || G > MAX_RANKS) return fail(where, "bad sizes");
    if (P == 0) return 0;
    if (!records || !radii || !depth_keys || !temp || !send_counts || !dst_records || !dst_depth_keys || !dst_radii ||
        !dst_row0 || !send_src)
        return fail(where, "NULL required pointer");
    long long row0[MAX_RANKS];
    for (int d = 0; d < G; d++) {
        if (!dst_records[d] || !dst_depth_keys[d] || !dst_radii[d] || dst_row0[d] < 0) return fail(where, "bad destination");
        if (!aligned(dst_records[d], 16)) return fail(where, "destination records must be 16-byte aligned");
        row0[d] = (long long)dst_row0[d];
    }
    if (!aligned(records, 16)) return fail(where, "alignment");
    const cudaError_t e = run_route_scatter_peers(P, G, records, radii, depth_keys, temp, send_counts, dst_records,
                                                  dst_depth_keys, dst_radii, row0, send_src,
                                                  static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda(where, e);
}

int surfel_window_push_grad_rows(int64_t n_rows, const float *grad_rows, int G, const int64_t *seg_count,
                                 float *const *dst_rows, const int64_t *dst_row0, void *stream)
{
    const char *where = "surfel_window_push_grad_rows";
    if (n_rows < 0 || G < 1 || G > MAX_RANKS) return fail(where, "bad sizes");
    if (n_rows == 0) return 0;
    if (!grad_rows || !seg_count || !dst_rows || !dst_row0) return fail(where, "NULL required pointer");
    long long cnt[MAX_RANKS], row0[MAX_RANKS];
    for (int s = 0; s < G; s++) {
        if (seg_count[s] < 0 || dst_row0[s] < 0 || (seg_count[s] > 0 && (!dst_rows[s] || !aligned(dst_rows[s], 16))))
            return fail(where, "bad destination");
        cnt[s] = (long long)seg_count[s];
        row0[s] = (long long)dst_row0[s];
    }
    if (!aligned(grad_rows, 16)) return fail(where, "alignment");
    const cudaError_t e = run_push_grad_rows((long long)n_rows, grad_rows, G, cnt, dst_rows, row0,
                                             static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda(where, e);
}

int surfel_window_unpack(int n, const float *rows, float *records, int *radii, uint32_t *depth_keys, void *stream)
{
    if (n < 0) return fail("surfel_window_unpack", "bad sizes");
    if (n == 0) return 0;
    if (!rows || !records || !radii || !depth_keys) return fail("surfel_window_unpack", "NULL required pointer");
    if (!aligned(rows, 16) || !aligned(records, 16)) return fail("surfel_window_unpack", "alignment");
    launch_unpack_rows(n, rows, records, depth_keys, radii, static_cast<cudaStream_t>(stream));
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail_cuda("surfel_window_unpack", e);
}

int surfel_shard_grad_accumulate(int P, int64_t n_rows, const float *grad_rows, const uint32_t *send_src,
                                 float *grad_records, void *stream)
{
    if (P < 0 || n_rows < 0) return fail("surfel_shard_grad_accumulate", "bad sizes");
    if (P == 0) return 0;
    if (!grad_records || (n_rows > 0 && (!grad_rows || !send_src))) return fail("surfel_shard_grad_accumulate", "NULL required pointer");
    if (!aligned(grad_records, 16) || (grad_rows && !aligned(grad_rows, 16))) return fail("surfel_shard_grad_accumulate", "alignment");
    const cudaError_t e = run_grad_accumulate(P, (long long)n_rows, grad_rows, send_src, grad_records,
                                              static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? 0 : fail_cuda("surfel_shard_grad_accumulate", e);
}

size_t surfel_window_bytes(int P_total)
{
    if (P_total < 0) { fail("surfel_window_bytes", "negative P"); return 0; }
    if (!have_device()) { fail("surfel_window_bytes", "no CUDA device (this library has no CPU path)"); return 0; }
    const size_t cub = depth_sort_temp_bytes(P_total);
    WindowView w = carve_window(nullptr, P_total, cub);
    return (size_t)(w.cub_temp + cub) + 128;
}

int surfel_window_prepare(int P_total, int width, int height, int tile_lo, int tile_hi, const float *records,
                          const int *radii, const uint32_t *depth_keys, char *window_buffer, int64_t *num_rendered,
                          void *stream, int debug)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (num_rendered) *num_rendered = 0;
    if (P_total < 0 || width <= 0 || height <= 0 || tile_lo < 0 || tile_hi < tile_lo)
        return fail("surfel_window_prepare", "bad sizes");
    if (P_total == 0) return 0;
    if (!records || !radii || !depth_keys || !window_buffer) return fail("surfel_window_prepare", "NULL required pointer");
    const int gx = (width + TILE_X - 1) / TILE_X, gy = (height + TILE_Y - 1) / TILE_Y;
    WindowView w = carve_window(window_buffer, P_total, depth_sort_temp_bytes(P_total));
    if (tile_hi > gx * gy) return fail("surfel_window_prepare", "tile range exceeds the tile grid");
    launch_count_window_tiles(P_total, gx, gy, tile_lo, tile_hi, records, radii, w.tiles_touched, w.idx_in, st);
    STAGE("window tile count");
    CK("depth order", run_depth_order(P_total, depth_keys, w.depth_key_sorted, w.idx_in, w.idx_sorted, w.tiles_touched,
                                      w.offsets, nullptr, w.cub_temp, w.cub_temp_bytes, st));
    STAGE("depth order");
    if (num_rendered) {
        uint32_t r32 = 0;
        CK("num_rendered readback", read_back_u32(w.offsets + (P_total - 1), &r32, st));
        *num_rendered = (int64_t)r32;
    }
    return 0;
}

static int window_render_impl(const char *where, int P_total, int width, int height, int tile_lo, int tile_hi,
                              int64_t num_rendered, const float *background, const float *records, const int *radii,
                              char *window_buffer, char *binning_buffer, char *image_buffer, float *out_color,
                              float *out_others, int n_peers, float *const *peer_planes, int multicast, void *stream,
                              int debug)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P_total < 0 || width <= 0 || height <= 0 || num_rendered < 0 || tile_lo < 0 || tile_hi < tile_lo)
        return fail(where, "bad sizes");
    if (!background || !image_buffer) return fail(where, "NULL required pointer");
    if (n_peers == 0 && (!out_color || !out_others)) return fail(where, "NULL required pointer");
    if (n_peers < 0 || n_peers > MAX_RANKS || (n_peers > 0 && !peer_planes)) return fail(where, "bad peer list");
    if (P_total > 0 && (!records || !radii || !window_buffer)) return fail(where, "NULL geometry");
    if (num_rendered > 0 && !binning_buffer) return fail(where, "NULL binning_buffer");
    const int gx = (width + TILE_X - 1) / TILE_X, gy = (height + TILE_Y - 1) / TILE_Y;
    if (tile_hi > gx * gy) return fail(where, "tile range exceeds the tile grid");
    ImageView iv = carve_image(image_buffer, width, height);
    WindowView w{};
    BinView bv{};
    if (P_total > 0) w = carve_window(window_buffer, P_total, depth_sort_temp_bytes(P_total));
    if (num_rendered > 0) bv = carve_bin(binning_buffer, num_rendered, tile_sort_temp_bytes(num_rendered));
    CK("tile binning", run_tile_binning(P_total, num_rendered, gx, gy, tile_lo, tile_hi, records, radii,
                                        w.idx_sorted, w.offsets, bv.keys_unsorted, bv.vals_unsorted, bv.keys_sorted,
                                        bv.point_list, iv.ranges, iv.tile_order, bv.cub_temp, bv.cub_temp_bytes, st));
    STAGE("tile binning");
    RenderFwdArgs r;
    r.W = width; r.H = height; r.gx = gx; r.gy = gy; r.tile_lo = tile_lo; r.tile_hi = tile_hi;
    r.ranges = iv.ranges; r.tile_order = iv.tile_order; r.point_list = bv.point_list; r.rec = records; r.bg = background;
    r.final_T = iv.final_T; r.n_contrib = iv.n_contrib; r.tile_max_contrib = iv.tile_max_contrib;
    r.out_color = out_color; r.out_others = out_others; r.subtile_cull = g_subtile_cull;
    r.n_peers = n_peers; r.peer_multicast = multicast;
    for (int q = 0; q < n_peers; q++) {
        if (!peer_planes[q] || !aligned(peer_planes[q], 4)) return fail(where, "bad peer plane pointer");
        r.peer_planes[q] = peer_planes[q];
    }
    launch_render_fwd(r, st);
    STAGE("render forward");
    return 0;
}

int surfel_window_render(int P_total, int width, int height, int tile_lo, int tile_hi, int64_t num_rendered,
                         const float *background, const float *records, const int *radii, char *window_buffer,
                         char *binning_buffer, char *image_buffer, float *out_color, float *out_others, void *stream,
                         int debug)
{
    return window_render_impl("surfel_window_render", P_total, width, height, tile_lo, tile_hi, num_rendered, background,
                              records, radii, window_buffer, binning_buffer, image_buffer, out_color, out_others, 0,
                              nullptr, 0, stream, debug);
}

int surfel_window_render_peers(int P_total, int width, int height, int tile_lo, int tile_hi, int64_t num_rendered,
                               const float *background, const float *records, const int *radii, char *window_buffer,
                               char *binning_buffer, char *image_buffer, int n_peers, float *const *peer_planes,
                               int multicast, void *stream, int debug)
{
    if (n_peers < 1) return fail("surfel_window_render_peers", "needs at least one destination");
    return window_render_impl("surfel_window_render_peers", P_total, width, height, tile_lo, tile_hi, num_rendered,
                              background, records, radii, window_buffer, binning_buffer, image_buffer, nullptr, nullptr,
                              n_peers, peer_planes, multicast, stream, debug);
}

int surfel_window_backward(int P_total, int width, int height, int tile_lo, int tile_hi, int64_t num_rendered,
                           const float *background, const float *records, char *binning_buffer, char *image_buffer,
                           const float *dL_dpix, const float *dL_dothers, float *grad_records, void *stream, int debug)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P_total < 0 || width <= 0 || height <= 0 || num_rendered < 0 || tile_lo < 0 || tile_hi < tile_lo)
        return fail("surfel_window_backward", "bad sizes");
    if (P_total == 0) return 0;
    if (!background || !records || !image_buffer || !dL_dpix || !dL_dothers || !grad_records)
        return fail("surfel_window_backward", "NULL required pointer");
    if (!aligned(grad_records, 16)) return fail("surfel_window_backward", "grad_records must be 16-byte aligned");
    CK("grad records clear", cudaMemsetAsync(grad_records, 0, (size_t)P_total * GACC_FLOATS * sizeof(float), st));
    if (num_rendered == 0) return 0;
    if (!binning_buffer) return fail("surfel_window_backward", "NULL binning_buffer");
    const int gx = (width + TILE_X - 1) / TILE_X, gy = (height + TILE_Y - 1) / TILE_Y;
    ImageView iv = carve_image(image_buffer, width, height);
    BinView bv = carve_bin(binning_buffer, num_rendered, tile_sort_temp_bytes(num_rendered));
    RenderBwdArgs r;
    r.W = width; r.H = height; r.gx = gx; r.gy = gy; r.tile_lo = tile_lo; r.tile_hi = tile_hi;
    r.ranges = iv.ranges; r.tile_order = iv.tile_order; r.point_list = bv.point_list; r.rec = records; r.bg = background;
    r.final_T = iv.final_T; r.n_contrib = iv.n_contrib; r.tile_max_contrib = iv.tile_max_contrib;
    r.dL_dpix = dL_dpix; r.dL_dothers = dL_dothers; r.gacc = grad_records; r.subtile_cull = g_subtile_cull;
    r.aux_flag = iv.aux_flag; r.variant = g_bwd_variant;
    launch_render_bwd(r, st);
    STAGE("render backward");
    return 0;
}

int surfel_shard_backward(int P, int D, int M, int width, int height, const float *means3D, const float *shs,
                          const float *scales, const float *rotations, const float *transMat_precomp,
                          const float *viewmatrix, const float *projmatrix, const float *cam_pos, float tan_fovx,
                          float tan_fovy, const int *radii, const float *records, const unsigned char *clamped,
                          const float *grad_records, const uint32_t *grad_slot, float *dL_dmean2D, float *dL_dnormal,
                          float *dL_dopacity, float *dL_dcolor, float *dL_dmean3D, float *dL_dtransMat, float *dL_dsh,
                          float *dL_dscale, float *dL_drot, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P < 0 || width <= 0 || height <= 0) return fail("surfel_shard_backward", "bad sizes");
    if (P == 0) return 0;
    if (!means3D || !viewmatrix || !projmatrix || !cam_pos || !radii || !records || !clamped || !grad_records ||
        !dL_dmean2D || !dL_dopacity || !dL_dcolor || !dL_dmean3D || !dL_dtransMat || !dL_dscale || !dL_drot)
        return fail("surfel_shard_backward", "NULL required pointer");
    if (shs && M > 0 && !dL_dsh) return fail("surfel_shard_backward", "dL_dsh is NULL but shs given");
    if (!aligned(grad_records, 16) || !aligned(dL_dscale, 8) || !aligned(dL_drot, 16))
        return fail("surfel_shard_backward", "alignment");
    PreprocessBwdArgs a;
    a.P = P; a.D = D; a.M = M;
    a.focal_y = height / (2.0f * tan_fovy);
    a.focal_x = width / (2.0f * tan_fovx);
    a.tan_fovx = tan_fovx; a.tan_fovy = tan_fovy;
    a.means3D = means3D; a.shs = shs; a.scales = scales; a.rotations = rotations;
    a.transMat_precomp = transMat_precomp; a.viewmatrix = viewmatrix; a.projmatrix = projmatrix; a.cam_pos = cam_pos;
    a.radii = radii; a.clamped = clamped; a.rec = records; a.gacc = grad_records; a.gacc_slot = grad_slot;
    a.dL_dmean2D = dL_dmean2D; a.dL_dnormal = dL_dnormal; a.dL_dopacity = dL_dopacity; a.dL_dcolor = dL_dcolor;
    a.dL_dmean3D = dL_dmean3D; a.dL_dtransMat = dL_dtransMat; a.dL_dsh = dL_dsh; a.dL_dscale = dL_dscale;
    a.dL_drot = dL_drot;
    launch_preprocess_bwd(a, st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda("surfel_shard_backward", e);
    return 0;
}

int surfel_debug_sort_pairs(int64_t n, int end_bit, const uint32_t *keys_in, const uint32_t *vals_in,
                            uint32_t *keys_out, uint32_t *vals_out, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n < 0 || end_bit < 1 || end_bit > 32) return fail("surfel_debug_sort_pairs", "bad arguments");
    if (n == 0) return 0;
    if (!keys_in || !vals_in || !keys_out || !vals_out) return fail("surfel_debug_sort_pairs", "NULL pointer");
    char *temp = nullptr;
    const size_t bytes = radix_sort_temp_bytes(n);
    CK("temp alloc", cudaMalloc(&temp, bytes));
    cudaError_t e = radix_sort_pairs(keys_in, vals_in, keys_out, vals_out, n, end_bit, temp, bytes, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(temp);
    if (e != cudaSuccess) return fail_cuda("surfel_debug_sort_pairs", e);
    return 0;
}

int surfel_debug_aux_flag(int P, const char *grad_scratch, int *flag_host, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P <= 0 || !grad_scratch || !flag_host) return fail("surfel_debug_aux_flag", "bad arguments");
    char *gp = const_cast<char *>(grad_scratch);
    carve<float>(gp, (size_t)P * GACC_FLOATS);
    const int *flag = carve<int>(gp, 1);
    CK("copy flag", cudaMemcpyAsync(flag_host, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK("copy flag", cudaStreamSynchronize(st));
    return 0;
}

int surfel_debug_copy_geometry(int P, const char *geometry_buffer, uint32_t *tiles_touched_out,
                               uint32_t *idx_sorted_out, uint32_t *offsets_out, float *records_out, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P <= 0 || !geometry_buffer) return fail("surfel_debug_copy_geometry", "bad arguments");
    GeomView g = carve_geom(const_cast<char *>(geometry_buffer), P, depth_sort_temp_bytes(P));
    const size_t n = (size_t)P;
    if (tiles_touched_out)
        CK("copy tiles", cudaMemcpyAsync(tiles_touched_out, g.tiles_touched, n * 4, cudaMemcpyDeviceToDevice, st));
    if (idx_sorted_out)
        CK("copy idx", cudaMemcpyAsync(idx_sorted_out, g.idx_sorted, n * 4, cudaMemcpyDeviceToDevice, st));
    if (offsets_out) CK("copy offsets", cudaMemcpyAsync(offsets_out, g.offsets, n * 4, cudaMemcpyDeviceToDevice, st));
    if (records_out)
        CK("copy records", cudaMemcpyAsync(records_out, g.rec, n * REC_BYTES, cudaMemcpyDeviceToDevice, st));
    return 0;
}

}  // extern "C"
