"""Host-side stand-in for the reference's pybind module ``diff_surfel_rasterization._C``.

Same three entry points, argument order and return tuples as RAST/ext.cpp:15-19 /
RAST/rasterize_points.cu:39-134 (forward), :136-233 (backward), :235-254 (mark_visible) -- but
implemented over the C ABI of libsurfel_b200.so (include/surfel_rasterizer.h).  torch is used only
to own device memory and to name the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib

NUM_CHANNELS = 3  # RAST/cuda_rasterizer/config.h:15

# tests only: when KEEP_LAST is set, the scratch buffers of the most recent forward stay reachable here
KEEP_LAST = False
LAST = None
LAST_BWD_AUX = None   # tests only (KEEP_LAST): 1 if the last backward ran the full blend specialisation, 0 if colour+alpha


def _dev_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:  # same check (and wording) as CHECK_INPUT, rasterize_points.cu:27-29
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.numel() and t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")
    return t.contiguous()


def _ptr(t):
    if t is None or t.numel() == 0:
        return None
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, transMat_precomp,
                        viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                        prefiltered, debug):
    """-> (num_rendered, out_color[3,H,W], out_others[7,H,W], radii[P] int32, geomBuffer, binningBuffer, imgBuffer)"""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:61-63
    L = _lib.lib()
    background = _dev_f32(background, "background")
    means3D = _dev_f32(means3D, "means3D")
    colors = _dev_f32(colors, "colors")
    opacity = _dev_f32(opacity, "opacity")
    scales = _dev_f32(scales, "scales")
    rotations = _dev_f32(rotations, "rotations")
    transMat_precomp = _dev_f32(transMat_precomp, "transMat_precomp")
    viewmatrix = _dev_f32(viewmatrix, "viewmatrix")
    projmatrix = _dev_f32(projmatrix, "projmatrix")
    sh = _dev_f32(sh, "sh")
    campos = _dev_f32(campos, "campos")

    P, H, W = int(means3D.size(0)), int(image_height), int(image_width)
    dev = means3D.device
    with torch.cuda.device(dev):
        out_color = torch.empty((NUM_CHANNELS, H, W), dtype=torch.float32, device=dev)
        out_others = torch.empty((7, H, W), dtype=torch.float32, device=dev)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        img_buf = torch.empty((_lib.size(L.surfel_image_bytes(W, H), "surfel_image_bytes"),), **u8)
        geom_buf = torch.empty((_lib.size(L.surfel_geometry_bytes(P), "surfel_geometry_bytes") if P else 0,), **u8)
        M = int(sh.size(1)) if sh.numel() != 0 else 0
        st = _stream()
        R = C.c_int64(0)
        if P:
            _lib.check(L.surfel_forward_prepare(
                P, int(degree), M, W, H, _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(opacity), _ptr(scales),
                float(scale_modifier), _ptr(rotations), _ptr(transMat_precomp), _ptr(viewmatrix), _ptr(projmatrix),
                _ptr(campos), float(tan_fovx), float(tan_fovy), int(bool(prefiltered)), _ptr(radii), _ptr(geom_buf),
                C.byref(R), st, int(bool(debug))), "surfel_forward_prepare")
        num_rendered = int(R.value)
        bin_buf = torch.empty(
            (_lib.size(L.surfel_binning_bytes(num_rendered), "surfel_binning_bytes") if num_rendered else 0,), **u8)
        _lib.check(L.surfel_forward_render(
            P, W, H, num_rendered, _ptr(background), _ptr(radii), _ptr(geom_buf), _ptr(bin_buf), _ptr(img_buf),
            _ptr(out_color), _ptr(out_others), st, int(bool(debug))), "surfel_forward_render")
    if KEEP_LAST:
        global LAST
        LAST = (num_rendered, geom_buf, bin_buf, img_buf)
    return num_rendered, out_color, out_others, radii, geom_buf, bin_buf, img_buf


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                 transMat_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color,
                                 dL_dout_others, sh, degree, campos, geomBuffer, R, binningBuffer, imageBuffer, debug):
    """-> (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dtransMat, dL_dsh, dL_dscales, dL_drotations)"""
    L = _lib.lib()
    background = _dev_f32(background, "background")
    means3D = _dev_f32(means3D, "means3D")
    colors = _dev_f32(colors, "colors")
    scales = _dev_f32(scales, "scales")
    rotations = _dev_f32(rotations, "rotations")
    transMat_precomp = _dev_f32(transMat_precomp, "transMat_precomp")
    viewmatrix = _dev_f32(viewmatrix, "viewmatrix")
    projmatrix = _dev_f32(projmatrix, "projmatrix")
    sh = _dev_f32(sh, "sh")
    campos = _dev_f32(campos, "campos")
    dL_dout_color = _dev_f32(dL_dout_color, "dL_dout_color")
    dL_dout_others = _dev_f32(dL_dout_others, "dL_dout_others")
    radii = radii.contiguous()

    P = int(means3D.size(0))
    H, W = int(dL_dout_color.size(1)), int(dL_dout_color.size(2))
    M = int(sh.size(1)) if sh.numel() != 0 else 0
    dev = means3D.device
    f32 = dict(dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        # torch.empty everywhere: the library writes every element (the reference zero-fills 304 B/Gaussian)
        dL_dmeans3D = torch.empty((P, 3), **f32)
        dL_dmeans2D = torch.empty((P, 3), **f32)
        dL_dcolors = torch.empty((P, NUM_CHANNELS), **f32)
        dL_dopacity = torch.empty((P, 1), **f32)
        dL_dtransMat = torch.empty((P, 9), **f32)
        dL_dsh = torch.empty((P, M, 3), **f32)
        dL_dscales = torch.empty((P, 2), **f32)
        dL_drotations = torch.empty((P, 4), **f32)
        if P:
            scratch = torch.empty((_lib.size(L.surfel_grad_scratch_bytes(P), "surfel_grad_scratch_bytes"),),
                                  dtype=torch.uint8, device=dev)
            _lib.check(L.surfel_backward(
                P, int(degree), M, int(R), _ptr(background), W, H, _ptr(means3D), _ptr(sh), _ptr(colors),
                _ptr(scales), float(scale_modifier), _ptr(rotations), _ptr(transMat_precomp), _ptr(viewmatrix),
                _ptr(projmatrix), _ptr(campos), float(tan_fovx), float(tan_fovy), _ptr(radii), _ptr(geomBuffer),
                _ptr(binningBuffer), _ptr(imageBuffer), _ptr(dL_dout_color), _ptr(dL_dout_others),
                _ptr(dL_dmeans2D), None, _ptr(dL_dopacity), _ptr(dL_dcolors), _ptr(dL_dmeans3D), _ptr(dL_dtransMat),
                _ptr(dL_dsh), _ptr(dL_dscales), _ptr(dL_drotations), _ptr(scratch), _stream(), int(bool(debug))),
                "surfel_backward")
            if KEEP_LAST:
                global LAST_BWD_AUX
                flag = C.c_int(-1)
                _lib.check(L.surfel_debug_aux_flag(P, _ptr(scratch), C.byref(flag), _stream()), "surfel_debug_aux_flag")
                LAST_BWD_AUX = int(flag.value)
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dtransMat, dL_dsh, dL_dscales, dL_drotations


def mark_visible(means3D, viewmatrix, projmatrix):
    L = _lib.lib()
    means3D = _dev_f32(means3D, "means3D")
    viewmatrix = _dev_f32(viewmatrix, "viewmatrix")
    projmatrix = _dev_f32(projmatrix, "projmatrix")
    P = int(means3D.size(0))
    present = torch.empty((P,), dtype=torch.bool, device=means3D.device)
    if P:
        with torch.cuda.device(means3D.device):
            _lib.check(L.surfel_mark_visible(P, _ptr(means3D), _ptr(viewmatrix), _ptr(projmatrix), _ptr(present),
                                             _stream()), "surfel_mark_visible")
    return present


def debug_binning(width, height, num_rendered, binningBuffer, imageBuffer):
    """(ranges [tiles,2] int64, point_list [R] int64) copied out of the private scratch layout (tests only)."""
    L = _lib.lib()
    tiles = ((width + 15) // 16) * ((height + 15) // 16)
    dev = imageBuffer.device
    ranges = torch.empty((tiles, 2), dtype=torch.int32, device=dev)
    plist = torch.empty((max(num_rendered, 1),), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.surfel_debug_copy_binning(width, height, num_rendered, _ptr(binningBuffer), _ptr(imageBuffer),
                                               _ptr(ranges), _ptr(plist), _stream()), "surfel_debug_copy_binning")
    return ranges.to(torch.int64) & 0xFFFFFFFF, (plist[:num_rendered].to(torch.int64) & 0xFFFFFFFF)


def debug_geometry(P, geomBuffer):
    """(tiles_touched [P], idx_sorted [P], offsets [P], records [P,24]) copied out of the geometry scratch (tests only)."""
    L = _lib.lib()
    dev = geomBuffer.device
    tiles = torch.empty((P,), dtype=torch.int32, device=dev)
    idx = torch.empty((P,), dtype=torch.int32, device=dev)
    offs = torch.empty((P,), dtype=torch.int32, device=dev)
    recs = torch.empty((P, 24), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.surfel_debug_copy_geometry(P, _ptr(geomBuffer), _ptr(tiles), _ptr(idx), _ptr(offs), _ptr(recs),
                                                _stream()), "surfel_debug_copy_geometry")
    m = 0xFFFFFFFF
    return tiles.to(torch.int64) & m, idx.to(torch.int64) & m, offs.to(torch.int64) & m, recs
