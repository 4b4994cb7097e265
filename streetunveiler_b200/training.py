"""One training iteration of the reference's loop on the fused rows (train.py:109-199 without data loading, the sky
model's own optimiser and logging):

    activations (row f5) -> rasterizer -> render() epilogue (f1) -> loss block (f3) -> backward
    -> densification statistics + Adam (f4)

``fused_training_step`` is what a ``train.py`` that has switched to this package calls per iteration; every piece is
also usable on its own.  The model object is the reference's ``GaussianModel`` (or anything
with its raw parameter attributes ``_xyz, _features_dc, _features_rest, _scaling, _rotation, _opacity`` and the
statistics tensors ``max_radii2D, xyz_gradient_accum, denom``).
"""
from __future__ import annotations

from .fused_adam import densification_stats
from .loss_block import training_loss
from .parameter_activation import ActivatedGaussians
from .surface_epilogue import render


def fused_training_step(gaussians, viewpoint_cam, pipe, background, gt_image, sky_image, lambda_dssim, lambda_normal=0.0,
                        lambda_dist=0.0, optimizer=None, update_densification_stats=True):
    """-> ``(loss, loss_dict, render_pkg)``.

    train.py:109   ``render_pkg = render(viewpoint_cam, gaussians, pipe, background)``
    train.py:113-136  composite with the sky image, L1 + D-SSIM, normal-consistency and distortion terms
    train.py:143   ``loss.backward()``
    train.py:168-169  ``max_radii2D`` / ``add_densification_stats``   (if ``update_densification_stats``)
    train.py:197-198  ``optimizer.step(); optimizer.zero_grad(set_to_none=True)``   (if ``optimizer`` is given)
    The shrink term (train.py:138-141), densify / prune / opacity reset and the sky model stay with the caller.
    """
    render_pkg = render(viewpoint_cam, ActivatedGaussians(gaussians), pipe, background)
    loss, loss_dict = training_loss(render_pkg, sky_image, gt_image, lambda_dssim, lambda_normal, lambda_dist)
    loss.backward()
    if update_densification_stats:
        densification_stats(render_pkg["radii"], render_pkg["viewspace_points"].grad, gaussians.max_radii2D,
                            gaussians.xyz_gradient_accum, gaussians.denom)
    if optimizer is not None:
        optimizer.step()
        optimizer.zero_grad(set_to_none=True)
    return loss, loss_dict, render_pkg
