"""B200-native differentiable 2D-Gaussian surfel rasterizer (drop-in for StreetUnveiler's
``diff_surfel_rasterization``).  See DESIGN.md / INTEGRATION.md."""
import sys

__all__ = ["install_dropin"]


def install_dropin() -> None:
    """Make ``import diff_surfel_rasterization`` resolve to this implementation, so the reference's
    gaussian_renderer/__init__.py:11 works unchanged."""
    from . import diff_surfel_rasterization as mod

    sys.modules["diff_surfel_rasterization"] = mod
    sys.modules["diff_surfel_rasterization._C"] = mod._C
