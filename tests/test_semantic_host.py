"""CPU checks of the host logic of the shared-binning semantic path (no GPU compute)."""
import torch


def test_one_hot_colors_match_reference_statements():
    from streetunveiler_b200.semantic_passes import one_hot_colors, CONCERNED_CLASSES
    g = torch.Generator().manual_seed(0)
    tag = torch.randint(0, 6, (500, 1), generator=g, dtype=torch.int32)
    n = len(CONCERNED_CLASSES)
    for i in range(0, n, 3):
        # gaussian_renderer/__init__.py:420-430
        semantic_3 = torch.zeros(500, 3)
        semantic_3[(tag == i).reshape(-1), 0] = 1.0
        if (i + 1) < n:
            semantic_3[(tag == (i + 1)).reshape(-1), 1] = 1.0
        if (i + 2) < n:
            semantic_3[(tag == (i + 2)).reshape(-1), 2] = 1.0
        assert torch.equal(one_hot_colors(tag, i, n), semantic_3)
    # five classes: the last pass has only two valid channels
    assert one_hot_colors(torch.tensor([[5], [4]], dtype=torch.int32), 3, 5)[:, 2].sum() == 0


def test_color_passes_validate_arguments_before_touching_the_device():
    import pytest
    from streetunveiler_b200.diff_surfel_rasterization import GaussianRasterizationSettings
    from streetunveiler_b200.diff_surfel_rasterization.color_passes import rasterize_color_passes
    z = torch.zeros(4, 3)
    s = GaussianRasterizationSettings(8, 8, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0, torch.zeros(3), False, False)
    with pytest.raises(Exception, match="scale/rotation pair"):
        rasterize_color_passes(s, z, z, torch.zeros(4, 1), [z])
    with pytest.raises(Exception, match="at least one"):
        rasterize_color_passes(s, z, z, torch.zeros(4, 1), [], scales=torch.zeros(4, 2), rotations=torch.zeros(4, 4))
    with pytest.raises(RuntimeError):   # CPU tensors: no fallback
        rasterize_color_passes(s, z, z, torch.zeros(4, 1), [z], scales=torch.zeros(4, 2), rotations=torch.zeros(4, 4))
