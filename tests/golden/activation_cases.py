"""Seeded inputs of the parameter-activation golden cases (see make_golden_activation.py)."""
import torch

ACTIVATION_CASES = {"p1000_sh3": (1000, 15, 31), "p36_sh1": (36, 3, 32), "p5_sh0": (5, 0, 33)}   # name -> (P, R, seed)
RAW = ["scaling_raw", "rotation_raw", "opacity_raw", "features_dc", "features_rest"]
OUT = ["scaling", "rotation", "opacity", "features"]


def build_activation_case(name):
    P, R, seed = ACTIVATION_CASES[name]
    g = torch.Generator().manual_seed(seed)
    raw = {"scaling_raw": torch.randn(P, 2, generator=g) - 3.0, "rotation_raw": torch.randn(P, 4, generator=g) * 2.0,
           "opacity_raw": torch.randn(P, 1, generator=g) * 3.0, "features_dc": torch.randn(P, 1, 3, generator=g),
           "features_rest": torch.randn(P, R, 3, generator=g) * 0.1}
    raw["rotation_raw"][0] = torch.tensor([1e-3, 0.0, 0.0, 0.0])        # tiny but non-zero quaternion
    raw["opacity_raw"][1] = 30.0                                         # saturated sigmoid
    up = {"scaling": torch.randn(P, 2, generator=g), "rotation": torch.randn(P, 4, generator=g),
          "opacity": torch.randn(P, 1, generator=g), "features": torch.randn(P, 1 + R, 3, generator=g)}
    return dict(P=P, R=R, raw=raw, upstream=up)
