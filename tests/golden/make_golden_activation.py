"""Golden vectors for the fused parameter activations, generated on CPU in this container by the reference's OWN
``GaussianModel`` properties (scene/gaussian_model.py:101-127: get_scaling, get_rotation, get_opacity, get_features)
under autograd; only modules that cannot be imported without a GPU toolchain are stubbed (none of them is on this path).

    python tests/golden/make_golden_activation.py        # needs /root/reference, no GPU

Outputs: tests/golden/activation_*.npz (inputs are regenerated from seeds by tests/golden/activation_cases.py).
"""
import os
import sys
import types

import numpy as np
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from activation_cases import ACTIVATION_CASES, OUT, build_activation_case  # noqa: E402

REF = "/root/reference"


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {})


def reference_model():
    for name in ["plyfile", "simple_knn", "simple_knn._C", "tinycudann", "sh_encoder", "sh_encoder._shencoder",
                 "scene.dataset_readers", "diff_surfel_rasterization"]:
        sys.modules[name] = _Stub(name)
    sys.path.insert(0, REF)
    from scene.gaussian_model import GaussianModel
    return GaussianModel


def run_reference(GaussianModel, case):
    m = GaussianModel(3)
    raw = {k: nn.Parameter(v.clone()) for k, v in case["raw"].items()}
    m._scaling, m._rotation, m._opacity = raw["scaling_raw"], raw["rotation_raw"], raw["opacity_raw"]
    m._features_dc, m._features_rest = raw["features_dc"], raw["features_rest"]
    outs = {"scaling": m.get_scaling, "rotation": m.get_rotation, "opacity": m.get_opacity, "features": m.get_features}
    sum((outs[k] * case["upstream"][k]).sum() for k in OUT).backward()
    res = {k: v.detach().numpy() for k, v in outs.items()}
    res.update({"g_" + k: v.grad.numpy() for k, v in raw.items()})
    return res


if __name__ == "__main__":
    GM = reference_model()
    for name in ACTIVATION_CASES:
        out = run_reference(GM, build_activation_case(name))
        np.savez_compressed(os.path.join(HERE, f"activation_{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if not k.startswith("g_")})
