"""Generate the golden vectors under tests/golden/*.npz from the UNMODIFIED reference extension.

Run on a B200 box (needs a GPU and oracle/_ref built by oracle/build_ref.py):

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   # then copy *.npz to tests/golden/

Inputs are regenerated from seeds (tests/golden/cases.py); each file stores the reference's
outputs (color, allmap, radii, num_rendered, gradients) plus two repeat runs' worth of gradient
noise (float atomics make the reference's backward order-dependent).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import harness as hz  # noqa: E402
from cases import CASES, build_case  # noqa: E402


def main(out_dir, only=None):
    os.makedirs(out_dir, exist_ok=True)
    for name in (only or CASES):
        c = build_case(name)
        runs = [hz.run_reference(c["scene"], c["cam"], bg=c["bg"], grads=c["grads"], **c["kw"]) for _ in range(3)]
        ref = runs[0]
        out = {k: v for k, v in ref.items() if not k.startswith("_") and k != "num_rendered"}
        out["num_rendered"] = np.int64(ref["num_rendered"])
        for k in ref:
            if k.startswith("g_"):
                out["noise_" + k] = np.float32(max(hz.rel_err(r[k], ref[k]) for r in runs[1:]))
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **out)
        print(name, "R", ref["num_rendered"], "visible", int((ref["radii"] > 0).sum()),
              {k: float(v) for k, v in out.items() if k.startswith("noise_")})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "_new"), sys.argv[2:] or None)
