import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, harness as hz
from cases import build_case
c = build_case(sys.argv[1] if len(sys.argv) > 1 else "box_sh0")
o = hz.run_ours(c["scene"], c["cam"], bg=c["bg"], grads=c["grads"], **c["kw"])
torch.cuda.synchronize()
print("ok R", o["num_rendered"])
