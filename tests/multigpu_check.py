"""Run under torchrun (one rank per GPU): sharded rasterization must equal the single-GPU operator.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py [P] [all|color_alpha]

The second argument picks the upstream gradients: "all" exercises the full backward-blend specialisation, "color_alpha"
the colour+alpha one that the device-side flag selects when the depth/normal/distortion gradients are zero.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as hz  # noqa: E402
from streetunveiler_b200 import synthetic as syn  # noqa: E402
from streetunveiler_b200.sharded import ShardedRasterizer  # noqa: E402


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    cam = syn.make_camera(960, 640, 1027.5, 1027.5)
    scene = syn.street_scene(P, 4, 3)
    mode = sys.argv[2] if len(sys.argv) > 2 else "all"
    grads = syn.upstream_grads(cam.width, cam.height, mode, seed=11)
    bg = torch.tensor([0.1, 0.3, 0.2])
    # uneven shards on purpose (padding path): rank r owns [cuts[r], cuts[r+1])
    cuts = [0] + [int(P * (r + 1) / world) - (7 * (r + 1) if r + 1 < world else 0) for r in range(world)]
    lo, hi = cuts[rank], cuts[rank + 1]
    mod = hz.ours_module()
    st = hz._settings(mod, cam, bg, 3, 1.0, dev)
    p = {k: v[lo:hi].to(dev).clone().requires_grad_(True) for k, v in scene.items() if isinstance(v, torch.Tensor)}
    m2 = torch.zeros_like(p["means3D"], requires_grad=True)
    rast = ShardedRasterizer()
    color, radii, allmap = rast(p["means3D"], m2, p["opacities"], p["shs"], p["scales"], p["rotations"], st)
    torch.autograd.backward([color, allmap], [grads[0].to(dev), grads[1].to(dev)])
    torch.cuda.synchronize()
    # single-GPU reference on every rank (full scene), compare own slice
    ref = hz.run_ours(scene, cam, bg=bg, grads=grads, device=str(dev))
    ok = True
    msgs = []

    def chk(name, a, b, tol):
        nonlocal ok
        e = hz.rel_err(a, b)
        msgs.append(f"{name}={e:.2e}")
        if not e <= tol:
            ok = False

    chk("color", color.detach().cpu().numpy(), ref["color"], 0.0)       # bit-identical forward
    chk("allmap", allmap.detach().cpu().numpy(), ref["allmap"], 0.0)
    if not np.array_equal(radii.cpu().numpy(), ref["radii"][lo:hi]):
        ok = False
        msgs.append("radii MISMATCH")
    chk("g_means3D", p["means3D"].grad.cpu().numpy(), ref["g_means3D"][lo:hi], 1e-4)
    chk("g_means2D", m2.grad.cpu().numpy(), ref["g_means2D"][lo:hi], 1e-4)
    chk("g_shs", p["shs"].grad.cpu().numpy(), ref["g_shs"][lo:hi], 1e-4)
    chk("g_opacities", p["opacities"].grad.cpu().numpy(), ref["g_opacities"][lo:hi], 1e-4)
    chk("g_scales", p["scales"].grad.cpu().numpy(), ref["g_scales"][lo:hi], 1e-4)
    chk("g_rotations", p["rotations"].grad.cpu().numpy(), ref["g_rotations"][lo:hi], 1e-4)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print(f"[rank {rank}/{world}] shard [{lo},{hi}) R_window={color.grad_fn.num_rendered} image exchange: "
          f"{rast.backend.image_exchange}, row exchange: {rast.backend.row_exchange} " + " ".join(msgs), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        print("MULTIGPU_CHECK FAILED", flush=True)
        sys.exit(1)
    if rank == 0:
        print("MULTIGPU_CHECK OK", flush=True)


if __name__ == "__main__":
    main()
