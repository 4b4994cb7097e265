"""Times NCCL all-gather / reduce-scatter / all-reduce at the message sizes of the sharded path (CUDA events)."""
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    dev = torch.device("cuda")
    for rows in (4096, 65536, 409600, 2_000_000):
        for name, width in (("all_gather", 24), ("reduce_scatter", 20)):
            x = torch.randn(rows * (1 if name == "all_gather" else world), width, device=dev)
            out = torch.empty(rows * (world if name == "all_gather" else 1), width, device=dev)
            fn = (lambda: dist.all_gather_into_tensor(out, x)) if name == "all_gather" else (lambda: dist.reduce_scatter_tensor(out, x))
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            if rank == 0:
                mb = rows * width * 4 / 1e6
                print(f"{name:15s} rows/rank={rows:8d} ({mb:7.1f} MB/rank): {ms:7.3f} ms  -> {mb * (world - 1) / ms / 1e3:6.1f} GB/s in per rank")
    img = torch.randn(10, 1280, 1920, device=dev)
    for _ in range(5):
        dist.all_reduce(img)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(img)
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"all_reduce 10 planes (98 MB): {e0.elapsed_time(e1) / 20:.3f} ms")
    dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
