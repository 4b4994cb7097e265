"""Host<->device copy bandwidth on the box (pinned memory): each direction alone and both at once -- the ceiling of
bench.py's `e2e` leg, which moves 562 MB in and 594 MB out per step."""
import json
import torch
dev = torch.device("cuda")
N = 512 * 1024 * 1024
h_in, h_out = torch.empty(N, dtype=torch.uint8).pin_memory(), torch.empty(N, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(N, dtype=torch.uint8, device=dev), torch.empty(N, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def run(h2d, d2h, reps=8):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return reps * N / (e0.elapsed_time(e1) * 1e-3) / 1e9

run(True, True, 2)
print(json.dumps({"h2d_alone_GBs": round(run(True, False), 1), "d2h_alone_GBs": round(run(False, True), 1),
                  "each_direction_when_concurrent_GBs": round(run(True, True), 1), "buffer_MiB": 512}))
