"""Probe: host->device copy bandwidth from pinned memory, one stream vs two (GPU box only)."""
import time, torch
n = 610 * 1024 * 1024 // 4
hs = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(4)]
ds = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(4)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(streams, reps=5):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        for i in range(4):
            with torch.cuda.stream(streams[i % len(streams)]):
                ds[i].copy_(hs[i], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    return reps * 4 * n * 4 / dt / 1e9
for name, st in (("one stream", [s1]), ("two streams", [s1, s2])):
    run(st, 1)
    print(name, round(run(st), 2), "GB/s")
big = torch.empty(4 * n, dtype=torch.float32).pin_memory(); dbig = torch.empty(4 * n, dtype=torch.float32, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5): dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize(); print("single 2.4 GB copies", round(5 * 4 * n * 4 / (time.perf_counter() - t) / 1e9, 2), "GB/s")
