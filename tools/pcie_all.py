"""Aggregate host<->device copy bandwidth with every rank copying at the same time (pinned memory), the ceiling the N-GPU
`e2e` leg of bench.py runs into.  Run under torchrun; prints per-rank and total GB/s for H2D, D2H and both directions."""
import os, sys, time, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device=dev); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(kind):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(8):
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d.copy_(h, non_blocking=True)
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = (16 if kind == "both" else 8) * n / dt / 1e9
    t = torch.tensor([gbs], dtype=torch.float64, device=dev)
    if world > 1:
        allg = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allg, t)
    else:
        allg = [t]
    return [round(float(x.item()), 1) for x in allg]


for kind in ("h2d", "d2h", "both"):
    run(kind)
    r = run(kind)
    if rank == 0:
        print(f"{kind:5s} per rank {r}  total {sum(r):.1f} GB/s", flush=True)
if world > 1:
    dist.destroy_process_group()
