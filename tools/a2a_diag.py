"""NCCL all-to-all / p2p bandwidth between the ranks of this box (developer diagnostic)."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 22_000_000
a = torch.empty(n, dtype=torch.int64, device="cuda"); b = torch.empty_like(a)
for it in range(3):
    dist.all_to_all_single(b, a)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(10):
    dist.all_to_all_single(b, a)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
if rank == 0:
    print(f"all_to_all_single {n * 8 / 1e6:.0f} MB per rank: {ms:.3f} ms  -> {n * 8 * (world - 1) / world / ms / 1e6:.1f} GB/s sent per rank", flush=True)
    print("p2p access 0->1:", torch.cuda.can_device_access_peer(0, 1) if world > 1 else None)
dist.destroy_process_group()
