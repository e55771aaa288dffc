"""Peer-memory vs NCCL transport of the sharded dedup: same answers, and time per step
(developer tool; run under torchrun on >= 2 GPUs)."""
import os, sys, torch, torch.distributed as dist
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from deal_yolo_daya_b200 import ops, sharding, synth_device

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
_, uoff, udata = synth_device.make_urls(0, rank * n, n, dev)
keys = ops.hash_strings(uoff, udata)
res = {}
for transport in ("nccl", "p2p"):
    os.environ["DYD_EXCHANGE"] = transport
    x = sharding.DedupExchange(n, world, dev)
    if rank == 0:
        print(f"requested {transport}: using {x.transport}" + (f" ({getattr(x, 'p2p_error', '')})" if x.transport != transport else ""), flush=True)
    for keep in ("first", "last", False):
        k, r = x.run(keys, rank * n, keep)
        res[(transport, keep)] = (k.clone(), r.clone())
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): x.run(keys, rank * n, "first", check_overflow=False)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a.record()
    for _ in range(10): x.run(keys, rank * n, "first", check_overflow=False)
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / 10], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"{x.transport}: {t.item():.3f} ms per exchange step ({n} rows per rank)", flush=True)
    del x
same = all(torch.equal(res[("nccl", k)][0], res[("p2p", k)][0]) and torch.equal(res[("nccl", k)][1], res[("p2p", k)][1]) for k in ("first", "last", False))
# ground truth on rank 0 from all ranks' keys (exact-size path)
kk, rr = sharding.dedup_global(keys, None, rank * n, "first")
same_exact = torch.equal(kk, res[("p2p", "first")][0]) and torch.equal(rr, res[("p2p", "first")][1])
flags = torch.tensor([int(same), int(same_exact)], device=dev); dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0: print("p2p == nccl:", bool(flags[0].item()), " p2p == exact-size path:", bool(flags[1].item()), flush=True)
dist.destroy_process_group()
