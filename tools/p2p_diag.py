"""Per-piece timing of the peer-memory dedup exchange (developer diagnostic; run under torchrun)."""
import os, sys, torch, torch.distributed as dist
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from deal_yolo_daya_b200 import _lib, ops, sharding, synth_device
from deal_yolo_daya_b200.ops import _ptr, _stream
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 10_000_000
_, uoff, udata = synth_device.make_urls(0, rank * n, n, dev)
keys = ops.hash_strings(uoff, udata)
x = sharding.DedupExchange(n, world, dev)
assert x.transport == "p2p", getattr(x, "p2p_error", "")
lib = x.lib; m = world * x.cap; s = _stream(dev)
names = ["fill", "barrier_a", "bucket_p2p", "barrier_b", "dedup_records", "pack_p2p", "barrier_c", "unpack"]
acc = {k: 0.0 for k in names}
def ev(): return torch.cuda.Event(enable_timing=True)
for it in range(10):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e = [ev() for _ in range(9)]
    e[0].record(); x.recv.fill_(-1)
    e[1].record(); x.h_recv.barrier(channel=0)
    e[2].record(); _lib.check(lib.dyd_shard_bucket_p2p(_ptr(keys), None, rank * n, n, world, rank, x.cap, _ptr(x.peer_recv), _ptr(x.sent_row), _ptr(x.cursors), _ptr(x.overflow), s), "b")
    e[3].record(); x.h_recv.barrier(channel=1)
    e[4].record(); _lib.check(lib.dyd_dedup_records(_ptr(x.recv), m, 0, _ptr(x.keep_r), _ptr(x.rep_r), _ptr(x.ws), x.ws.numel(), s), "d")
    e[5].record(); _lib.check(lib.dyd_shard_pack_reply_p2p(_ptr(x.recv), _ptr(x.keep_r), _ptr(x.rep_r), m, x.cap, rank, _ptr(x.peer_back), s), "p")
    e[6].record(); x.h_back.barrier(channel=0)
    e[7].record(); _lib.check(lib.dyd_shard_unpack_p2p(_ptr(x.back), _ptr(x.sent_row), _ptr(x.cursors), world, x.cap, n, _ptr(x.keep), _ptr(x.rep), s), "u")
    e[8].record(); torch.cuda.synchronize()
    if it >= 4:
        for i, k in enumerate(names): acc[k] += e[i].elapsed_time(e[i + 1]) / 6
t = torch.tensor([acc[k] for k in names], device=dev)
tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
if rank == 0:
    print("world", world, "m", m, "cap", x.cap)
    for i, k in enumerate(names): print(f"{k:14s} min {tmin[i].item():.3f}  max {tmax[i].item():.3f} ms")
    print("sum of max", round(float(tmax.sum()), 3), "ms")
dist.destroy_process_group()
