"""Per-piece timing of the joint dedup + anti-join peer-memory exchange (sharding.UrlFilterExchange.run, p2p transport);
developer diagnostic, run under torchrun.  Prints min / max over ranks of every piece."""
import os, sys, torch, torch.distributed as dist
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from deal_yolo_daya_b200 import _lib, ops, sharding, synth_device
from deal_yolo_daya_b200.ops import _ptr, _stream
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 10_000_000; n_ref = n // 2
_, uoff, udata = synth_device.make_urls(0, rank * n, n, dev)
_, roff, rdata = synth_device.make_urls(0, rank * n_ref, n_ref, dev, n_main_for_ref=world * n)
x = sharding.UrlFilterExchange(n, n_ref, world, dev)
assert x.transport == "p2p", getattr(x, "p2p_error", "")
lib = x.lib; m = world * x.cap; m_ref = world * x.cap_ref; s = _stream(dev)
names = ["hash_main", "hash_ref", "scatter_ref", "scatter_main", "barrier_1", "url_filter_records", "pack_both",
         "barrier_2", "unpack_both"]
acc = {k: 0.0 for k in names}
def ev(): return torch.cuda.Event(enable_timing=True)
R = 12
for it in range(R):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e = [ev() for _ in range(len(names) + 1)]
    i = 0
    e[i].record(); i += 1; keys = ops.hash_strings(uoff, udata)
    e[i].record(); i += 1; rkeys = ops.hash_strings(roff, rdata)
    e[i].record(); i += 1; x._scatter_p2p(x.ref, rkeys, None, rank * n_ref, x.overflow[1:], s)
    e[i].record(); i += 1; _lib.check(lib.dyd_shard_bucket_p2p_defaults(_ptr(keys), None, rank * n, n, world, rank, x.main.cap, _ptr(x.main.peers), _ptr(x.main.sent_row), _ptr(x.main.cursors), _ptr(x.overflow[:1]), _ptr(x.keep_d), _ptr(x.rep_d), _ptr(x.keep), _ptr(x.rep), s), "s")
    e[i].record(); i += 1; x.main.h.barrier(channel=1)
    e[i].record(); i += 1; _lib.check(lib.dyd_url_filter_records(_ptr(x.recv_ref), m_ref, _ptr(x.recv), m, 0, _ptr(x.keep_dr), _ptr(x.rep_dr), _ptr(x.keep_r), _ptr(x.rep_r), _ptr(x.ws_d), x.ws_d.numel(), 1, world * n, s), "j")
    e[i].record(); i += 1; _lib.check(lib.dyd_shard_pack_reply2_p2p(_ptr(x.recv), _ptr(x.keep_dr), _ptr(x.rep_dr), _ptr(x.keep_r), _ptr(x.rep_r), m, x.cap, rank, _ptr(x.peer_back2), 1, 1, s), "p")
    e[i].record(); i += 1; x.h_back2.barrier(channel=0)
    e[i].record(); i += 1; _lib.check(lib.dyd_shard_unpack2_p2p(_ptr(x.back2), _ptr(x.main.sent_row), _ptr(x.cursors), world, x.cap, n, _ptr(x.keep_d), _ptr(x.rep_d), _ptr(x.keep), _ptr(x.rep), 1, s), "u")
    e[i].record(); torch.cuda.synchronize()
    if it >= 4:
        for j, k in enumerate(names): acc[k] += e[j].elapsed_time(e[j + 1]) / (R - 4)
t = torch.tensor([acc[k] for k in names], device=dev)
tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
if rank == 0:
    print("world", world, "m", m, "cap", x.cap, "m_ref", m_ref)
    for j, k in enumerate(names): print(f"{k:18s} min {tmin[j].item():.3f}  max {tmax[j].item():.3f} ms")
    print("sum of max", round(float(tmax.sum()), 3), "ms;  sum of rank0", round(float(t.sum()), 3), "ms")
dist.destroy_process_group()
