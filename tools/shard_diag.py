"""Per-piece timing of the sharded dedup step (developer diagnostic; run under torchrun)."""
import os, sys, torch, torch.distributed as dist
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from deal_yolo_daya_b200 import _lib, ops, sharding, synth_device
from deal_yolo_daya_b200.ops import KEEP_MODES, _ptr, _stream
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 10_000_000
_, uoff, udata = synth_device.make_urls(0, rank * n, n, dev)
keys = ops.hash_strings(uoff, udata)
x = sharding.DedupExchange(n, world, dev)
lib = x.lib; m = world * x.cap; s = _stream(dev)
def ev(): return torch.cuda.Event(enable_timing=True)
names = ["hash", "bucket", "a2a", "dedup", "pack", "a2a_back", "unpack"]
acc = {k: 0.0 for k in names}
for it in range(8):
    e = [ev() for _ in range(8)]
    e[0].record(); k2 = ops.hash_strings(uoff, udata)
    e[1].record(); _lib.check(lib.dyd_shard_bucket(_ptr(keys), None, rank * n, n, world, x.cap, _ptr(x.send), _ptr(x.cursors), _ptr(x.overflow), s), "b")
    e[2].record(); dist.all_to_all_single(x.recv, x.send)
    e[3].record(); _lib.check(lib.dyd_dedup_records(_ptr(x.recv), m, 0, _ptr(x.keep_r), _ptr(x.rep_r), _ptr(x.ws), x.ws.numel(), s), "d")
    e[4].record(); _lib.check(lib.dyd_shard_pack_reply(_ptr(x.recv), _ptr(x.keep_r), _ptr(x.rep_r), m, _ptr(x.reply), s), "p")
    e[5].record(); dist.all_to_all_single(x.back, x.reply)
    e[6].record(); _lib.check(lib.dyd_shard_unpack(_ptr(x.back), m, rank * n, n, _ptr(x.keep), _ptr(x.rep), s), "u")
    e[7].record(); torch.cuda.synchronize()
    if it >= 3:
        for i, k in enumerate(names): acc[k] += e[i].elapsed_time(e[i + 1]) / 5
if rank == 0:
    print({k: round(v, 3) for k, v in acc.items()}, "total", round(sum(acc.values()), 3), "ms; overflow", int(x.overflow.item()), "dups", int(n - x.keep.sum().item()), flush=True)
# whole-step loops without host syncs, with and without the fused kernel
t = synth_device.make_table(0, rank * n, n, dev)
buf = ops.FusedBuffers(t.n_img, t.n_poly, dev)
def loop(with_fused, with_xch, steps=10):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = ev(), ev(); a.record()
    for _ in range(steps):
        if with_fused: ops.bbox_iou_fused(t.img_off, t.poly_off, t.xy, 2, 0.7, out=buf)
        k = ops.hash_strings(uoff, udata)
        if with_xch: x.run(k, rank * n, "first", check_overflow=False)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps
for wf, wx in ((True, False), (False, True), (True, True), (True, True)):
    ms = loop(wf, wx)
    if rank == 0: print(f"fused={wf} exchange={wx}: {ms:.3f} ms/step", flush=True)
dist.destroy_process_group()
