"""Where does the host-buffer path spend its time?  (developer diagnostic)"""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from deal_yolo_daya_b200 import ops, synth_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
dev = torch.device("cuda", 0)
t = synth_device.make_table(0, 0, n, dev)
_, uoff, udata = synth_device.make_urls(0, 0, n, dev)

def pinned(src):
    h = torch.empty(src.shape, dtype=src.dtype, pin_memory=True); h.copy_(src); return h
hp = {k: pinned(getattr(t, k)) for k in ("img_off", "poly_off", "xy")}
print("pinned?", {k: v.is_pinned() for k, v in hp.items()})
h = {k: v.numpy() for k, v in hp.items()}
out = {"pts": torch.empty(4 * t.n_poly, dtype=torch.float64, pin_memory=True).numpy(),
       "valid": torch.empty(t.n_poly, dtype=torch.uint8, pin_memory=True).numpy(),
       "high": torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy(),
       "count": torch.empty(n, dtype=torch.int32, pin_memory=True).numpy()}
gb = (h["xy"].nbytes + h["poly_off"].nbytes + h["img_off"].nbytes) / 1e9
d = torch.empty_like(t.xy)
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(hp["xy"], non_blocking=True); torch.cuda.synchronize()
print(f"torch H2D xy: {h['xy'].nbytes / 1e9 / (time.perf_counter() - t0):.1f} GB/s")
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(hp["xy"], non_blocking=True); torch.cuda.synchronize()
    print(f"torch H2D xy again: {h['xy'].nbytes / 1e9 / (time.perf_counter() - t0):.1f} GB/s")
chunks = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else [16384, 65536]
for chunk in chunks:
    for want_pts in (True, False):
        ts = []
        for rep in range(4):
            t0 = time.perf_counter()
            ops.bbox_iou_host(h["img_off"], h["poly_off"], h["xy"], 2, 0.7, want_pts=want_pts, out=out, chunk_images=chunk)
            ts.append(time.perf_counter() - t0)
        print(f"chunk {chunk:8d} pts={want_pts}: " + " ".join(f"{x * 1e3:7.1f}" for x in ts) + f" ms  best {gb / min(ts):6.1f} GB/s  {n / min(ts) / 1e6:6.1f} M images/s", flush=True)
ho, hd = pinned(uoff).numpy(), pinned(udata).numpy()
ops.dedup_host(ho, hd, None, "first")
for rep in range(3):
    t0 = time.perf_counter(); ops.dedup_host(ho, hd, None, "first"); dt = time.perf_counter() - t0
    print(f"dedup_host: {dt * 1e3:.1f} ms ({n / dt / 1e6:.1f} M rows/s)")
