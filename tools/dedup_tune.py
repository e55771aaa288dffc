"""Time K4 with different numbers of table passes (developer tool)."""
import os, sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from deal_yolo_daya_b200 import _lib, ops, synth_device
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
_, off, data = synth_device.make_urls(0, 0, n, dev)
keys = ops.hash_strings(off, data)
ws = torch.empty(_lib.load().dyd_dedup_workspace_bytes(n), dtype=torch.uint8, device=dev)
ref = None
for lp in ("0", "1", "2", "3", "4", "5"):
    os.environ["DYD_DEDUP_PASSES_LOG2"] = lp
    for _ in range(3): k, r = ops.dedup(keys, None, "first", workspace=ws)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): k, r = ops.dedup(keys, None, "first", workspace=ws)
    b.record(); torch.cuda.synchronize()
    if ref is None: ref = (k.clone(), r.clone())
    print(f"passes 2^{lp}: {a.elapsed_time(b) / 10:.3f} ms  same={torch.equal(k, ref[0]) and torch.equal(r, ref[1])}", flush=True)
