"""URL-chain micro-benchmark (developer tool): K0 hash, K4 dedup, K5 anti-join on device-generated URL columns,
timed alone with CUDA events.  python tools/url_bench.py [rows] [reps]"""
import sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from deal_yolo_daya_b200 import _lib, ops, synth_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
lib = _lib.load()
_, off, data = synth_device.make_urls(0, 0, n, dev)
_, roff, rdata = synth_device.make_urls(0, 0, n // 2, dev, n_main_for_ref=n)
keys = ops.hash_strings(off, data); rkeys = ops.hash_strings(roff, rdata)
ws = torch.empty(max(lib.dyd_dedup_workspace_bytes(n), lib.dyd_antijoin_fast_workspace_bytes(n, n // 2)), dtype=torch.uint8, device=dev)


def t(fn, label):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    print(f"{label:28s} median {ts[len(ts) // 2]:.3f} ms  best {ts[0]:.3f} ms", flush=True)


t(lambda: ops.hash_strings(off, data), f"hash {n} urls")
for keep in ("first", "last", False):
    t(lambda: ops.dedup(keys, None, keep, workspace=ws), f"dedup keep={keep}")
t(lambda: ops.antijoin(keys, None, rkeys, None, workspace=ws), "antijoin (ref = n/2)")
