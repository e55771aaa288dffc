"""K5 anti-join and K4 dedup at growing sizes (developer tool)."""
import sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from deal_yolo_daya_b200 import _lib, ops, synth_device
_lib.load()
dev = torch.device("cuda", 0)
def t_ms(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for n in (4_000_000, 20_000_000, 50_000_000):
    _, off, data = synth_device.make_urls(0, 0, n, dev)
    mk = ops.hash_strings(off, data)
    _, roff, rdata = synth_device.make_urls(0, 0, n // 2, dev, n_main_for_ref=n)
    rk = ops.hash_strings(roff, rdata)
    del off, data, roff, rdata
    ms = t_ms(lambda: ops.antijoin(mk, None, rk, None))
    md = t_ms(lambda: ops.dedup(mk, None, "first"))
    print(f"n_main {n:>11,d} n_ref {n // 2:>11,d}: antijoin {ms:7.3f} ms = {(n + n // 2) / ms / 1e6:6.1f} G rows/s; dedup {md:7.3f} ms = {n / md / 1e6:6.1f} G rows/s", flush=True)
    del mk, rk
