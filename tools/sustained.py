#!/usr/bin/env python
"""Burst vs sustained: per-launch time of the fused call and of a plain device copy over a few
hundred back-to-back launches, with nvidia-smi clocks / power sampled alongside (developer tool)."""
import subprocess, sys, threading, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from deal_yolo_daya_b200 import _lib, ops, synth_device

_lib.load()
dev = torch.device("cuda", 0)
samples, stop = [], False

def sampler():
    while not stop:
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active,temperature.gpu",
                                  "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True, timeout=5).stdout.strip()
            samples.append((time.perf_counter(), out))
        except Exception:  # noqa: BLE001
            pass
        time.sleep(0.02)

def run(name, fn, n, nbytes):
    global samples
    samples = []
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    evs[0].record()
    for i in range(n):
        fn(); evs[i + 1].record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]
    def gbs(x): return nbytes / (x * 1e-3) / 1e9
    for a, b in ((0, 10), (10, 20), (20, 40), (40, 80), (80, 160), (160, n)):
        if a < n:
            seg = sorted(ms[a:min(b, n)])
            print(f"{name}: launches {a:3d}-{min(b, n):3d}  median {seg[len(seg) // 2]:.3f} ms = {gbs(seg[len(seg) // 2]):7.1f} GB/s")
    mine = [s for t, s in samples if t0 <= t <= t1]
    print(f"{name}: nvidia-smi (sm MHz, mem MHz, W, reasons, C) first/middle/last: {mine[:1]} {mine[len(mine) // 2:len(mine) // 2 + 1]} {mine[-1:]}")

th = threading.Thread(target=sampler, daemon=True); th.start()
t = synth_device.make_table(0, 0, 10_000_000, dev)
buf = ops.FusedBuffers(t.n_img, t.n_poly, dev)
by = 16 * t.n_vert + 8 * (t.n_poly + 1) + 32 * t.n_poly + t.n_poly + 8 * (t.n_img + 1) + 5 * t.n_img
for _ in range(3):
    ops.bbox_iou_fused(t.img_off, t.poly_off, t.xy, 2, 0.7, out=buf)
torch.cuda.synchronize(); time.sleep(2.0)
run("fused", lambda: ops.bbox_iou_fused(t.img_off, t.poly_off, t.xy, 2, 0.7, out=buf), 300, by)
del buf, t
a = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev); b = torch.empty_like(a)
for _ in range(3):
    b.copy_(a)
torch.cuda.synchronize(); time.sleep(2.0)
run("copy ", lambda: b.copy_(a), 600, 2 * a.numel() * 2)
stop = True
