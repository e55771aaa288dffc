#!/usr/bin/env python
"""Does the fused call's time depend on the threshold or only on the order of the runs? (developer tool)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from deal_yolo_daya_b200 import _lib, ops, synth_device
from tools.kbench import time_ms

_lib.load()
dev = torch.device("cuda", 0)
t = synth_device.make_table(0, 0, 10_000_000, dev)
buf = ops.FusedBuffers(t.n_img, t.n_poly, dev)
for thr in (0.98, 0.7, 0.98, 0.7, 0.5, 0.98, 1.5):
    ms, best = time_ms(lambda: ops.bbox_iou_fused(t.img_off, t.poly_off, t.xy, 2, thr, out=buf), 8)
    print(f"thr {thr}: median {ms:.3f} ms best {best:.3f} ms high={int(buf.high.sum())}", flush=True)
