"""Kernel micro-benchmarks (developer tool; bench.py is the judged entry point).

    python tools/kbench.py [--images N] [--reps R] [--which fused,k1,k2,dedup,antijoin,hash,crowd]

Times each kernel family alone with CUDA events on the current stream over a device-resident
synthetic table (inputs >> L2 at the default size), and prints achieved algorithmic GB/s
against MEASURED_PEAKS.json.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from deal_yolo_daya_b200 import _lib, ops, synth_device  # noqa: E402


def peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text())["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def time_ms(fn, reps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=4_000_000)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--which", default="fused,k1,k2,hash,dedup,antijoin,crowd,labels")
    args = ap.parse_args()
    which = set(args.which.split(","))
    _lib.load()
    dev = torch.device("cuda", 0)
    peak, kind = peak_gbs()
    t0 = time.time()
    t = synth_device.make_table(0, 0, args.images, dev)
    torch.cuda.synchronize()
    print(f"table: {t.n_img} images, {t.n_poly} polygons, {t.n_vert} vertices ({t.xy.numel() * 8 / 1e9:.2f} GB xy) "
          f"generated in {time.time() - t0:.1f}s; peak {peak} GB/s ({kind})", flush=True)
    n_img, n_poly, n_vert = t.n_img, t.n_poly, t.n_vert
    rows = []

    def report(name, ms, best, nbytes, units, unit_name):
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append((name, ms, best, gbs, gbs / peak, units / (ms * 1e-3), unit_name))
        print(f"{name:28s} median {ms:8.3f} ms  best {best:8.3f} ms  {gbs:8.1f} GB/s  {100 * gbs / peak:5.1f}% of {kind} peak  "
              f"{units / (ms * 1e-3) / 1e6:10.1f} M {unit_name}/s", flush=True)

    import os
    if "fused" in which:
        buf = ops.FusedBuffers(n_img, n_poly, dev)
        by = 16 * n_vert + 8 * (n_poly + 1) + 32 * n_poly + n_poly + 8 * (n_img + 1) + 5 * n_img
        for fused, group in (("tma", "4"), ("direct", "4"), ("direct", "8")):
            os.environ["DYD_FUSED"] = fused; os.environ["DYD_GROUP"] = group
            for thr in (0.7, 0.98):
                ms, best = time_ms(lambda: ops.bbox_iou_fused(t.img_off, t.poly_off, t.xy, 2, thr, out=buf), args.reps)
                report(f"fused {fused}/G{group} thr={thr}", ms, best, by, n_img, "images")
        os.environ["DYD_FUSED"] = "tma"; os.environ["DYD_GROUP"] = "4"
        bufa = ops.FusedBuffers(n_img, n_poly, dev, want_arg=True)
        ms, best = time_ms(lambda: ops.bbox_iou_fused(t.img_off, t.poly_off, t.xy, 2, 0.7, want_arg=True, out=bufa), args.reps)
        report("fused tma +arg", ms, best, by + 16 * n_poly, n_img, "images")
        del bufa
    if "k1" in which:
        by = 16 * n_vert + 8 * (n_poly + 1) + 33 * n_poly
        for group in ("4", "8", "2"):
            os.environ["DYD_GROUP"] = group
            ms, best = time_ms(lambda: ops.bbox_minmax(t.poly_off, t.xy), args.reps)
            report(f"K1 bbox G{group} (allocs incl.)", ms, best, by, n_poly, "polygons")
        os.environ["DYD_GROUP"] = "4"
    if "k2" in which:
        pts, valid, _ = ops.bbox_minmax(t.poly_off, t.xy)
        by = 33 * n_poly + 8 * (n_img + 1) + 5 * n_img
        ms, best = time_ms(lambda: ops.iou_filter(t.img_off, pts, valid, 2, 0.7), args.reps)
        report("K2 iou sparse", ms, best, by, n_img, "images")
        del pts, valid
    if which & {"hash", "dedup", "antijoin"}:
        url_id, off, data = synth_device.make_urls(0, 0, n_img, dev)
        ms, best = time_ms(lambda: ops.hash_strings(off, data), args.reps)
        report("K0 hash urls", ms, best, data.numel() + 8 * (n_img + 1) + 8 * n_img, n_img, "rows")
        keys = ops.hash_strings(off, data)
        ws = torch.empty(_lib.load().dyd_dedup_workspace_bytes(n_img), dtype=torch.uint8, device=dev)
        ms, best = time_ms(lambda: ops.dedup(keys, None, "first", workspace=ws), args.reps)
        report("K4 dedup first", ms, best, 17 * n_img, n_img, "rows")
        if "antijoin" in which:
            rid, roff, rdata = synth_device.make_urls(0, 0, n_img // 2, dev, n_main_for_ref=n_img)
            rkeys = ops.hash_strings(roff, rdata)
            ms, best = time_ms(lambda: ops.antijoin(keys, None, rkeys, None, workspace=ws), args.reps)
            report("K5 antijoin (ref = n/2)", ms, best, 8 * (n_img // 2) + 17 * n_img, n_img, "rows")
            ws2 = torch.empty(_lib.load().dyd_url_filter_workspace_bytes(n_img, n_img // 2), dtype=torch.uint8, device=dev)
            ms, best = time_ms(lambda: ops.url_filter(keys, None, rkeys, None, "first", workspace=ws2), args.reps)
            report("K4+K5 joint url filter", ms, best, 8 * (n_img // 2) + 8 * n_img + 18 * n_img, n_img, "rows")
    if "crowd" in which:
        nc = max(1000, args.images // 400)
        io, pts = synth_device.make_crowd(0, 0, nc, device=dev)
        nb = pts.numel() // 4
        pairs = float(((io[1:] - io[:-1]).double() * ((io[1:] - io[:-1]).double() - 1) / 2).sum().item())
        for thr, tag in ((0.7, "natural"), (2.0, "worst case")):
            ms, best = time_ms(lambda: ops.iou_filter(io, pts, None, 2, thr), max(3, args.reps // 2))
            report(f"K2 crowd {tag}", ms, best, 32 * nb + 13 * nc, nc, "images")
            if thr == 2.0:
                print(f"    {pairs / (ms * 1e-3) / 1e9:.1f} G pairs/s over {nc} images, {nb} boxes", flush=True)
    if "labels" in which:
        import numpy as np
        nv, ncat = 100, 4                         # BASELINE config C5: 4 categories
        rng = np.random.RandomState(0)
        lut_new = torch.from_numpy(rng.randint(0, nv, nv).astype(np.int32)).to(dev)
        lut_ntok = torch.from_numpy(np.ones(nv, np.int32)).to(dev)
        lut_nrep = torch.from_numpy((rng.rand(nv) < 0.8).astype(np.int32)).to(dev)
        cat = torch.from_numpy(rng.randint(-1, ncat, nv).astype(np.int32)).to(dev)
        ms, best = time_ms(lambda: ops.label_lut(t.img_off, t.label_id, lut_new, lut_ntok, lut_nrep), args.reps)
        report("K3 label LUT", ms, best, 8 * n_poly + 9 * n_img, n_poly, "objects")
        ms, best = time_ms(lambda: ops.label_hist(t.label_id, nv), args.reps)
        report("K3 label histogram", ms, best, 4 * n_poly, n_poly, "objects")
        ei, eb, ec, co = ops.split_expand(t.img_off, t.label_id, cat, ncat)
        ms, best = time_ms(lambda: ops.split_expand(t.img_off, t.label_id, cat, ncat), max(3, args.reps // 2))
        report("K6 split expand (2 passes)", ms, best, 8 * n_poly + 8 * n_img + 20 * ei.numel(), n_poly, "objects")
        cat20 = torch.from_numpy(rng.randint(-1, 20, nv).astype(np.int32)).to(dev)
        ei, eb, ec, co = ops.split_expand(t.img_off, t.label_id, cat20, 20)
        ms, best = time_ms(lambda: ops.split_expand(t.img_off, t.label_id, cat20, 20), max(3, args.reps // 2))
        report("K6 split expand, 20 categories", ms, best, 8 * n_poly + 8 * n_img + 20 * ei.numel(), n_poly, "objects")
        pts, valid, _ = ops.bbox_minmax(t.poly_off, t.xy)
        wh = torch.tensor([1920.0, 1080.0], dtype=torch.float64, device=dev).repeat(n_img)
        ms, best = time_ms(lambda: ops.yolo_normalise(t.img_off, pts, valid, wh), args.reps)
        report("YOLO normalise", ms, best, 66 * n_poly + 24 * n_img, n_poly, "objects")
    print(json.dumps([{"name": r[0], "ms": r[1], "best_ms": r[2], "gbs": r[3], "frac": r[4]} for r in rows]))


if __name__ == "__main__":
    main()
