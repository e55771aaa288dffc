#!/usr/bin/env python
"""Benchmark of the hot path: ptList -> bbox + IoU filter + URL dedup, images/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input.  At N=1 the batch
is BASELINE.json configs[1]: 10 M images / ~80 M polygons (SURVEY §8d generator), fused
polygon->bbox + IoU flag (thr 0.7, min_boxes 2) plus hash + first-occurrence dedup of the
10 M `source` URLs (5 % duplicates).  At N>1 every rank holds its own 10 M-image shard
(weak scaling): bbox/IoU need no communication, dedup hash-partitions its keys with one
all-to-all each way (deal_yolo_daya_b200/sharding.py).

Prints ONE JSON line (rank 0).  `value` = whole-job images/s with inputs resident in HBM;
`e2e` = the same metric through the host-buffer C-ABI entry points (H2D and D2H inside the
timed region); `roofline` describes the fused kernel; `cpu_baseline` is the CPU port of the
reference timed on this box's host cores on a bounded sample.

`--impl reference` times the reference's CPU path.  The reference is pure Python and cannot
travel to the GPU box, so this arm runs oracle/pipeline_port.py -- the row-level port held
byte-for-byte to the real reference's outputs by tests/test_oracle_golden.py -- through CSV
files with every host core (kind "port").
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
emit = None

METRIC = "images/sec, ptList->bbox+IoU filter+dedup"
UNIT = "images/s"
IMAGES_PER_GPU = 10_000_000
SEED = 0
MIN_BOXES, THR = 2, 0.7


def kernels_per_step(n_rows: int, world: int) -> int:
    """tile_desc, fused_tma, iou_crowd, hash_strings + dedup (csrc/hash_dedup.cu): partition + resolve +
    the gated global-table fallback (fill, insert / lookup once per half of a table above 256 MB), which
    is launched every time and exits at once unless a partition overflowed; at N > 1 bucket / pack_reply
    / unpack in addition.  Checked against the ncu launch list in profiles/."""
    cap = 1024
    while cap < 2 * n_rows:
        cap *= 2
    passes = 2 if cap * 16 > (256 << 20) else 1
    if n_rows < (1 << 21):
        return 4 + 2 * passes + (3 if world > 1 else 0)
    return 4 + 2 + 1 + 2 * passes + (3 if world > 1 else 0)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------- CPU port legs
def _port_worker(args):
    """One process: the row-level port through CSV files on its own sample (single-threaded like the reference)."""
    seed, first, n_rows, reps = args
    import contextlib
    import io
    import pandas as pd
    from deal_yolo_daya_b200 import synth
    from oracle import pipeline_port as port
    t = synth.make_table(seed, first, n_rows)
    rows = synth.table_to_rows(t)
    best = float("inf")
    with tempfile.TemporaryDirectory() as td:
        merged = Path(td) / "merged.csv"
        pd.DataFrame(rows, columns=[port.COL_SRC, port.COL_ANN]).to_csv(merged, index=False, encoding="utf-8-sig")
        for _ in range(reps):
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                port.run_hot_path_files(td, merged, None, MIN_BOXES, THR)
            best = min(best, time.perf_counter() - t0)
    return n_rows, best


def port_throughput(rows_per_proc: int, procs: int, reps: int = 1, first: int = 0):
    """images/s of the CPU port with `procs` worker processes, each on its own row range."""
    jobs = [(SEED, first + i * rows_per_proc, rows_per_proc, reps) for i in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        res = [_port_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_port_worker, jobs)
    wall = time.perf_counter() - t0
    slowest = max(r[1] for r in res)
    return sum(r[0] for r in res) / slowest, slowest, wall


def csr_port_throughput(n_img: int):
    """images/s of the arithmetic-only C port (oracle/dyd_oracle.c, OpenMP) on CSR buffers."""
    import numpy as np
    from deal_yolo_daya_b200 import synth
    from oracle import oracle_c
    t = synth.make_table(SEED, 0, n_img)
    urls = [synth.url_of(i) for i in t.url_id]
    off, data = oracle_c.pack_strings(urls)
    best = float("inf")
    for _ in range(3):
        t0 = time.perf_counter()
        pts, valid, _ = oracle_c.bbox_fold(t.poly_off, t.xy, want_arg=False)
        oracle_c.iou_filter(t.img_off, pts, valid, MIN_BOXES, THR)
        keys = oracle_c.hash_strings_buf(off, data)
        oracle_c.dedup(keys, np.zeros(n_img, np.uint8), "first")
        best = min(best, time.perf_counter() - t0)
    return n_img / best, oracle_c.num_threads()


def dropin_files_throughput(n_rows: int, device: int):
    """images/s of the drop-in's own step functions on CSV files (dedup -> ptList->bbox -> IoU filter),
    the same chain and file contract the reference arm runs."""
    import contextlib
    import io
    import pandas as pd
    from deal_yolo_daya_b200 import processor as P, synth
    os.environ["DYD_DEVICE"] = str(device)
    t = synth.make_table(SEED, 0, n_rows)
    rows = synth.table_to_rows(t)
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        merged = td / "merged.csv"
        pd.DataFrame(rows, columns=[P.COL_SRC, P.COL_ANN]).to_csv(merged, index=False, encoding="utf-8-sig")
        best, steps = float("inf"), None
        for _ in range(2):
            with contextlib.redirect_stdout(io.StringIO()):
                t0 = time.perf_counter()
                P.deduplicate_csv_by_source(str(merged), str(td / "dedup.csv"))
                t1 = time.perf_counter()
                P.process_csv_replace_ptlist(str(td / "dedup.csv"), str(td / "rep.csv"), str(td / "exc.csv"))
                t2 = time.perf_counter()
                P.filter_by_box_count_and_iou(str(td / "rep.csv"), str(td / "hi.csv"), str(td / "other.csv"), MIN_BOXES, THR)
                t3 = time.perf_counter()
            if t3 - t0 < best:
                best, steps = t3 - t0, {"dedup_s": t1 - t0, "replace_ptlist_s": t2 - t1, "iou_filter_s": t3 - t2}
        # steps 5.5 / 6 on what the chain left over (informational; not part of `value`)
        try:
            from deal_yolo_daya_b200 import labels as L
            other = P._read_csv(str(td / "other.csv"), encoding="utf-8-sig")
            lm = {synth.label_name(i): f"grp{i % 20:02d}" for i in range(synth.N_LABELS)}
            l2c = {f"grp{g:02d}": f"cat{g // 5}" for g in range(20)}
            with contextlib.redirect_stdout(io.StringIO()):
                t0 = time.perf_counter()
                out, _, _, _ = L.remap_df(other, lm)
                t1 = time.perf_counter()
                res = L.split_df(out, l2c)
                t2 = time.perf_counter()
            steps["labels"] = {"rows": int(len(other)), "remap_s": t1 - t0, "split_s": t2 - t1,
                               "expanded_rows": int(res["summary"]["classified"]), "lanes": dict(L.LAST)}
        except Exception as e:  # noqa: BLE001
            steps["labels"] = {"error": str(e)[:200]}
    return n_rows / best, steps


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rows = 1500
    vals = []
    for _ in range(args.warmup):
        port_throughput(200, min(cores, 8))
    for _ in range(args.steps):
        v, slow, wall = port_throughput(rows, cores)
        vals.append((v, slow))
    vals.sort()
    v, slow = vals[len(vals) // 2]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": slow * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C2 rows as reference-format CSV: {rows} rows x {cores} processes per step, dedup -> ptList->bbox -> IoU filter "
                               f"(thr {THR}, min_boxes {MIN_BOXES}) through CSV files", "seed": SEED},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{rows} rows per process, {cores} processes, oracle/pipeline_port.run_hot_path_files"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=IMAGES_PER_GPU, help="images per GPU")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    # only the JSON line may reach stdout: library banners (e.g. "NCCL version ...") are sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    global emit

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from deal_yolo_daya_b200 import _lib, build, ops, sharding, synth_device

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    if rank == 0:
        build.build()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    lib = _lib.load()
    peak, peak_kind = peaks()

    n = args.images
    first = rank * n
    t = synth_device.make_table(SEED, first, n, dev)
    url_id, uoff, udata = synth_device.make_urls(SEED, first, n, dev)
    del url_id
    n_img, n_poly, n_vert = t.n_img, t.n_poly, t.n_vert
    buf = ops.FusedBuffers(n_img, n_poly, dev)
    dws = torch.empty(lib.dyd_dedup_workspace_bytes(int(n_img * (1.3 if world > 1 else 1.0))), dtype=torch.uint8, device=dev)
    fused_bytes = 16 * n_vert + 8 * (n_poly + 1) + 33 * n_poly + 8 * (n_img + 1) + 5 * n_img
    ev_f0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev_f1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    state = {}
    xch = sharding.DedupExchange(n_img, world, dev) if world > 1 else None

    def step(i=None):
        if i is not None:
            ev_f0[i].record()
        ops.bbox_iou_fused(t.img_off, t.poly_off, t.xy, MIN_BOXES, THR, out=buf)
        if i is not None:
            ev_f1[i].record()
        keys = ops.hash_strings(uoff, udata)
        if world == 1:
            state["keep"], state["rep"] = ops.dedup(keys, None, "first", workspace=dws)
        else:
            state["keep"], state["rep"] = xch.run(keys, first, "first", check_overflow=False)

    for _ in range(max(args.warmup, 1)):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start(); time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()                      # every rank enters the timed region together
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    fused_ms = sum(a.elapsed_time(b) for a, b in zip(ev_f0, ev_f1)) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms_step = ms_total / args.steps
    value = world * n_img / (ms_step * 1e-3)
    if xch is not None:
        assert int(xch.overflow.item()) == 0, "exchange bucket overflow: rerun with exact-size splits"
    n_high = int(buf.high.sum().item()); n_dup = int(n_img - state["keep"].sum().item())

    # ---------------- end to end through the host-buffer C ABI (H2D + D2H inside the timed region) ----------------
    e2e = None
    if not args.no_e2e:
        ne = n_img
        try:
            avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
        except Exception:  # noqa: BLE001
            avail = 0
        need = 16 * n_vert + 8 * n_poly + 37 * n_poly + 60 * n_img
        frac = 1.0
        if avail and need * world > 0.6 * avail:
            frac = max(0.05, 0.6 * avail / (need * world))
            ne = int(n_img * frac)

        def pinned(src, count=None):
            src = src if count is None else src[:count]
            h = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            h.copy_(src)
            return h.numpy()

        q_e = int(t.img_off[ne].item()); v_e = int(t.poly_off[q_e].item())
        h_img = pinned(t.img_off, ne + 1); h_poly = pinned(t.poly_off, q_e + 1); h_xy = pinned(t.xy, 2 * v_e)
        ub = int(uoff[ne].item())
        h_uoff = pinned(uoff, ne + 1); h_udata = pinned(udata, ub)
        out = {"pts": torch.empty(4 * q_e, dtype=torch.float64, pin_memory=True).numpy(),
               "valid": torch.empty(q_e, dtype=torch.uint8, pin_memory=True).numpy(),
               "high": torch.empty(ne, dtype=torch.uint8, pin_memory=True).numpy(),
               "count": torch.empty(ne, dtype=torch.int32, pin_memory=True).numpy()}
        h2d = h_img.nbytes + h_poly.nbytes + h_xy.nbytes + h_uoff.nbytes + h_udata.nbytes
        d2h = out["pts"].nbytes + out["valid"].nbytes + out["high"].nbytes + out["count"].nbytes + ne * 9

        def e2e_step():
            ops.bbox_iou_host(h_img, h_poly, h_xy, MIN_BOXES, THR, want_pts=True, out=out, device=local)
            return ops.dedup_host(h_uoff, h_udata, None, "first", device=local)

        e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            k_e, _ = e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e2e = {"value": world * ne * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": args.e2e_steps, "images_per_step_per_gpu": ne,
               "note": "dyd_bbox_iou_host + dyd_dedup_host on pinned host CSR/Arrow buffers; chunked H2D/kernel/D2H overlap inside the library"
                       + ("" if frac == 1.0 else f"; host memory allowed only {frac:.2f} of the shard") +
                       ("" if world == 1 else "; per-rank dedup (no cross-rank exchange on the host path)")}
        assert int(out["high"].sum()) == int(buf.high[:ne].sum().item()), "host-path result differs from the device-resident path"
        del h_xy, h_poly, h_img, out

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    traffic = None
    tp = ROOT / "profiles" / "roofline_traffic.json"
    if tp.exists():
        tj = json.loads(tp.read_text())
        if tj.get("images") == n_img:
            traffic = tj.get("dram_bytes_per_launch")
    achieved = fused_bytes / (fused_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C2: {n_img} images / {n_poly} polygons / {n_vert} vertices per GPU, fused ptList->bbox + IoU flag "
                               f"(thr {THR}, min_boxes {MIN_BOXES}) + URL hash + first-occurrence dedup (5% dupes)",
                   "seed": SEED, "l2": f"inputs ({16 * n_vert / 1e9:.1f} GB of vertices per step) are far larger than the 126 MB L2; no flush needed",
                   "parallelism": (f"{world} rank(s), rows partitioned by image; dedup keys hash-partitioned to owner ranks, "
                                   f"exchange transport: {xch.transport}") if world > 1 else "1 GPU",
                   "results": {"high_iou_images": n_high, "duplicate_rows_rank0": n_dup}},
        "clocks": clocks,
        "gpu_launches": kernels_per_step(n_img, world) * args.steps,
        "roofline": {"bound": "hbm", "kernel": "fused_tma_kernel (+ tile_desc pre-pass and crowd worklist kernel, timed together)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_kind, "algorithmic_bytes_per_launch": fused_bytes,
                     "bytes_per_image": fused_bytes / n_img, "ms_per_launch": fused_ms},
    }
    if e2e:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        rows = 1200
        v1, slow1, _ = port_throughput(rows, 1)
        vall, slow, wall = port_throughput(rows, cores)
        line["cpu_baseline"] = {"value": vall, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{rows} synthetic C2 rows per process as reference-format CSV, {cores} processes "
                                          f"(oracle/pipeline_port.run_hot_path_files: dedup -> ptList->bbox -> IoU filter through CSV files)",
                                "one_core_value": v1}
        try:
            vd, dsteps = dropin_files_throughput(20_000, local)
            line["dropin_files"] = {"value": vd, "unit": UNIT, "rows": 20_000, "seconds": dsteps,
                                    "note": "this repo's processor.py step functions on CSV files, one process (native CSV reader/writer, native JSON "
                                            "ingest/egress, CUDA kernels); value = rows / (dedup + ptList->bbox + IoU filter seconds); "
                                            "seconds.labels = label remap + split of the rows that remain (not in value); "
                                            "compare with cpu_baseline.one_core_value"}
        except Exception as e:  # noqa: BLE001
            line["dropin_files"] = {"error": str(e)[:200]}
        try:
            vc, thr_c = csr_port_throughput(400_000)
            line["cpu_baseline_csr"] = {"value": vc, "unit": UNIT, "cores": thr_c, "kind": "port",
                                        "sample": "400000 images, arithmetic-only C port on CSR buffers (oracle/dyd_oracle.c, OpenMP): "
                                                  "no JSON/CSV work, the kernel-for-kernel comparison"}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline_csr"] = {"error": str(e)[:200]}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
