#!/usr/bin/env python
"""Benchmark of the hot path: ptList -> bbox + IoU filter + URL dedup + reference filter, images/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input.  At N=1 the batch is
BASELINE.json configs[1] (C2): 10 M images / ~80 M polygons (SURVEY §8d generator), fused polygon->bbox +
IoU flag (thr 0.7, min_boxes 2), plus the URL chain of configs[2] (C3) on the same rows: hash of the 10 M
`source` URLs (5 % duplicates) and of a 5 M-row reference set (10 % of it inside the main id range),
first-occurrence dedup and the reference-set anti-join.  At N>1 every rank holds its own 10 M-image /
5 M-reference-row shard (weak scaling; N=8 is C3's 100 M-vs-50 M shape at 80 M / 40 M): bbox/IoU need no
communication; dedup and the anti-join hash-partition their keys to owner ranks over NVLink peer memory
(deal_yolo_daya_b200/sharding.py).  The URL chain depends on the `source` column only, so it runs on a second
stream next to the fused kernel, which leaves it a few SMs (dyd_bbox_iou_fused_ex); a step ends when both
streams are done.

Prints ONE JSON line (rank 0).  `value` = whole-job images/s with inputs resident in HBM; `e2e` = the same
work through the host-buffer entry points (H2D and D2H inside the timed region, cross-rank exchange
included); `roofline` describes the fused kernel; `results` carries the global duplicate / filtered counts
and `exchange_verified` -- the sharded answers compared bit for bit, after the timed region, with a
sort-based ground truth on the generator's integer url ids (deal_yolo_daya_b200/verify.py); `cpu_baseline`
is the reference's own CPU path timed on this box's host cores on a bounded sample; `c4` / `c5` are the
dense-crowd and remap+split legs (BASELINE.json configs[3], configs[4]).

`--impl reference` times the reference's CPU path: the UNMODIFIED reference step functions staged in
oracle/_ref by oracle/make_ref.py (kind "reference"), or -- when no staged copy travelled -- the row-level
port oracle/pipeline_port.py, held byte-for-byte to the reference's outputs by tests/test_oracle_golden.py
(kind "port"); through CSV files, all host cores.
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
URL_SMS = 0                      # SMs the fused kernel leaves to a second (URL) stream; 0 = one stream (overlap is zero-sum here)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_config(world: int, n_img: int):
    """The workload both arms are quoted on (the reference arm times a bounded sample of it)."""
    return {"workload": f"C2 + C3 URL chain: {n_img} images (~{8 * n_img} polygons, 4-32 vertices each) and {n_img // 2} reference rows per GPU; "
                        f"dedup by source (5% dupes) -> reference filter (10% of the reference set inside the main ids) -> ptList->bbox -> "
                        f"IoU filter (thr {THR}, min_boxes {MIN_BOXES})",
            "seed": SEED, "images_per_gpu": n_img, "reference_rows_per_gpu": n_img // 2,
            "l2": "inputs (23 GB of vertices per step at 10 M images) are far larger than the 126 MB L2; no flush needed",
            "parallelism": f"{world} rank(s), rows partitioned by image; dedup / anti-join keys hash-partitioned to owner ranks" if world > 1 else "1 GPU"}


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


# --------------------------------------------------------------------------------------- CPU legs
def _write_sample(td: Path, seed, first, n_rows):
    """Reference-format CSVs of one bounded sample: merged.csv (rows [first, first+n)) and ref.csv (n/2 rows, 10 % of
    them sources of the sample)."""
    import numpy as np
    import pandas as pd
    from deal_yolo_daya_b200 import synth
    from oracle import pipeline_port as port
    t = synth.make_table(seed, first, n_rows)
    rows = synth.table_to_rows(t)
    merged, ref = td / "merged.csv", td / "ref.csv"
    pd.DataFrame(rows, columns=[port.COL_SRC, port.COL_ANN]).to_csv(merged, index=False, encoding="utf-8-sig")
    n_ref = n_rows // 2
    inside = [int(i) for i in t.url_id[::20]][: max(1, n_ref // 10)]
    ref_ids = inside + [10 ** 12 + first + k for k in range(n_ref - len(inside))]
    ref_ids = [ref_ids[i] for i in np.random.RandomState(seed).permutation(len(ref_ids))]
    pd.DataFrame({port.COL_SRC: [synth.url_of(i) for i in ref_ids]}).to_csv(ref, index=False, encoding="utf-8-sig")
    return merged, ref


def _cpu_worker(args):
    """One process: the reference's step functions through CSV files on its own sample (single-threaded like the reference)."""
    seed, first, n_rows, reps, kind = args
    import contextlib
    import io
    if kind == "reference":
        from oracle import ref_loader as impl
    else:
        from oracle import pipeline_port as impl
    best = float("inf")
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        merged, ref = _write_sample(td, seed, first, n_rows)
        for _ in range(reps):
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                impl.run_hot_path_files(td, merged, ref, MIN_BOXES, THR)
            best = min(best, time.perf_counter() - t0)
    return n_rows, best


def cpu_kind():
    from oracle import ref_loader
    return "reference" if ref_loader.available() else "port"


def cpu_throughput(rows_per_proc: int, procs: int, kind: str, reps: int = 1, first: int = 0):
    """images/s of the CPU path with `procs` worker processes, each on its own row range."""
    jobs = [(SEED, first + i * rows_per_proc, rows_per_proc, reps, kind) for i in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    slowest = max(r[1] for r in res)
    return sum(r[0] for r in res) / slowest, slowest, wall


def cpu_sample_text(kind, rows, procs):
    what = ("the unmodified reference step functions (oracle/_ref via oracle/ref_loader.run_hot_path_files)" if kind == "reference"
            else "oracle/pipeline_port.run_hot_path_files")
    return (f"{rows} synthetic C2 rows (+ {rows // 2} reference rows) per process as reference-format CSV, {procs} process(es): {what}: "
            f"dedup -> reference filter -> ptList->bbox -> IoU filter through CSV files")


def csr_port_throughput(n_img: int):
    """images/s of the arithmetic-only C port (oracle/dyd_oracle.c, OpenMP) on CSR buffers."""
    import numpy as np
    from deal_yolo_daya_b200 import synth
    from oracle import oracle_c
    t = synth.make_table(SEED, 0, n_img)
    urls = [synth.url_of(i) for i in t.url_id]
    off, data = oracle_c.pack_strings(urls)
    rurls = [synth.url_of(i) for i in synth.ref_ids_of(SEED, np.arange(n_img // 2), n_img)]
    roff, rdata = oracle_c.pack_strings(rurls)
    best = float("inf")
    for _ in range(3):
        t0 = time.perf_counter()
        pts, valid, _ = oracle_c.bbox_fold(t.poly_off, t.xy, want_arg=False)
        oracle_c.iou_filter(t.img_off, pts, valid, MIN_BOXES, THR)
        keys = oracle_c.hash_strings_buf(off, data)
        rkeys = oracle_c.hash_strings_buf(roff, rdata)
        oracle_c.dedup(keys, np.zeros(n_img, np.uint8), "first")
        oracle_c.antijoin(keys, np.zeros(n_img, np.uint8), rkeys, np.zeros(len(rkeys), np.uint8))
        best = min(best, time.perf_counter() - t0)
    return n_img / best, oracle_c.num_threads()


def dropin_files_throughput(n_rows: int, device: int):
    """images/s of the drop-in's own step functions on CSV files (dedup -> reference filter -> ptList->bbox -> IoU
    filter), the same chain and file contract the reference arm runs."""
    import contextlib
    import io
    from deal_yolo_daya_b200 import processor as P, synth, tablecache
    os.environ["DYD_DEVICE"] = str(device)
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        merged, ref = _write_sample(td, SEED, 0, n_rows)
        best, steps = float("inf"), None
        for _ in range(3):
            tablecache.clear()                      # every pass starts cold: the first step really reads and parses merged.csv
            P.PHASES.clear()
            with contextlib.redirect_stdout(io.StringIO()):
                t0 = time.perf_counter()
                P.deduplicate_csv_by_source(str(merged), str(td / "dedup.csv"))
                t1 = time.perf_counter()
                P.remove_duplicates_between_csv(str(td / "dedup.csv"), str(ref), str(td / "filtered.csv"))
                t2 = time.perf_counter()
                P.process_csv_replace_ptlist(str(td / "filtered.csv"), str(td / "rep.csv"), str(td / "exc.csv"))
                t3 = time.perf_counter()
                P.filter_by_box_count_and_iou(str(td / "rep.csv"), str(td / "hi.csv"), str(td / "other.csv"), MIN_BOXES, THR)
                t4 = time.perf_counter()
            if t4 - t0 < best:
                best = t4 - t0
                steps = {"dedup_s": t1 - t0, "ref_filter_s": t2 - t1, "replace_ptlist_s": t3 - t2, "iou_filter_s": t4 - t3,
                         "phases": {k: round(v, 4) for k, v in P.PHASES.items()}, "table_cache": dict(tablecache.STATS)}
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
    kind = cpu_kind()
    rows = 1000 if kind == "reference" else 1500
    vals = []
    for _ in range(args.warmup):
        cpu_throughput(200, min(cores, 8), kind)
    for _ in range(args.steps):
        v, slow, wall = cpu_throughput(rows, cores, kind)
        vals.append((v, slow))
    vals.sort()
    v, slow = vals[len(vals) // 2]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": slow * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": make_config(world, args.images),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": cpu_sample_text(kind, rows, cores) + " per step"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------- NUMA
def bind_to_gpu_numa_node(local: int):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned allocation: cudaHostAlloc
    places pages on the allocating thread's node, and a pinned buffer on the far socket halves H2D bandwidth once
    several ranks copy at the same time.  No-op when the box exposes one node."""
    info = {"bound": False}
    try:
        pci = None
        out = subprocess.run(["nvidia-smi", f"--id={local}", "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True, timeout=20).stdout.strip()
        if out:
            pci = out.lower()
            if pci.startswith("00000000:"):
                pci = pci[4:]
        nodes = sorted(p.name for p in Path("/sys/devices/system/node").glob("node[0-9]*"))
        info["nodes"] = len(nodes)
        node = -1
        if pci and Path(f"/sys/bus/pci/devices/{pci}/numa_node").exists():
            node = int(Path(f"/sys/bus/pci/devices/{pci}/numa_node").read_text().strip())
        info["gpu_numa_node"] = node
        if len(nodes) > 1 and node >= 0:
            cpus = Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip()
            ids = set()
            for part in cpus.split(","):
                a, _, b = part.partition("-")
                ids.update(range(int(a), int(b or a) + 1))
            ids &= os.sched_getaffinity(0)
            if ids:
                os.sched_setaffinity(0, ids)
                info["bound"] = True
                info["cpus"] = cpus
    except Exception as e:  # noqa: BLE001
        info["error"] = str(e)[:120]
    return info


# --------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=IMAGES_PER_GPU, help="images per GPU")
    ap.add_argument("--url-sms", type=int, default=URL_SMS, help="SMs left to the URL stream by the fused kernel (0: one stream, no overlap)")
    ap.add_argument("--coresident", action="store_true", help="URL chain on a second stream next to ALL 148 fused CTAs (its CTAs share the SMs "
                    "with the persistent ones; needs a fused kernel built with room in shared memory, DESIGN.md §5.1)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = --steps")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the C4 / C5 legs")
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
    from deal_yolo_daya_b200 import _lib, build, ops, sharding, synth_device, verify

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    if rank == 0:
        build.build()
    numa = bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    lib = _lib.load()
    peak, peak_kind = peaks()

    n = args.images
    n_ref = n // 2
    first, ref_first = rank * n, rank * n_ref
    t = synth_device.make_table(SEED, first, n, dev)
    url_id, uoff, udata = synth_device.make_urls(SEED, first, n, dev)
    ref_id, roff, rdata = synth_device.make_urls(SEED, ref_first, n_ref, dev, n_main_for_ref=world * n)
    n_img, n_poly, n_vert = t.n_img, t.n_poly, t.n_vert
    buf = ops.FusedBuffers(n_img, n_poly, dev)
    fused_bytes = 16 * n_vert + 8 * (n_poly + 1) + 33 * n_poly + 8 * (n_img + 1) + 5 * n_img
    url_bytes = int(udata.numel() + rdata.numel()) + 16 * (n_img + n_ref) + 18 * n_img + 8 * n_ref + 17 * n_img   # K0 (both) + K4 + K5
    overlap = args.url_sms > 0 or args.coresident
    max_ctas = 148 - args.url_sms if args.url_sms > 0 else 0
    # The polygon stream has the higher priority: when the pre-pass ends its persistent CTAs are placed first and the URL
    # stream's kernels, released by the same event, fill the SMs that are left (and all of them once the fused kernel is done).
    s_poly = torch.cuda.Stream(dev, priority=-1) if overlap else torch.cuda.current_stream(dev)
    s_url = torch.cuda.Stream(dev, priority=0) if overlap else s_poly
    torch.cuda.synchronize()
    torch.cuda.set_stream(s_poly)
    ev_pre = torch.cuda.Event()
    ev_pre.record()                                 # creates the underlying cudaEvent
    K = args.steps
    mk = lambda: [torch.cuda.Event(enable_timing=True) for _ in range(K)]   # noqa: E731
    ev_f0, ev_f1, ev_u0, ev_u1, ev_a0 = mk(), mk(), mk(), mk(), mk()
    state = {}
    if world > 1:
        xj = sharding.UrlFilterExchange(n_img, n_ref, world, dev)       # dedup + anti-join in one exchange
    else:
        dws = torch.empty(lib.dyd_url_filter_workspace_bytes(n_img, n_ref), dtype=torch.uint8, device=dev)

    def url_chain(i):
        if i is not None:
            ev_u0[i].record()
        keys = ops.hash_strings(uoff, udata)
        rkeys = ops.hash_strings(roff, rdata)
        if world == 1:
            if i is not None:
                ev_a0[i].record()
            # both questions about the same keys in one call: common key partitions, one shared-memory table each (dyd_url_filter)
            state["keep"], state["rep"], state["keep_ref"], state["ref_row"] = ops.url_filter(keys, None, rkeys, None, "first", workspace=dws)
        else:
            if i is not None:
                ev_a0[i].record()
            state["keep"], state["rep"], state["keep_ref"], state["ref_row"] = xj.run(keys, first, rkeys, ref_first, "first", check_overflow=False)
        if i is not None:
            ev_u1[i].record()
        state["keys"] = keys

    def step(i=None):
        if i is not None:
            ev_f0[i].record()
        ops.bbox_iou_fused(t.img_off, t.poly_off, t.xy, MIN_BOXES, THR, out=buf, max_ctas=max_ctas, prepass_event=ev_pre if overlap else None)
        if i is not None:
            ev_f1[i].record()
        if overlap:
            with torch.cuda.stream(s_url):
                s_url.wait_event(ev_pre)            # the URL chain of a step starts when the fused kernel's CTAs are being placed ...
                url_chain(i)
            s_poly.wait_stream(s_url)               # ... and the step ends when both streams are done
        else:
            url_chain(i)

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
    launches0 = int(lib.dyd_launch_count())
    e0.record()
    for i in range(K):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    launches = int(lib.dyd_launch_count()) - launches0
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    fused_ms = sum(a.elapsed_time(b) for a, b in zip(ev_f0, ev_f1)) / K
    url_ms = sum(a.elapsed_time(b) for a, b in zip(ev_u0, ev_u1)) / K
    anti_ms = sum(a.elapsed_time(b) for a, b in zip(ev_a0, ev_u1)) / K
    clocks = sampler.stop() if rank == 0 else None

    def allmax(v):
        if world == 1:
            return float(v)
        tt = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    ct = ops.fused_cta_times().astype(np.float64)
    ct = ct[ct[:, 0] > 0]
    cta_times = {"ctas": int(len(ct)), "start_spread_ms": float((ct[:, 0].max() - ct[:, 0].min()) * 1e-6),
                 "end_spread_ms": float((ct[:, 1].max() - ct[:, 1].min()) * 1e-6),
                 "kernel_ms": float((ct[:, 1].max() - ct[:, 0].min()) * 1e-6)} if len(ct) else None
    ms_total = allmax(ms_total)
    ms_step = ms_total / K
    value = world * n_img / (ms_step * 1e-3)
    url_ms_max, anti_ms_max = allmax(url_ms), allmax(anti_ms)

    # ---------------- results + verification against the url-id ground truth (after the timed region) ----------------
    overflowed = 0
    if world > 1:
        fl = xj.overflow.max().reshape(1).clone()
        dist.all_reduce(fl, op=dist.ReduceOp.MAX)
        overflowed = int(fl.item())
    assert overflowed == 0, "exchange bucket overflow: rerun with exact-size splits"
    n_high = int(buf.high.sum().item())
    ek, er = verify.expected_dedup_first(url_id, first)
    ak, ar = verify.expected_antijoin(url_id, ref_id, ref_first)
    ok = [torch.equal(ek, state["keep"]), torch.equal(er, state["rep"]), torch.equal(ak, state["keep_ref"]), torch.equal(ar, state["ref_row"])]
    okt = torch.tensor([int(all(ok))], device=dev)
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    verified = bool(okt.item())
    dup_global = verify.global_counts(state["keep"])
    filt_global = verify.global_counts(state["keep_ref"])
    dup_expected = verify.global_counts(ek); filt_expected = verify.global_counts(ak)
    dropped_either = verify.global_counts(state["keep"] & state["keep_ref"])
    high_t = torch.tensor([n_high], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(high_t)
    del ek, er, ak, ar
    results = {"high_iou_images_global": int(high_t.item()), "duplicate_rows_global": dup_global, "duplicate_rows_expected": dup_expected,
               "duplicate_rows_rank0": int(n_img - state["keep"].sum().item()),
               "reference_filtered_rows_global": filt_global, "reference_filtered_rows_expected": filt_expected,
               "rows_dropped_by_either_global": dropped_either, "exchange_verified": verified,
               "verified_how": "keep / rep of the sharded dedup and keep / ref_row of the sharded anti-join of the LAST timed step compared bit for bit, on "
                               "every rank, with a torch sort-based ground truth over the all-gathered integer url ids (deal_yolo_daya_b200/verify.py)",
               "exchange_transport": (xj.transport + ", dedup + anti-join in one exchange (sharding.UrlFilterExchange)" if world > 1 else "none (1 GPU)")}
    assert verified, f"sharded dedup / anti-join differs from the url-id ground truth: {ok}"

    # ---------------- end to end through the host-buffer entry points (H2D + D2H inside the timed region) ----------------
    e2e = None
    if not args.no_e2e:
        ne = n_img
        try:
            avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
        except Exception:  # noqa: BLE001
            avail = 0
        need = 16 * n_vert + 8 * n_poly + 37 * n_poly + 120 * n_img
        frac = 1.0
        if avail and need * world > 0.6 * avail:
            frac = max(0.05, 0.6 * avail / (need * world))
            ne = int(n_img * frac)
        if world > 1:                               # every rank must hold the same number of rows (global row ids)
            tt = torch.tensor([ne], dtype=torch.int64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MIN)
            ne = int(tt.item())
        ne_ref = ne // 2

        def pinned(src, count=None):
            src = src if count is None else src[:count]
            h = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            h.copy_(src)
            return h.numpy()

        q_e = int(t.img_off[ne].item()); v_e = int(t.poly_off[q_e].item())
        h_img = pinned(t.img_off, ne + 1); h_poly = pinned(t.poly_off, q_e + 1); h_xy = pinned(t.xy, 2 * v_e)
        # the URL columns of this leg's table: rows [rank*ne, rank*ne + ne) and reference rows [rank*ne_ref, ...)
        e_uid, e_uoff, e_udata = synth_device.make_urls(SEED, rank * ne, ne, dev)
        e_rid, e_roff, e_rdata = synth_device.make_urls(SEED, rank * ne_ref, ne_ref, dev, n_main_for_ref=world * ne)
        h_uoff = pinned(e_uoff); h_udata = pinned(e_udata); h_roff = pinned(e_roff); h_rdata = pinned(e_rdata)
        out = {"pts": torch.empty(4 * q_e, dtype=torch.float64, pin_memory=True).numpy(),
               "valid": torch.empty(q_e, dtype=torch.uint8, pin_memory=True).numpy(),
               "high": torch.empty(ne, dtype=torch.uint8, pin_memory=True).numpy(),
               "count": torch.empty(ne, dtype=torch.int32, pin_memory=True).numpy()}
        url = sharding.ShardedUrlFilter(ne, ne_ref, h_udata.nbytes, h_rdata.nbytes, world, dev)
        h2d_poly = h_img.nbytes + h_poly.nbytes + h_xy.nbytes
        h2d = h2d_poly + h_uoff.nbytes + h_udata.nbytes + h_roff.nbytes + h_rdata.nbytes
        d2h = out["pts"].nbytes + out["valid"].nbytes + out["high"].nbytes + out["count"].nbytes + 18 * ne
        t_poly = [0.0]

        def e2e_step():
            a = time.perf_counter()
            r = ops.bbox_iou_host(h_img, h_poly, h_xy, MIN_BOXES, THR, want_pts=True, out=out, device=local, tile_modes=True)
            t_poly[0] += time.perf_counter() - a
            u = url.run(h_uoff, h_udata, h_roff, h_rdata, rank * ne, rank * ne_ref, "first")
            return r, u

        e2e_step()
        esteps = args.e2e_steps or K
        t_poly[0] = 0.0
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            r_e, u_e = e2e_step()
        torch.cuda.synchronize()
        dt_own = time.perf_counter() - t0
        dt = allmax(dt_own)
        gbs = torch.tensor([h2d_poly * esteps / t_poly[0] / 1e9, (h2d + d2h) * esteps / dt_own / 1e9], dtype=torch.float64, device=dev)
        if world > 1:
            allg = [torch.empty_like(gbs) for _ in range(world)]
            dist.all_gather(allg, gbs)
        else:
            allg = [gbs]
        # the host path must give the device-resident answers: polygons directly, the URL columns against the same ground truth
        assert int(out["high"].sum()) == int(buf.high[:ne].sum().item()), "host-path result differs from the device-resident path"
        ek, er = verify.expected_dedup_first(e_uid, rank * ne)
        ak, ar = verify.expected_antijoin(e_uid, e_rid, rank * ne_ref)
        e_ok = (torch.equal(ek.cpu(), u_e[0]) and torch.equal(er.cpu(), u_e[1]) and torch.equal(ak.cpu(), u_e[2]) and torch.equal(ar.cpu(), u_e[3]))
        okt = torch.tensor([int(e_ok)], device=dev)
        if world > 1:
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        assert bool(okt.item()), "host-path dedup / anti-join differs from the url-id ground truth"
        e2e = {"value": world * ne * esteps / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": esteps, "images_per_step_per_gpu": ne, "verified": True, "tile_modes": r_e.get("tile_modes"),
               "per_rank_polygon_h2d_gbs": [round(float(g[0]), 2) for g in allg],
               "per_rank_total_pcie_gbs": [round(float(g[1]), 2) for g in allg],
               "numa": numa,
               "note": "dyd_bbox_iou_host on pinned host CSR buffers (chunked H2D / kernel / D2H overlap inside the library), then "
                       "sharding.ShardedUrlFilter on pinned host Arrow buffers (H2D, hash, dedup + anti-join"
                       + (" with the cross-rank exchange" if world > 1 else "") + ", D2H)"
                       + ("" if frac == 1.0 else f"; host memory allowed only {frac:.2f} of the shard")}
        del h_xy, h_poly, h_img, out, url, e_uid, e_rid

    # ---------------- C4 / C5 legs (every rank takes part; rank 0 reports) ----------------
    legs = {}
    if not args.no_legs:
        import bench_legs
        try:
            legs["c4"] = bench_legs.c4_leg(dev, rank, world, peak)
        except Exception as e:  # noqa: BLE001
            legs["c4"] = {"error": repr(e)[:300]}
        try:
            legs["c5"] = bench_legs.c5_leg(dev, rank, world, t, peak)
        except Exception as e:  # noqa: BLE001
            legs["c5"] = {"error": repr(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    traffic, traffic_source = None, None
    tp = ROOT / "profiles" / "roofline_traffic.json"
    if tp.exists():
        tj = json.loads(tp.read_text())
        if tj.get("images") == n_img:
            traffic = tj.get("dram_bytes_per_launch")
            traffic_source = "not measured in this run (a run under ncu is never a bench value): " + str(tj.get("source"))
    achieved = fused_bytes / (fused_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(world, n_img),
        "results": results,
        "clocks": clocks,
        "gpu_launches": launches,
        "gpu_launches_how": "dyd_launch_count() before / after the timed region: every kernel launch of libdyd.so increments it",
        "streams": {"overlap": overlap, "fused_ctas": max_ctas or 148, "url_stream_sms": args.url_sms,
                    "fused_ms": fused_ms, "url_chain_ms": url_ms_max, ("url_filter_ms" if world == 1 else "joint_exchange_ms"): anti_ms_max, "fused_cta_times_last_launch": cta_times,
                    "note": "per-step CUDA-event times on each stream (max over ranks for the URL chain); a step ends when both streams are done"},
        "roofline": {"bound": "hbm", "kernel": "fused_tma_kernel (+ tile_desc pre-pass and crowd worklist kernel, timed together)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source,
                     "peak_source": peak_kind, "algorithmic_bytes_per_launch": fused_bytes,
                     "bytes_per_image": fused_bytes / n_img, "ms_per_launch": fused_ms,
                     "concurrent": "timed while the URL stream runs beside it" if overlap else "timed alone"},
        "url_chain": {"rows_per_s_per_gpu": (n_img + n_ref) / (url_ms_max * 1e-3), "dedup_plus_antijoin_main_rows_per_s_per_gpu": n_img / (anti_ms_max * 1e-3),
                      "algorithmic_bytes": url_bytes, "achieved_gbs": url_bytes / (url_ms_max * 1e-3) / 1e9,
                      "frac_of_peak": url_bytes / (url_ms_max * 1e-3) / 1e9 / peak,
                      "note": "K0 hash of main + reference URLs, K4 dedup, K5 anti-join" + (" incl. both cross-rank exchanges" if world > 1 else "")
                              + "; runs on the SMs the fused kernel leaves free, so its own fraction of peak is not the optimisation target"},
    }
    if e2e:
        line["e2e"] = e2e
    line.update(legs)
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        kind = cpu_kind()
        rows = 800 if kind == "reference" else 1200
        v1, slow1, _ = cpu_throughput(rows, 1, kind)
        vall, slow, wall = cpu_throughput(rows, cores, kind)
        line["cpu_baseline"] = {"value": vall, "unit": UNIT, "cores": cores, "kind": kind, "sample": cpu_sample_text(kind, rows, cores),
                                "one_core_value": v1}
        try:
            vd, dsteps = dropin_files_throughput(20_000, local)
            line["dropin_files"] = {"value": vd, "unit": UNIT, "rows": 20_000, "seconds": dsteps,
                                    "note": "this repo's processor.py step functions on CSV files, one process (native CSV reader/writer, native JSON "
                                            "ingest/egress, CUDA kernels; every output file written synchronously; a step finds the frame the previous step "
                                            "wrote in the process-level table cache instead of parsing the file again, the cache is emptied before every "
                                            "pass so the first step reads merged.csv from disk); value = rows / (dedup + reference filter + ptList->bbox + "
                                            "IoU filter seconds); seconds.phases = where the time went, summed over the four steps; "
                                            "seconds.labels = label remap + split of the rows that remain (not in value); "
                                            "the like-for-like figure against cpu_baseline (same files, same chain)"}
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
