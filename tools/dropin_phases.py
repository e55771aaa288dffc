#!/usr/bin/env python
"""Wall time of the drop-in's four file-level steps with the per-phase timers of processor.PHASES (developer tool).

    python tools/dropin_phases.py [rows] [--oracle] [--nocache]

--oracle emulates the kernels with the CPU oracle (tests/oracle_kernels.py) so the HOST phases can be looked at in a
container without a GPU; the kernel phase is then meaningless.  Without it the CUDA facade runs (GPU box).
"""
import contextlib, io, json, os, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import pandas as pd
from deal_yolo_daya_b200 import processor as P, synth

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if args else 20000
if "--oracle" in sys.argv:
    from tests.oracle_kernels import OracleKernels
    P.KERNELS = OracleKernels()
cold = "--cold" in sys.argv
if "--nocache" in sys.argv:
    os.environ["DYD_TABLE_CACHE"] = "0"
sys.path.insert(0, str(ROOT))
import bench

with tempfile.TemporaryDirectory(dir=os.environ.get("DYD_TMP")) as td:
    td = Path(td)
    merged, ref = bench._write_sample(td, 0, 0, n)
    print("csv bytes", merged.stat().st_size, "ref bytes", ref.stat().st_size)
    steps = [("dedup", lambda: P.deduplicate_csv_by_source(str(merged), str(td / "dedup.csv"))),
             ("ref_filter", lambda: P.remove_duplicates_between_csv(str(td / "dedup.csv"), str(ref), str(td / "filtered.csv"))),
             ("replace", lambda: P.process_csv_replace_ptlist(str(td / "filtered.csv"), str(td / "rep.csv"), str(td / "exc.csv"))),
             ("iou", lambda: P.filter_by_box_count_and_iou(str(td / "rep.csv"), str(td / "hi.csv"), str(td / "other.csv"), 2, 0.7))]
    from deal_yolo_daya_b200 import tablecache
    for rep in range(3):
        tot = 0.0
        if cold:
            tablecache.clear()
        for name, fn in steps:
            if hasattr(P, "PHASES"):
                P.PHASES.clear()
            with contextlib.redirect_stdout(io.StringIO()):
                t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
            tot += dt
            ph = {k: round(v, 4) for k, v in getattr(P, "PHASES", {}).items()}
            print(f"pass {rep} {name:10s} {dt:.3f} s  {json.dumps(ph)}")
        print(f"pass {rep} total {tot:.3f} s = {n / tot:.0f} images/s")
