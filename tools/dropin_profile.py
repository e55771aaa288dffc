#!/usr/bin/env python
"""cProfile of the drop-in's three file-level steps on a synthetic C2-style CSV (developer tool)."""
import contextlib, cProfile, io, os, pstats, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import pandas as pd
from deal_yolo_daya_b200 import processor as P, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
t = synth.make_table(0, 0, n)
rows = synth.table_to_rows(t)
with tempfile.TemporaryDirectory() as td:
    td = Path(td)
    merged = td / "merged.csv"
    pd.DataFrame(rows, columns=[P.COL_SRC, P.COL_ANN]).to_csv(merged, index=False, encoding="utf-8-sig")
    print("csv bytes", merged.stat().st_size)
    steps = [("dedup", lambda: P.deduplicate_csv_by_source(str(merged), str(td / "dedup.csv"))),
             ("replace", lambda: P.process_csv_replace_ptlist(str(td / "dedup.csv"), str(td / "rep.csv"), str(td / "exc.csv"))),
             ("iou", lambda: P.filter_by_box_count_and_iou(str(td / "rep.csv"), str(td / "hi.csv"), str(td / "other.csv"), 2, 0.7))]
    for name, fn in steps:                       # warm-up pass (library load, CUDA context)
        with contextlib.redirect_stdout(io.StringIO()):
            fn()
    for name, fn in steps:
        pr = cProfile.Profile()
        with contextlib.redirect_stdout(io.StringIO()):
            t0 = time.perf_counter(); pr.enable(); fn(); pr.disable(); dt = time.perf_counter() - t0
        print(f"==== {name}: {dt:.3f} s")
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
        print("\n".join(l for l in s.getvalue().splitlines() if l.strip() and not l.startswith("   Ordered"))[:3500])
