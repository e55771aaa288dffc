"""The drop-in layer against the reference's golden outputs (shared by the CPU host-logic test
and the GPU test; the only difference is which kernel facade is installed)."""
from __future__ import annotations

import contextlib
import gzip
import io
import json
from pathlib import Path

import pandas as pd
import pytest

from deal_yolo_daya_b200 import processor as P

G = Path(__file__).resolve().parent / "golden"
SUMMARY = json.loads((G / "expected" / "summary.json").read_text(encoding="utf-8"))


def gz_text(p):
    return gzip.open(p, "rt", encoding="utf-8").read()


def exp_text(name):
    return gz_text(G / "expected" / name)


def put(tmp, name, text):
    p = tmp / name
    p.write_text(text, encoding="utf-8-sig")
    return str(p)


def got(path):
    return Path(path).read_text(encoding="utf-8-sig")


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        r = fn(*a, **k)
    return r, buf.getvalue()


def check_dedup(tmp):
    src = put(tmp, "merged.csv", gz_text(G / "inputs" / "merged.csv.gz"))
    for keep in ("first", "last"):
        out = str(tmp / f"d_{keep}.csv")
        df, log = quiet(P.deduplicate_csv_by_source, src, out, "utf-8-sig", keep)
        assert got(out) == exp_text(f"dedup_{keep}.csv.gz")
        assert len(df) == SUMMARY[f"dedup_{keep}_rows"]
        assert "去重策略" in log and "去除重复数据行数" in log
    with pytest.raises(FileNotFoundError):
        P.deduplicate_csv_by_source(str(tmp / "nope.csv"))
    bad = tmp / "x.txt"; bad.write_text("a")
    with pytest.raises(ValueError):
        P.deduplicate_csv_by_source(str(bad))
    nosrc = put(tmp, "nosrc.csv", "a,b\n1,2\n")
    with pytest.raises(KeyError):
        quiet(P.deduplicate_csv_by_source, nosrc, None)


def check_ref_filter(tmp):
    main = put(tmp, "dedup.csv", exp_text("dedup_first.csv.gz"))
    ref = put(tmp, "reference.csv", gz_text(G / "inputs" / "reference.csv.gz"))
    out = str(tmp / "sub" / "filtered.csv")
    df, log = quiet(P.remove_duplicates_between_csv, main, ref, out)
    assert got(out) == exp_text("filtered_main.csv.gz")
    assert len(df) == SUMMARY["ref_filter_rows"] and "剔除重复行数" in log
    with pytest.raises(KeyError):
        quiet(P.remove_duplicates_between_csv, main, ref, out, "nocol")


def check_replace(tmp):
    src = put(tmp, "filtered.csv", exp_text("filtered_main.csv.gz"))
    out, exc = str(tmp / "rep.csv"), str(tmp / "exc.csv")
    res, _ = quiet(P.process_csv_replace_ptlist, src, out, exc)
    assert got(out) == exp_text("processed_replaced_ptlist.csv.gz")
    assert got(exc) == exp_text("processed_excluded.csv.gz")
    assert res["filtered_rows"] == SUMMARY["replace"]["filtered_rows"] and res["excluded_rows"] == SUMMARY["replace"]["excluded_rows"]
    r, log = quiet(P.process_csv_replace_ptlist, str(tmp / "missing.csv"), out)
    assert r is None and "未找到文件" in log
    r, log = quiet(P.process_csv_replace_ptlist, put(tmp, "nocol.csv", "source\na\n"), out)
    assert r is None and "缺少列" in log
    crash = put(tmp, "crash.csv", gz_text(G / "inputs" / "crash_none.csv.gz"))
    with pytest.raises(TypeError):
        quiet(P.process_csv_replace_ptlist, crash, out, None)


def check_iou(tmp):
    src = put(tmp, "rep.csv", exp_text("processed_replaced_ptlist.csv.gz"))
    for thr, mb in ((0.7, 2), (0.98, 2), (0.7, 3), (0.0, 1)):
        hi, ot = str(tmp / "hi.csv"), str(tmp / "ot.csv")
        r, _ = quiet(P.filter_by_box_count_and_iou, src, hi, ot, mb, thr)
        assert r is None
        tag = f"{thr:.2f}_{mb}"
        assert got(hi) == exp_text(f"high_iou_{tag}.csv.gz"), tag
        assert got(ot) == exp_text(f"other_{tag}.csv.gz"), tag
    r, log = quiet(P.filter_by_box_count_and_iou, put(tmp, "nonew.csv", "source\na\n"), str(tmp / "h"), str(tmp / "o"))
    assert r is None and "缺少必要列" in log


def check_remap():
    df = pd.read_csv(io.StringIO(exp_text("other_0.70_2.csv.gz")))
    mapping = pd.DataFrame(json.loads((G / "inputs" / "mapping.json").read_text(encoding="utf-8")))
    lm = P.mapping_from_frame(mapping)
    out, summary, diffs, unmatched = P.remap_df(df, lm)
    assert out.to_csv(index=False) == exp_text("other_data_label_replaced.csv.gz")
    assert summary == SUMMARY["remap"]["summary"]
    assert pd.DataFrame(diffs).to_csv(index=False) == exp_text("remap_diff.csv.gz")
    um = pd.DataFrame([{"标签": k, "数量": v} for k, v in unmatched.items()]).sort_values("数量", ascending=False)
    assert um.to_csv(index=False) == exp_text("remap_unmatched.csv.gz")
    assert diffs[:30] == SUMMARY["remap"]["sample_diff"]


def check_split():
    df = pd.read_csv(io.StringIO(exp_text("other_data_label_replaced.csv.gz")))
    rules = pd.DataFrame(json.loads((G / "inputs" / "rules.json").read_text(encoding="utf-8")))
    res = P.split_df(df, P.rules_from_frame(rules))
    assert res["summary"] == SUMMARY["split"]
    assert [f"{P._safe_filename(c)}.xlsx" for c in res["categories"]] == SUMMARY["split_files"]
    for cat, parts in res["categories"].items():
        for name, part in parts.items():
            assert part.to_csv(index=False) == exp_text(f"split__{cat}__{name}.csv.gz"), (cat, name)
    assert res["unclassified"].to_csv(index=False) == exp_text("split__unclassified__Sheet1.csv.gz")
    assert res["split_counts"].to_csv(index=False) == exp_text("split__split_counts__Sheet1.csv.gz")
