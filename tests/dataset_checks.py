"""The round-2 rows of the drop-in (merge, YOLO dataset writer, summaries) against fixtures produced by running the
unmodified reference (tests/golden/make_golden_r2.py).  Shared by the CPU host-logic test (oracle kernel stand-in) and the
GPU test (CUDA facade); the only difference is which kernel facade is installed."""
from __future__ import annotations

import contextlib
import inspect
import io
import json
from pathlib import Path

import pandas as pd
import pytest

from deal_yolo_daya_b200 import processor as P
from tests import excel_shim, r2_cases

FX = json.loads((r2_cases.G / "r2" / "fixtures.json").read_text(encoding="utf-8"))


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        r = fn(*a, **k)
    return r, buf.getvalue()


def tree_texts(root: Path):
    return {str(p.relative_to(root)): p.read_text(encoding="utf-8") for p in sorted(root.rglob("*")) if p.is_file() and p.suffix in (".txt", ".yaml")}


def check_signatures():
    """The drop-in surface equals the reference's, parameter for parameter (names, order, defaults, annotations)."""
    for name, sig in FX["signatures"].items():
        assert str(inspect.signature(getattr(P, name), eval_str=True)) == sig, name   # (this package defers annotations)


def check_merge(tmp: Path):
    r2_cases.write_merge_inputs(tmp / "in")
    for chunk in (100000, 4):
        want = FX[f"merge_{chunk}"]
        with r2_cases.sorted_glob():
            n, log = quiet(P.merge_all_csv_in_folder, str(tmp / "in"), str(tmp / "out" / f"merged_{chunk}.csv"), "utf-8-sig", chunk)
        assert n == want["rows"]
        assert (tmp / "out" / f"merged_{chunk}.csv").read_bytes().hex() == want["bytes_hex"], chunk
        assert log.replace(str(tmp / "out"), "{TMP}") == want["log"]
    (tmp / "emptydir").mkdir()
    n, log = quiet(P.merge_all_csv_in_folder, str(tmp / "emptydir"), str(tmp / "x.csv"))
    assert n is None and log.replace(str(tmp), "{TMP}") == FX["merge_empty"]["log"]
    with pytest.raises(FileNotFoundError):
        P.merge_all_csv_in_folder(str(tmp / "no_such_folder"))
    calls = []
    with r2_cases.sorted_glob():
        quiet(P.merge_all_csv_in_folder, str(tmp / "in"), str(tmp / "cb.csv"), "utf-8-sig", 4, lambda *a: calls.append(a))
    assert calls and calls[-1][3] == FX["merge_4"]["rows"] and all(len(c) == 10 for c in calls)


def check_yolo_and_summaries(tmp: Path):
    want = FX["yolo"]
    with excel_shim.installed():
        books = r2_cases.yolo_books(tmp / "imgs")
        paths = []
        for cat, sheets in books.items():
            excel_shim.put_book(tmp / "split" / f"{cat}.xlsx", sheets)
            paths.append(str(tmp / "split" / f"{cat}.xlsx"))
        kw = dict(download_images=False, class_order=["grp3", "grp1", "not_a_class"])
        res, _ = quiet(P.generate_yolo_datasets_from_excels, paths, str(tmp / "yolo"), None, **kw)
        files = {k: v.replace(str(tmp), "{TMP}") for k, v in tree_texts(tmp / "yolo").items()}
        assert sorted(files) == sorted(want["files"])
        for k in files:
            assert files[k] == want["files"][k], k
        assert excel_shim.BOOK[str(res["skipped"])]["Sheet1"].to_csv(index=False) == want["skipped_csv"]
        assert res["stats"] == want["stats"] and res["total"] == want["total"] and res["processed"] == want["processed"]
        assert res["downloaded"] == want["downloaded"] and res["dataset_name_map"] == want["dataset_name_map"]
        assert [str(Path(d).relative_to(tmp)) for d in res["datasets"]] == want["datasets"]
        # second run: resume keeps every label file that exists
        res2, _ = quiet(P.generate_yolo_datasets_from_excels, paths, str(tmp / "yolo"), None, **kw)
        w2 = FX["yolo_resume"]
        assert res2["stats"] == w2["stats"] and res2["processed"] == w2["processed"] and res2["downloaded"] == w2["downloaded"]
        assert excel_shim.BOOK[str(res2["skipped"])]["Sheet1"].to_csv(index=False) == w2["skipped_csv"]
        with pytest.raises(NameError):                          # the reference dies at the end when given a callback (:1076-1077)
            quiet(P.generate_yolo_datasets_from_excels, paths[:1], str(tmp / "yolo_cb"), None, progress_callback=lambda *a: None, **kw)
        # ---- label-count summary of what was written
        stats, flat = P.summarize_yolo_label_counts(res["datasets"] + [None, str(tmp / "no_such_dataset")])
        assert stats == FX["label_counts"]["stats"]
        assert flat.sort_values(list(flat.columns)).to_csv(index=False) == FX["label_counts"]["flat_sorted_csv"]
        # ---- unclassified summary
        unc = pd.read_csv(io.StringIO(r2_cases.gz_text(r2_cases.G / "expected" / "split__unclassified__Sheet1.csv.gz")))
        excel_shim.put_book(tmp / "split" / "unclassified.xlsx", {"Sheet1": unc})
        out = P.summarize_unclassified(str(tmp / "split" / "unclassified.xlsx"), str(tmp / "summ"))
        assert {k: v.to_csv(index=False) for k, v in excel_shim.BOOK[str(out)].items()} == FX["unclassified"]
        excel_shim.put_book(tmp / "split" / "unclassified2.xlsx", {"Sheet1": unc.drop(columns=["无法分类标签"])})
        out = P.summarize_unclassified(str(tmp / "split" / "unclassified2.xlsx"), str(tmp / "summ2"))
        assert {k: v.to_csv(index=False) for k, v in excel_shim.BOOK[str(out)].items()} == FX["unclassified_no_label_column"]
        with pytest.raises(FileNotFoundError):
            P.summarize_unclassified(str(tmp / "split" / "nope.xlsx"), str(tmp / "summ3"))
