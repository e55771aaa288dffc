"""Golden fixtures of the round-2 rows by RUNNING THE UNMODIFIED REFERENCE (authoring container only):

    python tests/golden/make_golden_r2.py

merge_all_csv_in_folder (processor.py:26-109), generate_yolo_datasets_from_excels (:893-1087),
summarize_yolo_label_counts (:1089-1163), summarize_unclassified (:833-891); plus the `inspect.signature` strings of every
function of the drop-in surface.  Excel I/O goes through tests/excel_shim.py (openpyxl is absent); Path.glob is pinned to
sorted order for the merge (tests/r2_cases.sorted_glob).  Nothing at test time reads /root/reference.
"""
from __future__ import annotations

import contextlib
import inspect
import io
import json
import sys
import tempfile
from pathlib import Path

import pandas as pd

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(ROOT))                     # the repo's `tests` package must win over the reference's empty one

from src.deal_yolo_data.core import processor as ref  # noqa: E402
from tests import excel_shim, r2_cases  # noqa: E402

OUT = HERE / "r2"
OUT.mkdir(exist_ok=True)


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        r = fn(*a, **k)
    return r, buf.getvalue()


def tree_texts(root: Path):
    return {str(p.relative_to(root)): p.read_text(encoding="utf-8") for p in sorted(root.rglob("*")) if p.is_file() and p.suffix in (".txt", ".yaml")}


def main():
    fx = {}
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        # ---- merge
        r2_cases.write_merge_inputs(td / "in")
        for chunk in (100000, 4):
            with r2_cases.sorted_glob():
                n, log = quiet(ref.merge_all_csv_in_folder, str(td / "in"), str(td / f"merged_{chunk}.csv"), "utf-8-sig", chunk)
            fx[f"merge_{chunk}"] = {"rows": n, "log": log.replace(str(td), "{TMP}"), "bytes_hex": (td / f"merged_{chunk}.csv").read_bytes().hex()}
        (td / "emptydir").mkdir()
        n, log = quiet(ref.merge_all_csv_in_folder, str(td / "emptydir"), str(td / "x.csv"))
        fx["merge_empty"] = {"rows": n, "log": log.replace(str(td), "{TMP}")}
        # ---- YOLO dataset writer + label-count summary
        with excel_shim.installed():
            books = r2_cases.yolo_books(td / "imgs")
            paths = []
            for cat, sheets in books.items():
                excel_shim.put_book(td / "split" / f"{cat}.xlsx", sheets)
                paths.append(str(td / "split" / f"{cat}.xlsx"))
            res, log = quiet(ref.generate_yolo_datasets_from_excels, paths, str(td / "yolo"), None, download_images=False,
                             class_order=["grp3", "grp1", "not_a_class"])
            skipped = excel_shim.BOOK[str(res["skipped"])]["Sheet1"]
            fx["yolo"] = {"files": tree_texts(td / "yolo"), "skipped_csv": skipped.to_csv(index=False), "stats": res["stats"],
                          "total": res["total"], "processed": res["processed"], "downloaded": res["downloaded"],
                          "dataset_name_map": res["dataset_name_map"], "datasets": [str(Path(d).relative_to(td)) for d in res["datasets"]]}
            fx["yolo"]["files"] = {k: v.replace(str(td), "{TMP}") for k, v in fx["yolo"]["files"].items()}
            res2, _ = quiet(ref.generate_yolo_datasets_from_excels, paths, str(td / "yolo"), None, download_images=False,
                            class_order=["grp3", "grp1", "not_a_class"])
            fx["yolo_resume"] = {"stats": res2["stats"], "processed": res2["processed"], "downloaded": res2["downloaded"],
                                 "skipped_csv": excel_shim.BOOK[str(res2["skipped"])]["Sheet1"].to_csv(index=False)}
            stats, flat = ref.summarize_yolo_label_counts(res["datasets"] + [None, str(td / "no_such_dataset")])
            fx["label_counts"] = {"stats": stats, "flat_sorted_csv": flat.sort_values(list(flat.columns)).to_csv(index=False)}
            # ---- unclassified summary
            unc = pd.read_csv(io.StringIO(r2_cases.gz_text(r2_cases.G / "expected" / "split__unclassified__Sheet1.csv.gz")))
            excel_shim.put_book(td / "split" / "unclassified.xlsx", {"Sheet1": unc})
            out = ref.summarize_unclassified(str(td / "split" / "unclassified.xlsx"), str(td / "summ"))
            fx["unclassified"] = {k: v.to_csv(index=False) for k, v in excel_shim.BOOK[str(out)].items()}
            unc2 = unc.drop(columns=["无法分类标签"]) if "无法分类标签" in unc.columns else unc
            excel_shim.put_book(td / "split" / "unclassified2.xlsx", {"Sheet1": unc2})
            out = ref.summarize_unclassified(str(td / "split" / "unclassified2.xlsx"), str(td / "summ2"))
            fx["unclassified_no_label_column"] = {k: v.to_csv(index=False) for k, v in excel_shim.BOOK[str(out)].items()}
    # ---- the drop-in surface: signature strings of the unmodified reference
    names = ["merge_all_csv_in_folder", "deduplicate_csv_by_source", "remove_duplicates_between_csv", "process_csv_replace_ptlist",
             "filter_by_box_count_and_iou", "replace_labels_by_mapping", "split_dataset_by_rules", "summarize_unclassified",
             "generate_yolo_datasets_from_excels", "summarize_yolo_label_counts"]
    fx["signatures"] = {n: str(inspect.signature(getattr(ref, n))) for n in names}
    (OUT / "fixtures.json").write_text(json.dumps(fx, ensure_ascii=False, indent=1), encoding="utf-8")
    print("wrote", OUT / "fixtures.json", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in fx.items()})


if __name__ == "__main__":
    main()
