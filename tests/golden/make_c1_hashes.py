"""SHA-256 pins of the UNMODIFIED REFERENCE's outputs on BASELINE.json's config C1 (10 k synthetic rows
as real CSV files): merged -> dedup -> ref-filter -> ptList->bbox -> IoU filter (thr 0.7 and 0.98).

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_c1_hashes.py

The inputs are regenerated at test time from the seeded generator (deal_yolo_daya_b200/synth.py) and
must hash to the recorded input digests before any output is compared; nothing at test time reads
/root/reference.  tests/c1_case.py holds the shared input builder."""
from __future__ import annotations

import contextlib
import hashlib
import io
import json
import sys
import tempfile
import time
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

from src.deal_yolo_data.core import processor as ref  # noqa: E402
import importlib.util  # noqa: E402
_spec = importlib.util.spec_from_file_location("c1_case", ROOT / "tests" / "c1_case.py")
c1_case = importlib.util.module_from_spec(_spec); _spec.loader.exec_module(c1_case)


def main():
    out = {"pandas": __import__("pandas").__version__, "rows": c1_case.ROWS, "files": {}, "seconds": {}}
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        paths = c1_case.write_inputs(td)
        for k in ("merged", "ref"):
            out["files"][k] = c1_case.sha256(paths[k])
        steps = c1_case.steps(ref, td, paths)
        for name, fn, produced in steps:
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                fn()
            out["seconds"][name] = round(time.perf_counter() - t0, 3)
            for key, p in produced.items():
                out["files"][key] = c1_case.sha256(p)
                out.setdefault("lines", {})[key] = sum(1 for _ in open(p, "rb"))
        # label remap + split through the unmodified reference; openpyxl is absent, so Excel I/O goes
        # through the same in-memory shim tests/golden/make_golden.py uses (reference code untouched)
        import pandas as pd
        book = {}

        class _Writer:
            def __init__(self, path, *a, **k):
                self.path = str(path); book[self.path] = {}

            def __enter__(self):
                return self

            def __exit__(self, *a):
                return False

        def _read_excel(path, sheet_name=None, **k):
            b = book[str(path)]
            return b[sheet_name or next(iter(b))].copy()

        def _to_excel(self, target, sheet_name="Sheet1", index=True, **k):
            if isinstance(target, _Writer):
                book[target.path][sheet_name] = self.copy()
            else:
                book.setdefault(str(target), {})[sheet_name] = self.copy()
        pd.read_excel, pd.ExcelWriter, pd.DataFrame.to_excel = _read_excel, _Writer, _to_excel
        p_map = td / "map.xlsx"; book[str(p_map)] = {"Sheet1": c1_case.mapping_frame()}
        p_rules = td / "rules.xlsx"; p_rules.touch(); book[str(p_rules)] = {"Sheet1": c1_case.rules_frame()}
        p_remap = td / "remapped.csv"
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            m = ref.replace_labels_by_mapping(str(td / "other70.csv"), str(p_map), str(p_remap), None, None, None, None,
                                              str(td / "diff.xlsx"), str(td / "unm.xlsx"))
        out["seconds"]["remap"] = round(time.perf_counter() - t0, 3)
        out["files"]["remapped"] = c1_case.sha256(p_remap)
        out["remap_summary"] = {k: (int(v) if hasattr(v, "__int__") and not isinstance(v, (str, bool)) else v) for k, v in m["summary"].items()}
        out["frames"] = {"remap_diff": c1_case.frame_digest(book[str(td / "diff.xlsx")]["Sheet1"]),
                         "remap_unmatched": c1_case.frame_digest(book[str(td / "unm.xlsx")]["Sheet1"])}
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            sres = ref.split_dataset_by_rules(str(p_remap), str(p_rules), str(td / "split"))
        out["seconds"]["split"] = round(time.perf_counter() - t0, 3)
        out["split_summary"] = json.loads(json.dumps(sres["summary"], default=lambda o: int(o) if hasattr(o, "__int__") else str(o), ensure_ascii=False))
        for path, sheets in book.items():
            if str(td / "split") in path:
                for sh, df in sheets.items():
                    out["frames"][f"split/{Path(path).stem}/{sh}"] = c1_case.frame_digest(df)
    (HERE / "c1_hashes.json").write_text(json.dumps(out, indent=1, ensure_ascii=False) + "\n")
    print(json.dumps(out, indent=1, ensure_ascii=False))


if __name__ == "__main__":
    main()
