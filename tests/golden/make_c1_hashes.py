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
    (HERE / "c1_hashes.json").write_text(json.dumps(out, indent=1, ensure_ascii=False) + "\n")
    print(json.dumps(out, indent=1, ensure_ascii=False))


if __name__ == "__main__":
    main()
