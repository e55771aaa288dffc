"""BASELINE.json config C1 (10 k rows as real CSV files) replayed through the oracle port (CPU) and
through the CUDA drop-in (GPU): every output file must have the SHA-256 the UNMODIFIED REFERENCE
produced for it (tests/golden/c1_hashes.json, written by tests/golden/make_c1_hashes.py), and so must every
frame of the label remap and the split that follow (the port's turn at those takes five minutes and runs
with DYD_SLOW_TESTS=1; the drop-in's runs with the GPU tests)."""
from __future__ import annotations

import contextlib
import io
import json
import os
from pathlib import Path

import pandas as pd
import pytest

from tests import c1_case

PINS = json.loads((Path(__file__).parent / "golden" / "c1_hashes.json").read_text())


def _replay(mod, tmp_path):
    if pd.__version__ != PINS["pandas"]:
        pytest.skip(f"pins were made with pandas {PINS['pandas']}; CSV text of other versions may differ")
    paths = c1_case.write_inputs(tmp_path)
    for k in ("merged", "ref"):
        assert c1_case.sha256(paths[k]) == PINS["files"][k], f"regenerated input {k} differs from the pinned one"
    for name, fn, produced in c1_case.steps(mod, tmp_path, paths):
        with contextlib.redirect_stdout(io.StringIO()):
            fn()
        for key, p in produced.items():
            assert c1_case.sha256(p) == PINS["files"][key], f"step {name}: {key}.csv differs from the reference's output"
    return tmp_path / "other70.csv"


def _replay_labels(mod, other70, tmp_path):
    """Label remap + split on the chain's other70.csv: every frame the reference wrote (to CSV or to an
    Excel sheet) must have the reference's digest, and the summaries must be equal."""
    df = pd.read_csv(other70, encoding="utf-8-sig")
    out, summary, diffs, unmatched = mod.remap_df(df, mod.mapping_from_frame(c1_case.mapping_frame()))
    remapped = tmp_path / "remapped.csv"
    out.to_csv(remapped, index=False, encoding="utf-8-sig")
    assert c1_case.sha256(remapped) == PINS["files"]["remapped"]
    assert {k: summary[k] for k in PINS["remap_summary"]} == PINS["remap_summary"]
    assert c1_case.frame_digest(pd.DataFrame(diffs)) == PINS["frames"]["remap_diff"]
    um = pd.DataFrame([{"标签": k, "数量": v} for k, v in unmatched.items()]).sort_values("数量", ascending=False)
    assert c1_case.frame_digest(um) == PINS["frames"]["remap_unmatched"]
    res = mod.split_df(pd.read_csv(remapped, encoding="utf-8-sig"), mod.rules_from_frame(c1_case.rules_frame()))
    assert json.loads(json.dumps(res["summary"], default=int, ensure_ascii=False)) == PINS["split_summary"]
    from deal_yolo_daya_b200.processor import _safe_filename as safe       # sheet files are named by it (utils.py:525-529)
    seen = set()
    for cat, parts in res["categories"].items():
        for name, part in parts.items():
            key = f"split/{safe(cat)}/{name}"
            assert c1_case.frame_digest(part) == PINS["frames"][key], key
            seen.add(key)
    assert c1_case.frame_digest(res["unclassified"]) == PINS["frames"]["split/unclassified/Sheet1"]
    assert c1_case.frame_digest(res["split_counts"]) == PINS["frames"]["split/split_counts/Sheet1"]
    assert seen == {k for k in PINS["frames"] if k.startswith("split/") and k.split("/")[1] not in ("unclassified", "split_counts")}


def test_oracle_port_reproduces_reference_files(tmp_path):
    from oracle import pipeline_port

    enc = "utf-8-sig"

    class Port:      # the port's frame-level steps behind the reference's file-level signatures
        @staticmethod
        def deduplicate_csv_by_source(src, out):
            pipeline_port.dedup_df(pd.read_csv(src, encoding=enc, parse_dates=False)).to_csv(out, index=False, encoding=enc)

        @staticmethod
        def remove_duplicates_between_csv(main, ref, out):
            pipeline_port.ref_filter_df(pd.read_csv(main, encoding=enc, parse_dates=False),
                                        pd.read_csv(ref, encoding=enc, parse_dates=False)).to_csv(out, index=False, encoding=enc)

        @staticmethod
        def process_csv_replace_ptlist(src, out, exc):
            r, _ = pipeline_port.replace_ptlist_df(pd.read_csv(src, encoding=enc))
            r.to_csv(out, index=False, encoding=enc)

        @staticmethod
        def filter_by_box_count_and_iou(src, hi, other, min_boxes, thr):
            h, o = pipeline_port.iou_split_df(pd.read_csv(src, encoding=enc), min_boxes, thr)
            h.to_csv(hi, index=False, encoding=enc); o.to_csv(other, index=False, encoding=enc)
    other70 = _replay(Port, tmp_path)
    if os.environ.get("DYD_SLOW_TESTS") == "1":      # the port expands rows one by one like the reference: ~5 minutes
        _replay_labels(pipeline_port, other70, tmp_path)


def test_dropin_label_host_logic_reproduces_reference_frames(tmp_path, monkeypatch):
    """The drop-in's remap / split host logic (dictionary encoding, LUTs, frame assembly) with the kernels
    emulated by the C oracle, on the reference's own C1 `other70.csv`: all 18 frame digests."""
    from deal_yolo_daya_b200 import processor as P
    from oracle import pipeline_port
    from tests.oracle_kernels import OracleKernels
    if pd.__version__ != PINS["pandas"]:
        pytest.skip(f"pins were made with pandas {PINS['pandas']}")
    monkeypatch.setattr(P, "KERNELS", OracleKernels())
    paths = c1_case.write_inputs(tmp_path)
    enc = "utf-8-sig"
    d = pipeline_port.dedup_df(pd.read_csv(paths["merged"], encoding=enc, parse_dates=False))
    f = pipeline_port.ref_filter_df(d, pd.read_csv(paths["ref"], encoding=enc, parse_dates=False))
    r, _ = pipeline_port.replace_ptlist_df(f)
    r.to_csv(tmp_path / "rep.csv", index=False, encoding=enc)
    _, other = pipeline_port.iou_split_df(pd.read_csv(tmp_path / "rep.csv", encoding=enc), 2, 0.7)
    other.to_csv(tmp_path / "other70.csv", index=False, encoding=enc)
    assert c1_case.sha256(tmp_path / "other70.csv") == PINS["files"]["other70"]
    _replay_labels(P, tmp_path / "other70.csv", tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("io", ["native-io", "pandas-io"])
def test_cuda_dropin_reproduces_reference_files(tmp_path, cuda_device, monkeypatch, io):
    from deal_yolo_daya_b200 import processor as P
    monkeypatch.setattr(P, "KERNELS", P.CudaKernels(cuda_device.index))
    if io == "pandas-io":
        monkeypatch.setenv("DYD_NATIVE_INGEST", "0")
    other70 = _replay(P, tmp_path)
    if io == "native-io":
        _replay_labels(P, other70, tmp_path)
