"""BASELINE.json config C1 (10 k rows as real CSV files) replayed through the oracle port (CPU) and
through the CUDA drop-in (GPU): every output file must have the SHA-256 the UNMODIFIED REFERENCE
produced for it (tests/golden/c1_hashes.json, written by tests/golden/make_c1_hashes.py)."""
from __future__ import annotations

import contextlib
import io
import json
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
    _replay(Port, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("io", ["native-io", "pandas-io"])
def test_cuda_dropin_reproduces_reference_files(tmp_path, cuda_device, monkeypatch, io):
    from deal_yolo_daya_b200 import processor as P
    monkeypatch.setattr(P, "KERNELS", P.CudaKernels(cuda_device.index))
    if io == "pandas-io":
        monkeypatch.setenv("DYD_NATIVE_INGEST", "0")
    _replay(P, tmp_path)
