"""Host logic of the drop-in (ingest, egress, file/error contract) on CPU, kernels replaced by the
oracle stand-in from tests/oracle_kernels.py.  Byte-for-byte against the reference's outputs."""
from __future__ import annotations

import pytest

from deal_yolo_daya_b200 import processor as P
from tests import dropin_checks as C
from tests.oracle_kernels import OracleKernels


@pytest.fixture(autouse=True)
def _oracle_facade(monkeypatch):
    monkeypatch.setattr(P, "KERNELS", OracleKernels())


def test_dedup(tmp_path):
    C.check_dedup(tmp_path)


def test_ref_filter(tmp_path):
    C.check_ref_filter(tmp_path)


def test_replace_ptlist(tmp_path):
    C.check_replace(tmp_path)


def test_iou_filter(tmp_path):
    C.check_iou(tmp_path)


def test_remap():
    C.check_remap()


def test_split():
    C.check_split()


def test_signatures_equal_the_reference():
    from tests import dataset_checks as D
    D.check_signatures()


@pytest.mark.parametrize("native_min", ["0", None])
def test_merge(tmp_path, monkeypatch, native_min):
    from tests import dataset_checks as D
    if native_min is not None:
        monkeypatch.setenv("DYD_CSV_NATIVE_MIN_BYTES", native_min)     # the small inputs take the native reader
    D.check_merge(tmp_path)


def test_yolo_writer_and_summaries(tmp_path):
    from tests import dataset_checks as D
    D.check_yolo_and_summaries(tmp_path)


def test_product_refuses_to_run_without_cuda(monkeypatch):
    """Without the oracle facade and without a GPU the drop-in must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pandas as pd
    from deal_yolo_daya_b200 import _lib
    monkeypatch.setattr(P, "KERNELS", P.CudaKernels())
    with pytest.raises(_lib.DydError):
        P.deduplicate_df(pd.DataFrame({"source": ["a", "b", "a"]}))
