"""The drop-in layer end to end through the CUDA kernels, byte-for-byte against the reference's
golden outputs (tests/golden/, produced by running the unmodified reference)."""
from __future__ import annotations

import pytest

from deal_yolo_daya_b200 import processor as P
from tests import dropin_checks as C

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["native-io", "pandas-io"])
def _cuda_facade(request, cuda_device, monkeypatch):
    """Every check runs twice: with the native CSV reader / writer / JSON lanes forced on even for the
    small golden files, and with pandas + CPython json doing the host work (kernels identical)."""
    monkeypatch.setattr(P, "KERNELS", P.CudaKernels(cuda_device.index))
    if request.param == "native-io":
        monkeypatch.setenv("DYD_CSV_NATIVE_MIN_BYTES", "0")
    else:
        monkeypatch.setenv("DYD_NATIVE_INGEST", "0")
    yield request.param


def test_dedup(tmp_path):
    C.check_dedup(tmp_path)


def test_ref_filter(tmp_path):
    C.check_ref_filter(tmp_path)


def test_replace_ptlist(tmp_path, _cuda_facade):
    from deal_yolo_daya_b200 import native
    before = native._READ_STATS["native"]
    C.check_replace(tmp_path)
    assert P.STATS["hostlane_objects"] == 0
    if _cuda_facade == "native-io" and native._pandas_infers_arrow_str():
        assert native._READ_STATS["native"] > before          # the golden input went through csrc/csv_read.cpp


def test_iou_filter(tmp_path):
    C.check_iou(tmp_path)
    assert P.STATS["hostlane_rows"] == 0


def test_remap():
    C.check_remap()


def test_split():
    C.check_split()


def test_signatures_equal_the_reference():
    from tests import dataset_checks as D
    D.check_signatures()


def test_merge(tmp_path):
    from tests import dataset_checks as D
    D.check_merge(tmp_path)


def test_yolo_writer_and_summaries(tmp_path):
    from tests import dataset_checks as D
    D.check_yolo_and_summaries(tmp_path)
