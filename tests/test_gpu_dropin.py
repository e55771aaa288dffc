"""The drop-in layer end to end through the CUDA kernels, byte-for-byte against the reference's
golden outputs (tests/golden/, produced by running the unmodified reference)."""
from __future__ import annotations

import pytest

from deal_yolo_daya_b200 import processor as P
from tests import dropin_checks as C

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _cuda_facade(cuda_device, monkeypatch):
    monkeypatch.setattr(P, "KERNELS", P.CudaKernels(cuda_device.index))


def test_dedup(tmp_path):
    C.check_dedup(tmp_path)


def test_ref_filter(tmp_path):
    C.check_ref_filter(tmp_path)


def test_replace_ptlist(tmp_path):
    C.check_replace(tmp_path)
    assert P.STATS["hostlane_objects"] == 0


def test_iou_filter(tmp_path):
    C.check_iou(tmp_path)
    assert P.STATS["hostlane_rows"] == 0


def test_remap():
    C.check_remap()


def test_split():
    C.check_split()
