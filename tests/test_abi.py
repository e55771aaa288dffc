"""The C-ABI library loads on a CPU-only box and exports every symbol include/dyd.h declares
(no compute calls here: those need a GPU)."""
from __future__ import annotations

import ctypes
import re
from pathlib import Path

from deal_yolo_daya_b200 import _lib, build

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols(header="dyd.h"):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dyd_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    build.build()
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dyd.h but not exported by libdyd.so"


def test_binding_matches_header():
    assert sorted(_lib.PROTOTYPES) == declared_symbols()
    lib = _lib.load()
    assert lib.dyd_version() == 100


def test_generator_library_is_separate_from_the_product_library():
    """The synthetic-table generator (bench / tests) has its own header and library; libdyd.so does not export it."""
    build.build()
    names = declared_symbols("dyd_synth.h")
    assert sorted(_lib.SYNTH_PROTOTYPES) == names and len(names) >= 6
    synth = ctypes.CDLL(str(_lib.SYNTH_LIB_PATH))
    product = ctypes.CDLL(str(_lib.LIB_PATH))
    for n in names:
        assert hasattr(synth, n)
        assert not hasattr(product, n), f"{n} leaked into the product ABI"


def test_argument_errors_are_reported_without_a_gpu():
    lib = _lib.load()
    rc = lib.dyd_bbox_minmax(None, None, -1, None, None, None, None)
    assert rc == -1 and "negative" in _lib.last_error()
    assert lib.dyd_dedup_workspace_bytes(1000) >= 2048 * 16
    assert lib.dyd_iou_workspace_bytes(0) > 0


def test_built_for_sm_100a():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
