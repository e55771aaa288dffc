"""ctypes binding of libdyd.so (include/dyd.h).  Fails loudly: there is no CPU fallback.

The library is built in-tree by ``deal_yolo_daya_b200.build`` (nvcc, sm_100a).  Loading it
does not need a GPU; calling a compute entry point does.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libdyd.so"

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_int = C.c_int
_u64 = C.c_uint64
_f64 = C.c_double
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/dyd.h one to one
PROTOTYPES = {
    "dyd_version": (_int, []),
    "dyd_launch_count": (_u64, []),
    "dyd_last_error": (_sz, [C.c_char_p, _sz]),
    "dyd_bbox_minmax": (_int, [_p, _p, _i64, _p, _p, _p, _p]),
    "dyd_iou_workspace_bytes": (_sz, [_i64]),
    "dyd_iou_filter": (_int, [_p, _p, _p, _i64, _i64, _f64, _p, _p, _p, _sz, _p]),
    "dyd_bbox_iou_fused": (_int, [_p, _p, _p, _i64, _i64, _i64, _f64, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "dyd_hash_strings": (_int, [_p, _p, _i64, _p, _p]),
    "dyd_dedup_workspace_bytes": (_sz, [_i64]),
    "dyd_dedup": (_int, [_p, _p, _i64, _int, _p, _p, _p, _sz, _p]),
    "dyd_dedup_ids": (_int, [_p, _p, _i64, _int, _p, _p, _p, _sz, _p]),
    "dyd_shard_bucket": (_int, [_p, _p, _i64, _i64, _i32, _i64, _p, _p, _p, _p]),
    "dyd_dedup_records": (_int, [_p, _i64, _int, _p, _p, _p, _sz, _p]),
    "dyd_shard_bucket_p2p": (_int, [_p, _p, _i64, _i64, _i32, _i32, _i64, _p, _p, _p, _p, _p]),
    "dyd_shard_unpack_p2p": (_int, [_p, _p, _p, _i32, _i64, _i64, _p, _p, _i32, _p]),
    "dyd_shard_pack_reply_p2p": (_int, [_p, _p, _p, _i64, _i64, _i32, _p, _i32, _i32, _p]),
    "dyd_shard_bucket_p2p_defaults": (_int, [_p, _p, _i64, _i64, _i32, _i32, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "dyd_shard_pack_reply2_p2p": (_int, [_p, _p, _p, _p, _p, _i64, _i64, _i32, _p, _i32, _i32, _p]),
    "dyd_shard_unpack2_p2p": (_int, [_p, _p, _p, _i32, _i64, _i64, _p, _p, _p, _p, _i32, _p]),
    "dyd_shard_pack_reply": (_int, [_p, _p, _p, _i64, _p, _i32, _p]),
    "dyd_shard_unpack": (_int, [_p, _i64, _i64, _i64, _p, _p, _i32, _p]),
    "dyd_url_filter_workspace_bytes": (_sz, [_i64, _i64]),
    "dyd_url_filter": (_int, [_p, _p, _i64, _p, _p, _i64, _int, _p, _p, _p, _p, _p, _sz, _p]),
    "dyd_url_filter_records": (_int, [_p, _i64, _p, _i64, _int, _p, _p, _p, _p, _p, _sz, _i32, _i64, _p]),
    "dyd_antijoin_records": (_int, [_p, _i64, _p, _i64, _p, _p, _p, _sz, _i32, _p]),
    "dyd_bbox_iou_fused_ex": (_int, [_p, _p, _p, _i64, _i64, _i64, _f64, _p, _p, _p, _p, _p, _p, _sz, _i32, _p, _p]),
    "dyd_fused_cta_times": (_int, [_p, _i32]),
    "dyd_fused_tile_modes": (_int, [_p, _i64, _p, _p]),
    "dyd_bbox_iou_host_ex": (_int, [_p, _p, _p, _i64, _i64, _f64, _p, _p, _p, _p, _p, _i64, _p]),
    "dyd_antijoin_workspace_bytes": (_sz, [_i64]),
    "dyd_antijoin": (_int, [_p, _p, _i64, _p, _p, _i64, _p, _p, _p, _sz, _p]),
    "dyd_label_lut": (_int, [_p, _p, _i64, _i64, _p, _p, _p, _i32, _p, _p, _p, _p]),
    "dyd_label_hist": (_int, [_p, _i64, _i32, _p, _p]),
    "dyd_label_presence": (_int, [_p, _p, _i64, _i32, _p, _p, _p]),
    "dyd_split_workspace_bytes": (_sz, [_i64, _i32]),
    "dyd_split_count": (_int, [_p, _i64, _p, _p, _i32, _i32, _p, _p, _sz, _p]),
    "dyd_split_fill": (_int, [_p, _i64, _p, _p, _i32, _i32, _p, _p, _p, _p, _p, _sz, _p]),
    "dyd_split_assign": (_int, [_p, _i32, _p, _i64, _p, _p, _p, _p, _p]),
    "dyd_split_assign_range": (_int, [_p, _i32, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p]),
    "dyd_numpy_permutation": (_int, [C.c_uint32, _i64, _p]),
    "dyd_yolo_normalise": (_int, [_p, _p, _p, _p, _i64, _i64, _p, _p, _p]),
    "dyd_bbox_iou_host": (_int, [_p, _p, _p, _i64, _i64, _f64, _p, _p, _p, _p, _p, _i64]),
    "dyd_dedup_host": (_int, [_p, _p, _p, _i64, _int, _p, _p]),
    "dyd_antijoin_host": (_int, [_p, _p, _p, _i64, _p, _p, _p, _i64, _p, _p]),
    "dyd_host_release": (_int, []),
    "dyd_ingest_cells": (_int, [_p, _p, _p, _i64, _int, _int, _p]),
    "dyd_ingest_free": (None, [_p]),
    "dyd_ingest_sizes": (_int, [_p, _p, _p, _p]),
    "dyd_ingest_export_polygons": (_int, [_p, _p, _p, _p, _p, _p, _p, _p, _int]),
    "dyd_ingest_export_boxes": (_int, [_p, _p, _p, _p, _p, _int]),
    "dyd_ingest_effective_text": (_int, [_p, _p, _p, _p, _p, _p, _int]),
    "dyd_json_canonical": (_i64, [_p, _i64, _p, _i64]),
    "dyd_ingest_export_names": (_int, [_p, _p, _p, _p, _p, _int]),
    "dyd_ingest_export_objects": (_int, [_p, _p, _p, _p, _int]),
    "dyd_egress_split": (_int, [_p, _p, _p, _i64, _p, _p, _p, _p, _p, _i64, _p, _p, _int]),
    "dyd_egress_names": (_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _p, _p, _int]),
    "dyd_egress_ptlist": (_int, [_p, _p, _p, _p, _p, _p, _p, _int]),
    "dyd_csv_write": (_int, [_p, _p, _p, _p, _i32, _i64, _p, _p, _int]),
    "dyd_py_float_repr": (_int, [_f64, C.c_char_p]),
    "dyd_yolo_format": (_int, [_p, _p, _p, _p, _i64, _p, _p, _int]),
    "dyd_csv_open": (_int, [_p, _i64, _p, _p, _i32, _i32, _p]),
    "dyd_csv_info": (_int, [_p, _p, _p, _p, _p, _p]),
    "dyd_csv_measure": (_int, [_p, _i64, _p, _p, _p, _p, _i32]),
    "dyd_csv_fill": (_int, [_p, _i32, _p, _p, _p, _p, _i32]),
    "dyd_csv_close": (None, [_p]),
    "dyd_read_file": (_int, [C.c_char_p, _p, _i64, _i32]),
    "dyd_csv_roundtrip_check": (_int, [_p, _p, _p, _i64, _i64, _p, _p, _i32, _i32, _i32]),
    "dyd_csv_write_file": (_int, [C.c_char_p, _i32, _p, _i64, _p, _p, _p, _p, _i32, _p, _i64, _int, _p]),
}

# libdyd_synth.so (include/dyd_synth.h): the synthetic-table generator of bench.py / the GPU tests, not product code
SYNTH_PROTOTYPES = {
    "dyd_synth_counts": (_int, [_u64, _i64, _i64, _p, _i32, _p, _p]),
    "dyd_synth_nvert": (_int, [_u64, _i64, _i64, _p, _p, _p]),
    "dyd_synth_fill": (_int, [_u64, _i64, _i64, _p, _p, _p, _p, _p]),
    "dyd_synth_urls": (_int, [_u64, _i64, _i64, _i64, _p, _p, _p]),
    "dyd_synth_url_bytes": (_int, [_p, _p, _i64, _p, _p]),
    "dyd_synth_crowd": (_int, [_u64, _i64, _i64, _i32, _i32, _p, _p, _p, _p]),
}
SYNTH_LIB_PATH = PKG / "libdyd_synth.so"

_lock = threading.Lock()
_lib = None
_synth = None


class DydError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libdyd.so (building nothing: run ``python -m deal_yolo_daya_b200.build`` first)."""
    global _lib
    with _lock:
        if _lib is None:
            if not LIB_PATH.exists():
                raise DydError(
                    f"{LIB_PATH} is missing: the CUDA hot path is not built. "
                    "Run `python -m deal_yolo_daya_b200.build` (needs nvcc); there is no CPU fallback.")
            lib = C.CDLL(str(LIB_PATH), use_errno=True)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(lib, name)          # AttributeError = header / library out of sync
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def load_synth() -> C.CDLL:
    """Load libdyd_synth.so (the generator used by bench.py and the tests)."""
    global _synth
    with _lock:
        if _synth is None:
            if not SYNTH_LIB_PATH.exists():
                raise DydError(f"{SYNTH_LIB_PATH} is missing: run `python -m deal_yolo_daya_b200.build`")
            lib = C.CDLL(str(SYNTH_LIB_PATH))
            for name, (res, args) in SYNTH_PROTOTYPES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _synth = lib
    return _synth


def last_error() -> str:
    buf = C.create_string_buffer(512)
    load().dyd_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        kind = "invalid argument" if rc < 0 else "CUDA error"
        raise DydError(f"{what} failed ({kind} {rc}): {last_error()}")
