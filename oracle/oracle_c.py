"""ctypes front-end of oracle/dyd_oracle.c -- TEST INFRASTRUCTURE ONLY.

Same functions and return conventions as oracle/oracle_np.py, fast enough for
multi-million-row parity checks and for bench.py's ``cpu_baseline`` leg.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build_oracle

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build_oracle.build()))
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_threads(n: int) -> None:
    lib().orc_set_threads(C.c_int(n))


def bbox_fold(poly_off, xy, want_arg=True):
    poly_off = _c(poly_off, np.int64); xy = _c(xy, np.float64)
    n = len(poly_off) - 1
    pts = np.empty(4 * n, np.float64); valid = np.empty(n, np.uint8)
    arg = np.empty(4 * n, np.int32) if want_arg else None
    lib().orc_bbox(_p(poly_off), _p(xy), C.c_int64(n), _p(pts), _p(valid), _p(arg))
    return pts, valid, arg


def iou_filter(img_off, pts, valid, min_boxes, thr):
    img_off = _c(img_off, np.int64); pts = _c(pts, np.float64)
    valid = _c(valid, np.uint8) if valid is not None else None
    n = len(img_off) - 1
    high = np.empty(n, np.uint8); count = np.empty(n, np.int32)
    lib().orc_iou_filter(_p(img_off), _p(pts), _p(valid), C.c_int64(n), C.c_int64(int(min_boxes)),
                         C.c_double(float(thr)), _p(high), _p(count))
    return high, count


def hash_strings_buf(off, data):
    off = _c(off, np.int64); data = _c(data, np.uint8)
    n = len(off) - 1
    out = np.empty(n, np.uint64)
    lib().orc_hash_strings(_p(off), _p(data), C.c_int64(n), _p(out))
    return out


def pack_strings(strings):
    bs = [s.encode("utf-8") for s in strings]
    off = np.zeros(len(bs) + 1, np.int64)
    np.cumsum([len(b) for b in bs], out=off[1:])
    data = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return off, data


def hash_strings(strings):
    return hash_strings_buf(*pack_strings(strings))


_KEEP = {"first": 0, "last": 1, False: 2}


def dedup(keys, null, keep="first"):
    keys = _c(keys, np.uint64); null = _c(null, np.uint8)
    n = len(keys)
    km = np.empty(n, np.uint8); rep = np.empty(n, np.int64)
    rc = lib().orc_dedup(_p(keys), _p(null), C.c_int64(n), C.c_int(_KEEP[keep]), _p(km), _p(rep))
    assert rc == 0
    return km, rep


def antijoin(mk, mnull, rk, rnull):
    mk = _c(mk, np.uint64); mnull = _c(mnull, np.uint8)
    rk = _c(rk, np.uint64); rnull = _c(rnull, np.uint8)
    n = len(mk)
    km = np.empty(n, np.uint8); rr = np.empty(n, np.int64)
    rc = lib().orc_antijoin(_p(mk), _p(mnull), C.c_int64(n), _p(rk), _p(rnull), C.c_int64(len(rk)), _p(km), _p(rr))
    assert rc == 0
    return km, rr


def label_lut(img_off, label_id, lut_new, lut_ntok, lut_nrep):
    img_off = _c(img_off, np.int64); label_id = _c(label_id, np.int32)
    lut_new = _c(lut_new, np.int32); lut_ntok = _c(lut_ntok, np.int32); lut_nrep = _c(lut_nrep, np.int32)
    n_img = len(img_off) - 1
    new_id = np.empty(len(label_id), np.int32); row_rep = np.empty(n_img, np.uint8)
    cnt = np.zeros(6, np.uint64)
    lib().orc_label_lut(_p(img_off), _p(label_id), C.c_int64(n_img), _p(lut_new), _p(lut_ntok), _p(lut_nrep),
                        _p(new_id), _p(row_rep), _p(cnt))
    names = ["total_objects", "missing_name_objects", "total_labels", "replaced_labels",
             "replaced_objects", "replaced_rows"]
    return new_id, row_rep, {k: int(v) for k, v in zip(names, cnt)}


def split_expand(img_off, label_id, cat_of_label, n_cat):
    img_off = _c(img_off, np.int64); label_id = _c(label_id, np.int32); cat = _c(cat_of_label, np.int32)
    n_img = len(img_off) - 1
    cat_off = np.zeros(n_cat + 1, np.int64)
    lib().orc_split_expand(_p(img_off), _p(label_id), C.c_int64(n_img), _p(cat), C.c_int32(n_cat),
                           _p(cat_off), None, None, None)
    n = int(cat_off[-1])
    ei = np.empty(n, np.int64); eb = np.empty(n, np.int64); ec = np.empty(n, np.int32)
    lib().orc_split_expand(_p(img_off), _p(label_id), C.c_int64(n_img), _p(cat), C.c_int32(n_cat),
                           _p(cat_off), _p(ei), _p(eb), _p(ec))
    return ei, eb, ec, cat_off
