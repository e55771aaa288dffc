"""Python side of the native (C++, multi-threaded) ingest / egress in csrc/ingest.cpp.

Cells go to the library as one UTF-8 buffer + int64 offsets (Arrow ``large_string`` layout, taken
zero-copy from pyarrow when the column allows it).  The parser returns the CSR buffers for the
kernels and, for step 4, splices the output cells natively from the original number literals.
Rows it is not certain about come back flagged SLOW; the caller runs those through the CPython
path in ``ingest.py`` / ``processor.py`` (json.loads, exact reference semantics).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib

ROW_OK, ROW_SLOW, ROW_NOT_TEXT = 0, 1, 2
K_INT, K_FLT, K_TRUE, K_FALSE, K_NULL = 3, 4, 5, 6, 7


def enabled() -> bool:
    return os.environ.get("DYD_NATIVE_INGEST", "1") != "0"


def _threads() -> int:
    return int(os.environ.get("DYD_INGEST_THREADS", "0"))


def pack_cells(cells):
    """list / Series of str-or-missing -> (text uint8[], off int64[n+1], is_text uint8[n])."""
    n = len(cells)
    try:
        import pyarrow as pa
        arr = pa.array(cells, type=pa.large_string(), from_pandas=True)
        if arr.offset != 0:
            arr = pa.concat_arrays([arr])
        bufs = arr.buffers()
        off = np.frombuffer(bufs[1], dtype=np.int64, count=n + 1)
        data = np.frombuffer(bufs[2], dtype=np.uint8) if bufs[2] is not None else np.zeros(0, np.uint8)
        is_text = np.ones(n, np.uint8) if arr.null_count == 0 else np.asarray(arr.is_valid()).astype(np.uint8)
        return data, off, is_text, arr          # keep `arr` alive: off/data alias its buffers
    except Exception:  # noqa: BLE001 - mixed-type column: encode by hand
        pass
    is_text = np.zeros(n, np.uint8)
    off = np.zeros(n + 1, np.int64)
    parts = []
    pos = 0
    for i, c in enumerate(cells):
        if isinstance(c, str):
            b = c.encode("utf-8", "surrogatepass")
            parts.append(b); pos += len(b); is_text[i] = 1
        off[i + 1] = pos
    data = np.frombuffer(b"".join(parts), dtype=np.uint8) if pos else np.zeros(0, np.uint8)
    return data, off, is_text, parts


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Ingest:
    """Owns a dyd_ingest handle (parse results) plus the packed text it refers to."""

    def __init__(self, cells, mode: int):
        self.lib = _lib.load()
        self.n = len(cells)
        self.mode = mode
        self.text, self.off, self.is_text, self._keep = pack_cells(cells)
        h = C.c_void_p()
        _lib.check(self.lib.dyd_ingest_cells(_p(self.text), _p(self.off), _p(self.is_text), self.n, mode, _threads(), C.byref(h)),
                   "dyd_ingest_cells")
        self.h = h
        no, nv, ns = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(self.lib.dyd_ingest_sizes(self.h, C.byref(no), C.byref(nv), C.byref(ns)), "dyd_ingest_sizes")
        self.n_obj, self.n_vert, self.n_slow = no.value, nv.value, ns.value

    def close(self):
        if getattr(self, "h", None):
            self.lib.dyd_ingest_free(self.h)
            self.h = None

    __del__ = close

    # ---- step 4 ----
    def polygons(self):
        n = self.n
        self.status = np.empty(n, np.uint8)
        self.img_off = np.empty(n + 1, np.int64)
        self.poly_off = np.empty(self.n_obj + 1, np.int64)
        self.xy = np.empty(2 * self.n_vert, np.float64)
        self.wh_off = np.empty(2 * n, np.int64); self.wh_len = np.empty(2 * n, np.int32); self.wh_kind = np.empty(2 * n, np.uint8)
        _lib.check(self.lib.dyd_ingest_export_polygons(self.h, _p(self.status), _p(self.img_off), _p(self.poly_off), _p(self.xy),
                                                       _p(self.wh_off), _p(self.wh_len), _p(self.wh_kind), _threads()),
                   "dyd_ingest_export_polygons")
        return self

    def egress_ptlist(self, arg, valid):
        """-> (out bytes uint8[], out_off int64[n+1]) : the new cell texts of the ROW_OK rows."""
        arg = np.ascontiguousarray(arg, np.int32); valid = np.ascontiguousarray(valid, np.uint8)
        out_off = np.empty(self.n + 1, np.int64)
        a = (self.h, _p(self.text), _p(self.off), _p(arg), _p(valid), _p(out_off))
        _lib.check(self.lib.dyd_egress_ptlist(*a, None, _threads()), "dyd_egress_ptlist(size)")
        out = np.empty(int(out_off[-1]), np.uint8)
        _lib.check(self.lib.dyd_egress_ptlist(*a, _p(out), _threads()), "dyd_egress_ptlist(write)")
        return out, out_off

    def scalar(self, r: int, which: int):
        """width (which=0) / height (which=1) of row r as the Python object json.loads would give."""
        i = 2 * r + which
        o = int(self.wh_off[i])
        if o < 0:
            return None
        k = int(self.wh_kind[i])
        if k == K_NULL:
            return None
        if k == K_TRUE:
            return True
        if k == K_FALSE:
            return False
        a = int(self.off[r]) + o
        lit = bytes(self.text[a:a + int(self.wh_len[i])]).decode("ascii")
        return int(lit) if k == K_INT else float(lit)

    # ---- step 5 ----
    def boxes(self):
        n = self.n
        self.status = np.empty(n, np.uint8)
        self.img_off = np.empty(n + 1, np.int64)
        self.pts = np.empty(4 * self.n_obj, np.float64)
        self.valid = np.empty(self.n_obj, np.uint8)
        _lib.check(self.lib.dyd_ingest_export_boxes(self.h, _p(self.status), _p(self.img_off), _p(self.pts), _p(self.valid), _threads()),
                   "dyd_ingest_export_boxes")
        return self


# ------------------------------------------------------------------------------------------------
# CSV egress: DataFrame.to_csv(path, index=False, encoding=...) with the body written natively
# ------------------------------------------------------------------------------------------------
def _csv_column(col):
    """-> (kind, off, data, valid, keepalive) or None when the column needs pandas' own formatting."""
    import pandas as pd
    import pyarrow as pa
    dt = col.dtype
    if dt == np.float64:
        a = np.ascontiguousarray(col.to_numpy())
        return 1, None, a.view(np.uint8), None, a
    if dt == np.int64:
        a = np.ascontiguousarray(col.to_numpy())
        return 2, None, a.view(np.uint8), None, a
    if dt == np.bool_:
        a = np.ascontiguousarray(col.to_numpy()).astype(np.uint8)
        return 3, None, a, None, a
    if isinstance(dt, pd.StringDtype) or dt == object:
        try:
            arr = pa.array(col, type=pa.large_string(), from_pandas=True)      # raises on non-str objects
        except Exception:  # noqa: BLE001
            return None
        if arr.offset != 0:
            arr = pa.concat_arrays([arr])
        bufs = arr.buffers()
        n = len(arr)
        off = np.frombuffer(bufs[1], dtype=np.int64, count=n + 1)
        data = np.frombuffer(bufs[2], dtype=np.uint8) if bufs[2] is not None and bufs[2].size else np.zeros(1, np.uint8)
        valid = None if arr.null_count == 0 else np.asarray(arr.is_valid()).astype(np.uint8)
        return 0, off, data, valid, arr
    return None


def to_csv(df, path, encoding="utf-8-sig") -> None:
    """``df.to_csv(path, index=False, encoding=encoding)``, byte-identical, body rows written by
    csrc/ingest.cpp (multi-threaded).  Falls back to pandas for frames it does not cover."""
    import csv
    import io
    enc = (encoding or "utf-8").lower().replace("_", "-")
    cols = None
    if enabled() and df.shape[1] >= 2 and enc in ("utf-8", "utf-8-sig", "utf8") and df.columns.is_unique:
        cols = [_csv_column(df[c]) for c in df.columns]
        if any(c is None for c in cols):
            cols = None
    if cols is None:
        df.to_csv(path, index=False, encoding=encoding)
        return
    lib = _lib.load()
    n, nc = len(df), len(cols)
    kinds = (C.c_int32 * nc)(*[c[0] for c in cols])
    offs = (C.c_void_p * nc)(*[c[1].ctypes.data if c[1] is not None else None for c in cols])
    datas = (C.c_void_p * nc)(*[c[2].ctypes.data for c in cols])
    valids = (C.c_void_p * nc)(*[c[3].ctypes.data if c[3] is not None else None for c in cols])
    row_off = np.empty(n + 1, np.int64)
    a = (kinds, offs, datas, valids, nc, n, _p(row_off))
    _lib.check(lib.dyd_csv_write(*a, None, _threads()), "dyd_csv_write(size)")
    body = np.empty(int(row_off[-1]), np.uint8)
    _lib.check(lib.dyd_csv_write(*a, _p(body), _threads()), "dyd_csv_write(write)")
    head = io.StringIO()
    csv.writer(head, lineterminator="\n", quoting=csv.QUOTE_MINIMAL).writerow([str(c) for c in df.columns])
    with open(path, "wb") as f:
        if enc == "utf-8-sig":
            f.write(b"\xef\xbb\xbf")
        f.write(head.getvalue().encode("utf-8"))
        f.write(body.tobytes() if body.size < (1 << 20) else memoryview(body))
