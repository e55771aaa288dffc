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

ROW_OK, ROW_SLOW, ROW_NOT_TEXT, ROW_NO_LIST, ROW_NOT_A_LIST = 0, 1, 2, 4, 5
K_INT, K_FLT, K_TRUE, K_FALSE, K_NULL = 3, 4, 5, 6, 7


def enabled() -> bool:
    return os.environ.get("DYD_NATIVE_INGEST", "1") != "0"


def _threads() -> int:
    return int(os.environ.get("DYD_INGEST_THREADS", "0"))


def pack_cells(cells):
    """list / Series of str-or-missing -> (text uint8[], off int64[n+1], is_text uint8[n]).
    An Arrow-backed pandas column is taken as it is (no Python string is created)."""
    n = len(cells)
    try:
        import pyarrow as pa
        pa_arr = getattr(getattr(cells, "array", None), "_pa_array", None)      # pandas str column (pyarrow storage)
        if pa_arr is not None and str(pa_arr.type) in ("large_string", "string"):
            arr = pa_arr.combine_chunks() if pa_arr.num_chunks != 1 else pa_arr.chunk(0)
            if str(arr.type) != "large_string":
                arr = arr.cast(pa.large_string())
        else:
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
        self.arrow_input = not isinstance(self._keep, list)          # valid UTF-8 by construction
        h = C.c_void_p()
        _lib.check(self.lib.dyd_ingest_cells(_p(self.text), _p(self.off), _p(self.is_text), self.n, mode, _threads(), C.byref(h)),
                   "dyd_ingest_cells")
        self.h = h
        no, nv, ns = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(self.lib.dyd_ingest_sizes(self.h, C.byref(no), C.byref(nv), C.byref(ns)), "dyd_ingest_sizes")
        self.n_obj, self.n_vert, self.n_slow = no.value, nv.value, ns.value
        # cells that were valid JSON in another style were rewritten as json.dumps would write them and parsed
        # from that form: every later call must see the rewritten texts
        nc = C.c_int64()
        _lib.check(self.lib.dyd_ingest_effective_text(self.h, None, None, C.byref(nc), None, None, 0), "dyd_ingest_effective_text(count)")
        self.n_canon = nc.value
        if self.n_canon:
            new_off = np.empty(self.n + 1, np.int64)
            a = (self.h, _p(self.text), _p(self.off), None, _p(new_off))
            _lib.check(self.lib.dyd_ingest_effective_text(*a, None, _threads()), "dyd_ingest_effective_text(size)")
            new_text = np.empty(max(int(new_off[-1]), 1), np.uint8)
            _lib.check(self.lib.dyd_ingest_effective_text(*a, _p(new_text), _threads()), "dyd_ingest_effective_text(write)")
            self.text, self.off = new_text, new_off

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

    # ---- step 5.5 ----
    def names(self):
        n = self.n
        self.status = np.empty(n, np.uint8)
        self.cell_off = np.empty(n + 1, np.int64)
        self.name_off = np.empty(self.n_obj, np.int64)
        self.name_len = np.empty(self.n_obj, np.int32)
        _lib.check(self.lib.dyd_ingest_export_names(self.h, _p(self.status), _p(self.cell_off), _p(self.name_off), _p(self.name_len),
                                                    _threads()), "dyd_ingest_export_names")
        return self

    def name_array(self):
        """Arrow large_string array of the objects' names (null where the object has none)."""
        import pyarrow as pa
        cnt = np.diff(self.cell_off)
        start = np.repeat(self.off[:-1], cnt) + self.name_off                  # absolute byte offset of each name
        ln = np.where(self.name_len < 0, 0, self.name_len).astype(np.int64)
        o = np.zeros(self.n_obj + 1, np.int64); np.cumsum(ln, out=o[1:])
        total = int(o[-1])
        if total:
            idx = np.repeat(start - o[:-1], ln) + np.arange(total, dtype=np.int64)
            data = self.text[idx]
        else:
            data = np.zeros(1, np.uint8)
        valid = np.packbits(self.name_len >= 0, bitorder="little")
        return pa.Array.from_buffers(pa.large_string(), self.n_obj, [pa.py_buffer(valid), pa.py_buffer(o), pa.py_buffer(data)])

    def objects(self):
        """list_len int32[n] (elements of "objects"), obj_off int64 / obj_len int32 [n_obj] (dict-object spans)."""
        self.list_len = np.empty(self.n, np.int32)
        self.obj_off = np.empty(self.n_obj, np.int64)
        self.obj_len = np.empty(self.n_obj, np.int32)
        _lib.check(self.lib.dyd_ingest_export_objects(self.h, _p(self.list_len), _p(self.obj_off), _p(self.obj_len), _threads()),
                   "dyd_ingest_export_objects")
        return self

    def egress_split(self, exp_cell, exp_obj, exp_tok, tok_bytes, tok_off):
        """-> (out bytes uint8[], out_off int64[n_exp+1]): the one-object cells of the split's expanded rows."""
        exp_cell = np.ascontiguousarray(exp_cell, np.int64); exp_obj = np.ascontiguousarray(exp_obj, np.int64)
        exp_tok = np.ascontiguousarray(exp_tok, np.int32)
        tok_bytes = np.ascontiguousarray(tok_bytes, np.uint8); tok_off = np.ascontiguousarray(tok_off, np.int64)
        n_exp = len(exp_cell)
        out_off = np.empty(n_exp + 1, np.int64)
        a = (self.h, _p(self.text), _p(self.off), n_exp, _p(exp_cell), _p(exp_obj), _p(exp_tok), _p(tok_bytes), _p(tok_off), len(tok_off) - 1, _p(out_off))
        _lib.check(self.lib.dyd_egress_split(*a, None, _threads()), "dyd_egress_split(size)")
        out = np.empty(max(int(out_off[-1]), 1), np.uint8)
        _lib.check(self.lib.dyd_egress_split(*a, _p(out), _threads()), "dyd_egress_split(write)")
        return out[:int(out_off[-1])], out_off

    def egress_names(self, obj_flag, obj_new, vocab_bytes, vocab_off):
        """-> (out bytes uint8[], out_off int64[n+1]): new texts of the ROW_OK cells (length 0 for the others)."""
        obj_flag = np.ascontiguousarray(obj_flag, np.uint8); obj_new = np.ascontiguousarray(obj_new, np.int32)
        vocab_bytes = np.ascontiguousarray(vocab_bytes, np.uint8); vocab_off = np.ascontiguousarray(vocab_off, np.int64)
        out_off = np.empty(self.n + 1, np.int64)
        a = (self.h, _p(self.text), _p(self.off), _p(obj_flag), _p(obj_new), _p(vocab_bytes), _p(vocab_off), len(vocab_off) - 1, _p(out_off))
        _lib.check(self.lib.dyd_egress_names(*a, None, _threads()), "dyd_egress_names(size)")
        out = np.empty(max(int(out_off[-1]), 1), np.uint8)
        _lib.check(self.lib.dyd_egress_names(*a, _p(out), _threads()), "dyd_egress_names(write)")
        return out[:int(out_off[-1])], out_off

    def int_columns(self):
        """(width int64[n], height int64[n]) when every row has both as plain JSON integers that fit
        int64 (what pandas turns a list of Python ints into), else None."""
        if self.n == 0 or not (self.wh_kind == K_INT).all() or (self.wh_off < 0).any() or (self.wh_len > 18).any():
            return None
        start = np.repeat(self.off[:-1], 2) + self.wh_off
        length = self.wh_len.astype(np.int64)
        out = np.zeros(2 * self.n, np.int64)
        for L in np.unique(length):                     # a handful of distinct literal lengths
            idx = np.nonzero(length == L)[0]
            chars = self.text[start[idx][:, None] + np.arange(L)[None, :]].astype(np.int64)
            neg = chars[:, 0] == ord("-")
            digits = chars - ord("0")
            digits[neg, 0] = 0
            if ((digits < 0) | (digits > 9)).any():
                return None
            val = digits @ (10 ** np.arange(L - 1, -1, -1, dtype=np.int64))
            out[idx] = np.where(neg, -val, val)
        return out[0::2].copy(), out[1::2].copy()

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


def arrow_strings(data: np.ndarray, off: np.ndarray):
    """(bytes, offsets int64[n+1]) -> pandas string array over the same buffers (the dtype pd.read_csv
    infers for text: ``str`` where pandas has it, else a list of Python strings)."""
    import pandas as pd
    n = len(off) - 1
    if not _pandas_infers_arrow_str():
        blob = data.tobytes()
        return [blob[off[r]:off[r + 1]].decode("utf-8", "surrogatepass") for r in range(n)]
    import pyarrow as pa
    from pandas.arrays import ArrowStringArray
    arr = pa.Array.from_buffers(pa.large_string(), n, [None, pa.py_buffer(np.ascontiguousarray(off, np.int64)),
                                                        pa.py_buffer(data if data.size else np.zeros(1, np.uint8))], null_count=0)
    return ArrowStringArray(pa.chunked_array([arr]), dtype=pd.StringDtype(storage="pyarrow", na_value=np.nan))


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
        pa_arr = getattr(getattr(col, "array", None), "_pa_array", None)          # Arrow-backed column: its own buffers
        try:
            if pa_arr is not None and str(pa_arr.type) == "large_string":
                arr = pa_arr.chunk(0) if pa_arr.num_chunks == 1 else pa_arr.combine_chunks()
            else:
                arr = pa.array(col, type=pa.large_string(), from_pandas=True)  # raises on non-str objects
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


def to_csv(df, path, encoding="utf-8-sig", mode="w", header=True, rows=None) -> bool:
    """``df.to_csv(path, index=False, encoding=encoding, mode=mode, header=header)``, byte-identical, written by
    csrc/ingest.cpp straight into the file (worker threads format blocks of rows, the calling thread writes them in
    order).  ``rows`` (int64 row positions, in output order) writes ``df.iloc[rows]`` without building that frame.
    Falls back to pandas for frames the native writer does not cover; returns True when the native writer wrote the
    file.  In append mode no BOM is written (Python's TextIOWrapper skips it when the file position is not 0, which is
    what pandas relies on)."""
    import csv
    import io
    enc = (encoding or "utf-8").lower().replace("_", "-")
    cols = None
    if enabled() and df.shape[1] >= 2 and enc in ("utf-8", "utf-8-sig", "utf8") and df.columns.is_unique and isinstance(path, (str, os.PathLike)):
        cols = [_csv_column(df[c]) for c in df.columns]
        if any(c is None for c in cols):
            cols = None
    if cols is None:
        (df if rows is None else df.iloc[rows]).to_csv(path, index=False, encoding=encoding, mode=mode, header=header)
        return False
    lib = _lib.load()
    nc = len(cols)
    kinds = (C.c_int32 * nc)(*[c[0] for c in cols])
    offs = (C.c_void_p * nc)(*[c[1].ctypes.data if c[1] is not None else None for c in cols])
    datas = (C.c_void_p * nc)(*[c[2].ctypes.data for c in cols])
    valids = (C.c_void_p * nc)(*[c[3].ctypes.data if c[3] is not None else None for c in cols])
    if rows is not None:
        rows = np.ascontiguousarray(rows, np.int64)
        if rows.size and (int(rows.min()) < 0 or int(rows.max()) >= len(df)):
            raise IndexError("to_csv: row position out of range")
    n_sel = len(df) if rows is None else int(rows.size)
    append = mode == "a"
    prefix = b""
    if enc == "utf-8-sig" and not (append and os.path.exists(path) and os.path.getsize(path) != 0):
        prefix += b"\xef\xbb\xbf"
    if header:
        head = io.StringIO()
        csv.writer(head, lineterminator="\n", quoting=csv.QUOTE_MINIMAL).writerow([str(c) for c in df.columns])
        prefix += head.getvalue().encode("utf-8")
    if mode not in ("w", "a"):
        raise ValueError(f"to_csv: unsupported mode {mode!r}")
    written = C.c_int64()
    rc = lib.dyd_csv_write_file(os.fsencode(path), 1 if append else 0, prefix, len(prefix), kinds, offs, datas, valids, nc,
                                _p(rows) if rows is not None else None, n_sel, _threads(), C.byref(written))
    if rc == -4:                                   # DYD_E_IO: the same exception open() / write() would have raised
        err = C.get_errno()
        raise OSError(err, os.strerror(err) if err else "write failed", os.fspath(path))
    _lib.check(rc, "dyd_csv_write_file")
    return True


def roundtrip_safe(df, check_cells=True) -> bool:
    """True when ``pd.read_csv(file, encoding="utf-8-sig")`` of the file ``to_csv(df, file)`` writes returns `df` again
    (RangeIndex assumed): what tablecache needs to know before it hands the written frame to the next step."""
    import pandas as pd
    n, nc = df.shape
    if n == 0 or nc < 2 or not df.columns.is_unique or not _pandas_infers_arrow_str():
        return False
    for name in df.columns:
        if not isinstance(name, str) or not name or name != name.strip() or any(ch in name for ch in '\r\n\0'):
            return False
    lib = _lib.load()
    na_bytes, na_off, n_na = _na_table()
    window = _buffer_lines(nc)
    for name in df.columns:
        col = df[name]
        dt = col.dtype
        if dt == np.float64 or dt == np.int64 or dt == np.bool_:
            continue
        if not (isinstance(dt, pd.StringDtype) and dt.storage == "pyarrow" and str(dt) == "str"):
            return False
        got = _csv_column(col)
        if got is None or got[0] != 0:
            return False
        _, off, data, valid, _keep = got
        rc = lib.dyd_csv_roundtrip_check(_p(off), _p(data), _p(valid), n, window, _p(na_bytes), _p(na_off), n_na,
                                         1 if check_cells else 0, _threads())
        if rc != 1:
            if rc < 0:
                _lib.check(rc, "dyd_csv_roundtrip_check")
            return False
    return True


def permutation(seed, n: int) -> np.ndarray:
    """``np.random.RandomState(seed).permutation(n)`` (the order ``DataFrame.sample(frac=1, random_state=seed)`` gives the rows,
    processor.py:800), bit for bit, from csrc/np_perm.cpp; seeds numpy treats differently (not an int in 0 .. 2^32-1) and small
    n go to numpy itself."""
    if n < 4096 or n > 0xFFFFFFFF or not enabled() or isinstance(seed, bool) or not isinstance(seed, (int, np.integer)) or not 0 <= int(seed) <= 0xFFFFFFFF:
        return np.random.RandomState(seed).permutation(n)
    out = np.empty(n, np.int64)
    _lib.check(_lib.load().dyd_numpy_permutation(int(seed), n, _p(out)), "dyd_numpy_permutation")
    return out


def yolo_label_texts(img_off, class_id, cxcywh, ok):
    """Label-file texts of processor.py:1045-1052 for every image: (text uint8[], off int64[n_img+1]).
    Inputs are host arrays: img_off int64[n_img+1], class_id int32[n_box], cxcywh float64[4*n_box]
    (dyd_yolo_normalise's output), ok uint8[n_box]."""
    img_off = np.ascontiguousarray(img_off, np.int64); class_id = np.ascontiguousarray(class_id, np.int32)
    cxcywh = np.ascontiguousarray(cxcywh, np.float64).reshape(-1); ok = np.ascontiguousarray(ok, np.uint8)
    n_img = len(img_off) - 1
    lib = _lib.load()
    off = np.empty(n_img + 1, np.int64)
    a = (_p(img_off), _p(class_id), _p(cxcywh), _p(ok), n_img, _p(off))
    _lib.check(lib.dyd_yolo_format(*a, None, _threads()), "dyd_yolo_format(size)")
    out = np.empty(max(int(off[-1]), 1), np.uint8)
    _lib.check(lib.dyd_yolo_format(*a, _p(out), _threads()), "dyd_yolo_format(write)")
    return out[:int(off[-1])], off


# ------------------------------------------------------------------------------------------------
# CSV ingest: pd.read_csv(path, encoding="utf-8[-sig]") with the text columns tokenised natively
# ------------------------------------------------------------------------------------------------
_READ_STATS = {"native": 0, "pandas": 0, "delegated_columns": 0, "wide": 0}
_STR_PROBE = None


def _pandas_infers_arrow_str() -> bool:
    """True when this pandas gives text columns the pyarrow-backed ``str`` dtype (pandas >= 3): only
    then can Arrow buffers become the column without creating one Python object per cell."""
    global _STR_PROBE
    if _STR_PROBE is None:
        import io
        import pandas as pd
        try:
            a = pd.read_csv(io.StringIO("a\nx\n"))["a"].array
            _STR_PROBE = type(a).__name__ == "ArrowStringArray" and str(a.dtype) == "str" and \
                str(a._pa_array.type) == "large_string"
        except Exception:  # noqa: BLE001
            _STR_PROBE = False
    return _STR_PROBE


def _buffer_lines(n_cols: int) -> int:
    """Rows per dtype-inference chunk of pandas' low-memory C reader (parsers.pyx, TextReader.__cinit__)."""
    heuristic = 2 ** 20 // max(n_cols, 1)
    b = 1
    while b * 2 < heuristic:
        b *= 2
    return b


def _na_table():
    from pandas._libs.parsers import STR_NA_VALUES
    vals = sorted(v.encode("utf-8") for v in STR_NA_VALUES)
    off = np.zeros(len(vals) + 1, np.int64)
    off[1:] = np.cumsum([len(v) for v in vals])
    return np.frombuffer(b"".join(vals) or b"\0", dtype=np.uint8), off, len(vals)


class NotNative(Exception):
    """read_csv(..., _strict_native=True): the input is outside what the native reader restates."""


def read_csv(path, encoding="utf-8", _strict_native=False, _window_rows=None, **kwargs):
    """``pd.read_csv(path, encoding=encoding, **kwargs)`` -- same frame, text columns tokenised by
    csrc/csv_read.cpp.  Anything the native reader does not cover (other encodings or keyword
    arguments, an unusual dialect, ragged rows, pandas without the Arrow ``str`` dtype, small files)
    is read by pandas itself -- or, with ``_strict_native``, reported with NotNative so that the caller
    can apply its own reading rules (the merge step reads with ``errors="ignore"`` in chunks of ``_window_rows`` rows,
    each inferred on its own: a longer file is then measured per such chunk)."""
    import io
    import pandas as pd
    enc = (encoding or "utf-8").lower().replace("_", "-")
    extra = {k: v for k, v in kwargs.items() if not (k == "parse_dates" and v is False)}

    def fallback():
        if _strict_native:
            raise NotNative(str(path))
        _READ_STATS["pandas"] += 1
        return pd.read_csv(path, encoding=encoding, **kwargs)

    if not enabled() or extra or enc not in ("utf-8", "utf8", "utf-8-sig") or not isinstance(path, (str, os.PathLike)):
        return fallback()
    try:
        size = os.path.getsize(path)
    except OSError:
        return fallback()
    if size < int(os.environ.get("DYD_CSV_NATIVE_MIN_BYTES", str(1 << 20))) or not _pandas_infers_arrow_str():
        return fallback()
    import pyarrow as pa
    lib = _lib.load()
    data = np.empty(size, np.uint8)
    rc = lib.dyd_read_file(os.fsencode(path), _p(data), size, _threads())
    if rc == -4:
        return fallback()                      # unreadable / changed while reading: pandas raises what it raises
    _lib.check(rc, "dyd_read_file")
    na_bytes, na_off, n_na = _na_table()
    h = C.c_void_p()
    _lib.check(lib.dyd_csv_open(_p(data), data.size, _p(na_bytes), _p(na_off), n_na, _threads(), C.byref(h)), "dyd_csv_open")
    try:
        nr, nc, hb, he, fl = C.c_int64(), C.c_int32(), C.c_int64(), C.c_int64(), C.c_int32()
        _lib.check(lib.dyd_csv_info(h, C.byref(nr), C.byref(nc), C.byref(hb), C.byref(he), C.byref(fl)), "dyd_csv_info")
        n_rows, n_cols = nr.value, nc.value
        if fl.value & 1 or n_rows == 0:
            return fallback()
        try:                                   # header names: pandas' own rules (duplicates, unnamed, quoting)
            names = list(pd.read_csv(io.BytesIO(data[hb.value:he.value].tobytes() + b"\n"), nrows=0, encoding="utf-8").columns)
        except Exception:  # noqa: BLE001
            return fallback()
        if len(names) != n_cols:
            return fallback()
        window = _buffer_lines(n_cols)
        if _window_rows is not None and n_rows > _window_rows:
            if _window_rows > window:
                return fallback()              # the caller's chunks would be cut again by pandas' own buffer: not restated
            window = int(_window_rows)
        col_bytes = np.zeros(n_cols, np.int64); col_nulls = np.zeros(n_cols, np.int64)
        col_text = np.zeros(n_cols, np.uint8); col_utf8 = np.zeros(n_cols, np.uint8)
        _lib.check(lib.dyd_csv_measure(h, window, _p(col_bytes), _p(col_nulls), _p(col_text), _p(col_utf8), _threads()),
                   "dyd_csv_measure")
        if not col_utf8.all():
            return fallback()                  # pandas raises UnicodeDecodeError
        uncertain = [j for j in range(n_cols) if not col_text[j]]
        # pandas infers dtypes per chunk of `window` rows and concatenates the chunks; columns that are not
        # certainly text in every chunk are inferred by pandas itself, chunk by chunk, from the same tokens
        windows = [(a, min(a + window, n_rows)) for a in range(0, n_rows, window)]
        sel = list(range(n_cols))              # every column's cell texts, one parallel pass
        offs = [np.empty(n_rows + 1, np.int64) for _ in sel]
        datas = [np.empty(max(int(col_bytes[j]), 1), np.uint8) for j in sel]
        maps = [np.empty((n_rows + 7) // 8, np.uint8) if col_nulls[j] else None for j in sel]
        cols = (C.c_int32 * len(sel))(*sel)
        a_off = (C.c_void_p * len(sel))(*[o.ctypes.data for o in offs])
        a_dat = (C.c_void_p * len(sel))(*[d.ctypes.data for d in datas])
        a_map = (C.c_void_p * len(sel))(*[m.ctypes.data if m is not None else None for m in maps])
        _lib.check(lib.dyd_csv_fill(h, len(sel), cols, a_off, a_dat, a_map, _threads()), "dyd_csv_fill")
        columns = {}
        str_dtype = pd.StringDtype(storage="pyarrow", na_value=np.nan)
        from pandas.arrays import ArrowStringArray
        import pyarrow.compute as pc
        for j in sel:
            bufs = [pa.py_buffer(maps[j]) if maps[j] is not None else None, pa.py_buffer(offs[j]), pa.py_buffer(datas[j])]
            arr = pa.Array.from_buffers(pa.large_string(), n_rows, bufs, null_count=int(col_nulls[j]))
            if col_text[j]:
                columns[j] = ArrowStringArray(pa.chunked_array([arr]), dtype=str_dtype)
                continue
            # numbers, booleans, empty columns ...: pandas infers the dtype from the same tokens, handed over
            # as a header-less one-column CSV with every cell quoted (a missing cell is an empty one)
            ls = pa.large_string()
            lines = pc.binary_join_element_wise(pa.scalar('"', ls), pc.replace_substring(pc.fill_null(arr, pa.scalar("", ls)), '"', '""'),
                                                pa.scalar('"\n', ls), pa.scalar("", ls))
            o = np.frombuffer(lines.buffers()[1], dtype=np.int64)[lines.offset:lines.offset + len(lines) + 1]
            buf = lines.buffers()[2]
            parts = []
            for a, b in windows:                                   # the cells' lines of one chunk, back to back
                text = buf.slice(int(o[a]), int(o[b] - o[a])).to_pybytes()
                one = pd.read_csv(io.BytesIO(text), header=None, encoding="utf-8", skip_blank_lines=False)
                if one.shape != (b - a, 1):
                    return fallback()
                parts.append(one.iloc[:, 0])
            one = parts[0].to_frame() if len(parts) == 1 else pd.concat(parts, ignore_index=True).to_frame()
            columns[j] = one.iloc[:, 0].array
            _READ_STATS["delegated_columns"] += 1
    finally:
        lib.dyd_csv_close(h)
    df = pd.DataFrame({i: columns[i] for i in range(n_cols)}, copy=False)
    df.columns = names
    _READ_STATS["native"] += 1
    _READ_STATS["wide"] += 1 if fl.value & 4 else 0
    return df
