"""Columnar ingest: annotation JSON cells -> ragged CSR buffers.

Round-1 implementation: CPython ``json.loads`` per cell (a native multi-threaded tokenizer is
SURVEY §8f-1, next).  The walk over each document follows the reference's own accessors so
that the same inputs raise the same exceptions:

  * step 4 (``parse_polygons``): processor.py:262-278 -- ``data.get("objects", [])``, dict
    objects only, ``obj.get("polygon", {}).get("ptList", [])``, valid points = dicts that
    have both "x" and "y" (:253);
  * step 5 (``parse_boxes``): processor.py:341-366 -- two-point ptLists only, a raising
    object ends the row's scan (prefix kept).

Coordinates go to the GPU as fp64.  A value is *device representable* when it is an int /
float / bool whose double conversion is exact (|int| <= 2**53).  Anything else (None, str,
huge ints, nested containers) cannot be expressed in an fp64 array; such polygons / rows are
flagged for the host exception lane (``_hostlane.py``), which evaluates them with CPython
semantics, raising exactly where the reference raises.  Synthetic and real-world numeric
tables never take that lane; the drop-in reports how many objects did.
"""
from __future__ import annotations

import json
from array import array
from dataclasses import dataclass, field

import numpy as np

from . import _hostlane

_EXACT = 2 ** 53


def _dev_ok(v, int_bound=_EXACT) -> bool:
    t = type(v)
    if t is float:
        return True
    if t is int or t is bool:
        return -int_bound <= v <= int_bound
    return False


# Step 5 does arithmetic on the coordinates.  CPython's is exact for ints; fp64 gives the same quotient only while every
# difference, product and sum is exactly representable, which |int| <= 2^25 guarantees (products <= 2^52, their sum <= 2^53).
_IOU_INT = 2 ** 25


@dataclass
class PolygonBatch:
    """CSR of the dict objects of every parsable row (step 4 input)."""
    n_rows: int
    docs: list                      # parsed document per row, None where the cell is not decodable JSON
    objs: list                      # per row: list of the dict objects (reference order)
    points: list                    # per object: its valid point dicts (host-lane objects: their two corner points)
    hostlane: np.ndarray            # uint8 per object: 1 = evaluated on the host lane
    img_off: np.ndarray             # int64[n_rows+1]
    poly_off: np.ndarray            # int64[n_obj+1]
    xy: np.ndarray                  # float64[2*n_vert]

    @property
    def n_obj(self):
        return len(self.poly_off) - 1


def parse_polygons(cells) -> PolygonBatch:
    docs, objs_per_row, points = [], [], []
    img_cnt, poly_cnt, lane = array("q"), array("q"), array("B")
    xy = array("d")
    for text in cells:
        if not isinstance(text, str):
            docs.append(None); objs_per_row.append([]); img_cnt.append(0)
            continue
        try:
            doc = json.loads(text)
        except json.JSONDecodeError:
            docs.append(None); objs_per_row.append([]); img_cnt.append(0)
            continue
        row_objs = []
        for obj in doc.get("objects", []):               # AttributeError / TypeError propagate like the reference
            if not isinstance(obj, dict):
                continue
            ptlist = obj.get("polygon", {}).get("ptList", [])
            good = [p for p in ptlist if isinstance(p, dict) and "x" in p and "y" in p]
            on_dev = True
            for p in good:
                if not (_dev_ok(p["x"]) and _dev_ok(p["y"])):
                    on_dev = False
                    break
            if on_dev:
                for p in good:
                    xy.append(p["x"]); xy.append(p["y"])
                poly_cnt.append(len(good)); lane.append(0)
                points.append(good)
            else:
                # host exception lane, evaluated here so that a raising polygon raises in row order
                poly_cnt.append(0); lane.append(1)
                points.append(_hostlane.corner_points(good))
            row_objs.append(obj)
        docs.append(doc); objs_per_row.append(row_objs); img_cnt.append(len(row_objs))
    n_rows = len(docs)
    img_off = np.zeros(n_rows + 1, np.int64)
    np.cumsum(np.frombuffer(img_cnt, dtype=np.int64) if len(img_cnt) else np.zeros(0, np.int64), out=img_off[1:])
    poly_off = np.zeros(len(poly_cnt) + 1, np.int64)
    np.cumsum(np.frombuffer(poly_cnt, dtype=np.int64) if len(poly_cnt) else np.zeros(0, np.int64), out=poly_off[1:])
    return PolygonBatch(n_rows, docs, objs_per_row, points,
                        np.frombuffer(lane, dtype=np.uint8).copy() if len(lane) else np.zeros(0, np.uint8),
                        img_off, poly_off,
                        np.frombuffer(xy, dtype=np.float64).copy() if len(xy) else np.zeros(0, np.float64))


@dataclass
class BoxBatch:
    """CSR of the two-point boxes of every row (step 5 input)."""
    n_rows: int
    img_off: np.ndarray             # int64[n_rows+1]
    pts: np.ndarray                 # float64[4*n_box]  (p1.x, p1.y, p2.x, p2.y)
    valid: np.ndarray               # uint8[n_box]      0 = the scan ends here (the object raised)
    host_rows: list = field(default_factory=list)   # rows holding a value fp64 cannot carry -> host lane


def parse_boxes(cells) -> BoxBatch:
    img_cnt, valid = array("q"), array("B")
    pts = array("d")
    host_rows = []
    for r, text in enumerate(cells):
        row_pts, row_valid = [], []
        exotic = False
        try:
            if isinstance(text, str):
                for obj in json.loads(text).get("objects", []):
                    if not isinstance(obj, dict):
                        continue
                    pl = obj.get("polygon", {}).get("ptList", [])
                    if len(pl) != 2:
                        continue
                    p, q = pl
                    if not (isinstance(p, dict) and isinstance(q, dict)
                            and "x" in p and "y" in p and "x" in q and "y" in q):
                        continue
                    vals = (p["x"], p["y"], q["x"], q["y"])
                    if vals[0] is None and vals[1] is None and vals[2] is None and vals[3] is None:
                        raise TypeError("null bbox")       # min(None, None) raises (processor.py:359)
                    if not all(_dev_ok(v, _IOU_INT) for v in vals):
                        exotic = True
                        break
                    row_pts.extend(vals); row_valid.append(1)
        except Exception:                                   # noqa: BLE001 - the reference swallows everything here
            row_pts.extend((0.0, 0.0, 0.0, 0.0)); row_valid.append(0)
        if exotic:
            host_rows.append(r)
            row_pts, row_valid = [], []
        pts.extend(row_pts); valid.extend(row_valid)
        img_cnt.append(len(row_valid))
    n_rows = len(img_cnt)
    img_off = np.zeros(n_rows + 1, np.int64)
    np.cumsum(np.frombuffer(img_cnt, dtype=np.int64) if n_rows else np.zeros(0, np.int64), out=img_off[1:])
    return BoxBatch(n_rows, img_off,
                    np.frombuffer(pts, dtype=np.float64).copy() if len(pts) else np.zeros(0, np.float64),
                    np.frombuffer(valid, dtype=np.uint8).copy() if len(valid) else np.zeros(0, np.uint8),
                    host_rows)


# ------------------------------------------------------------------ string columns
def pack_strings(values):
    """Column of Python objects -> Arrow-style (off int64[n+1], data uint8, null uint8[n]).

    A cell is null when it is not a ``str`` (NaN / None), matching pandas' isna on object or
    str columns produced by ``read_csv``; non-string scalars are keyed by ``str(value)`` like
    ``astype(str)`` (processor.py:194-198).
    """
    n = len(values)
    try:                                            # Arrow fast path: all cells are str or missing
        import pyarrow as pa
        arr = pa.array(values, type=pa.large_string(), from_pandas=True)
        if arr.offset != 0:
            arr = pa.concat_arrays([arr])
        bufs = arr.buffers()
        off = np.frombuffer(bufs[1], dtype=np.int64, count=n + 1).copy()
        data = np.frombuffer(bufs[2], dtype=np.uint8).copy() if bufs[2] is not None else np.zeros(0, np.uint8)
        null = np.zeros(n, np.uint8) if arr.null_count == 0 else (~np.asarray(arr.is_valid())).astype(np.uint8)
        if arr.null_count:                          # Arrow leaves the offsets of null slots untouched (zero length)
            pass
        return off, data, null
    except Exception:  # noqa: BLE001 - mixed-type column: CPython loop below
        pass
    null = np.zeros(n, np.uint8)
    chunks = []
    off = np.zeros(n + 1, np.int64)
    pos = 0
    for i, v in enumerate(values):
        if isinstance(v, str):
            b = v.encode("utf-8")
        elif v is None or (isinstance(v, float) and v != v):
            null[i] = 1; b = b""
        else:
            try:
                import pandas as pd
                if pd.isna(v):
                    null[i] = 1; b = b""
                else:
                    b = str(v).encode("utf-8")
            except (TypeError, ValueError):
                b = str(v).encode("utf-8")
        chunks.append(b); pos += len(b); off[i + 1] = pos
    data = np.frombuffer(b"".join(chunks), dtype=np.uint8).copy() if pos else np.zeros(0, np.uint8)
    return off, data, null
