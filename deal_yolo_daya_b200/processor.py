"""Drop-in for the hot-path steps of the reference's ``core/processor.py``.

Same function names, positional signatures, CSV/Excel file contract, printed log lines and error
behaviour as /root/reference/src/deal_yolo_data/core/processor.py (the Streamlit page calls
them positionally, ui/pages/processing.py:548-630), so a maintainer can point the page's import
at this module.  Each step also has a DataFrame-level core (``*_df``) for
callers that already hold frames.

What runs where
  * host (this file, ``ingest.py``): CSV/Excel I/O through pandas, JSON decode/encode, packing
    the table into CSR / Arrow buffers, assembling output frames from kernel results;
  * device (libdyd.so through ``KERNELS``): every per-row decision of the path -- string hashing,
    first-occurrence dedup, anti-join, polygon -> corner points with argmin indices, box count +
    any-pair IoU, label-id remap with counters and histogram, category expansion and split ids.

There is no CPU implementation of those decisions here; without the CUDA library the calls
raise.  64-bit hash collisions are ruled out by comparing the strings of every dropped row with
the row that caused the drop (the kernels return it).
"""
from __future__ import annotations

import json
import os
import time
from pathlib import Path
from typing import Optional

import numpy as np
import pandas as pd

from . import _hostlane, ingest, tablecache

COL_SRC = "source"
COL_ANN = "结果字段-目标检测标签配置"
COL_NEW = "新_结果字段-目标检测标签配置"


# =============================================================================================
# kernel facade: numpy in / numpy out, all work on the GPU
# =============================================================================================
class CudaKernels:
    """Moves packed host buffers to the device, runs the libdyd.so kernels, brings results back."""

    def __init__(self, device: Optional[int] = None):
        self._device = device

    def _dev(self):
        import torch
        from . import _lib
        if not torch.cuda.is_available():
            raise _lib.DydError("no CUDA device: deal_yolo_daya_b200.processor has no CPU fallback")
        idx = self._device if self._device is not None else int(os.environ.get("DYD_DEVICE", "0"))
        return torch.device("cuda", idx)

    @staticmethod
    def _up(a, dev):
        import torch
        t0 = time.perf_counter()
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        _phase("h2d", t0)
        return t

    @staticmethod
    def _down(*tensors):
        """Results back to host arrays; the wait for the kernels and the copy are booked separately."""
        import torch
        t0 = time.perf_counter()
        torch.cuda.synchronize()
        t1 = _phase("kernels", t0)
        out = tuple(t.cpu().numpy() for t in tensors)
        _phase("d2h", t1)
        return out

    def bbox_fused(self, img_off, poly_off, xy, min_boxes, thr):
        """K1 + K2 in one pass over the vertices (dyd_bbox_iou_fused, the sm_100a staged kernel):
        (pts, valid, arg, high, count)."""
        from . import ops
        d = self._dev()
        o = ops.bbox_iou_fused(self._up(img_off, d), self._up(poly_off, d), self._up(xy, d), int(min_boxes), float(thr), want_arg=True)
        return self._down(o.pts, o.valid, o.arg, o.high, o.count)

    def bbox(self, poly_off, xy):
        from . import ops
        d = self._dev()
        pts, valid, arg = ops.bbox_minmax(self._up(poly_off, d), self._up(xy, d), want_arg=True)
        return self._down(pts, valid, arg)

    def iou(self, img_off, pts, valid, min_boxes, thr):
        from . import ops
        d = self._dev()
        high, count = ops.iou_filter(self._up(img_off, d), self._up(pts, d), self._up(valid, d), min_boxes, thr)
        return self._down(high, count)

    def hash(self, off, data):
        from . import ops
        d = self._dev()
        return ops.hash_strings(self._up(off, d), self._up(data, d))

    def dedup(self, off, data, null, keep):
        from . import ops
        d = self._dev()
        km, rep = ops.dedup(self.hash(off, data), self._up(null, d), keep)
        return self._down(km, rep)

    def antijoin(self, moff, mdata, mnull, roff, rdata, rnull):
        from . import ops
        d = self._dev()
        km, rr = ops.antijoin(self.hash(moff, mdata), self._up(mnull, d), self.hash(roff, rdata), self._up(rnull, d))
        return self._down(km, rr)

    def label_lut(self, img_off, label_id, lut_new, lut_ntok, lut_nrep):
        from . import ops
        d = self._dev()
        lid = self._up(label_id, d)
        new, rr, cnt = ops.label_lut(self._up(img_off, d), lid, self._up(lut_new, d), self._up(lut_ntok, d), self._up(lut_nrep, d))
        hist = ops.label_hist(lid, len(lut_new))
        c = cnt.cpu().numpy()
        return new.cpu().numpy(), rr.cpu().numpy(), {k: int(v) for k, v in zip(ops.COUNTER_NAMES, c)}, hist.cpu().numpy()

    def split_expand(self, img_off, label_id, cat_of_label, n_cat):
        from . import ops
        d = self._dev()
        ei, eb, ec, co = ops.split_expand(self._up(img_off, d), self._up(label_id, d), self._up(cat_of_label, d), n_cat)
        return ei.cpu().numpy(), eb.cpu().numpy(), ec.cpu().numpy(), co.cpu().numpy()

    def split_assign(self, cat_off, perm, n_train, n_val):
        from . import ops
        d = self._dev()
        s, p = ops.split_assign(self._up(cat_off, d), self._up(perm, d), self._up(n_train, d), self._up(n_val, d))
        return s.cpu().numpy(), p.cpu().numpy()

    def yolo(self, img_off, pts, img_wh):
        """(cxcywh float64[4*n_box], ok uint8[n_box]) of processor.py:1045-1052 for boxes grouped by row."""
        from . import ops
        d = self._dev()
        out, ok = ops.yolo_normalise(self._up(img_off, d), self._up(pts, d), None, self._up(img_wh, d))
        return out.cpu().numpy(), ok.cpu().numpy()

    def label_presence(self, img_off, label_id, n_vocab):
        from . import ops
        d = self._dev()
        ih, bh = ops.label_presence(self._up(img_off, d), self._up(label_id, d), n_vocab)
        return ih.cpu().numpy().astype(np.int64), bh.cpu().numpy().astype(np.int64)

    def hist(self, ids, n_vocab):
        from . import ops
        return ops.label_hist(self._up(ids, self._dev()), n_vocab).cpu().numpy().astype(np.int64)


KERNELS = CudaKernels()


def _to_csv(df: pd.DataFrame, path, encoding, rows=None, clean=False, extras=None, returns_frame=True):
    """``df.iloc[rows].to_csv(path, index=False, encoding=encoding)`` (all rows when ``rows`` is None) through the native
    writer (byte-identical; pandas writes the frames the native writer does not cover), and the written frame remembered
    for the next step when reading the file back could not change it (tablecache).  Returns the written frame with a
    RangeIndex when ``returns_frame`` (the row gather runs on a second thread while this one writes the file); with
    ``returns_frame=False`` nothing is gathered unless a later step asks the cache for it."""
    from . import native
    t0 = time.perf_counter()
    fut = None
    if rows is not None and returns_frame:
        fut = _pool().submit(lambda: df.iloc[rows].reset_index(drop=True))
    try:
        native_wrote = native.to_csv(df, path, encoding, rows=rows)
    finally:
        out = fut.result() if fut is not None else None
    t1 = _phase("write", t0)
    if rows is None:
        out = df if isinstance(df.index, pd.RangeIndex) else df.reset_index(drop=True)
    enc = (encoding or "utf-8").lower().replace("_", "-")
    if native_wrote and enc in ("utf-8", "utf-8-sig", "utf8") and tablecache.enabled():
        if out is not None:
            if native.roundtrip_safe(out, check_cells=not clean):
                tablecache.put(path, out.copy(deep=False), extras)
            else:
                tablecache.forget(path); tablecache.declined()
        else:                                      # gathered (and checked) only if a later step reads this file
            tablecache.put(path, None, extras, lazy=(df, rows, lambda f: native.roundtrip_safe(f, check_cells=not clean)))
    else:
        tablecache.forget(path)
    _phase("cache_check", t1)
    return out


def _read_csv(path, **kwargs):
    """pd.read_csv(path, **kwargs) with the text columns tokenised natively (csrc/csv_read.cpp); the
    same frame, dtype for dtype (tests/test_native_csv_read.py)."""
    return _load(path, **kwargs)[0]


def _load(path, **kwargs):
    """-> (frame, cache entry or None, clean).  The frame of an unchanged file this process wrote or read a moment ago
    comes from tablecache (no parse); `clean` = its text cells are known to come from the CSV reader."""
    from . import native
    t0 = time.perf_counter()
    enc = str(kwargs.get("encoding", "utf-8") or "utf-8").lower().replace("_", "-")
    plain = enc == "utf-8-sig" and all(k == "encoding" or (k == "parse_dates" and v is False) for k, v in kwargs.items())
    if plain:
        ent = tablecache.get(path)
        if ent is not None:
            _phase("read_cached", t0)
            return ent.frame.copy(deep=False), ent, ent.clean
    before = native._READ_STATS["native"]
    df = native.read_csv(path, **kwargs)
    clean = native._READ_STATS["native"] > before
    if plain and isinstance(df.index, pd.RangeIndex):
        tablecache.put(path, df.copy(deep=False), None, clean=clean)
    _phase("read", t0)
    return df, None, clean


PHASES = {}            # seconds per phase of the calls since the caller last cleared it (observability; tools/dropin_phases.py)
_POOL = None


def _phase(name, t0):
    t1 = time.perf_counter()
    PHASES[name] = PHASES.get(name, 0.0) + (t1 - t0)
    return t1


def _pool():
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=2, thread_name_prefix="dyd-gather")
    return _POOL


STATS = {"hostlane_objects": 0, "hostlane_rows": 0, "hash_collisions": 0, "native_rows": 0, "slow_rows": 0}    # observability, last call


# =============================================================================================
# step 2: dedup by source                                            reference: processor.py:111-164
# =============================================================================================
def _source_buffers(col, normalise_zero=False):
    """`source` column -> (off int64[n+1], data uint8[], null uint8[n], values or None).  An Arrow-backed text column
    goes to the kernels as its own buffers (no Python string per row); anything else is packed from its values with
    str() like the reference's astype(str) (`values` is then the list the collision repair compares)."""
    from . import native
    pa_arr = getattr(getattr(col, "array", None), "_pa_array", None)
    if native.enabled() and pa_arr is not None and str(pa_arr.type) in ("large_string", "string"):
        data, off, is_text, keep = native.pack_cells(col)
        if data.size == 0:
            data = np.zeros(1, np.uint8)
        return off, data, (is_text == 0).astype(np.uint8), None, keep
    vals = col.tolist()
    if normalise_zero and col.dtype == np.float64:
        vals = (col.to_numpy() + 0.0).tolist()         # drop_duplicates compares values: -0.0 and 0.0 are one group
    off, data, null = ingest.pack_strings(vals)
    return off, data, null, vals, None


def _rows_equal(off_a, data_a, a, off_b, data_b, b) -> np.ndarray:
    """bool per pair: string a[i] of column A equals string b[i] of column B (byte compare on the packed buffers)."""
    la = off_a[a + 1] - off_a[a]; lb = off_b[b + 1] - off_b[b]
    eq = la == lb
    idx = np.nonzero(eq)[0]
    if idx.size:
        ln = la[idx]
        tot = int(ln.sum())
        if tot:
            starts = np.zeros(idx.size + 1, np.int64); np.cumsum(ln, out=starts[1:])
            within = np.arange(tot, dtype=np.int64) - np.repeat(starts[:-1], ln)
            pa_ = np.repeat(off_a[a[idx]], ln) + within; pb_ = np.repeat(off_b[b[idx]], ln) + within
            diff = data_a[pa_] != data_b[pb_]
            if diff.any():
                bad_pair = np.unique(np.searchsorted(starts, np.nonzero(diff)[0], side="right") - 1)
                eq[idx[bad_pair]] = False
    return eq


def dedup_keep_mask(col, keep="first") -> np.ndarray:
    """uint8 keep mask of ``drop_duplicates(subset=[source], keep=keep)`` for one `source` column (K0 + K4), every dropped
    row checked against the row that caused the drop (a 64-bit hash collision regroups on the strings)."""
    STATS["hash_collisions"] = 0
    t0 = time.perf_counter()
    off, data, null, vals, _keepalive = _source_buffers(col, normalise_zero=True)
    t1 = _phase("ingest", t0)
    km, rep = KERNELS.dedup(off, data, null, keep)
    t1 = time.perf_counter()
    # collision check: a row may only be dropped because of a row holding the SAME string
    dropped = np.nonzero((km == 0) & (rep != np.arange(len(km))))[0]
    if dropped.size:
        r = rep[dropped]
        na, nb = null[dropped] != 0, null[r] != 0
        same = np.where(na | nb, na & nb, _rows_equal(off, data, dropped, off, data, r))
        if not same.all():
            STATS["hash_collisions"] = int((~same).sum())
            if vals is None:
                vals = col.tolist()
            km = _dedup_exact(vals, null, keep)
    _phase("verify", t1)
    return km


def deduplicate_df(df: pd.DataFrame, keep="first") -> pd.DataFrame:
    """``df.drop_duplicates(subset=["source"], keep=keep, ignore_index=True)`` with the grouping on the GPU."""
    if keep not in ("first", "last", False):
        raise ValueError('keep must be either "first", "last" or False')
    if len(df) == 0:
        return df.reset_index(drop=True)
    km = dedup_keep_mask(df[COL_SRC], keep)
    return df[km.astype(bool)].reset_index(drop=True)


def _dedup_exact(vals, null, keep):
    """Collision repair: regroup by the strings themselves (runs only if two different strings share a 64-bit hash)."""
    groups = {}
    for r, v in enumerate(vals):
        groups.setdefault(None if null[r] else str(v), []).append(r)
    km = np.zeros(len(vals), np.uint8)
    for rows in groups.values():
        if keep == "first":
            km[rows[0]] = 1
        elif keep == "last":
            km[rows[-1]] = 1
        elif len(rows) == 1:
            km[rows[0]] = 1
    return km


def deduplicate_csv_by_source(
        csv_path: str,
        output_file: Optional[str] = "deduplicate_result.csv",
        encoding: str = "utf-8-sig",
        keep: str = "first",
        verbose: bool = True
) -> pd.DataFrame:
    if not os.path.exists(csv_path):
        raise FileNotFoundError(f"CSV文件不存在：{csv_path}")
    if not csv_path.endswith(".csv"):
        raise ValueError(f"文件不是CSV格式：{csv_path}（请传入.csv后缀的文件）")
    try:
        df, _, clean = _load(csv_path, encoding=encoding, parse_dates=False)
        if verbose:
            print(f"成功读取CSV文件：{os.path.basename(csv_path)}")
            print(f"读取后原始数据行数：{len(df)}")
    except Exception as e:
        raise Exception(f"读取CSV文件失败：{str(e)}") from e
    if COL_SRC not in df.columns:
        raise KeyError(f"CSV文件中未找到'source'列，请检查列名是否正确（当前列名：{list(df.columns)}）")
    if keep not in ("first", "last", False):
        raise ValueError('keep must be either "first", "last" or False')
    rows = np.nonzero(dedup_keep_mask(df[COL_SRC], keep))[0] if len(df) else np.zeros(0, np.int64)
    if verbose:
        print(f"去重策略：按'source'列保留{keep}条数据")
        print(f"去除重复数据行数：{len(df) - len(rows)}")
        print(f"去重后剩余数据行数：{len(rows)}")
    sel = None if len(rows) == len(df) else rows
    if output_file is not None:
        try:
            d = os.path.dirname(output_file)
            if d and not os.path.exists(d):
                os.makedirs(d, exist_ok=True)
            out = _to_csv(df, output_file, encoding, rows=sel, clean=clean)     # the file write and the row gather overlap
            if verbose:
                print(f"去重后的文件已保存至：{os.path.abspath(output_file)}")
        except Exception as e:
            raise Exception(f"保存去重文件失败：{str(e)}") from e
    else:
        out = df.reset_index(drop=True) if sel is None else df.iloc[sel].reset_index(drop=True)
    return out


# =============================================================================================
# step 3: anti-join against the reference set                        reference: processor.py:166-219
# =============================================================================================
def antijoin_keep_mask(main_col, ref_col):
    """(uint8 keep mask of the main rows, number of distinct reference values) of processor.py:194-199 (K0 + K5), every
    dropped row checked against the reference row that caused the drop."""
    STATS["hash_collisions"] = 0
    t0 = time.perf_counter()
    moff, mdata, mnull, mvals, _k1 = _source_buffers(main_col)
    roff, rdata, rnull, rvals, _k2 = _source_buffers(ref_col)
    t1 = _phase("ingest", t0)
    if rvals is None:                               # distinct non-null reference strings, counted on the Arrow column
        import pyarrow.compute as pc
        arr = ref_col.array._pa_array
        n_ref_unique = int(pc.count_distinct(arr, mode="only_valid").as_py())
    else:
        n_ref_unique = len(set(str(v) for v, z in zip(rvals, rnull) if not z))
    _phase("ref_unique", t1)
    if len(main_col) == 0:
        return np.zeros(0, np.uint8), n_ref_unique
    km, rr = KERNELS.antijoin(moff, mdata, mnull, roff, rdata, rnull)
    t1 = time.perf_counter()
    dropped = np.nonzero(km == 0)[0]
    if dropped.size:
        same = _rows_equal(moff, mdata, dropped, roff, rdata, rr[dropped])
        if not same.all():                          # 64-bit collision: decide those rows on the strings
            STATS["hash_collisions"] = int((~same).sum())
            mv = main_col.tolist() if mvals is None else mvals
            rv = ref_col.tolist() if rvals is None else rvals
            seen = set(str(v) for v, z in zip(rv, rnull) if not z)
            km = km.copy()
            for r in dropped[~same]:
                km[r] = 0 if str(mv[r]) in seen else 1
    _phase("verify", t1)
    return km, n_ref_unique


def remove_duplicates_df(df_main: pd.DataFrame, df_ref: pd.DataFrame, compare_col: str = COL_SRC):
    """``df_main[~df_main[col].astype(str).isin(set(df_ref[col].dropna().astype(str)))].copy()``.
    Returns (filtered frame, number of distinct reference values)."""
    km, n_ref_unique = antijoin_keep_mask(df_main[compare_col], df_ref[compare_col])
    if len(df_main) == 0:
        return df_main.copy(), n_ref_unique
    return df_main[km.astype(bool)].copy(), n_ref_unique


def remove_duplicates_between_csv(
        main_csv: str,
        ref_csv: str,
        output_csv: str = "filtered_main.csv",
        compare_col: str = "source",
        encoding: str = "utf-8-sig",
        verbose: bool = True
) -> pd.DataFrame:
    for p in [main_csv, ref_csv]:
        if not os.path.exists(p):
            raise FileNotFoundError(f"文件不存在：{p}")
        if not p.endswith(".csv"):
            raise ValueError(f"文件不是CSV格式：{p}（请传入.csv后缀文件）")
    try:
        df_main, _, clean = _load(main_csv, encoding=encoding, parse_dates=False)
        df_ref = _read_csv(ref_csv, encoding=encoding, parse_dates=False)
        if verbose:
            print(f"读取主文件：{len(df_main)}行")
            print(f"读取参考文件：{len(df_ref)}行")
    except Exception as e:
        raise Exception(f"读取CSV失败：{str(e)}") from e
    if compare_col not in df_main.columns:
        raise KeyError(f"主文件中未找到列 '{compare_col}'")
    if compare_col not in df_ref.columns:
        raise KeyError(f"参考文件中未找到列 '{compare_col}'")
    km, n_unique = antijoin_keep_mask(df_main[compare_col], df_ref[compare_col])
    rows = np.nonzero(km)[0]
    if verbose:
        print(f"去重依据列：{compare_col}")
        print(f"参考文件中唯一值数量：{n_unique}")
        print(f"剔除重复行数：{len(df_main) - len(rows)}")
        print(f"保留行数：{len(rows)}")
    try:
        d = os.path.dirname(output_csv)
        if d and not os.path.exists(d):
            os.makedirs(d, exist_ok=True)
        out = _to_csv(df_main, output_csv, encoding, rows=None if len(rows) == len(df_main) else rows, clean=clean)
        if verbose:
            print(f"结果已保存至：{os.path.abspath(output_csv)}")
    except Exception as e:
        raise Exception(f"保存结果失败：{str(e)}") from e
    # the reference returns df_main[~is_dup].copy(): the surviving rows keep their original index labels
    if len(rows) != len(df_main):
        out = out.set_axis(df_main.index[rows], axis=0)
    elif out is df_main or not out.index.equals(df_main.index):
        out = df_main.copy()
    return out


# =============================================================================================
# step 4: ptList -> two corner points                                reference: processor.py:229-319
# =============================================================================================
IOU_SPECULATION = [2, 0.98]      # (min_boxes, threshold) step 4 evaluates ahead of step 5: the reference's defaults
                                 # (processor.py:321-327) until a step-5 call has shown what this session uses


def replace_ptlist_cells(cells, side=None):
    """Annotation JSON texts -> (new JSON texts | None, widths, heights).

    ``side`` (a dict, optional) receives what step 5 can reuse when every row took the native lane: the boxes of the new
    cells as CSR (img_off, pts, valid) and the IoU flags the fused kernel computed on the way for IOU_SPECULATION.

    K1 finds, per polygon, which vertex supplies min_x / min_y / max_x / max_y; the output cell is
    the reference's ``json.dumps`` of the document with every dict object's ptList replaced by the
    two corner points built from the ORIGINAL JSON numbers at those indices (int stays int).

    Fast lane: the native parser (csrc/ingest.cpp) builds the CSR buffers and splices the output
    text for every row whose text is in json.dumps form; rows it flags go through the CPython lane
    below (json.loads / json.dumps), which is also the whole path when DYD_NATIVE_INGEST=0.
    """
    from . import native
    if not native.enabled() or len(cells) == 0:
        return _replace_ptlist_cells_python(list(cells))
    t0 = time.perf_counter()
    ing = native.Ingest(cells, 0).polygons()              # an Arrow-backed column goes in without a copy
    t0 = _phase("ingest", t0)
    cell_at = cells.iloc.__getitem__ if hasattr(cells, "iloc") else cells.__getitem__
    try:
        STATS["native_rows"] = int((ing.status == native.ROW_OK).sum())
        all_ok = bool(ing.n and ing.arrow_input and STATS["native_rows"] == ing.n)
        fused = None
        if ing.n_obj and all_ok and hasattr(KERNELS, "bbox_fused"):
            # one pass over the vertices: corner points with their vertex indices AND the IoU flags of the likely step 5
            mb, thr = IOU_SPECULATION
            pts, valid, arg, high, count = KERNELS.bbox_fused(ing.img_off, ing.poly_off, ing.xy, mb, thr)
            # step 5 may reuse these boxes only where fp64 arithmetic equals CPython's on the JSON numbers: an int coordinate
            # beyond 2^25 is computed exactly there (ingest._IOU_INT); such tables are parsed again by step 5's own lanes
            if not pts.size or float(np.nanmax(np.abs(np.where(np.isfinite(pts), pts, 0.0)))) <= float(ingest._IOU_INT):
                fused = {"img_off": ing.img_off.copy(), "pts": pts, "valid": valid, "flags": {(int(mb), float(thr)): high.astype(bool)}}
        elif ing.n_obj:
            _, valid, arg = KERNELS.bbox(ing.poly_off, ing.xy)
        else:
            valid = np.zeros(0, np.uint8); arg = np.zeros(0, np.int32)
        t0 = time.perf_counter()
        out_bytes, out_off = ing.egress_ptlist(arg, valid)
        t0 = _phase("egress", t0)
        if all_ok:
            wh = ing.int_columns()
            if wh is not None:                         # every row native, width / height plain ints: no per-row Python
                STATS["hostlane_objects"] = 0; STATS["slow_rows"] = 0
                if side is not None and fused is not None:
                    side["boxes"] = fused
                _phase("egress", t0)
                return native.arrow_strings(out_bytes, out_off), wh[0], wh[1]
        blob = out_bytes.tobytes()
        out, widths, heights = [None] * ing.n, [None] * ing.n, [None] * ing.n
        slow = []
        for r in range(ing.n):
            st = ing.status[r]
            if st == native.ROW_OK:
                out[r] = blob[out_off[r]:out_off[r + 1]].decode("utf-8", "surrogatepass")
                widths[r] = ing.scalar(r, 0); heights[r] = ing.scalar(r, 1)
            elif st == native.ROW_SLOW:
                slow.append(r)
    finally:
        ing.close()
    STATS["hostlane_objects"] = 0
    if slow:                                   # CPython lane for the rows the native parser declined
        o2, w2, h2 = _replace_ptlist_cells_python([cell_at(r) for r in slow])
        for i, r in enumerate(slow):
            out[r], widths[r], heights[r] = o2[i], w2[i], h2[i]
    STATS["slow_rows"] = len(slow)
    return out, widths, heights


def _replace_ptlist_cells_python(cells):
    """CPython lane of step 4: json.loads per cell, K1 on the packed vertices, json.dumps per cell."""
    batch = ingest.parse_polygons(cells)
    STATS["hostlane_objects"] = int(batch.hostlane.sum())
    if batch.n_obj:
        pts, valid, arg = KERNELS.bbox(batch.poly_off, batch.xy)
    else:
        valid = np.zeros(0, np.uint8); arg = np.zeros(0, np.int32)
    out, widths, heights = [], [], []
    q = 0
    for r in range(batch.n_rows):
        doc = batch.docs[r]
        if doc is None:
            out.append(None); widths.append(None); heights.append(None)
            continue
        new_objs = []
        for obj in batch.objs[r]:
            new = obj.copy()
            good = batch.points[q]
            if batch.hostlane[q]:
                corners = good                                    # evaluated in ingest (CPython semantics)
            elif not valid[q]:
                corners = [{"x": None, "y": None}, {"x": None, "y": None}]
            else:
                a = arg[4 * q:4 * q + 4]
                corners = [{"x": good[a[0]]["x"], "y": good[a[1]]["y"]}, {"x": good[a[2]]["x"], "y": good[a[3]]["y"]}]
            if "polygon" not in new:
                new["polygon"] = {}
            new["polygon"]["ptList"] = corners
            new_objs.append(new)
            q += 1
        widths.append(doc.get("width")); heights.append(doc.get("height"))
        doc["objects"] = new_objs
        out.append(json.dumps(doc, ensure_ascii=False))
    return out, widths, heights


def replace_ptlist_df(df: pd.DataFrame, side=None):
    """DataFrame core of step 4 -> (result frame with the reference's column subset, excluded frame)."""
    missing = df[COL_ANN].isna()
    if not bool(missing.any()):                    # nothing to exclude: no row gather at all
        kept = df.copy(deep=False)
        excluded = df.iloc[0:0].copy()
    else:
        kept = df[~missing].copy()
        excluded = df[missing].copy()
    new, w, h = replace_ptlist_cells(kept[COL_ANN], side)
    if isinstance(new, list):
        kept[COL_NEW] = pd.Series(new, index=kept.index, dtype=object) if len(new) else pd.Series([], index=kept.index, dtype=object)
    else:
        kept[COL_NEW] = pd.Series(new, index=kept.index)
    kept["width"] = w
    kept["height"] = h
    cols = [c for c in (COL_SRC, COL_ANN, COL_NEW, "width", "height") if c in kept.columns]
    return kept[cols], excluded


def process_csv_replace_ptlist(
        input_csv_path: str,
        output_csv_path: str = "processed_replaced_ptlist.csv",
        excluded_output_file: Optional[str] = "processed_excluded.csv"
):
    try:
        df, _, clean = _load(input_csv_path, encoding="utf-8-sig")
        print(f"成功读取CSV，共 {len(df)} 行数据")
    except FileNotFoundError:
        print(f"错误：未找到文件 {input_csv_path}")
        return None
    except Exception as e:
        print(f"读取失败：{e}")
        return None
    if COL_ANN not in df.columns:
        print(f"错误：CSV缺少列 '{COL_ANN}'")
        return None
    side = {}
    res, excluded = replace_ptlist_df(df, side)
    Path(output_csv_path).parent.mkdir(parents=True, exist_ok=True)
    extras = None
    if "boxes" in side and len(side["boxes"]["img_off"]) == len(res) + 1:
        extras = {"boxes": side["boxes"], "boxes_column": COL_NEW}        # step 5 on this very file needs no JSON parse
    _to_csv(res, output_csv_path, "utf-8-sig", clean=clean, extras=extras)
    if excluded_output_file is not None:
        Path(excluded_output_file).parent.mkdir(parents=True, exist_ok=True)
        _to_csv(excluded, excluded_output_file, "utf-8-sig", clean=clean)
    return {"filtered_rows": len(res), "excluded_rows": len(excluded), "excluded_output": excluded_output_file}


# =============================================================================================
# step 5: box-count + IoU quality filter                             reference: processor.py:321-407
# =============================================================================================
def high_iou_mask(cells, min_boxes: int = 2, iou_threshold: float = 0.98) -> np.ndarray:
    """bool per row: len(boxes) >= min_boxes and some pair has IoU >= threshold (K2).

    Fast lane: native parse of the two-point boxes (csrc/ingest.cpp, mode 1); rows it declines are
    parsed by the CPython lane; rows holding values fp64 cannot carry are evaluated by _hostlane.
    """
    from . import native
    n = len(cells)
    if n == 0:
        return np.zeros(0, bool)
    cell_at = cells.iloc.__getitem__ if hasattr(cells, "iloc") else cells.__getitem__
    mask = np.zeros(n, bool)
    slow = list(range(n))
    if native.enabled():
        t0 = time.perf_counter()
        ing = native.Ingest(cells, 1).boxes()
        _phase("ingest", t0)
        try:
            if ing.n > 0:
                high, _ = KERNELS.iou(ing.img_off, ing.pts, ing.valid, min_boxes, iou_threshold)
                ok = ing.status != native.ROW_SLOW          # non-text cells have no boxes: not high
                mask[ok] = high.astype(bool)[ok]
            slow = [int(r) for r in np.nonzero(ing.status == native.ROW_SLOW)[0]]
        finally:
            ing.close()
    STATS["slow_rows"] = len(slow)
    STATS["hostlane_rows"] = 0
    if slow:
        sub = [cell_at(r) for r in slow]
        batch = ingest.parse_boxes(sub)
        STATS["hostlane_rows"] = len(batch.host_rows)
        high, _ = KERNELS.iou(batch.img_off, batch.pts, batch.valid, min_boxes, iou_threshold)
        m2 = high.astype(bool)
        for i in batch.host_rows:                   # rows holding values fp64 cannot carry
            m2[i] = _hostlane.row_is_high_iou(sub[i], min_boxes, iou_threshold)
        mask[np.array(slow)] = m2
    return mask


def filter_by_box_count_and_iou_df(df: pd.DataFrame, min_boxes: int = 2, iou_threshold: float = 0.98):
    """DataFrame core of step 5 -> (high_iou frame, other frame); all columns, original order."""
    m = pd.Series(high_iou_mask(df[COL_NEW], min_boxes, iou_threshold), index=df.index, dtype=bool)
    return df[m], df[~m]


def _mask_from_boxes(boxes, min_boxes, iou_threshold) -> np.ndarray:
    """Step 5's row mask from the boxes step 4 left with the file it wrote (same values the JSON text holds: the corner
    points are the original number literals K1 selected): the flags the fused kernel already computed when the
    parameters are the speculated ones, else K2 on the cached boxes -- no JSON parse either way."""
    key = (int(min_boxes), float(iou_threshold))
    STATS["slow_rows"] = 0; STATS["hostlane_rows"] = 0
    hit = boxes["flags"].get(key) if isinstance(min_boxes, (int, np.integer)) and not isinstance(min_boxes, bool) else None
    if hit is None:
        high, _ = KERNELS.iou(boxes["img_off"], boxes["pts"], boxes["valid"], min_boxes, iou_threshold)
        hit = high.astype(bool)
        if len(boxes["flags"]) < 8:
            boxes["flags"][key] = hit
    return hit


def filter_by_box_count_and_iou(
        input_csv_path,
        high_iou_csv="high_iou_0.98.csv",
        other_csv="other_data.csv",
        min_boxes: int = 2,
        iou_threshold: float = 0.98
):
    try:
        df, ent, clean = _load(input_csv_path, encoding="utf-8-sig")
    except Exception as e:
        print(f"读取失败：{e}")
        return
    if COL_NEW not in df.columns:
        print(f"错误：缺少必要列 {COL_NEW}")
        return
    boxes = ent.extras.get("boxes") if ent is not None and ent.extras.get("boxes_column") == COL_NEW else None
    if boxes is not None and len(boxes["img_off"]) == len(df) + 1:
        m = _mask_from_boxes(boxes, min_boxes, iou_threshold)
    else:
        m = high_iou_mask(df[COL_NEW], min_boxes, iou_threshold)
    try:
        IOU_SPECULATION[:] = [int(min_boxes), float(iou_threshold)]          # what step 4 evaluates ahead next time
    except (TypeError, ValueError):
        pass
    Path(high_iou_csv).parent.mkdir(parents=True, exist_ok=True)
    Path(other_csv).parent.mkdir(parents=True, exist_ok=True)
    # both files straight from the frame by row numbers: the two filtered frames are never built (the step returns None)
    _to_csv(df, high_iou_csv, "utf-8-sig", rows=np.nonzero(m)[0], clean=clean, returns_frame=False)
    _to_csv(df, other_csv, "utf-8-sig", rows=np.nonzero(~m)[0], clean=clean, returns_frame=False)


# =============================================================================================
# step 5.5: label remap                                   reference: processor.py:516-652, utils.py:659-679
# =============================================================================================
from .labels import (default_json_columns, mapping_from_frame, remap_df, rules_from_frame, split_df)  # noqa: E402


def replace_labels_by_mapping(
        input_csv_path: str,
        mapping_excel_path: str,
        output_csv_path: str,
        sheet_name: Optional[str] = None,
        old_col: Optional[str] = None,
        new_col: Optional[str] = None,
        json_columns: Optional[list] = None,
        diff_excel_path: Optional[str] = None,
        unmatched_excel_path: Optional[str] = None,
        sample_size: int = 30,
):
    df = _read_csv(input_csv_path, encoding="utf-8-sig")
    mapping_df = pd.read_excel(mapping_excel_path, sheet_name=sheet_name) if sheet_name else pd.read_excel(mapping_excel_path)
    label_map = mapping_from_frame(mapping_df, old_col, new_col)
    out, summary, diff_rows, unmatched = remap_df(df, label_map, json_columns)
    output_csv_path = Path(output_csv_path)
    output_csv_path.parent.mkdir(parents=True, exist_ok=True)
    _to_csv(out, output_csv_path, "utf-8-sig")
    diff_path = None
    if diff_excel_path:
        diff_path = Path(diff_excel_path)
        diff_path.parent.mkdir(parents=True, exist_ok=True)
        pd.DataFrame(diff_rows).to_excel(diff_path, index=False)
    unmatched_path = None
    if unmatched_excel_path:
        unmatched_path = Path(unmatched_excel_path)
        unmatched_path.parent.mkdir(parents=True, exist_ok=True)
        if unmatched:
            pd.DataFrame([{"标签": k, "数量": v} for k, v in unmatched.items()]).sort_values("数量", ascending=False).to_excel(unmatched_path, index=False)
        else:
            pd.DataFrame(columns=["标签", "数量"]).to_excel(unmatched_path, index=False)
    return {"output_csv": output_csv_path, "summary": summary, "diff": diff_path, "unmatched": unmatched_path,
            "sample_diff": diff_rows[:sample_size]}


# =============================================================================================
# step 6: rule-based split                                           reference: processor.py:654-831
# =============================================================================================
def _safe_filename(value: str) -> str:
    """File stem of a category workbook, as utils.safe_filename does it (utils.py:525-529): runs of
    characters outside [A-Za-z0-9._-] become "_", edge underscores are stripped, empty -> "train"
    (so non-ASCII category names all map to the same file, exactly like the reference)."""
    import re
    if not value:
        return "train"
    return re.sub(r"[^A-Za-z0-9._-]+", "_", value).strip("_") or "train"


def split_dataset_by_rules(
        input_csv_path: str,
        rules_excel_path: str,
        output_dir: str,
        rule_mode: str = "wide",
        sheet_name: Optional[str] = None,
        label_col: Optional[str] = None,
        category_col: Optional[str] = None,
        json_columns: Optional[list] = None,
        train_ratio: float = 0.8,
        val_ratio: float = 0.1,
        test_ratio: float = 0.1,
        random_seed: int = 42,
):
    if not os.path.exists(input_csv_path):
        raise FileNotFoundError(f"输入CSV不存在：{input_csv_path}")
    if not os.path.exists(rules_excel_path):
        raise FileNotFoundError(f"规则Excel不存在：{rules_excel_path}")
    df = _read_csv(input_csv_path, encoding="utf-8-sig")
    rules_df = pd.read_excel(rules_excel_path, sheet_name=sheet_name) if sheet_name else pd.read_excel(rules_excel_path)
    l2c = rules_from_frame(rules_df, rule_mode, label_col, category_col)
    res = split_df(df, l2c, json_columns, train_ratio, val_ratio, test_ratio, random_seed)
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    category_files = []
    for cat, parts in res["categories"].items():
        out_path = output_dir / f"{_safe_filename(cat)}.xlsx"
        with pd.ExcelWriter(out_path) as writer:
            parts["train"].to_excel(writer, sheet_name="train", index=False)
            parts["val"].to_excel(writer, sheet_name="val", index=False)
            parts["test"].to_excel(writer, sheet_name="test", index=False)
        category_files.append(out_path)
    unclassified_path = output_dir / "unclassified.xlsx"
    res["unclassified"].to_excel(unclassified_path, index=False)
    split_counts_path = output_dir / "split_counts.xlsx"
    res["split_counts"].to_excel(split_counts_path, index=False)
    return {"output_dir": output_dir, "category_files": category_files, "unclassified": unclassified_path,
            "split_counts": split_counts_path, "summary": res["summary"]}


# =============================================================================================
# the callers either side of the path: merge in front, dataset writer + summaries behind
# reference: processor.py:26-109, 833-891, 893-1087, 1089-1163
# =============================================================================================
from .dataset import (generate_yolo_datasets_from_excels, merge_all_csv_in_folder, summarize_unclassified,  # noqa: E402,F401
                      summarize_yolo_label_counts)
