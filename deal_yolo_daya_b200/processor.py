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
from pathlib import Path
from typing import Optional

import numpy as np
import pandas as pd

from . import _hostlane, ingest

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
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def bbox(self, poly_off, xy):
        from . import ops
        d = self._dev()
        pts, valid, arg = ops.bbox_minmax(self._up(poly_off, d), self._up(xy, d), want_arg=True)
        return pts.cpu().numpy(), valid.cpu().numpy(), arg.cpu().numpy()

    def iou(self, img_off, pts, valid, min_boxes, thr):
        from . import ops
        d = self._dev()
        high, count = ops.iou_filter(self._up(img_off, d), self._up(pts, d), self._up(valid, d), min_boxes, thr)
        return high.cpu().numpy(), count.cpu().numpy()

    def hash(self, off, data):
        from . import ops
        d = self._dev()
        return ops.hash_strings(self._up(off, d), self._up(data, d))

    def dedup(self, off, data, null, keep):
        from . import ops
        d = self._dev()
        km, rep = ops.dedup(self.hash(off, data), self._up(null, d), keep)
        return km.cpu().numpy(), rep.cpu().numpy()

    def antijoin(self, moff, mdata, mnull, roff, rdata, rnull):
        from . import ops
        d = self._dev()
        km, rr = ops.antijoin(self.hash(moff, mdata), self._up(mnull, d), self.hash(roff, rdata), self._up(rnull, d))
        return km.cpu().numpy(), rr.cpu().numpy()

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


def _to_csv(df: pd.DataFrame, path, encoding) -> None:
    """df.to_csv(path, index=False, encoding=encoding) through the native body writer (byte-identical;
    pandas writes the frames the native writer does not cover)."""
    from . import native
    native.to_csv(df, path, encoding)


def _read_csv(path, **kwargs):
    """pd.read_csv(path, **kwargs) with the text columns tokenised natively (csrc/csv_read.cpp); the
    same frame, dtype for dtype (tests/test_native_csv_read.py)."""
    from . import native
    return native.read_csv(path, **kwargs)


STATS = {"hostlane_objects": 0, "hostlane_rows": 0, "hash_collisions": 0, "native_rows": 0, "slow_rows": 0}    # observability, last call


# =============================================================================================
# step 2: dedup by source                                            reference: processor.py:111-164
# =============================================================================================
def _strings_equal(vals, null, a, b):
    if null[a] or null[b]:
        return bool(null[a] and null[b])
    return str(vals[a]) == str(vals[b])


def deduplicate_df(df: pd.DataFrame, keep="first") -> pd.DataFrame:
    """``df.drop_duplicates(subset=["source"], keep=keep, ignore_index=True)`` with the grouping on the GPU."""
    if keep not in ("first", "last", False):
        raise ValueError('keep must be either "first", "last" or False')
    vals = df[COL_SRC].tolist()
    if not vals:
        return df.reset_index(drop=True)
    off, data, null = ingest.pack_strings(vals)
    km, rep = KERNELS.dedup(off, data, null, keep)
    # collision check: a row may only be dropped because of a row holding the SAME string
    dropped = np.nonzero(km == 0)[0]
    bad = [r for r in dropped if rep[r] != r and not _strings_equal(vals, null, r, rep[r])]
    if bad:
        STATS["hash_collisions"] = len(bad)
        km = _dedup_exact(vals, null, keep)
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
        df = _read_csv(csv_path, encoding=encoding, parse_dates=False)
        if verbose:
            print(f"成功读取CSV文件：{os.path.basename(csv_path)}")
            print(f"读取后原始数据行数：{len(df)}")
    except Exception as e:
        raise Exception(f"读取CSV文件失败：{str(e)}") from e
    if COL_SRC not in df.columns:
        raise KeyError(f"CSV文件中未找到'source'列，请检查列名是否正确（当前列名：{list(df.columns)}）")
    out = deduplicate_df(df, keep)
    if verbose:
        print(f"去重策略：按'source'列保留{keep}条数据")
        print(f"去除重复数据行数：{len(df) - len(out)}")
        print(f"去重后剩余数据行数：{len(out)}")
    if output_file is not None:
        try:
            d = os.path.dirname(output_file)
            if d and not os.path.exists(d):
                os.makedirs(d, exist_ok=True)
            _to_csv(out, output_file, encoding)
            if verbose:
                print(f"去重后的文件已保存至：{os.path.abspath(output_file)}")
        except Exception as e:
            raise Exception(f"保存去重文件失败：{str(e)}") from e
    return out


# =============================================================================================
# step 3: anti-join against the reference set                        reference: processor.py:166-219
# =============================================================================================
def remove_duplicates_df(df_main: pd.DataFrame, df_ref: pd.DataFrame, compare_col: str = COL_SRC):
    """``df_main[~df_main[col].astype(str).isin(set(df_ref[col].dropna().astype(str)))].copy()``.
    Returns (filtered frame, number of distinct reference values)."""
    mvals = df_main[compare_col].tolist(); rvals = df_ref[compare_col].tolist()
    moff, mdata, mnull = ingest.pack_strings(mvals)
    roff, rdata, rnull = ingest.pack_strings(rvals)
    n_ref_unique = len(set(str(v) for v, z in zip(rvals, rnull) if not z))
    if not mvals:
        return df_main.copy(), n_ref_unique
    km, rr = KERNELS.antijoin(moff, mdata, mnull, roff, rdata, rnull)
    bad = [r for r in np.nonzero(km == 0)[0] if str(mvals[r]) != str(rvals[rr[r]])]
    if bad:                                        # 64-bit collision: decide those rows on the strings
        STATS["hash_collisions"] = len(bad)
        seen = set(str(v) for v, z in zip(rvals, rnull) if not z)
        for r in bad:
            km[r] = 0 if str(mvals[r]) in seen else 1
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
        df_main = _read_csv(main_csv, encoding=encoding, parse_dates=False)
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
    out, n_unique = remove_duplicates_df(df_main, df_ref, compare_col)
    if verbose:
        print(f"去重依据列：{compare_col}")
        print(f"参考文件中唯一值数量：{n_unique}")
        print(f"剔除重复行数：{len(df_main) - len(out)}")
        print(f"保留行数：{len(out)}")
    try:
        d = os.path.dirname(output_csv)
        if d and not os.path.exists(d):
            os.makedirs(d, exist_ok=True)
        _to_csv(out, output_csv, encoding)
        if verbose:
            print(f"结果已保存至：{os.path.abspath(output_csv)}")
    except Exception as e:
        raise Exception(f"保存结果失败：{str(e)}") from e
    return out


# =============================================================================================
# step 4: ptList -> two corner points                                reference: processor.py:229-319
# =============================================================================================
def replace_ptlist_cells(cells):
    """Annotation JSON texts -> (new JSON texts | None, widths, heights).

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
    ing = native.Ingest(cells, 0).polygons()              # an Arrow-backed column goes in without a copy
    cell_at = cells.iloc.__getitem__ if hasattr(cells, "iloc") else cells.__getitem__
    try:
        STATS["native_rows"] = int((ing.status == native.ROW_OK).sum())
        if ing.n_obj:
            _, valid, arg = KERNELS.bbox(ing.poly_off, ing.xy)
        else:
            valid = np.zeros(0, np.uint8); arg = np.zeros(0, np.int32)
        out_bytes, out_off = ing.egress_ptlist(arg, valid)
        if ing.n and ing.arrow_input and bool((ing.status == native.ROW_OK).all()):
            wh = ing.int_columns()
            if wh is not None:                         # every row native, width / height plain ints: no per-row Python
                STATS["hostlane_objects"] = 0; STATS["slow_rows"] = 0
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


def replace_ptlist_df(df: pd.DataFrame):
    """DataFrame core of step 4 -> (result frame with the reference's column subset, excluded frame)."""
    kept = df.dropna(subset=[COL_ANN]).copy()
    excluded = df[df[COL_ANN].isna()].copy()
    new, w, h = replace_ptlist_cells(kept[COL_ANN])
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
        df = _read_csv(input_csv_path, encoding="utf-8-sig")
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
    res, excluded = replace_ptlist_df(df)
    Path(output_csv_path).parent.mkdir(parents=True, exist_ok=True)
    _to_csv(res, output_csv_path, "utf-8-sig")
    if excluded_output_file is not None:
        Path(excluded_output_file).parent.mkdir(parents=True, exist_ok=True)
        _to_csv(excluded, excluded_output_file, "utf-8-sig")
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
        ing = native.Ingest(cells, 1).boxes()
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


def filter_by_box_count_and_iou(
        input_csv_path,
        high_iou_csv="high_iou_0.98.csv",
        other_csv="other_data.csv",
        min_boxes: int = 2,
        iou_threshold: float = 0.98
):
    try:
        df = _read_csv(input_csv_path, encoding="utf-8-sig")
    except Exception as e:
        print(f"读取失败：{e}")
        return
    if COL_NEW not in df.columns:
        print(f"错误：缺少必要列 {COL_NEW}")
        return
    hi, ot = filter_by_box_count_and_iou_df(df, min_boxes, iou_threshold)
    Path(high_iou_csv).parent.mkdir(parents=True, exist_ok=True)
    Path(other_csv).parent.mkdir(parents=True, exist_ok=True)
    _to_csv(hi, high_iou_csv, "utf-8-sig")
    _to_csv(ot, other_csv, "utf-8-sig")


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
