"""Row-level CPU port of the reference's six hot-path steps -- TEST INFRASTRUCTURE ONLY.

The reference (/root/reference/src/deal_yolo_data/core/processor.py) is pure
Python and cannot travel to the GPU box, so this module restates its step
semantics at the DataFrame level.  It serves two purposes:

  * the checker for the drop-in layer's DataFrame / CSV outputs
    (tests compare it with the fixtures the real reference produced, then use it
    on fresh random tables);
  * the thing ``bench.py --impl reference`` and the ``cpu_baseline`` leg time on
    the GPU box's host cores (kind "port").

It deliberately keeps the reference's cost structure (one ``json.loads`` per row
and step, interpreter loops, builtin ``min``/``max``), because that *is* the
reference's CPU path.  Where the arithmetic lives in pandas (pinned 2.3.3 in the
reference's uv.lock; 3.0.2 in this image) -- ``drop_duplicates``, ``isin``,
``sample`` -- the port calls the same pandas entry points.

Each function cites the reference lines it follows.
"""
from __future__ import annotations

import copy
import json
import re

import pandas as pd

COL_SRC = "source"
COL_ANN = "结果字段-目标检测标签配置"
COL_NEW = "新_结果字段-目标检测标签配置"
_SEP = re.compile(r"[,，;；|]")


# ---- step 2: processor.py:140-144 -------------------------------------------
def dedup_df(df: pd.DataFrame, keep="first") -> pd.DataFrame:
    return df.drop_duplicates(subset=[COL_SRC], keep=keep, ignore_index=True)


# ---- step 3: processor.py:194-199 -------------------------------------------
def ref_filter_df(main: pd.DataFrame, ref: pd.DataFrame, col: str = COL_SRC) -> pd.DataFrame:
    seen = set(ref[col].dropna().astype(str))
    hit = main[col].astype(str).isin(seen)
    return main[~hit].copy()


# ---- step 4: processor.py:252-296 -------------------------------------------
def _corners(ptlist):
    """get_bbox_points (:252-260)."""
    good = [p for p in ptlist if isinstance(p, dict) and "x" in p and "y" in p]
    if not good:
        return [{"x": None, "y": None}, {"x": None, "y": None}]
    xs = [p["x"] for p in good]
    ys = [p["y"] for p in good]
    return [{"x": min(xs), "y": min(ys)}, {"x": max(xs), "y": max(ys)}]


def replace_cell(text):
    """parse_and_replace_ptlist (:262-281): JSON text -> JSON text with every
    dict object's ptList replaced by its two corner points; non-dict objects are
    dropped; undecodable JSON gives None."""
    if not isinstance(text, str):
        return None
    try:
        doc = json.loads(text)
    except json.JSONDecodeError:
        return None
    out = []
    for obj in doc.get("objects", []):
        if not isinstance(obj, dict):
            continue
        new = obj.copy()
        pts = _corners(obj.get("polygon", {}).get("ptList", []))
        if "polygon" not in new:
            new["polygon"] = {}
        new["polygon"]["ptList"] = pts       # shared inner dict, as in the reference
        out.append(new)
    doc["objects"] = out
    return json.dumps(doc, ensure_ascii=False)


def width_height_cell(text):
    """extract_width_height (:285-292)."""
    if not isinstance(text, str):
        return None, None
    try:
        doc = json.loads(text)
        return doc.get("width"), doc.get("height")
    except Exception:
        return None, None


def replace_ptlist_df(df: pd.DataFrame):
    """:249-250, 283, 294-306 -> (result frame with the reference's column subset,
    excluded frame)."""
    kept = df.dropna(subset=[COL_ANN]).copy()
    excluded = df[df[COL_ANN].isna()].copy()
    kept[COL_NEW] = kept[COL_ANN].apply(replace_cell)
    wh = [width_height_cell(t) for t in kept[COL_ANN]]
    kept["width"] = [w for w, _ in wh]
    kept["height"] = [h for _, h in wh]
    cols = [c for c in (COL_SRC, COL_ANN, COL_NEW, "width", "height") if c in kept.columns]
    return kept[cols], excluded


# ---- step 5: processor.py:328-376 -------------------------------------------
def _iou(a, b):
    """calculate_iou (:328-339)."""
    ix1 = max(a[0], b[0]); iy1 = max(a[1], b[1])
    ix2 = min(a[2], b[2]); iy2 = min(a[3], b[3])
    inter = max(0, ix2 - ix1) * max(0, iy2 - iy1)
    if inter == 0:
        return 0.0
    union = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    return inter / union if union != 0 else 0.0


def boxes_of_cell(text):
    """extract_boxes (:341-366); any exception ends the scan and keeps the prefix."""
    found = []
    try:
        if not isinstance(text, str):
            return found
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
            found.append((min(p["x"], q["x"]), min(p["y"], q["y"]),
                          max(p["x"], q["x"]), max(p["y"], q["y"])))
    except Exception:
        pass
    return found


def is_high_iou(boxes, min_boxes, thr):
    """meet_conditions (:368-376)."""
    n = len(boxes)
    if n < min_boxes:
        return False
    return any(_iou(boxes[i], boxes[j]) >= thr for i in range(n) for j in range(i + 1, n))


def iou_split_df(df: pd.DataFrame, min_boxes=2, thr=0.98):
    """:392-407 -> (high_iou frame, other frame); all columns, original order."""
    flags = [is_high_iou(boxes_of_cell(t), min_boxes, thr) for t in df[COL_NEW]]
    m = pd.Series(flags, index=df.index, dtype=bool)
    return df[m], df[~m]


# ---- step 5.5: processor.py:533-652, utils.py:659-679 -------------------------
def split_labels(raw):
    """_split_object_labels (utils.py:659-662)."""
    if not raw:
        return []
    return [t.strip() for t in _SEP.split(str(raw)) if t.strip()]


def rewrite_name(raw, label_map):
    """_replace_label_tokens (utils.py:664-679) -> (new_name, n_replaced, n_tokens)."""
    if not raw:
        return raw, 0, 0
    toks = split_labels(raw)
    mapped = [label_map.get(t, t) for t in toks]
    nrep = sum(1 for t in toks if t in label_map)
    return ",".join(sorted(set(mapped))), nrep, len(toks)


def mapping_from_frame(mapping_df: pd.DataFrame, old_col=None, new_col=None):
    """:533-545."""
    if not old_col or not new_col:
        cols = list(mapping_df.columns)
        if len(cols) < 2:
            raise ValueError("标签对照表至少需要两列")
        old_col = old_col or cols[0]
        new_col = new_col or cols[1]
    out = {}
    for _, r in mapping_df.iterrows():
        a = str(r.get(old_col, "")).strip()
        b = str(r.get(new_col, "")).strip()
        if a and a.lower() != "nan" and b and b.lower() != "nan":
            out[a] = b
    return out


def default_json_columns(df):
    return [c for c in (COL_NEW, COL_ANN) if c in df.columns]


def remap_df(df: pd.DataFrame, label_map: dict, json_columns=None):
    """:554-613 -> (rewritten frame, summary dict, diff rows, unmatched counter)."""
    df = df.copy()
    if json_columns is None:
        json_columns = default_json_columns(df)
    s = dict(total_rows=len(df), replaced_rows=0, total_objects=0, replaced_objects=0,
             total_labels=0, replaced_labels=0, invalid_json_rows=0, missing_name_objects=0)
    unmatched = {}
    diffs = []
    for idx, row in df.iterrows():
        touched = False
        for col in json_columns:
            if col not in df.columns:
                continue
            text = row.get(col)
            if not isinstance(text, str) or not text:
                continue
            try:
                doc = json.loads(text)
            except json.JSONDecodeError:
                s["invalid_json_rows"] += 1
                continue
            objs = doc.get("objects")
            if not isinstance(objs, list):
                continue
            pairs = []
            for obj in objs:
                if not isinstance(obj, dict):
                    continue
                s["total_objects"] += 1
                raw = obj.get("name")
                if raw is None:
                    s["missing_name_objects"] += 1
                    continue
                for t in split_labels(raw):
                    if t not in label_map:
                        unmatched[t] = unmatched.get(t, 0) + 1
                new, nrep, ntok = rewrite_name(raw, label_map)
                s["total_labels"] += ntok
                if nrep > 0:
                    obj["name"] = new
                    s["replaced_labels"] += nrep
                    s["replaced_objects"] += 1
                    touched = True
                if raw != new:
                    pairs.append((raw, new))
            doc["objects"] = objs
            df.at[idx, col] = json.dumps(doc, ensure_ascii=False)
            if pairs:
                diffs.append({"source": row.get("source"), "column": col,
                              "before": "；".join(p[0] for p in pairs),
                              "after": "；".join(p[1] for p in pairs)})
        if touched:
            s["replaced_rows"] += 1
    s["mapping_size"] = len(label_map)
    s["unmatched_labels"] = len(unmatched)
    return df, s, diffs, unmatched


# ---- step 6: processor.py:673-806, utils.py:635-657 ---------------------------
def split_label_cell(cell):
    """_split_label_cell (utils.py:635-643)."""
    if pd.isna(cell):
        return []
    text = str(cell).strip()
    if not text:
        return []
    return [t.strip() for t in _SEP.split(text) if t.strip()]


def rules_from_frame(rules_df, rule_mode="wide", label_col=None, category_col=None):
    """:690-703."""
    l2c = {}
    if rule_mode == "wide":
        for col in rules_df.columns:
            cat = str(col).strip()
            if not cat:
                continue
            for cell in rules_df[col].dropna():
                for lab in split_label_cell(cell):
                    l2c[lab] = cat
    elif rule_mode == "two_column":
        for _, r in rules_df.iterrows():
            lab = str(r.get(label_col, "")).strip()
            cat = str(r.get(category_col, "")).strip()
            if lab and cat and lab.lower() != "nan" and cat.lower() != "nan":
                l2c[lab] = cat
    return l2c


def parse_objects(text):
    """_parse_data_objects (utils.py:645-657)."""
    if not isinstance(text, str) or not text:
        return None, [], "空数据"
    try:
        doc = json.loads(text)
        objs = doc.get("objects", [])
        if not isinstance(objs, list):
            return doc, [], "objects不是列表"
        return doc, objs, None
    except json.JSONDecodeError:
        return None, [], "JSON解析失败"
    except Exception as e:  # noqa: BLE001 - the reference reports str(e)
        return None, [], str(e)


def split_df(df: pd.DataFrame, l2c: dict, json_columns=None,
             train_ratio=0.8, val_ratio=0.1, test_ratio=0.1, random_seed=42):
    """:673-676, 711-806 -> dict with per-category {train,val,test} frames, the
    unclassified frame, the split_counts frame and the summary."""
    tot = train_ratio + val_ratio + test_ratio
    train_ratio /= tot; val_ratio /= tot; test_ratio /= tot
    if json_columns is None:
        json_columns = default_json_columns(df)
    by_cat, unclassified, counts = {}, [], []
    for _, row in df.iterrows():
        text = None
        for col in json_columns:
            if col in row and isinstance(row[col], str) and row[col]:
                text = row[col]
                break
        doc, objs, err = parse_objects(text)
        if err or not objs:
            why = err or "标注字段objects为空"
            r = row.copy(); r["无法分类原因"] = why
            unclassified.append(r)
            counts.append({"source": row.get("source"), "原始标签组合": "", "拆分条数": 0,
                           "是否可分类": "否", "无法分类原因": why})
            continue
        labset = set()
        for obj in objs:
            if isinstance(obj, dict) and obj.get("name"):
                labset.update(split_labels(obj.get("name")))
        combo = "，".join(sorted(labset)) if labset else ""
        n_exp, reasons, any_ok = 0, set(), False
        for obj in objs:
            if not isinstance(obj, dict):
                continue
            labs = split_labels(obj.get("name"))
            if not labs:
                r = row.copy(); r["无法分类原因"] = "标注框缺少name字段"
                unclassified.append(r)
                continue
            for lab in labs:
                if lab not in l2c:
                    r = row.copy()
                    r["无法分类原因"] = f"标签{lab}未在规则中定义"
                    r["无法分类标签"] = lab
                    unclassified.append(r)
                    reasons.add(f"标签{lab}未在规则中定义")
                    continue
                cat = l2c[lab]
                r = row.copy()
                one = copy.deepcopy(obj); one["name"] = lab
                nd = {k: v for k, v in doc.items() if k != "objects"}
                nd["objects"] = [one]
                js = json.dumps(nd, ensure_ascii=False)
                for col in json_columns:
                    if col in df.columns:
                        r[col] = js
                r["分类标签"] = lab; r["分类类别"] = cat; r["原始标签组合"] = combo
                by_cat.setdefault(cat, []).append(r)
                any_ok = True; n_exp += 1
        if not any_ok:
            r = row.copy()
            r["无法分类原因"] = "；".join(sorted(reasons)) if reasons else "标签无法匹配规则"
            unclassified.append(r)
        status = "否" if not any_ok else ("部分可分类" if reasons else "是")
        counts.append({"source": row.get("source"), "原始标签组合": combo, "拆分条数": n_exp,
                       "是否可分类": status, "无法分类原因": "；".join(sorted(reasons))})
    cats, cat_counts = {}, {}
    for cat, rows in by_cat.items():
        if not rows:
            continue
        cat_counts[cat] = len(rows)
        cdf = pd.DataFrame(rows).sample(frac=1, random_state=random_seed).reset_index(drop=True)
        n = len(cdf); ntr = int(n * train_ratio); nva = int(n * val_ratio)
        cats[cat] = {"train": cdf.iloc[:ntr], "val": cdf.iloc[ntr:ntr + nva], "test": cdf.iloc[ntr + nva:]}
    return {
        "categories": cats,
        "unclassified": pd.DataFrame(unclassified),
        "split_counts": pd.DataFrame(counts),
        "summary": {"categories": len(by_cat), "classified": sum(cat_counts.values()),
                    "unclassified": len(unclassified), "category_counts": cat_counts},
    }


# ---- whole-path runner used by bench.py's reference arm -----------------------
def run_hot_path(df: pd.DataFrame, ref: pd.DataFrame | None, min_boxes=2, thr=0.7):
    """dedup -> (ref filter) -> ptList->bbox -> IoU filter on an in-memory frame.
    Returns the 'other' frame the pipeline continues with (processing.py:598-625)."""
    d = dedup_df(df)
    if ref is not None:
        d = ref_filter_df(d, ref)
    r, _ = replace_ptlist_df(d)
    _, other = iou_split_df(r, min_boxes, thr)
    return other


def run_hot_path_files(workdir, merged_csv, ref_csv=None, min_boxes=2, thr=0.7):
    """The same chain through CSV files, the way the reference's step functions hand data to each
    other (read_csv -> step -> to_csv, utf-8-sig, index=False; processor.py:125/158, 181/213,
    235/309, 379/404-407).  Returns the number of rows the pipeline continues with."""
    from pathlib import Path
    w = Path(workdir)
    enc = "utf-8-sig"
    d = dedup_df(pd.read_csv(merged_csv, encoding=enc, parse_dates=False))
    d.to_csv(w / "deduplicate_result.csv", index=False, encoding=enc)
    cur = w / "deduplicate_result.csv"
    if ref_csv is not None:
        f = ref_filter_df(pd.read_csv(cur, encoding=enc, parse_dates=False), pd.read_csv(ref_csv, encoding=enc, parse_dates=False))
        f.to_csv(w / "filtered_main.csv", index=False, encoding=enc)
        cur = w / "filtered_main.csv"
    r, exc = replace_ptlist_df(pd.read_csv(cur, encoding=enc))
    r.to_csv(w / "processed_replaced_ptlist.csv", index=False, encoding=enc)
    exc.to_csv(w / "processed_excluded.csv", index=False, encoding=enc)
    hi, ot = iou_split_df(pd.read_csv(w / "processed_replaced_ptlist.csv", encoding=enc), min_boxes, thr)
    hi.to_csv(w / f"high_iou_{thr:.2f}.csv", index=False, encoding=enc)
    ot.to_csv(w / "other_data.csv", index=False, encoding=enc)
    return len(ot)
