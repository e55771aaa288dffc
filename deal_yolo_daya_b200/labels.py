"""Label remap (step 5.5) and rule-based split (step 6): DataFrame cores.

Strings stay on the host: names are dictionary-encoded in the reference's traversal order
(row, JSON column, object), the per-vocabulary tables are built with the reference's string
rules (utils.py:635-679) and every per-object / per-row decision -- which name is rewritten,
all counters, the per-label histogram, the category grouping and the split ids -- is taken by
the K3 / K6 kernels on the encoded ids.  The host then re-serialises the documents exactly as
the reference does (``json.dumps(..., ensure_ascii=False)`` of the mutated parsed document).

reference: processor.py:516-652 (remap), :654-831 (split); utils.py:635-679 (string rules).
"""
from __future__ import annotations

import json
import re

import numpy as np
import pandas as pd

COL_ANN = "结果字段-目标检测标签配置"
COL_NEW = "新_结果字段-目标检测标签配置"
_SEP = re.compile(r"[,，;；|]")
LAST = {"remap_lane": None, "split_lane": None}          # observability: which lane the last call took


def _kernels():
    from . import processor
    return processor.KERNELS


def split_labels(raw):
    """_split_object_labels (utils.py:659-662)."""
    if not raw:
        return []
    return [t.strip() for t in _SEP.split(str(raw)) if t.strip()]


def split_label_cell(cell):
    """_split_label_cell (utils.py:635-643)."""
    if pd.isna(cell):
        return []
    text = str(cell).strip()
    if not text:
        return []
    return [t.strip() for t in _SEP.split(text) if t.strip()]


def default_json_columns(df):
    return [c for c in (COL_NEW, COL_ANN) if c in df.columns]


def mapping_from_frame(mapping_df: pd.DataFrame, old_col=None, new_col=None) -> dict:
    """Label map of processor.py:533-545."""
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


def rules_from_frame(rules_df: pd.DataFrame, rule_mode="wide", label_col=None, category_col=None) -> dict:
    """label -> category of processor.py:690-703."""
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


# =============================================================================================
# remap
# =============================================================================================
class _Vocab:
    def __init__(self):
        self.ids, self.names = {}, []

    def get(self, name: str) -> int:
        i = self.ids.get(name)
        if i is None:
            i = len(self.names)
            self.ids[name] = i
            self.names.append(name)
        return i


def _remap_tables(vocab, label_map):
    """Per-vocabulary tables with the reference's string rules (utils.py:659-679): tokens, normalised
    new name, and the three LUTs K3 reads (new id, token count, replaced-token count)."""
    n_raw = len(vocab.names)
    toks_of, norm_of = [], []
    for v in range(n_raw):
        raw = vocab.names[v]
        toks = split_labels(raw)
        toks_of.append(toks)
        norm_of.append(raw if not raw else ",".join(sorted(set(label_map.get(t, t) for t in toks))))
    lut_ntok = np.array([0 if not vocab.names[v] else len(toks_of[v]) for v in range(n_raw)], np.int32)
    lut_nrep = np.array([0 if not vocab.names[v] else sum(1 for t in toks_of[v] if t in label_map) for v in range(n_raw)], np.int32)
    lut_new = np.array([vocab.get(norm_of[v]) for v in range(n_raw)], np.int32)       # may append new names
    n_vocab = len(vocab.names)
    pad = n_vocab - n_raw
    lut_new = np.concatenate([lut_new, np.arange(n_raw, n_vocab, dtype=np.int32)])
    lut_ntok = np.concatenate([lut_ntok, np.zeros(pad, np.int32)])
    lut_nrep = np.concatenate([lut_nrep, np.zeros(pad, np.int32)])
    return n_raw, n_vocab, toks_of, norm_of, lut_new, lut_ntok, lut_nrep


def _unmatched_from_hist(vocab, n_raw, toks_of, hist, label_map):
    """Unmatched labels in first-encounter order with their object counts (processor.py:591-593)."""
    unmatched = {}
    for v in range(n_raw):
        h = int(hist[v]) if v < len(hist) else 0
        if h == 0 or not vocab.names[v]:
            continue
        for t in toks_of[v]:
            if t not in label_map:
                unmatched[t] = unmatched.get(t, 0) + h
    return unmatched


def _remap_native(df, label_map, cols):
    """Native lane of remap_df: every text cell of the JSON columns is already in json.dumps form (what
    step 4 / step 5 wrote), so the names are taken from the text by csrc/ingest.cpp (mode 2), K3 decides
    on the dictionary-encoded ids and the new texts are the old ones with names spliced.  Returns None
    when any cell needs CPython's json (the caller then runs the Python lane for the whole frame: the
    vocabulary order depends on every cell)."""
    from . import native
    n = len(df)
    if not native.enabled() or not cols or n == 0 or not native._pandas_infers_arrow_str():
        return None
    if any(getattr(df[c].array, "_pa_array", None) is None for c in cols):
        return None
    import pyarrow as pa
    import pyarrow.compute as pc
    ings = []
    try:
        for c in cols:
            ing = native.Ingest(df[c], 2)
            ings.append(ing)
            if ing.n_slow:
                return None
            ing.names()
        ncols = len(cols)
        counts = np.stack([np.diff(i.cell_off) for i in ings], axis=1)              # objects per (row, column) cell
        cell_off = np.zeros(n * ncols + 1, np.int64)
        np.cumsum(counts.reshape(-1), out=cell_off[1:])
        total = int(cell_off[-1])
        dests = []                                                                   # column-major object -> traversal position
        for ci, ing in enumerate(ings):
            dst_start = cell_off[np.arange(n, dtype=np.int64) * ncols + ci]
            dests.append(np.repeat(dst_start - ing.cell_off[:-1], counts[:, ci]) + np.arange(ing.n_obj, dtype=np.int64))
        dest = np.concatenate(dests) if dests else np.zeros(0, np.int64)
        perm = np.empty(total, np.int64); perm[dest] = np.arange(total, dtype=np.int64)
        if total:
            enc = pc.dictionary_encode(pa.concat_arrays([i.name_array() for i in ings]).take(pa.array(perm)))
            ids = np.asarray(enc.indices.fill_null(-1)).astype(np.int32)
            raw_names = enc.dictionary.to_pylist()
        else:
            ids = np.zeros(0, np.int32); raw_names = []
        vocab = _Vocab()
        for nm in raw_names:
            vocab.get(nm)
        n_raw, n_vocab, toks_of, norm_of, lut_new, lut_ntok, lut_nrep = _remap_tables(vocab, label_map)
        if total:
            new_id, cell_rep, cnt, hist = _kernels().label_lut(cell_off, ids, lut_new, lut_ntok, lut_nrep)
        else:
            new_id = np.zeros(0, np.int32); cell_rep = np.zeros(n * ncols, np.uint8); hist = np.zeros(n_vocab, np.uint64)
            cnt = dict(total_objects=0, missing_name_objects=0, total_labels=0, replaced_labels=0, replaced_objects=0, replaced_rows=0)
        unmatched = _unmatched_from_hist(vocab, n_raw, toks_of, hist, label_map)
        # ---- egress: splice the new names, one pass per column ----
        safe = np.where(ids >= 0, ids, n_vocab)                                     # objects without a name index a padding slot
        flags = (np.append(lut_nrep, 0)[safe] > 0).astype(np.uint8) if total else np.zeros(0, np.uint8)
        esc = [json.dumps(nm, ensure_ascii=False)[1:-1].encode("utf-8") for nm in vocab.names]
        v_off = np.zeros(len(esc) + 1, np.int64)
        if esc:
            np.cumsum([len(e) for e in esc], out=v_off[1:])
        v_bytes = np.frombuffer(b"".join(esc) or b"\0", dtype=np.uint8)
        for ci, (c, ing) in enumerate(zip(cols, ings)):
            out, out_off = ing.egress_names(flags[dests[ci]], np.asarray(new_id)[dests[ci]], v_bytes, v_off)
            ok = ing.status == native.ROW_OK
            if ok.all():
                df[c] = pd.Series(native.arrow_strings(out, out_off), index=df.index)
            elif ok.any():
                vals = df[c].tolist()
                blob = out.tobytes()
                for r in np.nonzero(ok)[0]:
                    vals[r] = blob[out_off[r]:out_off[r + 1]].decode("utf-8")
                df[c] = pd.Series(vals, index=df.index, dtype=df[c].dtype)
        # ---- diff rows in traversal order ----
        diff_rows = []
        if total:
            changed_v = np.array([vocab.names[v] != norm_of[v] for v in range(n_raw)] + [False] * (n_vocab - n_raw + 1), bool)
            obj_changed = changed_v[safe]
            cell_of_obj = np.repeat(np.arange(n * ncols, dtype=np.int64), counts.reshape(-1))
            has = np.bincount(cell_of_obj[obj_changed], minlength=n * ncols) > 0
            sources = df["source"].tolist() if "source" in df.columns else [None] * n
            names = vocab.names
            for ci in np.nonzero(has)[0]:
                a, b = int(cell_off[ci]), int(cell_off[ci + 1])
                sel = ids[a:b][obj_changed[a:b]]
                diff_rows.append({"source": sources[ci // ncols], "column": cols[ci % ncols],
                                  "before": "；".join(names[v] for v in sel), "after": "；".join(norm_of[v] for v in sel)})
        row_touched = np.asarray(cell_rep).reshape(n, ncols).any(axis=1) if n else np.zeros(0, bool)
    finally:
        for ing in ings:
            ing.close()
    summary = {
        "total_rows": n, "replaced_rows": int(row_touched.sum()), "total_objects": cnt["total_objects"],
        "replaced_objects": cnt["replaced_objects"], "total_labels": cnt["total_labels"],
        "replaced_labels": cnt["replaced_labels"], "invalid_json_rows": 0,
        "missing_name_objects": cnt["missing_name_objects"], "mapping_size": len(label_map),
        "unmatched_labels": len(unmatched),
    }
    return df, summary, diff_rows, unmatched


def remap_df(df: pd.DataFrame, label_map: dict, json_columns=None):
    """DataFrame core of replace_labels_by_mapping -> (frame, summary, diff_rows, unmatched counter)."""
    df = df.copy()
    if json_columns is None:
        json_columns = default_json_columns(df)
    cols = [c for c in json_columns if c in df.columns]
    fast = _remap_native(df, label_map, cols)
    LAST["remap_lane"] = "native" if fast is not None else "python"
    if fast is not None:
        return fast
    n_rows = len(df)
    invalid_json_rows = 0
    # ---- ingest: decode cells, dictionary-encode names in reference traversal order ----
    cells = []          # (row position, column, doc, objects) for every re-serialised cell
    cell_cnt, label_ids = [], []
    exotic = []         # (cell index, object index) whose name is neither str nor None
    vocab = _Vocab()
    col_values = {c: df[c].tolist() for c in cols}
    for r in range(n_rows):
        for c in cols:
            text = col_values[c][r]
            if not isinstance(text, str) or not text:
                continue
            try:
                doc = json.loads(text)
            except json.JSONDecodeError:
                invalid_json_rows += 1
                continue
            objs = doc.get("objects")
            if not isinstance(objs, list):
                continue
            n = 0
            for k, obj in enumerate(objs):
                if not isinstance(obj, dict):
                    continue
                name = obj.get("name")
                if name is None:
                    label_ids.append(-1)
                elif isinstance(name, str):
                    label_ids.append(vocab.get(name))
                else:
                    label_ids.append(-2); exotic.append((len(cells), k))
                n += 1
            cells.append((r, c, doc, objs)); cell_cnt.append(n)
    # ---- per-vocabulary tables (string rules of utils.py:659-679) ----
    n_raw, n_vocab, toks_of, norm_of, lut_new, lut_ntok, lut_nrep = _remap_tables(vocab, label_map)
    # ---- device: rewrite decisions, counters, histogram ----
    cell_off = np.zeros(len(cells) + 1, np.int64)
    np.cumsum(np.array(cell_cnt, np.int64), out=cell_off[1:])
    ids = np.array(label_ids, np.int32) if label_ids else np.zeros(0, np.int32)
    dev_ids = np.where(ids == -2, -1, ids).astype(np.int32)       # exotic names are settled on the host lane below
    if len(cells):
        new_id, cell_rep, cnt, hist = _kernels().label_lut(cell_off, dev_ids, lut_new, lut_ntok, lut_nrep)
    else:
        new_id = np.zeros(0, np.int32); cell_rep = np.zeros(0, np.uint8); hist = np.zeros(n_vocab, np.uint64)
        cnt = dict(total_objects=0, missing_name_objects=0, total_labels=0, replaced_labels=0, replaced_objects=0, replaced_rows=0)
    cnt["missing_name_objects"] -= len(exotic)                    # they were sent as "no name"; corrected here
    # ---- unmatched labels in first-encounter order (processor.py:591-593) ----
    unmatched = _unmatched_from_hist(vocab, n_raw, toks_of, hist, label_map)
    # ---- host lane: names that are not strings (numbers, lists ...) follow CPython semantics ----
    row_touched = np.zeros(n_rows, bool)
    exotic_set = {}
    for ci, k in exotic:
        exotic_set.setdefault(ci, set()).add(k)
    # ---- egress: rewrite names, re-serialise, collect diff rows ----
    diff_rows = []
    touched_cols = set()
    sources = df["source"].tolist() if "source" in df.columns else [None] * n_rows
    for ci, (r, c, doc, objs) in enumerate(cells):
        q = int(cell_off[ci])
        pairs = []
        for k, obj in enumerate(objs):
            if not isinstance(obj, dict):
                continue
            v = int(ids[q])
            if v >= 0:
                if lut_nrep[v] > 0:
                    obj["name"] = vocab.names[int(new_id[q])]
                if vocab.names[v] != norm_of[v]:
                    pairs.append((vocab.names[v], norm_of[v]))
            elif v == -2:
                raw = obj.get("name")
                for t in split_labels(raw):
                    if t not in label_map:
                        unmatched[t] = unmatched.get(t, 0) + 1
                if raw:
                    toks = split_labels(raw)
                    new = ",".join(sorted(set(label_map.get(t, t) for t in toks)))
                    nrep = sum(1 for t in toks if t in label_map)
                    cnt["total_labels"] += len(toks)
                else:
                    new, nrep = raw, 0
                if nrep > 0:
                    obj["name"] = new
                    cnt["replaced_labels"] += nrep; cnt["replaced_objects"] += 1
                    row_touched[r] = True
                if raw != new:
                    pairs.append((raw, new))
            q += 1
        if cell_rep[ci]:
            row_touched[r] = True
        doc["objects"] = objs
        col_values[c][r] = json.dumps(doc, ensure_ascii=False)     # columns are assigned once below: a per-cell
        touched_cols.add(c)                                         # setitem on an Arrow string column copies the column
        if pairs:
            diff_rows.append({"source": sources[r], "column": c,
                              "before": "；".join(p[0] for p in pairs), "after": "；".join(p[1] for p in pairs)})
    for c in cols:
        if c in touched_cols:
            df[c] = pd.Series(col_values[c], index=df.index, dtype=df[c].dtype)
    summary = {
        "total_rows": n_rows, "replaced_rows": int(row_touched.sum()), "total_objects": cnt["total_objects"],
        "replaced_objects": cnt["replaced_objects"], "total_labels": cnt["total_labels"],
        "replaced_labels": cnt["replaced_labels"], "invalid_json_rows": invalid_json_rows,
        "missing_name_objects": cnt["missing_name_objects"], "mapping_size": len(label_map),
        "unmatched_labels": len(unmatched),
    }
    return df, summary, diff_rows, unmatched


# =============================================================================================
# split
# =============================================================================================
def _parse_objects(text):
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


def _split_native(df, l2c, json_columns, train_ratio, val_ratio, random_seed):
    """Native lane of split_df: the JSON column every row uses holds json.dumps-form text, so the objects'
    names and spans come from csrc/ingest.cpp (mode 2), the (object, label) entries, label sets and the
    unclassified / split_counts bookkeeping are built on arrays, K6 groups and assigns, and the one-object
    cells are spliced natively (dyd_egress_split).  Returns None when a row needs CPython's json or picks a
    different JSON column: the caller then runs the Python lane for the whole frame."""
    from . import native
    n = len(df)
    if not native.enabled() or n == 0 or not native._pandas_infers_arrow_str():
        return None
    c0 = next((c for c in json_columns if c in df.columns), None)
    if c0 is None:
        return None
    pa_arr = getattr(df[c0].array, "_pa_array", None)
    if pa_arr is None or pa_arr.null_count:
        return None
    import pyarrow.compute as pc
    if not pc.all(pc.greater(pc.binary_length(pa_arr), 0)).as_py():
        return None                                            # some row would fall through to the next JSON column
    ing = native.Ingest(df[c0], 2)
    try:
        if ing.n_slow:
            return None
        ing.names().objects()
        st = ing.status
        cnt_obj = np.diff(ing.cell_off)
        n_obj = ing.n_obj
        obj_row = np.repeat(np.arange(n, dtype=np.int64), cnt_obj)
        # ---- names -> label tokens (string rules on the distinct names only) ----
        if n_obj:
            enc = pc.dictionary_encode(ing.name_array())
            ids = np.asarray(enc.indices.fill_null(-1)).astype(np.int64)
            uniq = enc.dictionary.to_pylist()
        else:
            ids = np.zeros(0, np.int64); uniq = []
        tok_ids, tok_names, flat, name_off = {}, [], [], [0]
        for nm in uniq:
            for t in split_labels(nm):
                i = tok_ids.get(t)
                if i is None:
                    i = tok_ids[t] = len(tok_names); tok_names.append(t)
                flat.append(i)
            name_off.append(len(flat))
        n_tok = len(tok_names)
        name_off = np.array(name_off, np.int64); flat = np.array(flat, np.int64)
        ntok_name = np.diff(name_off)
        ntok_obj = np.where(ids >= 0, np.append(ntok_name, 0)[np.where(ids >= 0, ids, len(ntok_name))], 0).astype(np.int64)
        n_ent = int(ntok_obj.sum())
        ent_obj = np.repeat(np.arange(n_obj, dtype=np.int64), ntok_obj)
        ent_first = np.cumsum(ntok_obj) - ntok_obj
        ent_tok = flat[name_off[ids[ent_obj]] + (np.arange(n_ent, dtype=np.int64) - ent_first[ent_obj])] if n_ent else np.zeros(0, np.int64)
        ent_row = obj_row[ent_obj]
        row_off = np.zeros(n + 1, np.int64)
        np.cumsum(np.bincount(ent_row, minlength=n), out=row_off[1:])
        # ---- categories in first-encounter order over the entries ----
        tok_cat = [l2c.get(t) for t in tok_names]
        known_tok = np.array([c is not None for c in tok_cat], bool) if n_tok else np.zeros(0, bool)
        first_of_tok = np.full(n_tok, n_ent, np.int64)
        if n_ent:
            np.minimum.at(first_of_tok, ent_tok, np.arange(n_ent, dtype=np.int64))
        cat_first = {}
        for t in range(n_tok):
            if tok_cat[t] is not None:
                cat_first[tok_cat[t]] = min(cat_first.get(tok_cat[t], n_ent), int(first_of_tok[t]))
        cat_names = sorted(cat_first, key=cat_first.get)
        cat_ids = {c: i for i, c in enumerate(cat_names)}
        n_cat = len(cat_names)
        cat_of_label = np.array([cat_ids[c] if c is not None else -1 for c in tok_cat], np.int32) if n_tok else np.zeros(0, np.int32)
        # ---- label-set string of every row: sorted distinct tokens joined by the full-width comma ----
        combo = [""] * n
        if n_ent:
            by_name = sorted(range(n_tok), key=tok_names.__getitem__)
            rank = np.empty(n_tok, np.int64); rank[by_name] = np.arange(n_tok)
            names_by_rank = [tok_names[t] for t in by_name]
            uk = np.unique(ent_row * n_tok + rank[ent_tok])
            urow, urank = uk // n_tok, uk % n_tok
            cut = np.searchsorted(urow, np.arange(n + 1))
            for r in np.nonzero(np.diff(cut))[0]:
                combo[r] = "，".join(names_by_rank[k] for k in urank[cut[r]:cut[r + 1]])
        # ---- device: stable grouping by category, split ids from the host permutation ----
        labels = ent_tok.astype(np.int32)
        if n_cat and n_ent:
            exp_row, exp_entry, exp_cat, cat_off = _kernels().split_expand(row_off, labels, cat_of_label, n_cat)
        else:
            exp_row = np.zeros(0, np.int64); exp_entry = np.zeros(0, np.int64); cat_off = np.zeros(n_cat + 1, np.int64)
        sizes = np.diff(cat_off)
        n_train = np.array([int(k * train_ratio) for k in sizes], np.int64)
        n_val = np.array([int(k * val_ratio) for k in sizes], np.int64)
        perms = [native.permutation(random_seed, int(k)) for k in sizes]
        perm = np.concatenate(perms).astype(np.int64) if perms else np.zeros(0, np.int64)
        if len(perm):
            split_id, pos = _kernels().split_assign(cat_off, perm, n_train, n_val)
        else:
            split_id = np.zeros(0, np.uint8); pos = np.zeros(0, np.int64)
        # ---- egress: per-category frames, cells spliced natively ----
        esc = [json.dumps(t, ensure_ascii=False)[1:-1].encode("utf-8") for t in tok_names]
        t_off = np.zeros(n_tok + 1, np.int64)
        if esc:
            np.cumsum([len(e) for e in esc], out=t_off[1:])
        t_bytes = np.frombuffer(b"".join(esc) or b"\0", dtype=np.uint8)
        keep_cols = [c for c in df.columns if c not in json_columns]
        categories, cat_counts = {}, {}
        for ci, cat in enumerate(cat_names):
            a, b = int(cat_off[ci]), int(cat_off[ci + 1])
            if b == a:
                continue
            cat_counts[cat] = b - a
            order = np.empty(b - a, np.int64)
            order[pos[a:b]] = np.arange(b - a)             # shuffled position -> original expanded row
            rows_idx = np.asarray(exp_row[a:b])[order]
            ents = np.asarray(exp_entry[a:b])[order]
            out, out_off = ing.egress_split(rows_idx, ent_obj[ents], ent_tok[ents], t_bytes, t_off)
            cells = native.arrow_strings(out, out_off)
            frame = df[keep_cols].iloc[rows_idx].reset_index(drop=True)
            for c in json_columns:
                if c in df.columns:
                    frame[c] = cells
            frame = frame[list(df.columns)]
            frame["分类标签"] = [tok_names[t] for t in ent_tok[ents]]
            frame["分类类别"] = cat
            frame["原始标签组合"] = [combo[r] for r in rows_idx]
            sp = np.asarray(split_id[a:b])[order]
            categories[cat] = {"train": frame[sp == 0], "val": frame[sp == 1], "test": frame[sp == 2]}
        # ---- unclassified rows and split_counts in reference order, on arrays ----
        bad = (st != native.ROW_OK) | (ing.list_len == 0)
        bad_reason = np.where(st == native.ROW_NOT_A_LIST, "objects不是列表", "标注字段objects为空")
        ent_known = known_tok[ent_tok] if n_ent else np.zeros(0, bool)
        n_ok_row = np.bincount(ent_row[ent_known], minlength=n) if n_ent else np.zeros(n, np.int64)
        reason_tok = [f"标签{t}未在规则中定义" for t in tok_names]
        b_ent = np.nonzero(~ent_known)[0]
        row_reasons = [""] * n                                  # "；".join(sorted(set(reasons))) per row
        has_reason = np.zeros(n, bool)
        if len(b_ent):
            by_reason = sorted(range(n_tok), key=reason_tok.__getitem__)
            rrank = np.empty(n_tok, np.int64); rrank[by_reason] = np.arange(n_tok)
            reason_by_rank = [reason_tok[t] for t in by_reason]
            uk = np.unique(ent_row[b_ent] * n_tok + rrank[ent_tok[b_ent]])
            urow, urank = uk // n_tok, uk % n_tok
            cut = np.searchsorted(urow, np.arange(n + 1))
            for r in np.nonzero(np.diff(cut))[0]:
                row_reasons[r] = "；".join(reason_by_rank[k] for k in urank[cut[r]:cut[r + 1]])
                has_reason[r] = True
        a_obj = np.nonzero(ntok_obj == 0)[0]
        c_row = np.nonzero(~bad & (n_ok_row == 0))[0]
        d_row = np.nonzero(bad)[0]
        ev_row = np.concatenate([d_row, obj_row[a_obj], ent_row[b_ent], c_row])
        ev_k2 = np.concatenate([np.full(len(d_row), -1, np.int64), a_obj, ent_obj[b_ent], np.full(len(c_row), n_obj + 1, np.int64)])
        ev_k3 = np.concatenate([np.zeros(len(d_row), np.int64), np.full(len(a_obj), -1, np.int64), b_ent, np.zeros(len(c_row), np.int64)])
        ev_reason = ([str(x) for x in bad_reason[d_row]] + ["标注框缺少name字段"] * len(a_obj) + [reason_tok[t] for t in ent_tok[b_ent]] +
                     [row_reasons[r] if has_reason[r] else "标签无法匹配规则" for r in c_row])
        ev_label = [None] * (len(d_row) + len(a_obj)) + [tok_names[t] for t in ent_tok[b_ent]] + [None] * len(c_row)
        o = np.lexsort((ev_k3, ev_k2, ev_row))
        unc_rows = ev_row[o]
        unc_reason = [ev_reason[i] for i in o]
        unc_label = [ev_label[i] for i in o]
        if len(unc_rows):
            unc = df.iloc[unc_rows].reset_index(drop=True)
            unc["无法分类原因"] = unc_reason
            if len(b_ent):
                unc["无法分类标签"] = pd.Series(unc_label, dtype=object).where(pd.Series([x is not None for x in unc_label]), np.nan)
        else:
            unc = pd.DataFrame()
        sources = df["source"].tolist() if "source" in df.columns else [None] * n
        status = np.where(bad | (n_ok_row == 0), "否", np.where(has_reason, "部分可分类", "是"))
        counts = pd.DataFrame({
            "source": sources,
            "原始标签组合": [("" if bad[r] else combo[r]) for r in range(n)],
            "拆分条数": np.where(bad, 0, n_ok_row).astype(np.int64),
            "是否可分类": [str(x) for x in status],
            "无法分类原因": [(str(bad_reason[r]) if bad[r] else row_reasons[r]) for r in range(n)],
        }) if n else pd.DataFrame([])
    finally:
        ing.close()
    return {
        "categories": categories,
        "unclassified": unc,
        "split_counts": counts,
        "summary": {"categories": n_cat, "classified": int(sum(cat_counts.values())), "unclassified": int(len(unc_rows)),
                    "category_counts": cat_counts},
    }


def split_df(df: pd.DataFrame, l2c: dict, json_columns=None,
             train_ratio=0.8, val_ratio=0.1, test_ratio=0.1, random_seed=42):
    """DataFrame core of split_dataset_by_rules -> dict(categories={cat: {train,val,test}}, unclassified,
    split_counts, summary).  Category grouping (K6) and the split ids come from the device; the
    permutation is numpy's legacy MT19937 shuffle, which is what DataFrame.sample applies."""
    tot = train_ratio + val_ratio + test_ratio
    train_ratio /= tot; val_ratio /= tot; test_ratio /= tot
    if json_columns is None:
        json_columns = default_json_columns(df)
    fast = _split_native(df, l2c, list(json_columns), train_ratio, val_ratio, random_seed)
    LAST["split_lane"] = "native" if fast is not None else "python"
    if fast is not None:
        return fast
    jcols_present = [c for c in json_columns if c in df.columns]
    n_rows = len(df)
    col_values = {c: df[c].tolist() for c in jcols_present}
    sources = df["source"].tolist() if "source" in df.columns else [None] * n_rows
    # ---- ingest: (object, label) entries per row, labels dictionary-encoded ----
    vocab = _Vocab()
    cat_ids, cat_names = {}, []
    entry_cnt = np.zeros(n_rows, np.int64)
    entry_label, entry_obj = [], []          # per entry: label id / (row, object) payload
    labs_cache = {}

    def labs_of(name):                       # split_labels per distinct name string
        if type(name) is str:
            got = labs_cache.get(name)
            if got is None:
                got = labs_cache[name] = split_labels(name)
            return got
        return split_labels(name)
    row_info = [None] * n_rows               # (doc, combo) for classifiable rows
    events = []                              # unclassified rows & split_counts rows in reference order
    for r in range(n_rows):
        text = None
        for c in json_columns:
            if c in col_values and isinstance(col_values[c][r], str) and col_values[c][r]:
                text = col_values[c][r]
                break
        doc, objs, err = _parse_objects(text)
        if err or not objs:
            events.append(("bad_row", r, err or "标注字段objects为空"))
            continue
        labset = set()
        for obj in objs:
            if isinstance(obj, dict) and obj.get("name"):
                labset.update(labs_of(obj.get("name")))
        combo = "，".join(sorted(labset)) if labset else ""
        row_info[r] = (doc, combo)
        n = 0
        per_obj = []
        for obj in objs:
            if not isinstance(obj, dict):
                continue
            labs = labs_of(obj.get("name"))
            per_obj.append((obj, labs))
            for lab in labs:
                v = vocab.get(lab)
                entry_label.append(v); entry_obj.append(obj); n += 1
                cat = l2c.get(lab)
                if cat is not None and cat not in cat_ids:
                    cat_ids[cat] = len(cat_names); cat_names.append(cat)
        entry_cnt[r] = n
        events.append(("row", r, per_obj))
    n_cat = len(cat_names)
    cat_of_label = np.array([cat_ids[l2c[name]] if name in l2c else -1 for name in vocab.names], np.int32) \
        if vocab.names else np.zeros(0, np.int32)
    row_off = np.zeros(n_rows + 1, np.int64)
    np.cumsum(entry_cnt, out=row_off[1:])
    labels = np.array(entry_label, np.int32) if entry_label else np.zeros(0, np.int32)
    # ---- device: stable grouping by category, then split ids from the host permutation ----
    if n_cat and len(labels):
        exp_row, exp_entry, exp_cat, cat_off = _kernels().split_expand(row_off, labels, cat_of_label, n_cat)
    else:
        exp_row = np.zeros(0, np.int64); exp_entry = np.zeros(0, np.int64); cat_off = np.zeros(n_cat + 1, np.int64)
    sizes = np.diff(cat_off)
    n_train = np.array([int(n * train_ratio) for n in sizes], np.int64)
    n_val = np.array([int(n * val_ratio) for n in sizes], np.int64)
    from . import native
    perms = [native.permutation(random_seed, int(n)) for n in sizes]
    perm = np.concatenate(perms).astype(np.int64) if perms else np.zeros(0, np.int64)
    if len(perm):
        split_id, pos = _kernels().split_assign(cat_off, perm, n_train, n_val)
    else:
        split_id = np.zeros(0, np.uint8); pos = np.zeros(0, np.int64)
    # ---- egress: per-category frames ----
    categories, cat_counts = {}, {}
    for ci, cat in enumerate(cat_names):
        a, b = int(cat_off[ci]), int(cat_off[ci + 1])
        if b == a:
            continue
        cat_counts[cat] = b - a
        order = np.empty(b - a, np.int64)
        order[pos[a:b]] = np.arange(b - a)             # shuffled position -> original expanded row
        rows_idx = exp_row[a:b][order]
        keep_cols = [c for c in df.columns if c not in json_columns]        # the JSON columns are rewritten below:
        frame = df[keep_cols].iloc[rows_idx].reset_index(drop=True)         # do not gather their old texts first
        cells, labs, combos = [], [], []
        for e, r in zip(exp_entry[a:b][order], rows_idx):
            doc, combo = row_info[int(r)]
            lab = vocab.names[int(labels[e])]
            one = dict(entry_obj[int(e)]); one["name"] = lab      # serialised at once: a shallow copy is enough
            nd = {k: v for k, v in doc.items() if k != "objects"}
            nd["objects"] = [one]
            cells.append(json.dumps(nd, ensure_ascii=False)); labs.append(lab); combos.append(combo)
        for c in json_columns:
            if c in df.columns:
                frame[c] = cells
        frame = frame[list(df.columns)]                                      # original column order
        frame["分类标签"] = labs
        frame["分类类别"] = cat
        frame["原始标签组合"] = combos
        sp = split_id[a:b][order]
        categories[cat] = {"train": frame[sp == 0], "val": frame[sp == 1], "test": frame[sp == 2]}
    # ---- egress: unclassified rows and split_counts in reference order ----
    unc_rows, unc_reason, unc_label = [], [], []
    counts = []
    has_label_col = False
    for ev in events:
        if ev[0] == "bad_row":
            _, r, why = ev
            unc_rows.append(r); unc_reason.append(why); unc_label.append(None)
            counts.append({"source": sources[r], "原始标签组合": "", "拆分条数": 0, "是否可分类": "否", "无法分类原因": why})
            continue
        _, r, per_obj = ev
        combo = row_info[r][1]
        n_exp, reasons, any_ok = 0, set(), False
        for obj, labs in per_obj:
            if not labs:
                unc_rows.append(r); unc_reason.append("标注框缺少name字段"); unc_label.append(None)
                continue
            for lab in labs:
                if lab not in l2c:
                    why = f"标签{lab}未在规则中定义"
                    unc_rows.append(r); unc_reason.append(why); unc_label.append(lab); has_label_col = True
                    reasons.add(why)
                else:
                    any_ok = True; n_exp += 1
        if not any_ok:
            unc_rows.append(r); unc_label.append(None)
            unc_reason.append("；".join(sorted(reasons)) if reasons else "标签无法匹配规则")
        status = "否" if not any_ok else ("部分可分类" if reasons else "是")
        counts.append({"source": sources[r], "原始标签组合": combo, "拆分条数": n_exp, "是否可分类": status,
                       "无法分类原因": "；".join(sorted(reasons))})
    if unc_rows:
        unc = df.iloc[unc_rows].reset_index(drop=True)
        unc["无法分类原因"] = unc_reason
        if has_label_col:
            unc["无法分类标签"] = pd.Series(unc_label, dtype=object).where(pd.Series([x is not None for x in unc_label]), np.nan)
    else:
        unc = pd.DataFrame()
    return {
        "categories": categories,
        "unclassified": unc,
        "split_counts": pd.DataFrame(counts),
        "summary": {"categories": n_cat, "classified": int(sum(cat_counts.values())), "unclassified": len(unc_rows),
                    "category_counts": cat_counts},
    }
