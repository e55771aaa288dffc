"""CPU oracle (numpy / plain Python) for the hot path -- TEST INFRASTRUCTURE ONLY.

This module restates, on the CSR buffers of DESIGN.md §3, the arithmetic the
reference performs per row in /root/reference/src/deal_yolo_data/core/
processor.py.  It is the checker: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
The product path (``deal_yolo_daya_b200``) never does and fails loudly when
the CUDA library is missing.

Parity status: PINNED by execution.  The reference ships no tests or golden
vectors (its tests/ directory holds one empty file), so the pins are fixtures
produced by importing and running the unmodified reference step functions in
the authoring container (tests/golden/make_golden.py, pandas 3.0.2 /
numpy 2.3.5 / CPython 3.12.3); tests/test_oracle_golden.py holds this module
to those fixtures.

Everything here is scalar-loop Python so that it has *exactly* CPython's
``min`` / ``max`` / float semantics (first occurrence wins on ties, NaN is
order dependent, -0.0 == 0.0).  Use oracle/dyd_oracle.c for large inputs.
"""
from __future__ import annotations

import numpy as np

MASK64 = (1 << 64) - 1

# ----------------------------------------------------------------------------
# K1  polygon -> two corner points        reference: processor.py:252-260
# ----------------------------------------------------------------------------

def bbox_fold(poly_off, xy):
    """Left folds ``min``/``max`` over each polygon's vertices.

    Mirrors get_bbox_points (processor.py:252-260): four independent builtin
    min/max scans over the valid points, i.e. ``cur = v if v < cur else cur``
    (resp. ``>``) from the first vertex on.  Returns

        pts   float64[4*n_poly]  (min_x, min_y, max_x, max_y)
        valid uint8[n_poly]      0 where the polygon has no valid point
                                 (the reference emits {x: None, y: None} twice)
        arg   int32[4*n_poly]    vertex index (within the polygon) each value
                                 came from; lets the host re-emit the original
                                 JSON number (``10`` vs ``10.0``)
    """
    poly_off = np.asarray(poly_off, dtype=np.int64)
    xy = np.asarray(xy, dtype=np.float64)
    n = len(poly_off) - 1
    pts = np.zeros(4 * n, dtype=np.float64)
    arg = np.full(4 * n, -1, dtype=np.int32)
    valid = np.zeros(n, dtype=np.uint8)
    for p in range(n):
        a, b = int(poly_off[p]), int(poly_off[p + 1])
        if b <= a:
            continue
        valid[p] = 1
        mnx = mxx = float(xy[2 * a]); mny = mxy = float(xy[2 * a + 1])
        imnx = imxx = imny = imxy = 0
        for v in range(a + 1, b):
            x = float(xy[2 * v]); y = float(xy[2 * v + 1]); k = v - a
            if x < mnx: mnx, imnx = x, k
            if x > mxx: mxx, imxx = x, k
            if y < mny: mny, imny = y, k
            if y > mxy: mxy, imxy = y, k
        pts[4 * p:4 * p + 4] = (mnx, mny, mxx, mxy)
        arg[4 * p:4 * p + 4] = (imnx, imny, imxx, imxy)
    return pts, valid, arg


# ----------------------------------------------------------------------------
# K2  box count + any-pair IoU             reference: processor.py:328-376
# ----------------------------------------------------------------------------

def _py_min(a, b):   # builtin min(a, b): b only if strictly smaller
    return b if b < a else a


def _py_max(a, b):   # builtin max(a, b): b only if strictly larger
    return b if b > a else a


def iou_pair(b1, b2):
    """calculate_iou (processor.py:328-339), operation for operation."""
    xi1 = _py_max(b1[0], b2[0]); yi1 = _py_max(b1[1], b2[1])
    xi2 = _py_min(b1[2], b2[2]); yi2 = _py_min(b1[3], b2[3])
    inter = _py_max(0, xi2 - xi1) * _py_max(0, yi2 - yi1)
    if inter == 0:
        return 0.0
    area1 = (b1[2] - b1[0]) * (b1[3] - b1[1])
    area2 = (b2[2] - b2[0]) * (b2[3] - b2[1])
    union = area1 + area2 - inter
    return inter / union if union != 0 else 0.0


def effective_boxes(pts, valid, a, b):
    """extract_boxes (processor.py:341-366) on CSR: boxes of objects [a, b) up to
    the first null bbox (the TypeError it raises is swallowed by the function-wide
    ``except`` and the prefix collected so far is returned)."""
    out = []
    for q in range(a, b):
        if not valid[q]:
            break
        p1x, p1y, p2x, p2y = (float(pts[4 * q + k]) for k in range(4))
        out.append((_py_min(p1x, p2x), _py_min(p1y, p2y), _py_max(p1x, p2x), _py_max(p1y, p2y)))
    return out


def iou_filter(img_off, pts, valid, min_boxes, thr):
    """meet_conditions (processor.py:368-376) per image -> (high uint8, count int32)."""
    img_off = np.asarray(img_off, dtype=np.int64)
    n = len(img_off) - 1
    high = np.zeros(n, dtype=np.uint8)
    count = np.zeros(n, dtype=np.int32)
    for i in range(n):
        boxes = effective_boxes(pts, valid, int(img_off[i]), int(img_off[i + 1]))
        count[i] = len(boxes)
        if len(boxes) < min_boxes:
            continue
        hit = False
        for s in range(len(boxes)):
            for t in range(s + 1, len(boxes)):
                if iou_pair(boxes[s], boxes[t]) >= thr:
                    hit = True
                    break
            if hit:
                break
        high[i] = 1 if hit else 0
    return high, count


# ----------------------------------------------------------------------------
# K0  64-bit string hash (this repo's own definition)
# ----------------------------------------------------------------------------

HM = 0xC6A4A7935BD1E995
HSEED = 0x8445D61A4E774912


def hash_bytes(b: bytes) -> int:
    """MurmurHash64A-style hash of a byte string (little-endian 8-byte words)."""
    n = len(b)
    h = (HSEED ^ (n * HM)) & MASK64
    nblk = n // 8
    for i in range(nblk):
        k = int.from_bytes(b[8 * i:8 * i + 8], "little")
        k = (k * HM) & MASK64; k ^= k >> 47; k = (k * HM) & MASK64
        h ^= k; h = (h * HM) & MASK64
    tail = b[8 * nblk:]
    if tail:
        h ^= int.from_bytes(tail, "little")
        h = (h * HM) & MASK64
    h ^= h >> 47; h = (h * HM) & MASK64; h ^= h >> 47
    return h


def hash_strings(strings):
    return np.array([hash_bytes(s.encode("utf-8")) for s in strings], dtype=np.uint64)


# ----------------------------------------------------------------------------
# K4  drop_duplicates(subset=["source"])   reference: processor.py:140-144
# ----------------------------------------------------------------------------

def dedup(keys, null, keep="first"):
    """Row keep-mask of ``DataFrame.drop_duplicates(subset=[col], keep=keep)``.

    ``keys`` are the 64-bit hashes of the strings, ``null`` marks NaN cells (all
    NaN cells are one group -- pandas treats NaN == NaN when de-duplicating).
    Returns (keep uint8[n], rep int64[n]) where ``rep`` is the row the group is
    represented by (first row for keep="first"/False, last row for "last").
    """
    keys = np.asarray(keys, dtype=np.uint64); null = np.asarray(null, dtype=np.uint8)
    n = len(keys)
    first, last, cnt = {}, {}, {}
    for r in range(n):
        k = None if null[r] else int(keys[r])
        first.setdefault(k, r); last[k] = r; cnt[k] = cnt.get(k, 0) + 1
    keepm = np.zeros(n, dtype=np.uint8); rep = np.zeros(n, dtype=np.int64)
    for r in range(n):
        k = None if null[r] else int(keys[r])
        if keep == "first":
            rep[r] = first[k]; keepm[r] = rep[r] == r
        elif keep == "last":
            rep[r] = last[k]; keepm[r] = rep[r] == r
        else:
            rep[r] = first[k]; keepm[r] = cnt[k] == 1
    return keepm, rep


# ----------------------------------------------------------------------------
# K5  anti-join                            reference: processor.py:194-199
# ----------------------------------------------------------------------------

def antijoin(main_keys, main_null, ref_keys, ref_null):
    """keep[i] = main value not in set(ref.dropna()); NaN main rows never match.

    Returns (keep uint8[n_main], ref_row int64[n_main]) with ref_row = the first
    reference row holding the matched key (-1 when kept).
    """
    ref_first = {}
    for r in range(len(ref_keys)):
        if not ref_null[r]:
            ref_first.setdefault(int(ref_keys[r]), r)
    n = len(main_keys)
    keepm = np.ones(n, dtype=np.uint8); ref_row = np.full(n, -1, dtype=np.int64)
    for r in range(n):
        if main_null[r]:
            continue
        hit = ref_first.get(int(main_keys[r]))
        if hit is not None:
            keepm[r] = 0; ref_row[r] = hit
    return keepm, ref_row


# ----------------------------------------------------------------------------
# K3  label remap through a lookup table   reference: processor.py:582-602,
#                                          utils.py:659-679
# ----------------------------------------------------------------------------

def label_lut(img_off, label_id, lut_new, lut_ntok, lut_nrep):
    """Per-object name rewrite with the counters of replace_labels_by_mapping.

    ``label_id`` indexes the table's vocabulary of distinct raw ``name`` strings
    (-1 = object without a name).  The three per-vocabulary tables are built on
    the host with the reference's string rules (utils.py:659-679):
    ``lut_new[v]`` id of the rewritten name, ``lut_ntok[v]`` number of tokens,
    ``lut_nrep[v]`` number of tokens found in the mapping.  A name is rewritten
    only when at least one token was replaced (processor.py:596-600).
    Returns (new_id int32[n_box], row_replaced uint8[n_img], counters dict).
    """
    label_id = np.asarray(label_id, dtype=np.int32)
    n_img = len(img_off) - 1
    new_id = label_id.copy()
    row_rep = np.zeros(n_img, dtype=np.uint8)
    c = dict(total_objects=0, missing_name_objects=0, total_labels=0,
             replaced_labels=0, replaced_objects=0)
    for i in range(n_img):
        for q in range(int(img_off[i]), int(img_off[i + 1])):
            c["total_objects"] += 1
            v = int(label_id[q])
            if v < 0:
                c["missing_name_objects"] += 1
                continue
            c["total_labels"] += int(lut_ntok[v])
            if lut_nrep[v] > 0:
                new_id[q] = lut_new[v]
                c["replaced_labels"] += int(lut_nrep[v])
                c["replaced_objects"] += 1
                row_rep[i] = 1
    return new_id, row_rep, c


# ----------------------------------------------------------------------------
# K6  label -> category expansion + split  reference: processor.py:741-806
# ----------------------------------------------------------------------------

def split_expand(img_off, label_id, cat_of_label, n_cat):
    """One expanded row per (object, label) whose label has a category
    (processor.py:751-775), grouped by category in encounter order.

    Single-token names only at this level (multi-token names are expanded on the
    host before ids are assigned).  Returns (exp_img int64, exp_box int64,
    exp_cat int32, cat_off int64[n_cat+1]) with rows of category c occupying
    [cat_off[c], cat_off[c+1]) in original (image, object) order.
    """
    rows = [[] for _ in range(n_cat)]
    n_img = len(img_off) - 1
    for i in range(n_img):
        for q in range(int(img_off[i]), int(img_off[i + 1])):
            v = int(label_id[q])
            if v < 0:
                continue
            c = int(cat_of_label[v])
            if c >= 0:
                rows[c].append((i, q))
    cat_off = np.zeros(n_cat + 1, dtype=np.int64)
    for c in range(n_cat):
        cat_off[c + 1] = cat_off[c] + len(rows[c])
    flat = [rc for c in range(n_cat) for rc in rows[c]]
    exp_img = np.array([f[0] for f in flat], dtype=np.int64)
    exp_box = np.array([f[1] for f in flat], dtype=np.int64)
    exp_cat = np.repeat(np.arange(n_cat, dtype=np.int32), np.diff(cat_off))
    return exp_img, exp_box, exp_cat, cat_off


def split_assign(cat_off, train_ratio, val_ratio, test_ratio, seed):
    """Per-category shuffle + cut (processor.py:673-676, 796-806).

    ``DataFrame.sample(frac=1, random_state=seed)`` orders the rows by
    ``np.random.RandomState(seed).permutation(n)`` with a fresh generator per
    category; the first int(n*train) rows are train, the next int(n*val) val,
    the rest test.  Returns (split uint8[n_exp] 0/1/2, pos int64[n_exp]) where
    ``pos`` is each expanded row's position inside its category's shuffled order.
    """
    s = train_ratio + val_ratio + test_ratio
    tr, va = train_ratio / s, val_ratio / s
    n_exp = int(cat_off[-1])
    split = np.zeros(n_exp, dtype=np.uint8); pos = np.zeros(n_exp, dtype=np.int64)
    for c in range(len(cat_off) - 1):
        a, b = int(cat_off[c]), int(cat_off[c + 1]); n = b - a
        if n == 0:
            continue
        perm = np.random.RandomState(seed).permutation(n)
        ntr, nva = int(n * tr), int(n * va)
        inv = np.empty(n, dtype=np.int64); inv[perm] = np.arange(n)
        pos[a:b] = inv
        split[a:b] = np.where(inv < ntr, 0, np.where(inv < ntr + nva, 1, 2))
    return split, pos


# ----------------------------------------------------------------------------
# YOLO label arithmetic ("next" row f-3)   reference: processor.py:1045-1052
# ----------------------------------------------------------------------------

def yolo_norm(box, width, height):
    """(cx, cy, w, h) of one kept box, or None when degenerate."""
    x1, y1, x2, y2 = box
    x1, x2 = _py_min(x1, x2), _py_max(x1, x2)
    y1, y2 = _py_min(y1, y2), _py_max(y1, y2)
    bw = _py_max(x2 - x1, 0.0); bh = _py_max(y2 - y1, 0.0)
    if bw <= 0 or bh <= 0:
        return None
    return ((x1 + x2) / 2 / width, (y1 + y2) / 2 / height, bw / width, bh / height)
