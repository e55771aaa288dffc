"""Native (C++) ingest / egress against the CPython lane and the reference's golden outputs.

Runs on CPU: the kernels are replaced by the oracle stand-in, the subject under test is host code
(csrc/ingest.cpp through deal_yolo_daya_b200/native.py).  Every case is checked both ways: the
native lane must give byte-identical results to the CPython lane, and for the cases meant to be
canonical it must actually have taken them (otherwise the comparison would be vacuous).
"""
from __future__ import annotations

import gzip
import io
import json
import random
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

from deal_yolo_daya_b200 import ingest, native, processor as P, synth
from tests.oracle_kernels import OracleKernels

G = Path(__file__).resolve().parent / "golden"
ANN, NEW = P.COL_ANN, P.COL_NEW


@pytest.fixture(autouse=True)
def _oracle_facade(monkeypatch):
    monkeypatch.setattr(P, "KERNELS", OracleKernels())


def _as_lists(res):
    """(cells, widths, heights) as plain lists: the native lane answers with an Arrow string array and
    int64 arrays when every row took it (what pandas makes of the lists anyway)."""
    cells, w, h = res
    un = lambda v: [x.item() if isinstance(x, np.generic) else x for x in v]   # noqa: E731
    return [c if isinstance(c, str) else None for c in list(cells)], un(list(w)), un(list(h))


def both_lanes_replace(cells, monkeypatch):
    monkeypatch.setenv("DYD_NATIVE_INGEST", "1")
    a = _as_lists(P.replace_ptlist_cells(cells))
    stats = dict(P.STATS)
    a_series = _as_lists(P.replace_ptlist_cells(pd.Series(cells)))            # Arrow-backed column in
    assert a_series[0] == a[0] and a_series[1] == a[1] and a_series[2] == a[2]
    monkeypatch.setenv("DYD_NATIVE_INGEST", "0")
    b = _as_lists(P.replace_ptlist_cells(cells))
    return a, b, stats


def same_scalars(x, y):
    return len(x) == len(y) and all(type(p) is type(q) and (p == q or (p != p and q != q)) for p, q in zip(x, y))


def test_float_repr_is_cpython_repr():
    import ctypes as C
    from deal_yolo_daya_b200 import _lib
    lib = _lib.load(); buf = C.create_string_buffer(40)
    rng = random.Random(3)
    vals = [0.0, -0.0, 1e15, 1e16, 9999999999999998.0, 1e-4, 9.999e-5, 1e-5, 5e-324, 1.7976931348623157e308, 0.1, 1 / 3, 1920.0, 123.456]
    vals += [rng.uniform(-3000, 3000) for _ in range(20000)] + [round(rng.uniform(0, 2000), rng.randint(0, 8)) for _ in range(20000)]
    vals += [np.frombuffer(rng.getrandbits(64).to_bytes(8, "little"), np.float64)[0].item() for _ in range(20000)]
    for v in vals:
        n = lib.dyd_py_float_repr(C.c_double(v), buf)
        want = "NaN" if v != v else json.dumps(v)
        assert buf.raw[:n].decode() == want, v


def test_golden_rows_take_the_native_lane_and_match(monkeypatch):
    df = pd.read_csv(io.StringIO(gzip.open(G / "expected" / "filtered_main.csv.gz", "rt", encoding="utf-8").read()))
    cells = df[ANN].tolist()
    a, b, stats = both_lanes_replace(cells, monkeypatch)
    assert a[0] == b[0] and same_scalars(a[1], b[1]) and same_scalars(a[2], b[2])
    want = pd.read_csv(io.StringIO(gzip.open(G / "expected" / "processed_replaced_ptlist.csv.gz", "rt", encoding="utf-8").read()))
    kept = [c for c, src in zip(a[0], cells) if isinstance(src, str)]
    assert [c if c is not None else None for c in kept] == [c if isinstance(c, str) else None for c in want[NEW].tolist()]
    # the synthetic rows and most hand-written rows are canonical: the native lane must have taken them
    assert stats["native_rows"] >= 55 and stats["slow_rows"] <= 12


def test_synthetic_table_full_precision(monkeypatch):
    t = synth.make_table(11, 500, 300)
    cells = [r[1] for r in synth.table_to_rows(t)]             # 17-digit doubles: repr round trip
    a, b, stats = both_lanes_replace(cells, monkeypatch)
    assert a == b and stats["native_rows"] == 300 and stats["slow_rows"] == 0


CANON = lambda d: json.dumps(d, ensure_ascii=False)  # noqa: E731


def edge_docs():
    pt = lambda x, y: {"x": x, "y": y}  # noqa: E731
    docs = [
        {"objects": []},
        {"width": 5, "height": 6.5, "objects": [{}]},
        {"objects": [{"name": "a"}]},                                         # polygon appended to the object
        {"objects": [{"polygon": {}}]},                                       # ptList appended to an empty polygon
        {"objects": [{"polygon": {"k": 1}}]},                                 # ptList appended after other keys
        {"objects": [{"polygon": {"ptList": []}}]},
        {"objects": [{"polygon": {"ptList": [pt(1, 2)]}, "z": [1, {"q": None}]}]},
        {"objects": [{"polygon": {"ptList": [pt(True, False), pt(2, 3)]}}]},  # bools are ints for min/max
        {"objects": [{"polygon": {"ptList": [pt(1e22, -1e-7), pt(1.5e300, 5e-324), pt(-0.0, 0.0)]}}]},
        {"objects": [{"polygon": {"ptList": [pt(float("nan"), 1.0), pt(float("inf"), float("-inf"))]}}]},
        {"objects": [{"polygon": {"ptList": [pt(10.0, 5), pt(10, 5.0), pt(10, 7)]}}]},
        {"objects": [{"polygon": {"ptList": [{"x": 1}, 5, "s", None, [1], pt(3, 4)]}}]},
        {"objects": [{"name": "含\"引号\n换行\t制表\x01控制", "polygon": {"ptList": [pt(1, 1)]}}], "路径": "C:\\x"},
        {"width": None, "height": True, "objects": [{"polygon": {"ptList": [pt(9007199254740992, 1)]}}]},
        {"id": 123456789012345678901234567890, "objects": [{"polygon": {"ptList": [pt(1, 2), pt(3, 4)]}}]},
    ]
    return docs


SLOW_TEXTS = [
    '{"objects":[{"polygon":{"ptList":[{"x":1,"y":2}]}}]}',                  # compact separators
    '{"objects": [{"polygon": {"ptList": [{"x": 1.50, "y": 2}]}}]}',         # 1.50 is not repr(1.5)
    '{"objects": [{"polygon": {"ptList": [{"x": 1e3, "y": 2}]}}]}',          # 1e3 -> 1000.0
    '{"objects": [{"polygon": {"ptList": [{"x": -0, "y": 2}]}}]}',           # -0 -> 0
    '{"objects": [{"polygon": {"ptList": [{"x": 1, "y": 2}]}}], "a": "\\u4e2d"}',   # escaped non-ASCII
    '{"objects": [{"polygon": {"ptList": [{"x": 1, "y": 2}]}}], "a": "\\/"}',
    ' {"objects": []}',                                                       # leading whitespace
    '{"objects": [1, {"polygon": {"ptList": [{"x": 1, "y": 2}]}}]}',          # non-dict object is dropped
    '{"objects": [{"polygon": {"ptList": [{"x": 1, "y": 2}]}, "polygon": {}}]}',    # duplicate key
    '{"objects": [{"polygon": {"ptList": [{"x": 1, "x": 3, "y": 2}]}}]}',
    '{"a": 1}',                                                               # no objects key
    '{"objects": [{"polygon": {"ptList": [{"x": 9007199254740993, "y": 2}]}}]}',    # int beyond 2^53 as coordinate
    '{"objects": [{"polygon": {"ptList": [{"x": 1, "y": 2}]}}], "width": "w"}',
    '{"objects": [oops',
    '[1, 2]',
    '{"objects": [{"polygon": {"ptList": [{"x": 01, "y": 2}]}}]}',
]


def test_edge_documents_both_lanes_agree(monkeypatch):
    cells = [CANON(d) for d in edge_docs()]
    a, b, stats = both_lanes_replace(cells, monkeypatch)
    assert a[0] == b[0] and same_scalars(a[1], b[1]) and same_scalars(a[2], b[2])
    assert stats["native_rows"] >= len(cells) - 1                              # only the 2^53+ row may go slow
    from oracle import pipeline_port as port                                   # and they equal the reference's algorithm
    assert a[0] == [port.replace_cell(c) for c in cells]


def test_non_canonical_rows(monkeypatch):
    """Documents that differ from json.dumps form only in style (first seven) are rewritten canonically inside the
    library and take the native lane; the others (dropped objects, duplicate keys, values fp64 cannot carry,
    invalid JSON) go to the CPython lane.  Either way the cell equals the reference's."""
    from oracle import pipeline_port as port
    ok = [t for t in SLOW_TEXTS if t not in ('[1, 2]',)]                       # a list document crashes the reference
    a, b, stats = both_lanes_replace(ok, monkeypatch)
    assert a[0] == b[0] == [port.replace_cell(c) for c in ok]
    style_only = SLOW_TEXTS[:7]
    a, b, stats = both_lanes_replace(style_only, monkeypatch)
    assert stats["native_rows"] == len(style_only) and stats["slow_rows"] == 0
    assert a[0] == b[0] == [port.replace_cell(c) for c in style_only]
    rest = [t for t in SLOW_TEXTS[7:] if t != '[1, 2]']
    a, b, stats = both_lanes_replace(rest, monkeypatch)
    assert stats["native_rows"] == 0 and stats["slow_rows"] == len(rest)
    with pytest.raises(AttributeError):
        monkeypatch.setenv("DYD_NATIVE_INGEST", "1")
        P.replace_ptlist_cells(['[1, 2]'])
    with pytest.raises(TypeError):
        P.replace_ptlist_cells([CANON({"objects": [{"polygon": {"ptList": [{"x": None, "y": 1}, {"x": 2, "y": 3}]}}]})])


def rand_doc(rng):
    def num():
        k = rng.random()
        if k < 0.3:
            return rng.randint(-50, 4000)
        if k < 0.9:
            return round(rng.uniform(-10, 2000), rng.randint(0, 12))
        return rng.choice([0.0, -0.0, 1e-7, 123456789.0, 1e16, float("inf"), float("nan"), True])
    objs = []
    for _ in range(rng.randint(0, 6)):
        o = {}
        if rng.random() < 0.8:
            o["name"] = rng.choice(["cls01", "行人", "a,b；c", "", 'q"uote', "tab\t"])
        if rng.random() < 0.9:
            pg = {}
            if rng.random() < 0.2:
                pg["kind"] = "poly"
            if rng.random() < 0.9:
                pg["ptList"] = [({"x": num(), "y": num()} if rng.random() < 0.93 else rng.choice([{"x": 1}, {}, 7, "p", None]))
                                for _ in range(rng.randint(0, 9))]
            o["polygon"] = pg
        if rng.random() < 0.3:
            o["extra"] = {"k": [1, 2.5, None, {"z": "w"}], "e": {}}
        objs.append(o)
    d = {}
    if rng.random() < 0.8:
        d["width"] = rng.choice([1920, 1080.0, None, 0])
    d["objects"] = objs
    if rng.random() < 0.8:
        d["height"] = rng.choice([1080, 720.5, True])
    return d


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_documents(monkeypatch, seed):
    from oracle import pipeline_port as port
    rng = random.Random(seed)
    cells = []
    for _ in range(400):
        d = rand_doc(rng)
        t = CANON(d)
        if rng.random() < 0.15:                                # a non-canonical rendering of the same document
            t = json.dumps(d, ensure_ascii=rng.random() < 0.5, separators=rng.choice([(",", ":"), (", ", ": "), (" , ", " : ")]),
                           indent=rng.choice([None, 1]))
        cells.append(t)
    cells += [float("nan"), None]
    a, b, stats = both_lanes_replace(cells, monkeypatch)
    assert a[0] == b[0] and same_scalars(a[1], b[1]) and same_scalars(a[2], b[2])
    assert a[0][:-2] == [port.replace_cell(c) for c in cells[:-2]] and a[0][-2:] == [None, None]
    assert stats["native_rows"] > 250


def test_boxes_mode_matches_cpython_lane(monkeypatch):
    rng = random.Random(5)
    rows = []
    for _ in range(600):
        objs = []
        for _ in range(rng.randint(0, 7)):
            k = rng.random()
            if k < 0.7:
                pl = [{"x": rng.uniform(0, 50), "y": rng.randint(0, 50)}, {"x": rng.uniform(0, 60), "y": rng.uniform(0, 60)}]
            elif k < 0.78:
                pl = [{"x": None, "y": None}, {"x": None, "y": None}]
            elif k < 0.84:
                pl = [{"x": 1, "y": 2}]
            elif k < 0.9:
                pl = [{"x": 1, "y": 2}, 5]
            elif k < 0.95:
                pl = [{"x": None, "y": 3}, {"x": 2, "y": 3}]
            else:
                pl = [{"x": "a", "y": 3}, {"x": 2, "y": 3}]
            o = {"name": "n", "polygon": {"ptList": pl}} if rng.random() < 0.95 else rng.choice([7, {"name": "x"}, {"polygon": None}])
            objs.append(o)
        text = json.dumps({"objects": objs}, separators=rng.choice([(",", ":"), (", ", ": ")]), indent=rng.choice([None, None, 2]))
        rows.append(text)
    rows += ["{bad", float("nan"), '{"objects": 5}', '{"a": 1}', "[]"]
    from oracle import pipeline_port as port
    want = np.array([port.is_high_iou(port.boxes_of_cell(t), 2, 0.02) for t in rows])
    monkeypatch.setenv("DYD_NATIVE_INGEST", "1")
    got = P.high_iou_mask(rows, 2, 0.02)
    native_slow = P.STATS["slow_rows"]
    monkeypatch.setenv("DYD_NATIVE_INGEST", "0")
    got2 = P.high_iou_mask(rows, 2, 0.02)
    assert np.array_equal(got, want) and np.array_equal(got2, want)
    assert native_slow < 0.2 * len(rows) and want.sum() > 20


def test_remap_native_lane_equals_python_lane(monkeypatch):
    """Step 5.5 on json.dumps-form cells: names spliced by csrc/ingest.cpp (mode 2) vs the CPython lane --
    same frame, summary, diff rows and unmatched counter; awkward cells send the whole frame to CPython."""
    from deal_yolo_daya_b200 import labels
    from oracle import pipeline_port as port
    t = synth.make_table(5, 0, 400)
    df = pd.DataFrame(synth.table_to_rows(t), columns=["source", ANN])
    rep, _ = port.replace_ptlist_df(df)
    rep = pd.read_csv(io.StringIO(rep.to_csv(index=False)))                 # as step 5.5 reads it: Arrow-backed str columns
    # a few name shapes: multi-token, unknown, empty, None, missing key, unicode
    docs = [json.loads(x) for x in rep[NEW]]
    docs[0]["objects"][0]["name"] = "cls01,cls02；cls01"
    docs[1]["objects"][0]["name"] = "未知标签"
    docs[2]["objects"][0]["name"] = ""
    docs[3]["objects"][0]["name"] = None
    docs[4]["objects"][0].pop("name", None)
    docs[5]["objects"] = []
    docs[6].pop("objects")
    rep[NEW] = [json.dumps(d, ensure_ascii=False) for d in docs]
    rep.loc[7, NEW] = np.nan
    rep = pd.read_csv(io.StringIO(rep.to_csv(index=False)))
    lm = {synth.label_name(i): f"grp{i % 20:02d}" for i in range(synth.N_LABELS) if i % 7}
    lm["cls02"] = "cls01"

    def run(native_on):
        monkeypatch.setenv("DYD_NATIVE_INGEST", "1" if native_on else "0")
        out = labels.remap_df(rep, lm)
        return out, labels.LAST["remap_lane"]
    (a, lane_a), (b, lane_b) = run(True), run(False)
    assert (lane_a, lane_b) == ("native", "python")
    pd.testing.assert_frame_equal(a[0], b[0])
    assert a[1] == b[1] and a[2] == b[2] and a[3] == b[3] and list(a[3]) == list(b[3])
    # a cell in another style (extra space) is rewritten canonically inside the library: still the native lane
    rep2 = rep.copy()
    rep2[NEW] = rep2[NEW].astype(object)
    rep2.loc[9, NEW] = rep2.loc[9, NEW].replace('{"', '{ "', 1)
    rep2 = pd.read_csv(io.StringIO(rep2.to_csv(index=False)))
    monkeypatch.setenv("DYD_NATIVE_INGEST", "1")
    c = labels.remap_df(rep2, lm)
    assert labels.LAST["remap_lane"] == "native"
    monkeypatch.setenv("DYD_NATIVE_INGEST", "0")
    d = labels.remap_df(rep2, lm)
    pd.testing.assert_frame_equal(c[0], d[0]); assert c[1:] == d[1:]
    # a name that needs an escape when written -> the whole call takes the CPython lane, same answer
    docs3 = [json.loads(x) if isinstance(x, str) else None for x in rep[NEW]]
    docs3[11]["objects"][0]["name"] = 'say "cheese"'
    rep3 = rep.copy()
    rep3[NEW] = [json.dumps(x, ensure_ascii=False) if x is not None else np.nan for x in docs3]
    rep3 = pd.read_csv(io.StringIO(rep3.to_csv(index=False)))
    monkeypatch.setenv("DYD_NATIVE_INGEST", "1")
    e = labels.remap_df(rep3, lm)
    assert labels.LAST["remap_lane"] == "python"
    monkeypatch.setenv("DYD_NATIVE_INGEST", "0")
    f = labels.remap_df(rep3, lm)
    pd.testing.assert_frame_equal(e[0], f[0]); assert e[1:] == f[1:]


def test_remap_lanes_agree_on_random_documents(monkeypatch):
    """Random documents (shuffled keys, missing / non-list "objects", objects without names, odd name
    values, junk members, missing cells): the native lane, where it applies, and the CPython lane give the
    same frame / summary / diff rows / unmatched counter, or raise the same exception."""
    from deal_yolo_daya_b200 import labels
    rng = random.Random(99)
    names_pool = ["cat", "dog", "cat,dog", "猫", "cat；bird", " cat ", "", "x|y", "bird", "a,b,a"]
    odd_names = [None, 5, 2.5, True, ["cat"], {"a": 1}, 'q"uote', "back\\slash", "new\nline"]

    def val(d=0):
        t = rng.random()
        if t < 0.3: return rng.randint(-5, 2000)
        if t < 0.5: return round(rng.uniform(-10, 3000), rng.randint(0, 6))
        if t < 0.6: return rng.choice([None, True, False])
        if t < 0.8: return rng.choice(["s", "中文", "", "a b"])
        if d > 1: return 1
        if t < 0.9: return [val(d + 1) for _ in range(rng.randint(0, 3))]
        return {rng.choice("abcxyz"): val(d + 1) for _ in range(rng.randint(0, 3))}

    def obj(clean):
        if not clean and rng.random() < 0.1:
            return rng.choice([1, "str", None, [1]])
        o, keys = {}, ["name", "polygon", "id", "score"]
        rng.shuffle(keys)
        for k in keys:
            if k == "name":
                if rng.random() < 0.85:
                    o["name"] = rng.choice(names_pool) if (clean or rng.random() < 0.8) else rng.choice(odd_names)
            elif rng.random() < 0.6:
                o[k] = val()
        return o

    def doc(clean):
        d, keys = {}, ["width", "objects", "height", "meta"]
        rng.shuffle(keys)
        for k in keys:
            if k == "objects":
                r = rng.random()
                if r < 0.85: d[k] = [obj(clean) for _ in range(rng.randint(0, 5))]
                elif r < 0.92: d[k] = val()
            elif rng.random() < 0.7:
                d[k] = val()
        return d
    lm = {"cat": "animal", "dog": "animal", "猫": "animal", "bird": "bird2", "x": "y"}
    lanes = {"native": 0, "python": 0}
    for _ in range(120):
        clean = rng.random() < 0.7
        n = rng.randint(1, 10)
        col_new = [json.dumps(doc(clean), ensure_ascii=False) if rng.random() < 0.93 else rng.choice([None, ""] if clean else [None, "", "not json", "[1, 2]"])
                   for _ in range(n)]
        col_ann = [json.dumps(doc(clean), ensure_ascii=False) if rng.random() < 0.9 else None for _ in range(n)]
        df = pd.read_csv(io.StringIO(pd.DataFrame({"source": [f"u{i}" for i in range(n)], ANN: col_ann, NEW: col_new}).to_csv(index=False)))
        for c in (ANN, NEW):
            if str(df[c].dtype) != "str" and df[c].notna().any():
                df[c] = df[c].astype("str")
        res = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("DYD_NATIVE_INGEST", mode)
            try:
                res[mode] = (labels.remap_df(df, lm), None)
            except Exception as e:  # noqa: BLE001
                res[mode] = (None, type(e).__name__)
            if mode == "1":
                lanes[labels.LAST["remap_lane"]] += 1
        (a, ea), (b, eb) = res["1"], res["0"]
        assert ea == eb
        if ea is None:
            pd.testing.assert_frame_equal(a[0], b[0])
            assert a[1] == b[1] and a[2] == b[2] and a[3] == b[3] and list(a[3]) == list(b[3])
    assert lanes["native"] > 40 and lanes["python"] > 10


def test_split_lanes_agree_on_random_documents(monkeypatch):
    """split_df: the native / array lane (names and spans from csrc/ingest.cpp, cells spliced by
    dyd_egress_split, bookkeeping on arrays) against the CPython lane on random documents -- same category
    frames (values, dtypes, index), unclassified sheet, split_counts sheet and summary, or the same exception."""
    from deal_yolo_daya_b200 import labels
    rng = random.Random(4242)
    names_pool = ["cat", "dog", "cat,dog", "猫", "cat；bird", " cat ", "", "x|y", "bird", "a,b,a", "ab", "a"]
    odd_names = [None, 5, 2.5, True, ["cat"], {"a": 1}, 'q"uote', "back\\slash", "new\nline"]

    def val(d=0):
        t = rng.random()
        if t < 0.3: return rng.randint(-5, 2000)
        if t < 0.5: return round(rng.uniform(-10, 3000), rng.randint(0, 6))
        if t < 0.6: return rng.choice([None, True, False])
        if t < 0.8: return rng.choice(["s", "中文", "", "a b"])
        if d > 1: return 1
        if t < 0.9: return [val(d + 1) for _ in range(rng.randint(0, 3))]
        return {rng.choice("abcxyz"): val(d + 1) for _ in range(rng.randint(0, 3))}

    def obj(clean):
        if rng.random() < (0.05 if clean else 0.1):
            return rng.choice([1, "str", None, [1]])
        o, keys = {}, ["name", "polygon", "id", "score"]
        rng.shuffle(keys)
        for k in keys:
            if k == "name":
                if rng.random() < 0.85:
                    o["name"] = rng.choice(names_pool) if (clean or rng.random() < 0.8) else rng.choice(odd_names)
                    if clean and rng.random() < 0.05:
                        o["name"] = None
            elif rng.random() < 0.6:
                o[k] = val()
        return o

    def doc(clean):
        d, keys = {}, ["width", "objects", "height", "meta"]
        rng.shuffle(keys)
        for k in keys:
            if k == "objects":
                r = rng.random()
                if r < 0.85: d[k] = [obj(clean) for _ in range(rng.randint(0, 5))]
                elif r < 0.92: d[k] = val()
            elif rng.random() < 0.7:
                d[k] = val()
        return d

    def same_frame(a, b):
        if len(a) == 0 and len(b) == 0:
            return
        pd.testing.assert_frame_equal(a.reset_index(drop=True), b.reset_index(drop=True))
        assert list(a.index) == list(b.index)
    l2c = {"cat": "动物", "dog": "动物", "猫": "动物", "bird": "鸟类", "x": "其它", "a": "其它"}
    lanes = {"native": 0, "python": 0}
    for _ in range(150):
        clean = rng.random() < 0.75
        n = rng.randint(1, 14)
        col_new = [json.dumps(doc(clean), ensure_ascii=False) if (clean or rng.random() < 0.93) else rng.choice([None, "", "not json", "[1, 2]"])
                   for _ in range(n)]
        col_ann = [json.dumps(doc(clean), ensure_ascii=False) if rng.random() < 0.9 else None for _ in range(n)]
        cols = {"source": [f"u{i}" for i in range(n)], ANN: col_ann, NEW: col_new}
        if rng.random() < 0.3:
            cols["width"] = [rng.randint(1, 9) for _ in range(n)]
        df = pd.read_csv(io.StringIO(pd.DataFrame(cols).to_csv(index=False)))
        for c in (ANN, NEW):
            if str(df[c].dtype) != "str" and df[c].notna().any():
                df[c] = df[c].astype("str")
        res = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("DYD_NATIVE_INGEST", mode)
            try:
                res[mode] = (labels.split_df(df, l2c), None)
            except Exception as e:  # noqa: BLE001
                res[mode] = (None, type(e).__name__)
            if mode == "1":
                lanes[labels.LAST["split_lane"]] += 1
        (a, ea), (b, eb) = res["1"], res["0"]
        assert ea == eb
        if ea is None:
            assert a["summary"] == b["summary"] and list(a["categories"]) == list(b["categories"])
            for cat in a["categories"]:
                for part in ("train", "val", "test"):
                    same_frame(a["categories"][cat][part], b["categories"][cat][part])
            same_frame(a["unclassified"], b["unclassified"]); same_frame(a["split_counts"], b["split_counts"])
    assert lanes["native"] > 60 and lanes["python"] > 15


def test_canonical_rewriter_equals_cpython_json():
    """dyd_json_canonical(text) == json.dumps(json.loads(text), ensure_ascii=False) on random values written in
    random styles (compact / padded separators, indentation, \\u escapes, alternative number spellings,
    truncations); documents json.loads rejects must be declined, and nothing valid may be declined here."""
    import ctypes as C
    import re
    from deal_yolo_daya_b200 import _lib
    lib = _lib.load()
    rng = random.Random(2718)

    def canon(text):
        b = text.encode("utf-8")
        out = C.create_string_buffer(len(b) * 6 + 64)
        n = lib.dyd_json_canonical(b, len(b), out, len(out))
        return None if n < 0 else out.raw[:n].decode("utf-8")
    chars = ["a", "b", " ", "中", "é", "😀", '"', "\\", "\n", "\t", "\r", "\b", "\f", "\x01", "\x1f", "\x7f", "/", "x", "1", "{", "}", "[", ","]

    def rstr():
        return "".join(rng.choice(chars) for _ in range(rng.randint(0, 8)))

    def rnum():
        t = rng.random()
        if t < 0.25: return rng.randint(-10**6, 10**6)
        if t < 0.3: return rng.choice([0, -0.0, 0.0, 10**20, -10**25, 2**53, 2**53 + 1, 1e16, 1e15, 123456789012345678])
        if t < 0.6: return round(rng.uniform(-5000, 5000), rng.randint(0, 10))
        if t < 0.7: return rng.uniform(-1, 1) * 10 ** rng.randint(-320, 308)
        if t < 0.75: return rng.choice([float("inf"), float("-inf"), float("nan"), 5e-324, 1.7976931348623157e308])
        return rng.random()

    def rval(d=0):
        t = rng.random()
        if d > 3 or t < 0.35: return rnum()
        if t < 0.5: return rstr()
        if t < 0.6: return rng.choice([True, False, None])
        if t < 0.8: return [rval(d + 1) for _ in range(rng.randint(0, 4))]
        return {rstr() if rng.random() < 0.3 else rng.choice(["x", "y", "name", "objects", "k1", "键"]): rval(d + 1) for _ in range(rng.randint(0, 4))}

    def style(v):
        t, kw = rng.random(), {}
        if t < 0.3: kw["separators"] = (",", ":")
        elif t < 0.5: kw["indent"] = rng.choice([1, 2, 4, "\t"])
        elif t < 0.6: kw["separators"] = (" , ", " : ")
        kw["ensure_ascii"] = rng.random() < 0.5
        s = json.dumps(v, **kw)
        if rng.random() < 0.3:
            s = rng.choice([" ", "\n", "\t\r\n"]) + s + rng.choice(["", " ", "\n"])
        return s

    def respell(m):
        lit, r = m.group(0), rng.random()
        plain_int = "." not in lit and "e" not in lit.lower()
        if r < 0.2 and "." in lit and "e" not in lit.lower(): return lit + "0"
        if r < 0.3 and plain_int and len(lit) < 8: return lit + "E0"
        if r < 0.4 and "e" in lit: return lit.replace("e", "E")
        if r < 0.45 and plain_int: return lit + ".0"
        return lit
    n_equal = n_invalid = 0
    for _ in range(4000):
        v = rval()
        text = style(v)
        if '"' not in text and rng.random() < 0.5:
            text = re.sub(r"-?\d+(?:\.\d+)?(?:[eE][+-]?\d+)?", respell, text)
        if rng.random() < 0.03:
            text = text[:rng.randint(0, len(text))]
        try:
            want = json.dumps(json.loads(text), ensure_ascii=False)
        except (json.JSONDecodeError, RecursionError):
            want = None
        got = canon(text)
        assert got == want, (text, want, got)
        n_equal += want is not None; n_invalid += want is None
    assert n_equal > 3500 and n_invalid > 20
    # what it must leave to CPython: duplicate keys (also when spelled differently), lone surrogates
    for text in ('{"a": 1, "a": 2}', '{"a": 1, "\\u0061": 2}', '"\\ud800"', '"\\udc00x"', '{"a": }', "[1, 2", "01", "- 1", "'a'"):
        assert canon(text) is None, text


def test_other_json_styles_take_the_native_lanes(monkeypatch):
    """Cells written compactly, indented, or with \\u escapes (what an annotation platform exports) are rewritten
    canonically inside the library and take the native lanes of steps 4, 5.5 and 6, with the CPython lanes'
    (and the oracle port's) results."""
    from deal_yolo_daya_b200 import labels
    from oracle import pipeline_port as port
    t = synth.make_table(3, 0, 240)
    rows = synth.table_to_rows(t)
    docs = [json.loads(r[1]) for r in rows]
    for d in docs[:40]:
        d["objects"][0]["name"] = "猫,狗"
    styles = [lambda d: json.dumps(d, separators=(",", ":"), ensure_ascii=True), lambda d: json.dumps(d, indent=2, ensure_ascii=False),
              lambda d: json.dumps(d, ensure_ascii=False)]
    cells = [styles[i % 3](d) for i, d in enumerate(docs)]
    df = pd.read_csv(io.StringIO(pd.DataFrame({"source": [r[0] for r in rows], ANN: cells}).to_csv(index=False)))

    def both(fn):
        out = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("DYD_NATIVE_INGEST", mode)
            out[mode] = fn()
        return out["1"], out["0"]
    (a, sa), (b, _) = both(lambda: (P.replace_ptlist_df(df)[0], dict(P.STATS)))
    assert sa["native_rows"] == len(df) and sa["slow_rows"] == 0
    assert a.to_csv(index=False) == b.to_csv(index=False) == port.replace_ptlist_df(df)[0].to_csv(index=False)
    rep = pd.read_csv(io.StringIO(a.to_csv(index=False)))
    lm = {synth.label_name(i): f"grp{i % 20:02d}" for i in range(synth.N_LABELS)}
    lm["猫"] = "animal"
    (ra, lane_a), (rb, lane_b) = both(lambda: (labels.remap_df(rep, lm), labels.LAST["remap_lane"]))
    assert (lane_a, lane_b) == ("native", "python")
    pd.testing.assert_frame_equal(ra[0], rb[0]); assert ra[1:] == rb[1:]
    l2c = {f"grp{g:02d}": f"cat{g // 5}" for g in range(20)}
    (xa, lane_a), (xb, lane_b) = both(lambda: (labels.split_df(ra[0], l2c), labels.LAST["split_lane"]))
    assert (lane_a, lane_b) == ("native", "python") and xa["summary"] == xb["summary"]
    for c in xa["categories"]:
        for part in ("train", "val", "test"):
            pd.testing.assert_frame_equal(xa["categories"][c][part], xb["categories"][c][part])
    pd.testing.assert_frame_equal(xa["split_counts"], xb["split_counts"])


def test_split_egress_cells_equal_json_dumps():
    """dyd_egress_split at the ABI level: for every position of the "objects" member (first, middle, last,
    only) and for cells rewritten from another style, the spliced one-object cell equals json.dumps of the
    document with "objects" replaced by the renamed object (processor.py:760-775)."""
    docs = [{"width": 10, "objects": [{"name": "a,b", "id": 1}, {"id": 2}, {"name": "c", "polygon": {"ptList": []}}], "height": 5},
            {"objects": [{"name": "x"}]},
            {"objects": [{"name": "y", "k": [1, 2]}], "w": 1},
            {"w": 1, "objects": [{"k": 0, "name": "z"}, 7, {"name": "中"}]}]
    texts = [json.dumps(docs[0], ensure_ascii=False), json.dumps(docs[1], separators=(",", ":")),
             json.dumps(docs[2], indent=2), json.dumps(docs[3], ensure_ascii=True)]
    ing = native.Ingest(pd.Series(texts), 2).names().objects()
    try:
        assert list(ing.status) == [0, 0, 0, 0] and ing.n_canon == 3 and list(ing.list_len) == [3, 1, 1, 3]
        labels_ = ["L0", 'q"uote', "标签"]
        esc = [json.dumps(t, ensure_ascii=False)[1:-1].encode() for t in labels_]
        off = np.zeros(len(esc) + 1, np.int64); off[1:] = np.cumsum([len(e) for e in esc])
        exp = []                                    # (cell, global dict-object index, label)
        for r, d in enumerate(docs):
            dict_objs = [o for o in d["objects"] if isinstance(o, dict)]
            for k, o in enumerate(dict_objs):
                if o.get("name"):
                    exp.append((r, int(ing.cell_off[r]) + k, (r + k) % 3, dict_objs[k]))
        out, oo = ing.egress_split([e[0] for e in exp], [e[1] for e in exp], [e[2] for e in exp], np.frombuffer(b"".join(esc), np.uint8), off)
        blob = out.tobytes()
        for i, (r, q, t, obj) in enumerate(exp):
            one = dict(obj); one["name"] = labels_[t]
            nd = {k: v for k, v in docs[r].items() if k != "objects"}; nd["objects"] = [one]
            assert blob[oo[i]:oo[i + 1]].decode() == json.dumps(nd, ensure_ascii=False)
    finally:
        ing.close()
