"""The frames the drop-in keeps between steps (deal_yolo_daya_b200/tablecache.py) are exactly what pd.read_csv would give,
the row-selected / straight-to-file CSV writer equals DataFrame.to_csv, and the cache never changes a result.

Kernels are emulated by the CPU oracle (tests/oracle_kernels.py): this file checks HOST logic."""
from __future__ import annotations

import contextlib
import io
import os
import random
import subprocess
import sys
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

from deal_yolo_daya_b200 import native, processor as P, synth, tablecache
from tests.oracle_kernels import OracleKernels

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.skipif(not native._pandas_infers_arrow_str(), reason="needs pandas' Arrow-backed str dtype")


@pytest.fixture(autouse=True)
def _fresh(monkeypatch):
    monkeypatch.setattr(P, "KERNELS", OracleKernels())
    monkeypatch.delenv("DYD_TABLE_CACHE", raising=False)
    tablecache.clear()
    yield
    tablecache.clear()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def write_table(tmp, n=300, seed=5):
    t = synth.make_table(seed, 0, n)
    rows = synth.table_to_rows(t)
    merged, ref = tmp / "merged.csv", tmp / "ref.csv"
    pd.DataFrame(rows, columns=[P.COL_SRC, P.COL_ANN]).to_csv(merged, index=False, encoding="utf-8-sig")
    ids = [int(i) for i in t.url_id[::7]] + [10 ** 12 + k for k in range(40)]
    pd.DataFrame({P.COL_SRC: [synth.url_of(i) for i in ids], "note": ["r"] * len(ids)}).to_csv(ref, index=False, encoding="utf-8-sig")
    return merged, ref


def run_chain(tmp, merged, ref, thr=0.7):
    out = {}
    out["dedup"] = quiet(P.deduplicate_csv_by_source, str(merged), str(tmp / "dedup.csv"))
    out["filtered"] = quiet(P.remove_duplicates_between_csv, str(tmp / "dedup.csv"), str(ref), str(tmp / "filtered.csv"))
    out["rep"] = quiet(P.process_csv_replace_ptlist, str(tmp / "filtered.csv"), str(tmp / "rep.csv"), str(tmp / "exc.csv"))
    quiet(P.filter_by_box_count_and_iou, str(tmp / "rep.csv"), str(tmp / "hi.csv"), str(tmp / "other.csv"), 2, thr)
    return out


FILES = ["dedup.csv", "filtered.csv", "rep.csv", "exc.csv", "hi.csv", "other.csv"]


def test_chain_with_cache_equals_chain_without(tmp_path, monkeypatch):
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(); b.mkdir()
    ma, ra = write_table(a)
    mb, rb = write_table(b)
    hits0 = tablecache.STATS["hits"]
    ra_out = run_chain(a, ma, ra)
    assert tablecache.STATS["hits"] >= hits0 + 3, "steps 3, 4 and 5 should have found the previous step's frame"
    monkeypatch.setenv("DYD_TABLE_CACHE", "0")
    rb_out = run_chain(b, mb, rb)
    for f in FILES:
        assert (a / f).read_bytes() == (b / f).read_bytes(), f
    pd.testing.assert_frame_equal(ra_out["dedup"], rb_out["dedup"])
    pd.testing.assert_frame_equal(ra_out["filtered"], rb_out["filtered"])          # index labels of the surviving rows included
    assert ra_out["rep"] == {**rb_out["rep"], "excluded_output": str(a / "exc.csv")}


def test_every_remembered_frame_is_what_read_csv_returns(tmp_path):
    merged, ref = write_table(tmp_path)
    run_chain(tmp_path, merged, ref)
    seen = 0
    for f in FILES + ["merged.csv", "ref.csv"]:
        ent = tablecache.get(tmp_path / f)
        if ent is None:
            continue
        seen += 1
        want = pd.read_csv(tmp_path / f, encoding="utf-8-sig")
        pd.testing.assert_frame_equal(ent.frame, want, check_exact=True)
        assert list(ent.frame.dtypes) == list(want.dtypes)
    assert seen >= 6            # everything except the (empty) excluded file


def test_step5_reuses_step4_boxes_and_any_threshold_is_right(tmp_path, monkeypatch):
    merged, ref = write_table(tmp_path)
    run_chain(tmp_path, merged, ref, thr=0.7)
    ent = tablecache.get(tmp_path / "rep.csv")
    assert ent is not None and "boxes" in ent.extras
    parsed = []
    real = native.Ingest
    monkeypatch.setattr(native, "Ingest", lambda *a, **k: parsed.append(1) or real(*a, **k))
    for thr, mb in ((0.7, 2), (0.98, 2), (0.5, 3), (0.0, 1)):
        quiet(P.filter_by_box_count_and_iou, str(tmp_path / "rep.csv"), str(tmp_path / "hi2.csv"), str(tmp_path / "ot2.csv"), mb, thr)
        df = pd.read_csv(tmp_path / "rep.csv", encoding="utf-8-sig")
        got_hi = pd.read_csv(tmp_path / "hi2.csv", encoding="utf-8-sig")
        got_ot = pd.read_csv(tmp_path / "ot2.csv", encoding="utf-8-sig")
        assert len(got_hi) + len(got_ot) == len(df)
        # the file-level answer must equal the JSON-parsing DataFrame core
        tablecache.clear()
        hi, ot = P.filter_by_box_count_and_iou_df(df, mb, thr)
        pd.testing.assert_frame_equal(got_hi, hi.reset_index(drop=True), check_dtype=len(hi) > 0)    # a header-only file reads as object columns
        pd.testing.assert_frame_equal(got_ot, ot.reset_index(drop=True), check_dtype=len(ot) > 0)
        # put the step-4 entry back for the next parameters
        quiet(P.process_csv_replace_ptlist, str(tmp_path / "filtered.csv"), str(tmp_path / "rep.csv"), str(tmp_path / "exc.csv"))
        parsed.clear()
    assert not parsed


def test_a_file_changed_by_someone_else_is_read_from_disk(tmp_path):
    merged, ref = write_table(tmp_path)
    quiet(P.deduplicate_csv_by_source, str(merged), str(tmp_path / "dedup.csv"))
    assert tablecache.get(tmp_path / "dedup.csv") is not None
    df = pd.read_csv(tmp_path / "dedup.csv", encoding="utf-8-sig").iloc[:-3]
    df.to_csv(tmp_path / "dedup.csv", index=False, encoding="utf-8-sig")              # an external edit
    assert tablecache.get(tmp_path / "dedup.csv") is None
    out = quiet(P.remove_duplicates_between_csv, str(tmp_path / "dedup.csv"), str(ref), str(tmp_path / "filtered.csv"))
    assert len(out) <= len(df)
    again = pd.read_csv(tmp_path / "filtered.csv", encoding="utf-8-sig")
    assert set(again[P.COL_SRC]) <= set(df[P.COL_SRC])


@pytest.mark.parametrize("case", ["numeric_text", "na_string", "empty_string", "all_missing", "nul", "object_column", "int_name", "ok"])
def test_only_round_trip_safe_frames_are_remembered(tmp_path, case):
    n = 50
    base = {"source": [f"https://x/{i}.jpg" for i in range(n)], "k": [f"v{i}" for i in range(n)]}
    df = pd.DataFrame(base)
    if case == "numeric_text":
        df["k"] = pd.Series([str(i) for i in range(n)], dtype="str")         # comes back as int64
    elif case == "na_string":
        df.loc[7, "k"] = "NA"                                                # comes back as NaN
    elif case == "empty_string":
        df.loc[7, "k"] = ""
    elif case == "all_missing":
        df["k"] = pd.Series([None] * n, dtype="str")                         # comes back as float64 NaN
    elif case == "nul":
        df.loc[7, "k"] = "a\0b"
    elif case == "object_column":
        df["k"] = pd.Series([f"v{i}" for i in range(n)], dtype=object)
    elif case == "int_name":
        df = df.rename(columns={"k": 7})
    path = tmp_path / "t.csv"
    P._to_csv(df, path, "utf-8-sig")
    back = pd.read_csv(path, encoding="utf-8-sig")
    ent = tablecache.get(path)
    if case == "ok":
        assert ent is not None
        pd.testing.assert_frame_equal(ent.frame, back)
    else:
        assert ent is None, case
    # whatever the cache decided, the reading function returns what pandas returns
    pd.testing.assert_frame_equal(P._read_csv(path, encoding="utf-8-sig"), back)


def test_window_rule_is_checked_per_inference_chunk(tmp_path, monkeypatch):
    """A text column whose later dtype-inference chunk holds only number-like cells comes back mixed: not remembered."""
    monkeypatch.setattr(native, "_buffer_lines", lambda n_cols: 16)
    df = pd.DataFrame({"a": [f"v{i}" for i in range(16)] + [str(i) for i in range(16)], "b": ["x"] * 32})
    assert not native.roundtrip_safe(df)
    df2 = pd.DataFrame({"a": [f"v{i}" for i in range(32)], "b": ["x"] * 32})
    assert native.roundtrip_safe(df2) and native.roundtrip_safe(df2, check_cells=False)


def test_row_selected_writer_equals_pandas(tmp_path):
    rng = random.Random(3)
    alphabet = ['a', 'b', ',', '"', '\n', '\r', ' ', "'", '中', '😀', '""', ',"', 'xyzxyzxyzxyz' * 4, '"' * 40]
    n = 5000
    strs = ["".join(rng.choice(alphabet) for _ in range(rng.randint(0, 30))) for _ in range(n)]
    strs[5] = None; strs[6] = ""
    df = pd.DataFrame({"s": pd.Series(strs, dtype="str"), "f": np.random.RandomState(1).rand(n), "i": np.arange(n, dtype=np.int64) * 977 - 5000,
                       "b": np.arange(n) % 3 == 0, "t": pd.Series(strs[::-1], dtype="str")})
    for rows in (None, np.arange(0, n, 3), np.array([], np.int64), np.array([n - 1, 0, 17, 17, 4]), np.nonzero(df["b"].to_numpy())[0]):
        a, b = tmp_path / "a.csv", tmp_path / "b.csv"
        assert native.to_csv(df, a, "utf-8-sig", rows=rows)
        (df if rows is None else df.iloc[rows]).to_csv(b, index=False, encoding="utf-8-sig")
        assert a.read_bytes() == b.read_bytes()
    # append mode: no second BOM, optional header
    a, b = tmp_path / "ap_a.csv", tmp_path / "ap_b.csv"
    for k, (path, wr) in enumerate(((a, native.to_csv), (b, None))):
        for j, rows in enumerate((np.arange(10), np.arange(10, 30))):
            if wr:
                wr(df, path, "utf-8-sig", mode="a" if j else "w", header=not j, rows=rows)
            else:
                df.iloc[rows].to_csv(path, index=False, encoding="utf-8-sig", mode="a" if j else "w", header=not j)
    assert a.read_bytes() == b.read_bytes()
    with pytest.raises(OSError):
        native.to_csv(df, tmp_path / "no_such_dir" / "x.csv", "utf-8-sig")


def test_wide_and_scalar_writers_agree(tmp_path):
    """The AVX-512 quote-doubling copy and the scalar one produce the same file (the second process disables the wide forms)."""
    code = (
        "import sys, numpy as np, pandas as pd, random\n"
        f"sys.path.insert(0, {str(ROOT)!r})\n"
        "from deal_yolo_daya_b200 import native\n"
        "rng = random.Random(11)\n"
        "al = ['a', '\"', '\"\"', ',', 'json \"key\": 1.5, ', 'x' * 70, '\\n']\n"
        "s = [''.join(rng.choice(al) for _ in range(rng.randint(0, 40))) for _ in range(4000)]\n"
        "df = pd.DataFrame({'a': pd.Series(s, dtype='str'), 'b': pd.Series(s[::-1], dtype='str')})\n"
        "native.to_csv(df, sys.argv[1], 'utf-8-sig')\n"
    )
    outs = []
    for simd in ("0", "1"):
        path = tmp_path / f"w{simd}.csv"
        env = dict(os.environ, DYD_NO_SIMD=simd)
        subprocess.run([sys.executable, "-c", code, str(path)], check=True, env=env, timeout=300)
        outs.append(path.read_bytes())
    assert outs[0] == outs[1] and len(outs[0]) > 100000
    df = pd.read_csv(tmp_path / "w0.csv", encoding="utf-8-sig", keep_default_na=False)
    b = tmp_path / "p.csv"
    df.to_csv(b, index=False, encoding="utf-8-sig")
    assert b.read_bytes() == outs[0]


def test_float_source_column_groups_signed_zeros_like_pandas():
    df = pd.DataFrame({"source": [0.0, -0.0, 1.5, float("nan"), float("nan"), 1.5, -0.0], "v": list("abcdefg")})
    for keep in ("first", "last", False):
        want = df.drop_duplicates(subset=["source"], keep=keep, ignore_index=True)
        pd.testing.assert_frame_equal(P.deduplicate_df(df, keep), want)


def test_int_coordinates_beyond_2_pow_25_are_computed_like_cpython(tmp_path):
    """CPython's IoU on int coordinates is exact big-int arithmetic; fp64 agrees only up to |int| <= 2^25.  Rows beyond that
    bound take the CPython lane in step 5 (native parser and Python parser alike), and step 4 does not hand its fp64 boxes over."""
    import json
    big = 2 ** 27

    def cell(boxes):
        return json.dumps({"width": 10, "height": 10, "objects": [
            {"name": "a", "polygon": {"ptList": [{"x": b[0], "y": b[1]}, {"x": b[2], "y": b[1]}, {"x": b[2], "y": b[3]}, {"x": b[0], "y": b[3]}]}} for b in boxes]})
    rows = [cell([(0, 0, big + 3, big + 1), (1, 0, big + 3, big + 1)]),           # IoU just below / above thresholds only exact arithmetic separates
            cell([(0, 0, 100, 100), (0, 0, 100, 70)]),
            cell([(0.5, 0.5, 10.5, 10.5), (0.5, 0.5, 10.5, 9.5)])]
    df = pd.DataFrame({P.COL_SRC: [f"u{i}" for i in range(len(rows))], P.COL_ANN: rows})
    src = tmp_path / "in.csv"
    df.to_csv(src, index=False, encoding="utf-8-sig")
    quiet(P.process_csv_replace_ptlist, str(src), str(tmp_path / "rep.csv"), str(tmp_path / "exc.csv"))
    ent = tablecache.get(tmp_path / "rep.csv")
    assert ent is not None and "boxes" not in ent.extras, "fp64 boxes of a table with huge int coordinates must not be reused"
    rep = pd.read_csv(tmp_path / "rep.csv", encoding="utf-8-sig")

    def reference_mask(cells, mb, thr):          # processor.py:328-376 restated on Python objects (exact int arithmetic)
        out = []
        for c in cells:
            bx = []
            for o in json.loads(c)["objects"]:
                p, q = o["polygon"]["ptList"]
                bx.append((min(p["x"], q["x"]), min(p["y"], q["y"]), max(p["x"], q["x"]), max(p["y"], q["y"])))
            hit = False
            for i in range(len(bx)):
                for j in range(i + 1, len(bx)):
                    a, b = bx[i], bx[j]
                    iw = max(0, min(a[2], b[2]) - max(a[0], b[0])); ih = max(0, min(a[3], b[3]) - max(a[1], b[1]))
                    inter = iw * ih
                    if inter == 0:
                        continue
                    union = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
                    hit = hit or (union != 0 and inter / union >= thr)
            out.append(len(bx) >= mb and hit)
        return np.array(out)
    exact = ((big + 2) * (big + 1)) / ((big + 3) * (big + 1))
    for thr in (0.7, exact, float(np.nextafter(exact, 1.0)), float(np.nextafter(exact, 0.0))):
        want = reference_mask(rep[P.COL_NEW], 2, thr)
        assert np.array_equal(P.high_iou_mask(rep[P.COL_NEW], 2, thr), want), thr
        assert P.STATS["slow_rows"] >= 1
