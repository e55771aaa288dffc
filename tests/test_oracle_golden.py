"""The CPU oracle against the fixtures produced by the real reference (tests/golden/).

Pins (a) the row-level port oracle/pipeline_port.py byte-for-byte on every step's output,
(b) the CSR-level oracles oracle_np.py / dyd_oracle.c through the ingest layer.
"""
from __future__ import annotations

import gzip
import io
import json
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

from deal_yolo_daya_b200 import ingest
from oracle import oracle_c, oracle_np, pipeline_port as port
from tests import tables

G = Path(__file__).resolve().parent / "golden"
ANN, NEW = port.COL_ANN, port.COL_NEW


def gz_text(p: Path) -> str:
    return gzip.open(p, "rt", encoding="utf-8").read()


def read_csv_text(text: str) -> pd.DataFrame:
    return pd.read_csv(io.StringIO(text))


def csv_text(df: pd.DataFrame) -> str:
    return df.to_csv(index=False)


def inp(name):
    return read_csv_text(gz_text(G / "inputs" / name))


def exp_text(name):
    return gz_text(G / "expected" / name)


SUMMARY = json.loads((G / "expected" / "summary.json").read_text(encoding="utf-8"))


def test_versions_recorded():
    assert SUMMARY["versions"]["pandas"]


@pytest.mark.parametrize("keep", ["first", "last"])
def test_port_dedup(keep):
    out = port.dedup_df(inp("merged.csv.gz"), keep)
    assert csv_text(out) == exp_text(f"dedup_{keep}.csv.gz")


def test_port_ref_filter():
    main = read_csv_text(exp_text("dedup_first.csv.gz"))
    out = port.ref_filter_df(main, inp("reference.csv.gz"))
    assert csv_text(out) == exp_text("filtered_main.csv.gz")


def test_port_replace_ptlist():
    df = read_csv_text(exp_text("filtered_main.csv.gz"))
    res, exc = port.replace_ptlist_df(df)
    assert csv_text(res) == exp_text("processed_replaced_ptlist.csv.gz")
    assert csv_text(exc) == exp_text("processed_excluded.csv.gz")
    assert len(res) == SUMMARY["replace"]["filtered_rows"] and len(exc) == SUMMARY["replace"]["excluded_rows"]


@pytest.mark.parametrize("thr,mb", [(0.7, 2), (0.98, 2), (0.7, 3), (0.0, 1)])
def test_port_iou(thr, mb):
    df = read_csv_text(exp_text("processed_replaced_ptlist.csv.gz"))
    hi, ot = port.iou_split_df(df, mb, thr)
    tag = f"{thr:.2f}_{mb}"
    assert csv_text(hi) == exp_text(f"high_iou_{tag}.csv.gz")
    assert csv_text(ot) == exp_text(f"other_{tag}.csv.gz")


def test_port_crash_on_none_coordinate():
    want = json.loads((G / "expected" / "crash_none.json").read_text())["raises"]
    assert want == "TypeError"
    with pytest.raises(TypeError):
        port.replace_ptlist_df(inp("crash_none.csv.gz"))


def _mapping():
    return pd.DataFrame(json.loads((G / "inputs" / "mapping.json").read_text(encoding="utf-8")))


def _rules():
    return pd.DataFrame(json.loads((G / "inputs" / "rules.json").read_text(encoding="utf-8")))


def test_port_remap():
    df = read_csv_text(exp_text("other_0.70_2.csv.gz"))
    lm = port.mapping_from_frame(_mapping())
    out, summary, diffs, unmatched = port.remap_df(df, lm)
    assert csv_text(out) == exp_text("other_data_label_replaced.csv.gz")
    assert summary == SUMMARY["remap"]["summary"]
    assert csv_text(pd.DataFrame(diffs)) == exp_text("remap_diff.csv.gz")
    um = pd.DataFrame([{"标签": k, "数量": v} for k, v in unmatched.items()]).sort_values("数量", ascending=False)
    assert csv_text(um) == exp_text("remap_unmatched.csv.gz")
    assert diffs[:30] == SUMMARY["remap"]["sample_diff"]


def test_port_split():
    df = read_csv_text(exp_text("other_data_label_replaced.csv.gz"))
    l2c = port.rules_from_frame(_rules())
    res = port.split_df(df, l2c)
    assert res["summary"] == SUMMARY["split"]
    assert [f"{c}.xlsx" for c in res["categories"]] == SUMMARY["split_files"]
    for cat, parts in res["categories"].items():
        for name, part in parts.items():
            assert csv_text(part) == exp_text(f"split__{cat}__{name}.csv.gz"), (cat, name)
    assert csv_text(res["unclassified"]) == exp_text("split__unclassified__Sheet1.csv.gz")
    assert csv_text(res["split_counts"]) == exp_text("split__split_counts__Sheet1.csv.gz")


# ---------------------------------------------------------------- CSR-level oracles vs golden
def _corners_from_text(text):
    """(min_x, min_y, max_x, max_y) per dict object out of a replaced cell; None for a null bbox."""
    out = []
    for obj in json.loads(text)["objects"]:
        p, q = obj["polygon"]["ptList"]
        out.append(None if p["x"] is None else (p["x"], p["y"], q["x"], q["y"]))
    return out


def _same_double(a, b):
    return np.float64(a).tobytes() == np.float64(b).tobytes()


@pytest.mark.parametrize("impl", ["np", "c"])
def test_csr_bbox_matches_reference_cells(impl):
    df = read_csv_text(exp_text("processed_replaced_ptlist.csv.gz"))
    batch = ingest.parse_polygons(df[ANN].tolist())
    assert batch.hostlane.sum() == 0
    fold = oracle_np.bbox_fold if impl == "np" else oracle_c.bbox_fold
    pts, valid, arg = fold(batch.poly_off, batch.xy)
    q = 0
    checked = 0
    for r, text in enumerate(df[NEW].tolist()):
        if not isinstance(text, str):
            assert batch.docs[r] is None
            continue
        want = _corners_from_text(text)
        assert len(want) == batch.img_off[r + 1] - batch.img_off[r]
        for w in want:
            if w is None:
                assert valid[q] == 0
            else:
                assert valid[q] == 1
                for k in range(4):
                    assert _same_double(pts[4 * q + k], w[k]), (r, q, k, pts[4 * q + k], w[k])
                # arg indices select the original JSON numbers (int vs float text)
                good = batch.points[q]
                src = [good[arg[4 * q]]["x"], good[arg[4 * q + 1]]["y"], good[arg[4 * q + 2]]["x"], good[arg[4 * q + 3]]["y"]]
                assert [type(s) for s in src] == [type(v) for v in w] and all(_same_double(s, v) for s, v in zip(src, w))
                checked += 1
            q += 1
    assert q == batch.n_obj and checked > 300


@pytest.mark.parametrize("impl", ["np", "c"])
@pytest.mark.parametrize("thr,mb", [(0.7, 2), (0.98, 2), (0.7, 3), (0.0, 1)])
def test_csr_iou_matches_reference_split(impl, thr, mb):
    df = read_csv_text(exp_text("processed_replaced_ptlist.csv.gz"))
    hi = read_csv_text(exp_text(f"high_iou_{thr:.2f}_{mb}.csv.gz"))
    batch = ingest.parse_boxes(df[NEW].tolist())
    assert not batch.host_rows
    f = oracle_np.iou_filter if impl == "np" else oracle_c.iou_filter
    high, count = f(batch.img_off, batch.pts, batch.valid, mb, thr)
    got = df[high.astype(bool)]
    assert csv_text(got) == csv_text(hi)


@pytest.mark.parametrize("impl", ["np", "c"])
@pytest.mark.parametrize("keep", ["first", "last"])
def test_csr_dedup_matches_reference(impl, keep):
    df = inp("merged.csv.gz")
    off, data, null = ingest.pack_strings(df["source"].tolist())
    keys = oracle_c.hash_strings_buf(off, data)
    assert np.array_equal(keys, oracle_np.hash_strings([s if isinstance(s, str) else "" for s in df["source"].tolist()]))
    f = oracle_np.dedup if impl == "np" else oracle_c.dedup
    km, rep = f(keys, null, keep)
    got = df[km.astype(bool)].reset_index(drop=True)
    assert csv_text(got) == exp_text(f"dedup_{keep}.csv.gz")


@pytest.mark.parametrize("impl", ["np", "c"])
def test_csr_antijoin_matches_reference(impl):
    main = read_csv_text(exp_text("dedup_first.csv.gz")); ref = inp("reference.csv.gz")
    mo, md, mn = ingest.pack_strings(main["source"].tolist())
    ro, rd, rn = ingest.pack_strings(ref["source"].tolist())
    mk, rk = oracle_c.hash_strings_buf(mo, md), oracle_c.hash_strings_buf(ro, rd)
    f = oracle_np.antijoin if impl == "np" else oracle_c.antijoin
    km, rr = f(mk, mn, rk, rn)
    assert csv_text(main[km.astype(bool)]) == exp_text("filtered_main.csv.gz")
    for r in np.where(km == 0)[0]:
        assert ref["source"][rr[r]] == main["source"][r]


# ---------------------------------------------------------------- the two CSR oracles agree with each other
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_np_and_c_oracles_agree(seed):
    img_off, poly_off, xy = tables.random_polygon_table(seed, 60)
    p1, v1, a1 = oracle_np.bbox_fold(poly_off, xy)
    p2, v2, a2 = oracle_c.bbox_fold(poly_off, xy)
    assert p1.tobytes() == p2.tobytes() and np.array_equal(v1, v2) and np.array_equal(a1, a2)
    for thr, mb in [(0.7, 2), (0.98, 2), (0.0, 1), (0.5, 4)]:
        h1, c1 = oracle_np.iou_filter(img_off, p1, v1, mb, thr)
        h2, c2 = oracle_c.iou_filter(img_off, p2, v2, mb, thr)
        assert np.array_equal(h1, h2) and np.array_equal(c1, c2)


def test_np_and_c_oracles_agree_on_edges():
    img_off, poly_off, xy = tables.edge_polygon_table()
    p1, v1, a1 = oracle_np.bbox_fold(poly_off, xy)
    p2, v2, a2 = oracle_c.bbox_fold(poly_off, xy)
    assert p1.tobytes() == p2.tobytes() and np.array_equal(v1, v2) and np.array_equal(a1, a2)
    h1, _ = oracle_np.iou_filter(img_off, p1, v1, 2, 0.7)
    h2, _ = oracle_c.iou_filter(img_off, p2, v2, 2, 0.7)
    assert np.array_equal(h1, h2)
    # the hand-made expectations of SURVEY §8a
    assert list(h1[5:11]) == [1, 0, 0, 1, 1, 0]


def test_split_assign_equals_dataframe_sample():
    cat_off = np.array([0, 7, 7, 30, 130], np.int64)
    split, pos = oracle_np.split_assign(cat_off, 0.8, 0.1, 0.1, 42)
    for c in range(4):
        a, b = cat_off[c], cat_off[c + 1]
        if b == a:
            continue
        df = pd.DataFrame({"j": np.arange(b - a)}).sample(frac=1, random_state=42).reset_index(drop=True)
        n = b - a; ntr, nva = int(n * 0.8), int(n * 0.1)
        assert list(df["j"][:ntr]) == [j for j in np.argsort(pos[a:b]) if split[a + j] == 0]
        assert set(df["j"][ntr:ntr + nva]) == {j for j in range(n) if split[a + j] == 1}
