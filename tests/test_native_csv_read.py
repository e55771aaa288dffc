"""Native CSV ingest (csrc/csv_read.cpp + native.read_csv) against pd.read_csv itself: same frame,
same dtypes, same column labels, or the same exception.  CPU-only: the tokenizer is host code."""
import io
import os
import random
import warnings
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

from deal_yolo_daya_b200 import native

GOLDEN = Path(__file__).parent / "golden" / "inputs"


@pytest.fixture(autouse=True)
def _always_native(monkeypatch):
    monkeypatch.setenv("DYD_CSV_NATIVE_MIN_BYTES", "0")
    if not native._pandas_infers_arrow_str():
        pytest.skip("this pandas does not infer the Arrow-backed str dtype; native.read_csv defers to pandas")


def _both(tmp_path, text, name="t.csv", encoding="utf-8-sig", binary=False):
    p = tmp_path / name
    if binary:
        p.write_bytes(text)
    else:
        p.write_text(text, encoding="utf-8", newline="")
    before = native._READ_STATS["native"]

    def run(fn):
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                return fn(), None
        except Exception as e:  # noqa: BLE001
            return None, type(e).__name__
    exp, exp_err = run(lambda: pd.read_csv(p, encoding=encoding))
    got, got_err = run(lambda: native.read_csv(str(p), encoding=encoding))
    took_native = native._READ_STATS["native"] > before
    assert exp_err == got_err
    if exp_err is None:
        pd.testing.assert_frame_equal(got, exp)
        assert [str(d) for d in got.dtypes] == [str(d) for d in exp.dtypes]
        assert list(got.columns) == list(exp.columns)
    return took_native


CASES = {
    "basic": ('source,标注,n\nhttp://a/1,"{""k"": [1, 2]}",5\nhttp://a/2,"{""k"": ""x,y""}",7\n', True),
    "crlf_and_embedded_newlines": ('a,b\r\nhello,"x\r\ny"\r\nworld,"p\nq"\r\n', True),
    "na_strings": ('a,b\nNA,foo\n,bar\nnull,"NA"\nhi,"n/a"\n<NA>,#N/A\n', True),
    "blank_and_whitespace_lines": ('a,b\n\nx,y\n   \n\t\np,q\n\n', True),
    "quote_inside_unquoted_and_tail_after_quote": ('a,b\nx"y,5" pipe\n"ab"c,"d""e"f\n', True),
    "bom": ('﻿a,b\nxx,yy\n', True),
    "no_final_newline": ('a,b\nxx,yy', True),
    "numeric_bool_columns_delegated": ('id,url,score,flag\n1,http://x,0.5,True\n2,http://y,,False\n', True),
    "all_missing_column": ('a,b\nxx,\nyy,\n', True),
    "duplicate_and_unnamed_headers": ('a,a,\nxx,yy,zz\n', True),
    "leading_whitespace_cell": ('a,b\n  x,y\n', True),
    "unicode": ('a,b\n标注,héllo\nzz,"多\n行"\n', True),
    "inf_like_text": ('a,b\ninfo,1\nnope,2\n', True),
    "ragged_short_row": ('a,b\nx\ny,z\n', False),
    "implicit_index": ('a,b\nx,y,z\nq,w,e\n', False),
    "bare_cr_terminators": ('a,b\rx,y\rp,q\r', False),
    "eof_inside_quotes": ('a,b\nx,"unterminated\n', False),
    "header_only": ('a,b\n', False),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_edge_cases_match_pandas(tmp_path, name):
    text, native_expected = CASES[name]
    took_native = _both(tmp_path, text)
    assert took_native == native_expected


def test_empty_file_and_invalid_utf8_raise_like_pandas(tmp_path):
    assert _both(tmp_path, b"", binary=True) is False
    assert _both(tmp_path, b"a,b\nx,\xff\xfe\n", name="bad.csv", binary=True) is False
    assert _both(tmp_path, b"a,b\nx,\xed\xa0\x80\n", name="surrogate.csv", binary=True) is False
    assert _both(tmp_path, b"a,b\nx\x00y,z\n", name="nul.csv", binary=True) is False


def test_other_encodings_and_keywords_go_to_pandas(tmp_path):
    p = tmp_path / "g.csv"
    p.write_bytes("a,b\n标,注\n".encode("gbk"))
    before = dict(native._READ_STATS)
    got = native.read_csv(str(p), encoding="gbk")
    pd.testing.assert_frame_equal(got, pd.read_csv(p, encoding="gbk"))
    got = native.read_csv(str(p), encoding="gbk", nrows=1)
    assert native._READ_STATS["native"] == before["native"] and native._READ_STATS["pandas"] == before["pandas"] + 2
    assert len(got) == 1


def test_golden_inputs_read_identically(tmp_path):
    import gzip
    n = 0
    for src in sorted(GOLDEN.glob("*.csv*")):
        raw = gzip.open(src, "rb").read() if src.suffix == ".gz" else src.read_bytes()
        assert _both(tmp_path, raw, name=f"g{n}.csv", binary=True) in (True, False)
        n += 1
    assert n > 0


def test_fuzz_against_pandas(tmp_path):
    rng = random.Random(20260101)
    atoms = [",", ",", '"', '""', "\n", "\n", "\r\n", " ", "\t", "a", "b", "http://x", '{"k": 1}', "1", "2.5", "-3", "NA",
             "null", "nan", "True", "é", "标", "x y", "'", "e5", "inf", "0x1", "1_0"]
    n_native = 0
    for it in range(400):
        if rng.random() < 0.5:
            ncols, rows = rng.randint(1, 4), rng.randint(1, 8)
            lines = []
            for _ in range(rows + 1):
                cells = []
                for _ in range(ncols):
                    cell = "".join(rng.choice(atoms[6:]) for _ in range(rng.randint(0, 3)))
                    if rng.random() < 0.3 or any(ch in cell for ch in ',"\n\r'):
                        if rng.random() < 0.9:
                            cell = '"' + cell.replace('"', '""') + '"'
                    cells.append(cell)
                lines.append(",".join(cells))
            eol = rng.choice(["\n", "\r\n"])
            text = eol.join(lines) + (eol if rng.random() < 0.8 else "")
        else:
            text = "".join(rng.choice(atoms) for _ in range(rng.randint(1, 40)))
        if rng.random() < 0.1:
            text = "﻿" + text
        n_native += bool(_both(tmp_path, text, name=f"f{it % 4}.csv"))
    assert n_native > 100


def test_chunked_dtype_inference_boundary(tmp_path):
    """pandas infers dtypes per chunk of _buffer_lines(n_cols) rows and concatenates the chunks: a column that
    is numeric for a whole chunk and text later comes back mixed.  The native reader hands exactly those
    columns to pandas chunk by chunk and must return the very same frame."""
    ncols = 16
    w = native._buffer_lines(ncols)
    assert w == 32768
    head = ",".join(f"c{k}" for k in range(ncols))
    for k_numeric in (w, w - 1, 0):
        rows = [head]
        for r in range(w + 5):
            rows.append(",".join(["http://t"] * (ncols - 1) + ["1" if r < k_numeric else "word"]))
        assert _both(tmp_path, "\n".join(rows) + "\n", name=f"chunk{k_numeric}.csv") is True


def test_mixed_columns_across_chunks_match_pandas(tmp_path):
    """A wide file (400 columns -> 2048-row chunks) whose typed columns change character from chunk to chunk:
    int -> float, int -> text, a missing chunk, bool -> int, text starting with 't' ...: same dtypes and the same
    Python element types as pd.read_csv, chunk for chunk."""
    ncols = 400
    w = native._buffer_lines(ncols)
    n = 3 * w + 37
    kinds = {
        "text": lambda i: f"http://t/{i}", "int": lambda i: str(i), "float": lambda i: f"{i / 7:.3f}",
        "int_then_float": lambda i: str(i) if i < w else f"{i:.1f}", "int_then_text": lambda i: str(i) if i < 2 * w else f"word{i}",
        "text_then_int": lambda i: f"word{i}" if i < w else str(i), "int_with_na_chunk": lambda i: "" if w <= i < 2 * w else str(i),
        "bool_then_int": lambda i: ("True" if i % 2 else "False") if i < w else str(i), "all_na": lambda i: "",
        "na_then_text": lambda i: "" if i < w else "word", "text_t": lambda i: f"text{i}",
        "float_na_text": lambda i: "1.5" if i < w else ("" if i < 2 * w else "zz"),
    }
    names = list(kinds)
    cols = {f"c{k}": ([kinds[names[k % len(names)]](i) for i in range(n)] if k < 2 * len(names) else ["x"] * n) for k in range(ncols)}
    p = tmp_path / "wide.csv"
    pd.DataFrame(cols).to_csv(p, index=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exp = pd.read_csv(p)
        before = native._READ_STATS["native"]
        got = native.read_csv(str(p))
    assert native._READ_STATS["native"] == before + 1
    pd.testing.assert_frame_equal(got, exp)
    assert [str(a) for a in got.dtypes] == [str(a) for a in exp.dtypes]
    for k in range(len(names)):
        a, b = got[f"c{k}"], exp[f"c{k}"]
        assert [type(x) for x in a.iloc[[0, w, 2 * w, n - 1]]] == [type(x) for x in b.iloc[[0, w, 2 * w, n - 1]]], names[k]


def test_parallel_stitch_with_quoted_newlines(tmp_path, monkeypatch):
    """Chunks cut inside quoted fields: both hypotheses are scanned and stitched."""
    monkeypatch.setenv("DYD_INGEST_THREADS", "8")
    rng = random.Random(7)
    rows = ["source,标注"]
    for r in range(90000):
        body = "\n".join('{"k": %d, "s": "a,b"}' % rng.randint(0, 9) for _ in range(rng.randint(1, 6)))
        rows.append(f'http://h/{r},"{body.replace(chr(34), chr(34) * 2)}"')
    text = "\n".join(rows) + "\n"
    assert len(text) > 8 * (1 << 20)
    assert _both(tmp_path, text) is True


def test_wide_scanner_equals_bytewise_scanner_and_pandas(tmp_path, monkeypatch):
    """The AVX-512 tokenizer (quote parity by carry-less multiply) against the byte-wise state machine and against pandas on
    texts long enough to cross many 64-byte blocks and several chunks, with quotes, doubled quotes, separators and line feeds
    at every alignment; texts it must decline (quotes inside unquoted fields, text behind a closing quote, CR) included."""
    from deal_yolo_daya_b200 import simd_available
    if not simd_available():
        pytest.skip("host CPU without AVX-512 VBMI2: only the byte-wise scanner exists here")
    rng = random.Random(77)
    pieces = ["a", "bb", "http://x/1.jpg", '{"k": [1, 2.5], "name": "v"}', "标注", "é", "x" * 61, "y" * 64, "z" * 130, '"', '""', '"""', ",", "\n", " ",
              '{"objects": [{"name": "' + "q" * 50 + '", "p": [{"x": 1.5, "y": 2}]}]}']
    wide_before = native._READ_STATS["wide"]
    took_wide = declined = 0
    for it in range(300):
        ncols = rng.randint(2, 4)
        rows = rng.randint(1, 30)
        odd_file = rng.random() < 0.3               # some files hold cells the wide scanner must decline
        lines = [",".join(f"c{j}" for j in range(ncols))]
        for _ in range(rows):
            cells = []
            for _ in range(ncols):
                cell = "".join(rng.choice(pieces) for _ in range(rng.randint(0, 5)))
                mode = rng.random() if odd_file else 0.0
                if mode < 0.93:                         # what csv.writer would write
                    if any(ch in cell for ch in ',"\n'):
                        cell = '"' + cell.replace('"', '""') + '"'
                    elif rng.random() < 0.2:
                        cell = '"' + cell + '"'
                elif mode < 0.96:
                    cell = cell.replace(",", ";").replace("\n", " ")           # raw: quotes inside an unquoted field
                else:
                    cell = '"' + cell.replace('"', '""') + '"' + rng.choice(["tail", " ", ""])   # text behind the closing quote
                cells.append(cell)
            lines.append(",".join(cells))
        text = "\n".join(lines) + ("\n" if rng.random() < 0.8 else "")
        if rng.random() < 0.05:
            text = text.replace("\n", "\r\n")
        if rng.random() < 0.2:
            text = "﻿" + text
        monkeypatch.setenv("DYD_CSV_CHUNK_MIN", str(rng.choice([1, 7, 64, 100, 1000, 1 << 20])))
        before = native._READ_STATS["wide"]
        monkeypatch.setenv("DYD_CSV_WIDE", "1")
        took = _both(tmp_path, text, name="w.csv")
        was_wide = native._READ_STATS["wide"] > before
        a = native.read_csv(str(tmp_path / "w.csv"), encoding="utf-8-sig") if took else None
        monkeypatch.setenv("DYD_CSV_WIDE", "0")
        took2 = _both(tmp_path, text, name="w.csv")
        if took and took2:
            b = native.read_csv(str(tmp_path / "w.csv"), encoding="utf-8-sig")
            pd.testing.assert_frame_equal(a, b)
        took_wide += was_wide
        declined += (took2 and not was_wide)
    assert took_wide > 150 and declined > 20, (took_wide, declined)
    assert native._READ_STATS["wide"] > wide_before
