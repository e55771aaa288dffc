"""Native CSV writer == pandas.DataFrame.to_csv(index=False), byte for byte."""
from __future__ import annotations

import gzip
import io
import random
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

from deal_yolo_daya_b200 import native

G = Path(__file__).resolve().parent / "golden"


def both(df, tmp_path, enc="utf-8-sig"):
    a, b = tmp_path / "a.csv", tmp_path / "b.csv"
    native.to_csv(df, a, enc)
    df.to_csv(b, index=False, encoding=enc)
    return a.read_bytes(), b.read_bytes()


@pytest.mark.parametrize("name", ["processed_replaced_ptlist", "other_0.70_2", "dedup_first", "high_iou_0.70_2", "other_data_label_replaced"])
def test_golden_frames(tmp_path, name):
    df = pd.read_csv(io.StringIO(gzip.open(G / "expected" / f"{name}.csv.gz", "rt", encoding="utf-8").read()))
    x, y = both(df, tmp_path)
    assert x == y
    assert x.decode("utf-8-sig") == gzip.open(G / "expected" / f"{name}.csv.gz", "rt", encoding="utf-8").read()


def test_adversarial_strings_and_numbers(tmp_path):
    rng = random.Random(1)
    alphabet = ['a', 'b', ',', '"', '\n', '\r', ' ', '\t', "'", '中', '😀', ';', '\\', '\x00', '\x0b', '""', ',"']
    strs = ["".join(rng.choice(alphabet) for _ in range(rng.randint(0, 12))) for _ in range(3000)]
    strs[5] = None; strs[6] = ""; strs[7] = "nan"; strs[8] = '"'; strs[9] = ","
    f = np.array([rng.choice([0.0, -0.0, 1920.0, 0.1, 1e-5, 1e16, 1e22, 5e-324, 1.7976931348623157e308, float("nan"), float("inf"), -float("inf"),
                              rng.uniform(-1e6, 1e6), round(rng.uniform(0, 2000), 3)]) for _ in range(3000)])
    i = np.array([rng.choice([0, -1, 1, 2 ** 63 - 1, -2 ** 63, rng.randint(-10 ** 12, 10 ** 12)]) for _ in range(3000)], dtype=np.int64)
    b = np.array([rng.random() < 0.5 for _ in range(3000)])
    df = pd.DataFrame({"s": strs, "f": f, "i": i, "b": b, "o": pd.Series(strs[::-1], dtype=object), "weird,name": strs, 'q"': f})
    for enc in ("utf-8-sig", "utf-8"):
        x, y = both(df, tmp_path, enc)
        assert x == y
    # empty frame, frame with no rows kept
    x, y = both(df.iloc[0:0], tmp_path)
    assert x == y


def test_falls_back_to_pandas_for_uncovered_frames(tmp_path):
    one = pd.DataFrame({"a": ["", None, "x"]})                               # single column: csv quotes lone empty fields
    mixed = pd.DataFrame({"a": [1, "x", 2.5], "b": [1, 2, 3]})               # object column with non-str values
    dt = pd.DataFrame({"a": pd.to_datetime(["2024-01-01", "2024-01-02"]), "b": [1, 2]})
    i32 = pd.DataFrame({"a": np.array([1, 2], np.int32), "b": ["x", "y"]})
    for df in (one, mixed, dt, i32):
        x, y = both(df, tmp_path)
        assert x == y


def test_yolo_label_text_matches_python_formatting():
    """dyd_yolo_format vs the reference's f-string (processor.py:1052), incl. rounding ties, tiny,
    huge, signed zero, nan and inf values."""
    import numpy as np
    from deal_yolo_daya_b200 import native
    rng = np.random.RandomState(3)
    special = [0.0, -0.0, 0.5e-6, 1.5e-6, 2.5e-6, 0.1234565, 0.1234575, 1e-7, 0.9999995, 0.99999949999, 1.0, 123456.7890125,
               1e22, -3.25, float("nan"), float("inf"), float("-inf"), 5e-324, 0.0000005, 0.0000015]
    vals = np.concatenate([np.array(special), rng.rand(4000), rng.rand(200) * 1e6, np.round(rng.rand(800), 6) + 5e-7])
    vals = vals[: len(vals) // 4 * 4]
    n_box = len(vals) // 4
    counts = rng.randint(0, 5, size=n_box)
    counts = counts[np.cumsum(counts) <= n_box]
    img_off = np.zeros(len(counts) + 2, np.int64); img_off[1:-1] = np.cumsum(counts); img_off[-1] = n_box
    cls = rng.randint(0, 90, size=n_box).astype(np.int32)
    ok = (rng.rand(n_box) < 0.8).astype(np.uint8)
    text, off = native.yolo_label_texts(img_off, cls, vals, ok)
    for i in range(len(img_off) - 1):
        want = "\n".join(f"{cls[q]} {vals[4*q]:.6f} {vals[4*q+1]:.6f} {vals[4*q+2]:.6f} {vals[4*q+3]:.6f}"
                         for q in range(img_off[i], img_off[i + 1]) if ok[q])
        assert bytes(text[off[i]:off[i + 1]]).decode() == want


def test_native_permutation_is_numpys_legacy_permutation():
    """dyd_numpy_permutation == np.random.RandomState(seed).permutation(n), bit for bit (the order DataFrame.sample(frac=1,
    random_state=seed) gives the rows of a category, processor.py:800); other seeds and small n are numpy's own."""
    for seed in (0, 1, 42, 2024, 2 ** 32 - 1):
        for n in (4096, 4097, 65535, 65536, 65537, 300001, (1 << 20) + 3):
            assert np.array_equal(native.permutation(seed, n), np.random.RandomState(seed).permutation(n)), (seed, n)
    for seed in (42, np.int64(7), None.__class__ and 5):
        assert np.array_equal(native.permutation(seed, 100), np.random.RandomState(seed).permutation(100))
    df = pd.DataFrame({"a": np.arange(50000)})
    assert np.array_equal(df.sample(frac=1, random_state=42)["a"].to_numpy(), native.permutation(42, 50000))
    with pytest.raises((ValueError, TypeError)):
        native.permutation(-1, 10)
