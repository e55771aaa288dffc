"""Inputs of the round-2 golden cases (merge, YOLO dataset writer, summaries): built the same way by the generator
(tests/golden/make_golden_r2.py, around the unmodified reference) and by the tests (around the drop-in)."""
from __future__ import annotations

import contextlib
import gzip
import io
import json
from pathlib import Path

import numpy as np
import pandas as pd

G = Path(__file__).resolve().parent / "golden"
ANN = "结果字段-目标检测标签配置"
NEW = "新_结果字段-目标检测标签配置"


def gz_text(p):
    return gzip.open(p, "rt", encoding="utf-8").read()


@contextlib.contextmanager
def sorted_glob():
    """Path.glob in sorted order: the merge step takes its inputs in filesystem order (reference :36), which differs between
    machines; the fixture pins one order for both the reference run and the test."""
    orig = Path.glob

    def glob(self, pattern, **k):
        return iter(sorted(orig(self, pattern, **k)))
    Path.glob = glob
    try:
        yield
    finally:
        Path.glob = orig


def write_merge_inputs(folder: Path):
    """Four inputs: plain text columns; an extra column + an int column with a gap; an empty file (read failure); a file whose
    numeric column changes type between chunks of 4 rows."""
    folder.mkdir(parents=True, exist_ok=True)
    doc = lambda i: json.dumps({"width": 640, "height": 480, "objects": [{"name": f"猫{i}", "polygon": {"ptList": [{"x": i, "y": 1.5}, {"x": 9, "y": 7}]}}]},  # noqa: E731
                               ensure_ascii=False)
    a = pd.DataFrame({"source": [f"https://x/{i}.jpg" for i in range(6)], ANN: [doc(i) for i in range(6)]})
    a.loc[2, ANN] = np.nan
    a.loc[4, "source"] = 'quoted "name", with comma.jpg'
    a.to_csv(folder / "a_first.csv", index=False, encoding="utf-8-sig")
    b = pd.DataFrame({"source": [f"b{i}" for i in range(5)], ANN: [doc(i + 10) for i in range(5)], "extra": [1, 2, None, 4, 5], "flag": [True, False, True, True, False]})
    b.to_csv(folder / "b_second.csv", index=False, encoding="utf-8-sig")
    (folder / "c_empty.csv").write_bytes(b"")
    d = pd.DataFrame({"source": [f"d{i}" for i in range(10)], "score": ["1", "2", "3", "4", "5.5", "6", "7", "8", "x", "10"]})
    d.to_csv(folder / "d_chunks.csv", index=False, encoding="utf-8-sig")
    (folder / "notes.txt").write_text("not a csv")


def yolo_books(img_dir: Path):
    """Two category workbooks from the golden split sheets, `source` pointing at local stand-in image files, with rows edited to
    reach every skip reason of processor.py:1009-1069."""
    img_dir.mkdir(parents=True, exist_ok=True)
    books = {}
    for cat in ("catA", "catB"):
        sheets = {}
        for split in ("train", "val", "test"):
            df = pd.read_csv(io.StringIO(gz_text(G / "expected" / f"split__{cat}__{split}.csv.gz")))
            names = []
            for k, src in enumerate(df["source"]):
                stem = str(src).rsplit("/", 1)[-1] if isinstance(src, str) else f"none{k}.jpg"
                p = img_dir / stem
                p.write_bytes(b"not really a jpeg " + stem.encode())
                names.append(str(p))
            df["source"] = names
            sheets[split] = df
        books[cat] = sheets
    t = books["catA"]["train"]
    if len(t) >= 8:
        t.loc[0, "source"] = np.nan                                   # 缺少source
        t.loc[1, "分类标签"] = np.nan                                   # 缺少或无效分类标签
        t.loc[2, "width"] = 0                                          # 缺少图像尺寸 (the stand-in file is no image)
        t.loc[3, "source"] = str(img_dir / "missing_file.jpg")          # 图片下载失败
        t.loc[4, "分类标签"] = "no_such_label_in_row"                    # 无匹配标签框 (label is a class, but no box carries it)
        t.loc[5, NEW] = json.dumps({"width": 10, "height": 10, "objects": [
            {"name": t.loc[5, "分类标签"], "polygon": {"ptList": [{"x": 3, "y": 1}, {"x": 3, "y": 9}]}}]}, ensure_ascii=False)   # 标注框无效
        t.loc[6, NEW] = json.dumps({"width": 10, "height": 10, "objects": [
            {"name": t.loc[6, "分类标签"], "polygon": {"ptList": [{"x": 1, "y": 2}, {"x": 8.5, "y": 7}]}},
            {"name": t.loc[6, "分类标签"], "polygon": {"ptList": [{"x": 5, "y": 5}, {"x": 5, "y": 6}]}},     # zero width: skipped line
            {"name": t.loc[6, "分类标签"], "polygon": {"ptList": [{"x": 2, "y": 9}, {"x": 0, "y": 1}, {"x": 4, "y": 3}]}}]}, ensure_ascii=False)
    v = books["catB"]["val"]
    if len(v) >= 2:
        v.loc[0, "width"] = 1919.5                                     # float sizes
        v.loc[1, NEW] = np.nan                                         # falls back to the original column's polygons
    return books
