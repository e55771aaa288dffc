"""The callers either side of the hot path: CSV merge in front of it, YOLO dataset writer and
the two summaries behind it.  Same names, signatures, printed lines, return values and on-disk results as the reference's
``core/processor.py`` (merge :26-109, unclassified summary :833-891, dataset writer :893-1087, label-count summary
:1089-1163); ``processor.py`` of this package re-exports them so the drop-in surface stays one module.

What runs where
  * files, folders, Excel sheets, JSON text, image bytes, ``data.yaml``: host, in the reference's own order (the order of
    side effects is observable: the `skipped` sheet, `resume`, the counters);
  * the per-box arithmetic of the label files -- ``(x1+x2)/2/W, (y1+y2)/2/H, bw/W, bh/H`` with the ``bw <= 0 or bh <= 0``
    skip rule -- and every histogram of the summaries: device (``dyd_yolo_normalise``, ``dyd_label_presence``,
    ``dyd_label_hist`` through ``processor.KERNELS``); the ``%.6f`` label text is written by the native formatter
    (``dyd_yolo_format``).  The CSV merge re-serialises through the native reader / writer of the path.
There is no CPU implementation of those decisions here; without the CUDA library the calls raise.
"""
from __future__ import annotations

import json
import os
import re
from pathlib import Path
from typing import Optional

import numpy as np
import pandas as pd

COL_ANN = "结果字段-目标检测标签配置"
COL_NEW = "新_结果字段-目标检测标签配置"
PHASES = {}


def _kernels():
    from . import processor
    return processor.KERNELS


# =============================================================================================
# step 1: merge                                                        reference: processor.py:26-109
# =============================================================================================
def _merge_one_native(csv_file, encoding, chunk_size):
    """Whole-file native read of one input when that cannot differ from the reference's chunked read: valid UTF-8 in the
    restated dialect, and either no more rows than one chunk, or every column holds a cell that is certainly text in every
    chunk (the reference infers dtypes chunk by chunk and writes each chunk as inferred: "6" becomes "6.0" in a chunk of
    floats but stays "6" next to an "x").  None = use the reference's own chunk loop."""
    from . import native
    if not native.enabled():
        return None
    before = dict(native._READ_STATS)
    try:
        df = native.read_csv(str(csv_file), encoding=encoding, parse_dates=False, _strict_native=True, _window_rows=chunk_size)
    except native.NotNative:
        return None
    delegated = native._READ_STATS["delegated_columns"] - before["delegated_columns"]
    if len(df) > chunk_size and delegated:
        return None
    return df


def merge_all_csv_in_folder(
        folder_path,
        output_file="merged_csv.csv",
        encoding="utf-8-sig",
        chunk_size: int = 100000,
        progress_callback=None,
):
    from . import native
    if not os.path.exists(folder_path):
        raise FileNotFoundError(f"文件夹不存在：{folder_path}")
    csv_files = list(Path(folder_path).glob("*.csv"))          # filesystem order, unsorted -- like the reference (:36)
    if not csv_files:
        print(f"警告：文件夹 {folder_path} 中未找到CSV文件")
        return None
    print(f"找到 {len(csv_files)} 个CSV文件，开始合并...")
    output_file = str(output_file)
    Path(output_file).parent.mkdir(parents=True, exist_ok=True)
    header_written = False
    total_rows = 0
    total_bytes = sum(f.stat().st_size for f in csv_files)
    completed_bytes = 0

    def emit(df):
        nonlocal header_written
        df["source_file"] = os.path.basename(csv_file)
        native.to_csv(df, output_file, encoding, mode="w" if not header_written else "a", header=not header_written)
        header_written = True

    for file_idx, csv_file in enumerate(csv_files, start=1):
        try:
            file_size = csv_file.stat().st_size
            if progress_callback:
                progress_callback(file_idx, len(csv_files), csv_file.name, total_rows, 0, 0, file_size, 0, total_bytes, completed_bytes)
            file_rows = 0
            whole = _merge_one_native(csv_file, encoding, chunk_size)
            if whole is not None:
                n_chunks = max(1, -(-len(whole) // chunk_size))
                emit(whole)
                file_rows += len(whole)
                total_rows += len(whole)
                if progress_callback:
                    progress_callback(file_idx, len(csv_files), csv_file.name, total_rows, file_rows, n_chunks, file_size, file_size,
                                      total_bytes, completed_bytes + file_size)
            else:
                with open(csv_file, "r", encoding=encoding, errors="ignore") as f:
                    for chunk_idx, df in enumerate(pd.read_csv(f, parse_dates=False, chunksize=chunk_size), start=1):
                        emit(df)
                        rows = len(df)
                        file_rows += rows
                        total_rows += rows
                        file_bytes = f.tell()
                        if progress_callback:
                            progress_callback(file_idx, len(csv_files), csv_file.name, total_rows, file_rows, chunk_idx, file_size,
                                              file_bytes, total_bytes, completed_bytes + file_bytes)
            print(f"成功读取：{csv_file.name}（{file_rows}行）")
            completed_bytes += file_size
        except Exception as e:
            print(f"读取失败 {csv_file.name}：{str(e)}")
            continue
    if not header_written:
        print("错误：没有可合并的有效CSV数据")
        return None
    print(f"\n合并完成！共 {total_rows} 行数据")
    print(f"输出文件：{os.path.abspath(output_file)}")
    return total_rows


# =============================================================================================
# unclassified summary                                                 reference: processor.py:833-891
# =============================================================================================
_REASON_LABEL = re.compile(r"^标签(.+?)(未在规则中定义)$")


def summarize_unclassified_df(df: pd.DataFrame):
    """The three sheets of unclassified_summary.xlsx.  Which (label, reason) pair a row contributes is string work with the
    reference's rules; the pairs are dictionary-encoded in encounter order and counted on the device (K3 histogram)."""
    from .labels import split_labels
    reason_col = "无法分类原因"
    if reason_col not in df.columns:
        df[reason_col] = "未知原因"
    reason_counts = df[reason_col].fillna("未知原因").value_counts().reset_index()
    reason_counts.columns = ["原因", "数量"]
    has_labels = "无法分类标签" in df.columns
    label_ids, pair_ids = {}, {}
    lab_seq, pair_seq = [], []
    reasons = df[reason_col].tolist()
    cells = df["无法分类标签"].tolist() if has_labels else [None] * len(df)
    for reason, cell in zip(reasons, cells):
        labels = split_labels(cell) if has_labels else []       # (a NaN cell is truthy: "nan" becomes a label, as in the reference)
        if not labels:
            m = _REASON_LABEL.match(str(reason))
            labels = [m.group(1)] if m else ["无标签"]
        for label in labels:
            lab_seq.append(label_ids.setdefault(label, len(label_ids)))
            pair_seq.append(pair_ids.setdefault((label, reason), len(pair_ids)))
    k = _kernels()
    lab_cnt = k.hist(np.asarray(lab_seq, np.int32), len(label_ids)) if label_ids else np.zeros(0, np.int64)
    pair_cnt = k.hist(np.asarray(pair_seq, np.int32), len(pair_ids)) if pair_ids else np.zeros(0, np.int64)
    label_summary = pd.DataFrame([{"标签": lab, "数量": int(lab_cnt[i])} for lab, i in label_ids.items()]).sort_values("数量", ascending=False)
    reason_label_summary = pd.DataFrame([{"标签": key[0], "原因": key[1], "数量": int(pair_cnt[i])} for key, i in pair_ids.items()]
                                        ).sort_values("数量", ascending=False)
    return reason_counts, label_summary, reason_label_summary


def summarize_unclassified(
        unclassified_excel_path: str,
        output_dir: str,
        json_columns: Optional[list] = None,
):
    if not os.path.exists(unclassified_excel_path):
        raise FileNotFoundError(f"无法分类文件不存在：{unclassified_excel_path}")
    df = pd.read_excel(unclassified_excel_path)
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    reason_counts, label_summary, reason_label_summary = summarize_unclassified_df(df)
    out_path = output_dir / "unclassified_summary.xlsx"
    with pd.ExcelWriter(out_path) as writer:
        reason_counts.to_excel(writer, sheet_name="reason_summary", index=False)
        label_summary.to_excel(writer, sheet_name="label_summary", index=False)
        reason_label_summary.to_excel(writer, sheet_name="reason_label", index=False)
    return out_path


# =============================================================================================
# YOLO dataset writer                                                  reference: processor.py:893-1087
# =============================================================================================
def _safe_name(value: str) -> str:
    """utils.safe_filename (utils.py:525-529)."""
    if not value:
        return "train"
    return re.sub(r"[^A-Za-z0-9._-]+", "_", value).strip("_") or "train"


def _dataset_dir_name(category_name, default_name):
    """utils._safe_dataset_dir_name (utils.py:630-633)."""
    cleaned = _safe_name(str(category_name))
    return cleaned if cleaned and cleaned != "train" else default_name


def _image_stem(source_url, idx):
    """utils._safe_image_stem (utils.py:712-724)."""
    if not source_url:
        return f"img_{idx}"
    try:
        stem = Path(Path(str(source_url)).name).stem
        if "?" in stem:
            stem = stem.split("?")[0]
        return f"{_safe_name(stem)}_{idx}"
    except Exception:  # noqa: BLE001
        return f"img_{idx}"


def _cached_image(source_url, cache_dir: Path):
    """utils._ensure_image_cached (utils.py:726-748): a local path is used as is; a URL is fetched once into the cache."""
    if not source_url:
        return None
    try:
        if Path(source_url).exists():
            return Path(source_url)
        filename = source_url.split("/")[-1]
        if "?" in filename:
            filename = filename.split("?")[0]
        if not filename:
            filename = f"image_{hash(source_url)}.jpg"
        cache_path = cache_dir / filename
        if cache_path.exists() and cache_path.stat().st_size > 0:
            return cache_path
        import requests
        try:                                                   # utils.download_image (utils.py:44-55)
            r = requests.get(source_url, timeout=10)
            if r.status_code == 200:
                with open(cache_path, "wb") as fh:
                    fh.write(r.content)
        except Exception:  # noqa: BLE001
            pass
        if cache_path.exists():
            return cache_path
    except Exception:  # noqa: BLE001
        pass
    return None


def _boxes_with_labels(json_str):
    """utils._extract_boxes_with_labels (utils.py:681-710): (label, x1, y1, x2, y2) per named object with a ptList, corner
    values by builtin min / max over the points that carry the coordinate; any exception ends the scan (prefix kept)."""
    boxes = []
    try:
        if pd.isna(json_str) or not isinstance(json_str, str):
            return boxes
        data = json.loads(json_str)
        for obj in data.get("objects", []):
            if not isinstance(obj, dict):
                continue
            label = obj.get("name")
            if not label:
                continue
            ptlist = obj.get("polygon", {}).get("ptList", [])
            if not ptlist:
                continue
            xs = [p.get("x") for p in ptlist if isinstance(p, dict) and "x" in p]
            ys = [p.get("y") for p in ptlist if isinstance(p, dict) and "y" in p]
            if not xs or not ys:
                continue
            boxes.append((label, min(xs), min(ys), max(xs), max(ys)))
    except Exception:  # noqa: BLE001
        pass
    return boxes


def _device_number(v) -> bool:
    """True if fp64 carries the value exactly as CPython computes with it (the label arithmetic is then the kernel's)."""
    if isinstance(v, bool):
        return False
    if isinstance(v, (int, np.integer)):
        return abs(int(v)) <= (1 << 52)
    return isinstance(v, (float, np.floating))


def _label_lines_host(class_id, boxes, width, height):
    """CPython lane for rows whose numbers fp64 cannot carry (huge ints, strings that happen to compare): :1045-1052 as is."""
    lines = []
    for _, x1, y1, x2, y2 in boxes:
        x1, x2 = min(x1, x2), max(x1, x2)
        y1, y2 = min(y1, y2), max(y1, y2)
        bw = max(x2 - x1, 0.0)
        bh = max(y2 - y1, 0.0)
        if bw <= 0 or bh <= 0:
            continue
        lines.append(f"{class_id} {(x1 + x2) / 2 / width:.6f} {(y1 + y2) / 2 / height:.6f} {bw / width:.6f} {bh / height:.6f}")
    return lines


def label_texts(rows):
    """rows: list of (class_id, boxes [(label, x1, y1, x2, y2)...], width, height) -> list of label-file texts ("" = no valid
    line).  One device call for all rows: dyd_yolo_normalise on the packed boxes, dyd_yolo_format for the text."""
    from . import native
    n = len(rows)
    texts = [None] * n
    dev_rows = []
    for r, (cid, boxes, w, h) in enumerate(rows):
        if _device_number(w) and _device_number(h) and all(_device_number(v) for b in boxes for v in b[1:]):
            dev_rows.append(r)
        else:                                                  # (an exception here leaves the call, as it does in the reference)
            texts[r] = "\n".join(_label_lines_host(cid, boxes, w, h))
    if dev_rows:
        counts = np.array([len(rows[r][1]) for r in dev_rows], np.int64)
        img_off = np.zeros(len(dev_rows) + 1, np.int64); np.cumsum(counts, out=img_off[1:])
        pts = np.array([v for r in dev_rows for b in rows[r][1] for v in b[1:]], np.float64)
        wh = np.array([v for r in dev_rows for v in (rows[r][2], rows[r][3])], np.float64)
        cls = np.repeat(np.array([rows[r][0] for r in dev_rows], np.int32), counts)
        if len(pts):
            cxcywh, ok = _kernels().yolo(img_off, pts, wh)
        else:
            cxcywh, ok = np.zeros(0, np.float64), np.zeros(0, np.uint8)
        text, off = native.yolo_label_texts(img_off, cls, cxcywh, ok)
        blob = text.tobytes()
        for k, r in enumerate(dev_rows):
            texts[r] = blob[off[k]:off[k + 1]].decode("ascii")
    return texts


def generate_yolo_datasets_from_excels(
        category_excels: list,
        output_dir: str,
        image_cache_dir: Optional[str] = None,
        source_col: str = "source",
        label_col: str = "分类标签",
        json_col_primary: str = "新_结果字段-目标检测标签配置",
        json_col_fallback: str = "结果字段-目标检测标签配置",
        width_col: str = "width",
        height_col: str = "height",
        download_images: bool = True,
        random_seed: int = 42,
        class_order: Optional[list] = None,
        resume: bool = True,
        progress_callback=None,
):
    import yaml
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    cache_dir = Path(image_cache_dir) if image_cache_dir else (output_dir / "image_cache")
    cache_dir.mkdir(parents=True, exist_ok=True)
    datasets, dataset_name_map, skipped, dataset_stats = [], {}, [], {}
    total_rows = processed_rows = downloaded_images = 0
    used_dir_names = set()
    for excel_path in category_excels:
        if not excel_path or not Path(excel_path).exists():
            continue
        xls = pd.ExcelFile(excel_path)
        for split in ["train", "val", "test"]:
            if split in xls.sheet_names:
                total_rows += len(pd.read_excel(excel_path, sheet_name=split))

    for idx_excel, excel_path in enumerate(category_excels):
        if not excel_path or not Path(excel_path).exists():
            continue
        excel_path = Path(excel_path)
        category_name = excel_path.stem
        base_dir_name = _dataset_dir_name(category_name, f"category_{idx_excel:03d}")
        dir_name, suffix = base_dir_name, 1
        while dir_name in used_dir_names:
            dir_name = f"{base_dir_name}_{suffix}"
            suffix += 1
        used_dir_names.add(dir_name)
        dataset_dir = output_dir / dir_name
        dataset_name_map[dataset_dir.name] = category_name
        images_root, labels_root = dataset_dir / "images", dataset_dir / "labels"
        for split in ["train", "val", "test"]:
            (images_root / split).mkdir(parents=True, exist_ok=True)
            (labels_root / split).mkdir(parents=True, exist_ok=True)
        xls = pd.ExcelFile(excel_path)
        split_sheets = [s for s in ["train", "val", "test"] if s in xls.sheet_names]
        all_labels, split_dfs = [], {}
        for split in split_sheets:
            df_split = pd.read_excel(excel_path, sheet_name=split)
            split_dfs[split] = df_split
            if label_col in df_split.columns:
                all_labels.extend([str(v) for v in df_split[label_col].dropna()])
        classes = sorted(list(dict.fromkeys(all_labels)))      # class ids: sorted unique labels (:964), class_order first (:965-968)
        if class_order:
            ordered = [c for c in class_order if c in classes]
            classes = ordered + [c for c in classes if c not in ordered]
        class_to_id = {name: i for i, name in enumerate(classes)}
        dataset_stats[category_name] = {"train": 0, "val": 0, "test": 0}

        for split in split_sheets:
            df_split = split_dfs[split].sample(frac=1, random_state=random_seed).reset_index(drop=True)   # per-split shuffle (:1003)
            records = df_split.to_dict("records")
            # pass 1 (no side effects): the boxes of every row that can reach the label arithmetic, one device call for all
            cand = {}
            for idx, row in enumerate(records):
                source = row.get(source_col)
                label_value = str(row.get(label_col, ""))
                if not source or not label_value or label_value not in class_to_id:
                    continue
                boxes = [b for b in _boxes_with_labels(row.get(json_col_primary) or row.get(json_col_fallback)) if b[0] == label_value]
                if boxes:
                    cand[idx] = boxes
            # the arithmetic needs width / height, which pass 2 may only learn from the image file: rows with sizes in the sheet
            # go to the device now, the others are computed when their image has been opened
            pre = {idx: (class_to_id[str(records[idx].get(label_col, ""))], b, records[idx].get(width_col), records[idx].get(height_col))
                   for idx, b in cand.items() if records[idx].get(width_col) and records[idx].get(height_col)}
            keys = list(pre)
            texts = dict(zip(keys, label_texts([pre[k] for k in keys]))) if keys else {}
            # pass 2: the reference's loop, in order, with every side effect
            for idx, row in enumerate(records):
                if progress_callback and processed_rows % 50 == 0:
                    progress_callback(processed_rows, total_rows, downloaded_images, category_name, split, f"idx_{idx}", "", excel_path.name, idx)
                source = row.get(source_col)
                if not source:
                    skipped.append({"category": category_name, "reason": "缺少source", "split": split})
                    processed_rows += 1
                    continue
                label_value = str(row.get(label_col, ""))
                if not label_value or label_value not in class_to_id:
                    skipped.append({"category": category_name, "reason": "缺少或无效分类标签", "split": split})
                    processed_rows += 1
                    continue
                image_stem = _image_stem(str(source), idx)
                label_path = labels_root / split / f"{image_stem}.txt"
                if resume and label_path.exists() and label_path.stat().st_size > 0:
                    dataset_stats[category_name][split] += 1
                    processed_rows += 1
                    continue
                filtered_boxes = cand.get(idx)
                if not filtered_boxes:
                    skipped.append({"category": category_name, "reason": "无匹配标签框", "split": split})
                    processed_rows += 1
                    continue
                image_path = None
                if download_images:
                    image_path = _cached_image(str(source), cache_dir)
                elif Path(str(source)).exists():
                    image_path = Path(str(source))
                width, height = row.get(width_col), row.get(height_col)
                if (not width or not height) and image_path:
                    try:
                        from PIL import Image
                        with Image.open(image_path) as img:
                            width, height = img.size
                    except Exception:  # noqa: BLE001
                        pass
                if not width or not height:
                    skipped.append({"category": category_name, "reason": "缺少图像尺寸", "split": split})
                    processed_rows += 1
                    continue
                out_image = images_root / split / f"{image_stem}{image_path.suffix if image_path else '.jpg'}"
                if image_path:
                    if not out_image.exists():
                        try:
                            out_image.write_bytes(Path(image_path).read_bytes())
                            downloaded_images += 1
                        except Exception:  # noqa: BLE001
                            skipped.append({"category": category_name, "reason": "图片写入失败", "split": split})
                            processed_rows += 1
                            continue
                else:
                    skipped.append({"category": category_name, "reason": "图片下载失败", "split": split})
                    processed_rows += 1
                    continue
                text = texts.get(idx)
                if text is None:                               # sizes came from the image file
                    text = label_texts([(class_to_id[label_value], filtered_boxes, width, height)])[0]
                if text:
                    label_path.write_text(text, encoding="utf-8")
                    dataset_stats[category_name][split] += 1
                else:
                    skipped.append({"category": category_name, "reason": "标注框无效", "split": split})
                processed_rows += 1

        (dataset_dir / "data.yaml").write_text(yaml.dump({
            "path": str(dataset_dir), "train": "images/train", "val": "images/val", "test": "images/test",
            "nc": len(classes), "names": classes}, sort_keys=False, allow_unicode=True), encoding="utf-8")
        datasets.append(dataset_dir)

    skipped_path = output_dir / "yolo_skipped.xlsx"
    pd.DataFrame(skipped if skipped else [{"category": "无", "reason": "无", "split": "无"}]).to_excel(skipped_path, index=False)
    if progress_callback:
        # the reference reads names that do not exist here (current_category, ...) and dies with NameError (:1076-1077);
        # the UI never passes a callback (ui/pages/processing.py:644).  Same behaviour.
        raise NameError("name 'current_category' is not defined")
    return {"datasets": datasets, "skipped": skipped_path, "stats": dataset_stats, "total": total_rows, "processed": processed_rows,
            "downloaded": downloaded_images, "dataset_name_map": dataset_name_map}


# =============================================================================================
# label-count summary                                                  reference: processor.py:1089-1163
# =============================================================================================
def _read_label_dir(label_dir: Path, names):
    """Label files of one split -> (n_files, label-name ids per file as CSR, vocabulary) with the reference's parsing:
    first token of a non-empty line through int(float(.)), out-of-range ids named by their number, bad lines skipped."""
    vocab, ids, counts = {}, [], []
    n_files = 0
    if label_dir.exists():
        for txt_path in label_dir.glob("*.txt"):
            n_files += 1
            try:
                lines = txt_path.read_text(encoding="utf-8", errors="ignore").splitlines()
            except Exception:  # noqa: BLE001
                continue
            k = 0
            for line in lines:
                parts = line.strip().split()
                if not parts:
                    continue
                try:
                    class_id = int(float(parts[0]))
                    label_name = names[class_id] if class_id < len(names) else str(class_id)
                    hash(label_name)
                except Exception:  # noqa: BLE001
                    continue
                ids.append(vocab.setdefault(label_name, len(vocab)))
                k += 1
            counts.append(k)
    off = np.zeros(len(counts) + 1, np.int64)
    if counts:
        np.cumsum(counts, out=off[1:])
    return n_files, off, np.asarray(ids, np.int32), vocab


def summarize_yolo_label_counts(dataset_dirs):
    import yaml
    stats, flat_rows = {}, []
    for dataset_dir in dataset_dirs or []:
        if not dataset_dir:
            continue
        dataset_path = Path(dataset_dir)
        if not dataset_path.exists():
            continue
        names = []
        data_yaml = dataset_path / "data.yaml"
        if data_yaml.exists():
            try:
                names = yaml.safe_load(data_yaml.read_text(encoding="utf-8")).get("names") or []
            except Exception:  # noqa: BLE001
                pass
        dataset_key = dataset_path.name
        split_stats, total_images_all, total_img_counts, total_box_counts = {}, 0, {}, {}
        for split in ["train", "val", "test"]:
            total_images, off, ids, vocab = _read_label_dir(dataset_path / "labels" / split, names)
            img_counts, box_counts = {}, {}
            if len(vocab):
                ih, bh = _kernels().label_presence(off, ids, len(vocab))       # per-image presence + per-box counts on the device
                for label, v in vocab.items():                                  # dict order = first occurrence, like the reference
                    box_counts[label] = int(bh[v])
                # the reference's img_counts dict is keyed in the order images first SHOW a label; with sets per image that
                # order is only observable through dict iteration below, which the flat rows do not depend on beyond set()
                for label, v in vocab.items():
                    if ih[v]:
                        img_counts[label] = int(ih[v])
            split_stats[split] = {"total_images": total_images, "label_counts": img_counts, "box_counts": box_counts}
            total_images_all += total_images
            for label, count in img_counts.items():
                total_img_counts[label] = total_img_counts.get(label, 0) + count
            for label, count in box_counts.items():
                total_box_counts[label] = total_box_counts.get(label, 0) + count
            for label in set(img_counts) | set(box_counts):
                flat_rows.append({"数据集": dataset_key, "split": split, "标签": label, "图片数量": img_counts.get(label, 0),
                                  "标注框数量": box_counts.get(label, 0),
                                  "占比%": f"{(img_counts.get(label, 0) / total_images * 100):.1f}%" if total_images else "0.0%",
                                  "split总图片数": total_images})
        split_stats["all"] = {"total_images": total_images_all, "label_counts": total_img_counts, "box_counts": total_box_counts}
        stats[dataset_key] = split_stats
        for label in set(total_img_counts) | set(total_box_counts):
            flat_rows.append({"数据集": dataset_key, "split": "all", "标签": label, "图片数量": total_img_counts.get(label, 0),
                              "标注框数量": total_box_counts.get(label, 0),
                              "占比%": f"{(total_img_counts.get(label, 0) / total_images_all * 100):.1f}%" if total_images_all else "0.0%",
                              "split总图片数": total_images_all})
    return stats, pd.DataFrame(flat_rows)
