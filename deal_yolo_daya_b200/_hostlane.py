"""Host exception lane: values an fp64 device array cannot carry.

The CUDA kernels see coordinates as IEEE doubles.  JSON can also hold ``null``, strings,
integers beyond 2**53 and nested containers where a number is expected; CPython's ``min`` /
``max`` / arithmetic give those their own results or exceptions (a ``None`` coordinate makes
the reference's whole step raise ``TypeError``, processor.py:256).  ``ingest`` routes exactly
those polygons / rows here, where they are evaluated with CPython semantics so the drop-in
stays observably identical.  This is not a CPU fallback of the hot path: numeric tables never
reach it, it has no vectorised code, and the drop-in counts and reports what it handled.
"""
from __future__ import annotations

import json


def corner_points(points):
    """Two corner points of one polygon's valid points (processor.py:254-260)."""
    if not points:
        return [{"x": None, "y": None}, {"x": None, "y": None}]
    xs = [p["x"] for p in points]
    ys = [p["y"] for p in points]
    return [{"x": min(xs), "y": min(ys)}, {"x": max(xs), "y": max(ys)}]


def _iou(a, b):
    ix1 = max(a[0], b[0]); iy1 = max(a[1], b[1])
    ix2 = min(a[2], b[2]); iy2 = min(a[3], b[3])
    inter = max(0, ix2 - ix1) * max(0, iy2 - iy1)
    if inter == 0:
        return 0.0
    union = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    return inter / union if union != 0 else 0.0


def row_is_high_iou(text, min_boxes, thr) -> bool:
    """One row of step 5 with CPython semantics (processor.py:341-376)."""
    boxes = []
    try:
        if isinstance(text, str):
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
                boxes.append((min(p["x"], q["x"]), min(p["y"], q["y"]), max(p["x"], q["x"]), max(p["y"], q["y"])))
    except Exception:  # noqa: BLE001
        pass
    n = len(boxes)
    if n < min_boxes:
        return False
    for i in range(n):
        for j in range(i + 1, n):
            if _iou(boxes[i], boxes[j]) >= thr:      # may raise, as in the reference (uncaught there too)
                return True
    return False
