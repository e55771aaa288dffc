"""Seeded synthetic annotation tables, host (numpy) twin.

Every value is a pure function of ``(seed, global image id, tag, k)`` through a
64-bit integer mixer, so any row can be regenerated anywhere: the CUDA generator
in ``csrc/synth.cu`` evaluates the same integer recipe and the same IEEE-754
double operations (no FMA, no transcendental functions), which makes the device
table and this numpy table bit-identical.  That is what lets the at-scale
device-resident benchmark inputs be parity-checked on sampled slices against
the CPU oracle.

Layout produced (the ragged CSR the kernels consume):

    img_off  int64[n_img+1]   polygon range of each image
    poly_off int64[n_poly+1]  vertex range of each polygon
    xy       float64[2*n_vert] interleaved x,y (only valid points)
    label_id int32[n_poly]
    img_wh   float64[2*n_img]
    url ids  int64[n_img]     -> "https://img.example.com/<id>.jpg"

The reference has no generator; its row format is the JSON document handled at
/root/reference/src/deal_yolo_data/core/processor.py:262-281 (``objects[*].
polygon.ptList[*].{x,y}``, top-level ``width``/``height``).
"""
from __future__ import annotations

import json
from dataclasses import dataclass

import numpy as np

U64 = np.uint64
GAMMA = U64(0x9E3779B97F4A7C15)
M1 = U64(0xBF58476D1CE4E5B9)
M2 = U64(0x94D049BB133111EB)

IMG_W = 1920.0
IMG_H = 1080.0
N_LABELS = 80
MAX_POLYS = 50
MIN_VERTS = 4
VERT_SPAN = 29  # vertices per polygon ~ U{4..32}

# tags (third argument of rnd); keep in sync with csrc/synth.cu
TAG_NPOLY = 0x0A
TAG_DUPBOX = 0x0B
TAG_NVERT = 0x0C
TAG_CX = 0x0D
TAG_CY = 0x0E
TAG_R = 0x0F
TAG_LABEL = 0x10
TAG_URLDUP = 0x11
TAG_URLPICK = 0x12
TAG_REFHIT = 0x13
TAG_REFPICK = 0x14
TAG_NBOX_CROWD = 0x15
TAG_VERT = 0x1000    # + polygon slot
TAG_JITTER = 0x2000  # + nothing (the copy polygon is unique per image)

P_DUPBOX = 0.10
P_URLDUP = 0.05
P_REFHIT = 0.10


def mix64(z):
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        z = (np.asarray(z, dtype=U64) + GAMMA).astype(U64)
        z = ((z ^ (z >> U64(30))) * M1).astype(U64)
        z = ((z ^ (z >> U64(27))) * M2).astype(U64)
        return (z ^ (z >> U64(31))).astype(U64)


def rnd(seed, a, tag, k):
    """64 random bits for (seed, a, tag, k); all arguments broadcast as uint64."""
    with np.errstate(over="ignore"):
        h = mix64(np.asarray(seed, dtype=U64) + np.asarray(a, dtype=U64))
        h = mix64(h + np.asarray(tag, dtype=U64))
        return mix64(h + np.asarray(k, dtype=U64))


def u01(h):
    """Top 53 bits -> double in [0,1); exact."""
    return (np.asarray(h, dtype=U64) >> U64(11)).astype(np.float64) * (2.0 ** -53)


def poisson8_thresholds():
    """Inverse-CDF thresholds of Poisson(8) as 53-bit integers.

    Count = number of thresholds <= (h >> 11); evaluated with integers only so
    host and device agree.  The table itself is computed here once in double
    precision and handed to the device kernel as data.
    """
    lam = 8.0
    p = np.exp(-lam)
    cdf = []
    acc = 0.0
    for k in range(64):
        acc += p
        cdf.append(acc)
        p = p * lam / (k + 1)
    thr = np.minimum(np.floor(np.array(cdf) * 2.0 ** 53), 2.0 ** 53).astype(np.uint64)
    return thr


_P8 = poisson8_thresholds()


def n_polys_of(seed, img_ids):
    h = rnd(seed, img_ids, TAG_NPOLY, 0) >> U64(11)
    n = np.searchsorted(_P8, h, side="right").astype(np.int64)
    return np.clip(n, 1, MAX_POLYS)


def dupbox_of(seed, img_ids, n_polys):
    return (u01(rnd(seed, img_ids, TAG_DUPBOX, 0)) < P_DUPBOX) & (n_polys >= 2)


@dataclass
class SynthTable:
    seed: int
    first_img: int
    img_off: np.ndarray
    poly_off: np.ndarray
    xy: np.ndarray
    label_id: np.ndarray
    img_wh: np.ndarray
    url_id: np.ndarray

    @property
    def n_img(self):
        return len(self.img_off) - 1

    @property
    def n_poly(self):
        return len(self.poly_off) - 1


def make_table(seed: int, first_img: int, n_img: int) -> SynthTable:
    """Sparse-image table (configs C1/C2/C3/C5): Poisson(8) polygons, 4..32 vertices."""
    ids = np.arange(first_img, first_img + n_img, dtype=np.uint64)
    npoly = n_polys_of(seed, ids)
    dup = dupbox_of(seed, ids, npoly)
    img_off = np.zeros(n_img + 1, dtype=np.int64)
    np.cumsum(npoly, out=img_off[1:])
    n_poly = int(img_off[-1])

    poly_img = np.repeat(ids, npoly)                       # global image id of each polygon
    poly_slot = (np.arange(n_poly, dtype=np.int64) - np.repeat(img_off[:-1], npoly)).astype(np.uint64)
    is_copy = np.repeat(dup, npoly) & (poly_slot == np.repeat(npoly - 1, npoly).astype(np.uint64))
    src_slot = np.where(is_copy, U64(0), poly_slot)        # geometry comes from slot 0 for the copy

    nvert = (MIN_VERTS + (rnd(seed, poly_img, TAG_NVERT, src_slot) % U64(VERT_SPAN))).astype(np.int64)
    poly_off = np.zeros(n_poly + 1, dtype=np.int64)
    np.cumsum(nvert, out=poly_off[1:])
    n_vert = int(poly_off[-1])

    cx = u01(rnd(seed, poly_img, TAG_CX, src_slot)) * IMG_W
    cy = u01(rnd(seed, poly_img, TAG_CY, src_slot)) * IMG_H
    r = 5.0 + u01(rnd(seed, poly_img, TAG_R, src_slot)) * 195.0

    v_poly = np.repeat(np.arange(n_poly, dtype=np.int64), nvert)
    v_k = (np.arange(n_vert, dtype=np.int64) - np.repeat(poly_off[:-1], nvert)).astype(np.uint64)
    v_img = poly_img[v_poly]
    v_src = src_slot[v_poly]
    v_r = r[v_poly]
    tag = U64(TAG_VERT) + v_src
    ux = u01(rnd(seed, v_img, tag, v_k * U64(2)))
    uy = u01(rnd(seed, v_img, tag, v_k * U64(2) + U64(1)))
    x = cx[v_poly] + (2.0 * ux - 1.0) * v_r
    y = cy[v_poly] + (2.0 * uy - 1.0) * v_r
    x = np.minimum(np.maximum(x, 0.0), IMG_W)
    y = np.minimum(np.maximum(y, 0.0), IMG_H)
    cp = is_copy[v_poly]
    if cp.any():
        jx = u01(rnd(seed, v_img, TAG_JITTER, v_k * U64(2)))
        jy = u01(rnd(seed, v_img, TAG_JITTER, v_k * U64(2) + U64(1)))
        amp = 0.005 * v_r
        xj = np.minimum(np.maximum(x + (2.0 * jx - 1.0) * amp, 0.0), IMG_W)
        yj = np.minimum(np.maximum(y + (2.0 * jy - 1.0) * amp, 0.0), IMG_H)
        x = np.where(cp, xj, x)
        y = np.where(cp, yj, y)
    xy = np.empty(2 * n_vert, dtype=np.float64)
    xy[0::2] = x
    xy[1::2] = y

    label_id = (rnd(seed, poly_img, TAG_LABEL, poly_slot) % U64(N_LABELS)).astype(np.int32)
    img_wh = np.tile(np.array([IMG_W, IMG_H]), n_img)
    url_id = url_ids_of(seed, ids)
    return SynthTable(seed, first_img, img_off, poly_off, xy, label_id, img_wh, url_id)


def url_ids_of(seed, img_ids):
    """5 % of rows re-use the id of an earlier row (duplicate ``source``)."""
    ids = np.asarray(img_ids, dtype=np.uint64)
    isdup = (u01(rnd(seed, ids, TAG_URLDUP, 0)) < P_URLDUP) & (ids > 0)
    pick = rnd(seed, ids, TAG_URLPICK, 0) % np.maximum(ids, U64(1))
    return np.where(isdup, pick, ids).astype(np.int64)


def ref_ids_of(seed, ref_rows, n_main):
    """Reference-set ids: 10 % fall inside the main id range [0, n_main)."""
    q = np.asarray(ref_rows, dtype=np.uint64)
    hit = u01(rnd(seed, q, TAG_REFHIT, 0)) < P_REFHIT
    pick = rnd(seed, q, TAG_REFPICK, 0) % U64(max(n_main, 1))
    return np.where(hit, pick, U64(n_main) + q).astype(np.int64)


def make_crowd_boxes(seed: int, first_img: int, n_img: int, lo: int = 200, hi: int = 500):
    """Dense-crowd table (config C4): boxes given directly as 2-point ptLists.

    Returns (img_off, pts[4*n_box]) with pts = (p1x, p1y, p2x, p2y) per box.
    """
    ids = np.arange(first_img, first_img + n_img, dtype=np.uint64)
    nbox = (lo + (rnd(seed, ids, TAG_NBOX_CROWD, 0) % U64(hi - lo + 1))).astype(np.int64)
    img_off = np.zeros(n_img + 1, dtype=np.int64)
    np.cumsum(nbox, out=img_off[1:])
    n_box = int(img_off[-1])
    b_img = np.repeat(ids, nbox)
    b_slot = (np.arange(n_box, dtype=np.int64) - np.repeat(img_off[:-1], nbox)).astype(np.uint64)
    cx = u01(rnd(seed, b_img, TAG_CX, b_slot)) * IMG_W
    cy = u01(rnd(seed, b_img, TAG_CY, b_slot)) * IMG_H
    hw = 4.0 + u01(rnd(seed, b_img, TAG_R, b_slot)) * 36.0
    hh = 4.0 + u01(rnd(seed, b_img, TAG_NVERT, b_slot)) * 36.0
    pts = np.empty(4 * n_box, dtype=np.float64)
    pts[0::4] = np.maximum(cx - hw, 0.0)
    pts[1::4] = np.maximum(cy - hh, 0.0)
    pts[2::4] = np.minimum(cx + hw, IMG_W)
    pts[3::4] = np.minimum(cy + hh, IMG_H)
    return img_off, pts


def url_of(url_id: int) -> str:
    return f"https://img.example.com/{int(url_id)}.jpg"


def label_name(label_id: int) -> str:
    return f"cls{int(label_id):02d}"


def table_to_rows(t: SynthTable, decimals: int | None = None):
    """Reference-format rows: (source, annotation JSON text) per image.

    ``decimals`` rounds coordinates (keeps committed fixtures small); ``None``
    keeps the full doubles (``float.__repr__`` round-trips them exactly).
    """
    rows = []
    for i in range(t.n_img):
        objs = []
        for p in range(int(t.img_off[i]), int(t.img_off[i + 1])):
            a, b = int(t.poly_off[p]), int(t.poly_off[p + 1])
            pts = []
            for v in range(a, b):
                x, y = float(t.xy[2 * v]), float(t.xy[2 * v + 1])
                if decimals is not None:
                    x, y = round(x, decimals), round(y, decimals)
                pts.append({"x": x, "y": y})
            objs.append({"name": label_name(t.label_id[p]), "polygon": {"ptList": pts}})
        doc = {"width": int(t.img_wh[2 * i]), "height": int(t.img_wh[2 * i + 1]), "objects": objs}
        rows.append((url_of(t.url_id[i]), json.dumps(doc, ensure_ascii=False)))
    return rows
