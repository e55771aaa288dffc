"""Seeded CSR test tables (shared by the CPU-oracle and GPU parity tests)."""
from __future__ import annotations

import numpy as np

from deal_yolo_daya_b200 import synth

NAN = float("nan")
INF = float("inf")


def csr_from_polygons(images):
    """images: list of images, each a list of polygons, each a list of (x, y)."""
    img_off = [0]; poly_off = [0]; xy = []
    for polys in images:
        for pts in polys:
            for x, y in pts:
                xy += [x, y]
            poly_off.append(poly_off[-1] + len(pts))
        img_off.append(img_off[-1] + len(polys))
    return (np.array(img_off, np.int64), np.array(poly_off, np.int64), np.array(xy, np.float64))


def edge_polygon_table():
    """Hand-made polygons covering ties, signed zeros, NaN/inf placement, V=0/1, V>32."""
    sq = lambda x0, y0, x1, y1: [(x0, y0), (x1, y0), (x1, y1), (x0, y1)]  # noqa: E731
    rng = np.random.RandomState(3)
    long_poly = [(float(a), float(b)) for a, b in rng.uniform(-50, 50, size=(77, 2))]
    long_poly[40] = (-60.0, 70.0)
    long_zero = [(0.0, 5.0)] * 20 + [(-0.0, 5.0)] * 20            # V=40: first zero is +0.0
    long_nzero = [(-0.0, -0.0)] + [(0.0, 0.0)] * 39
    images = [
        [[(10.0, 5.0), (10.0, 5.0), (10.0, 7.0)]],                                   # ties
        [[(0.0, -0.0), (-0.0, 0.0), (0.0, -0.0)], [(-0.0, 0.0), (0.0, -0.0)]],       # signed zeros
        [[(-0.0, 1.0), (0.0, 2.0), (3.0, 0.0), (2.0, -0.0)]],                        # zero is min only
        [[(-3.0, -1.0), (-0.0, -0.0), (0.0, 0.0)]],                                  # zero is max only
        [[], [(1.0, 2.0)], []],                                                      # empty / single vertex
        [sq(10, 10, 110, 110), sq(10, 10, 110, 110), []],                            # [A, A, null] -> high
        [[], sq(10, 10, 110, 110), sq(10, 10, 110, 110)],                            # [null, A, A] -> other
        [sq(5, 5, 5, 20), sq(5, 5, 5, 20)],                                          # zero-area
        [sq(0, 0, 10, 10), sq(0, 0, 10, 7)],                                         # 70/100
        [sq(0, 0, 10, 10), sq(0, 0, 7, 10)],
        [sq(1, 2, 3, 4)],                                                            # single box
        [],                                                                          # no objects
        [[(NAN, 1.0), (2.0, NAN), (3.0, 0.5)]],                                      # NaN first / middle
        [[(4.0, 1.0), (NAN, NAN), (3.0, 2.5), (INF, -INF)], [(4.0, 1.0), (3.0, 2.5), (INF, -INF)]],
        [[(NAN, NAN)] * 5],                                                          # all NaN
        [[(1.0, 1.0), (NAN, NAN), (NAN, NAN)]],
        [long_poly, long_zero, long_nzero],                                          # V > 32 paths
        [[(float(i % 7), float((i * 3) % 5)) for i in range(32)]],                   # V == 32 exactly
        [[(float(i % 7), float((i * 3) % 5)) for i in range(33)]],                   # V == 33
        [sq(-5, -7, 3999, 2999), sq(-5, -7, 3999, 2990)],
        [sq(1e-200, 1e-200, 3e-200, 3e-200), sq(1e-200, 1e-200, 3e-200, 2.9e-200)],  # tiny areas: exact-division path
        [sq(0, 0, 1e160, 1e160), sq(0, 0, 1e160, 0.9e160)],                          # overflowing areas
        [sq(0, 0, INF, 10), sq(0, 0, INF, 10)],
    ]
    return csr_from_polygons(images)


def random_polygon_table(seed, n_img, max_polys=12, max_verts=40, p_empty=0.05, p_special=0.03):
    """Random ragged table with a sprinkling of special values at random positions."""
    rng = np.random.RandomState(seed)
    images = []
    specials = [0.0, -0.0, NAN, INF, -INF, 1920.0]
    for _ in range(n_img):
        polys = []
        for _ in range(rng.randint(0, max_polys + 1)):
            if rng.rand() < p_empty:
                polys.append([]); continue
            v = rng.randint(1, max_verts + 1)
            pts = rng.uniform(0, 1920, size=(v, 2)).round(rng.randint(0, 4))
            pts = [(float(a), float(b)) for a, b in pts]
            if rng.rand() < 0.3:                      # near-duplicate of the previous polygon
                if polys and polys[-1]:
                    pts = [(a + rng.uniform(-0.5, 0.5), b) for a, b in polys[-1]]
            for k in range(len(pts)):
                if rng.rand() < p_special:
                    pts[k] = (specials[rng.randint(len(specials))], pts[k][1])
                if rng.rand() < p_special:
                    pts[k] = (pts[k][0], specials[rng.randint(len(specials))])
            polys.append(pts)
        images.append(polys)
    return csr_from_polygons(images)


def random_box_table(seed, n_img, lo, hi, p_invalid=0.02, span=400.0, p_dup=0.02):
    """Boxes given directly as two points (possibly unordered), with invalid markers."""
    rng = np.random.RandomState(seed)
    nb = rng.randint(lo, hi + 1, size=n_img)
    img_off = np.zeros(n_img + 1, np.int64); np.cumsum(nb, out=img_off[1:])
    n = int(img_off[-1])
    c = rng.uniform(0, span, size=(n, 2)); h = rng.uniform(0.5, 12, size=(n, 2))
    pts = np.concatenate([c - h, c + h], axis=1)
    swap = rng.rand(n) < 0.3
    pts[swap] = pts[swap][:, [2, 3, 0, 1]]
    dup = np.where(rng.rand(n) < p_dup)[0]
    dup = dup[dup > 0]
    pts[dup] = pts[dup - 1] + rng.uniform(-0.01, 0.01, size=(len(dup), 4))
    valid = (rng.rand(n) >= p_invalid).astype(np.uint8)
    return img_off, pts.reshape(-1).astype(np.float64), valid


def synth_csr(seed, first, n):
    t = synth.make_table(seed, first, n)
    return t
