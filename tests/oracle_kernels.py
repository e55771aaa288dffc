"""Stand-in for deal_yolo_daya_b200.processor.KERNELS backed by the CPU oracle.

TEST INFRASTRUCTURE: lets the -m "not gpu" suite exercise the drop-in's HOST logic (ingest,
egress, file contract) without a GPU.  The product never imports this; on the GPU box
tests/test_gpu_dropin.py runs the same checks through the real CUDA facade.
"""
from __future__ import annotations

import numpy as np

from oracle import oracle_c, oracle_np


class OracleKernels:
    def bbox(self, poly_off, xy):
        return oracle_c.bbox_fold(poly_off, xy)

    def bbox_fused(self, img_off, poly_off, xy, min_boxes, thr):
        pts, valid, arg = oracle_c.bbox_fold(poly_off, xy)
        high, count = oracle_c.iou_filter(img_off, pts, valid, min_boxes, thr)
        return pts, valid, arg, high, count

    def iou(self, img_off, pts, valid, min_boxes, thr):
        return oracle_c.iou_filter(img_off, pts, valid, min_boxes, thr)

    def dedup(self, off, data, null, keep):
        return oracle_c.dedup(oracle_c.hash_strings_buf(off, data), null, keep)

    def antijoin(self, moff, mdata, mnull, roff, rdata, rnull):
        return oracle_c.antijoin(oracle_c.hash_strings_buf(moff, mdata), mnull, oracle_c.hash_strings_buf(roff, rdata), rnull)

    def label_lut(self, img_off, label_id, lut_new, lut_ntok, lut_nrep):
        new, rr, cnt = oracle_c.label_lut(img_off, label_id, lut_new, lut_ntok, lut_nrep)
        lid = np.asarray(label_id)
        hist = np.bincount(lid[lid >= 0], minlength=len(lut_new)).astype(np.uint64)
        return new, rr, cnt, hist

    def split_expand(self, img_off, label_id, cat_of_label, n_cat):
        return oracle_c.split_expand(img_off, label_id, cat_of_label, n_cat)

    def split_assign(self, cat_off, perm, n_train, n_val):
        n_exp = int(cat_off[-1])
        split = np.zeros(n_exp, np.uint8); pos = np.zeros(n_exp, np.int64)
        for c in range(len(cat_off) - 1):
            a, b = int(cat_off[c]), int(cat_off[c + 1])
            p = perm[a:b]
            pos[a + p] = np.arange(b - a)
            r = np.arange(b - a)
            split[a + p] = np.where(r < n_train[c], 0, np.where(r < n_train[c] + n_val[c], 1, 2))
        return split, pos

    def yolo(self, img_off, pts, img_wh):
        img_off = np.asarray(img_off); pts = np.asarray(pts, np.float64).reshape(-1, 4); wh = np.asarray(img_wh, np.float64).reshape(-1, 2)
        out = np.zeros((len(pts), 4), np.float64); ok = np.zeros(len(pts), np.uint8)
        for i in range(len(img_off) - 1):
            for q in range(int(img_off[i]), int(img_off[i + 1])):
                r = oracle_np.yolo_norm(tuple(pts[q]), wh[i, 0], wh[i, 1])
                if r is not None:
                    out[q] = r; ok[q] = 1
                else:                              # the kernel still writes the arithmetic; only `ok` is observable
                    x1, y1, x2, y2 = pts[q]
                    with np.errstate(all="ignore"):
                        out[q] = ((x1 + x2) / 2 / wh[i, 0], (y1 + y2) / 2 / wh[i, 1], max(x2 - x1, 0.0) / wh[i, 0], max(y2 - y1, 0.0) / wh[i, 1])
        return out.reshape(-1), ok

    def label_presence(self, img_off, label_id, n_vocab):
        ih = np.zeros(n_vocab, np.int64); bh = np.zeros(n_vocab, np.int64)
        lid = np.asarray(label_id)
        for i in range(len(img_off) - 1):
            seg = lid[int(img_off[i]):int(img_off[i + 1])]
            seg = seg[(seg >= 0) & (seg < n_vocab)]
            for v in seg:
                bh[v] += 1
            for v in set(seg.tolist()):
                ih[v] += 1
        return ih, bh

    def hist(self, ids, n_vocab):
        ids = np.asarray(ids)
        return np.bincount(ids[(ids >= 0) & (ids < n_vocab)], minlength=n_vocab).astype(np.int64)
