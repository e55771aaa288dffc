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
