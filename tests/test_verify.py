"""The bench's ground-truth checker (deal_yolo_daya_b200/verify.py, torch sort-based on integer url ids) against
the CPU oracle on the hashed URL strings of the same rows -- so the checker itself is pinned."""
import numpy as np
import torch

from deal_yolo_daya_b200 import synth, verify
from oracle import oracle_c


def test_expected_dedup_and_antijoin_equal_the_oracle():
    n, n_ref = 20000, 9000
    ids = synth.url_ids_of(5, np.arange(n))
    ref = synth.ref_ids_of(5, np.arange(n_ref), n)
    keys = oracle_c.hash_strings([synth.url_of(i) for i in ids])
    rkeys = oracle_c.hash_strings([synth.url_of(i) for i in ref])
    wk, wr = oracle_c.dedup(keys, np.zeros(n, np.uint8), "first")
    k, r = verify.expected_dedup_first(torch.from_numpy(ids), 0)
    assert np.array_equal(k.numpy(), wk) and np.array_equal(r.numpy(), wr)
    assert verify.global_counts(k) == int(n - wk.sum()) > 0
    ak, ar = oracle_c.antijoin(keys, np.zeros(n, np.uint8), rkeys, np.zeros(n_ref, np.uint8))
    k, r = verify.expected_antijoin(torch.from_numpy(ids), torch.from_numpy(ref), 0)
    assert np.array_equal(k.numpy(), ak) and np.array_equal(r.numpy(), ar)
    assert 0 < int(ak.sum()) < n
