"""Full-size checks (BASELINE.json configs 2-4) through size-independent properties.

The oracle cannot run 10 M images in seconds, so at full size the kernels are held to
  * sampled-slice parity: the device table is bit-identical to the numpy generator, so any
    image range can be rebuilt on the host and pushed through the C oracle;
  * cross-implementation identity: TMA-staged fused kernel == direct fused kernel == K1 then K2;
  * algebraic properties of dedup / anti-join (representative idempotence, key equality,
    first-occurrence order, keep <-> hit equivalence) and of the IoU flag (monotone in thr).
"""
from __future__ import annotations

import os

import numpy as np
import pytest
import torch

from deal_yolo_daya_b200 import ops, synth, synth_device
from oracle import oracle_c

pytestmark = pytest.mark.gpu

N_C2 = 10_000_000


@pytest.fixture(scope="module")
def c2(cuda_device):
    t = synth_device.make_table(0, 0, N_C2, cuda_device)
    yield t
    del t
    torch.cuda.empty_cache()


def _slice_oracle(seed, first, n, mb, thr):
    h = synth.make_table(seed, first, n)
    pts, valid, arg = oracle_c.bbox_fold(h.poly_off, h.xy)
    high, count = oracle_c.iou_filter(h.img_off, pts, valid, mb, thr)
    return h, pts, valid, high, count


def test_c2_fused_full_size(cuda_device, c2):
    t = c2
    assert t.n_img == N_C2 and 79_000_000 < t.n_poly < 81_000_000 and 1.43e9 < t.n_vert < 1.45e9
    os.environ["DYD_FUSED"] = "tma"
    a = ops.bbox_iou_fused(t.img_off, t.poly_off, t.xy, 2, 0.7)
    # sampled slices (start, middle, ragged end) against the oracle
    for first, n in ((0, 20000), (4_999_990, 20011), (N_C2 - 15001, 15001)):
        h, pts, valid, high, count = _slice_oracle(0, first, n, 2, 0.7)
        q0 = int(t.img_off[first].item()); q1 = int(t.img_off[first + n].item())
        assert q1 - q0 == h.n_poly
        assert a.pts[4 * q0:4 * q1].cpu().numpy().tobytes() == pts.tobytes()
        assert np.array_equal(a.valid[q0:q1].cpu().numpy(), valid)
        assert np.array_equal(a.high[first:first + n].cpu().numpy(), high)
        assert np.array_equal(a.count[first:first + n].cpu().numpy(), count)
    # every polygon of this generator is non-empty; counts add up to the object count
    assert int(a.valid.sum(dtype=torch.int64).item()) == t.n_poly
    assert int(a.count.sum(dtype=torch.int64).item()) == t.n_poly
    n_high_07 = int(a.high.sum(dtype=torch.int64).item())
    assert 0.05 * N_C2 < n_high_07 < 0.2 * N_C2
    # identity across implementations
    pts_a, high_a, count_a = a.pts.clone(), a.high.clone(), a.count.clone()
    del a
    os.environ["DYD_FUSED"] = "direct"
    b = ops.bbox_iou_fused(t.img_off, t.poly_off, t.xy, 2, 0.7)
    assert torch.equal(b.pts.view(torch.int64), pts_a.view(torch.int64)) and torch.equal(b.high, high_a) and torch.equal(b.count, count_a)
    del b
    os.environ.pop("DYD_FUSED", None)
    p1, v1, _ = ops.bbox_minmax(t.poly_off, t.xy)
    assert torch.equal(p1.view(torch.int64), pts_a.view(torch.int64))
    h2, c2_ = ops.iou_filter(t.img_off, p1, v1, 2, 0.7)
    assert torch.equal(h2, high_a) and torch.equal(c2_, count_a)
    # the flag is monotone in the threshold and in min_boxes
    h98, _ = ops.iou_filter(t.img_off, p1, v1, 2, 0.98)
    assert bool((h98 <= high_a).all()) and int(h98.sum(dtype=torch.int64).item()) < n_high_07
    h5, _ = ops.iou_filter(t.img_off, p1, v1, 5, 0.7)
    assert bool((h5 <= high_a).all())
    assert bool((h5[count_a < 5] == 0).all())


def test_c3_dedup_and_antijoin_properties(cuda_device):
    d = cuda_device
    n, n_ref = 20_000_000, 10_000_000
    url_id, off, data = synth_device.make_urls(0, 0, n, d)
    keys = ops.hash_strings(off, data)
    del off, data
    for keep in ("first", "last"):
        km, rep = ops.dedup(keys, None, keep)
        idx = torch.arange(n, device=d)
        assert torch.equal(km.bool(), rep == idx)                      # kept rows represent themselves
        assert torch.equal(rep[rep], rep)                              # representatives are fixed points
        assert torch.equal(keys.view(torch.int64)[rep], keys.view(torch.int64))   # same key as the representative
        assert bool((rep <= idx).all()) if keep == "first" else bool((rep >= idx).all())
        # the URL ids are the ground truth of equality for this generator (hashes of distinct ids differ)
        n_distinct = int(torch.unique(url_id).numel())
        assert int(km.sum(dtype=torch.int64).item()) == n_distinct
        assert torch.equal(url_id[rep], url_id)
    kf, _ = ops.dedup(keys, None, False)
    km, rep = ops.dedup(keys, None, "first")
    cnt = torch.bincount(rep, minlength=n)
    assert torch.equal(kf.bool(), cnt[rep] == 1)
    assert 0.04 < 1 - float(km.float().mean().item()) < 0.06
    # anti-join
    rid, roff, rdata = synth_device.make_urls(0, 0, n_ref, d, n_main_for_ref=n)
    rkeys = ops.hash_strings(roff, rdata)
    keep, rr = ops.antijoin(keys, None, rkeys, None)
    assert torch.equal(keep.bool(), rr < 0)
    hit = rr >= 0
    assert torch.equal(rkeys.view(torch.int64)[rr[hit]], keys.view(torch.int64)[hit])
    in_ref = torch.isin(url_id, rid)
    assert torch.equal(hit, in_ref)
    # first matching reference row
    first_row = torch.full((int(max(url_id.max().item(), rid.max().item())) + 1,), n_ref, dtype=torch.int64, device=d)
    first_row.scatter_reduce_(0, rid, torch.arange(n_ref, device=d), reduce="amin")
    assert torch.equal(rr[hit], first_row[url_id[hit]])


def test_c4_dense_crowd(cuda_device):
    d = cuda_device
    n = 100_000                                    # 35 M boxes; the full config is 10x this, same code path
    io, pts = synth_device.make_crowd(0, 0, n, device=d)
    nbox = io[1:] - io[:-1]
    assert int(nbox.min().item()) >= 200 and int(nbox.max().item()) <= 500
    h0, c0 = ops.iou_filter(io, pts, None, 2, 0.0)            # every pair hits at thr 0
    assert bool((h0 == 1).all()) and torch.equal(c0.long(), nbox)
    h2, _ = ops.iou_filter(io, pts, None, 2, 2.0)             # no pair can reach 2.0: worst case, all pairs evaluated
    assert int(h2.sum().item()) == 0
    h7, _ = ops.iou_filter(io, pts, None, 2, 0.7)
    h9, _ = ops.iou_filter(io, pts, None, 2, 0.9)
    assert bool((h9 <= h7).all())
    hbig, _ = ops.iou_filter(io, pts, None, 501, 0.0)         # min_boxes above every image
    assert int(hbig.sum().item()) == 0
    # sampled images against the oracle
    hio, hpts = synth.make_crowd_boxes(0, 5000, 64)
    a, b = int(io[5000].item()), int(io[5064].item())
    assert pts[4 * a:4 * b].cpu().numpy().tobytes() == hpts.tobytes()
    for thr, got in ((0.7, h7), (0.9, h9)):
        want, _ = oracle_c.iou_filter(hio, hpts, None, 2, thr)
        assert np.array_equal(got[5000:5064].cpu().numpy(), want)
