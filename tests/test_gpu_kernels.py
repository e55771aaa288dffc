"""CUDA kernels (through the C ABI) against the CPU oracle -- bit-exact.

Every comparison is on raw bytes for doubles (signed zeros and NaN payloads included) and
exact equality for masks / indices.  Sizes are chosen so the C oracle finishes in seconds;
the full-size properties live in test_gpu_scale.py.
"""
from __future__ import annotations

import os

import numpy as np
import pytest
import torch

from deal_yolo_daya_b200 import ops, synth, synth_device
from oracle import oracle_c, oracle_np
from tests import tables

pytestmark = pytest.mark.gpu


def dev(a, d):
    return torch.from_numpy(np.ascontiguousarray(a)).to(d)


def host(t):
    return t.cpu().numpy()


def assert_bits(a, b, what=""):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.tobytes() != b.tobytes():
        bad = np.where(a.view(np.uint64) != b.view(np.uint64))[0] if a.dtype == np.float64 else np.where(a != b)[0]
        raise AssertionError(f"{what}: {len(bad)} mismatches, first at {bad[:8]}: got {a[bad[:8]]} want {b[bad[:8]]}")


VARIANTS = (("tma", "4"), ("direct", "4"), ("direct", "8"), ("direct", "2"))   # (DYD_FUSED, DYD_GROUP)


def set_variant(fused, group):
    os.environ["DYD_FUSED"] = fused
    os.environ["DYD_GROUP"] = group


@pytest.fixture(autouse=True)
def _default_variant():
    yield
    os.environ.pop("DYD_FUSED", None); os.environ.pop("DYD_GROUP", None)


def check_bbox(poly_off, xy, d):
    want_pts, want_valid, want_arg = oracle_c.bbox_fold(poly_off, xy)
    for group in ("4", "8", "2"):
        set_variant("tma", group)
        for want_arg_flag in (False, True):
            pts, valid, arg = ops.bbox_minmax(dev(poly_off, d), dev(xy, d), want_arg=want_arg_flag)
            assert_bits(host(pts), want_pts, f"pts(arg={want_arg_flag}, G={group})")
            assert_bits(host(valid), want_valid, "valid")
            if want_arg_flag:
                assert_bits(host(arg), want_arg, "arg")
    return want_pts, want_valid


def check_fused(img_off, poly_off, xy, d, params=((2, 0.7), (2, 0.98), (1, 0.0), (3, 0.5))):
    want_pts, want_valid, want_arg = oracle_c.bbox_fold(poly_off, xy)
    for mb, thr in params:
        want_high, want_count = oracle_c.iou_filter(img_off, want_pts, want_valid, mb, thr)
        for fused, group in VARIANTS:
            set_variant(fused, group)
            for flag in (False, True):
                tag = f"{fused}/G{group}/arg={flag} mb={mb} thr={thr}"
                out = ops.bbox_iou_fused(dev(img_off, d), dev(poly_off, d), dev(xy, d), mb, thr, want_arg=flag)
                assert_bits(host(out.pts), want_pts, "fused pts " + tag)
                assert_bits(host(out.valid), want_valid, "fused valid " + tag)
                assert_bits(host(out.high), want_high, "fused high " + tag)
                assert_bits(host(out.count), want_count, "fused count " + tag)
                if flag:
                    assert_bits(host(out.arg), want_arg, "fused arg " + tag)
        high, count = ops.iou_filter(dev(img_off, d), dev(want_pts, d), dev(want_valid, d), mb, thr)
        assert_bits(host(high), want_high, f"k2 high mb={mb} thr={thr}")
        assert_bits(host(count), want_count, "k2 count")


def test_edge_polygons(cuda_device):
    img_off, poly_off, xy = tables.edge_polygon_table()
    check_bbox(poly_off, xy, cuda_device)
    check_fused(img_off, poly_off, xy, cuda_device)
    # the small table is also checked by the pure-Python oracle (CPython semantics, not C)
    pts, valid, arg = ops.bbox_minmax(dev(poly_off, cuda_device), dev(xy, cuda_device), want_arg=True)
    p, v, a = oracle_np.bbox_fold(poly_off, xy)
    assert_bits(host(pts), p); assert_bits(host(valid), v); assert_bits(host(arg), a)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_polygons_with_specials(cuda_device, seed):
    img_off, poly_off, xy = tables.random_polygon_table(seed, 400, max_polys=14, max_verts=70)
    check_bbox(poly_off, xy, cuda_device)
    check_fused(img_off, poly_off, xy, cuda_device)


@pytest.mark.parametrize("n_img,max_polys,max_verts", [(997, 3, 9), (240, 80, 6), (61, 6, 900), (13, 300, 5), (1201, 10, 33)])
def test_fused_tile_shapes(cuda_device, n_img, max_polys, max_verts):
    """Tile geometry of the staged kernel: ragged tails, stages that overflow (vertex or object
    capacity -> fallback lane), crowded images inside fast tiles."""
    img_off, poly_off, xy = tables.random_polygon_table(n_img + max_polys, n_img, max_polys=max_polys,
                                                        max_verts=max_verts, p_empty=0.02, p_special=0.01)
    check_fused(img_off, poly_off, xy, cuda_device, params=((2, 0.7), (1, 0.0)))


def test_empty_and_tiny(cuda_device):
    d = cuda_device
    z = np.zeros(1, np.int64)
    pts, valid, arg = ops.bbox_minmax(dev(z, d), dev(np.zeros(0), d), want_arg=True)
    assert pts.numel() == 0 and valid.numel() == 0
    out = ops.bbox_iou_fused(dev(z, d), dev(z, d), dev(np.zeros(0), d))
    assert out.high.numel() == 0
    img_off, poly_off, xy = tables.csr_from_polygons([[], [], [[(1.0, 1.0)]]])
    check_fused(img_off, poly_off, xy, d)


@pytest.mark.parametrize("seed,n", [(0, 20000), (5, 3000)])
def test_synth_table_device_equals_numpy_and_oracle(cuda_device, seed, n):
    d = cuda_device
    t = synth.make_table(seed, 1000, n)
    g = synth_device.make_table(seed, 1000, n, d)
    assert_bits(host(g.img_off), t.img_off, "img_off")
    assert_bits(host(g.poly_off), t.poly_off, "poly_off")
    assert_bits(host(g.xy), t.xy, "xy")
    assert_bits(host(g.label_id), t.label_id, "label_id")
    want_pts, want_valid, _ = oracle_c.bbox_fold(t.poly_off, t.xy)
    for mb, thr in ((2, 0.7), (2, 0.98)):
        want_high, want_count = oracle_c.iou_filter(t.img_off, want_pts, want_valid, mb, thr)
        for fused, group in VARIANTS[:2]:
            set_variant(fused, group)
            out = ops.bbox_iou_fused(g.img_off, g.poly_off, g.xy, mb, thr)
            assert_bits(host(out.pts), want_pts, fused); assert_bits(host(out.valid), want_valid, fused)
            assert_bits(host(out.high), want_high, fused); assert_bits(host(out.count), want_count, fused)
        assert 0.02 < want_high.mean() < 0.5          # the generator's jittered copies do trigger the filter


@pytest.mark.parametrize("lo,hi,n", [(0, 9, 3000), (30, 70, 600), (60, 130, 300), (200, 500, 60), (1000, 1100, 6)])
def test_iou_box_tables(cuda_device, lo, hi, n):
    d = cuda_device
    img_off, pts, valid = tables.random_box_table(lo * 7 + n, n, lo, hi)
    for mb, thr in ((2, 0.7), (2, 0.98), (5, 0.3)):
        want_high, want_count = oracle_c.iou_filter(img_off, pts, valid, mb, thr)
        high, count = ops.iou_filter(dev(img_off, d), dev(pts, d), dev(valid, d), mb, thr)
        assert_bits(host(high), want_high, f"high {lo}-{hi} mb={mb} thr={thr}")
        assert_bits(host(count), want_count, "count")
    high, count = ops.iou_filter(dev(img_off, d), dev(pts, d), None, 2, 0.7)
    want_high, want_count = oracle_c.iou_filter(img_off, pts, None, 2, 0.7)
    assert_bits(host(high), want_high); assert_bits(host(count), want_count)


def test_crowd_binned_form_edge_images(cuda_device):
    """The binned form of the block-per-image kernel (33 .. 512 boxes) on images built to stress its counting sort: every box
    with the same x1 (one bin), x1 spread over 1e300 (scale underflow), infinite coordinates, boxes that only touch (x2 == x1 of
    the next: no overlap), exact duplicates far apart in box order, a hit only between the first and the last box."""
    d = cuda_device
    rng = np.random.RandomState(3)
    imgs = []

    def boxes(n, x1, w, y1=None, h=None):
        y1 = rng.uniform(0, 1000, n) if y1 is None else y1
        h = rng.uniform(5, 40, n) if h is None else h
        return np.stack([x1, y1, x1 + w, y1 + h], axis=1)
    n = 200
    imgs.append(boxes(n, np.full(n, 7.0), rng.uniform(5, 40, n)))                                # one x1 value
    imgs.append(boxes(n, rng.uniform(0, 1900, n) * 1e298, rng.uniform(5, 40, n) * 1e298))       # huge range
    b = boxes(n, rng.uniform(0, 1900, n), rng.uniform(5, 40, n)); b[5, 0] = -np.inf; b[9, 2] = np.inf; imgs.append(b)
    x = np.arange(n) * 10.0
    imgs.append(boxes(n, x, np.full(n, 10.0), np.zeros(n), np.full(n, 10.0)))                    # a row of touching boxes
    b = boxes(n, rng.uniform(0, 1900, n), rng.uniform(5, 40, n)); b[199] = b[0]; imgs.append(b)  # first == last
    b = boxes(n, rng.uniform(0, 1900, n), rng.uniform(5, 40, n)); b[120] = b[17] + np.array([0.0, 0.0, 0.001, 0.0]); imgs.append(b)
    imgs.append(boxes(40, rng.uniform(0, 50, 40), rng.uniform(5, 40, 40), rng.uniform(0, 50, 40)))   # everything overlaps: the queue fills
    imgs.append(boxes(512, rng.uniform(0, 1900, 512), rng.uniform(5, 40, 512)))                  # the largest binned image
    imgs.append(boxes(513, rng.uniform(0, 1900, 513), rng.uniform(5, 40, 513)))                  # the smallest all-pairs image
    img_off = np.zeros(len(imgs) + 1, np.int64); np.cumsum([len(b) for b in imgs], out=img_off[1:])
    pts = np.concatenate(imgs).reshape(-1).astype(np.float64)
    for mb, thr in ((2, 0.7), (2, 0.98), (2, 1.0), (2, 1e-300), (300, 0.5)):
        want_high, want_count = oracle_c.iou_filter(img_off, pts, None, mb, thr)
        high, count = ops.iou_filter(dev(img_off, d), dev(pts, d), None, mb, thr)
        assert_bits(host(high), want_high, f"high mb={mb} thr={thr}"); assert_bits(host(count), want_count, "count")
    assert oracle_c.iou_filter(img_off, pts, None, 2, 0.98)[0][4] == 1


def test_crowd_generator_and_worst_case(cuda_device):
    d = cuda_device
    io, pts = synth.make_crowd_boxes(3, 10, 40)
    gio, gpts = synth_device.make_crowd(3, 10, 40, device=d)
    assert_bits(host(gio), io); assert_bits(host(gpts), pts)
    for thr in (0.7, 2.0):                      # thr 2.0: no pair can hit -> every pair is evaluated
        want_high, want_count = oracle_c.iou_filter(io, pts, None, 2, thr)
        high, count = ops.iou_filter(gio, gpts, None, 2, thr)
        assert_bits(host(high), want_high); assert_bits(host(count), want_count)


def test_near_threshold_pairs_take_exact_division(cuda_device):
    """Pairs whose IoU sits within a few ulps of the threshold must agree with the oracle's division."""
    d = cuda_device
    rng = np.random.RandomState(11)
    imgs = []
    thr = 0.7
    for _ in range(4000):
        w = rng.uniform(1, 50); h = rng.uniform(1, 50)
        # box B = same width, height scaled so that IoU = inter/union is ~thr
        k = thr * (1 + rng.uniform(-3, 3) * 2.0 ** -52)
        imgs.append([[(0.0, 0.0), (w, h)], [(0.0, 0.0), (w, h * k)]])
    img_off, poly_off, xy = tables.csr_from_polygons(imgs)
    pts, valid, _ = oracle_c.bbox_fold(poly_off, xy)
    want_high, _ = oracle_c.iou_filter(img_off, pts, valid, 2, thr)
    assert 0.2 < want_high.mean() < 0.8
    high, _ = ops.iou_filter(dev(img_off, d), dev(pts, d), dev(valid, d), 2, thr)
    assert_bits(host(high), want_high)


# ---------------------------------------------------------------- K0 / K4 / K5
def test_hash_strings(cuda_device):
    d = cuda_device
    strs = ["", "a", "ab", "abcdefg", "abcdefgh", "abcdefghi", "https://img.example.com/123.jpg", "行人/车.jpg" * 3]
    strs += [synth.url_of(i) for i in range(0, 3000, 7)] + ["x" * k for k in range(1, 40)]
    off, data = oracle_c.pack_strings(strs)
    want = oracle_c.hash_strings_buf(off, data)
    assert np.array_equal(want, oracle_np.hash_strings(strs))
    got = ops.hash_strings(dev(off, d), dev(data, d))
    assert_bits(host(got), want, "hash")


@pytest.fixture
def small_partitions(monkeypatch):
    """The partitioned dedup path normally starts at 2 M rows; let 64 K-row inputs take it."""
    monkeypatch.setenv("DYD_DEDUP_PARTITION_MIN", "65536")


@pytest.mark.parametrize("keep", ["first", "last", False])
@pytest.mark.parametrize("n,groups,pnull", [(1, 1, 0.0), (5000, 700, 0.02), (200000, 150000, 0.001), (3000, 3000, 1.0)])
def test_dedup(cuda_device, small_partitions, keep, n, groups, pnull):
    d = cuda_device
    rng = np.random.RandomState(n + groups)
    pool = rng.randint(0, 2 ** 63, size=groups, dtype=np.int64).astype(np.uint64)
    pool[0] = np.uint64(0xFFFFFFFFFFFFFFFF)             # the table's EMPTY sentinel as a real key
    keys = pool[rng.randint(0, groups, size=n)]
    null = (rng.rand(n) < pnull).astype(np.uint8)
    want_keep, want_rep = oracle_c.dedup(keys, null, keep)
    got_keep, got_rep = ops.dedup(dev(keys, d), dev(null, d), keep)
    assert_bits(host(got_keep), want_keep, "keep"); assert_bits(host(got_rep), want_rep, "rep")
    got_keep, got_rep = ops.dedup(dev(keys, d), None, keep)
    want_keep, want_rep = oracle_c.dedup(keys, np.zeros(n, np.uint8), keep)
    assert_bits(host(got_keep), want_keep); assert_bits(host(got_rep), want_rep)


def test_dedup_with_row_ids(cuda_device):
    d = cuda_device
    rng = np.random.RandomState(4)
    n = 50000
    keys = rng.randint(0, 9000, size=n).astype(np.uint64)
    ids = rng.permutation(n).astype(np.int64) * 3 + 11
    keep, rep = ops.dedup(dev(keys, d), None, "first", row_id=dev(ids, d))
    first = {}
    for k, i in zip(keys, ids):
        first[k] = min(first.get(k, 1 << 62), i)
    want_rep = np.array([first[k] for k in keys], np.int64)
    assert_bits(host(rep), want_rep); assert_bits(host(keep), (want_rep == ids).astype(np.uint8))


@pytest.mark.parametrize("keep", ["first", "last", False])
def test_dedup_partition_overflow_falls_back_on_device(cuda_device, small_partitions, keep):
    """Large inputs go through the partitioned path; one key repeated thousands of times overflows its
    partition and the gated global-table kernels must produce the answer inside the same call."""
    d = cuda_device
    rng = np.random.RandomState(77)
    n = 300000
    keys = rng.randint(0, 2 ** 63, size=n, dtype=np.int64).astype(np.uint64)
    hot = rng.choice(n, size=6000, replace=False)
    keys[hot] = np.uint64(0x1234567890ABCDEF)
    keys[rng.choice(n, size=20000)] = keys[rng.choice(n, size=20000)]
    null = (rng.rand(n) < 0.003).astype(np.uint8)
    want_keep, want_rep = oracle_c.dedup(keys, null, keep)
    got_keep, got_rep = ops.dedup(dev(keys, d), dev(null, d), keep)
    assert_bits(host(got_keep), want_keep, "keep"); assert_bits(host(got_rep), want_rep, "rep")


@pytest.mark.parametrize("keep", ["first", "last", False])
def test_dedup_with_row_ids_partitioned(cuda_device, small_partitions, keep):
    d = cuda_device
    rng = np.random.RandomState(41)
    n = 150000
    keys = rng.randint(0, 40000, size=n).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    ids = rng.permutation(n).astype(np.int64) * 5 + (1 << 33)
    ids[rng.choice(n, size=500, replace=False)] = -1                  # bucket padding
    got_keep, got_rep = ops.dedup(dev(keys, d), None, keep, row_id=dev(ids, d))
    live = ids >= 0
    first, last, cnt = {}, {}, {}
    for k, i in zip(keys[live], ids[live]):
        first[k] = min(first.get(k, 1 << 62), i); last[k] = max(last.get(k, -1), i); cnt[k] = cnt.get(k, 0) + 1
    want_rep = np.full(n, -1, np.int64); want_keep = np.zeros(n, np.uint8)
    pick = last if keep == "last" else first
    want_rep[live] = [pick[k] for k in keys[live]]
    want_keep[live] = [cnt[k] == 1 for k in keys[live]] if keep is False else (want_rep[live] == ids[live])
    assert_bits(host(got_rep), want_rep, "rep"); assert_bits(host(got_keep), want_keep, "keep")


@pytest.mark.parametrize("n,nr", [(4000, 0), (4000, 900), (100000, 60000), (300000, 0), (250000, 170000)])
def test_antijoin(cuda_device, n, nr):
    d = cuda_device
    rng = np.random.RandomState(n + nr)
    mk = rng.randint(0, 120000, size=n).astype(np.uint64)
    rk = rng.randint(0, 120000, size=nr).astype(np.uint64)
    if nr:
        mk[7] = rk[3] = np.uint64(0xFFFFFFFFFFFFFFFF)                        # the EMPTY sentinel as a real key
    mn = (rng.rand(n) < 0.01).astype(np.uint8); rn = (rng.rand(nr) < 0.05).astype(np.uint8)
    want_keep, want_rr = oracle_c.antijoin(mk, mn, rk, rn)
    keep, rr = ops.antijoin(dev(mk, d), dev(mn, d), dev(rk, d), dev(rn, d))
    assert_bits(host(keep), want_keep); assert_bits(host(rr), want_rr)
    keep, rr = ops.antijoin(dev(mk, d), None, dev(rk, d), None)
    want_keep, want_rr = oracle_c.antijoin(mk, np.zeros(n, np.uint8), rk, np.zeros(nr, np.uint8))
    assert_bits(host(keep), want_keep); assert_bits(host(rr), want_rr)


@pytest.mark.parametrize("keep", ["first", "last", False])
@pytest.mark.parametrize("n,nr,hot", [(3000, 1200, 0), (200000, 90000, 0), (300000, 0, 0), (250000, 170000, 5000), (120000, 400000, 0)])
def test_url_filter_joint_equals_dedup_plus_antijoin(cuda_device, small_partitions, keep, n, nr, hot):
    """dyd_url_filter (one shared-memory table per key partition answering both questions) against the oracle's dedup and
    anti-join: null cells on both sides, the EMPTY sentinel as a key, duplicate reference keys (smallest row wins), a key
    repeated thousands of times (partition overflow -> both gated global-table fallbacks inside the same call)."""
    d = cuda_device
    rng = np.random.RandomState(n + nr + hot)
    mk = rng.randint(0, 150000, size=n).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    rk = rng.randint(100000, 260000, size=nr).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    if nr:
        mk[7] = rk[3] = np.uint64(0xFFFFFFFFFFFFFFFF)
    if hot:
        mk[rng.choice(n, size=hot, replace=False)] = np.uint64(0x1234567890ABCDEF)
        rk[rng.choice(nr, size=300, replace=False)] = np.uint64(0x1234567890ABCDEF)
    mn = (rng.rand(n) < 0.01).astype(np.uint8); rn = (rng.rand(nr) < 0.05).astype(np.uint8)
    for main_null, ref_null in ((mn, rn), (None, None)):
        zm = main_null if main_null is not None else np.zeros(n, np.uint8)
        zr = ref_null if ref_null is not None else np.zeros(nr, np.uint8)
        want_keep, want_rep = oracle_c.dedup(mk, zm, keep)
        want_ka, want_rr = oracle_c.antijoin(mk, zm, rk, zr)
        km, rep, ka, rr = ops.url_filter(dev(mk, d), None if main_null is None else dev(main_null, d), dev(rk, d),
                                         None if ref_null is None else dev(ref_null, d), keep)
        assert_bits(host(km), want_keep, "keep"); assert_bits(host(rep), want_rep, "rep")
        assert_bits(host(ka), want_ka, "keep_ref"); assert_bits(host(rr), want_rr, "ref_row")


def test_url_filter_records_equals_separate_record_kernels(cuda_device, small_partitions):
    """The sharded form on (key, id) records with bucket padding: same four answers as dyd_dedup_records followed by
    dyd_antijoin_records; reference records are padding afterwards when asked."""
    from deal_yolo_daya_b200 import _lib
    from deal_yolo_daya_b200.ops import _ptr, _stream
    lib = _lib.load(); d = cuda_device
    rng = np.random.RandomState(5)
    m, mr = 180000, 90000
    rec = np.empty((m, 2), np.int64); ref = np.empty((mr, 2), np.int64)
    rec[:, 0] = (rng.randint(0, 70000, size=m).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)).view(np.int64)
    rec[:, 1] = rng.permutation(m).astype(np.int64) * 7 + (1 << 34)
    ref[:, 0] = (rng.randint(50000, 120000, size=mr).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)).view(np.int64)
    ref[:, 1] = rng.permutation(mr).astype(np.int64) * 3 + (1 << 35)
    rec[rng.choice(m, 9000, replace=False), 1] = -1; ref[rng.choice(mr, 4000, replace=False), 1] = -1          # bucket padding
    s = _stream(d)
    t = lambda n, dt=torch.int64: torch.empty(n, dtype=dt, device=d)   # noqa: E731
    big_rec, big_ref = rec.copy(), ref.copy()
    for keep, small_ids in (("first", False), ("last", False), (False, False), ("first", True), ("last", True), (False, True)):
        rec, ref = big_rec.copy(), big_ref.copy()
        id_bound = 0                                   # ids beyond 2^31: 64-bit rows in the shared-memory tables
        if small_ids:                                  # ids below 2^31 and a bound that says so: the 32-bit form
            rec[:, 1] = np.where(rec[:, 1] >= 0, (rec[:, 1] - (1 << 34)) // 7 * 5 + 11, -1)
            ref[:, 1] = np.where(ref[:, 1] >= 0, (ref[:, 1] - (1 << 35)) // 3 * 9 + 4, -1)
            id_bound = int(max(rec[:, 1].max(), ref[:, 1].max())) + 1
        mode = ops.KEEP_MODES[keep]
        drec, dref = dev(rec.reshape(-1), d), dev(ref.reshape(-1), d)
        k1, r1, k2, r2 = t(m, torch.uint8), t(m), t(m, torch.uint8), t(m)
        ws_d = t(lib.dyd_dedup_workspace_bytes(m), torch.uint8); ws_a = t(lib.dyd_antijoin_workspace_bytes(mr), torch.uint8)
        _lib.check(lib.dyd_dedup_records(_ptr(drec), m, mode, _ptr(k1), _ptr(r1), _ptr(ws_d), ws_d.numel(), s), "dedup_records")
        _lib.check(lib.dyd_antijoin_records(_ptr(dref), mr, _ptr(drec), m, _ptr(k2), _ptr(r2), _ptr(ws_a), ws_a.numel(), 0, s), "antijoin_records")
        j = [t(m, torch.uint8), t(m), t(m, torch.uint8), t(m)]
        ws = t(lib.dyd_url_filter_workspace_bytes(m, mr), torch.uint8)
        _lib.check(lib.dyd_url_filter_records(_ptr(dref), mr, _ptr(drec), m, mode, _ptr(j[0]), _ptr(j[1]), _ptr(j[2]), _ptr(j[3]), _ptr(ws), ws.numel(), 1, id_bound, s),
                   "url_filter_records")
        assert_bits(host(j[0]), host(k1), f"keep {keep}"); assert_bits(host(j[1]), host(r1), f"rep {keep}")
        assert_bits(host(j[2]), host(k2), f"keep_ref {keep}"); assert_bits(host(j[3]), host(r2), f"ref_row {keep}")
        assert bool((dref.view(-1, 2)[:, 1] == -1).all()), "reference records must be padding after the call"


def test_url_pipeline_device(cuda_device):
    """Device-generated URL bytes -> hash -> dedup equals the oracle on the numpy twin's URLs."""
    d = cuda_device
    n = 30000
    url_id, off, data = synth_device.make_urls(9, 0, n, d)
    ids = synth.url_ids_of(9, np.arange(n))
    assert_bits(host(url_id), ids)
    strs = [synth.url_of(i) for i in ids]
    woff, wdata = oracle_c.pack_strings(strs)
    assert_bits(host(off), woff); assert_bits(host(data), wdata)
    keys = ops.hash_strings(off, data)
    want_keys = oracle_c.hash_strings_buf(woff, wdata)
    assert_bits(host(keys), want_keys)
    keep, rep = ops.dedup(keys, None, "first")
    want_keep, want_rep = oracle_c.dedup(want_keys, np.zeros(n, np.uint8), "first")
    assert_bits(host(keep), want_keep); assert_bits(host(rep), want_rep)
    assert 0.03 < 1 - want_keep.mean() < 0.07
    # reference set with 10 % overlap
    rid, roff, rdata = synth_device.make_urls(9, 0, n // 2, d, n_main_for_ref=n)
    assert_bits(host(rid), synth.ref_ids_of(9, np.arange(n // 2), n))
    rkeys = ops.hash_strings(roff, rdata)
    k2, rr = ops.antijoin(keys, None, rkeys, None)
    wk2, wrr = oracle_c.antijoin(want_keys, np.zeros(n, np.uint8), host(rkeys), np.zeros(n // 2, np.uint8))
    assert_bits(host(k2), wk2); assert_bits(host(rr), wrr)


# ---------------------------------------------------------------- K3 / K6 / YOLO
def test_label_lut(cuda_device):
    d = cuda_device
    t = synth.make_table(2, 0, 5000)
    lab = t.label_id.copy()
    lab[::97] = -1
    nv = 100
    rng = np.random.RandomState(0)
    lut_new = rng.randint(0, nv, nv).astype(np.int32); lut_ntok = rng.randint(0, 4, nv).astype(np.int32)
    lut_nrep = np.minimum(lut_ntok, rng.randint(0, 3, nv)).astype(np.int32)
    want_new, want_rr, want_c = oracle_c.label_lut(t.img_off, lab, lut_new, lut_ntok, lut_nrep)
    new, rr, cnt = ops.label_lut(dev(t.img_off, d), dev(lab, d), dev(lut_new, d), dev(lut_ntok, d), dev(lut_nrep, d))
    assert_bits(host(new), want_new); assert_bits(host(rr), want_rr)
    assert dict(zip(ops.COUNTER_NAMES, [int(x) for x in host(cnt)])) == want_c
    w2, r2, c2 = oracle_np.label_lut(t.img_off[:200], lab[:t.img_off[199]], lut_new, lut_ntok, lut_nrep)
    assert np.array_equal(w2, want_new[:t.img_off[199]])


@pytest.mark.parametrize("n_img,n_cat", [(1, 1), (300, 4), (70000, 20), (5000, 256), (60000, 16), (40000, 3), (9000, 1), (6000, 257), (8000, 700)])
def test_split_expand_and_assign(cuda_device, n_img, n_cat):
    d = cuda_device
    t = synth.make_table(n_cat, 0, n_img)
    lab = t.label_id.copy(); lab[::53] = -1
    rng = np.random.RandomState(n_cat)
    cat = rng.randint(-1, n_cat, synth.N_LABELS).astype(np.int32)
    if n_cat > 256:                                      # more categories than one pass holds: a wide label vocabulary
        lab = (lab.astype(np.int64) * 37 + np.arange(len(lab)) % 1500).astype(np.int32) % 1500; lab[::53] = -1
        cat = rng.randint(-1, n_cat, 1500).astype(np.int32)
    wi, wb, wc, woff = oracle_c.split_expand(t.img_off, lab, cat, n_cat)
    ei, eb, ec, coff = ops.split_expand(dev(t.img_off, d), dev(lab, d), dev(cat, d), n_cat)
    assert_bits(host(coff), woff); assert_bits(host(ei), wi); assert_bits(host(eb), wb); assert_bits(host(ec), wc)
    if n_img <= 300:
        ni, nb, nc, noff = oracle_np.split_expand(t.img_off, lab, cat, n_cat)
        assert np.array_equal(ni, wi) and np.array_equal(nb, wb) and np.array_equal(noff, woff)
    want_split, want_pos = oracle_np.split_assign(woff, 0.8, 0.1, 0.1, 42)
    perm = np.concatenate([np.random.RandomState(42).permutation(int(woff[c + 1] - woff[c])) for c in range(n_cat)] + [np.zeros(0, np.int64)]).astype(np.int64)
    sizes = np.diff(woff)
    ntr = np.array([int(n * 0.8) for n in sizes], np.int64); nva = np.array([int(n * 0.1) for n in sizes], np.int64)
    if len(perm):
        split, pos = ops.split_assign(coff, dev(perm, d), dev(ntr, d), dev(nva, d))
        assert_bits(host(split), want_split); assert_bits(host(pos), want_pos)
        # sharded form: three pretend ranks own consecutive slices of every category and together reproduce the single-table answer
        cuts = [np.array([int(n * f) for n in sizes], np.int64) for f in (0.0, 0.3, 0.85, 1.0)]
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            cnt = hi - lo
            loc = np.zeros(n_cat, np.int64); loc[1:] = np.cumsum(cnt)[:-1]
            if cnt.sum() == 0:
                continue
            s2, p2 = ops.split_assign_range(coff, dev(perm, d), dev(ntr, d), dev(nva, d), dev(lo, d), dev(cnt, d), dev(loc, d))
            s2, p2 = host(s2), host(p2)
            for c in range(n_cat):
                a = int(woff[c] + lo[c])
                assert_bits(s2[loc[c]:loc[c] + cnt[c]], want_split[a:a + cnt[c]]); assert_bits(p2[loc[c]:loc[c] + cnt[c]], want_pos[a:a + cnt[c]])


def test_yolo_normalise(cuda_device):
    d = cuda_device
    img_off, poly_off, xy = tables.random_polygon_table(8, 300)
    pts, valid, _ = oracle_c.bbox_fold(poly_off, xy)
    n_img = len(img_off) - 1
    wh = np.tile(np.array([1920.0, 1080.0]), n_img); wh[10] = 0.0
    out, ok = ops.yolo_normalise(dev(img_off, d), dev(pts, d), dev(valid, d), dev(wh, d))
    out, ok = host(out), host(ok)
    for i in range(n_img):
        for q in range(img_off[i], img_off[i + 1]):
            w = oracle_np.yolo_norm(tuple(pts[4 * q:4 * q + 4]), wh[2 * i], wh[2 * i + 1]) if valid[q] and wh[2 * i] and wh[2 * i + 1] else None
            assert bool(ok[q]) == (w is not None), (i, q)
            if w is not None:
                assert np.array(w).tobytes() == out[4 * q:4 * q + 4].tobytes()


# ---------------------------------------------------------------- host-buffer entry points
def test_host_entry_points(cuda_device):
    t = synth.make_table(12, 0, 40000)
    want_pts, want_valid, want_arg = oracle_c.bbox_fold(t.poly_off, t.xy)
    want_high, want_count = oracle_c.iou_filter(t.img_off, want_pts, want_valid, 2, 0.7)
    for chunk in (0, 1777, 1111, 100000):
        out = ops.bbox_iou_host(t.img_off, t.poly_off, t.xy, 2, 0.7, want_pts=True, want_arg=True, chunk_images=chunk, tile_modes=True)
        assert_bits(out["pts"], want_pts); assert_bits(out["valid"], want_valid); assert_bits(out["arg"], want_arg)
        assert_bits(out["high"], want_high); assert_bits(out["count"], want_count)
        # every chunk -- also those that start at an odd object number or deep inside the table -- takes the bulk-copy staged
        # kernel: at most the last tile of a chunk may fall back to direct loads (its 16-byte slices would leave the arrays)
        tm = out["tile_modes"]
        n_chunks = -(-40000 // (chunk or 16384))
        assert tm["deferred"] == 0 and tm["direct"] <= 2 * n_chunks and tm["staged"] > 6000, tm
    strs = [synth.url_of(i) for i in synth.url_ids_of(12, np.arange(20000))]
    off, data = oracle_c.pack_strings(strs)
    null = np.zeros(len(strs), np.uint8); null[5] = null[77] = 1
    keep, rep = ops.dedup_host(off, data, null, "first")
    wk, wr = oracle_c.dedup(oracle_c.hash_strings_buf(off, data), null, "first")
    assert_bits(keep, wk); assert_bits(rep, wr)
    rstrs = [synth.url_of(i) for i in synth.ref_ids_of(12, np.arange(9000), 20000)]
    roff, rdata = oracle_c.pack_strings(rstrs)
    rnull = np.zeros(len(rstrs), np.uint8); rnull[3] = 1
    keep, row = ops.antijoin_host(off, data, null, roff, rdata, rnull)
    wk, wr = oracle_c.antijoin(oracle_c.hash_strings_buf(off, data), null, oracle_c.hash_strings_buf(roff, rdata), rnull)
    assert_bits(keep, wk); assert_bits(row, wr)
    assert 0 < int(wk.sum()) < len(strs)


def test_fused_with_fewer_ctas_and_dynamic_segments(cuda_device):
    """The fused kernel claims its segments dynamically: any CTA cap gives the same bits (dyd_bbox_iou_fused_ex)."""
    d = cuda_device
    t = synth.make_table(21, 5, 30000)
    want_pts, want_valid, _ = oracle_c.bbox_fold(t.poly_off, t.xy)
    want_high, want_count = oracle_c.iou_filter(t.img_off, want_pts, want_valid, 2, 0.7)
    io, po, xy = dev(t.img_off, d), dev(t.poly_off, d), dev(t.xy, d)
    for ctas in (0, 1, 3, 47, 140, 148, 1000):
        out = ops.bbox_iou_fused(io, po, xy, 2, 0.7, max_ctas=ctas)
        assert_bits(host(out.pts), want_pts, f"pts ctas={ctas}"); assert_bits(host(out.valid), want_valid)
        assert_bits(host(out.high), want_high, f"high ctas={ctas}"); assert_bits(host(out.count), want_count)
    modes = ops.fused_tile_modes(out, 30000)
    assert modes[0] > 0 and sum(modes) >= 30000 // 6


def test_exchange_kernels_single_rank_roundtrip(cuda_device, small_partitions):
    """bucket -> (identity exchange) -> dedup on records -> reply -> unpack equals plain dedup;
    world=4 bucket layout is exercised by treating the 4 buckets as arriving from 4 ranks."""
    import ctypes as C
    from deal_yolo_daya_b200 import _lib
    from deal_yolo_daya_b200.ops import _ptr, _stream
    lib = _lib.load(); d = cuda_device
    rng = np.random.RandomState(5)
    n, world = 60000, 4
    keys = (rng.randint(0, 20000, size=n).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15))
    for keep in ("first", "last", False):
        want_keep, want_rep = oracle_c.dedup(keys, np.zeros(n, np.uint8), keep)
        want_rep = want_rep + 1000                               # global ids = row_base + row
        cap = n // world + 2000
        m = world * cap
        t = lambda *shape, dt=torch.int64: torch.empty(*shape, dtype=dt, device=d)   # noqa: E731
        send, reply = t(2 * m), t(2 * m)
        cursors, overflow = t(world, dt=torch.uint64), t(1, dt=torch.int32)
        keep_r, rep_r = t(m, dt=torch.uint8), t(m)
        ws = t(lib.dyd_dedup_workspace_bytes(m), dt=torch.uint8)
        out_keep, out_rep = t(n, dt=torch.uint8), t(n)
        k = dev(keys, d); s = _stream(d)
        _lib.check(lib.dyd_shard_bucket(_ptr(k), None, 1000, n, world, cap, _ptr(send), _ptr(cursors), _ptr(overflow), s), "bucket")
        assert int(overflow.item()) == 0 and int(cursors.cpu().numpy().astype(np.int64).sum()) == n
        _lib.check(lib.dyd_dedup_records(_ptr(send), m, ops.KEEP_MODES[keep], _ptr(keep_r), _ptr(rep_r), _ptr(ws), ws.numel(), s), "records")
        _lib.check(lib.dyd_shard_pack_reply(_ptr(send), _ptr(keep_r), _ptr(rep_r), m, _ptr(reply), 0, s), "pack")
        _lib.check(lib.dyd_shard_unpack(_ptr(reply), m, 1000, n, _ptr(out_keep), _ptr(out_rep), 0, s), "unpack")
        assert_bits(host(out_keep), want_keep, f"keep {keep}"); assert_bits(host(out_rep), want_rep, f"rep {keep}")
        # peer-memory forms with this GPU as all four "ranks": one rank's records fill region 0 of a 4-region buffer
        # per owner; here world = 1 keeps it a self-exchange (region 0 of the own buffer), run twice so that the
        # second pass works on records the first pass's pack kernel reset to padding
        cap1 = n + 512
        recv, back = torch.full((2 * cap1,), -1, dtype=torch.int64, device=d), torch.full((cap1,), -1, dtype=torch.int64, device=d)
        sent = t(cap1, dt=torch.int32); cur1 = t(1, dt=torch.uint64)
        peers_recv = torch.tensor([recv.data_ptr()], dtype=torch.int64, device=d); peers_back = torch.tensor([back.data_ptr()], dtype=torch.int64, device=d)
        kr1, rr1 = t(cap1, dt=torch.uint8), t(cap1)
        ws1 = t(lib.dyd_dedup_workspace_bytes(cap1), dt=torch.uint8)
        for _ in range(2):
            _lib.check(lib.dyd_shard_bucket_p2p(_ptr(k), None, 1000, n, 1, 0, cap1, _ptr(peers_recv), _ptr(sent), _ptr(cur1), _ptr(overflow), s), "bucket_p2p")
            _lib.check(lib.dyd_dedup_records(_ptr(recv), cap1, ops.KEEP_MODES[keep], _ptr(kr1), _ptr(rr1), _ptr(ws1), ws1.numel(), s), "records")
            _lib.check(lib.dyd_shard_pack_reply_p2p(_ptr(recv), _ptr(kr1), _ptr(rr1), cap1, cap1, 0, _ptr(peers_back), 0, 1, s), "pack_p2p")
            out_keep.fill_(7); out_rep.fill_(-7)
            _lib.check(lib.dyd_shard_unpack_p2p(_ptr(back), _ptr(sent), _ptr(cur1), 1, cap1, n, _ptr(out_keep), _ptr(out_rep), 0, s), "unpack_p2p")
            assert_bits(host(out_keep), want_keep, f"p2p keep {keep}"); assert_bits(host(out_rep), want_rep, f"p2p rep {keep}")
            assert bool((recv.view(-1, 2)[:, 1] == -1).all()), "pack kernel must leave the receive buffer padded"
    # overflow is reported, never silent
    _lib.check(lib.dyd_shard_bucket(_ptr(k), None, 0, n, world, 100, _ptr(send), _ptr(cursors), _ptr(overflow), s), "bucket")
    assert int(overflow.item()) == 1


def test_antijoin_records_roundtrip(cuda_device):
    """Sharded anti-join on one GPU: both tables bucketed into 4 fixed-capacity regions of (key, id) records, the owner-side
    table step on records (dyd_antijoin_records), answers packed in anti-join mode and unpacked -- equals the plain anti-join
    with global reference rows; the reference records are reset to padding by the build kernel."""
    from deal_yolo_daya_b200 import _lib
    from deal_yolo_daya_b200.ops import _ptr, _stream
    lib = _lib.load(); d = cuda_device
    rng = np.random.RandomState(11)
    n, nr, world = 50000, 21000, 4
    keys = rng.randint(0, 30000, size=n).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    ref = rng.randint(20000, 45000, size=nr).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    mnull = (rng.rand(n) < 0.01).astype(np.uint8); rnull = (rng.rand(nr) < 0.03).astype(np.uint8)
    want_keep, want_row = oracle_c.antijoin(keys, mnull, ref, rnull)
    want_row = np.where(want_row >= 0, want_row + 500, -1)                   # global reference rows = 500 + row
    cap, capr = n // world + 2000, nr // world + 2000
    m, mr = world * cap, world * capr
    t = lambda *shape, dt=torch.int64: torch.empty(*shape, dtype=dt, device=d)   # noqa: E731
    send, sendr, reply = t(2 * m), t(2 * mr), t(2 * m)
    cursors, overflow = t(world, dt=torch.uint64), t(2, dt=torch.int32)
    keep_r, row_r = t(m, dt=torch.uint8), t(m)
    ws = t(lib.dyd_antijoin_workspace_bytes(mr), dt=torch.uint8)
    s = _stream(d)
    _lib.check(lib.dyd_shard_bucket(_ptr(dev(ref, d)), _ptr(dev(rnull, d)), 500, nr, world, capr, _ptr(sendr), _ptr(cursors), _ptr(overflow[1:]), s), "bucket ref")
    _lib.check(lib.dyd_shard_bucket(_ptr(dev(keys, d)), _ptr(dev(mnull, d)), 9000, n, world, cap, _ptr(send), _ptr(cursors), _ptr(overflow[:1]), s), "bucket main")
    assert overflow.cpu().tolist() == [0, 0]
    _lib.check(lib.dyd_antijoin_records(_ptr(sendr), mr, _ptr(send), m, _ptr(keep_r), _ptr(row_r), _ptr(ws), ws.numel(), 1, s), "antijoin_records")
    assert bool((sendr.view(-1, 2)[:, 1] == -1).all()), "build kernel must reset the reference records"
    _lib.check(lib.dyd_shard_pack_reply(_ptr(send), _ptr(keep_r), _ptr(row_r), m, _ptr(reply), 1, s), "pack")
    out_keep = torch.ones(n, dtype=torch.uint8, device=d); out_row = torch.full((n,), -1, dtype=torch.int64, device=d)   # NaN rows never travel
    _lib.check(lib.dyd_shard_unpack(_ptr(reply), m, 9000, n, _ptr(out_keep), _ptr(out_row), 1, s), "unpack")
    assert_bits(host(out_keep), want_keep, "keep"); assert_bits(host(out_row), want_row, "ref_row")


def test_label_presence(cuda_device):
    """Per-image label presence + per-box counts (summarize_yolo_label_counts) against a plain count on the numpy twin."""
    from tests.oracle_kernels import OracleKernels
    d = cuda_device
    t = synth.make_table(31, 0, 3000)
    lid = t.label_id.copy()
    lid[::17] = -1; lid[5::23] = 200                        # ids outside the vocabulary are ignored
    ih, bh = ops.label_presence(dev(t.img_off, d), dev(lid, d), synth.N_LABELS)
    wi, wb = OracleKernels().label_presence(t.img_off, lid, synth.N_LABELS)
    assert_bits(host(ih).astype(np.int64), wi, "image counts"); assert_bits(host(bh).astype(np.int64), wb, "box counts")
    ih, bh = ops.label_presence(dev(np.zeros(1, np.int64), d), dev(np.zeros(0, np.int32), d), 4)
    assert host(ih).sum() == 0 and host(bh).sum() == 0
