"""Python wrappers over the C ABI (include/dyd.h).

PyTorch is used for device memory and streams only: every function takes / returns
``torch.Tensor`` buffers on a CUDA device (or numpy arrays for the ``*_host`` calls) and
forwards raw pointers to libdyd.so.  No computation happens in torch or numpy here, and
there is no CPU path: without a GPU these functions raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

KEEP_MODES = {"first": 0, "last": 1, False: 2}


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not (isinstance(t, torch.Tensor) and t.is_cuda):
            raise _lib.DydError("deal_yolo_daya_b200.ops works on CUDA tensors only (no CPU fallback)")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _chk(t, dtype, name):
    if t is None:
        return None
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def _ws(nbytes, dev):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)


# ------------------------------------------------------------------ K1
def bbox_minmax(poly_off, xy, want_arg=False):
    """Polygon -> two corner points (processor.py:252-260).  Returns (pts, valid, arg|None)."""
    _need_cuda(poly_off, xy)
    lib = _lib.load()
    poly_off = _chk(poly_off, torch.int64, "poly_off"); xy = _chk(xy, torch.float64, "xy")
    dev = poly_off.device
    n = poly_off.numel() - 1
    pts = torch.empty(4 * n, dtype=torch.float64, device=dev)
    valid = torch.empty(n, dtype=torch.uint8, device=dev)
    arg = torch.empty(4 * n, dtype=torch.int32, device=dev) if want_arg else None
    with torch.cuda.device(dev):
        _lib.check(lib.dyd_bbox_minmax(_ptr(poly_off), _ptr(xy), n, _ptr(pts), _ptr(valid), _ptr(arg), _stream(dev)),
                   "dyd_bbox_minmax")
    return pts, valid, arg


# ------------------------------------------------------------------ K2
def iou_filter(img_off, pts, valid, min_boxes=2, thr=0.98, workspace=None):
    """Box-count + any-pair IoU flag per image (processor.py:328-376).  Returns (high, count)."""
    _need_cuda(img_off, pts, valid)
    lib = _lib.load()
    img_off = _chk(img_off, torch.int64, "img_off"); pts = _chk(pts, torch.float64, "pts")
    valid = _chk(valid, torch.uint8, "valid")
    dev = img_off.device
    n = img_off.numel() - 1
    high = torch.empty(n, dtype=torch.uint8, device=dev)
    count = torch.empty(n, dtype=torch.int32, device=dev)
    need = lib.dyd_iou_workspace_bytes(n)
    ws = workspace if workspace is not None and workspace.numel() >= need else _ws(need, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.dyd_iou_filter(_ptr(img_off), _ptr(pts), _ptr(valid), n, int(min_boxes), float(thr),
                                      _ptr(high), _ptr(count), _ptr(ws), ws.numel(), _stream(dev)), "dyd_iou_filter")
    return high, count


# ------------------------------------------------------------------ K1+K2
class FusedBuffers:
    """Reusable outputs + workspace of the fused call (allocation kept out of timed loops)."""

    def __init__(self, n_img, n_poly, device, want_arg=False):
        lib = _lib.load()
        self.pts = torch.empty(4 * n_poly, dtype=torch.float64, device=device)
        self.valid = torch.empty(n_poly, dtype=torch.uint8, device=device)
        self.arg = torch.empty(4 * n_poly, dtype=torch.int32, device=device) if want_arg else None
        self.high = torch.empty(n_img, dtype=torch.uint8, device=device)
        self.count = torch.empty(n_img, dtype=torch.int32, device=device)
        self.ws = _ws(lib.dyd_iou_workspace_bytes(n_img), device)


def bbox_iou_fused(img_off, poly_off, xy, min_boxes=2, thr=0.98, want_arg=False, out: FusedBuffers | None = None,
                   max_ctas: int = 0, prepass_event=None):
    """Fused K1+K2 over a device-resident CSR table.  Returns a FusedBuffers.

    ``max_ctas`` (0 = every SM) caps the persistent grid so that a second stream -- the URL hash / dedup /
    exchange chain -- finds free SMs while this kernel runs (dyd_bbox_iou_fused_ex)."""
    _need_cuda(img_off, poly_off, xy)
    lib = _lib.load()
    img_off = _chk(img_off, torch.int64, "img_off"); poly_off = _chk(poly_off, torch.int64, "poly_off")
    xy = _chk(xy, torch.float64, "xy")
    dev = img_off.device
    n_img = img_off.numel() - 1
    n_poly = poly_off.numel() - 1
    if out is None:
        out = FusedBuffers(n_img, n_poly, dev, want_arg)
    with torch.cuda.device(dev):
        _lib.check(lib.dyd_bbox_iou_fused_ex(_ptr(img_off), _ptr(poly_off), _ptr(xy), n_img, n_poly, int(min_boxes),
                                             float(thr), _ptr(out.pts), _ptr(out.valid), _ptr(out.arg), _ptr(out.high),
                                             _ptr(out.count), _ptr(out.ws), out.ws.numel(), int(max_ctas),
                                             C.c_void_p(prepass_event.cuda_event) if prepass_event is not None else None, _stream(dev)),
                   "dyd_bbox_iou_fused_ex")
    return out


def fused_cta_times():
    """(start, end) globaltimer ns of every CTA of the last staged fused launch (diagnostics); rows of never-used CTAs are 0."""
    lib = _lib.load()
    a = np.zeros(296, np.uint64)
    _lib.check(lib.dyd_fused_cta_times(C.c_void_p(a.ctypes.data), 296), "dyd_fused_cta_times")
    return a.reshape(148, 2)


def fused_tile_modes(buf: FusedBuffers, n_img: int):
    """(staged, direct, deferred) tile counts of the last fused call that used ``buf`` (diagnostics)."""
    lib = _lib.load()
    dev = buf.ws.device
    counts = torch.empty(3, dtype=torch.uint64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.dyd_fused_tile_modes(_ptr(buf.ws), n_img, _ptr(counts), _stream(dev)), "dyd_fused_tile_modes")
    return tuple(int(v) for v in counts.cpu().numpy())


# ------------------------------------------------------------------ K0 / K4 / K5
def hash_strings(off, data):
    _need_cuda(off, data)
    lib = _lib.load()
    off = _chk(off, torch.int64, "off"); data = _chk(data, torch.uint8, "data")
    dev = off.device
    n = off.numel() - 1
    out = torch.empty(n, dtype=torch.uint64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.dyd_hash_strings(_ptr(off), _ptr(data), n, _ptr(out), _stream(dev)), "dyd_hash_strings")
    return out


def dedup(keys, null=None, keep="first", row_id=None, workspace=None):
    """drop_duplicates keep-mask (processor.py:140-144).  Returns (keep uint8, rep int64)."""
    _need_cuda(keys, null, row_id)
    lib = _lib.load()
    keys = _chk(keys, torch.uint64, "keys"); null = _chk(null, torch.uint8, "null")
    row_id = _chk(row_id, torch.int64, "row_id")
    dev = keys.device
    n = keys.numel()
    km = torch.empty(n, dtype=torch.uint8, device=dev)
    rep = torch.empty(n, dtype=torch.int64, device=dev)
    need = lib.dyd_dedup_workspace_bytes(n)
    ws = workspace if workspace is not None and workspace.numel() >= need else _ws(need, dev)
    mode = KEEP_MODES[keep]
    with torch.cuda.device(dev):
        if row_id is None:
            rc = lib.dyd_dedup(_ptr(keys), _ptr(null), n, mode, _ptr(km), _ptr(rep), _ptr(ws), ws.numel(), _stream(dev))
        else:
            if null is not None:
                raise ValueError("null cells are resolved by the caller in the sharded (row_id) form")
            rc = lib.dyd_dedup_ids(_ptr(keys), _ptr(row_id), n, mode, _ptr(km), _ptr(rep), _ptr(ws), ws.numel(), _stream(dev))
        _lib.check(rc, "dyd_dedup")
    return km, rep


def antijoin(main_keys, main_null, ref_keys, ref_null, workspace=None):
    """~main.isin(set(ref.dropna())) (processor.py:194-199).  Returns (keep uint8, ref_row int64)."""
    _need_cuda(main_keys, main_null, ref_keys, ref_null)
    lib = _lib.load()
    main_keys = _chk(main_keys, torch.uint64, "main_keys"); ref_keys = _chk(ref_keys, torch.uint64, "ref_keys")
    main_null = _chk(main_null, torch.uint8, "main_null"); ref_null = _chk(ref_null, torch.uint8, "ref_null")
    dev = main_keys.device
    n, nr = main_keys.numel(), ref_keys.numel()
    km = torch.empty(n, dtype=torch.uint8, device=dev)
    rr = torch.empty(n, dtype=torch.int64, device=dev)
    need = lib.dyd_antijoin_workspace_bytes(nr)
    ws = workspace if workspace is not None and workspace.numel() >= need else _ws(need, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.dyd_antijoin(_ptr(main_keys), _ptr(main_null), n, _ptr(ref_keys), _ptr(ref_null), nr,
                                    _ptr(km), _ptr(rr), _ptr(ws), ws.numel(), _stream(dev)), "dyd_antijoin")
    return km, rr


def url_filter(main_keys, main_null, ref_keys, ref_null, keep="first", workspace=None):
    """Steps 2 + 3 on the same main keys in one call: (keep, rep) of `dedup` and (keep_ref, ref_row) of `antijoin`."""
    _need_cuda(main_keys, main_null, ref_keys, ref_null)
    lib = _lib.load()
    main_keys = _chk(main_keys, torch.uint64, "main_keys"); ref_keys = _chk(ref_keys, torch.uint64, "ref_keys")
    main_null = _chk(main_null, torch.uint8, "main_null"); ref_null = _chk(ref_null, torch.uint8, "ref_null")
    dev = main_keys.device
    n, nr = main_keys.numel(), ref_keys.numel()
    km = torch.empty(n, dtype=torch.uint8, device=dev); rep = torch.empty(n, dtype=torch.int64, device=dev)
    ka = torch.empty(n, dtype=torch.uint8, device=dev); rr = torch.empty(n, dtype=torch.int64, device=dev)
    need = lib.dyd_url_filter_workspace_bytes(n, nr)
    ws = workspace if workspace is not None and workspace.numel() >= need else _ws(need, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.dyd_url_filter(_ptr(main_keys), _ptr(main_null), n, _ptr(ref_keys), _ptr(ref_null), nr, KEEP_MODES[keep],
                                      _ptr(km), _ptr(rep), _ptr(ka), _ptr(rr), _ptr(ws), ws.numel(), _stream(dev)), "dyd_url_filter")
    return km, rep, ka, rr


# ------------------------------------------------------------------ K3 / K6 / YOLO
COUNTER_NAMES = ("total_objects", "missing_name_objects", "total_labels", "replaced_labels",
                 "replaced_objects", "replaced_rows")


def label_lut(img_off, label_id, lut_new, lut_ntok, lut_nrep):
    """Name rewrite through LUTs (processor.py:582-602).  Returns (new_id, row_replaced, counters tensor[6])."""
    _need_cuda(img_off, label_id, lut_new, lut_ntok, lut_nrep)
    lib = _lib.load()
    dev = img_off.device
    n_img = img_off.numel() - 1
    n_box = label_id.numel()
    new_id = torch.empty(n_box, dtype=torch.int32, device=dev)
    row_rep = torch.empty(n_img, dtype=torch.uint8, device=dev)
    counters = torch.empty(6, dtype=torch.uint64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.dyd_label_lut(_ptr(_chk(img_off, torch.int64, "img_off")), _ptr(_chk(label_id, torch.int32, "label_id")),
                                     n_img, n_box, _ptr(_chk(lut_new, torch.int32, "lut_new")),
                                     _ptr(_chk(lut_ntok, torch.int32, "lut_ntok")), _ptr(_chk(lut_nrep, torch.int32, "lut_nrep")),
                                     lut_new.numel(), _ptr(new_id), _ptr(row_rep), _ptr(counters), _stream(dev)),
                   "dyd_label_lut")
    return new_id, row_rep, counters


def label_hist(label_id, n_vocab):
    """Occurrences of each vocabulary id (uint64[n_vocab])."""
    _need_cuda(label_id)
    lib = _lib.load()
    dev = label_id.device
    hist = torch.empty(max(n_vocab, 1), dtype=torch.uint64, device=dev)[:n_vocab]
    with torch.cuda.device(dev):
        _lib.check(lib.dyd_label_hist(_ptr(_chk(label_id, torch.int32, "label_id")), label_id.numel(), n_vocab,
                                      _ptr(hist), _stream(dev)), "dyd_label_hist")
    return hist


def label_presence(img_off, label_id, n_vocab):
    """(img_hist, box_hist) uint64[n_vocab]: images holding each id at least once / objects per id (processor.py:1113-1131)."""
    _need_cuda(img_off, label_id)
    lib = _lib.load()
    dev = img_off.device
    hists = torch.empty(2, max(n_vocab, 1), dtype=torch.uint64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.dyd_label_presence(_ptr(_chk(img_off, torch.int64, "img_off")), _ptr(_chk(label_id, torch.int32, "label_id")),
                                          img_off.numel() - 1, n_vocab, _ptr(hists[0]), _ptr(hists[1]), _stream(dev)), "dyd_label_presence")
    return hists[0, :n_vocab], hists[1, :n_vocab]


MAX_CAT_PER_PASS = 256          # dyd_split_count / _fill keep their per-category counters in shared memory


def split_expand(img_off, label_id, cat_of_label, n_cat):
    """Category expansion (processor.py:751-775).  Returns (exp_img, exp_box, exp_cat, cat_off).
    More than 256 categories (a rules table in two_column mode can hold any number; the reference has no limit) are
    expanded in blocks of 256: the rows of a block's categories with the other labels masked out, blocks in category order."""
    _need_cuda(img_off, label_id, cat_of_label)
    if n_cat > MAX_CAT_PER_PASS:
        parts, offs, base = [], [], 0
        for c0 in range(0, n_cat, MAX_CAT_PER_PASS):
            c1 = min(n_cat, c0 + MAX_CAT_PER_PASS)
            sub = torch.where((cat_of_label >= c0) & (cat_of_label < c1), cat_of_label - c0, torch.full_like(cat_of_label, -1))
            ei, eb, ec, co = split_expand(img_off, label_id, sub.contiguous(), c1 - c0)
            parts.append((ei, eb, ec + c0))
            offs.append(co[:-1] + base)
            base += int(co[-1].item())
        offs.append(torch.tensor([base], dtype=torch.int64, device=img_off.device))
        return (torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts]), torch.cat([p[2] for p in parts]), torch.cat(offs))
    lib = _lib.load()
    dev = img_off.device
    n_img = img_off.numel() - 1
    n_vocab = cat_of_label.numel()
    cat_off = torch.empty(n_cat + 1, dtype=torch.int64, device=dev)
    ws = _ws(lib.dyd_split_workspace_bytes(n_img, n_cat), dev)
    a = (_ptr(_chk(img_off, torch.int64, "img_off")), n_img, _ptr(_chk(label_id, torch.int32, "label_id")),
         _ptr(_chk(cat_of_label, torch.int32, "cat_of_label")), n_vocab, n_cat)
    with torch.cuda.device(dev):
        _lib.check(lib.dyd_split_count(*a, _ptr(cat_off), _ptr(ws), ws.numel(), _stream(dev)), "dyd_split_count")
        n_exp = int(cat_off[-1].item())          # the one host read the API needs (sizes the outputs)
        exp_img = torch.empty(n_exp, dtype=torch.int64, device=dev)
        exp_box = torch.empty(n_exp, dtype=torch.int64, device=dev)
        exp_cat = torch.empty(n_exp, dtype=torch.int32, device=dev)
        if n_exp:
            _lib.check(lib.dyd_split_fill(*a, _ptr(cat_off), _ptr(exp_img), _ptr(exp_box), _ptr(exp_cat),
                                          _ptr(ws), ws.numel(), _stream(dev)), "dyd_split_fill")
    return exp_img, exp_box, exp_cat, cat_off


def split_assign(cat_off, perm, n_train, n_val):
    """Split id + shuffled position per expanded row from the host permutation (processor.py:800-806)."""
    _need_cuda(cat_off, perm, n_train, n_val)
    lib = _lib.load()
    dev = cat_off.device
    n_cat = cat_off.numel() - 1
    n_exp = perm.numel()
    split = torch.empty(n_exp, dtype=torch.uint8, device=dev)
    pos = torch.empty(n_exp, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.dyd_split_assign(_ptr(_chk(cat_off, torch.int64, "cat_off")), n_cat, _ptr(_chk(perm, torch.int64, "perm")),
                                        n_exp, _ptr(_chk(n_train, torch.int64, "n_train")), _ptr(_chk(n_val, torch.int64, "n_val")),
                                        _ptr(split), _ptr(pos), _stream(dev)), "dyd_split_assign")
    return split, pos


def split_assign_range(cat_off, perm, n_train, n_val, own_lo, own_cnt, local_off):
    """Sharded split_assign: split id + shuffled position of THIS rank's expanded rows (category c: rows own_lo[c] ..
    own_lo[c] + own_cnt[c] of the global category, kept at local_off[c] .. in the outputs) from the global permutation."""
    _need_cuda(cat_off, perm, n_train, n_val, own_lo, own_cnt, local_off)
    lib = _lib.load()
    dev = cat_off.device
    n_cat = cat_off.numel() - 1
    n_local = int((own_cnt.sum()).item())
    split = torch.empty(n_local, dtype=torch.uint8, device=dev)
    pos = torch.empty(n_local, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.dyd_split_assign_range(_ptr(_chk(cat_off, torch.int64, "cat_off")), n_cat, _ptr(_chk(perm, torch.int64, "perm")), perm.numel(),
                                              _ptr(_chk(n_train, torch.int64, "n_train")), _ptr(_chk(n_val, torch.int64, "n_val")),
                                              _ptr(_chk(own_lo, torch.int64, "own_lo")), _ptr(_chk(own_cnt, torch.int64, "own_cnt")),
                                              _ptr(_chk(local_off, torch.int64, "local_off")), _ptr(split), _ptr(pos), _stream(dev)),
                   "dyd_split_assign_range")
    return split, pos


def yolo_normalise(img_off, pts, valid, img_wh):
    """cx, cy, w, h per box (processor.py:1045-1052).  Returns (cxcywh float64[4*n_box], ok uint8)."""
    _need_cuda(img_off, pts, valid, img_wh)
    lib = _lib.load()
    dev = img_off.device
    n_img = img_off.numel() - 1
    n_box = pts.numel() // 4
    out = torch.empty(4 * n_box, dtype=torch.float64, device=dev)
    ok = torch.empty(n_box, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.dyd_yolo_normalise(_ptr(_chk(img_off, torch.int64, "img_off")), _ptr(_chk(pts, torch.float64, "pts")),
                                          _ptr(_chk(valid, torch.uint8, "valid")), _ptr(_chk(img_wh, torch.float64, "img_wh")),
                                          n_img, n_box, _ptr(out), _ptr(ok), _stream(dev)), "dyd_yolo_normalise")
    return out, ok


# ------------------------------------------------------------------ host-buffer entry points
def _np(a, dtype, name):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        if a.is_cuda:
            raise ValueError(f"{name}: host entry points take host buffers")
        a = a.numpy()
    a = np.asarray(a)
    if a.dtype != dtype or not a.flags.c_contiguous:
        raise TypeError(f"{name} must be a C-contiguous {np.dtype(dtype).name} array")
    return a


def _hp(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def bbox_iou_host(img_off, poly_off, xy, min_boxes=2, thr=0.98, want_pts=True, want_arg=False,
                  chunk_images=0, out=None, device=0, tile_modes=False):
    """Host CSR in -> host results out, H2D / kernels / D2H pipelined inside the library.

    Returns dict(pts, valid, arg, high, count) of numpy arrays (pinned if ``out`` supplies them).
    """
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise _lib.DydError("no CUDA device: the hot path has no CPU fallback")
    img_off = _np(img_off, np.int64, "img_off"); poly_off = _np(poly_off, np.int64, "poly_off")
    xy = _np(xy, np.float64, "xy")
    n_img = len(img_off) - 1
    n_poly = len(poly_off) - 1
    out = dict(out or {})
    if want_pts and out.get("pts") is None:
        out["pts"] = np.empty(4 * n_poly, np.float64)
    if out.get("valid") is None:
        out["valid"] = np.empty(n_poly, np.uint8)
    if want_arg and out.get("arg") is None:
        out["arg"] = np.empty(4 * n_poly, np.int32)
    if out.get("high") is None:
        out["high"] = np.empty(n_img, np.uint8)
    if out.get("count") is None:
        out["count"] = np.empty(n_img, np.int32)
    pts = _np(out.get("pts"), np.float64, "pts") if want_pts else None
    arg = _np(out.get("arg"), np.int32, "arg") if want_arg else None
    modes = np.zeros(3, np.int64) if tile_modes else None
    with torch.cuda.device(device):
        _lib.check(lib.dyd_bbox_iou_host_ex(_hp(img_off), _hp(poly_off), _hp(xy), n_img, int(min_boxes), float(thr),
                                            _hp(pts), _hp(_np(out["valid"], np.uint8, "valid")), _hp(arg),
                                            _hp(_np(out["high"], np.uint8, "high")), _hp(_np(out["count"], np.int32, "count")),
                                            int(chunk_images), _hp(modes)), "dyd_bbox_iou_host_ex")
    if tile_modes:
        out["tile_modes"] = {"staged": int(modes[0]), "direct": int(modes[1]), "deferred": int(modes[2])}
    return out


def dedup_host(off, data, null=None, keep="first", device=0, out=None):
    """Host Arrow string buffers in -> (keep uint8, rep int64) numpy arrays out (``out`` = preallocated, e.g. pinned, pair)."""
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise _lib.DydError("no CUDA device: the hot path has no CPU fallback")
    off = _np(off, np.int64, "off"); data = _np(data, np.uint8, "data"); null = _np(null, np.uint8, "null")
    n = len(off) - 1
    km, rep = out if out is not None else (np.empty(n, np.uint8), np.empty(n, np.int64))
    with torch.cuda.device(device):
        _lib.check(lib.dyd_dedup_host(_hp(off), _hp(data), _hp(null), n, KEEP_MODES[keep], _hp(km), _hp(rep)),
                   "dyd_dedup_host")
    return km, rep


def antijoin_host(off, data, null, ref_off, ref_data, ref_null, device=0, out=None):
    """Host Arrow string buffers of the main and the reference `source` columns in -> (keep uint8, ref_row int64) out."""
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise _lib.DydError("no CUDA device: the hot path has no CPU fallback")
    off = _np(off, np.int64, "off"); data = _np(data, np.uint8, "data"); null = _np(null, np.uint8, "null")
    ref_off = _np(ref_off, np.int64, "ref_off"); ref_data = _np(ref_data, np.uint8, "ref_data"); ref_null = _np(ref_null, np.uint8, "ref_null")
    n, n_ref = len(off) - 1, len(ref_off) - 1
    km, rr = out if out is not None else (np.empty(n, np.uint8), np.empty(n, np.int64))
    with torch.cuda.device(device):
        _lib.check(lib.dyd_antijoin_host(_hp(off), _hp(data), _hp(null), n, _hp(ref_off), _hp(ref_data), _hp(ref_null), n_ref,
                                         _hp(km), _hp(rr)), "dyd_antijoin_host")
    return km, rr
