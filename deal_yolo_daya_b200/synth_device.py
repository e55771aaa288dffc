"""Device-side synthetic tables: the CUDA twin of ``synth.py`` (bit-identical output).

Used by bench.py and the at-scale GPU tests so that 10 M-image tables (23 GB of vertices)
are created directly in HBM.  torch supplies buffers and the two prefix sums; every value
comes from the dyd_synth_* kernels (csrc/synth.cu).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib, synth
from .ops import _ptr, _stream


@dataclass
class DeviceTable:
    seed: int
    first_img: int
    img_off: torch.Tensor
    poly_off: torch.Tensor
    xy: torch.Tensor
    label_id: torch.Tensor

    @property
    def n_img(self):
        return self.img_off.numel() - 1

    @property
    def n_poly(self):
        return self.poly_off.numel() - 1

    @property
    def n_vert(self):
        return self.xy.numel() // 2


def _excl_offsets(counts: torch.Tensor) -> torch.Tensor:
    off = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=counts.device)
    torch.cumsum(counts, 0, out=off[1:])
    return off


def make_table(seed: int, first_img: int, n_img: int, device="cuda") -> DeviceTable:
    lib = _lib.load_synth()
    dev = torch.device(device)
    thr = torch.from_numpy(synth.poisson8_thresholds()).to(dev)
    with torch.cuda.device(dev):
        s = _stream(dev)
        npoly = torch.empty(n_img, dtype=torch.int64, device=dev)
        _lib.check(lib.dyd_synth_counts(C.c_uint64(seed), first_img, n_img, _ptr(thr), thr.numel(), _ptr(npoly), s), "dyd_synth_counts")
        img_off = _excl_offsets(npoly)
        del npoly
        n_poly = int(img_off[-1].item())
        nvert = torch.empty(n_poly, dtype=torch.int64, device=dev)
        _lib.check(lib.dyd_synth_nvert(C.c_uint64(seed), first_img, n_img, _ptr(img_off), _ptr(nvert), s), "dyd_synth_nvert")
        poly_off = _excl_offsets(nvert)
        del nvert
        n_vert = int(poly_off[-1].item())
        xy = torch.empty(2 * n_vert, dtype=torch.float64, device=dev)
        label_id = torch.empty(n_poly, dtype=torch.int32, device=dev)
        _lib.check(lib.dyd_synth_fill(C.c_uint64(seed), first_img, n_img, _ptr(img_off), _ptr(poly_off), _ptr(xy), _ptr(label_id), s),
                   "dyd_synth_fill")
    return DeviceTable(seed, first_img, img_off, poly_off, xy, label_id)


def make_urls(seed: int, first_row: int, n: int, device="cuda", n_main_for_ref: int = -1):
    """URL column as Arrow buffers on the device: (url_id int64[n], off int64[n+1], bytes uint8)."""
    lib = _lib.load_synth()
    dev = torch.device(device)
    with torch.cuda.device(dev):
        s = _stream(dev)
        url_id = torch.empty(n, dtype=torch.int64, device=dev)
        ln = torch.empty(n, dtype=torch.int64, device=dev)
        _lib.check(lib.dyd_synth_urls(C.c_uint64(seed), first_row, n, n_main_for_ref, _ptr(url_id), _ptr(ln), s), "dyd_synth_urls")
        off = _excl_offsets(ln)
        del ln
        nbytes = int(off[-1].item())
        data = torch.empty(nbytes + 8, dtype=torch.uint8, device=dev)[:nbytes]
        _lib.check(lib.dyd_synth_url_bytes(_ptr(url_id), _ptr(off), n, _ptr(data), s), "dyd_synth_url_bytes")
    return url_id, off, data


def make_crowd(seed: int, first_img: int, n_img: int, lo: int = 200, hi: int = 500, device="cuda"):
    """Dense-crowd boxes (config C4): (img_off int64[n_img+1], pts float64[4*n_box])."""
    lib = _lib.load_synth()
    dev = torch.device(device)
    with torch.cuda.device(dev):
        s = _stream(dev)
        nbox = torch.empty(n_img, dtype=torch.int64, device=dev)
        _lib.check(lib.dyd_synth_crowd(C.c_uint64(seed), first_img, n_img, lo, hi, None, _ptr(nbox), None, s), "dyd_synth_crowd")
        img_off = _excl_offsets(nbox)
        n_box = int(img_off[-1].item())
        pts = torch.empty(4 * n_box, dtype=torch.float64, device=dev)
        _lib.check(lib.dyd_synth_crowd(C.c_uint64(seed), first_img, n_img, lo, hi, _ptr(img_off), None, _ptr(pts), s), "dyd_synth_crowd")
    return img_off, pts
