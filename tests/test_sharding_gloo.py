"""World-size-2 (and 3) CPU tests of the multi-GPU exchange logic over gloo.

The hash-partition all-to-all, the reverse all-to-all and the null-group reduction run for
real; the owner-local table step is a numpy stand-in with the semantics of dyd_dedup_ids /
dyd_antijoin (the CUDA kernels themselves are covered by the -m gpu tests).  The sharded
result must equal the single-table oracle on the concatenated rows.
"""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deal_yolo_daya_b200 import sharding
from oracle import oracle_c


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _local_dedup(keys, ids, mode):
    k = keys.view(torch.int64).numpy(); i = ids.numpy()
    first, last, cnt = {}, {}, {}
    for kk, ii in zip(k, i):
        first[kk] = min(first.get(kk, 1 << 62), ii); last[kk] = max(last.get(kk, -1), ii); cnt[kk] = cnt.get(kk, 0) + 1
    if mode == "first":
        rep = np.array([first[kk] for kk in k], np.int64); keep = rep == i
    elif mode == "last":
        rep = np.array([last[kk] for kk in k], np.int64); keep = rep == i
    else:
        rep = np.array([first[kk] for kk in k], np.int64); keep = np.array([cnt[kk] == 1 for kk in k])
    return torch.from_numpy(keep.astype(np.uint8)), torch.from_numpy(rep)


def _local_antijoin(mk, rk):
    m = mk.view(torch.int64).numpy(); r = rk.view(torch.int64).numpy()
    first = {}
    for j, kk in enumerate(r):
        first.setdefault(kk, j)
    hit = np.array([first.get(kk, -1) for kk in m], np.int64)
    return torch.from_numpy((hit < 0).astype(np.uint8)), torch.from_numpy(hit)


def _table(world):
    rng = np.random.RandomState(17)
    n = 4000
    keys = rng.randint(0, 900, size=n).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    null = (rng.rand(n) < 0.02).astype(np.uint8)
    ref = rng.randint(600, 1500, size=1500).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    rnull = (rng.rand(1500) < 0.05).astype(np.uint8)
    cuts = np.linspace(0, n, world + 1).astype(int); cuts[1] -= 37 if world > 1 else 0
    rcuts = np.linspace(0, 1500, world + 1).astype(int)
    return keys, null, ref, rnull, cuts, rcuts


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        keys, null, ref, rnull, cuts, rcuts = _table(world)
        a, b = cuts[rank], cuts[rank + 1]; ra, rb = rcuts[rank], rcuts[rank + 1]
        res = {}
        for mode in ("first", "last", False):
            keep, rep = sharding.dedup_global(torch.from_numpy(keys[a:b].copy()), torch.from_numpy(null[a:b].copy()),
                                              int(a), mode, local_dedup=_local_dedup)
            res[str(mode)] = (keep.numpy(), rep.numpy())
        keep, rr = sharding.antijoin_global(torch.from_numpy(keys[a:b].copy()), torch.from_numpy(null[a:b].copy()),
                                            torch.from_numpy(ref[ra:rb].copy()), torch.from_numpy(rnull[ra:rb].copy()),
                                            int(ra), local_antijoin=_local_antijoin)
        res["anti"] = (keep.numpy(), rr.numpy())
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_dedup_and_antijoin_equal_single_table(world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    keys, null, ref, rnull, cuts, rcuts = _table(world)
    for mode in ("first", "last", False):
        want_keep, want_rep = oracle_c.dedup(keys, null, mode)
        got_keep = np.concatenate([out[r][str(mode)][0] for r in range(world)])
        got_rep = np.concatenate([out[r][str(mode)][1] for r in range(world)])
        assert np.array_equal(got_keep, want_keep), mode
        assert np.array_equal(got_rep, want_rep), mode
    want_keep, want_rr = oracle_c.antijoin(keys, null, ref, rnull)
    assert np.array_equal(np.concatenate([out[r]["anti"][0] for r in range(world)]), want_keep)
    assert np.array_equal(np.concatenate([out[r]["anti"][1] for r in range(world)]), want_rr)


def test_image_ranges_balance_vertices():
    rng = np.random.RandomState(0)
    sizes = rng.randint(1, 500, size=10000)
    sizes[100:140] = 20000                                   # a crowded stretch
    vp = np.concatenate([[0], np.cumsum(sizes)])
    for world in (1, 2, 4, 8):
        rngs = sharding.image_ranges(vp, world)
        assert rngs[0][0] == 0 and rngs[-1][1] == len(sizes)
        assert all(rngs[i][1] == rngs[i + 1][0] for i in range(world - 1))
        loads = [vp[b] - vp[a] for a, b in rngs]
        assert max(loads) <= vp[-1] / world + 20000


# ---- K6 across ranks: per-category base offsets from one all_gather ------------------------------
def _split_table():
    rng = np.random.RandomState(5)
    n_img, n_vocab, n_cat = 300, 12, 4
    counts = rng.randint(0, 7, size=n_img)
    img_off = np.zeros(n_img + 1, np.int64); img_off[1:] = np.cumsum(counts)
    label = rng.randint(-1, n_vocab, size=int(img_off[-1])).astype(np.int32)
    cat_of_label = rng.randint(-1, n_cat, size=n_vocab).astype(np.int32)
    return img_off, label, cat_of_label, n_cat


def _split_worker(rank, world, port, out):
    from oracle import oracle_np
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        img_off, label, cat_of_label, n_cat = _split_table()
        n_img = len(img_off) - 1
        i0, i1 = sharding.image_ranges(img_off, world)[rank]
        q0 = int(img_off[i0])
        loc_off = img_off[i0:i1 + 1] - q0
        e_img, e_box, e_cat, loc_cat_off = oracle_np.split_expand(loc_off, label[q0:int(img_off[i1])], cat_of_label, n_cat)
        base, cat_off = sharding.split_category_bases(torch.from_numpy(np.diff(loc_cat_off)))
        # global position of every local expanded row: base of its category + its rank inside the local group
        pos = base.numpy()[e_cat] + (np.arange(len(e_cat)) - loc_cat_off[e_cat])
        out[rank] = (pos, e_img + i0, e_box + q0, e_cat, cat_off.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_split_bases_reassemble_the_single_table_order(world):
    from oracle import oracle_np
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_split_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    img_off, label, cat_of_label, n_cat = _split_table()
    w_img, w_box, w_cat, w_off = oracle_np.split_expand(img_off, label, cat_of_label, n_cat)
    g_img = np.full(len(w_img), -1, np.int64); g_box = g_img.copy(); g_cat = np.full(len(w_img), -1, np.int32)
    for r in range(world):
        pos, e_img, e_box, e_cat, cat_off = out[r]
        assert np.array_equal(cat_off, w_off)
        g_img[pos] = e_img; g_box[pos] = e_box; g_cat[pos] = e_cat
    assert np.array_equal(g_img, w_img) and np.array_equal(g_box, w_box) and np.array_equal(g_cat, w_cat)
