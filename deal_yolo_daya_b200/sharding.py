"""Multi-GPU form of the hot path: one process per GPU over torch.distributed.

Rows are partitioned across ranks by image (contiguous global ranges).  K1 / K2 / K3 are
independent per image and need no communication.  Dedup (K4) and the
anti-join (K5) have one real exchange step each: every rank buckets its (hash64, global row
id) records by owner rank = mix(hash) mod P, the buckets travel with one variable-size
all-to-all (NCCL over NVLink on GPUs, gloo in the CPU tests), the owner runs the local hash
table kernel -- global row ids make ``atomicMin`` pick the global first occurrence -- and the
keep bit + representative row travel back with the reverse all-to-all.

torch is plumbing here: the bucket permutation (a stable sort of owner ids), the collectives
and buffer ownership.  The table work itself is the CUDA kernels behind ``ops.dedup`` /
``ops.antijoin``; the ``local_*`` parameters exist so that the CPU (gloo) tests can exercise
this exchange logic with a stand-in defined under tests/.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

_GOLD = -7046029254386353131        # 0x9E3779B97F4A7C15 as int64


def owner_of(keys_i64: torch.Tensor, world: int) -> torch.Tensor:
    """Owner rank of each key (keys viewed as int64; wrap-around multiply, top bits)."""
    mixed = keys_i64 * _GOLD
    top = (mixed >> 33) & 0x7FFFFFFF
    return (top % world).to(torch.int64)


def _exchange(send: torch.Tensor, send_counts: torch.Tensor, group=None):
    """Variable-size all-to-all of rows of `send` (already ordered by destination rank)."""
    world = dist.get_world_size(group)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc = send_counts.tolist(); rc = recv_counts.tolist()
    recv = torch.empty((sum(rc),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=sc, group=group)
    assert len(sc) == world
    return recv, sc, rc


def _bucket(keys_i64, world):
    owner = owner_of(keys_i64, world)
    order = torch.sort(owner, stable=True).indices
    counts = torch.bincount(owner, minlength=world).to(torch.int64)
    return order, counts


def dedup_global(keys: torch.Tensor, null: torch.Tensor | None, row_base: int, keep="first",
                 group=None, local_dedup=None):
    """Global drop_duplicates keep-mask for this rank's rows (processor.py:140-144).

    keys uint64[n] (this rank's rows, global row id = row_base + local index), null uint8[n].
    Returns (keep uint8[n], rep int64[n]) exactly as the single-GPU ``ops.dedup`` would on the
    concatenated table.
    """
    if local_dedup is None:
        from . import ops
        local_dedup = lambda k, ids, mode: ops.dedup(k, None, mode, row_id=ids)  # noqa: E731
    world = dist.get_world_size(group)
    dev = keys.device
    n = keys.numel()
    ids = torch.arange(row_base, row_base + n, dtype=torch.int64, device=dev)
    k64 = keys.view(torch.int64)
    live = torch.ones(n, dtype=torch.bool, device=dev) if null is None else (null == 0)
    live_idx = torch.nonzero(live).squeeze(1)
    order, counts = _bucket(k64[live_idx], world)
    sel = live_idx[order]
    send = torch.stack([k64[sel], ids[sel]], dim=1).contiguous()
    recv, sc, rc = _exchange(send, counts.to(dev), group)
    if recv.shape[0]:
        rkeep, rrep = local_dedup(recv[:, 0].contiguous().view(torch.uint64), recv[:, 1].contiguous(), keep)
    else:
        rkeep = torch.empty(0, dtype=torch.uint8, device=dev); rrep = torch.empty(0, dtype=torch.int64, device=dev)
    back = torch.stack([rkeep.to(torch.int64), rrep], dim=1).contiguous()
    ans = torch.empty((send.shape[0], 2), dtype=torch.int64, device=dev)
    dist.all_to_all_single(ans, back, output_split_sizes=sc, input_split_sizes=rc, group=group)
    keep_out = torch.zeros(n, dtype=torch.uint8, device=dev)
    rep_out = torch.zeros(n, dtype=torch.int64, device=dev)
    keep_out[sel] = ans[:, 0].to(torch.uint8)
    rep_out[sel] = ans[:, 1]
    # null cells are one global group: first / last / count across ranks
    if null is not None:
        nidx = torch.nonzero(~live).squeeze(1)
        big = torch.iinfo(torch.int64).max
        stats = torch.tensor([ids[nidx].min().item() if nidx.numel() else big,
                              -(ids[nidx].max().item()) if nidx.numel() else big,
                              ], dtype=torch.int64, device=dev)
        dist.all_reduce(stats, op=dist.ReduceOp.MIN, group=group)
        cnt = torch.tensor([nidx.numel()], dtype=torch.int64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
        first, last, total = int(stats[0]), -int(stats[1]), int(cnt[0])
        if nidx.numel():
            if keep == "first":
                rep_out[nidx] = first; keep_out[nidx] = (ids[nidx] == first).to(torch.uint8)
            elif keep == "last":
                rep_out[nidx] = last; keep_out[nidx] = (ids[nidx] == last).to(torch.uint8)
            else:
                rep_out[nidx] = first; keep_out[nidx] = 1 if total == 1 else 0
    return keep_out, rep_out


def _want_p2p(world, device):
    import os
    return (os.environ.get("DYD_EXCHANGE", "p2p") == "p2p" and world > 1 and dist.is_initialized()
            and torch.device(device).type == "cuda")


class _Region:
    """One direction of the peer-memory exchange: a symmetric receive buffer of `world` regions of `cap`
    (key, id) records (region s is written by rank s), plus this rank's cursors / slot map for what it sent."""

    def __init__(self, cap, world, device, group, want_sent_row):
        import torch.distributed._symmetric_memory as symm
        m = world * cap
        self.cap, self.m = cap, m
        self.recv = symm.empty(2 * m, dtype=torch.int64, device=device)
        self.h = symm.rendezvous(self.recv, group)
        self.peers = torch.tensor(list(self.h.buffer_ptrs), dtype=torch.int64, device=device)
        self.cursors = torch.empty(world, dtype=torch.uint64, device=device)
        self.sent_row = torch.empty(m, dtype=torch.int32, device=device) if want_sent_row else None
        self.recv.fill_(-1)                     # padding once; afterwards the last reader of a record resets it


class DedupExchange:
    """Sync-free sharded dedup (the production multi-GPU form of K4, processor.py:140-144 over all ranks' rows).

    Fixed-capacity buckets make every exchange equal-sized, so nothing has to come back to the host
    between kernels.  Two transports:

    * ``p2p`` (default when the ranks can map each other's memory): the receive and reply buffers live
      in symmetric memory (torch.distributed._symmetric_memory, NVLink peer access).  The bucket kernel
      stores every record straight into its owner's receive buffer and the owner's pack kernel stores
      every answer straight into the origin's reply buffer -- the compute kernels are the all-to-all.
      Two device-side barriers per step order them (records landed, answers landed); the pack kernel, the
      last reader of a received record, turns it back into padding, so no fill and no third barrier.
    * ``nccl``: bucket kernel -> all_to_all_single -> dedup -> reply pack -> reverse all_to_all_single
      -> unpack (DYD_EXCHANGE=nccl forces it).

    The one host read is the overflow flag at the end; if a bucket overflowed (heavily skewed keys)
    the exact-size path `dedup_global` redoes the step.  Buffers are allocated once and reused.
    Null cells (``null`` uint8[n]) never travel: they form one global group resolved by two all-reduces.
    """
    MODE = 0

    def __init__(self, n_local: int, world: int, device, slack: float = 1.10, group=None):
        from . import _lib
        self.lib = _lib.load()
        self.n, self.world, self.dev = n_local, world, device
        self.cap = int(n_local / world * slack) + 4096
        m = world * self.cap
        self.transport = "nccl"
        if _want_p2p(world, device):
            try:
                self._init_p2p(m, group)
                self.transport = "p2p"
            except Exception as e:  # noqa: BLE001 - no peer access / no symmetric memory in this build
                self.p2p_error = repr(e)
        if self.transport == "nccl":
            self.send = torch.empty(2 * m, dtype=torch.int64, device=device)
            self.recv = torch.empty(2 * m, dtype=torch.int64, device=device)
            self.reply = torch.empty(2 * m, dtype=torch.int64, device=device)
            self.back = torch.empty(2 * m, dtype=torch.int64, device=device)
            self.cursors = torch.empty(world, dtype=torch.uint64, device=device)
        self.overflow = torch.zeros(2, dtype=torch.int32, device=device)
        self.keep_r = torch.empty(m, dtype=torch.uint8, device=device)
        self.rep_r = torch.empty(m, dtype=torch.int64, device=device)
        self.ws = torch.empty(self._workspace_bytes(m), dtype=torch.uint8, device=device)
        self.keep = torch.empty(n_local, dtype=torch.uint8, device=device)
        self.rep = torch.empty(n_local, dtype=torch.int64, device=device)

    def _workspace_bytes(self, m):
        return self.lib.dyd_dedup_workspace_bytes(m)

    def _init_p2p(self, m: int, group):
        import torch.distributed._symmetric_memory as symm
        grp = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(grp)
        self.main = _Region(self.cap, self.world, self.dev, grp, True)
        self.back = symm.empty(m, dtype=torch.int64, device=self.dev)          # 8-byte answers, region per owner
        self.h_back = symm.rendezvous(self.back, grp)
        assert self.main.h.world_size == self.world and self.main.h.rank == self.rank
        self.peer_back = torch.tensor(list(self.h_back.buffer_ptrs), dtype=torch.int64, device=self.dev)
        self.back.fill_(-1)
        self.recv, self.cursors = self.main.recv, self.main.cursors
        torch.cuda.synchronize(self.dev)
        self.main.h.barrier(channel=0)                                         # every buffer is padded before anyone writes

    # ---- steps shared with AntiJoinExchange -------------------------------------------------------------
    def _scatter_p2p(self, region, keys, null, row_base, overflow, s):
        from . import _lib
        from .ops import _ptr
        _lib.check(self.lib.dyd_shard_bucket_p2p(_ptr(keys), _ptr(null), row_base, keys.numel(), self.world, self.rank, region.cap,
                                                 _ptr(region.peers), _ptr(region.sent_row), _ptr(region.cursors),
                                                 _ptr(overflow), s), "dyd_shard_bucket_p2p")

    def _reply_p2p(self, s):
        from . import _lib
        from .ops import _ptr
        lib, m = self.lib, self.world * self.cap
        _lib.check(lib.dyd_shard_pack_reply_p2p(_ptr(self.recv), _ptr(self.keep_r), _ptr(self.rep_r), m, self.cap, self.rank,
                                                _ptr(self.peer_back), self.MODE, 1, s), "dyd_shard_pack_reply_p2p")
        self.h_back.barrier(channel=0)                    # all answers have landed (and every rank is done with its records)
        _lib.check(lib.dyd_shard_unpack_p2p(_ptr(self.back), _ptr(self.main.sent_row), _ptr(self.cursors), self.world, self.cap,
                                            self.n, _ptr(self.keep), _ptr(self.rep), self.MODE, s), "dyd_shard_unpack_p2p")

    def _reply_nccl(self, row_base, group, s):
        from . import _lib
        from .ops import _ptr
        lib, m = self.lib, self.world * self.cap
        _lib.check(lib.dyd_shard_pack_reply(_ptr(self.recv), _ptr(self.keep_r), _ptr(self.rep_r), m, _ptr(self.reply), self.MODE, s),
                   "dyd_shard_pack_reply")
        dist.all_to_all_single(self.back, self.reply, group=group)
        _lib.check(lib.dyd_shard_unpack(_ptr(self.back), m, row_base, self.n, _ptr(self.keep), _ptr(self.rep), self.MODE, s),
                   "dyd_shard_unpack")

    def _overflowed(self, group):
        flag = self.overflow.clone()
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        return bool(flag.max().item())

    def run(self, keys: torch.Tensor, row_base: int, keep="first", group=None, check_overflow=True, null=None):
        from . import _lib
        from .ops import KEEP_MODES, _ptr, _stream
        lib, dev, m = self.lib, self.dev, self.world * self.cap
        assert keys.numel() == self.n and keys.dtype == torch.uint64
        with torch.cuda.device(dev):
            s = _stream(dev)
            if self.transport == "p2p":
                self._scatter_p2p(self.main, keys, null, row_base, self.overflow, s)
                self.main.h.barrier(channel=1)                    # all records have landed
            else:
                _lib.check(lib.dyd_shard_bucket(_ptr(keys), _ptr(null), row_base, self.n, self.world, self.cap, _ptr(self.send),
                                                _ptr(self.cursors), _ptr(self.overflow), s), "dyd_shard_bucket")
                dist.all_to_all_single(self.recv, self.send, group=group)
            _lib.check(lib.dyd_dedup_records(_ptr(self.recv), m, KEEP_MODES[keep], _ptr(self.keep_r), _ptr(self.rep_r),
                                             _ptr(self.ws), self.ws.numel(), s), "dyd_dedup_records")
            if self.transport == "p2p":
                self._reply_p2p(s)
            else:
                self._reply_nccl(row_base, group, s)
            if null is not None:
                _resolve_null_group(self.keep, self.rep, null, row_base, keep, group)
        if check_overflow and self._overflowed(group):
            return dedup_global(keys, null, row_base, keep, group)
        return self.keep, self.rep


def _resolve_null_group(keep_out, rep_out, null, row_base, keep, group):
    """Null (NaN) cells are one global group (drop_duplicates treats NaN == NaN): first / last / count across ranks."""
    dev = null.device
    nidx = torch.nonzero(null != 0).squeeze(1)
    big = torch.iinfo(torch.int64).max
    gid = nidx + row_base
    stats = torch.stack([gid.min() if nidx.numel() else torch.tensor(big, device=dev),
                         -gid.max() if nidx.numel() else torch.tensor(big, device=dev)]).to(torch.int64)
    dist.all_reduce(stats, op=dist.ReduceOp.MIN, group=group)
    cnt = torch.tensor([nidx.numel()], dtype=torch.int64, device=dev)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    if nidx.numel():
        first, last = stats[0], -stats[1]
        if keep == "first":
            rep_out[nidx] = first; keep_out[nidx] = (gid == first).to(torch.uint8)
        elif keep == "last":
            rep_out[nidx] = last; keep_out[nidx] = (gid == last).to(torch.uint8)
        else:
            rep_out[nidx] = first; keep_out[nidx] = (cnt[0] == 1).to(torch.uint8)


class AntiJoinExchange(DedupExchange):
    """Sync-free sharded anti-join (the production multi-GPU form of K5, processor.py:194-199 with the main
    table AND the reference set sharded by row across the ranks).

    Both tables are hash-partitioned to owner ranks with the same peer-memory scatter as the dedup: reference
    records (key, global reference row) and main records (key, global row) land in the owner's two receive
    buffers, the owner builds its table from the reference records, probes the main records and stores each
    8-byte answer (kept, or the smallest global reference row holding the key -- what the host needs for the
    string check of a dropped row) straight into the origin's reply buffer.  Two device-side barriers per step;
    the build / pack kernels reset the records they consumed.  ``DYD_EXCHANGE=nccl`` (or ranks without peer
    access) uses bucket -> all_to_all_single for each table and the reverse all_to_all_single for the answers.
    A bucket overflow falls back to `antijoin_global`.  Null main cells are kept (never match); null reference
    cells are dropped (``dropna``)."""
    MODE = 1

    def __init__(self, n_main_local: int, n_ref_local: int, world: int, device, slack: float = 1.10, group=None):
        self.n_ref = n_ref_local
        self.cap_ref = int(n_ref_local / world * slack) + 4096
        super().__init__(n_main_local, world, device, slack, group)
        m_ref = world * self.cap_ref
        if self.transport == "p2p":
            grp = group if group is not None else dist.group.WORLD
            self.ref = _Region(self.cap_ref, world, device, grp, False)
            torch.cuda.synchronize(device)
            self.ref.h.barrier(channel=0)
            self.recv_ref, self.cursors_ref = self.ref.recv, self.ref.cursors
        else:
            self.send_ref = torch.empty(2 * m_ref, dtype=torch.int64, device=device)
            self.recv_ref = torch.empty(2 * m_ref, dtype=torch.int64, device=device)
            self.cursors_ref = torch.empty(world, dtype=torch.uint64, device=device)

    def _workspace_bytes(self, m):
        return self.lib.dyd_antijoin_workspace_bytes(self.world * self.cap_ref)

    def run(self, main_keys, row_base: int, ref_keys, ref_row_base: int, group=None, check_overflow=True,
            main_null=None, ref_null=None):
        from . import _lib
        from .ops import _ptr, _stream
        lib, dev = self.lib, self.dev
        m, m_ref = self.world * self.cap, self.world * self.cap_ref
        assert main_keys.numel() == self.n and ref_keys.numel() == self.n_ref
        with torch.cuda.device(dev):
            s = _stream(dev)
            if self.transport == "p2p":
                self._scatter_p2p(self.ref, ref_keys, ref_null, ref_row_base, self.overflow[1:], s)
                self._scatter_p2p(self.main, main_keys, main_null, row_base, self.overflow[:1], s)
                self.main.h.barrier(channel=1)                    # both tables' records have landed
            else:
                _lib.check(lib.dyd_shard_bucket(_ptr(ref_keys), _ptr(ref_null), ref_row_base, self.n_ref, self.world, self.cap_ref,
                                                _ptr(self.send_ref), _ptr(self.cursors_ref), _ptr(self.overflow[1:]), s), "dyd_shard_bucket")
                dist.all_to_all_single(self.recv_ref, self.send_ref, group=group)
                _lib.check(lib.dyd_shard_bucket(_ptr(main_keys), _ptr(main_null), row_base, self.n, self.world, self.cap,
                                                _ptr(self.send), _ptr(self.cursors), _ptr(self.overflow[:1]), s), "dyd_shard_bucket")
                dist.all_to_all_single(self.recv, self.send, group=group)
            _lib.check(lib.dyd_antijoin_records(_ptr(self.recv_ref), m_ref, _ptr(self.recv), m, _ptr(self.keep_r), _ptr(self.rep_r),
                                                _ptr(self.ws), self.ws.numel(), 1 if self.transport == "p2p" else 0, s),
                       "dyd_antijoin_records")
            if main_null is not None:                         # rows that never travel: a NaN cell never matches
                self.keep.fill_(1); self.rep.fill_(-1)
            if self.transport == "p2p":
                self._reply_p2p(s)
            else:
                self._reply_nccl(row_base, group, s)
        if check_overflow and self._overflowed(group):
            return antijoin_global(main_keys, main_null, ref_keys, ref_null, ref_row_base, group)
        return self.keep, self.rep


def antijoin_global(main_keys, main_null, ref_keys, ref_null, ref_row_base: int, group=None, local_antijoin=None):
    """Global anti-join for this rank's main rows against the union of all ranks' reference rows.

    Both tables are hash-partitioned to owners (one all-to-all each), probed locally, and the
    keep bit + first matching global reference row travel back (reverse all-to-all).
    """
    if local_antijoin is None:
        from . import ops
        local_antijoin = lambda mk, rk: ops.antijoin(mk, None, rk, None)  # noqa: E731
    world = dist.get_world_size(group)
    dev = main_keys.device
    n = main_keys.numel()
    m64 = main_keys.view(torch.int64); r64 = ref_keys.view(torch.int64)
    # reference side: ref.dropna()
    rlive = torch.nonzero(ref_null == 0).squeeze(1) if ref_null is not None else torch.arange(r64.numel(), device=dev)
    rorder, rcounts = _bucket(r64[rlive], world)
    rsel = rlive[rorder]
    rsend = torch.stack([r64[rsel], rsel + ref_row_base], dim=1).contiguous()
    rrecv, _, _ = _exchange(rsend, rcounts.to(dev), group)
    # main side: null cells never match
    mlive = torch.nonzero(main_null == 0).squeeze(1) if main_null is not None else torch.arange(n, device=dev)
    morder, mcounts = _bucket(m64[mlive], world)
    msel = mlive[morder]
    msend = m64[msel].contiguous().unsqueeze(1)
    mrecv, sc, rc = _exchange(msend, mcounts.to(dev), group)
    # owner-local probe; reference rows keep their arrival order, so map the local hit index back
    # to the smallest global reference row holding that key
    rk = rrecv[:, 0].contiguous(); rid = rrecv[:, 1].contiguous()
    if rk.numel():
        order = torch.sort(rid, stable=True).indices          # ascending global row -> first row wins ties
        rk = rk[order].contiguous(); rid = rid[order].contiguous()
    keep_l, hit_l = local_antijoin(mrecv[:, 0].contiguous().view(torch.uint64), rk.view(torch.uint64))
    grow = torch.where(hit_l >= 0, rid[hit_l.clamp(min=0)] if rid.numel() else hit_l, hit_l)
    back = torch.stack([keep_l.to(torch.int64), grow], dim=1).contiguous()
    ans = torch.empty((msend.shape[0], 2), dtype=torch.int64, device=dev)
    dist.all_to_all_single(ans, back, output_split_sizes=sc, input_split_sizes=rc, group=group)
    keep_out = torch.ones(n, dtype=torch.uint8, device=dev)
    ref_out = torch.full((n,), -1, dtype=torch.int64, device=dev)
    keep_out[msel] = ans[:, 0].to(torch.uint8)
    ref_out[msel] = ans[:, 1]
    return keep_out, ref_out


def image_ranges(vertex_prefix, world: int):
    """Contiguous image ranges balanced by vertex count (bytes, not image counts, balance).

    vertex_prefix int64[n_img+1] = poly_off[img_off[i]] (host numpy or CPU tensor).
    Returns a list of (i0, i1) per rank.
    """
    import numpy as np
    vp = np.asarray(vertex_prefix, dtype=np.int64)
    n = len(vp) - 1
    total = int(vp[-1] - vp[0])
    cuts = [0]
    for r in range(1, world):
        target = vp[0] + (total * r) // world
        cuts.append(int(np.searchsorted(vp, target, side="left")))
    cuts.append(n)
    for r in range(1, len(cuts)):
        cuts[r] = max(cuts[r], cuts[r - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def split_category_bases(local_counts, group=None):
    """K6 across ranks: rows are sharded by image in rank order, so category c's expanded
    rows of rank r start at  cat_off[c] + sum over ranks r' < r of counts[r'][c]  in the global,
    category-grouped order the reference builds (processor.py:760-775 appends in row order).

    local_counts int64[n_cat] = this rank's expanded rows per category (dyd_split_count's totals).
    Returns (base int64[n_cat], cat_off int64[n_cat+1]): this rank's first global position per
    category and the global category offsets.  One all_gather of n_cat integers; no data moves."""
    import torch.distributed as dist
    counts = local_counts.to(torch.int64).contiguous()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        all_counts = counts[None]
    else:
        bufs = [torch.empty_like(counts) for _ in range(world)]
        dist.all_gather(bufs, counts, group=group)
        all_counts = torch.stack(bufs)
    totals = all_counts.sum(0)
    cat_off = torch.zeros(len(counts) + 1, dtype=torch.int64, device=counts.device)
    cat_off[1:] = torch.cumsum(totals, 0)
    base = cat_off[:-1] + all_counts[:rank].sum(0)
    return base, cat_off


class UrlFilterExchange(AntiJoinExchange):
    """Steps 2 + 3 in ONE exchange: the main (key, row) records that travel to their owners for the anti-join are the
    same records the dedup needs, so the owner runs both table steps on what it received and sends two 8-byte answers
    per record back.  Compared with `DedupExchange` followed by `AntiJoinExchange` this saves a scatter of the main table
    over NVLink and a barrier pair per step; results are identical (tests/test_gpu_multi.py)."""

    def __init__(self, n_main_local: int, n_ref_local: int, world: int, device, slack: float = 1.10, group=None):
        super().__init__(n_main_local, n_ref_local, world, device, slack, group)
        m = world * self.cap
        self.keep_dr = torch.empty(m, dtype=torch.uint8, device=device)
        self.rep_dr = torch.empty(m, dtype=torch.int64, device=device)
        self.ws_d = torch.empty(self.lib.dyd_url_filter_workspace_bytes(m, world * self.cap_ref), dtype=torch.uint8, device=device)
        self.keep_d = torch.empty(n_main_local, dtype=torch.uint8, device=device)
        self.rep_d = torch.empty(n_main_local, dtype=torch.int64, device=device)
        self._bounds = {}
        if self.transport == "p2p":
            import torch.distributed._symmetric_memory as symm
            grp = group if group is not None else dist.group.WORLD
            self.back2 = symm.empty(2 * m, dtype=torch.int64, device=device)     # (dedup answer, anti-join answer) per slot
            self.h_back2 = symm.rendezvous(self.back2, grp)
            self.peer_back2 = torch.tensor(list(self.h_back2.buffer_ptrs), dtype=torch.int64, device=device)
            self.back2.fill_(-1)
            torch.cuda.synchronize(device)
            self.h_back2.barrier(channel=0)
        else:
            self.reply_d = torch.empty(2 * m, dtype=torch.int64, device=device)
            self.back_d = torch.empty(2 * m, dtype=torch.int64, device=device)

    def run(self, main_keys, row_base: int, ref_keys, ref_row_base: int, keep="first", group=None, check_overflow=True,
            main_null=None, ref_null=None, id_bound=None):
        """Returns (keep, rep, keep_ref, ref_row) for this rank's main rows.  ``id_bound``: an upper bound of every global row id
        of either table over ALL ranks (default: the equal-shard layout, world * rows per rank)."""
        if id_bound is None:                          # agreed on once per (row_base, ref_row_base): one small all-reduce, then cached
            key = (row_base, ref_row_base)
            if key not in self._bounds:
                b = torch.tensor([max(row_base + self.n, ref_row_base + self.n_ref)], dtype=torch.int64, device=self.dev)
                dist.all_reduce(b, op=dist.ReduceOp.MAX, group=group)
                self._bounds[key] = int(b.item())
            id_bound = self._bounds[key]
        from . import _lib
        from .ops import KEEP_MODES, _ptr, _stream
        lib, dev = self.lib, self.dev
        m, m_ref = self.world * self.cap, self.world * self.cap_ref
        assert main_keys.numel() == self.n and ref_keys.numel() == self.n_ref
        p2p = self.transport == "p2p"
        with torch.cuda.device(dev):
            s = _stream(dev)
            sparse = p2p and self.world <= 64
            if p2p:
                self._scatter_p2p(self.ref, ref_keys, ref_null, ref_row_base, self.overflow[1:], s)
                if sparse:        # the scatter also writes "kept by both" for every row; only other answers travel back
                    _lib.check(lib.dyd_shard_bucket_p2p_defaults(_ptr(main_keys), _ptr(main_null), row_base, self.n, self.world, self.rank, self.main.cap,
                                                                 _ptr(self.main.peers), _ptr(self.main.sent_row), _ptr(self.main.cursors),
                                                                 _ptr(self.overflow[:1]), _ptr(self.keep_d), _ptr(self.rep_d), _ptr(self.keep),
                                                                 _ptr(self.rep), s), "dyd_shard_bucket_p2p_defaults")
                else:
                    self._scatter_p2p(self.main, main_keys, main_null, row_base, self.overflow[:1], s)
                self.main.h.barrier(channel=1)                    # both tables' records have landed
            else:
                _lib.check(lib.dyd_shard_bucket(_ptr(ref_keys), _ptr(ref_null), ref_row_base, self.n_ref, self.world, self.cap_ref,
                                                _ptr(self.send_ref), _ptr(self.cursors_ref), _ptr(self.overflow[1:]), s), "dyd_shard_bucket")
                dist.all_to_all_single(self.recv_ref, self.send_ref, group=group)
                _lib.check(lib.dyd_shard_bucket(_ptr(main_keys), _ptr(main_null), row_base, self.n, self.world, self.cap,
                                                _ptr(self.send), _ptr(self.cursors), _ptr(self.overflow[:1]), s), "dyd_shard_bucket")
                dist.all_to_all_single(self.recv, self.send, group=group)
            # owner side: both questions about every received main record from one shared-memory table per key partition
            _lib.check(lib.dyd_url_filter_records(_ptr(self.recv_ref), m_ref, _ptr(self.recv), m, KEEP_MODES[keep], _ptr(self.keep_dr), _ptr(self.rep_dr),
                                                  _ptr(self.keep_r), _ptr(self.rep_r), _ptr(self.ws_d), self.ws_d.numel(), 1 if p2p else 0,
                                                  int(id_bound), s),
                       "dyd_url_filter_records")
            if main_null is not None and not sparse:          # rows that never travel: a NaN cell never matches
                self.keep.fill_(1); self.rep.fill_(-1)
            if p2p:
                # both answers of a record leave in one 16-byte store; the pack kernel, the last reader of the records, resets them
                _lib.check(lib.dyd_shard_pack_reply2_p2p(_ptr(self.recv), _ptr(self.keep_dr), _ptr(self.rep_dr), _ptr(self.keep_r), _ptr(self.rep_r),
                                                         m, self.cap, self.rank, _ptr(self.peer_back2), 1, 1 if sparse else 0, s),
                           "dyd_shard_pack_reply2_p2p")
                self.h_back2.barrier(channel=0)               # all answers have landed
                _lib.check(lib.dyd_shard_unpack2_p2p(_ptr(self.back2), _ptr(self.main.sent_row), _ptr(self.cursors), self.world, self.cap,
                                                     self.n, _ptr(self.keep_d), _ptr(self.rep_d), _ptr(self.keep), _ptr(self.rep),
                                                     1 if sparse else 0, s), "dyd_shard_unpack2_p2p")
            else:
                _lib.check(lib.dyd_shard_pack_reply(_ptr(self.recv), _ptr(self.keep_dr), _ptr(self.rep_dr), m, _ptr(self.reply_d), 0, s),
                           "dyd_shard_pack_reply")
                dist.all_to_all_single(self.back_d, self.reply_d, group=group)
                _lib.check(lib.dyd_shard_unpack(_ptr(self.back_d), m, row_base, self.n, _ptr(self.keep_d), _ptr(self.rep_d), 0, s),
                           "dyd_shard_unpack")
                self._reply_nccl(row_base, group, s)
            if main_null is not None:
                _resolve_null_group(self.keep_d, self.rep_d, main_null, row_base, keep, group)
        if check_overflow and self._overflowed(group):
            k, r = dedup_global(main_keys, main_null, row_base, keep, group)
            k2, r2 = antijoin_global(main_keys, main_null, ref_keys, ref_null, ref_row_base, group)
            return k, r, k2, r2
        return self.keep_d, self.rep_d, self.keep, self.rep


class ShardedUrlFilter:
    """Steps 2 + 3 (dedup by `source`, then the reference filter; processor.py:140-144, 194-199) for a row-sharded
    table whose `source` columns live in HOST memory on every rank -- the N-GPU form of dyd_dedup_host /
    dyd_antijoin_host.  Per call: H2D of the Arrow buffers (pinned memory gives full PCIe speed), K0 hash of both
    columns, `DedupExchange` and `AntiJoinExchange` across the ranks (plain K4 / K5 on one GPU), D2H of the four
    result columns into pinned host arrays.  Device buffers are allocated once for the stated capacities."""

    def __init__(self, n_local: int, n_ref_local: int, max_bytes: int, max_ref_bytes: int, world: int, device, group=None):
        from . import _lib
        self.lib = _lib.load()
        self.n, self.n_ref, self.world, self.dev, self.group = n_local, n_ref_local, world, device, group
        d = device
        self.d_off = torch.empty(n_local + 1, dtype=torch.int64, device=d)
        self.d_data = torch.empty(max_bytes + 8, dtype=torch.uint8, device=d)
        self.d_roff = torch.empty(n_ref_local + 1, dtype=torch.int64, device=d)
        self.d_rdata = torch.empty(max_ref_bytes + 8, dtype=torch.uint8, device=d)
        if world > 1:
            self.xchg = UrlFilterExchange(n_local, n_ref_local, world, d, group=group)
        else:
            self.ws = torch.empty(self.lib.dyd_url_filter_workspace_bytes(n_local, n_ref_local), dtype=torch.uint8, device=d)
        pin = lambda n, dt: torch.empty(n, dtype=dt, pin_memory=True)   # noqa: E731
        self.h_keep, self.h_rep = pin(n_local, torch.uint8), pin(n_local, torch.int64)
        self.h_keep_ref, self.h_ref_row = pin(n_local, torch.uint8), pin(n_local, torch.int64)
        self.h2d_bytes = 0
        self.d2h_bytes = 18 * n_local

    def run(self, h_off, h_data, h_ref_off, h_ref_data, row_base: int, ref_row_base: int, keep="first"):
        from . import ops
        t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(a)   # noqa: E731
        h_off, h_data, h_ref_off, h_ref_data = t(h_off), t(h_data), t(h_ref_off), t(h_ref_data)
        nb, nrb = h_data.numel(), h_ref_data.numel()
        self.h2d_bytes = 8 * (h_off.numel() + h_ref_off.numel()) + nb + nrb
        with torch.cuda.device(self.dev):
            self.d_off.copy_(h_off, non_blocking=True); self.d_data[:nb].copy_(h_data, non_blocking=True)
            self.d_roff.copy_(h_ref_off, non_blocking=True); self.d_rdata[:nrb].copy_(h_ref_data, non_blocking=True)
            keys = ops.hash_strings(self.d_off, self.d_data[:nb])
            rkeys = ops.hash_strings(self.d_roff, self.d_rdata[:nrb])
            if self.world > 1:
                k, r, k2, r2 = self.xchg.run(keys, row_base, rkeys, ref_row_base, keep, group=self.group)
            else:
                k, r, k2, r2 = ops.url_filter(keys, None, rkeys, None, keep, workspace=self.ws)
            self.h_keep.copy_(k, non_blocking=True); self.h_rep.copy_(r, non_blocking=True)
            self.h_keep_ref.copy_(k2, non_blocking=True); self.h_ref_row.copy_(r2, non_blocking=True)
            torch.cuda.current_stream(self.dev).synchronize()
        return self.h_keep, self.h_rep, self.h_keep_ref, self.h_ref_row
