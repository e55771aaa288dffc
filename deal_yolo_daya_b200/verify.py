"""Ground truth for the sharded dedup / anti-join from the generator's integer url ids.

CHECKER, not product: bench.py and tools/xchg_check.py call it after their timed regions to prove that
the cross-rank exchange produced the global answer.  It shares nothing with the kernels it checks: no
string hashing, no hash tables -- the synthetic `source` of a row is a pure function of its integer url id
(synth.url_of), so equality of ids is equality of strings, and the expected masks come from torch's
sort-based ops over the all-gathered ids.

Semantics restated: processor.py:140-144 (drop_duplicates keep="first": a row survives iff no earlier row
of the concatenated table holds the same source; rep = that first row) and :194-199 (a main row is dropped
iff its source occurs in the reference set; the row reported is the smallest reference row holding it).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def _gather(x: torch.Tensor, group=None) -> torch.Tensor:
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    out = torch.empty(dist.get_world_size(group) * x.numel(), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def expected_dedup_first(url_id: torch.Tensor, row_base: int, group=None):
    """(keep uint8[n], rep int64[n]) of this rank's rows; every rank holds the same number of rows."""
    n = url_id.numel()
    all_ids = _gather(url_id, group)
    _, inverse = torch.unique(all_ids, return_inverse=True)
    first = torch.full((int(inverse.max().item()) + 1,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=url_id.device)
    first.scatter_reduce_(0, inverse, torch.arange(all_ids.numel(), dtype=torch.int64, device=url_id.device), reduce="amin")
    rep = first[inverse[row_base:row_base + n]]
    rows = torch.arange(row_base, row_base + n, dtype=torch.int64, device=url_id.device)
    return (rep == rows).to(torch.uint8), rep


def expected_antijoin(url_id: torch.Tensor, ref_id: torch.Tensor, ref_row_base: int, group=None):
    """(keep uint8[n], ref_row int64[n]); ref rows are global (every rank holds the same number of reference rows)."""
    all_ref = _gather(ref_id, group)
    if all_ref.numel() == 0:
        return torch.ones_like(url_id, dtype=torch.uint8), torch.full_like(url_id, -1)
    sorted_ref, order = torch.sort(all_ref, stable=True)          # equal ids stay in ascending row order
    idx = torch.searchsorted(sorted_ref, url_id).clamp(max=all_ref.numel() - 1)
    hit = sorted_ref[idx] == url_id
    ref_row = torch.where(hit, order[idx], torch.full_like(idx, -1))
    return (~hit).to(torch.uint8), ref_row


def global_counts(keep: torch.Tensor, group=None) -> int:
    """Rows dropped over all ranks."""
    t = (keep.numel() - keep.sum(dtype=torch.int64)).reshape(1)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, group=group)
    return int(t.item())
