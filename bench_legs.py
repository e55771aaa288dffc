"""The C4 and C5 legs of bench.py (BASELINE.json configs[3] and configs[4]); imported by bench.py only.

c4_leg  dense-crowd stress: 1 M images x 200-500 boxes (boxes given directly as two-point ptLists), K2 block-per-image
        kernel at thr 0.7 (the generator's natural mix, early exit allowed) and at a threshold no pair can reach (worst
        case: no early exit, every candidate pair evaluated); images/s and the brute-force-equivalent pairs/s; a sample of the
        full-size result is compared with the C oracle.  Single-GPU configuration: reported on the N=1 line only.
c5_leg  label remap 80 -> 20 (K3) + rule-based split (K6 expand, global category offsets over NCCL, host numpy permutation,
        K6 assign) on the C2 table of every rank; results verified against the single-table order on a slice and by
        size-independent properties (processor.py:582-602, 751-806).
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.distributed as dist

FP64_PEAK_TFLOPS = 37.0          # B200 non-tensor fp64; the guide's figure, not measured here
FLOPS_PER_PAIR = 14              # calculate_iou: 4 min/max selects, 7 sub/mul/add, 1 div, 2 compares (processor.py:328-339)


def _time(fn, reps):
    fn(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]


def c4_leg(dev, rank, world, peak, n_img=1_000_000):
    if world > 1:
        return {"skipped": "single-GPU configuration: reported on the N=1 line"}
    from deal_yolo_daya_b200 import ops, synth_device
    from oracle import oracle_c
    io, pts = synth_device.make_crowd(0, 0, n_img, device=dev)
    nb = pts.numel() // 4
    cnt = (io[1:] - io[:-1]).double()
    pairs = float((cnt * (cnt - 1) / 2).sum().item())
    ws = torch.empty(ops._lib.load().dyd_iou_workspace_bytes(n_img), dtype=torch.uint8, device=dev)
    out = {}
    res = {}
    for thr, tag in ((0.7, "natural_mix"), (2.0, "worst_case")):
        ms = _time(lambda: res.__setitem__(tag, ops.iou_filter(io, pts, None, 2, thr, workspace=ws)), 3)
        out[tag] = {"thr": thr, "ms": ms, "images_per_s": n_img / (ms * 1e-3), "high_images": int(res[tag][0].sum().item())}
        if tag == "worst_case":
            pps = pairs / (ms * 1e-3)
            out[tag].update({"equivalent_pairs_per_s": pps,
                             "equivalent_note": "n(n-1)/2 pairs of every image / time: the binned kernel decides all of them but evaluates only the pairs "
                                                "that can overlap in x (DESIGN.md 4.3), so this is not an fp64 rate; the all-pairs kernel of this round, "
                                                "which evaluates every pair, ran at 1.26e12 pairs/s = 0.48 of the 37 TFLOP/s non-tensor fp64 figure "
                                                "(profiles/r2_ncu_crowd_1M.json)"})
    out["natural_mix"]["hbm_gbs"] = (32 * nb + 13 * n_img) / (out["natural_mix"]["ms"] * 1e-3) / 1e9
    out["natural_mix"]["frac_of_hbm_peak"] = out["natural_mix"]["hbm_gbs"] / peak
    # sampled parity at the full size: every 2500th image against the C oracle (both thresholds)
    pick = torch.arange(0, n_img, 2500, device=dev)
    a, b = io[pick].cpu().numpy(), io[pick + 1].cpu().numpy()
    sub_off = np.zeros(len(a) + 1, np.int64); np.cumsum(b - a, out=sub_off[1:])
    sub_pts = np.concatenate([pts[4 * x:4 * y].cpu().numpy() for x, y in zip(a, b)])
    ok = True
    for thr, tag in ((0.7, "natural_mix"), (2.0, "worst_case")):
        wh, wc = oracle_c.iou_filter(sub_off, sub_pts, np.ones(len(sub_pts) // 4, np.uint8), 2, thr)
        ok = ok and np.array_equal(res[tag][0][pick].cpu().numpy(), wh) and np.array_equal(res[tag][1][pick].cpu().numpy(), wc)
    out.update({"images": n_img, "boxes": nb, "pairs": pairs, "flops_per_pair": FLOPS_PER_PAIR, "fp64_peak_tflops": FP64_PEAK_TFLOPS,
                "parity_sample": {"images": int(len(a)), "equal_to_oracle": bool(ok)},
                "bound": "per-image latency (counting sort, barriers, short walks) in the binned form; the all-pairs form was fp64 / shared-memory "
                         "bound (77 flop/B, SURVEY.md 8d).  Neither is HBM-bound: the low HBM fraction is expected"})
    assert ok, "C4 sample differs from the oracle"
    return out


def c5_leg(dev, rank, world, t, peak, seed=42, ratios=(0.8, 0.1, 0.1)):
    from deal_yolo_daya_b200 import native, ops, sharding, synth
    from oracle import oracle_c
    n_img, n_obj = t.n_img, t.n_poly
    n_lab, n_grp, n_cat = synth.N_LABELS, 20, 4
    # vocabulary: ids 0..79 = cls00..cls79 (the table's names), 80..99 = grp00..grp19 (the mapping's targets)
    lut_new = torch.tensor([n_lab + i % n_grp for i in range(n_lab)] + list(range(n_lab, n_lab + n_grp)), dtype=torch.int32, device=dev)
    lut_ntok = torch.ones(n_lab + n_grp, dtype=torch.int32, device=dev)
    lut_nrep = torch.tensor([1] * n_lab + [0] * n_grp, dtype=torch.int32, device=dev)
    cat_of = torch.tensor([-1] * n_lab + [g // 5 for g in range(n_grp)], dtype=torch.int32, device=dev)
    st = {}

    def k3():
        st["new"], st["rowrep"], st["cnt"] = ops.label_lut(t.img_off, t.label_id, lut_new, lut_ntok, lut_nrep)

    def k6():
        st["ei"], st["eb"], st["ec"], st["co"] = ops.split_expand(t.img_off, st["new"], cat_of, n_cat)

    ms3 = _time(k3, 5)
    ms6 = _time(k6, 3)
    local_counts = (st["co"][1:] - st["co"][:-1]).contiguous()
    torch.cuda.synchronize()
    a = time.perf_counter()
    base, cat_off_g = sharding.split_category_bases(local_counts)
    torch.cuda.synchronize()
    ms_bases = (time.perf_counter() - a) * 1e3
    n_c = (cat_off_g[1:] - cat_off_g[:-1]).cpu().numpy()
    n_exp_g = int(cat_off_g[-1].item())
    # host permutations (numpy's legacy MT19937 shuffle, what DataFrame.sample(frac=1, random_state=seed) applies; native.permutation): category c is shuffled by
    # rank c % world and broadcast, so the wall time is one category's, not the sum
    a = time.perf_counter()
    perm = torch.empty(n_exp_g, dtype=torch.int64, device=dev)
    for c in range(n_cat):
        seg = perm[int(cat_off_g[c].item()):int(cat_off_g[c + 1].item())]
        if c % world == rank:
            seg.copy_(torch.from_numpy(native.permutation(seed, int(n_c[c]))))
    if world > 1:
        for c in range(n_cat):
            dist.broadcast(perm[int(cat_off_g[c].item()):int(cat_off_g[c + 1].item())], src=c % world)
    torch.cuda.synchronize()
    s_perm = time.perf_counter() - a
    n_train = torch.tensor([int(x * ratios[0]) for x in n_c], dtype=torch.int64, device=dev)
    n_val = torch.tensor([int(x * ratios[1]) for x in n_c], dtype=torch.int64, device=dev)
    # own rows of category c are rows base[c] - cat_off_g[c] .. of that category; every rank sweeps the global permutation once
    # and keeps the answers of its own rows (dyd_split_assign_range)
    own_lo = (base - cat_off_g[:-1]).contiguous()
    ms_assign = _time(lambda: st.__setitem__("sp", ops.split_assign_range(cat_off_g, perm, n_train, n_val, own_lo, local_counts, st["co"][:-1].contiguous())), 3)
    split_own, pos_own = st["sp"]
    own_cat = st["ec"].long()
    # ---- verification
    checks = {}
    hist = torch.zeros(n_cat * 3, dtype=torch.int64, device=dev)
    hist.scatter_add_(0, own_cat * 3 + split_own.long(), torch.ones_like(own_cat))
    possum = torch.zeros(n_cat, dtype=torch.float64, device=dev)
    possum.scatter_add_(0, own_cat, pos_own.double())          # pos = shuffled position inside the category
    tot = local_counts.clone()
    if world > 1:
        dist.all_reduce(hist); dist.all_reduce(possum); dist.all_reduce(tot)
    hist = hist.view(n_cat, 3).cpu().numpy()
    checks["split_sizes"] = bool(all(hist[c, 0] == int(n_c[c] * ratios[0]) and hist[c, 1] == int(n_c[c] * ratios[1]) and hist[c].sum() == n_c[c] for c in range(n_cat)))
    checks["positions_are_a_permutation"] = bool(all(abs(possum[c].item() - n_c[c] * (n_c[c] - 1) / 2) < 0.5 for c in range(n_cat)))
    checks["counts_add_up"] = bool(np.array_equal(tot.cpu().numpy(), n_c))
    if rank == 0:                               # the single-table order on a slice: rank 0's first images come first in every category
        ns = 4000
        h = synth.make_table(t.seed, t.first_img, ns)
        w_new, w_rr, w_cnt = oracle_c.label_lut(h.img_off, h.label_id, lut_new.cpu().numpy(), lut_ntok.cpu().numpy(), lut_nrep.cpu().numpy())
        wi, wb, wc, wo = oracle_c.split_expand(h.img_off, w_new, cat_of.cpu().numpy(), n_cat)
        ok = np.array_equal(st["new"][:h.n_poly].cpu().numpy(), w_new) and np.array_equal(st["rowrep"][:ns].cpu().numpy(), w_rr)
        for c in range(n_cat):
            k = int(wo[c + 1] - wo[c]); a0 = int(st["co"][c].item())
            ok = ok and np.array_equal(st["ei"][a0:a0 + k].cpu().numpy(), wi[wo[c]:wo[c + 1]]) and np.array_equal(st["eb"][a0:a0 + k].cpu().numpy(), wb[wo[c]:wo[c + 1]])
            ok = ok and int(base[c].item()) == int(cat_off_g[c].item())          # rank 0 starts every category
        checks["slice_equals_single_table_order"] = bool(ok)
    cnt = st["cnt"].cpu().numpy()
    out = {"images_per_gpu": n_img, "objects_per_gpu": n_obj, "expanded_rows_global": n_exp_g, "categories": n_cat,
           "k3_remap_ms": ms3, "k3_frac_of_hbm_peak": (8 * n_obj + 9 * n_img) / (ms3 * 1e-3) / 1e9 / peak,
           "k6_expand_ms": ms6, "k6_frac_of_hbm_peak": (8 * n_obj + 8 * n_img + 20 * int(st["co"][-1].item())) / (ms6 * 1e-3) / 1e9 / peak,
           "category_bases_allgather_ms": ms_bases, "host_permutation_s": s_perm, "k6_assign_ms": ms_assign,
           "images_per_s_kernels": world * n_img / ((ms3 + ms6 + ms_assign) * 1e-3),
           "replaced_labels_rank0": int(cnt[3]), "checks": checks, "verified": bool(all(checks.values())),
           "note": "K3 label LUT 80->20 + K6 expand / global category offsets (NCCL all_gather) / assign on every rank's C2 table; the host permutation "
                   "(np.random.RandomState(42).permutation per category, reference semantics, produced bit for bit by dyd_numpy_permutation) is timed separately and is not in images_per_s_kernels"}
    ok_t = torch.tensor([int(all(checks.values()))], device=dev)
    if world > 1:
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
    assert bool(ok_t.item()), f"C5 verification failed: {checks}"
    return out
