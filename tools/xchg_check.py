"""Peer-memory vs NCCL transport of the sharded dedup and the sharded anti-join: same answers as each
other, as the exact-size paths and as an independent torch (sort-based) ground truth on the generator's
url ids; and time per step.  Run under torchrun on >= 2 GPUs (tests/test_gpu_multi.py does)."""
import os, sys, torch, torch.distributed as dist
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from deal_yolo_daya_b200 import ops, sharding, synth_device
from deal_yolo_daya_b200.verify import expected_antijoin, expected_dedup_first

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
n_ref = n // 2
url_id, uoff, udata = synth_device.make_urls(0, rank * n, n, dev)
keys = ops.hash_strings(uoff, udata)
ref_id, roff, rdata = synth_device.make_urls(0, rank * n_ref, n_ref, dev, n_main_for_ref=world * n)
rkeys = ops.hash_strings(roff, rdata)
null = (torch.arange(n, device=dev) % 997 == 5).to(torch.uint8)          # a few NaN cells on every rank
rnull = (torch.arange(n_ref, device=dev) % 1013 == 7).to(torch.uint8)
res, ares, jres = {}, {}, {}


def timed(fn, label):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / 10], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"{label}: {t.item():.3f} ms per exchange step ({n} rows per rank)", flush=True)


for transport in ("nccl", "p2p"):
    os.environ["DYD_EXCHANGE"] = transport
    x = sharding.DedupExchange(n, world, dev)
    if rank == 0:
        print(f"requested {transport}: using {x.transport}" + (f" ({getattr(x, 'p2p_error', '')})" if x.transport != transport else ""), flush=True)
    for rep in range(2):                                   # twice: the second pass runs on buffers the first one reset
        for keep in ("first", "last", False):
            k, r = x.run(keys, rank * n, keep)
            res[(transport, keep, rep)] = (k.clone(), r.clone())
    k, r = x.run(keys, rank * n, "first", null=null)
    res[(transport, "nulls")] = (k.clone(), r.clone())
    timed(lambda: x.run(keys, rank * n, "first", check_overflow=False), f"dedup {x.transport}")
    del x
    y = sharding.AntiJoinExchange(n, n_ref, world, dev)
    for rep in range(2):
        k, r = y.run(keys, rank * n, rkeys, rank * n_ref)
        ares[(transport, rep)] = (k.clone(), r.clone())
    k, r = y.run(keys, rank * n, rkeys, rank * n_ref, main_null=null, ref_null=rnull)
    ares[(transport, "nulls")] = (k.clone(), r.clone())
    timed(lambda: y.run(keys, rank * n, rkeys, rank * n_ref, check_overflow=False), f"antijoin {y.transport}")
    del y
    z = sharding.UrlFilterExchange(n, n_ref, world, dev)
    for rep in range(2):
        for keep in ("first", "last", False):
            jres[(transport, keep, rep)] = [x.clone() for x in z.run(keys, rank * n, rkeys, rank * n_ref, keep)]
    jres[(transport, "nulls")] = [x.clone() for x in z.run(keys, rank * n, rkeys, rank * n_ref, "first", main_null=null, ref_null=rnull)]
    timed(lambda: z.run(keys, rank * n, rkeys, rank * n_ref, "first", check_overflow=False), f"joint {z.transport}")
    del z


def eq(a, b):
    return torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


same = all(eq(res[("nccl", k, rep)], res[("p2p", k, rep2)]) for k in ("first", "last", False) for rep in (0, 1) for rep2 in (0, 1))
same = same and eq(res[("nccl", "nulls")], res[("p2p", "nulls")])
kk, rr = sharding.dedup_global(keys, None, rank * n, "first")
same_exact = eq((kk, rr), res[("p2p", "first", 1)])
kk, rr = sharding.dedup_global(keys, null, rank * n, "first")
same_exact = same_exact and eq((kk, rr), res[("p2p", "nulls")])
ek, er = expected_dedup_first(url_id, rank * n)
truth = eq((ek, er), res[("p2p", "first", 1)])
a_same = all(eq(ares[("nccl", r1)], ares[("p2p", r2)]) for r1 in (0, 1) for r2 in (0, 1)) and eq(ares[("nccl", "nulls")], ares[("p2p", "nulls")])
kk, rr = sharding.antijoin_global(keys, None, rkeys, None, rank * n_ref)
a_exact = eq((kk, rr), ares[("p2p", 1)])
kk, rr = sharding.antijoin_global(keys, null, rkeys, rnull, rank * n_ref)
a_exact = a_exact and eq((kk, rr), ares[("p2p", "nulls")])
ek, er = expected_antijoin(url_id, ref_id, rank * n_ref)
a_truth = eq((ek, er), ares[("p2p", 1)])
j_ok = True
for transport in ("nccl", "p2p"):
    for rep in (0, 1):
        for keep in ("first", "last", False):
            j = jres[(transport, keep, rep)]
            j_ok = j_ok and eq(j[:2], res[("p2p", keep, 1)]) and eq(j[2:], ares[("p2p", 1)])
    j = jres[(transport, "nulls")]
    j_ok = j_ok and eq(j[:2], res[("p2p", "nulls")]) and eq(j[2:], ares[("p2p", "nulls")])
flags = torch.tensor([int(same), int(same_exact), int(truth), int(a_same), int(a_exact), int(a_truth), int(j_ok)], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    f = [bool(v) for v in flags.tolist()]
    print("p2p == nccl:", f[0], " p2p == exact-size path:", f[1], " dedup == url-id ground truth:", f[2], flush=True)
    print("antijoin p2p == nccl:", f[3], " antijoin p2p == exact-size path:", f[4], " antijoin == url-id ground truth:", f[5], flush=True)
    print("joint exchange == separate exchanges (both transports):", f[6], flush=True)
dist.destroy_process_group()
