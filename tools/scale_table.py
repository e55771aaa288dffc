#!/usr/bin/env python
"""Markdown scaling table from profiles/r2_bench_n{1,2,4,8}.json (developer tool)."""
import json, sys
from pathlib import Path
P = Path(__file__).resolve().parent.parent / "profiles"
rows, base = [], None
for n in (1, 2, 4, 8):
    f = P / (f"r2_bench_n{n}.json" if n > 1 else "r2_bench_n1_steps20.json")
    if not f.exists():
        continue
    d = json.loads(f.read_text())
    s, e = d["streams"], d.get("e2e") or {}
    if n == 1:
        base = d["value"]; ebase = e.get("value")
    x = s.get("joint_exchange_ms", s.get("url_filter_ms"))
    rows.append(f"| {n} | {d['steps']}+{d['warmup']} | {d['value'] / 1e9:.2f} | {d['ms_per_step']:.2f} | {s['fused_ms']:.2f} | {d['roofline']['frac']:.3f} | {s['url_chain_ms']:.2f} ({x:.2f}) | "
                f"{d['value'] / (n * base):.3f} | {d['results']['exchange_verified']} | {e.get('value', 0) / 1e6:.1f} | {e.get('value', 0) / (n * ebase) if ebase else 0:.2f} |")
print("| N | steps+warmup | G images/s | ms/step | fused ms | fused frac of peak | URL chain ms (of which dedup + anti-join / exchange) | efficiency vs N=1 | exchange verified | e2e M images/s | e2e per-GPU vs N=1 |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
print("\n".join(rows))
