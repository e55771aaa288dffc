#!/usr/bin/env python
"""Per-source-line share of executed instructions / stall samples / shared-memory wavefronts from an
.ncu-rep captured with --import-source on (developer tool).  usage: ncu_lines.py report.ncu-rep [min_pct]"""
import csv, io, subprocess, sys

def main():
    rep = sys.argv[1]
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    cur, hdr, out = None, None, []
    def I(x):
        try: return int(x)
        except ValueError: return 0
    for r in csv.reader(io.StringIO(txt)):
        if not r: continue
        if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
        if r[0] == "Function Name": continue
        if r[0] == "Line No": hdr = r; continue
        if hdr and r[0].isdigit():
            g = lambda k: I(r[hdr.index(k)])
            out.append((cur, int(r[0]), r[1].strip()[:90], g("Instructions Executed"), g("# Samples"),
                        g("L1 Wavefronts Shared"), g("L1 Wavefronts Shared Ideal")))
    ti, ts, tw = (sum(o[k] for o in out) or 1 for k in (3, 4, 5))
    print(f"total: {ti} warp instructions, {ts} samples, {tw} shared wavefronts")
    for o in sorted(out, key=lambda o: (o[0], o[1])):
        if 100 * o[3] / ti < thr and 100 * o[4] / ts < thr: continue
        print(f"{o[0][:14]:14s} {o[1]:4d} inst {100*o[3]/ti:5.1f}% smp {100*o[4]/ts:5.1f}% wf {100*o[5]/tw:5.1f}% (ideal {100*o[6]/tw:4.1f}) | {o[2]}")

main()
