#!/usr/bin/env python
"""Top stall sites of each kernel in an .ncu-rep captured with --set full --import-source on.
Usage: python tools/ncu_stalls.py rep.ncu-rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kern = None; hdr = None; rows = []
def flush():
    if not rows: return
    tot = sum(r[1] for r in rows)
    print(f"== {kern[:100]}  total samples {tot}")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = {hdr[i]: 0 for i in stall_cols}
    for r in rows:
        for i in stall_cols:
            agg[hdr[i]] += int(r[2][i] or 0)
    print("   by reason:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
    for r in sorted(rows, key=lambda r: -r[1])[:topn]:
        why = sorted(((int(r[2][i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(f"   {r[1]:7d} {100.0*r[1]/max(tot,1):5.1f}%  #{r[0]:5d} {r[2][1].strip()[:80]:80s} {why}")
for line in csv.reader(io.StringIO(out)):
    if not line: continue
    if line[0] == "Kernel Name":
        flush(); kern = line[1]; hdr = None; rows = []; continue
    if line[0] == "Address":
        hdr = line; continue
    if hdr is None: continue
    try: s = int(line[2])
    except ValueError: continue
    rows.append((len(rows), s, line))
flush()
