#!/usr/bin/env python
"""Top stall sites from `ncu -i rep --page source --csv`.  Usage: ncu_source_top.py rep.ncu-rep <kernel-regex> [N]"""
import csv, subprocess, sys, re
rep, pat = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}; secs.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = r; continue
    if len(r) == len(cur["hdr"]): cur["data"].append(r)
keys = ['stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_math', 'stall_branch_resolving', 'stall_not_selected', 'stall_mio', 'stall_lg',
        'stall_no_inst', 'stall_dispatch', 'stall_selected', 'stall_barrier']
for s in secs:
    if not re.search(pat, s["name"]): continue
    idx = {h: i for i, h in enumerate(s["hdr"])}
    data = s["data"]
    tot = sum(int(r[idx['# Samples']] or 0) for r in data)
    print("==", s["name"][:80], "samples", tot, "sass instrs", len(data))
    agg = {k: sum(int(r[idx[k]] or 0) for r in data) for k in keys if k in idx}
    print("  stall mix %:", {k.replace('stall_', ''): round(100 * v / max(tot, 1), 1) for k, v in agg.items()})
    ex = sum(int(r[idx['Instructions Executed']] or 0) for r in data)
    print("  warp instrs executed:", ex)
    for r in sorted(data, key=lambda r: -int(r[idx['# Samples']] or 0))[:N]:
        sm = int(r[idx['# Samples']])
        st = sorted(((k, int(r[idx[k]] or 0)) for k in keys if k in idx), key=lambda kv: -kv[1])[:2]
        print(f"  {r[idx['Address']][-5:]} {100 * sm / max(tot, 1):5.1f}%  {r[idx['Source']][:66]:66s} {st}")
