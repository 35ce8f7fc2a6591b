#!/usr/bin/env python
"""Instruction mix of the loops of one kernel: sass_loops.py lib.so <substring of the mangled name> [min loop size]."""
import re, subprocess, sys
from collections import Counter
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
funcs, cur = {}, None
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = funcs.setdefault(m.group(1), [])
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;", l)
    if m and cur is not None:
        cur.append((int(m.group(1), 16), m.group(2)))
minsz = int(sys.argv[3]) if len(sys.argv) > 3 else 40
for name, body in funcs.items():
    if sys.argv[2] not in name:
        continue
    print("==", name, len(body), "instructions")
    addr = {a: i for i, (a, _) in enumerate(body)}
    loops = []
    for i, (a, ins) in enumerate(body):
        m = re.search(r"BRA\S*\s+(?:!?U?P\d+,\s*)?(?:P\d+,\s*)?(0x[0-9a-f]+)", ins)
        if m and int(m.group(1), 16) in addr and addr[int(m.group(1), 16)] < i:
            loops.append((addr[int(m.group(1), 16)], i))
    for lo, hi in sorted(loops, key=lambda t: t[0] - t[1]):
        if hi - lo + 1 < minsz:
            continue
        ops = [re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0] for _, ins in body[lo:hi + 1]]
        fma = sum(o.startswith(("IMAD", "FFMA", "FMUL", "FADD", "HFMA")) for o in ops)
        lsu = sum(o.startswith(("LDS", "STS", "LDG", "STG", "LD.", "ST.", "ATOM", "RED")) for o in ops)
        xu = sum(o.startswith(("MUFU", "I2F.", "F2I", "POPC", "FLO")) for o in ops)
        ctl = sum(o.startswith(("BRA", "BSSY", "BSYNC", "WARPSYNC", "BAR", "SYNCS", "NOP", "EXIT", "CALL", "RET")) for o in ops)
        print(f"  loop {body[lo][0]:#x}..{body[hi][0]:#x}: {len(ops)} instr  fma {fma}  lsu {lsu}  xu {xu}  ctl {ctl}  alu/other {len(ops) - fma - lsu - xu - ctl}")
        print("     ", Counter(ops).most_common(16))
