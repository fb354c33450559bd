#!/usr/bin/env python3
"""Static schedule of a kernel's loops: python tools/sass_stalls.py [function-substring] [min] [max]
Per loop (backward branch): instructions, sum of the stall counts ptxas encoded (the cycles ONE warp
needs for an iteration if no scoreboard wait binds: B300_MICROARCH.md, single-warp model), waits."""
import re
import subprocess
import sys

import os
so = os.environ.get("SKB_SO", "skred_b200/libskred_b200.so")
fn = sys.argv[1] if len(sys.argv) > 1 else "k_render_free"
mn = int(sys.argv[2]) if len(sys.argv) > 2 else 100
mx = int(sys.argv[3]) if len(sys.argv) > 3 else 700
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
for b in re.split(r"\n\s+Function : ", txt)[1:]:
    name = b.split("\n", 1)[0]
    if fn not in name:
        continue
    lines = b.splitlines()
    ins = []
    i = 0
    while i < len(lines):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", lines[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ins.append((int(m.group(1), 16), m.group(2), (hi >> 41) & 0xf, (hi >> 52) & 0x3f))
                i += 2
                continue
        i += 1
    print(name[:60], len(ins), "instructions")
    addr2i = {a: k for k, (a, *_) in enumerate(ins)}
    for k, (a, t, *_) in enumerate(ins):
        m = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr2i:
            s = addr2i[tgt]
            n = k - s + 1
            if mn <= n <= mx:
                tot = sum(x[2] for x in ins[s:k + 1])
                ops = {}
                for x in ins[s:k + 1]:
                    w = x[1].split()
                    op = (w[1] if w[0].startswith("@") else w[0]).split(".")[0]
                    ops[op] = ops.get(op, 0) + 1
                top = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
                print("  loop @%5d n=%4d stall-sum=%5d (%.2f cyc/instr) waits=%3d  %s" %
                      (s, n, tot, tot / n, sum(1 for x in ins[s:k + 1] if x[3]), top))
