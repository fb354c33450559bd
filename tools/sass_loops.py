#!/usr/bin/env python3
"""Static look at the hot loops of a kernel: python tools/sass_loops.py [function-substring] [min-instrs]
Dumps every loop (backward branch) of the kernel with its instruction count and opcode mix."""
import collections
import re
import subprocess
import sys

so = "skred_b200/libskred_b200.so"
fn = sys.argv[1] if len(sys.argv) > 1 else "k_render_free"
mn = int(sys.argv[2]) if len(sys.argv) > 2 else 100
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s+Function : ", txt)
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if fn not in name:
        continue
    ins = []
    for l in b.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    print(name[:70], "total instrs", len(ins))
    addr2i = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr2i:
            s = addr2i[tgt]
            n = i - s + 1
            if n < mn:
                continue
            ops = collections.Counter()
            for _, tt in ins[s:i + 1]:
                w = tt.split()
                op = w[1] if w[0].startswith("@") else w[0]
                ops[op.split(".")[0]] += 1
            print("  loop @%d..%d: %d instrs  %s" % (s, i, n, dict(ops.most_common(16))))
