#!/usr/bin/env python3
"""Summarise an ncu report (raw + source pages) into text: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [warp_frames]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
wf = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'l1tex__t_sector_hit_rate.pct',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fmaheavy.sum',
        'smsp__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__sass_inst_executed_op_shared.sum']
for r in rows[2:]:
    print("=" * 100)
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print("%-70s %s %s" % (k, r[i], units[i]))
    st = [(hdr[i], r[i]) for i in range(len(hdr)) if 'issue_stalled' in hdr[i] and hdr[i].endswith('per_warp_active.pct')]
    for n, v in sorted(st, key=lambda x: -float(x[1].replace(',', '') or 0))[:8]:
        print("   %-80s %s" % (n.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__warp_issue_stalled_', ''), v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# several kernels may be present: split on "Kernel Name" rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks[:1]:
    h = b["rows"][0]
    ia, isrc, isamp = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
    ops, samp, tot = collections.Counter(), collections.Counter(), 0
    for r in b["rows"][1:]:
        if len(r) <= ia or not r[ia].isdigit():
            continue
        t = r[isrc].split()
        op = t[1] if t[0].startswith('@') else t[0]
        op = op.split('.')[0]
        ops[op] += int(r[ia]); samp[op] += int(r[isamp]); tot += int(r[ia])
    print("-" * 100)
    print(b["name"][:60], "total warp-instructions", tot, ("per warp-frame %.1f" % (tot / wf)) if wf else "")
    for k, v in ops.most_common(28):
        print("  %-10s %12d %s   samples %d" % (k, v, ("%7.2f/frame" % (v / wf)) if wf else "", samp[k]))
    sc = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
    st = collections.Counter()
    for r in b["rows"][1:]:
        for i in sc:
            if i < len(r) and r[i].isdigit():
                st[h[i]] += int(r[i])
    print("  stall samples:", st.most_common(10))
