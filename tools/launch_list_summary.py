#!/usr/bin/env python3
"""Per-kernel totals and shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
python tools/launch_list_summary.py profiles/r02_final_launches.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
tot = collections.defaultdict(list)
for r in rows:
    ns = float(r[-1].replace(",", ""))
    unit = r[-2]
    us = ns / 1000.0 if unit in ("ns", "nsecond") else (ns if unit in ("us", "usecond") else ns * 1000.0)
    tot[r[4]].append(us)
allus = sum(sum(v) for v in tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -sum(kv[1])):
    v2 = sorted(v)
    print("%-60s n=%4d total %9.1f us  median %8.1f us  share %5.1f%%" % (k[:60], len(v), sum(v), v2[len(v2) // 2], 100.0 * sum(v) / allus))
