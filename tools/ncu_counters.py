#!/usr/bin/env python3
"""Write profiles/r02_ncu_counters.json from one `ncu --set full` capture of the dominant kernel of the bench workload:
DRAM bytes and warp instructions of ONE launch, stamped with the hash of the kernel sources the capture was taken from.
bench.py reports roofline.traffic / roofline_issue from this file ONLY while that hash still matches the tree.
  python tools/ncu_counters.py gpurun_out/prof.ncu-rep [voices_on_gpu] [frames] [out.json]
Run it ON THE GPU BOX right after the capture (tools/gpu_r2_full.sh does), so the hash is that of the tree the capture ran."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rep = sys.argv[1]
voices = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, r = rows[0], rows[2]


def val(k):
    return float(r[hdr.index(k)].replace(",", ""))


def scaled(k):
    u = rows[1][hdr.index(k)].lower()
    m = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
    return val(k) * m


out = {
    "source": "ncu --set full --clock-control none -k regex:k_render_free -s 6 -c 1 python bench.py --steps 2 --warmup 3 "
              "--no-cpu --no-latency --no-fast --min-timed-s 0: one %d-frame launch of the bench workload" % frames,
    "src_hash": bench.kernel_source_hash(),
    "kernel": r[hdr.index("Kernel Name")],
    "frames": frames, "voices_on_gpu": voices,
    "dram_bytes_read": scaled("dram__bytes_read.sum"),
    "dram_bytes_write": scaled("dram__bytes_write.sum"),
    "warp_instructions": val("smsp__inst_executed.sum"),
    "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "l1tex_hit_rate_pct": val("l1tex__t_sector_hit_rate.pct"),
    "sm_cycles_active_avg": val("smsp__cycles_active.avg"),
    "sm_cycles_elapsed_max": val("sm__cycles_elapsed.max"),
    "duration_us_under_ncu": val("gpu__time_duration.sum") / (1000.0 if rows[1][hdr.index("gpu__time_duration.sum")] in ("nsecond", "ns") else 1.0),
}
p = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", "r02_ncu_counters.json")
json.dump(out, open(p, "w"), indent=1)
print(json.dumps(out))
