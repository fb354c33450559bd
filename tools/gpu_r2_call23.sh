#!/bin/bash
# 0.sk inside long launches: where the time of a levelled launch goes (launch list of tools/osk_probe.py), modulated-voice probe
mkdir -p gpurun_out
timeout 300 python tools/bins_bench.py 1024 4096 2>&1 | grep -v "^#" > gpurun_out/bins_bench.txt; cat gpurun_out/bins_bench.txt
timeout 200 python tools/osk_probe.py 8192 8 2>&1 | grep -v "^#"
timeout 200 python tools/osk_probe.py 512 40 2>&1 | grep -v "^#" | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/osk_launches.csv python tools/osk_probe.py 8192 6 > gpurun_out/osk_ncu.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/osk_launches.csv")) if len(r) > 10 and r[0].isdigit()]
for r in rows[-16:]:
    print(r[4][:60], r[-1], r[-2])
PY
