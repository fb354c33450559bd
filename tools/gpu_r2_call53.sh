#!/bin/bash
# round 2, session 3, call 14: what the driver runs at round end, with its defaults: smoke, bench.py (no flags), bench.py --impl reference
mkdir -p gpurun_out
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | grep -v "^#" | tail -5
( time timeout 900 python bench.py --impl reference 2>gpurun_out/bench_ref_default.err > gpurun_out/bench_ref_default.json ) 2>&1 | tail -3; tail -c 300 gpurun_out/bench_ref_default.json
( time timeout 900 python bench.py 2>gpurun_out/bench_default.err > gpurun_out/bench_default.json ) 2>&1 | tail -3
python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print({k: d[k] for k in ('value','steps','warmup','ms_per_step','block_latency_ms_p50','block_latency_ms_p50_64_voices','gpu_launches')}); print(d['roofline']['frac'], d['roofline']['traffic'], (d['roofline_issue'] or {}).get('frac')); print(d['e2e']['value'], d['clocks']); print(d['cpu_baseline'])"
